"""ctypes wrapper of ``oracle/cheby_ref.c`` (the C restatement of the reference's CPU Chebyshev
step).  Test / measurement infrastructure only: used by ``bench.py``'s CPU legs and
cross-checked against the NumPy oracle in ``tests/test_oracle_pins.py``."""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libcheby_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB)
        _lib.cheby_ref_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    """Host threads the OpenMP variant may use: the cores this process is allowed to run on.
    (Not ``omp_get_max_threads()``: torchrun exports ``OMP_NUM_THREADS=1`` to every rank, which
    would silently turn the multi-threaded CPU baseline into a single-threaded one at N > 1.)"""
    import os

    _load()
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(int(n), 256))


class ChebyRef:
    """Holds the operators in the reference's own storage (CSC, Int64 indices, like
    ``SparseMatrixCSC{ComplexF64,Int64}``) and as CSR for the OpenMP variant."""

    def __init__(self, ops, n_coeffs: int):
        self.n = ops[0].shape[0]
        self.n_ops = len(ops)
        self.n_coeffs = n_coeffs
        self._keep = []
        self.csc = self._pack([sp.csc_matrix(A) for A in ops])
        self.csr = self._pack([sp.csr_matrix(A) for A in ops])
        self.v = [np.zeros(self.n, dtype=np.complex128) for _ in range(3)]

    def _pack(self, mats):
        ptrs, idxs, vals = [], [], []
        for A in mats:
            p = np.ascontiguousarray(A.indptr, dtype=np.int64)
            i = np.ascontiguousarray(A.indices, dtype=np.int64)
            v = np.ascontiguousarray(A.data, dtype=np.complex128)
            self._keep += [p, i, v]
            ptrs.append(p.ctypes.data)
            idxs.append(i.ctypes.data)
            vals.append(v.ctypes.data)
        arr = lambda xs: (C.c_void_p * len(xs))(*xs)  # noqa: E731
        return arr(ptrs), arr(idxs), arr(vals)

    def step(self, psi, coeffs, a, Delta, E_min, dt, threads=0):
        """In-place ``cheby!`` on ``psi``; ``threads == 0``: the faithful single-thread CSC
        scatter form; ``threads > 0``: the OpenMP row-parallel CSR variant."""
        lib = _load()
        coeffs = np.ascontiguousarray(coeffs, dtype=np.complex128)
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert psi.dtype == np.complex128 and psi.flags.c_contiguous and coeffs.size == self.n_coeffs
        common = (
            C.c_void_p(coeffs.ctypes.data), C.c_int(self.n_coeffs), C.c_void_p(psi.ctypes.data),
            C.c_void_p(self.v[0].ctypes.data), C.c_void_p(self.v[1].ctypes.data), C.c_void_p(self.v[2].ctypes.data),
            C.c_void_p(a.ctypes.data), C.c_int(len(a)), C.c_double(Delta), C.c_double(E_min), C.c_double(dt),
        )
        if threads == 0:
            return lib.cheby_step_csc(C.c_int64(self.n), C.c_int(self.n_ops), *self.csc, *common)
        return lib.cheby_step_csr_omp(C.c_int64(self.n), C.c_int(self.n_ops), *self.csr, *common, C.c_int(threads))
