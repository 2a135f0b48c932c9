"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every
symbol include/qprop.h declares (no compute calls without a GPU), compute calls fail loudly
without a device, and the host-side mirror (controls, generators, coefficient / Leja tables)
agrees with the oracle."""

import ctypes
import os
import re

import numpy as np
import pytest

import oracle as O
import qprop_b200 as qp
from qprop_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "qprop.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    names = _header_functions()
    assert len(names) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"libqprop_b200.so does not export {name}"
    # the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().qp_version() == 100


def test_header_cites_reference_and_has_plain_c_signatures():
    text = open(os.path.join(ROOT, "include", "qprop.h")).read()
    for cite in ("src/cheby.jl:150-213", "src/arnoldi.jl:60-100", "src/generators.jl:634-645", "src/newton.jl"):
        assert cite in text
    assert "torch" not in text.lower() and 'extern "C"' in text


def test_status_strings():
    lib = _lib.load()
    assert lib.qp_status_string(0) == b"ok"
    assert lib.qp_status_string(-5) == b"incorrect normalization"
    assert lib.qp_status_string(-4) == b"not converged"


def _no_gpu():
    import torch

    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device."""
    with pytest.raises(qp.QPropError) as exc:
        qp.Context(0)
    assert exc.value.status == -2 and "no CPU fallback" in str(exc.value)
    with pytest.raises(qp.QPropError):
        qp.propagate(np.array([1, 0], dtype=complex), (np.eye(2, dtype=complex),), np.linspace(0, 1, 5), "cheby")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "quantumpropagators.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f


# ---------------------------------------------------------------------------------------
# host mirror vs oracle
# ---------------------------------------------------------------------------------------


def test_cheby_coeffs_host():
    for Delta, dt in [(20.2, 0.02), (2.0, 1.0), (12.0, 1.0), (60.0, 1.0), (99.99, 0.1), (460.0, 1.0)]:
        a, b = qp.cheby_coeffs(Delta, dt), O.cheby_coeffs(Delta, dt)
        assert np.array_equal(a, b)
        assert abs(a[-1]) <= 1e-12 < abs(a[-2])  # the first coefficient <= limit is kept
    assert len(qp.cheby_coeffs(20.2, 0.02)) == 9 and len(qp.cheby_coeffs(2.0, 1.0)) == 13
    n, buf = qp.cheby_coeffs_(np.zeros(4), 60.0, 1.0)
    assert np.array_equal(buf[:n], O.cheby_coeffs(60.0, 1.0)) and len(buf) >= n


def test_discretization_host():
    assert np.array_equal(qp.get_tlist_midpoints([1, 3, 5, 6, 7]), [1, 4, 5.5, 7])
    tlist = np.linspace(0, 3, 16)
    f = lambda t: np.cos(3 * t)  # noqa: E731
    assert np.array_equal(qp.discretize_on_midpoints(f, tlist), O.discretize_on_midpoints(f, tlist))
    assert np.allclose(qp.discretize(f, tlist), O.discretize(f, tlist), atol=1e-15)
    v = np.random.default_rng(0).standard_normal(16)
    assert np.allclose(qp.discretize_on_midpoints(v, tlist), O.discretize_on_midpoints(v, tlist), atol=1e-14)
    assert np.allclose(qp.discretize(v[:15], tlist), O.discretize(v[:15], tlist), atol=1e-15)
    for n in (1, 2, 8, 15):
        assert qp.t_mid(tlist, n) == O.t_mid(tlist, n)
    with pytest.raises(ValueError):
        qp.discretize(np.zeros(3), tlist)


def test_generators_host():
    rng = np.random.default_rng(1)
    mats = [rng.standard_normal((6, 6)) + 0j for _ in range(4)]
    e1 = lambda t: t  # noqa: E731
    e2 = np.linspace(0, 1, 10)
    gen = qp.hamiltonian(mats[0], (mats[1], e1), mats[2], (mats[3], e2))
    assert isinstance(gen, qp.Generator) and len(gen.ops) == 3 and len(gen.amplitudes) == 2
    assert np.array_equal(gen.ops[0], mats[0] + mats[2])  # drift terms are summed
    ctrls = qp.get_controls(gen)
    assert len(ctrls) == 2 and ctrls[0] is e1 and ctrls[1] is e2
    tlist = np.linspace(0, 1, 11)
    op = qp.evaluate(gen, tlist, 4)
    assert isinstance(op, qp.Operator) and op.ops is gen.ops
    assert np.allclose(op.toarray(), mats[0] + mats[2] + 0.35 * mats[1] + e2[3] * mats[3])
    qp.evaluate_(op, gen, tlist, 1, vals_dict=qp.IdDict([(e1, 2.0), (e2, -1.0)]))
    assert op.coeffs == [2.0, -1.0]
    merged = qp.hamiltonian((mats[1], e1), (mats[2], e1))
    assert len(merged.ops) == 1 and np.array_equal(merged.ops[0], mats[1] + mats[2])
    assert isinstance(qp.hamiltonian(mats[0], (mats[1], 3.0)), qp.Operator)
    assert qp.hamiltonian(mats[0]) is mats[0]
    assert qp.get_controls((mats[0],)) == ()
    with pytest.raises(ValueError):
        qp.Operator(mats[:1], [1.0, 2.0])
    with pytest.raises(ValueError):
        qp.Generator(mats[:2], [])
    with pytest.raises(AssertionError):
        qp.evaluate_(qp.Operator([mats[0], mats[1]], [1.0]), gen, tlist, 1)


def test_leja_and_newton_coeffs_host():
    rng = np.random.default_rng(2)
    func = lambda z: np.exp(-1j * z)  # noqa: E731
    leja_a, leja_b = np.zeros(5, dtype=complex), np.zeros(5, dtype=complex)
    a_a, a_b = np.zeros(3, dtype=complex), np.zeros(3, dtype=complex)
    n_a = n_b = na_a = na_b = 0
    radius = None
    for restart in range(4):
        cand = rng.standard_normal(15) + 0.1j * rng.standard_normal(15)
        if restart == 1:
            cand[3] = cand[7]  # duplicates: exercises the tie / zero-product path
        if radius is None:
            radius = qp.leja_radius(cand)
            assert radius == O.leja_radius(cand)
        n_a, leja_a = qp.extend_leja_(leja_a, n_a, cand.copy(), 5)
        n_b, leja_b = O.extend_leja(leja_b, n_b, cand.copy(), 5)
        assert n_a == n_b == 5 * (restart + 1)
        assert np.array_equal(leja_a[:n_a], leja_b[:n_b])
        na_a, a_a = qp.extend_newton_coeffs_(a_a, na_a, leja_a, func, n_a, radius)
        na_b, a_b = O.extend_newton_coeffs(a_b, na_b, leja_b, func, n_b, radius)
        assert na_a == na_b == n_a and np.array_equal(a_a[:na_a], a_b[:na_b])
    # the Newton interpolant reproduces func at the Leja points
    z = leja_a[:8]
    for zi in z:
        acc, prod = 0j, 1.0 + 0j
        for k in range(8):
            acc += a_a[k] * prod
            prod *= (zi - leja_a[k]) / radius
        assert abs(acc - func(zi)) < 1e-10


def test_hessenberg_eigenvalues_host():
    rng = np.random.default_rng(3)
    Hs = np.triu(rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6)), -1)
    for acc in (False, True):
        assert np.allclose(qp.diagonalize_hessenberg_matrix(Hs, 5, accumulate=acc),
                           O.diagonalize_hessenberg_matrix(Hs, 5, accumulate=acc))
    two = qp.diagonalize_hessenberg_matrix(Hs, 2)
    assert np.allclose(sorted(two, key=lambda z: (z.real, z.imag)),
                       sorted(np.linalg.eigvals(Hs[:2, :2]), key=lambda z: (z.real, z.imag)))
    assert len(qp.diagonalize_hessenberg_matrix(Hs, 4, accumulate=True)) == 10


def test_workloads_shapes():
    W = qp.workloads
    H0, H1, H2 = W.tfim_chain(6)
    assert H0.nnz == 64 and H1.nnz == 6 * 64 and H2.nnz == 64
    dense = (H0 + 0.3 * H1 - 0.2 * H2).toarray()
    assert np.allclose(dense, dense.conj().T)
    w = W.config2_tfim(6)
    ev = np.linalg.eigvalsh((H0 + 1.0 * H1 + 0.5 * H2).toarray())
    assert w["E_min"] <= ev[0] and ev[-1] <= w["E_max"]
    T0, T1, T2 = W.transmon_chain(3, 3)
    for T in (T0, T1, T2):
        assert abs(T - T.conj().T).max() < 1e-14
    L = W.config4_liouvillian(n_spins=3, nt=4)
    L0, L1 = L["ops"]
    assert L0.shape == (64, 64)
    # TDSE convention: d/dt vec(rho) = -i L vec(rho); trace preservation <=> vec(1)^T L = 0
    one = np.eye(8).reshape(-1, order="F")
    assert np.abs(one @ L0.toarray()).max() < 1e-13 and np.abs(one @ L1.toarray()).max() < 1e-13
    # commutator part: -i L0 rho matches -i[H, rho] + D(rho) on a random density matrix
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(8) + 1j * rng.standard_normal(8)
    rho = np.outer(psi, psi.conj())
    H0s, H1s, _ = W.tfim_chain(3)
    lhs = (L1.toarray() @ rho.reshape(-1, order="F")).reshape(8, 8, order="F")
    assert np.allclose(lhs, H1s.toarray() @ rho - rho @ H1s.toarray())
    assert W.optomech().shape == (55, 55)


# ---------------------------------------------------------------------------------------
# host-side math of qp_newton_step (pure C++, no device): against the oracle's restatement of
# src/arnoldi.jl:143-170 and src/newton.jl:97-214
# ---------------------------------------------------------------------------------------


def _host_lib():
    from qprop_b200 import _lib

    return _lib, _lib.load()


def test_library_ritz_values_match_oracle():
    import ctypes as C
    import importlib

    OA = importlib.import_module("oracle.arnoldi")
    _lib, lib = _host_lib()
    rng = np.random.default_rng(1)
    for trial in range(120):
        m = int(rng.integers(1, 41))
        ld = m + 1
        A = rng.standard_normal((ld, ld)) + 1j * rng.standard_normal((ld, ld))
        if trial % 3 == 0:  # Hermitian tridiagonal (Lanczos-like, real spectrum)
            A = (A + A.conj().T) / 2
            H = np.tril(np.triu(A, -1), 1)
        elif trial % 3 == 1:  # general upper Hessenberg
            H = np.triu(A, -1)
        else:  # nearly decoupled blocks (tiny sub-diagonal entries)
            H = np.triu(A, -1)
            for k in range(1, ld, 3):
                H[k, k - 1] *= 1e-13
        Hf = np.asfortranarray(H, dtype=np.complex128)
        for acc in (0, 1):
            ref = OA.diagonalize_hessenberg_matrix(H, m, accumulate=bool(acc))
            out = np.zeros(m * (m + 1) // 2, dtype=np.complex128)
            n = C.c_int32()
            rc = lib.qp_diagonalize_hessenberg(Hf.ctypes.data_as(C.c_void_p), ld, m, acc, out.ctypes.data_as(C.c_void_p), C.byref(n))
            assert rc == 0
            got = out[: n.value]
            assert n.value == len(ref)
            scale = max(1.0, float(np.max(np.abs(ref))))
            # as multisets (ties in the (real, imag) order may resolve differently at rounding level)
            used = np.zeros(len(ref), dtype=bool)
            for z in got:
                d = np.abs(ref - z)
                d[used] = np.inf
                k = int(np.argmin(d))
                assert d[k] < 1e-9 * scale * (m if trial % 3 else 1), (trial, m, d[k])
                used[k] = True


def test_library_leja_and_newton_coeffs_match_oracle():
    import ctypes as C
    import importlib

    ON = importlib.import_module("oracle.newton")
    _lib, lib = _host_lib()
    rng = np.random.default_rng(2)
    for trial in range(30):
        m = int(rng.integers(3, 12))
        cap = 10 * m + 1
        leja_ref = np.zeros(cap, dtype=np.complex128)
        a_ref = np.zeros(cap, dtype=np.complex128)
        leja_lib, a_lib = leja_ref.copy(), a_ref.copy()
        n_ref = na_ref = 0
        n_lib, na_lib = C.c_int32(0), C.c_int32(0)
        radius = None
        for restart in range(3):  # three restarts extend the same sequences
            pts = (rng.standard_normal(m * (m + 1) // 2) + 0.3j * rng.standard_normal(m * (m + 1) // 2)) * 2.0
            if radius is None:
                radius = ON.leja_radius(pts)
            p1, p2 = pts.copy(), pts.copy()
            n_ref, leja_ref = ON.extend_leja(leja_ref, n_ref, p1, m)
            assert lib.qp_extend_leja(leja_lib.ctypes.data_as(C.c_void_p), cap, C.byref(n_lib),
                                      p2.ctypes.data_as(C.c_void_p), len(p2), m) == 0
            assert n_lib.value == n_ref
            np.testing.assert_array_equal(leja_lib[:n_ref], leja_ref[:n_ref])  # same selection, bit for bit
            na_ref, a_ref = ON.extend_newton_coeffs(a_ref, na_ref, leja_ref, lambda z: np.exp(-1j * z), n_ref, radius)
            assert lib.qp_extend_newton_coeffs(a_lib.ctypes.data_as(C.c_void_p), cap, C.byref(na_lib),
                                               leja_lib.ctypes.data_as(C.c_void_p), n_ref, _lib.QP_FUNC_EXPMI, None, None,
                                               float(radius)) == 0
            assert na_lib.value == na_ref
            # divided differences of high order are ill-conditioned coefficient by coefficient
            # (the two sides round differently), the interpolant they define is not: compare
            # the Newton polynomials at points inside the cloud, and the leading coefficients
            assert np.max(np.abs(a_lib[:m] - a_ref[:m])) <= 1e-10 * np.max(np.abs(a_ref[:m]))
            for z in pts[:5] * 0.7:
                vals = []
                for a in (a_lib, a_ref):
                    acc, prod = 0.0j, 1.0 + 0.0j
                    for k in range(na_ref):
                        acc += a[k] * prod
                        prod *= (z - leja_ref[k]) / radius
                    vals.append(acc)
                assert abs(vals[0] - vals[1]) <= 1e-9 * max(1.0, abs(vals[1]))
        # callback form of func
        cb = _lib.NEWTON_FUNC(lambda z, out, user: (setattr(out[0], "re", float(np.exp(z[0].re + 1j * z[0].im).real)),
                                                     setattr(out[0], "im", float(np.exp(z[0].re + 1j * z[0].im).imag)), None)[2])
        a_cb, a_exp = np.zeros(cap, dtype=np.complex128), np.zeros(cap, dtype=np.complex128)
        for arr, fid, fn in ((a_cb, _lib.QP_FUNC_CALLBACK, C.cast(cb, C.c_void_p)), (a_exp, _lib.QP_FUNC_EXP, None)):
            na = C.c_int32(0)
            assert lib.qp_extend_newton_coeffs(arr.ctypes.data_as(C.c_void_p), cap, C.byref(na),
                                               leja_lib.ctypes.data_as(C.c_void_p), m, fid, fn, None, float(radius)) == 0
        assert np.max(np.abs(a_cb[:m] - a_exp[:m])) <= 1e-12 * np.max(np.abs(a_exp[:m]))


def test_no_vector_load_is_scheduled_above_the_pdl_wait():
    """Kernels launched with programmatic dependent launch may touch the vectors written by the
    previous term only after griddepcontrol.wait (SASS: ACQBULK).  ptxas is free to move
    NON-COHERENT loads (LDG...CONSTANT: __ldg, const __restrict__, ld.global.nc) above the wait --
    it did so with the first own-x load of the bit-flip kernel, a rare stale read -- so the vectors are
    read with plain ld.global (csrc/spmv.cuh: ld_x) and every load that legitimately sits in the
    prologue (matrix, tables, coefficients) is non-coherent or narrower than 128 bits.  This scans
    the built objects: no plain 128-bit global load (= a vector element) before the wait, and in the
    Krylov kernels (no prologue at all) no load whatsoever."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    csrc = os.path.dirname(_lib.LIB_PATH)
    objs = {name: os.path.join(csrc, name + ".o") for name in ("sparse", "bitflip", "krylov")}
    if not os.path.exists(cuobjdump) or not all(os.path.exists(p) for p in objs.values()):
        pytest.skip("cuobjdump or the built objects are not available")
    total = 0
    for name, path in objs.items():
        sass = subprocess.run([cuobjdump, "-sass", path], capture_output=True, text=True, check=True).stdout
        fn, before_wait, seen_wait = None, [], False
        kernels = {}
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if fn is not None and seen_wait:
                    kernels[fn] = before_wait
                fn, before_wait, seen_wait = m.group(1), [], False
                continue
            if "ACQBULK" in line:
                seen_wait = True
            elif not seen_wait and "LDG" in line:
                before_wait.append(line.strip())
        if fn is not None and seen_wait:
            kernels[fn] = before_wait
        total += len(kernels)
        for fn, loads in kernels.items():
            if name == "krylov":
                assert not loads, f"{fn}: load above griddepcontrol.wait: {loads[0]}"
            bad = [l for l in loads if re.search(r"LDG\.E(\.NA)?\.128 ", l)]
            assert not bad, f"{fn}: plain 128-bit load above griddepcontrol.wait: {bad[0]}"
    assert total >= 100   # the CSR, SELL-D, bit-flip and Krylov kernels all use the attribute


def test_format_constants_match_the_header():
    """The storage-format codes of the ctypes mirror are the header's."""
    text = open(os.path.join(ROOT, "include", "qprop.h")).read()
    header = {name: int(val) for name, val in re.findall(r"#define\s+(QP_FORMAT_[A-Z]+)\s+(\d+)", text)}
    assert len(header) >= 7
    for name, val in header.items():
        assert getattr(_lib, name) == val, name
        assert val in _lib.FORMAT_NAMES
