"""Chebyshev propagation: host mirror of the reference's ``Cheby`` module
(``src/cheby.jl``).  The coefficient table is computed on the host (Bessel functions), the
recursion runs in ``qp_cheby_step`` (one fused CUDA kernel per Chebyshev term).
"""

from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
from scipy.special import jv

from . import _lib as L
from .device import DeviceGenerator, DeviceState
from .generators import Generator, Operator, ScaledOperator, _as_operator

__all__ = ["cheby_propagate_", "cheby_coeffs", "cheby_coeffs_", "ChebyWrk", "cheby_", "cheby"]


def cheby_coeffs(Delta, dt, limit=1e-12) -> np.ndarray:
    """``cheby_coeffs(Δ, dt; limit)`` (reference ``src/cheby.jl:25-39``): a_1 = J_0(α),
    a_k = 2 J_{k-1}(α) with α = |Δ dt / 2|, kept up to and including the first coefficient
    whose magnitude is <= limit."""
    alpha = abs(0.5 * Delta * dt)
    out = [float(jv(0, alpha))]
    k = 1
    while abs(out[-1]) > limit:
        out.append(2.0 * float(jv(k, alpha)))
        k += 1
    return np.array(out, dtype=np.float64)


def cheby_coeffs_(coeffs, Delta, dt, limit=1e-12):
    """``cheby_coeffs!(coeffs, Δ, dt, limit)`` (reference ``src/cheby.jl:54-72``).  Returns
    ``(n, coeffs)``: NumPy arrays cannot grow in place, so the (possibly re-allocated, doubled)
    array is returned next to the count."""
    new = cheby_coeffs(Delta, dt, limit)
    n = len(new)
    size = max(len(coeffs), 1)  # an empty array would never grow under doubling
    while size <= n:
        size *= 2
    if size != len(coeffs):
        coeffs = np.concatenate([coeffs, np.zeros(size - len(coeffs))])
    coeffs[:n] = new
    return n, coeffs


class ChebyWrk:
    """``ChebyWrk(Ψ, Δ, E_min, dt; limit)`` (reference ``src/cheby.jl:87-124``): device work
    vectors + the coefficient table.  ``Ψ`` is a DeviceState (only its shape is used);
    ``H`` names the generator whose device form the workspace is bound to."""

    def __init__(self, Psi: DeviceState, H, Delta, E_min, dt, limit=1e-12):
        if not isinstance(Psi, DeviceState):
            raise TypeError("ChebyWrk needs a DeviceState")
        self.ctx = Psi.ctx
        self.gen = _device_generator(H, Psi.ctx)
        # identity of the component operators the workspace is bound to (the device matrices are
        # immutable for its lifetime, like genop.ops in the reference, src/generators.jl:759)
        self._op_ids = [id(op) for op in H.ops] if isinstance(H, (Operator, Generator)) else None
        lib = self.ctx._lib
        h = C.c_void_p()
        L.check(lib.qp_cheby_create(self.gen.handle, Psi.handle, C.byref(h)), self.ctx.handle)
        self.handle = h
        self._finalizer = weakref.finalize(self, lib.qp_cheby_destroy, h)
        self.limit = float(limit)
        self.set_spectral_range(Delta, E_min, dt)

    def set_spectral_range(self, Delta, E_min, dt):
        """Recompute and re-upload the coefficients; the operators stay where they are
        (``reinit_prop!``, reference ``src/cheby_propagator.jl:272-292``)."""
        self.Delta = float(Delta)
        self.E_min = float(E_min)
        self.dt = float(dt)
        self.coeffs = cheby_coeffs(self.Delta, self.dt, limit=self.limit)
        self.n_coeffs = len(self.coeffs)
        L.check(
            self.ctx._lib.qp_cheby_set_coeffs(
                self.handle, L.ptr(self.coeffs), self.n_coeffs, self.Delta, self.E_min, abs(self.dt), self.limit
            ),
            self.ctx.handle,
        )

    @property
    def step_bytes(self) -> int:
        """Algorithmic bytes of one step: (n_coeffs-1)(M + 80 N B)."""
        b = C.c_int64()
        L.check(self.ctx._lib.qp_cheby_step_bytes(self.handle, C.byref(b)), self.ctx.handle)
        return b.value


def _device_generator(H, ctx) -> DeviceGenerator:
    if isinstance(H, DeviceGenerator):
        return H
    if isinstance(H, ScaledOperator):
        raise TypeError("Chebyshev propagation of a ScaledOperator is not supported; scale the coefficients")
    if isinstance(H, (Operator, Generator)):  # (not duck-typed: ndarray has its own to_device)
        return H.to_device(ctx)
    return _as_operator(H).to_device(ctx)


def _upload_table(wrk, E_min):
    L.check(
        wrk.ctx._lib.qp_cheby_set_coeffs(wrk.handle, L.ptr(wrk.coeffs), wrk.n_coeffs, wrk.Delta, float(E_min), abs(wrk.dt), wrk.limit),
        wrk.ctx.handle,
    )


def cheby_(Psi: DeviceState, H, dt, wrk: ChebyWrk, check_normalization=False, coeffs=None, per_trajectory=False,
           E_min=None):
    """``cheby!(Ψ, H, dt, wrk; check_normalization)`` (reference ``src/cheby.jl:150-213``):
    Ψ ← exp(-i H dt) Ψ in place on the device.

    ``H`` is an ``Operator`` (its ``coeffs`` are the per-interval numbers) or a bare matrix;
    ``coeffs`` overrides them, and with ``per_trajectory`` is an [n_coeffs][B] array giving each
    trajectory of a batched state its own amplitudes.  ``E_min`` overrides ``wrk.E_min`` for this
    call (reference ``src/cheby.jl:152``).  The matrices are the ones the workspace was built
    with: passing an ``Operator`` over different component matrices is an error.
    """
    if isinstance(H, Operator) and wrk._op_ids is not None and [id(op) for op in H.ops] != wrk._op_ids:
        raise ValueError("cheby!: H is made of different operators than the workspace was created with")
    if E_min is not None and float(E_min) != wrk.E_min:
        _upload_table(wrk, E_min)
        try:
            return cheby_(Psi, H, dt, wrk, check_normalization, coeffs, per_trajectory)
        finally:
            _upload_table(wrk, wrk.E_min)
    if coeffs is None:
        coeffs = H.coeffs if isinstance(H, Operator) else []
    c = L.as_c128_array(coeffs)
    n_c = wrk.gen.n_coeffs
    if per_trajectory:
        if c.shape != (n_c, Psi.batch):
            raise ValueError(f"per-trajectory coefficients must have shape ({n_c}, {Psi.batch})")
    elif c.size != n_c:
        raise ValueError(f"expected {n_c} operator coefficients, got {c.size}")
    L.check(
        wrk.ctx._lib.qp_cheby_step(
            wrk.handle, Psi.handle, L.ptr(c), 1 if per_trajectory else 0, float(dt), 1 if check_normalization else 0
        ),
        wrk.ctx.handle,
    )
    return Psi


def cheby_propagate_(Psi: DeviceState, wrk: ChebyWrk, coeff_table, dt, observables=(), norms=False,
                     per_trajectory=False):
    """The step loop of ``propagate`` (reference ``src/propagate.jl:283-344``) in one library
    call (``qp_cheby_propagate``): ``coeff_table[s]`` holds the operator coefficients of step
    ``s`` (shape ``[n_steps][n_coeffs]``, or ``[n_steps][n_coeffs][B]`` with ``per_trajectory``).
    ``observables`` are :class:`DeviceGenerator` objects without free coefficients whose
    expectation values are recorded on the device before the first and after every step.

    Returns ``(expvals, norms)``: ``expvals[s, k(, b)]`` complex, ``norms[s(, b)]`` float (``None``
    when not requested)."""
    n_c = wrk.gen.n_coeffs
    B = Psi.batch
    tbl = L.as_c128_array(coeff_table)
    n_steps = tbl.shape[0] if tbl.ndim > 1 or n_c > 0 else int(tbl.size)
    want = (n_steps, n_c, B) if per_trajectory else (n_steps, n_c)
    if n_c == 0:
        tbl = np.zeros(want, dtype=np.complex128)
    if tbl.shape != want:
        raise ValueError(f"coefficient table must have shape {want}, got {tbl.shape}")
    obs = list(observables)
    arr = (C.c_void_p * max(len(obs), 1))(*[o.handle for o in obs])
    ev = np.zeros((n_steps + 1, len(obs), B), dtype=np.complex128) if obs else None
    nr = np.zeros((n_steps + 1, B), dtype=np.float64) if norms else None
    L.check(
        wrk.ctx._lib.qp_cheby_propagate(
            wrk.handle, Psi.handle, L.ptr(tbl), 1 if per_trajectory else 0, int(n_steps), float(dt), len(obs),
            arr if obs else None, L.ptr(ev) if obs else None, L.ptr(nr) if norms else None,
        ),
        wrk.ctx.handle,
    )
    if B == 1:
        ev = ev[:, :, 0] if ev is not None else None
        nr = nr[:, 0] if nr is not None else None
    return ev, nr


def cheby(Psi: DeviceState, H, dt, wrk: ChebyWrk, **kwargs) -> DeviceState:
    """``cheby(Ψ, H, dt, wrk)`` (reference ``src/cheby.jl:224-276``): non-mutating form; returns
    a fresh state object and leaves ``Ψ`` untouched."""
    return cheby_(Psi.copy(), H, dt, wrk, **kwargs)
