"""Newton propagation with restarted Arnoldi: host mirror of the reference's ``Arnoldi`` and
``Newton`` modules (``src/arnoldi.jl``, ``src/newton.jl``).

Split of work (north_star): the Krylov vectors live on the GPU and every pass over them is a
CUDA kernel (``qp_arnoldi``, ``qp_krylov_combine``); the small dense step -- Ritz values of
the (m+1)x(m+1) Hessenberg matrix, Leja ordering, divided differences, the Newton polynomial
in the Hessenberg matrix -- stays on the host, where ``func`` is an arbitrary closure that is
only ever evaluated at Leja points.
"""

from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib as L
from .cheby import _device_generator
from .device import DeviceState
from .generators import Operator

__all__ = [
    "KrylovWrk",
    "arnoldi_",
    "extend_arnoldi_",
    "diagonalize_hessenberg_matrix",
    "NewtonWrk",
    "newton_",
    "extend_leja_",
    "extend_newton_coeffs_",
    "leja_radius",
]


def _op_coeffs(H, gen, coeffs=None):
    if coeffs is None:
        coeffs = H.coeffs if isinstance(H, Operator) else []
    c = L.as_c128_array(coeffs)
    if c.size != gen.n_coeffs:
        raise ValueError(f"expected {gen.n_coeffs} operator coefficients, got {c.size}")
    return c


class KrylovWrk:
    """m_max+1 device-resident Arnoldi vectors bound to one generator (``qp_krylov_create``);
    the ``q`` array of ``arnoldi!`` (reference ``src/arnoldi.jl:60-77``)."""

    def __init__(self, like: DeviceState, H, m_max: int):
        self.ctx = like.ctx
        self.gen = _device_generator(H, like.ctx)
        self.m_max = int(m_max)
        self.n = like.n
        self.batch = like.batch
        lib = self.ctx._lib
        h = C.c_void_p()
        L.check(lib.qp_krylov_create(self.gen.handle, like.handle, self.m_max, C.byref(h)), self.ctx.handle)
        self.handle = h
        self._finalizer = weakref.finalize(self, lib.qp_krylov_destroy, h)

    def combine(self, weights, first: int, st: DeviceState, accumulate: bool) -> DeviceState:
        """st ← (accumulate ? st : 0) + Σ_i weights[i] q_{first+i}; for a bundle of B states ``weights``
        is (n_w, B): state b uses ``weights[:, b]``."""
        w = L.as_c128_array(weights)
        n_w = w.shape[0] if self.batch > 1 else w.size
        if self.batch > 1 and w.shape != (n_w, self.batch):
            raise ValueError(f"weights must have shape (n_w, {self.batch}) for a bundle")
        L.check(
            self.ctx._lib.qp_krylov_combine(self.handle, L.ptr(w), int(first), int(n_w), st.handle, 1 if accumulate else 0),
            self.ctx.handle,
        )
        return st

    def get(self, index: int, dst: DeviceState) -> DeviceState:
        L.check(self.ctx._lib.qp_krylov_get(self.handle, int(index), dst.handle), self.ctx.handle)
        return dst


def arnoldi_(Hess: np.ndarray, K: KrylovWrk, m: int, Psi: DeviceState, H, dt=1.0, extended=True, norm_min=1e-15,
             coeffs=None) -> int:
    """``m = arnoldi!(Hess, q, m, Ψ, H, dt; extended, norm_min)`` (reference
    ``src/arnoldi.jl:60-100``).  ``Hess`` is a square complex128 host array (any order); it is
    overwritten (zero outside the computed block).  Returns the possibly reduced ``m``."""
    c = _op_coeffs(H, K.gen, coeffs)
    if K.batch > 1:
        # bundle of states: Hess is (B, ld, ld), one Hessenberg matrix per state; returns the B dimensions
        if Hess.ndim != 3 or Hess.shape[0] != K.batch or Hess.shape[1] != Hess.shape[2]:
            raise ValueError(f"Hess must have shape ({K.batch}, ld, ld) for a bundle of {K.batch} states")
        ld = Hess.shape[1]
        buf = np.zeros((K.batch, ld, ld), dtype=np.complex128)  # [b][column-major ld x ld]
        m_out = np.zeros(K.batch, dtype=np.int32)
        L.check(
            K.ctx._lib.qp_arnoldi(
                K.handle, L.ptr(c), Psi.handle, int(m), float(dt), 1 if extended else 0, float(norm_min),
                buf.ctypes.data_as(C.c_void_p), ld, m_out.ctypes.data_as(C.POINTER(C.c_int32)),
            ),
            K.ctx.handle,
        )
        Hess[...] = buf.transpose(0, 2, 1)  # column-major matrices -> Hess[b, i, j]
        return m_out
    ld = Hess.shape[0]
    if Hess.shape[0] != Hess.shape[1]:
        raise ValueError("Hess must be square")
    buf = np.zeros((ld, ld), dtype=np.complex128, order="F")
    m_out = C.c_int32()
    L.check(
        K.ctx._lib.qp_arnoldi(
            K.handle, L.ptr(c), Psi.handle, int(m), float(dt), 1 if extended else 0, float(norm_min),
            buf.ctypes.data_as(C.c_void_p), ld, C.byref(m_out),
        ),
        K.ctx.handle,
    )
    Hess[...] = buf
    return m_out.value


def extend_arnoldi_(Hess: np.ndarray, K: KrylovWrk, m: int, H, dt=1.0, norm_min=1e-15, coeffs=None) -> bool:
    """``extend_arnoldi!(Hess, q, m, H, dt; norm_min)`` (reference ``src/arnoldi.jl:115-129``):
    grow an (m-1)x(m-1) decomposition to m x m.  Returns False if the Krylov space was
    exhausted (nothing changed)."""
    ld = Hess.shape[0]
    buf = np.array(Hess, dtype=np.complex128, order="F")
    done = C.c_int32()
    c = _op_coeffs(H, K.gen, coeffs)
    L.check(
        K.ctx._lib.qp_arnoldi_extend(
            K.handle, L.ptr(c), int(m), float(dt), float(norm_min), buf.ctypes.data_as(C.c_void_p), ld, C.byref(done)
        ),
        K.ctx.handle,
    )
    Hess[...] = buf
    if done.value and m >= 3 and not np.all(Hess[m - 1, : m - 2] == 0.0):
        raise AssertionError("Hessenberg matrix has entries below the first sub-diagonal")
    return bool(done.value)


def _eigvals_sorted(A) -> np.ndarray:
    # LAPACK eigenvalues in Julia's default order: by (real, imag)
    ev = np.linalg.eigvals(A)
    return ev[np.lexsort((ev.imag, ev.real))]


def diagonalize_hessenberg_matrix(Hess, m: int, accumulate=False) -> np.ndarray:
    """Eigenvalues of the leading m x m block, or of all leading blocks 1..m concatenated
    (reference ``src/arnoldi.jl:143-170``)."""
    sizes = range(1, m + 1) if accumulate else (m,)
    out = []
    for j in sizes:
        if j == 1:
            out.append(np.array([Hess[0, 0]], dtype=np.complex128))
        elif j == 2:
            a, b, c, d = Hess[0, 0], Hess[0, 1], Hess[1, 0], Hess[1, 1]
            s = np.sqrt(complex(a * a + 4 * b * c - 2 * a * d + d * d))
            out.append(np.array([0.5 * (a + d - s), 0.5 * (a + d + s)], dtype=np.complex128))
        else:
            out.append(_eigvals_sorted(Hess[:j, :j]))
    return np.concatenate(out)


def leja_radius(z) -> float:
    """reference ``src/newton.jl:67-70``"""
    return 1.2 * float(np.max(np.abs(z)))


def extend_leja_(leja: np.ndarray, n: int, newpoints: np.ndarray, n_use: int):
    """``extend_leja!(leja, n, newpoints, n_use)`` (reference ``src/newton.jl:97-148``): append
    ``n_use`` of the candidate points in Leja order (each maximising the product of distances
    to all points chosen so far).  Returns ``(n + n_use, leja)``; ``newpoints`` is clobbered."""
    if len(leja) < n + n_use:
        grown = np.zeros(2 * (n + n_use), dtype=np.complex128)
        grown[:n] = leja[:n]
        leja = grown
    cand = newpoints
    u = len(cand) - 1
    start = 0
    if n == 0:
        # the candidate of largest magnitude starts the sequence; the reference bubbles it to
        # the end of `newpoints` with pairwise swaps -- replay them so ties resolve identically
        z_last = cand[u]
        for i in range(u):
            if abs(cand[i]) > abs(z_last):
                cand[u], cand[i] = cand[i], z_last
                z_last = cand[u]
        leja[0] = cand[u]
        start = 1
    exponent = 1.0 / (n + n_use)
    for i_add in range(start, n_use):
        pool = cand[: u - i_add + 1]
        dist = np.abs(pool[:, None] - leja[None, : n + i_add]) ** exponent
        p = np.ones(len(pool))
        for j in range(dist.shape[1]):  # same left-to-right product order as the reference
            p = p * dist[:, j]
        i_max = int(np.argmax(p)) if np.max(p) > 0.0 else 0
        leja[n + i_add] = pool[i_max]
        cand[i_max] = cand[u - i_add]
    return n + n_use, leja


def extend_newton_coeffs_(a: np.ndarray, n_a: int, leja: np.ndarray, func, n_leja: int, radius: float):
    """``extend_newton_coeffs!(a, n_a, leja, func, n_leja, radius)`` (reference
    ``src/newton.jl:176-214``): divided differences normalised by ``radius``.  Returns
    ``(n_leja, a)``."""
    if len(a) < n_leja:
        grown = np.zeros(2 * n_leja, dtype=np.complex128)
        grown[:n_a] = a[:n_a]
        a = grown
    if len(leja) < n_leja:
        raise AssertionError("not enough Leja points")
    if not radius > 0:
        raise AssertionError("radius must be positive")
    k0 = n_a
    if n_a == 0:
        a[0] = func(leja[0])
        k0 = 1
    for k in range(k0, n_leja):
        d = 1.0 + 0.0j
        pn = 0.0j
        for n in range(1, k):
            d = d * (leja[k] - leja[n - 1]) / radius
            pn = pn + a[n] * d
        d = d * (leja[k] - leja[k - 1]) / radius
        if not abs(d) > 1e-200:
            raise AssertionError("Divided differences too small")
        a[k] = (func(leja[k]) - a[0] - pn) / d
    return n_leja, a


class NewtonWrk:
    """``NewtonWrk(v0; m_max=10)`` (reference ``src/newton.jl:23-60``): m_max+1 Arnoldi vectors
    and the restart vector ``v`` on the device, Newton coefficients / Leja points on the host."""

    def __init__(self, v0: DeviceState, H, m_max: int = 10):
        if m_max <= 2:
            raise ValueError("Newton propagation requires m_max > 2")
        if m_max >= v0.n:
            m_max = v0.n - 1
            if m_max <= 2:
                raise ValueError("Newton propagation requires state dimension > 2")
        self.batch = v0.batch  # > 1: a bundle of states sharing the generator (library host step only)
        self.m_max = m_max
        self.krylov = KrylovWrk(v0, H, m_max)
        self.v = v0.similar()
        self.a = np.zeros(10 * m_max + 1, dtype=np.complex128)
        self.leja = np.zeros(10 * m_max + 1, dtype=np.complex128)
        self.radius = 0.0
        self.n_a = 0
        self.n_leja = 0
        self.restarts = 0


def _default_func(z):
    return np.exp(-1j * z)


def _newton_library(Psi, H, dt, wrk, func, norm_min, relerr, max_restarts, coeffs):
    """One ``qp_newton_step`` call: Arnoldi + vector updates on the device, the Ritz / Leja /
    divided-difference step in C++ inside the library (no per-restart host-language work)."""
    K = wrk.krylov
    c = _op_coeffs(H, K.gen, coeffs)
    cb = None
    if func is None or func is _default_func:
        func_id = L.QP_FUNC_EXPMI
    elif func is np.exp:
        func_id = L.QP_FUNC_EXP
    else:
        func_id = L.QP_FUNC_CALLBACK

        def _trampoline(z_ptr, out_ptr, _user):
            f = complex(func(complex(z_ptr[0].re, z_ptr[0].im)))
            out_ptr[0].re, out_ptr[0].im = f.real, f.imag

        cb = L.NEWTON_FUNC(_trampoline)
    restarts = C.c_int32()
    L.check(
        K.ctx._lib.qp_newton_step(
            K.handle, Psi.handle, wrk.v.handle, L.ptr(c), float(dt), func_id,
            C.cast(cb, C.c_void_p) if cb is not None else None, None,
            float(norm_min), float(relerr), int(max_restarts), C.byref(restarts),
        ),
        K.ctx.handle,
    )
    wrk.restarts = restarts.value
    # NewtonWrk bookkeeping like the reference (src/newton.jl:381-383)
    n_a, n_leja, radius = C.c_int32(), C.c_int32(), C.c_double()
    L.check(K.ctx._lib.qp_newton_last(K.handle, C.byref(n_a), C.byref(n_leja), C.byref(radius), None, None, 0), K.ctx.handle)
    if n_a.value > len(wrk.a):
        wrk.a = np.zeros(2 * n_a.value, dtype=np.complex128)
    if n_leja.value > len(wrk.leja):
        wrk.leja = np.zeros(2 * n_leja.value, dtype=np.complex128)
    cap = min(len(wrk.a), len(wrk.leja))
    L.check(K.ctx._lib.qp_newton_last(K.handle, None, None, None, L.ptr(wrk.a), L.ptr(wrk.leja), cap), K.ctx.handle)
    wrk.n_a, wrk.n_leja, wrk.radius = n_a.value, n_leja.value, radius.value
    return Psi


def newton_(Psi: DeviceState, H, dt, wrk: NewtonWrk, func=None, norm_min=1e-14, relerr=1e-12, max_restarts=50,
            coeffs=None, host_step="library") -> DeviceState:
    """``newton!(Ψ, H, dt, wrk; func, norm_min, relerr, max_restarts)`` (reference
    ``src/newton.jl:246-385``): Ψ ← func(H dt) Ψ in place on the device.

    ``host_step="library"`` (default) runs the whole restart loop inside ``qp_newton_step``;
    ``host_step="python"`` keeps the small dense step in this module (the fine-grained ABI the
    Julia wrapper uses: ``qp_arnoldi`` + ``qp_krylov_combine``), filling ``wrk.a`` / ``wrk.leja``."""
    if float(dt) == 0.0:
        raise AssertionError("dt must be non-zero")
    if host_step == "library":
        return _newton_library(Psi, H, dt, wrk, func, norm_min, relerr, max_restarts, coeffs)
    if Psi.batch != 1:
        raise ValueError('host_step="python" handles a single state; bundles run with host_step="library"')
    func = _default_func if func is None else func
    K = wrk.krylov
    m = wrk.m_max
    wrk.a[:] = 0
    wrk.leja[:] = 0
    Hess = np.zeros((wrk.m_max + 1, wrk.m_max + 1), dtype=np.complex128)
    dt = float(dt)
    if dt == 0.0:
        raise AssertionError("dt must be non-zero")
    n_a = n_leja = 0
    wrk.v.copyto(Psi)
    s = 0
    beta = wrk.v.norm()
    wrk.v.lmul(1.0 / beta)

    while True:
        m = arnoldi_(Hess, K, m, wrk.v, H, dt, extended=True, norm_min=norm_min, coeffs=coeffs)
        if m == 1 and s == 0:
            # v is an eigenvector: f(H dt) Ψ = f(λ) Ψ   (:289-295)
            Psi.lmul(func(beta * Hess[0, 0]))
            break
        ritz = diagonalize_hessenberg_matrix(Hess, m, accumulate=True)
        if s == 0:
            wrk.radius = leja_radius(ritz)
        n_s = n_leja
        n_leja, wrk.leja = extend_leja_(wrk.leja, n_leja, ritz, m)
        n_a, wrk.a = extend_newton_coeffs_(wrk.a, n_a, wrk.leja, func, n_leja, wrk.radius)

        # Newton polynomial in the extended Hessenberg matrix (:330-343)
        Hm = Hess[: m + 1, : m + 1]
        R = np.zeros(m + 1, dtype=np.complex128)
        R[0] = beta
        P = wrk.a[n_s] * R
        for k in range(1, m):
            R = (Hm @ R - wrk.leja[n_s + k - 1] * R) / wrk.radius
            P = P + wrk.a[n_s + k] * R

        # Ψ (+)= Σ_{i<m} P_i q_i   (:346-352)
        K.combine(P[:m], 0, Psi, accumulate=(s > 0))

        # restart vector v ← Σ_{i<=m} R_i q_i / β   (:356-367; q_0 still holds the old v)
        R = (Hm @ R - wrk.leja[n_s + m - 1] * R) / wrk.radius
        beta = float(np.linalg.norm(R))
        R = R / beta
        K.combine(R, 0, wrk.v, accumulate=False)

        # convergence: relative size of the last Newton term (:370-376)
        if beta * abs(wrk.a[n_a - 1]) / (1.0 + Psi.norm()) < relerr:
            break
        s += 1
        if s > max_restarts:
            raise L.QPropError(L.QP_ERR_NOT_CONVERGED, f"newton!: no convergence within max_restarts={max_restarts}")

    wrk.restarts = s
    wrk.n_leja = n_leja
    wrk.n_a = n_a
    return Psi
