"""Oracle restatement of ``src/newton.jl`` (module ``Newton``).  Test infrastructure only.

The reference's ``a`` and ``leja`` are 0-based ``OffsetVector``s, so NumPy indices here
equal the reference's indices.
"""

from __future__ import annotations

import numpy as np

from .arnoldi import arnoldi, diagonalize_hessenberg_matrix


class NewtonWrk:
    """``NewtonWrk(v0; m_max=10)`` -- ``src/newton.jl:23-60``."""

    def __init__(self, v0, m_max: int = 10):
        if m_max <= 2:
            raise ValueError("Newton propagation requires m_max > 2")
        if m_max >= v0.shape[0]:
            m_max = v0.shape[0] - 1
            if m_max <= 2:
                raise ValueError("Newton propagation requires state dimension > 2")
        self.arnoldi_vecs = [np.empty_like(v0) for _ in range(m_max + 1)]
        self.v = np.empty_like(v0)
        self.a = np.zeros(10 * m_max + 1, dtype=np.complex128)
        self.leja = np.zeros(10 * m_max + 1, dtype=np.complex128)
        self.radius = 0.0
        self.n_a = 0
        self.n_leja = 0
        self.restarts = 0
        self.m_max = m_max
        self.n_matvec = [0]


def leja_radius(z) -> float:
    """``leja_radius(z)`` -- ``src/newton.jl:67-70``."""
    return 1.2 * float(np.max(np.abs(z)))


def extend_leja(leja, n: int, newpoints, n_use: int):
    """``extend_leja!(leja, n, newpoints, n_use)`` -- ``src/newton.jl:97-148``.

    Returns ``(n + n_use, leja)`` (the array may have been re-allocated).
    ``newpoints`` is clobbered like in the reference.
    """
    if len(leja) < n + n_use:
        new = np.zeros(2 * (n + n_use), dtype=np.complex128)
        new[:n] = leja[:n]
        leja = new
    u = len(newpoints) - 1
    i_add_start = 0
    if n == 0:
        # move the point of largest magnitude to the end of `newpoints` (:120-129)
        z_last = newpoints[u]
        for i in range(0, u):
            if abs(newpoints[i]) > abs(z_last):
                newpoints[u] = newpoints[i]
                newpoints[i] = z_last
                z_last = newpoints[u]
        leja[0] = newpoints[-1]
        i_add_start = 1
    exponent = 1.0 / (n + n_use)
    for i_add in range(i_add_start, n_use):
        p_max = 0.0
        i_max = 0
        for i in range(0, u - i_add + 1):
            p = 1.0
            for j in range(0, n + i_add):
                p = p * abs(newpoints[i] - leja[j]) ** exponent
            if p > p_max:
                p_max = p
                i_max = i
        leja[n + i_add] = newpoints[i_max]
        newpoints[i_max] = newpoints[u - i_add]
    return n + n_use, leja


def extend_newton_coeffs(a, n_a: int, leja, func, n_leja: int, radius: float):
    """``extend_newton_coeffs!(a, n_a, leja, func, n_leja, radius)`` --
    ``src/newton.jl:176-214``; divided differences normalised by ``radius``.
    Returns ``(n_a, a)``."""
    m = n_leja - n_a
    n0 = n_a
    if len(a) < n_a + m:
        new = np.zeros(2 * n_leja, dtype=np.complex128)
        new[:n_a] = a[:n_a]
        a = new
    assert len(leja) >= n_leja
    assert radius > 0
    if n_a == 0:
        a[0] = func(leja[0])
        n0 = 1
    for k in range(n0, n_a + m):
        d = 1.0 + 0j
        pn = 0j
        for n in range(1, k):
            zd = leja[k] - leja[n - 1]
            d = d * zd / radius
            pn = pn + a[n] * d
        zd = leja[k] - leja[k - 1]
        d = d * zd / radius
        assert abs(d) > 1e-200, "Divided differences too small"
        a[k] = (func(leja[k]) - a[0] - pn) / d
    return n_a + m, a


def newton_inplace(
    Psi,
    H,
    dt: float,
    wrk: NewtonWrk,
    func=None,
    norm_min: float = 1e-14,
    relerr: float = 1e-12,
    max_restarts: int = 50,
):
    """``newton!(Ψ, H, dt, wrk; func, norm_min, relerr, max_restarts)`` --
    ``src/newton.jl:246-385``.  Mutates and returns ``Psi``."""
    if func is None:
        func = lambda z: np.exp(-1j * z)  # noqa: E731   (:247)

    m_max = len(wrk.arnoldi_vecs) - 1
    m = m_max
    wrk.a[...] = 0
    wrk.leja[...] = 0
    Hess = np.zeros((m_max + 1, m_max + 1), dtype=np.complex128)
    _dt = float(dt)
    assert _dt != 0.0

    n_a = 0
    n_leja = 0
    wrk.v[...] = Psi

    s = 0
    beta = float(np.linalg.norm(wrk.v))
    wrk.v *= 1 / beta

    while True:  # restart loop :274
        m = arnoldi(
            Hess,
            wrk.arnoldi_vecs,
            m,
            wrk.v,
            H,
            _dt,
            extended=True,
            norm_min=norm_min,
            counter=wrk.n_matvec,
        )
        if m == 1 and s == 0:  # :289-295
            lam = beta * Hess[0, 0]
            Psi *= func(lam)
            break
        ritz = diagonalize_hessenberg_matrix(Hess, m, accumulate=True)  # :297

        if s == 0:
            wrk.radius = leja_radius(ritz)  # :301-303

        n_s = n_leja
        n_leja, wrk.leja = extend_leja(wrk.leja, n_leja, ritz, m)  # :308

        n_a, wrk.a = extend_newton_coeffs(wrk.a, n_a, wrk.leja, func, n_leja, wrk.radius)  # :314
        assert n_a == n_leja

        # Newton polynomial in the extended (m+1)x(m+1) Hessenberg matrix :330-343
        Hm = Hess[: m + 1, : m + 1]
        R = np.zeros(m + 1, dtype=np.complex128)
        P = np.zeros(m + 1, dtype=np.complex128)
        R[0] = beta
        P[0] = wrk.a[n_s] * beta
        for k in range(1, m):
            z = wrk.leja[n_s + k - 1]
            R = (Hm @ R - z * R) / wrk.radius
            P += wrk.a[n_s + k] * R

        if s == 0:
            Psi[...] = 0  # :346-348
        for i in range(m):  # :350-352
            Psi += P[i] * wrk.arnoldi_vecs[i]

        # starting vector for the next restart :356-367
        R = (Hm @ R - wrk.leja[n_s + m - 1] * R) / wrk.radius
        beta = np.float64(np.linalg.norm(np.abs(R)))
        with np.errstate(all="ignore"):  # Julia: 1/0.0 == Inf (exhausted Krylov space), no exception
            R *= 1 / beta
        wrk.arnoldi_vecs[0][...] = wrk.v
        wrk.v *= R[0]
        for i in range(1, m + 1):
            wrk.v += R[i] * wrk.arnoldi_vecs[i]

        Psi_relerr = beta * abs(wrk.a[n_a - 1]) / (1 + np.linalg.norm(Psi))  # :370
        if Psi_relerr < relerr:
            break
        s += 1
        assert s <= max_restarts, "Newton propagation did not converge within max_restarts"

    wrk.restarts = s
    wrk.n_leja = n_leja
    wrk.n_a = n_a
    return Psi
