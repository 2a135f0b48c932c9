"""Oracle restatement of ``src/specrad.jl`` (module ``SpectralRange``).  Test infra only."""

from __future__ import annotations

import numpy as np

from .arnoldi import arnoldi, extend_arnoldi, diagonalize_hessenberg_matrix, _sorted_eigvals
from .generators import toarray


def random_state(H, rng=None) -> np.ndarray:
    """``random_state(H; rng)`` -- ``src/specrad.jl:153-158`` (draws differ from Julia's RNG;
    the reference's tests pin only properties of the result)."""
    rng = np.random.default_rng() if rng is None else rng
    N = H.shape[1]
    Psi = rng.random(N) * np.exp(2j * np.pi * rng.random(N))
    Psi /= np.linalg.norm(Psi)
    return Psi


def ritzvals(G, state, m_min: int, m_max=None, prec=1e-5, norm_min=1e-15, counter=None):
    """``ritzvals(G, state, m_min, m_max; prec, norm_min)`` -- ``src/specrad.jl:170-220``."""
    if m_max is None:
        m_max = 2 * m_min
    if m_max <= m_min:
        raise ValueError(f"m_max={m_max} must be smaller than m_min={m_min}")
    m = max(5, min(m_min, m_max - 1))

    Hess = np.zeros((m_max, m_max), dtype=np.complex128)
    q = [np.empty_like(state) for _ in range(m_max + 1)]

    m0 = m - 1
    m0 = arnoldi(Hess, q, m0, state, G, extended=False, norm_min=norm_min, counter=counter)
    eigenvals = diagonalize_hessenberg_matrix(Hess, m0)
    vr_lo0 = np.min(eigenvals.real)
    vr_hi0 = np.max(eigenvals.real)
    vi_hi0 = np.max(np.abs(eigenvals.imag))
    if m0 == m - 1:
        extend_arnoldi(Hess, q, m, G, norm_min=norm_min, counter=counter)
        eigenvals = diagonalize_hessenberg_matrix(Hess, m)
        vr_lo = np.min(eigenvals.real)
        vr_hi = np.max(eigenvals.real)
        vi_hi = np.max(np.abs(eigenvals.imag))
        e_lo = abs(1.0 - vr_lo / vr_lo0) if vr_lo0 != 0.0 else 0.0
        e_hi = abs(1.0 - vr_hi / vr_hi0) if vr_hi0 != 0.0 else 0.0
        e_im = abs(1.0 - vi_hi / vi_hi0) if vi_hi0 != 0.0 else 0.0
        while (e_lo > prec) or (e_hi > prec) or ((vi_hi0 > 1e-14) and e_im > prec):
            vr_lo0, vr_hi0, vi_hi0 = vr_lo, vr_hi, vi_hi
            m0 = m
            m = m + 1
            extend_arnoldi(Hess, q, m, G, norm_min=norm_min, counter=counter)
            if m == m0:
                break
            eigenvals = diagonalize_hessenberg_matrix(Hess, m)
            vr_lo = np.min(eigenvals.real)
            vr_hi = np.max(eigenvals.real)
            vi_hi = np.max(np.abs(eigenvals.imag))
            with np.errstate(divide="ignore", invalid="ignore"):
                e_lo = abs(1.0 - (vr_lo / vr_lo0))
                e_hi = abs(1.0 - (vr_hi / vr_hi0))
                e_im = abs(1.0 - (vi_hi / vi_hi0))
            if m == m_max:
                break
    return eigenvals


def specrange(H, method="auto", **kwargs):
    """``specrange(H; method, kwargs...)`` -- ``src/specrad.jl:36-140``.

    ``:auto`` -> ``:manual`` if both E_min/E_max are given, ``:diag`` for size <= 32,
    else ``:arnoldi``.
    """
    method = str(method).lstrip(":")
    if method == "auto":  # :44-61
        if "E_min" in kwargs and "E_max" in kwargs:
            return specrange(H, "manual", **kwargs)
        if H.shape[0] <= 32:
            return specrange(H, "diag", **kwargs)
        return specrange(H, "arnoldi", **kwargs)
    if method == "arnoldi":  # :88-112
        rng = kwargs.get("rng", None)
        state = kwargs.get("state", None)
        if state is None:
            state = random_state(H, rng=rng)
        m_max = kwargs.get("m_max", 60)
        m_min = max(5, min(kwargs.get("m_min", 25), m_max - 1))
        prec = kwargs.get("prec", 1e-3)
        norm_min = kwargs.get("norm_min", 1e-15)
        enlarge = kwargs.get("enlarge", True)
        R = ritzvals(
            H, state, m_min, m_max, prec=prec, norm_min=norm_min, counter=kwargs.get("counter")
        )
        E_min = float(R[0].real)
        E_max = float(R[-1].real)
        if enlarge and len(R) > 1:
            E_min = 2 * E_min - float(R[1].real)
            E_max = 2 * E_max - float(R[-2].real)
        return E_min, E_max
    if method == "diag":  # :122-126
        evals = _sorted_eigvals(toarray(H)).real
        return float(evals[0]), float(evals[-1])
    if method == "manual":  # :137-139
        if "E_min" not in kwargs or "E_max" not in kwargs:
            raise TypeError("specrange(H, :manual) requires keyword arguments E_min and E_max")
        return float(kwargs["E_min"]), float(kwargs["E_max"])
    raise ValueError(f"Unknown specrange method {method!r}")
