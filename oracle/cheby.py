"""Oracle restatement of ``src/cheby.jl`` (module ``Cheby``).  Test infrastructure only."""

from __future__ import annotations

import cmath

import numpy as np
from scipy.special import jv

from .generators import matvec


def cheby_coeffs(Delta: float, dt: float, limit: float = 1e-12) -> np.ndarray:
    """``cheby_coeffs(Δ, dt; limit)`` -- ``src/cheby.jl:25-39``.

    a_1 = J_0(α), a_k = 2 J_{k-1}(α), α = |Δ dt / 2|; coefficients are appended while the
    *previous* magnitude exceeds ``limit`` (so the first one <= limit is kept).
    """
    alpha = abs(0.5 * Delta * dt)
    coeffs = []
    a = float(jv(0, alpha))
    coeffs.append(a)
    eps = abs(a)
    i = 1
    while eps > limit:
        a = 2.0 * float(jv(i, alpha))
        coeffs.append(a)
        eps = abs(a)
        i += 1
    return np.asarray(coeffs, dtype=np.float64)


def cheby_coeffs_inplace(coeffs: np.ndarray, Delta: float, dt: float, limit: float = 1e-12):
    """``cheby_coeffs!(coeffs, Δ, dt, limit)`` -- ``src/cheby.jl:54-72``.

    Returns ``(n, coeffs)``; the array is grown by doubling like the reference's
    ``resize!`` (NumPy arrays cannot be resized in place, so it is returned).
    """
    alpha = abs(0.5 * Delta * dt)
    a = float(jv(0, alpha))
    coeffs[0] = a
    N = len(coeffs)
    eps = abs(a)
    n = 1
    while eps > limit:
        n += 1
        a = 2.0 * float(jv(n - 1, alpha))
        coeffs[n - 1] = a
        eps = abs(a)
        if n >= N:
            N *= 2
            coeffs = np.concatenate([coeffs, np.zeros(N - len(coeffs))])
    return n, coeffs


class ChebyWrk:
    """``ChebyWrk(Ψ, Δ, E_min, dt; limit)`` -- ``src/cheby.jl:87-124``."""

    def __init__(self, Psi, Delta: float, E_min: float, dt: float, limit: float = 1e-12):
        self.v0 = np.empty_like(Psi)
        self.v1 = np.empty_like(Psi)
        self.v2 = np.empty_like(Psi)
        self.coeffs = cheby_coeffs(Delta, dt, limit=limit)
        self.n_coeffs = len(self.coeffs)
        self.Delta = float(Delta)
        self.E_min = float(E_min)
        self.dt = float(dt)
        self.limit = float(limit)
        self.n_matvec = 0  # stands in for timing_data["matrix-vector product"].ncalls


def _isapprox(a: float, b: float) -> bool:
    # Julia's `≈` for Float64: rtol = sqrt(eps)
    return abs(a - b) <= np.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))


def cheby_inplace(Psi, H, dt: float, wrk: ChebyWrk, E_min=None, check_normalization=False):
    """``cheby!(Ψ, H, dt, wrk; E_min, check_normalization)`` -- ``src/cheby.jl:150-213``.

    Mutates ``Psi`` (ndarray complex128, shape (N,) or (N,B)) and returns it.
    """
    if E_min is None:
        E_min = wrk.E_min
    Delta = wrk.Delta
    beta = (Delta / 2) + E_min  # :156
    assert _isapprox(abs(dt), abs(wrk.dt)), (
        f"wrk was initialized for dt={wrk.dt}, not dt=abs({dt})"
    )  # :157
    c = -2j / Delta if dt > 0 else 2j / Delta  # :158-162
    a = wrk.coeffs
    eps = wrk.limit
    assert len(a) > 1, "Need at least 2 Chebychev coefficients"  # :165
    v0, v1, v2 = wrk.v0, wrk.v1, wrk.v2

    v0[...] = Psi  # :171
    Psi *= a[0]  # :172

    v1[...] = matvec(H, v0)  # :175-177
    wrk.n_matvec += 1
    v1 -= beta * v0  # :178
    v1 *= c  # :179

    Psi += a[1] * v1  # :182
    c *= 2  # :184

    for i in range(2, wrk.n_coeffs):  # :186
        v2[...] = matvec(H, v1)  # :189-191
        wrk.n_matvec += 1
        v2 -= beta * v1  # :192
        v2 *= c  # :193
        if check_normalization:  # :194-200
            map_norm = abs(np.vdot(v1, v2)) / (2 * np.linalg.norm(v1) ** 2)
            assert map_norm <= (1.0 + eps), f"Incorrect normalization (E_min={E_min}, Δ={Delta})"
        v2 += v0  # :202
        Psi += a[i] * v2  # :205
        v0, v1, v2 = v1, v2, v0  # :207

    Psi *= cmath.exp(-1j * beta * dt)  # :211
    return Psi


def cheby(Psi, H, dt: float, wrk: ChebyWrk, E_min=None, check_normalization=False):
    """``cheby(Ψ, H, dt, wrk)`` -- ``src/cheby.jl:224-276`` (non-mutating form)."""
    if E_min is None:
        E_min = wrk.E_min
    Delta = wrk.Delta
    beta = (Delta / 2) + E_min
    assert _isapprox(abs(dt), wrk.dt), f"wrk was initialized for dt={wrk.dt}, not dt=abs({dt})"
    c = -2j / Delta if dt > 0 else 2j / Delta
    a = wrk.coeffs
    eps = wrk.limit
    assert len(a) > 1, "Need at least 2 Chebychev coefficients"

    v0 = Psi
    out = a[0] * v0
    v1 = c * (matvec(H, v0) - beta * v0)
    wrk.n_matvec += 1
    out = out + a[1] * v1
    c *= 2
    for i in range(2, wrk.n_coeffs):
        v2 = matvec(H, v1)
        wrk.n_matvec += 1
        if check_normalization:
            v2 = c * (v2 - v1 * beta)
            map_norm = abs(np.vdot(v1, v2)) / (2 * np.linalg.norm(v1) ** 2)
            assert map_norm <= (1.0 + eps), f"Incorrect normalization (E_min={E_min}, Δ={Delta})"
            v2 = v2 + v0
        else:
            v2 = c * (v2 - beta * v1) + v0
        out = out + a[i] * v2
        v0, v1 = v1, v2
    return cmath.exp(-1j * beta * dt) * out
