"""Oracle restatement of ``src/arnoldi.jl`` (module ``Arnoldi``).  Test infrastructure only.

``Hess`` is a complex128 ndarray; ``q`` a list of m+1 state arrays.  Indices in comments
are the reference's 1-based ones.
"""

from __future__ import annotations

import numpy as np

from .generators import matvec


def arnoldi(Hess, q, m: int, Psi, H, dt: float = 1.0, extended=True, norm_min=1e-15, counter=None):
    """``arnoldi!(Hess, q, m, Ψ, H, dt; extended, norm_min)`` -- ``src/arnoldi.jl:60-100``.

    Modified Gram-Schmidt; returns the (possibly reduced) Krylov dimension ``m``.
    """
    dim_hess = m + 1 if extended else m
    assert Hess.shape[0] >= dim_hess and Hess.shape[1] >= dim_hess  # :76
    assert len(q) >= m + 1  # :77
    Hess[...] = 0  # :78
    q[0][...] = Psi  # :79
    for j in range(m):  # j+1 is the reference's j
        q[j + 1][...] = matvec(H, q[j])  # :81-83
        if counter is not None:
            counter[0] += 1
        for i in range(j + 1):  # :84-87
            Hess[i, j] = dt * np.vdot(q[i], q[j + 1])
            q[j + 1] -= (Hess[i, j] / dt) * q[i]
        if (j + 1 < m) or extended:  # :88
            h = np.linalg.norm(q[j + 1])
            Hess[j + 1, j] = dt * h
            if h < norm_min:  # :91-95 dimensionality exhausted
                m = j + 1
                break
            q[j + 1] *= 1 / h  # :96
    return m


def extend_arnoldi(Hess, q, m: int, H, dt: float = 1.0, norm_min=1e-15, counter=None):
    """``extend_arnoldi!(Hess, q, m, H, dt; norm_min)`` -- ``src/arnoldi.jl:115-129``:
    grow an (m-1)x(m-1) Hessenberg matrix (from ``extended=false``) to m x m."""
    h = np.linalg.norm(q[m - 1])
    if h < norm_min:
        return m
    Hess[m - 1, m - 2] = dt * h
    q[m - 1] *= 1 / h
    q[m][...] = matvec(H, q[m - 1])
    if counter is not None:
        counter[0] += 1
    for i in range(m):
        Hess[i, m - 1] = dt * np.vdot(q[i], q[m])
        q[m] -= (Hess[i, m - 1] / dt) * q[i]
    assert np.all(Hess[m - 1, : m - 2] == 0.0)  # :127
    return Hess


def _sorted_eigvals(A) -> np.ndarray:
    """LAPACK ``eigvals`` with Julia's default ordering by (real, imag)."""
    ev = np.linalg.eigvals(A)
    order = np.lexsort((ev.imag, ev.real))
    return ev[order]


def diagonalize_hessenberg_matrix(Hess, m: int, accumulate=False) -> np.ndarray:
    """``diagonalize_hessenberg_matrix(Hess, m; accumulate)`` -- ``src/arnoldi.jl:143-170``."""
    j_min = 1 if accumulate else m
    size = (m * (m + 1)) // 2 if accumulate else m
    eigenvals = np.zeros(size, dtype=np.complex128)
    offset = 0
    for j in range(j_min, m + 1):
        if j == 1:
            eigenvals[0] = Hess[0, 0]
        elif j == 2:  # closed form :156-163
            a, c, b, d = Hess[0, 0], Hess[1, 0], Hess[0, 1], Hess[1, 1]
            s = np.sqrt(complex(a**2 + 4 * b * c - 2 * a * d + d**2))
            eigenvals[offset + 0] = 0.5 * (a + d - s)
            eigenvals[offset + 1] = 0.5 * (a + d + s)
        else:
            eigenvals[offset : offset + j] = _sorted_eigvals(Hess[:j, :j])
        offset += j
    return eigenvals
