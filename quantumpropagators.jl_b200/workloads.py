"""Synthetic inputs for the BASELINE.json configurations.

Pure host-side data construction (SciPy sparse matrices, NumPy vectors) with seeded
``numpy.random.default_rng``; no device code.  The reference's own random fixtures
(``QuantumControlTestUtils.RandomObjects``) are not vendored, so these draw their own
inputs with the same stated properties.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# ---------------------------------------------------------------------------------------
# config 1: random sparse Hermitian H0 + one control (test_specrad.jl:147-163 style)
# ---------------------------------------------------------------------------------------


def random_sparse_hermitian(N, density, spectral_radius, rng):
    """Random sparse Hermitian matrix whose spectrum lies within ±spectral_radius
    (scaled by a Gershgorin-free estimate: the exact dense spectrum for N <= 2000,
    a norm bound otherwise)."""
    nnz_target = int(density * N * N / 2)
    rows = rng.integers(0, N, nnz_target)
    cols = rng.integers(0, N, nnz_target)
    vals = rng.standard_normal(nnz_target) + 1j * rng.standard_normal(nnz_target)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(N, N)).tocsr()
    H = (A + A.conj().T) * 0.5
    H = H.tocsr()
    if N <= 2000:
        ev = np.linalg.eigvalsh(H.toarray())
        rho = max(abs(ev[0]), abs(ev[-1]))
    else:
        rho = abs(H).sum(axis=1).max()
    H = H * (spectral_radius / rho)
    H = H.tocsr()
    H.sort_indices()
    return H


def config1_random(N=1000, density=0.1, rho=10.0, seed=1000, nt=501, T=10.0):
    """Random dynamic generator H0 + u(t) H1 with spectral envelope rho for |u| <= 1,
    500 steps, manual spectral range [-rho, rho]."""
    rng = np.random.default_rng(seed)
    H0 = random_sparse_hermitian(N, density, rho / 2, rng)
    H1 = random_sparse_hermitian(N, density, rho / 2, rng)
    psi0 = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    psi0 /= np.linalg.norm(psi0)
    tlist = np.linspace(0.0, T, nt)

    def u1(t):
        return np.sin(2 * np.pi * t / T)

    return dict(
        ops=[H0, H1],
        controls=[u1],
        psi0=psi0,
        tlist=tlist,
        E_min=-rho,
        E_max=rho,
        name=f"config1_random_N{N}",
    )


# ---------------------------------------------------------------------------------------
# config 2: transverse-field Ising chain
# ---------------------------------------------------------------------------------------


def tfim_chain(n_spins, J=1.0, dtype=np.complex128):
    """H0 = -J Σ_{i<n-1} Z_i Z_{i+1} (diagonal), H1 = Σ_i X_i (n nnz/row, col = row XOR 2^i,
    value 1), H2 = Σ_i Z_i (diagonal).  Bit i of the basis index is spin i (|0> = Z=+1)."""
    N = 1 << n_spins
    idx = np.arange(N, dtype=np.int64)
    z = [1.0 - 2.0 * ((idx >> i) & 1) for i in range(n_spins)]
    d0 = np.zeros(N)
    for i in range(n_spins - 1):
        d0 -= J * z[i] * z[i + 1]
    d2 = np.zeros(N)
    for i in range(n_spins):
        d2 += z[i]
    # diagonal operators store all N entries (nnz = N), including exact zeros
    ptr = np.arange(N + 1, dtype=np.int64)
    H0 = sp.csr_matrix((d0.astype(dtype), idx.copy(), ptr), shape=(N, N))
    H2 = sp.csr_matrix((d2.astype(dtype), idx.copy(), ptr.copy()), shape=(N, N))
    # H1: row r has columns r ^ (1<<i); build CSR directly with sorted columns
    cols = np.empty((N, n_spins), dtype=np.int64)
    for i in range(n_spins):
        cols[:, i] = idx ^ (1 << i)
    cols.sort(axis=1)
    indptr = np.arange(0, (N + 1) * n_spins, n_spins, dtype=np.int64)
    data = np.ones(N * n_spins, dtype=dtype)
    H1 = sp.csr_matrix((data, cols.reshape(-1), indptr), shape=(N, N))
    return H0, H1, H2


def config2_tfim(n_spins=20, nt=101, dt=0.1, J=1.0, seed=2000, random_state=True):
    """TFIM chain with two PWC controls.  Spectral envelope bound
    |E| <= J(n-1) + max|u1| n + max|u2| n."""
    H0, H1, H2 = tfim_chain(n_spins, J)
    N = 1 << n_spins
    T = dt * (nt - 1)
    tlist = np.linspace(0.0, T, nt)

    def u1(t):  # flat-top-like, in [0, 1]
        return float(np.sin(np.pi * t / T) ** 2) if T > 0 else 0.0

    def u2(t):  # in [-0.5, 0.5]
        return 0.5 * float(np.sin(4 * np.pi * t / T)) if T > 0 else 0.0

    rng = np.random.default_rng(seed)
    if random_state:
        psi0 = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        psi0 /= np.linalg.norm(psi0)
    else:
        psi0 = np.zeros(N, dtype=np.complex128)
        psi0[0] = 1.0
    bound = J * (n_spins - 1) + 1.0 * n_spins + 0.5 * n_spins
    return dict(
        ops=[H0, H1, H2],
        controls=[u1, u2],
        psi0=psi0,
        tlist=tlist,
        E_min=-bound,
        E_max=bound,
        name=f"config2_tfim_n{n_spins}",
    )


# ---------------------------------------------------------------------------------------
# config 3: transmon chain (n transmons x `levels` levels)
# ---------------------------------------------------------------------------------------


def _destroy(n):
    return sp.diags(np.sqrt(np.arange(1, n)).astype(np.complex128), 1, format="csr")


def _embed(op, q, n_sites, levels):
    """op on site q (site 0 = slowest index)."""
    left = sp.identity(levels**q, dtype=np.complex128, format="csr")
    right = sp.identity(levels ** (n_sites - q - 1), dtype=np.complex128, format="csr")
    return sp.kron(sp.kron(left, op, format="csr"), right, format="csr")


def transmon_chain(n_sites=8, levels=4, alpha=-0.3, J=0.05):
    """H0 = Σ_q [δ_q n_q + (α/2) n_q (n_q - 1)] + J Σ_q (a_q† a_{q+1} + h.c.), δ_q = 0.1 q;
    H1 = Σ_q (a_q + a_q†); H2 = i Σ_q (a_q† - a_q)."""
    a = _destroy(levels)
    ad = a.conj().T.tocsr()
    num = (ad @ a).tocsr()
    ident = sp.identity(levels, dtype=np.complex128, format="csr")
    N = levels**n_sites
    H0 = sp.csr_matrix((N, N), dtype=np.complex128)
    H1 = sp.csr_matrix((N, N), dtype=np.complex128)
    H2 = sp.csr_matrix((N, N), dtype=np.complex128)
    A = [_embed(a, q, n_sites, levels) for q in range(n_sites)]
    for q in range(n_sites):
        nq = _embed(num, q, n_sites, levels)
        anh = _embed((num @ (num - ident)).tocsr(), q, n_sites, levels)
        H0 = H0 + (0.1 * q) * nq + (alpha / 2) * anh
        if q + 1 < n_sites:
            hop = A[q].conj().T @ A[q + 1]
            H0 = H0 + J * (hop + hop.conj().T)
        H1 = H1 + (A[q] + A[q].conj().T)
        H2 = H2 + 1j * (A[q].conj().T - A[q])
    out = []
    for H in (H0, H1, H2):
        H = H.tocsr()
        H.eliminate_zeros()
        H.sort_indices()
        out.append(H)
    return tuple(out)


def config3_transmon(n_sites=8, levels=4, B=1024, nt=101, dt=0.5):
    """Ensemble of B trajectories; trajectory b scales both controls by
    s_b = 0.5 + b/(B-1)."""
    H0, H1, H2 = transmon_chain(n_sites, levels)
    N = levels**n_sites
    T = dt * (nt - 1)
    tlist = np.linspace(0.0, T, nt)

    def u1(t):
        return 0.05 * float(np.sin(np.pi * t / T) ** 2)

    def u2(t):
        return 0.05 * float(np.sin(2 * np.pi * t / T))

    scales = 0.5 + np.arange(B) / max(B - 1, 1)
    psi0 = np.zeros(N, dtype=np.complex128)
    psi0[0] = 1.0
    return dict(
        ops=[H0, H1, H2],
        controls=[u1, u2],
        scales=scales,
        psi0=psi0,
        tlist=tlist,
        name=f"config3_transmon_{n_sites}x{levels}_B{B}",
    )


# ---------------------------------------------------------------------------------------
# config 4: dissipative Liouvillian (vectorised density matrix, column stacking)
# ---------------------------------------------------------------------------------------


def ham_to_superop(H, convention="TDSE"):
    from .generators import ham_to_superop as f

    return f(H, convention)


def lindblad_to_superop(A, convention="TDSE"):
    from .generators import lindblad_to_superop as f

    return f(A, convention)


def config4_liouvillian(n_spins=12, gamma=0.05, J=1.0, nt=21, dt=0.05, seed=4000, matrix_free=False):
    """n-spin TFIM + local decay A_k = sqrt(γ) σ⁻_k, TDSE convention (func = exp(-i z));
    L0 = commutator(H0) + dissipator, L1 = commutator(Σ X_i).
    ``matrix_free``: the two super-operators as ``LeftRightOperator`` (n × n factors) instead of
    4^n × 4^n sparse matrices."""
    from .generators import liouvillian

    H0, H1, _ = tfim_chain(n_spins, J)
    NH = 1 << n_spins
    sm = sp.csr_matrix(np.array([[0, 1], [0, 0]], dtype=np.complex128))  # |0><1|
    c_ops = []
    for k in range(n_spins):
        left = sp.identity(1 << (n_spins - 1 - k), dtype=np.complex128, format="csr")
        right = sp.identity(1 << k, dtype=np.complex128, format="csr")
        c_ops.append(np.sqrt(gamma) * sp.kron(sp.kron(left, sm, format="csr"), right, format="csr"))
    T = dt * (nt - 1)
    tlist = np.linspace(0.0, T, nt)

    def u1(t):
        return float(np.sin(np.pi * t / T) ** 2)

    Lgen = liouvillian((H0, (H1, u1)), c_ops, convention="TDSE", matrix_free=matrix_free)
    L0, L1 = Lgen.ops
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal(NH) + 1j * rng.standard_normal(NH)
    psi /= np.linalg.norm(psi)
    rho0 = np.outer(psi, psi.conj()).reshape(-1, order="F")  # column stacking
    return dict(
        ops=[L0, L1],
        controls=[u1],
        psi0=rho0,
        tlist=tlist,
        name=f"config4_liouvillian_n{n_spins}" + ("_matrix_free" if matrix_free else ""),
    )


# ---------------------------------------------------------------------------------------
# config 5 / optomech fixture (reference test/optomech.jl:1-44)
# ---------------------------------------------------------------------------------------


def optomech(N_cav=4, N_mech=10, omega_mech=10.0, g=1.0, eta=2.0):
    """Sparse optomechanics Hamiltonian of the reference's deterministic test fixture
    (``test/optomech.jl``): Fock cut-offs N_cav, N_mech -> dimension (N_cav+1)(N_mech+1)."""
    Delta = -omega_mech

    def destroy(N):
        return sp.diags(np.sqrt(np.arange(1, N + 1)).astype(np.complex128), 1, format="csr")

    def ident(N):
        return sp.identity(N + 1, dtype=np.complex128, format="csr")

    a = sp.kron(destroy(N_cav), ident(N_mech), format="csr")
    at = a.conj().T.tocsr()
    b = sp.kron(ident(N_cav), destroy(N_mech), format="csr")
    bt = b.conj().T.tocsr()
    H_cav = -Delta * (at @ a) + eta * (a + at)
    H_mech = omega_mech * (bt @ b)
    H_int = -g * ((bt + b) @ at @ a)
    H = (H_cav + H_mech + H_int).tocsr()
    H.sort_indices()
    return H


def optomech_ket(n_cav, n_mech, N_cav=4, N_mech=10):
    psi = np.zeros((N_cav + 1) * (N_mech + 1), dtype=np.complex128)
    psi[n_cav * (N_mech + 1) + n_mech] = 1.0
    return psi


def config5_optomech_dense(N_cav=63, N_mech=127):
    """test/optomech.jl scaled to (N_cav+1)(N_mech+1) = 8192 and densified."""
    H = optomech(N_cav, N_mech)
    return np.asarray(H.toarray(), dtype=np.complex128, order="F")
