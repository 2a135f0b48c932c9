"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerance (north_star): relative ‖ψ_gpu − ψ_ref‖ ≤ 1e-10 after the full time
grid; norm conservation 1e-12 per step for Hermitian generators.
"""

import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import oracle as O
from oracle.controls import IdDict as OIdDict

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def rand_state(rng, n, B=None):
    shape = (n,) if B is None else (n, B)
    psi = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return psi / np.linalg.norm(psi, axis=0)


def rand_sparse(rng, n, density, hermitian=True):
    A = sp.random(n, n, density=density, random_state=np.random.RandomState(rng.integers(1 << 30)), format="csr")
    A = A + 1j * sp.random(n, n, density=density, random_state=np.random.RandomState(rng.integers(1 << 30)), format="csr")
    if hermitian:
        A = (A + A.conj().T) * 0.5
    A = A.tocsr()
    A.sort_indices()
    return A


# ---------------------------------------------------------------------------------------
# level-1 verbs (src/interfaces/state.jl:24-47)
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("n,B", [(1, 1), (7, 1), (1000, 1), (4097, 1), (257, 3), (64, 40), (1 << 16, 1)])
def test_state_verbs(qp, ctx, n, B):
    rng = np.random.default_rng(n * 31 + B)
    x = rand_state(rng, n, None if B == 1 else B)
    y = rand_state(rng, n, None if B == 1 else B)
    dx, dy = qp.DeviceState.from_host(ctx, x), qp.DeviceState.from_host(ctx, y)
    np.testing.assert_array_equal(dx.to_host(), x)
    assert np.allclose(dx.norm(), np.linalg.norm(x, axis=0), rtol=1e-14)
    ref_dot = np.vdot(x, y) if B == 1 else np.einsum("ib,ib->b", x.conj(), y)
    assert np.allclose(dx.dot(dy), ref_dot, rtol=1e-12, atol=1e-15)
    a = 0.3 - 1.7j
    dy.axpy(a, dx)
    assert rel(dy.to_host(), y + a * x) < 1e-15
    dx.lmul(a)
    assert rel(dx.to_host(), a * x) < 1e-15
    z = dx.zero()
    assert np.all(z.to_host() == 0)
    c = dx.copy()
    assert c is not dx and np.array_equal(c.to_host(), dx.to_host())
    s = (dx + dy) - dy
    assert rel(s.to_host(), dx.to_host()) < 1e-14
    dx.fill(2.0 + 1j)
    assert np.all(dx.to_host() == 2.0 + 1j)
    if B > 1:  # partial column transfer
        sub = dy.download(1, B - 1)
        np.testing.assert_array_equal(sub, dy.to_host()[:, 1:])


# ---------------------------------------------------------------------------------------
# the multi-operator mul! and 3-arg dot (test/test_operator_linalg.jl:30-64)
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("fmt", ["csr", "sell"])
@pytest.mark.parametrize("layout", ["csr", "csc"])
@pytest.mark.parametrize("n,density", [(5, 0.6), (100, 0.1), (1000, 0.01), (777, 0.2)])
def test_operator_mul(qp, ctx, fmt, layout, n, density):
    rng = np.random.default_rng(n)
    H0 = rand_sparse(rng, n, density, hermitian=False)
    H1 = rand_sparse(rng, n, density, hermitian=True)
    H2 = sp.diags(rng.standard_normal(n) + 0j, 0, format="csr")
    ops = [H0, H1, H2]
    if layout == "csc":
        ops = [A.tocsc() for A in ops]
    coeffs = [0.7 - 0.2j, -1.3]
    gen = qp.DeviceGenerator(ctx, ops, 2, fmt)
    assert gen.format == fmt
    dense = H0.toarray() + coeffs[0] * H1.toarray() + coeffs[1] * H2.toarray()
    x = rand_state(rng, n)
    y0 = rand_state(rng, n)
    dx = qp.DeviceState.from_host(ctx, x)
    for alpha in (1.0, 2.0, 0.5 - 1j):
        for beta in (0.0, 1.0, 2.0):
            dy = qp.DeviceState.from_host(ctx, y0)
            gen.mul(dy, dx, coeffs, alpha, beta)
            assert rel(dy.to_host(), beta * y0 + alpha * (dense @ x)) < 1e-12
    # beta == 0 must not read y (NaN-safe like BLAS)
    dy = qp.DeviceState.from_host(ctx, np.full(n, np.nan + 0j))
    gen.mul(dy, dx, coeffs, 1.0, 0.0)
    assert rel(dy.to_host(), dense @ x) < 1e-12
    dy0 = qp.DeviceState.from_host(ctx, y0)
    assert abs(gen.dot(dy0, dx, coeffs) - np.vdot(y0, dense @ x)) < 1e-12 * n


def _structured_ops(rng, n, n_vals, offsets):
    """Operators with few distinct (value, column - row) pairs: band matrices whose entries are
    drawn from a table of `n_vals` complex values -- the structure SELL-D compresses."""
    table = rng.standard_normal(n_vals) + 1j * rng.standard_normal(n_vals)
    ops = []
    for offs in offsets:
        rows, cols, vals = [], [], []
        for d in offs:
            r = np.arange(max(0, -d), min(n, n - d))
            keep = rng.random(r.size) < 0.8  # ragged rows, some rows empty
            r = r[keep]
            rows.append(r)
            cols.append(r + d)
            vals.append(table[rng.integers(0, n_vals, r.size)])
        A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
        A.sort_indices()
        ops.append(A)
    return ops


@pytest.mark.parametrize("n,n_vals", [(33, 3), (1000, 20), (4099, 200), (5000, 400)])
def test_operator_mul_selld(qp, ctx, n, n_vals):
    """Dictionary-compressed storage: 8-bit (< 256 entries) and 16-bit codes, ragged and empty
    rows, N not a multiple of the slice height; exact same mul! as the uncompressed formats."""
    rng = np.random.default_rng(n + n_vals)
    offsets = [[0], [-7, 1, 2, 64], [-n // 2, -1, 3, n // 3]]
    ops = _structured_ops(rng, n, n_vals, offsets)
    coeffs = [0.7 - 0.2j, -1.3]
    gen = qp.DeviceGenerator(ctx, ops, 2, "selld")
    assert gen.format == "selld"
    ref_gen = qp.DeviceGenerator(ctx, ops, 2, "csr")
    dense = ops[0].toarray() + coeffs[0] * ops[1].toarray() + coeffs[1] * ops[2].toarray()
    x, y0 = rand_state(rng, n), rand_state(rng, n)
    dx = qp.DeviceState.from_host(ctx, x)
    for alpha, beta in ((1.0, 0.0), (0.5 - 1j, 2.0)):
        dy = qp.DeviceState.from_host(ctx, y0)
        gen.mul(dy, dx, coeffs, alpha, beta)
        assert rel(dy.to_host(), beta * y0 + alpha * (dense @ x)) < 1e-13
        dz = qp.DeviceState.from_host(ctx, y0)
        ref_gen.mul(dz, dx, coeffs, alpha, beta)
        assert rel(dy.to_host(), dz.to_host()) < 1e-14


@pytest.mark.parametrize("B", [1, 33])
def test_selld_explicit_diagonal(qp, ctx, B):
    """A diagonal drift with thousands of distinct energies on top of few-valued couplings (the
    transmon-chain shape): the diagonals are kept as explicit vectors, everything else is coded."""
    rng = np.random.default_rng(77 + B)
    n = 6000
    D0 = sp.diags(rng.standard_normal(n) + 0j, 0, format="csr")          # n distinct values
    D2 = sp.diags(rng.standard_normal(n) * (1 + 0.5j), 0, format="csr")   # controlled op with a diagonal too
    band = _structured_ops(rng, n, 6, [[-9, -1, 1, 9, 500]])[0]
    ops = [D0 + band, band.conj().T.tocsr(), D2]
    coeffs = [0.3 + 0.1j, -0.8]
    gen = qp.DeviceGenerator(ctx, ops, 2, "selld" if B == 1 else "auto")
    assert gen.n_dict > 0
    dense = sum(c * A for c, A in zip([1.0] + coeffs, ops)).tocsr()
    X, Y = rand_state(rng, n, None if B == 1 else B), rand_state(rng, n, None if B == 1 else B)
    dx = qp.DeviceState.from_host(ctx, X)
    for alpha, beta in ((1.0, 0.0), (0.5 - 1j, 2.0)):
        dy = qp.DeviceState.from_host(ctx, Y)
        gen.mul(dy, dx, coeffs, alpha, beta)
        assert rel(dy.to_host(), beta * Y + alpha * (dense @ X)) < 1e-13


def test_selld_refuses_incompressible(qp, ctx):
    """More than 4095 distinct entries: forcing the format is an error, AUTO falls back."""
    rng = np.random.default_rng(8)
    A = rand_sparse(rng, 2000, 0.01)
    with pytest.raises(qp.QPropError, match="distinct"):
        qp.DeviceGenerator(ctx, [A], 0, "selld")
    assert qp.DeviceGenerator(ctx, [A], 0).format == "csr"


@pytest.mark.parametrize("n,n_vals,B", [(100, 5, 16), (1000, 300, 33), (333, 40, 70), (500, 30, 129), (200, 300, 300)])
def test_operator_mul_batched_selld(qp, ctx, n, n_vals, B):
    """Trajectory-batched dictionary kernel (B >= 16, AUTO format): real, imaginary and complex
    entries, 8- and 16-bit codes, partial trajectory chunks; 1, 2 and 4 chunks of 32 trajectories
    per lane (B <= 32, <= 64, > 64)."""
    rng = np.random.default_rng(n + B)
    offsets = [[0], [-7, 1, 2, 64], [-n // 2, -1, 3, n // 3]]
    ops = _structured_ops(rng, n, n_vals, offsets)
    ops[0] = sp.csr_matrix(ops[0].real.astype(complex))   # purely real operator
    ops[1] = sp.csr_matrix(1j * ops[1].imag)              # purely imaginary operator
    coeffs = [0.7 - 0.2j, -1.3]
    gen = qp.DeviceGenerator(ctx, ops, 2)
    assert gen.n_dict > 0
    dense = ops[0].toarray() + coeffs[0] * ops[1].toarray() + coeffs[1] * ops[2].toarray()
    X, Y = rand_state(rng, n, B), rand_state(rng, n, B)
    dx = qp.DeviceState.from_host(ctx, X)
    for alpha, beta in ((1.0, 0.0), (0.5 - 1j, 2.0)):
        dy = qp.DeviceState.from_host(ctx, Y)
        gen.mul(dy, dx, coeffs, alpha, beta)
        assert rel(dy.to_host(), beta * Y + alpha * (dense @ X)) < 1e-13


def test_operator_mul_batched(qp, ctx):
    rng = np.random.default_rng(5)
    n, B = 300, 6
    H0 = rand_sparse(rng, n, 0.05)
    H1 = rand_sparse(rng, n, 0.05)
    gen = qp.DeviceGenerator(ctx, [H0, H1], 1)
    X = rand_state(rng, n, B)
    Y = rand_state(rng, n, B)
    dx, dy = qp.DeviceState.from_host(ctx, X), qp.DeviceState.from_host(ctx, Y)
    gen.mul(dy, dx, [0.4j], 2.0, -1.0)
    dense = H0.toarray() + 0.4j * H1.toarray()
    assert rel(dy.to_host(), -Y + 2.0 * dense @ X) < 1e-12


def test_host_operator_verbs(qp, ctx):
    """Operator / ScaledOperator objects on DeviceStates (src/generators.jl:634-708)."""
    rng = np.random.default_rng(9)
    n = 64
    H0, H1 = rand_sparse(rng, n, 0.2), rand_sparse(rng, n, 0.2)
    op = qp.Operator([H0, H1], [0.5])
    x = rand_state(rng, n)
    dx = qp.DeviceState.from_host(ctx, x)
    dense = H0.toarray() + 0.5 * H1.toarray()
    assert rel((op @ dx).to_host(), dense @ x) < 1e-12
    sop = 2.0j * op
    assert isinstance(sop, qp.ScaledOperator)
    assert rel((sop @ dx).to_host(), 2.0j * dense @ x) < 1e-12
    assert abs(sop.dot(dx, dx) - 2.0j * np.vdot(x, dense @ x)) < 1e-12
    assert (1.0 * op) is op


# ---------------------------------------------------------------------------------------
# Chebyshev
# ---------------------------------------------------------------------------------------


def _oracle_generator(ops, controls):
    terms = [ops[0]] + [(op, c) for op, c in zip(ops[1:], controls)]
    return O.hamiltonian(*terms)


def _product_generator(qp, ops, controls):
    terms = [ops[0]] + [(op, c) for op, c in zip(ops[1:], controls)]
    return qp.hamiltonian(*terms)


@pytest.mark.parametrize("fmt", ["csr", "sell"])
@pytest.mark.parametrize("backward", [False, True])
def test_cheby_config1_vs_oracle(qp, ctx, fmt, backward):
    """BASELINE config 1 shape (random sparse Hermitian + 1 control), reduced to N=300 / 60
    steps so the oracle finishes in seconds; manual spectral range so both sides use the same
    coefficient table."""
    w = qp.workloads.config1_random(N=300, density=0.1, nt=61, T=1.2, seed=11)
    kw = dict(E_min=w["E_min"], E_max=w["E_max"], backward=backward)
    ref = O.propagate(w["psi0"], _oracle_generator(w["ops"], w["controls"]), w["tlist"], "cheby", **kw)
    gen = _product_generator(qp, w["ops"], w["controls"])
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, matrix_format=fmt, **kw)
    assert p.wrk.gen.format == fmt
    out = qp.propagate(p)
    assert rel(out, ref) < RTOL
    assert abs(np.linalg.norm(out) - 1) < 1e-11


def test_cheby_single_step_norm_and_coeff_count(qp, ctx):
    w = qp.workloads.config1_random(N=1000, density=0.1, seed=1000)
    p = qp.init_prop(w["psi0"], _product_generator(qp, w["ops"], w["controls"]), w["tlist"], "cheby",
                     ctx=ctx, E_min=w["E_min"], E_max=w["E_max"])
    # test/test_specrad.jl:187-191: manual E=±10 -> E_min = -10.1, Δ = 20.2; n_coeffs = 9
    assert abs(p.wrk.E_min + 10.1) < 1e-12 and abs(p.wrk.Delta - 20.2) < 1e-12
    assert p.wrk.n_coeffs == 9
    for _ in range(5):
        st = qp.prop_step(p)
        assert st is p.state
        assert abs(st.norm() - 1.0) < 1e-12


def test_cheby_tls_analytic(qp, ctx):
    """test/test_propagate.jl:74-150: Rabi 3π/2 pulse, forward and back, 1e-12."""
    H = np.array([[0, 0.5], [0.5, 0]], dtype=complex)
    tlist = np.linspace(0, 1.5 * np.pi, 101)
    psi0 = np.array([1, 0], dtype=complex)
    out = qp.propagate(psi0, (H,), tlist, "cheby", ctx=ctx, inplace=False)
    expected = np.array([-1 / np.sqrt(2), -1j / np.sqrt(2)])
    assert np.linalg.norm(out - expected) < 1e-12
    back = qp.propagate(out, (H,), tlist, "cheby", ctx=ctx, inplace=False, backward=True)
    assert np.linalg.norm(back - psi0) < 1e-12


@pytest.mark.parametrize("n_spins", [6, 10, 14])
def test_cheby_tfim_vs_oracle_and_expm(qp, ctx, n_spins):
    """BASELINE config 2 shape at reduced size: H0 + u1 ΣX + u2 ΣZ, both storage formats."""
    from scipy.sparse.linalg import expm_multiply

    w = qp.workloads.config2_tfim(n_spins, nt=11, dt=0.1)
    kw = dict(E_min=w["E_min"], E_max=w["E_max"])
    ref = O.propagate(w["psi0"], _oracle_generator(w["ops"], w["controls"]), w["tlist"], "cheby", **kw)
    for fmt in ("csr", "sell", "selld", "bitflip"):
        gen = _product_generator(qp, w["ops"], w["controls"])
        p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, matrix_format=fmt, **kw)
        assert p.wrk.gen.format == fmt
        out = qp.propagate(p)
        assert rel(out, ref) < RTOL, fmt
    # independent ground truth: exact PWC propagation with expm_multiply
    psi = w["psi0"].copy()
    tl = w["tlist"]
    mids = O.get_tlist_midpoints(tl)
    for k in range(len(tl) - 1):
        Hk = w["ops"][0] + w["controls"][0](mids[k]) * w["ops"][1] + w["controls"][1](mids[k]) * w["ops"][2]
        psi = expm_multiply(-1j * (tl[k + 1] - tl[k]) * Hk.tocsc(), psi)
    assert rel(out, psi) < 1e-9


@pytest.mark.parametrize("n,n_vals,n_ops,B", [(300, 6, 3, 129), (1000, 200, 3, 200), (257, 10, 2, 300), (300, 6, 3, 40)])
def test_operator_mul_batched_shared_columns(qp, ctx, n, n_vals, n_ops, B):
    """Operators that hit the same columns (quadrature control pairs): the rows are ordered pair
    by pair and, for B > 64, the pair kernel gathers each shared column once.  Odd numbers of
    pairs, a column shared by all three operators, ragged rows, 8- and 16-bit codes."""
    rng = np.random.default_rng(n + B)
    shared = [-17, -2, 1, 5, 40]
    base = _structured_ops(rng, n, n_vals, [[0, 3, -n // 3, 1], shared, shared + [7]])
    ops = [sp.csr_matrix(base[0].real.astype(complex)),                  # drift: real, shares column +1 with both controls
           sp.csr_matrix(base[1].real.astype(complex)),                  # real control operator
           sp.csr_matrix(1j * base[2].imag)][:n_ops]                      # imaginary control operator on the same columns
    coeffs = ([0.7 - 0.2j, -1.3] if n_ops == 3 else [0.4 + 0.3j])
    gen = qp.DeviceGenerator(ctx, ops, len(coeffs))
    assert gen.n_dict > 0
    dense = ops[0].toarray() + sum(c * A.toarray() for c, A in zip(coeffs, ops[1:]))
    X, Y = rand_state(rng, n, B), rand_state(rng, n, B)
    dx = qp.DeviceState.from_host(ctx, X)
    for alpha, beta in ((1.0, 0.0), (0.5 - 1j, 2.0)):
        dy = qp.DeviceState.from_host(ctx, Y)
        gen.mul(dy, dx, coeffs, alpha, beta)
        assert rel(dy.to_host(), beta * Y + alpha * (dense @ X)) < 1e-13
    ev = gen.expval(dx, coeffs)
    assert np.max(np.abs(ev - np.einsum("ib,ib->b", X.conj(), dense @ X))) < 1e-12 * n
    # single state through the B = 1 kernel on the same (reordered) rows
    x1 = rand_state(rng, n)
    d1, dy1 = qp.DeviceState.from_host(ctx, x1), qp.DeviceState.from_host(ctx, x1)
    gen.mul(dy1, d1, coeffs, 1.0, 0.0)
    assert rel(dy1.to_host(), dense @ x1) < 1e-13


@pytest.mark.parametrize("B", [5, 40, 150])
def test_cheby_batched_per_trajectory(qp, ctx, B):
    """Ensemble: B trajectories with their own control scale share one coefficient table
    (B = 5: merged-CSR SpMM kernel, B = 40 / 150: dictionary SpMM kernel with 2 / 4 trajectory
    chunks per lane)."""
    rng = np.random.default_rng(3)
    w = qp.workloads.config3_transmon(n_sites=3, levels=3, B=B, nt=9, dt=0.5)
    H0, H1, H2 = w["ops"]
    tl = w["tlist"]
    dt = tl[1] - tl[0]
    psi0 = rand_state(rng, H0.shape[0], B)
    scales = w["scales"]
    # common spectral envelope (Gershgorin bound over the whole ensemble)
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
    Delta, E_min = 2 * bound, -bound
    gen = qp.DeviceGenerator(ctx, [H0, H1, H2], 2)
    st = qp.DeviceState.from_host(ctx, psi0)
    wrk = qp.ChebyWrk(st, gen, Delta, E_min, dt)
    mids = O.get_tlist_midpoints(tl)
    ref = psi0.copy()
    owrk = O.ChebyWrk(ref[:, 0].copy(), Delta, E_min, dt)
    for k in range(len(tl) - 1):
        u = np.array([[w["controls"][0](mids[k]) * s for s in scales], [w["controls"][1](mids[k]) * s for s in scales]])
        qp.cheby_(st, None, dt, wrk, coeffs=u, per_trajectory=True)
        for b in range(B):
            Hb = O.Operator([H0, H1, H2], [u[0, b], u[1, b]])
            col = ref[:, b].copy()
            O.cheby_inplace(col, Hb, dt, owrk)
            ref[:, b] = col
    out = st.to_host()
    for b in range(B):
        assert rel(out[:, b], ref[:, b]) < RTOL


# ---------------------------------------------------------------------------------------
# two-pass tiled path for batched states (csrc/tile.cu): batch a multiple of 32
# ---------------------------------------------------------------------------------------


def _dense_apply(ops, u, X):
    """sum_l u[l, b] (H_l X)[:, b]"""
    return sum(np.asarray(u[l])[None, :] * (ops[l] @ X) for l in range(len(ops)))


@pytest.mark.parametrize("sites,levels,B", [(2, 4, 32), (3, 4, 64), (4, 4, 96), (4, 3, 32), (4, 4, 256)])
def test_tiled_batched_operator_verbs(qp, ctx, sites, levels, B):
    """mul! (alpha, beta), the fused expectation value and 3-argument dot on the tiled path against
    NumPy, for splits S x N/S = 4x4 ... 16x16 and 1 ... 8 trajectory chunks; then the same generator
    with other batch sizes (the completion counters start a new epoch)."""
    rng = np.random.default_rng(sites * 100 + B)
    H0, H1, H2 = qp.workloads.transmon_chain(sites, levels)
    N = H0.shape[0]
    gen = qp.DeviceGenerator(ctx, [H0, H1, H2], 2)
    info = gen.tile_info()
    if levels == 4 and sites >= 3:  # (2 sites: every hop straddles the split -> too many class-O entries, one-pass kernels)
        assert info["available"] and info["split"] * info["blocks"] == N and info["entries"]["other"] > 0
    ops = [H0, H1, H2]
    for Bk in (B, 32, B):
        X = rand_state(rng, N, Bk)
        Y0 = rand_state(rng, N, Bk)
        c = [0.3 - 0.2j, -1.1 + 0.4j]
        alpha, beta = 0.7 - 0.1j, -0.4 + 0.9j
        x = qp.DeviceState.from_host(ctx, X)
        y = qp.DeviceState.from_host(ctx, Y0)
        gen.mul(y, x, c, alpha, beta)
        u = np.array([[1.0] * Bk, [c[0]] * Bk, [c[1]] * Bk])
        want = beta * Y0 + alpha * _dense_apply(ops, u, X)
        assert np.max(np.abs(y.to_host() - want)) < 1e-13 * np.max(np.abs(want)) * N ** 0.5
        gen.mul(y, x, c, 1.0, 0.0)
        assert np.max(np.abs(y.to_host() - _dense_apply(ops, u, X))) < 1e-13 * N ** 0.5
        ev = gen.expval(x, c)
        assert np.max(np.abs(ev - np.einsum("nb,nb->b", X.conj(), _dense_apply(ops, u, X)))) < 1e-12
        d = gen.dot(y, x, c)
        assert np.max(np.abs(d - np.einsum("nb,nb->b", y.to_host().conj(), _dense_apply(ops, u, X)))) < 1e-11


@pytest.mark.parametrize("B", [64, 160])
@pytest.mark.parametrize("backward", [False, True])
def test_tiled_cheby_per_trajectory(qp, ctx, B, backward):
    """Ensemble on the tiled path: per-trajectory amplitudes, all Chebyshev epilogues (FIRST / MID /
    LAST), forward and backward, the normalization check, against the oracle trajectory by trajectory."""
    rng = np.random.default_rng(5 + B)
    w = qp.workloads.config3_transmon(n_sites=4, levels=4, B=B, nt=6, dt=0.5)
    H0, H1, H2 = w["ops"]
    tl = w["tlist"]
    dt = (tl[1] - tl[0]) * (-1 if backward else 1)
    psi0 = rand_state(rng, H0.shape[0], B)
    scales = w["scales"]
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
    Delta, E_min = 2 * bound, -bound
    gen = qp.DeviceGenerator(ctx, [H0, H1, H2], 2)
    assert gen.tile_info()["available"]
    st = qp.DeviceState.from_host(ctx, psi0)
    wrk = qp.ChebyWrk(st, gen, Delta, E_min, abs(dt))
    assert wrk.n_coeffs > 3
    mids = O.get_tlist_midpoints(tl)
    ref = psi0.copy()
    owrk = O.ChebyWrk(ref[:, 0].copy(), Delta, E_min, abs(dt))
    for k in range(len(tl) - 1):
        u = np.array([[w["controls"][0](mids[k]) * s for s in scales], [w["controls"][1](mids[k]) * s for s in scales]])
        qp.cheby_(st, None, dt, wrk, coeffs=u, per_trajectory=True, check_normalization=(k == 0))
        for b in range(0, B, 7):
            Hb = O.Operator([H0, H1, H2], [u[0, b], u[1, b]])
            col = ref[:, b].copy()
            O.cheby_inplace(col, Hb, dt, owrk)
            ref[:, b] = col
    out = st.to_host()
    for b in range(0, B, 7):
        assert rel(out[:, b], ref[:, b]) < RTOL
    assert np.max(np.abs(st.norm() - 1)) < 1e-12


@pytest.mark.parametrize("n_ops", [1, 2])
def test_tiled_one_and_two_operators(qp, ctx, n_ops):
    """The NOPS = 1 / 2 instantiations of the tiled kernel: a drift-only generator and drift + one
    control (kind P = the pair (drift, control) when their entries share columns with |v| equal), all
    epilogues through a two-coefficient (CHEB_ONLY) and a many-coefficient Chebyshev step."""
    rng = np.random.default_rng(40 + n_ops)
    H0, H1, _ = qp.workloads.transmon_chain(4, 4)
    ops = [H0 + H1] if n_ops == 1 else [H0 + 0.5 * H1, H1]     # shared columns, equal and unequal magnitudes
    N, B = 256, 64
    gen = qp.DeviceGenerator(ctx, ops, n_ops - 1)
    assert gen.tile_info()["available"]
    X = rand_state(rng, N, B)
    x = qp.DeviceState.from_host(ctx, X)
    y = x.similar()
    c = [] if n_ops == 1 else [0.4 - 0.3j]
    gen.mul(y, x, c, 1.0, 0.0)
    Hd = (ops[0] if n_ops == 1 else ops[0] + c[0] * ops[1]).toarray()
    assert np.max(np.abs(y.to_host() - Hd @ X)) < 1e-12
    assert np.max(np.abs(gen.expval(x, c) - np.einsum("nb,nb->b", X.conj(), Hd @ X))) < 1e-12
    # Chebyshev with a real coefficient (Hermitian): tiny dt -> 2 coefficients (CHEB_ONLY), larger dt -> FIRST/MID/LAST
    cr = [] if n_ops == 1 else [0.4]
    Hh = (ops[0] if n_ops == 1 else ops[0] + cr[0] * ops[1]).toarray()
    bound = float(np.abs(Hh).sum(axis=1).max())
    for dt, n_min in ((1e-14, 2), (0.7, 4)):
        st = qp.DeviceState.from_host(ctx, X)
        wrk = qp.ChebyWrk(st, gen, 2 * bound, -bound, dt)
        assert (wrk.n_coeffs == 2) if n_min == 2 else (wrk.n_coeffs >= n_min)
        qp.cheby_(st, None, dt, wrk, coeffs=cr, check_normalization=True)
        want = sla.expm(-1j * dt * Hh) @ X
        assert np.max(np.linalg.norm(st.to_host() - want, axis=0)) < 1e-10


def test_tiled_normalization_check_fails_loudly(qp, ctx):
    """A spectral range that is too small makes the Chebyshev recursion blow up: with
    check_normalization the batched tiled path reports it (src/cheby.jl:194-200) instead of returning
    garbage."""
    rng = np.random.default_rng(3)
    H0, H1, H2 = qp.workloads.transmon_chain(4, 4)
    gen = qp.DeviceGenerator(ctx, [H0, H1, H2], 2)
    st = qp.DeviceState.from_host(ctx, rand_state(rng, 256, 32))
    wrk = qp.ChebyWrk(st, gen, 0.05, -0.025, 5.0)                # true spectral width is ~ 4
    with pytest.raises(qp.QPropError) as exc:
        qp.cheby_(st, None, 5.0, wrk, coeffs=[1.0, 1.0], check_normalization=True)
    assert exc.value.status == -5 and "normalization" in str(exc.value).lower()


def test_tiled_real_operators_and_unqualified_generators(qp, ctx):
    """TFIM (real operators, two diagonals, XOR couplings) on the tiled path; an unstructured random
    generator does not qualify and silently uses the one-pass kernels -- same results."""
    rng = np.random.default_rng(8)
    H0, H1, H2 = qp.workloads.tfim_chain(8)
    N, B = 256, 64
    X = rand_state(rng, N, B)
    gen = qp.DeviceGenerator(ctx, [H0, H1, H2], 2)
    info = gen.tile_info()
    assert info["available"] and info["entries"]["other"] == 0 and info["entries"]["diag"] == N
    x = qp.DeviceState.from_host(ctx, X)
    y = x.similar()
    gen.mul(y, x, [0.6, -0.3])
    u = np.array([[1.0] * B, [0.6] * B, [-0.3] * B])
    assert np.max(np.abs(y.to_host() - _dense_apply([H0, H1, H2], u, X))) < 1e-12
    R = rand_sparse(rng, 256, 0.1)
    g2 = qp.DeviceGenerator(ctx, [R], 0)
    assert not g2.tile_info()["available"]
    g2.mul(y, x, [])
    assert np.max(np.abs(y.to_host() - R @ X)) < 1e-12


def test_cheby_check_normalization_and_errors(qp, ctx):
    rng = np.random.default_rng(4)
    n = 200
    H = rand_sparse(rng, n, 0.1)
    ev = np.linalg.eigvalsh(H.toarray())
    psi = rand_state(rng, n)
    st = qp.DeviceState.from_host(ctx, psi)
    gen = qp.DeviceGenerator(ctx, [H], 0)
    good = qp.ChebyWrk(st, gen, (ev[-1] - ev[0]) * 1.01, ev[0] - 0.005 * (ev[-1] - ev[0]), 0.1)
    qp.cheby_(st, None, 0.1, good, check_normalization=True, coeffs=[])
    assert abs(st.norm() - 1) < 1e-12
    # spectral radius underestimated by 3x: the reference asserts "Incorrect normalization"
    bad = qp.ChebyWrk(st, gen, (ev[-1] - ev[0]) / 3, ev[0] / 3, 0.1)
    with pytest.raises(qp.QPropError) as exc:
        qp.cheby_(st, None, 0.1, bad, check_normalization=True, coeffs=[])
    assert exc.value.status == -5 and "Incorrect normalization" in str(exc.value)
    # the same check on a batch (sums flushed per row by the batched kernels), all trajectory counts per lane
    for B in (40, 100):
        w3 = qp.workloads.config3_transmon(n_sites=3, levels=3, B=B, nt=3, dt=0.5)
        gen3 = qp.DeviceGenerator(ctx, w3["ops"], 2)
        bound = float((abs(w3["ops"][0]) + 0.1 * abs(w3["ops"][1]) + 0.1 * abs(w3["ops"][2])).sum(axis=1).max())
        stB = qp.DeviceState.from_host(ctx, rand_state(rng, w3["ops"][0].shape[0], B))
        wB = qp.ChebyWrk(stB, gen3, 2 * bound, -bound, 0.5)
        qp.cheby_(stB, None, 0.5, wB, check_normalization=True, coeffs=[0.05, -0.02])
        assert np.max(np.abs(stB.norm() - 1)) < 1e-12
        wBad = qp.ChebyWrk(stB, gen3, 2 * bound / 40, -bound / 40, 0.5)
        with pytest.raises(qp.QPropError, match="Incorrect normalization"):
            qp.cheby_(stB, None, 0.5, wBad, check_normalization=True, coeffs=[0.05, -0.02])
    # wrong dt (src/cheby.jl:157)
    with pytest.raises(qp.QPropError) as exc:
        qp.cheby_(st, None, 0.2, good, coeffs=[])
    assert exc.value.status == -1 and "initialized for dt" in str(exc.value)
    # coefficient-count mismatch
    with pytest.raises(ValueError):
        qp.cheby_(st, None, 0.1, good, coeffs=[1.0])
    # dimension mismatch
    with pytest.raises(qp.QPropError):
        qp.ChebyWrk(qp.DeviceState(ctx, n + 1), gen, 1.0, 0.0, 0.1)


def test_cheby_dense_random_hermitian(qp, ctx):
    """test/test_cheby.jl:6-49 at N=400: dense Hermitian, spectral range from eigvals, vs exp."""
    rng = np.random.default_rng(6)
    N = 400
    X = rng.random((N, N)) + 1j * rng.random((N, N))
    H = (X + X.conj().T) / 2
    dt = 0.5
    psi0 = rand_state(rng, N)
    ev = np.linalg.eigvalsh(H)
    expected = sla.expm(-1j * H * dt) @ psi0
    st = qp.DeviceState.from_host(ctx, psi0)
    gen = qp.DeviceGenerator(ctx, [H], 0)
    assert gen.format == "dense"
    wrk = qp.ChebyWrk(st, gen, ev[-1] - ev[0], ev[0], dt)
    qp.cheby_(st, None, dt, wrk, coeffs=[])
    assert np.linalg.norm(st.to_host() - expected) < 1e-10


@pytest.mark.parametrize("N,B", [(64, 2), (203, 5), (203, 9), (130, 17), (301, 40), (97, 70), (256, 64)])
def test_dense_batched_dmma_mul(qp, ctx, N, B):
    """Dense generators on a batch of states (FP64 tensor-core kernel): mul! with two operators,
    shared coefficients; N and B not multiples of the tile sizes."""
    rng = np.random.default_rng(N * 7 + B)
    H0 = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    H1 = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    gen = qp.DeviceGenerator(ctx, [H0, H1], 1)
    assert gen.format == "dense"
    X, Y = rand_state(rng, N, B), rand_state(rng, N, B)
    dx = qp.DeviceState.from_host(ctx, X)
    for alpha, beta in ((1.0, 0.0), (0.3 - 2j, -1.5)):
        dy = qp.DeviceState.from_host(ctx, Y)
        gen.mul(dy, dx, [0.4 - 0.9j], alpha, beta)
        ref = beta * Y + alpha * ((H0 + (0.4 - 0.9j) * H1) @ X)
        assert rel(dy.to_host(), ref) < 1e-13


@pytest.mark.parametrize("B", [3, 16, 33])
def test_cheby_dense_batched_per_trajectory(qp, ctx, B):
    """BASELINE config 5 shape at reduced size: dense Hermitian H0 + u_b H1 on B states, each
    trajectory with its own control amplitude; every column against the oracle's cheby!."""
    rng = np.random.default_rng(60 + B)
    N = 150
    A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    C = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    H0, H1 = (A + A.conj().T) / 2, (C + C.conj().T) / 2
    u = rng.uniform(-1, 1, (1, B))
    bound = np.linalg.norm(H0, 2) + np.linalg.norm(H1, 2)
    Delta, E_min, dt = 2 * bound, -bound, 0.05
    psi0 = rand_state(rng, N, B)
    st = qp.DeviceState.from_host(ctx, psi0)
    gen = qp.DeviceGenerator(ctx, [H0, H1], 1)
    wrk = qp.ChebyWrk(st, gen, Delta, E_min, dt)
    owrk = O.ChebyWrk(psi0[:, 0].copy(), Delta, E_min, dt)
    ref = psi0.copy()
    for _ in range(3):
        qp.cheby_(st, None, dt, wrk, coeffs=u, per_trajectory=True, check_normalization=True)
        for b in range(B):
            col = ref[:, b].copy()
            O.cheby_inplace(col, O.Operator([H0, H1], [u[0, b]]), dt, owrk)
            ref[:, b] = col
    out = st.to_host()
    for b in range(B):
        assert rel(out[:, b], ref[:, b]) < RTOL
        assert abs(np.linalg.norm(out[:, b]) - 1) < 1e-12


# ---------------------------------------------------------------------------------------
# Newton / Arnoldi / specrange
# ---------------------------------------------------------------------------------------


def test_arnoldi_matches_oracle(qp, ctx):
    rng = np.random.default_rng(7)
    n, m = 500, 12
    A = rand_sparse(rng, n, 0.05, hermitian=False)
    v = rand_state(rng, n)
    Hess_ref = np.zeros((m + 1, m + 1), dtype=complex)
    q = [np.empty(n, dtype=complex) for _ in range(m + 1)]
    m_ref = O.arnoldi(Hess_ref, q, m, v, A, 0.3, extended=True)
    st = qp.DeviceState.from_host(ctx, v)
    K = qp.KrylovWrk(st, A, m)
    Hess = np.zeros((m + 1, m + 1), dtype=complex)
    m_gpu = qp.arnoldi_(Hess, K, m, st, A, 0.3, extended=True)
    assert m_gpu == m_ref == m
    assert np.linalg.norm(Hess - Hess_ref) < 1e-10 * np.linalg.norm(Hess_ref)
    tmp = st.similar()
    for i in range(m + 1):
        assert rel(K.get(i, tmp).to_host(), q[i]) < 1e-9
    # eigenvector start: Krylov dimension collapses to 1 (src/arnoldi.jl:91-95)
    Hh = rand_sparse(rng, 50, 0.3)
    w_, V = np.linalg.eigh(Hh.toarray())
    st2 = qp.DeviceState.from_host(ctx, V[:, 3])
    K2 = qp.KrylovWrk(st2, Hh, 5)
    Hess2 = np.zeros((6, 6), dtype=complex)
    assert qp.arnoldi_(Hess2, K2, 5, st2, Hh, 1.0, extended=True, norm_min=1e-10) == 1
    assert abs(Hess2[0, 0] - w_[3]) < 1e-12


def test_newton_optomech_vs_cheby_and_oracle(qp, ctx):
    """test/test_propagate.jl:153-163: ‖Ψ_newton‖−1 < 1e-12, ‖Ψ_newton − Ψ_cheby‖ < 1e-10."""
    H = qp.workloads.optomech()
    psi0 = qp.workloads.optomech_ket(0, 2)
    tlist = np.arange(0, 50 + 1e-9, 0.2)
    p1 = qp.propagate(psi0, (H,), tlist, "newton", ctx=ctx)
    p2 = qp.propagate(psi0, (H,), tlist, "cheby", ctx=ctx)
    assert abs(np.linalg.norm(p1) - 1.0) < 1e-12
    assert np.linalg.norm(p1 - p2) < 1e-10
    ref = O.propagate(psi0, (H,), tlist, "newton")
    assert rel(p1, ref) < RTOL


@pytest.mark.parametrize("hermitian,m_max", [(True, 5), (False, 50)])
def test_newton_random_vs_expm(qp, ctx, hermitian, m_max):
    """test/test_newton.jl:7-127 at N=300: vs dense exp(-i H dt), 1e-10."""
    rng = np.random.default_rng(8 + m_max)
    N = 300
    X = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    if hermitian:
        X = (X + X.conj().T) / 2
    X *= 10 / np.max(np.abs(np.linalg.eigvals(X)))
    psi0 = rand_state(rng, N)
    expected = sla.expm(-1j * X * 0.5) @ psi0
    st = qp.DeviceState.from_host(ctx, psi0)
    wrk = qp.NewtonWrk(st, X, m_max=m_max)
    qp.newton_(st, X, 0.5, wrk, max_restarts=200, coeffs=[])
    assert np.linalg.norm(st.to_host() - expected) < 1e-10
    ref = psi0.copy()
    O.newton_inplace(ref, X, 0.5, O.NewtonWrk(ref, m_max=m_max), max_restarts=200)
    assert rel(st.to_host(), ref) < RTOL


def test_newton_sparse_liouvillian_func_exp(qp, ctx):
    """test/test_newton.jl:130-177: sparse non-Hermitian super-operator, func = exp(L dt)."""
    rng = np.random.default_rng(12)
    N = 16
    Lm = rand_sparse(rng, N * N, 0.5, hermitian=False)
    Lm = (Lm * (10 / np.max(np.abs(np.linalg.eigvals(Lm.toarray()))))).tocsr()
    psi = rand_state(rng, N)
    rho0 = np.outer(psi, psi.conj()).reshape(-1)
    expected = sla.expm(Lm.toarray() * 0.5) @ rho0
    st = qp.DeviceState.from_host(ctx, rho0)
    wrk = qp.NewtonWrk(st, Lm, m_max=50)
    qp.newton_(st, Lm, 0.5, wrk, func=np.exp, max_restarts=20, coeffs=[])
    assert np.linalg.norm(st.to_host() - expected) < 1e-10


def test_newton_config4_small_vs_oracle(qp, ctx):
    """BASELINE config 4 shape at n_spins=4 (super-operator dimension 256), PWC control."""
    w = qp.workloads.config4_liouvillian(n_spins=4, nt=6, dt=0.05)
    ref = O.propagate(w["psi0"], _oracle_generator(w["ops"], w["controls"]), w["tlist"], "newton")
    out = qp.propagate(w["psi0"], _product_generator(qp, w["ops"], w["controls"]), w["tlist"], "newton", ctx=ctx)
    assert rel(out, ref) < RTOL
    rho = out.reshape(16, 16, order="F")
    assert abs(np.trace(rho) - 1) < 1e-10  # trace preserved by the Lindblad generator


def test_newton_eigenstate_shortcut_and_max_restarts(qp, ctx):
    rng = np.random.default_rng(13)
    H = rand_sparse(rng, 40, 0.3)
    w_, V = np.linalg.eigh(H.toarray())
    st = qp.DeviceState.from_host(ctx, V[:, 0])
    wrk = qp.NewtonWrk(st, H, m_max=5)
    qp.newton_(st, H, 0.7, wrk, coeffs=[], norm_min=1e-10)
    assert rel(st.to_host(), np.exp(-1j * w_[0] * 0.7) * V[:, 0]) < 1e-12
    assert wrk.restarts == 0
    big = rand_sparse(rng, 400, 0.2) * 200.0
    st = qp.DeviceState.from_host(ctx, rand_state(rng, 400))
    with pytest.raises(qp.QPropError) as exc:
        qp.newton_(st, big, 1.0, qp.NewtonWrk(st, big, m_max=3), max_restarts=2, coeffs=[])
    assert exc.value.status == -4
    with pytest.raises(RuntimeError, match="only implemented in-place"):
        qp.init_prop(V[:, 0], (H,), np.linspace(0, 1, 5), "newton", ctx=ctx, inplace=False)


@pytest.mark.parametrize("hermitian", [True, False])
def test_newton_exact_krylov_breakdown(qp, ctx, hermitian):
    """A basis-state start inside a small invariant block of a sparse generator: the Krylov space
    is exhausted after 3 vectors with a residue of norm EXACTLY zero (src/arnoldi.jl:92-95).  The
    restart vector of newton! combines that residue too (src/newton.jl:360-366), so it must stay
    finite (zero), not inf * 0; a large dt forces restarts."""
    rng = np.random.default_rng(77)
    N = 24
    H = np.zeros((N, N), dtype=complex)
    blk = np.array([[0.3, 1.0, 0.0], [1.0, -0.2, 0.7], [0.0, 0.7, 0.5]], dtype=complex)
    if not hermitian:
        blk = blk + np.array([[0, 0.4j, 0], [-0.1, 0.05j, 0.3], [0, 0.2, -0.1j]])
    H[:3, :3] = blk
    rest = rng.standard_normal((N - 3, N - 3)) + 1j * rng.standard_normal((N - 3, N - 3))
    H[3:, 3:] = rest + rest.conj().T
    Hs = sp.csr_matrix(H)
    psi = np.zeros(N, dtype=complex)
    psi[0] = 1.0
    dt = 6.0
    expected = sla.expm(-1j * H * dt) @ psi
    for host_step in ("library", "python"):
        st = qp.DeviceState.from_host(ctx, psi)
        wrk = qp.NewtonWrk(st, Hs, m_max=8)
        qp.newton_(st, Hs, dt, wrk, max_restarts=200, coeffs=[], host_step=host_step)
        out = st.to_host()
        assert np.all(np.isfinite(out))
        assert rel(out, expected) < RTOL
        assert wrk.restarts >= 1
    # the fine-grained Arnoldi reports the reduced dimension and finite vectors
    st = qp.DeviceState.from_host(ctx, psi)
    wrk = qp.NewtonWrk(st, Hs, m_max=8)
    Hess = np.zeros((9, 9), dtype=np.complex128)
    m = qp.arnoldi_(Hess, wrk.krylov, 8, st, Hs, 1.0, coeffs=[])
    assert m == 3 and abs(Hess[3, 2]) < 1e-14 and np.all(np.isfinite(Hess))
    q3 = wrk.krylov.get(3, st.similar()).to_host()
    assert np.all(np.isfinite(q3)) and np.linalg.norm(q3) < 1e-14


@pytest.mark.parametrize("hermitian", [True, False])
def test_state_bundle_arnoldi_and_newton(qp, ctx, hermitian):
    """a bundle of B states sharing one generator through the batched Krylov workspace:
    per-state Hessenberg matrices equal the single-state ones, Arnoldi vectors are orthonormal per
    state, and the batched newton! equals B single-state propagations and the dense exponential
    (different states converge after different numbers of restarts; one is an exact eigenvector and
    takes the shortcut of src/newton.jl:289-295)."""
    rng = np.random.default_rng(21 + hermitian)
    N, B, m = 60, 5, 6
    A = rand_sparse(rng, N, 0.15, hermitian=hermitian)
    A = (A * (4.0 / np.max(np.abs(np.linalg.eigvals(A.toarray()))))).tocsr()
    psi = rand_state(rng, N, B)
    psi[:, 1] *= 0.1                                     # different norms (the convergence test of newton! is absolute)
    if hermitian:
        ev, U = np.linalg.eigh(A.toarray())
        psi[:, 3] = U[:, 7]                              # an exact eigenvector
    st = qp.DeviceState.from_host(ctx, psi)
    wrk = qp.NewtonWrk(st, A, m_max=m)
    # Arnoldi, bundle vs one state at a time
    v = st.copy()
    nrm = v.norm()
    vh = psi / nrm
    v.upload(vh)
    Hess = np.zeros((B, m + 1, m + 1), dtype=np.complex128)
    m_out = qp.arnoldi_(Hess, wrk.krylov, m, v, A, 0.7, norm_min=1e-12, coeffs=[])
    if hermitian:
        assert m_out[3] == 1                             # the eigenvector exhausts its Krylov space at once
    for b in range(B):
        s1 = qp.DeviceState.from_host(ctx, vh[:, b].copy())
        w1 = qp.NewtonWrk(s1, A, m_max=m)
        H1 = np.zeros((m + 1, m + 1), dtype=np.complex128)
        m1 = qp.arnoldi_(H1, w1.krylov, m, s1, A, 0.7, norm_min=1e-12, coeffs=[])
        assert m1 == m_out[b]
        if m1 == m:
            assert np.max(np.abs(Hess[b] - H1)) < 1e-12
    Q = np.stack([wrk.krylov.get(j, st.similar()).to_host() for j in range(m + 1)])   # (m+1, N, B)
    for b in range(B):
        k = m_out[b] + (1 if m_out[b] == m else 0)
        G = Q[:k, :, b].conj() @ Q[:k, :, b].T
        assert np.max(np.abs(G - np.eye(k))) < 1e-12
    # newton! on the bundle
    dt = 0.9
    out = qp.newton_(st, A, dt, wrk, coeffs=[], max_restarts=200)
    got = out.to_host()
    want = sla.expm(-1j * dt * A.toarray()) @ psi
    for b in range(B):
        assert rel(got[:, b], want[:, b]) < RTOL
        s1 = qp.DeviceState.from_host(ctx, psi[:, b].copy())
        w1 = qp.NewtonWrk(s1, A, m_max=m)
        qp.newton_(s1, A, dt, w1, coeffs=[], max_restarts=200)
        assert rel(got[:, b], s1.to_host()) < 1e-11


def test_state_bundle_forward_store_backward_consume(qp, ctx):
    """The sweep pattern of GRAPE / Krotov (reference src/cheby_propagator.jl:147-152, 353-356;
    consumer pattern test/test_exputils.jl:148-172): a bundle of states is propagated forward with
    every time slot stored (device-resident copies), a bundle of co-states is propagated BACKWARD from
    the final time, and at every slot the stored forward states are consumed: the overlaps
    <chi_k(t_n)|psi_k(t_n)> must not depend on n (unitary evolution), for Chebyshev and for Newton."""
    rng = np.random.default_rng(77)
    w = qp.workloads.config2_tfim(6, nt=9, dt=0.1)
    N, B = w["psi0"].shape[0], 4
    H0, H1, H2 = w["ops"]
    tl = w["tlist"]
    mids = O.get_tlist_midpoints(tl)
    coeffs = [[w["controls"][0](t), w["controls"][1](t)] for t in mids]
    psi0, chiT = rand_state(rng, N, B), rand_state(rng, N, B)
    gen = qp.DeviceGenerator(ctx, [H0, H1, H2], 2)
    dt = tl[1] - tl[0]
    for method in ("cheby", "newton"):
        psi = qp.DeviceState.from_host(ctx, psi0)
        chi = qp.DeviceState.from_host(ctx, chiT)
        if method == "cheby":
            wrk_f = qp.ChebyWrk(psi, gen, 2.02 * w["E_max"], -1.01 * w["E_max"], dt)
            step = lambda s, c, sign: qp.cheby_(s, None, sign * dt, wrk_f, coeffs=c)                         # noqa: E731
        else:
            wrk_n = qp.NewtonWrk(psi, gen, m_max=8)
            step = lambda s, c, sign: qp.newton_(s, gen, sign * dt, wrk_n, coeffs=c)                         # noqa: E731
        storage = qp.init_storage(psi, tl)                 # list of nt slots (device-resident bundle copies)
        qp.write_to_storage(storage, 1, psi)
        for n in range(1, len(tl)):
            step(psi, coeffs[n - 1], +1)
            qp.write_to_storage(storage, n + 1, psi)
        fwd = qp.DeviceState(ctx, N, B)
        tau = [chi.dot(qp.get_from_storage_(fwd, storage, len(tl)))]
        for n in range(len(tl) - 1, 0, -1):               # backward sweep consuming the stored forward states
            step(chi, coeffs[n - 1], -1)
            tau.append(chi.dot(qp.get_from_storage_(fwd, storage, n)))
        tau = np.array(tau)                               # (nt, B)
        assert np.max(np.abs(tau - tau[0])) < 1e-10
        assert np.max(np.abs(tau[0] - np.einsum("nb,nb->b", chiT.conj(), psi.to_host()))) < 1e-12
        # and the backward-propagated co-states returned to chi(0) = U^+ chi(T): compare with the oracle per state
        U = np.eye(N, dtype=complex)
        for c in coeffs:
            U = sla.expm(-1j * dt * (H0 + c[0] * H1 + c[1] * H2).toarray()) @ U
        assert rel(chi.to_host(), U.conj().T @ chiT) < RTOL


def test_specrange_arnoldi_brackets_spectrum(qp, ctx):
    """test/test_specrad.jl:80-144: :arnoldi within 5% of Δ outside the true spectrum;
    :diag exact; :manual / :auto dispatch."""
    w = qp.workloads.config1_random(N=600, density=0.1, seed=77)
    H = w["ops"][0]
    ev = np.linalg.eigvalsh(H.toarray())
    D = ev[-1] - ev[0]
    E_min, E_max = qp.specrange(H, "arnoldi", ctx=ctx, prec=1e-4, rng=np.random.default_rng(1))
    assert ev[0] - 0.05 * D <= E_min <= ev[0]
    assert ev[-1] <= E_max < ev[-1] + 0.05 * D
    lo, hi = qp.specrange(H, "diag")
    assert abs(lo - ev[0]) < 1e-12 and abs(hi - ev[-1]) < 1e-12
    assert qp.specrange(H, E_min=-10, E_max=10) == (-10.0, 10.0)
    with pytest.raises(TypeError):
        qp.specrange(H, "manual", E_min=-1.0)
    # same start vector -> same Ritz values as the oracle
    psi = O.random_state(H, rng=np.random.default_rng(5))
    R_ref = O.ritzvals(H, psi, 25, 60, prec=1e-3)
    R = qp.ritzvals(H, psi, 25, 60, prec=1e-3, ctx=ctx)
    assert len(R) == len(R_ref)
    assert abs(R[0] - R_ref[0]) < 1e-8 and abs(R[-1] - R_ref[-1]) < 1e-8


def test_init_prop_spectral_arithmetic(qp, ctx):
    """test/test_specrad.jl:147-223."""
    w = qp.workloads.config1_random(N=200, density=0.1, seed=5, nt=21, T=1.0)
    gen = _product_generator(qp, w["ops"], w["controls"])
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, E_min=-10, E_max=10,
                     specrange_method="manual", specrange_buffer=0.1)
    assert abs(p.wrk.E_min + 11.0) < 1e-12 and abs(p.wrk.Delta - 22.0) < 1e-12
    u = w["controls"][0]
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, specrange_method="diag",
                     specrange_buffer=0.0, control_ranges=qp.IdDict([(u, (-1, 1))]))
    H0, H1 = w["ops"]
    ev_m, ev_p = np.linalg.eigvalsh((H0 - H1).toarray()), np.linalg.eigvalsh((H0 + H1).toarray())
    assert abs(p.wrk.E_min - min(ev_m[0], ev_p[0])) < 1e-10
    assert abs(p.wrk.Delta - (max(ev_m[-1], ev_p[-1]) - p.wrk.E_min)) < 1e-10
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, rng=np.random.default_rng(0))  # :arnoldi
    ev = [np.linalg.eigvalsh((H0 + s * H1).toarray()) for s in (-1, 1)]
    assert p.wrk.E_min <= min(e[0] for e in ev) and p.wrk.E_min + p.wrk.Delta >= max(e[-1] for e in ev)


# ---------------------------------------------------------------------------------------
# the propagator protocol (src/interfaces/propagator.jl:55-338, test/test_prop_interfaces.jl)
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("method", ["cheby", "newton"])
@pytest.mark.parametrize("backward", [False, True])
def test_propagator_protocol(qp, ctx, method, backward):
    w = qp.workloads.config1_random(N=10, density=0.5, seed=21, nt=101, T=5.0)
    gen = _product_generator(qp, w["ops"], w["controls"])
    kw = dict(E_min=w["E_min"], E_max=w["E_max"]) if method == "cheby" else {}
    st0 = qp.DeviceState.from_host(ctx, w["psi0"])
    p = qp.init_prop(st0, gen, w["tlist"], method, ctx=ctx, backward=backward, **kw)
    tl = w["tlist"]
    assert p.state is not st0  # in-place propagators work on a copy
    assert p.t == (tl[-1] if backward else tl[0])
    assert set(p.propertynames()) == {"state", "tlist", "t", "parameters", "backward", "inplace"}
    with pytest.raises(AttributeError):
        p.generator
    s1 = qp.prop_step(p)
    assert s1 is p.state
    assert p.t == (tl[-2] if backward else tl[1])
    assert abs(s1.norm() - 1) < 1e-12
    # parameters: control -> nt-1 values, mutable between steps
    u = w["controls"][0]
    assert len(p.parameters[u]) == len(tl) - 1
    # set_t! to the end: prop_step! returns nothing and leaves the propagator untouched
    qp.set_t(p, tl[0] if backward else tl[-1])
    before = p.state.to_host()
    assert qp.prop_step(p) is None
    assert np.array_equal(p.state.to_host(), before)
    # set_state! overwrites in place and returns the same object
    other = qp.DeviceState.from_host(ctx, np.roll(w["psi0"], 1))
    held = p.state
    assert qp.set_state(p, other) is held
    assert np.array_equal(held.to_host(), other.to_host())
    # reinit_prop! is idempotent
    qp.reinit_prop(p, st0)
    t0 = p.t
    a = qp.prop_step(p).to_host()
    qp.reinit_prop(p, st0)
    assert p.t == t0
    b = qp.prop_step(p).to_host()
    assert np.array_equal(a, b)
    with pytest.warns(UserWarning, match="Snapping"):
        qp.set_t(p, 0.5 * (tl[3] + tl[4]) + 1e-3)


def test_propagate_storage_and_parameters_mutation(qp, ctx):
    w = qp.workloads.config1_random(N=50, density=0.2, seed=31, nt=21, T=1.0)
    gen = _product_generator(qp, w["ops"], w["controls"])
    ogen = _oracle_generator(w["ops"], w["controls"])
    kw = dict(E_min=w["E_min"], E_max=w["E_max"])
    pops = [lambda s: float(abs(s.to_host()[0]) ** 2), lambda s: s.norm()]
    store = qp.propagate(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, storage=True, observables=pops, **kw)
    ref = O.propagate(w["psi0"], ogen, w["tlist"], "cheby", storage=True,
                      observables=[lambda s: float(abs(s[0]) ** 2), lambda s: np.linalg.norm(s)], **kw)
    assert store.shape == (2, 21) and np.allclose(store, ref, atol=1e-11)
    full = qp.propagate(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, storage=True, **kw)
    full_bw = qp.propagate(full[:, -1], gen, w["tlist"], "cheby", ctx=ctx, storage=True, backward=True, **kw)
    assert np.linalg.norm(full - full_bw) < 1e-10  # stored back to front, same trajectory
    # the caller may rewrite propagator.parameters between steps (src/propagator.jl:100-104)
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, control_ranges=qp.IdDict([(w["controls"][0], (-1, 1))]), **kw)
    po = O.init_prop(w["psi0"], ogen, w["tlist"], "cheby", control_ranges=OIdDict([(w["controls"][0], (-1, 1))]), **kw)
    u = w["controls"][0]
    p.parameters[u][:] = 0.25
    po.parameters[u][:] = 0.25
    for _ in range(5):
        qp.prop_step(p)
        O.prop_step(po)
    assert rel(p.state.to_host(), po.state) < RTOL


def test_reinit_prop_recomputes_coefficients(qp, ctx):
    """src/cheby_propagator.jl:243-299: larger amplitudes than the stored ranges -> new
    coefficients, re-uploaded without touching the operators."""
    w = qp.workloads.config1_random(N=80, density=0.2, seed=41, nt=11, T=1.0)
    gen = _product_generator(qp, w["ops"], w["controls"])
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, specrange_method="diag")
    n0, dev0 = p.wrk.n_coeffs, p.wrk.gen
    u = w["controls"][0]
    p.parameters[u] *= 5.0
    qp.reinit_prop(p, p.state)
    assert p.wrk.gen is dev0 and p.wrk.n_coeffs > n0
    ogen = _oracle_generator(w["ops"], [5.0 * qp.discretize_on_midpoints(u, w["tlist"])])
    ref = O.propagate(w["psi0"], ogen, w["tlist"], "cheby", specrange_method="diag")
    for _ in range(10):
        qp.prop_step(p)
    assert rel(p.state.to_host(), ref) < 1e-9  # different (both valid) spectral envelopes


def test_timings_labels(qp):
    """test/test_timings.jl:8-39: > 200 "matrix-vector product" calls for 100 steps."""
    c = qp.Context(0)
    c.enable_timings()
    w = qp.workloads.config1_random(N=100, density=0.2, seed=51, nt=101, T=5.0)
    gen = _product_generator(qp, w["ops"], w["controls"])
    qp.propagate(w["psi0"], gen, w["tlist"], "cheby", ctx=c, E_min=w["E_min"], E_max=w["E_max"])
    n_step, t_step = c.timing("prop_step!")
    n_mv, t_mv = c.timing("matrix-vector product")
    assert n_step == 100 and n_mv > 200 and 0 < t_mv <= t_step
    # Newton: the labels of src/newton.jl:276-343 (test/test_timings.jl:8-39), one call of each host
    # section per restart, and the NewtonWrk bookkeeping of src/newton.jl:381-383
    c.reset_timings()
    p = qp.init_prop(w["psi0"], gen, w["tlist"][:11], "newton", ctx=c, m_max=6)
    restarts = 0
    while qp.prop_step(p) is not None:
        restarts += p.wrk.restarts + 1
    counts = {lab: c.timing(lab)[0] for lab in ("arnoldi!", "diagonalize_hessenberg_matrix", "get Leja points",
                                                "get Newton coeffs", "evaluate polynomial", "matrix-vector product")}
    assert counts["arnoldi!"] == restarts == counts["diagonalize_hessenberg_matrix"] == counts["get Leja points"]
    assert counts["get Newton coeffs"] == restarts == counts["evaluate polynomial"]
    assert counts["matrix-vector product"] == 6 * restarts
    assert all(c.timing(lab)[1] > 0 for lab in counts)
    assert p.wrk.n_a == p.wrk.n_leja == 6 * (p.wrk.restarts + 1) and p.wrk.radius > 0
    assert np.all(p.wrk.a[: p.wrk.n_a] != 0) and np.all(np.abs(p.wrk.leja[: p.wrk.n_leja]) <= p.wrk.radius)


# ---------------------------------------------------------------------------------------
# on-device observables and the one-call propagation loop (src/propagate.jl:283-344,
# src/storage.jl:100-123)
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("fmt,n,B", [("csr", 300, 1), ("sell", 4100, 1), ("selld", 4100, 1), ("auto", 500, 5), ("auto", 500, 40)])
def test_expval_fused(qp, ctx, fmt, n, B):
    rng = np.random.default_rng(n + B)
    ops = _structured_ops(rng, n, 7, [[0, -3, 1], [5, -5, 40]])
    coeffs = [0.3 - 0.4j]
    gen = qp.DeviceGenerator(ctx, ops, 1, fmt)
    X = rand_state(rng, n, None if B == 1 else B)
    dx = qp.DeviceState.from_host(ctx, X)
    H = (ops[0] + coeffs[0] * ops[1]).tocsr()
    ref = np.vdot(X, H @ X) if B == 1 else np.einsum("ib,ib->b", X.conj(), H @ X)
    assert np.allclose(gen.expval(dx, coeffs), ref, rtol=1e-12, atol=1e-13)
    np.testing.assert_array_equal(dx.to_host(), X)  # the state is untouched


@pytest.mark.parametrize("fmt,n_spins,B", [("selld", 14, 1), ("sell", 14, 1), ("csr", 12, 1), ("auto", 8, 64), ("bitflip", 14, 1)])
def test_expval_bitwise_reproducible(qp, ctx, fmt, n_spins, B):
    """The fused expectation value adds per-warp (single states) / per-tile (tiled batched kernel)
    partial sums in a fixed order: repeated calls agree to the last bit (no atomics)."""
    rng = np.random.default_rng(n_spins)
    H0, H1, H2 = qp.workloads.tfim_chain(n_spins)
    N = H0.shape[0]
    gen = qp.DeviceGenerator(ctx, [H0, H1, H2], 2, fmt)
    x = qp.DeviceState.from_host(ctx, rand_state(rng, N, None if B == 1 else B))
    vals = [np.atleast_1d(gen.expval(x, [0.3, -0.7])) for _ in range(6)]
    for v in vals[1:]:
        assert np.array_equal(v.view(np.float64), vals[0].view(np.float64))
    X = x.to_host().reshape(N, -1)
    want = np.einsum("nb,nb->b", X.conj(), (H0 + 0.3 * H1 - 0.7 * H2) @ X)
    assert np.max(np.abs(vals[0] - want)) < 1e-11


def test_expval_dense(qp, ctx):
    rng = np.random.default_rng(21)
    n = 130
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    for B in (1, 9):
        X = rand_state(rng, n, None if B == 1 else B)
        dx = qp.DeviceState.from_host(ctx, X)
        ref = np.vdot(X, A @ X) if B == 1 else np.einsum("ib,ib->b", X.conj(), A @ X)
        assert np.allclose(qp.DeviceGenerator(ctx, [A], 0).expval(dx), ref, rtol=1e-12)


def test_propagate_storage_follows_storage_module(qp, ctx):
    """propagate routes its storage through init_storage / map_observables / write_to_storage
    (reference src/propagate.jl:283-351): observables of (state, tlist, i), a single observable
    storing its bare value (length-nt vector), list storage for device-resident states, and a
    caller-supplied storage."""
    w = qp.workloads.config2_tfim(8, nt=11, dt=0.1)
    gen = _product_generator(qp, w["ops"], w["controls"])
    kw = dict(E_min=w["E_min"], E_max=w["E_max"], ctx=ctx)
    tl = w["tlist"]
    # 3-argument observable
    seen = []
    def obs3(state, tlist, i):
        seen.append(i)
        return tlist[i - 1] * state.norm()
    vec = qp.propagate(w["psi0"], gen, tl, "cheby", storage=True, observables=(obs3,), **kw)
    assert vec.shape == (11,) and seen[1:] == list(range(1, 12)) and np.allclose(vec, tl, atol=1e-11)
    # single matrix observable: bare expectation values, fast path == loop path
    Oz = w["ops"][2]
    fast = qp.propagate(w["psi0"], gen, tl, "cheby", storage=True, observables=(Oz,), **kw)
    slow = qp.propagate(w["psi0"], gen, tl, "cheby", storage=True, observables=(Oz,), callback=lambda p, o: None, **kw)
    assert fast.shape == (11,) and slow.shape == (11,) and np.max(np.abs(fast - slow)) < 1e-11
    ref = O.propagate(w["psi0"], _oracle_generator(w["ops"], w["controls"]), tl, "cheby", storage=True,
                      observables=(lambda psi: np.vdot(psi, Oz @ psi),), E_min=w["E_min"], E_max=w["E_max"])
    assert np.max(np.abs(fast - ref[0])) < 1e-10
    # device-resident initial state: default storage is a list of per-slot device copies
    st0 = qp.DeviceState.from_host(ctx, w["psi0"])
    slots = qp.propagate(st0, gen, tl, "cheby", storage=True, **kw)
    assert isinstance(slots, list) and len(slots) == 11 and all(isinstance(s, qp.DeviceState) for s in slots)
    full = qp.propagate(w["psi0"], gen, tl, "cheby", storage=True, **kw)
    assert full.shape == (w["psi0"].shape[0], 11)
    assert rel(slots[-1].to_host(), full[:, -1]) < 1e-14 and rel(slots[0].to_host(), w["psi0"]) < 1e-15
    # caller-supplied list storage with matrix observables on the one-call path
    mine = [None] * 11
    qp.propagate(w["psi0"], gen, tl, "cheby", storage=mine, observables=(Oz, Oz), **kw)
    assert np.allclose([m[0] for m in mine], fast, atol=1e-11)


@pytest.mark.parametrize("backward", [False, True])
def test_propagate_one_call_with_observables(qp, ctx, backward):
    """propagate(...; storage=true, observables=(O1, O2)) with matrix observables runs as ONE
    library call with the expectation values recorded on the device; it must agree with the
    oracle's step loop and with this package's own per-step loop."""
    w = qp.workloads.config2_tfim(10, nt=21, dt=0.1)
    N = w["psi0"].shape[0]
    O1 = w["ops"][2]                                   # sum Z_i (diagonal)
    O2 = (w["ops"][1] + 0.5j * sp.eye(N)).tocsr()      # sum X_i + 0.5i: complex expectation values
    kw = dict(E_min=w["E_min"], E_max=w["E_max"], backward=backward)
    obs_host = (lambda psi: np.vdot(psi, O1 @ psi), lambda psi: np.vdot(psi, O2 @ psi))
    ref = O.propagate(w["psi0"], _oracle_generator(w["ops"], w["controls"]), w["tlist"], "cheby", storage=True,
                      observables=obs_host, **kw)
    gen = _product_generator(qp, w["ops"], w["controls"])
    launches0 = ctx.launch_count
    fast = qp.propagate(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, storage=True, observables=(O1, O2), **kw)
    assert fast.shape == (2, 21)
    assert np.max(np.abs(fast - ref)) < 1e-10 * N ** 0.5
    # the slow loop (a callback forces it) gives the same numbers
    slow = qp.propagate(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, storage=True, observables=(O1, O2),
                        callback=lambda p, obs: None, **kw)
    assert np.max(np.abs(fast - slow)) < 1e-11
    # and the final state of the one-call path equals the oracle's
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, **kw)
    out = qp.propagate(p)
    ref_state = O.propagate(w["psi0"], _oracle_generator(w["ops"], w["controls"]), w["tlist"], "cheby", **kw)
    assert rel(out, ref_state) < RTOL
    assert p.t == (w["tlist"][0] if backward else w["tlist"][-1])
    assert qp.prop_step(p) is None  # grid exhausted, like after the step loop
    assert ctx.launch_count > launches0


def test_cheby_propagate_norms_batched(qp, ctx):
    """qp_cheby_propagate on a batch: per-trajectory coefficient table, norms recorded on device."""
    rng = np.random.default_rng(31)
    w = qp.workloads.config3_transmon(n_sites=3, levels=3, B=20, nt=6, dt=0.5)
    H0, H1, H2 = w["ops"]
    B, n_steps = 20, 5
    psi0 = rand_state(rng, H0.shape[0], B)
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
    st = qp.DeviceState.from_host(ctx, psi0)
    gen = qp.DeviceGenerator(ctx, [H0, H1, H2], 2)
    wrk = qp.ChebyWrk(st, gen, 2 * bound, -bound, 0.5)
    table = rng.uniform(-0.07, 0.07, (n_steps, 2, B)).astype(complex)
    obs = qp.DeviceGenerator(ctx, [H0], 0)
    ev, norms = qp.cheby_propagate_(st, wrk, table, 0.5, observables=[obs], norms=True, per_trajectory=True)
    assert ev.shape == (n_steps + 1, 1, B) and norms.shape == (n_steps + 1, B)
    assert np.max(np.abs(norms - 1)) < 1e-12
    st2 = qp.DeviceState.from_host(ctx, psi0)
    for s in range(n_steps):
        assert np.allclose(ev[s, 0], np.einsum("ib,ib->b", st2.to_host().conj(), H0 @ st2.to_host()), atol=1e-12)
        qp.cheby_(st2, None, 0.5, wrk, coeffs=table[s], per_trajectory=True)
    assert rel(st.to_host(), st2.to_host()) < 1e-14


@pytest.mark.parametrize("hermitian", [True, False])
def test_newton_library_vs_python_host_step(qp, ctx, hermitian):
    """qp_newton_step (dense step in C++ inside the library) against the fine-grained path
    (qp_arnoldi + qp_krylov_combine with the dense step in this package's Python mirror) and
    against dense exp; also a func given as a callback."""
    rng = np.random.default_rng(90 + hermitian)
    n = 300
    A = rand_sparse(rng, n, 0.05, hermitian=hermitian)
    if not hermitian:
        A = A - 0.2j * sp.identity(n)  # dissipative
    psi = rand_state(rng, n)
    dt = 0.3
    expected = sla.expm(-1j * dt * A.toarray()) @ psi
    outs = {}
    for mode in ("library", "python"):
        st = qp.DeviceState.from_host(ctx, psi)
        wrk = qp.NewtonWrk(st, A, m_max=10)
        qp.newton_(st, A, dt, wrk, coeffs=[], host_step=mode)
        outs[mode] = (st.to_host(), wrk.restarts)
        assert rel(outs[mode][0], expected) < RTOL
    assert outs["library"][1] == outs["python"][1]  # same number of restarts
    assert rel(outs["library"][0], outs["python"][0]) < 1e-12
    # arbitrary func through the callback: exp(-i z) written by hand must give the same answer
    st = qp.DeviceState.from_host(ctx, psi)
    wrk = qp.NewtonWrk(st, A, m_max=10)
    qp.newton_(st, A, dt, wrk, coeffs=[], func=lambda z: np.cos(z) - 1j * np.sin(z))
    assert rel(st.to_host(), expected) < RTOL


# ---------------------------------------------------------------------------------------
# edge cases: empty operators and rows, 1 x 1 systems, slice boundaries, every storage format
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("fmt", ["csr", "sell", "selld", "auto"])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 95])
def test_edge_shapes_all_formats(qp, ctx, fmt, n):
    """N around the slice height (32), a 1 x 1 system, an all-zero controlled operator, rows that
    are empty in every operator, for every sparse storage format and for B = 1 / 3 / 40."""
    rng = np.random.default_rng(100 + n)
    diag = sp.diags(np.arange(1, n + 1, dtype=complex) * 0.1, 0, format="csr")
    band = sp.diags([np.full(max(n - 1, 0), 0.5 + 0.25j)], [1], shape=(n, n), format="lil")
    if n > 4:
        band[n // 2, :] = 0  # an empty row in the coupling operator
    band = sp.csr_matrix(band)
    H1 = (band + band.conj().T).tocsr()
    zero = sp.csr_matrix((n, n), dtype=complex)
    ops = [diag, H1, zero]
    coeffs = [0.3 - 0.2j, 5.0]
    gen = qp.DeviceGenerator(ctx, ops, 2, fmt)
    dense = (diag + coeffs[0] * H1).toarray()
    for B in (None, 3, 40):
        X = rand_state(rng, n, B)
        Y = rand_state(rng, n, B)
        dx, dy = qp.DeviceState.from_host(ctx, X), qp.DeviceState.from_host(ctx, Y)
        gen.mul(dy, dx, coeffs, 0.5 - 1j, 2.0)
        ref = 2.0 * Y + (0.5 - 1j) * (dense @ X)
        assert np.linalg.norm(dy.to_host() - ref) <= 1e-13 * max(1.0, np.linalg.norm(ref))
        ev = np.atleast_1d(gen.expval(dx, coeffs))
        want = np.atleast_1d(np.einsum("i...,i...->...", np.conj(X), dense @ X))
        assert np.max(np.abs(ev - want)) < 1e-12
    # a generator made of zero operators only: mul! gives beta * y, Chebyshev gives a pure phase
    gz = qp.DeviceGenerator(ctx, [zero], 0, fmt if fmt != "selld" else "auto")
    x = rand_state(rng, n)
    dx, dy = qp.DeviceState.from_host(ctx, x), qp.DeviceState.from_host(ctx, x)
    gz.mul(dy, dx, [], 1.0, 0.0)
    assert np.linalg.norm(dy.to_host()) == 0.0
    wrk = qp.ChebyWrk(dx, gz, 2.0, -1.0, 0.1)
    qp.cheby_(dx, None, 0.1, wrk, coeffs=[])
    assert np.linalg.norm(dx.to_host() - x) < 1e-13


@pytest.mark.parametrize("row_len", [2, 4, 6, 10, 12, 18, 20, 22, 23])
@pytest.mark.parametrize("n_vals", [3, 60])
def test_selld_uniform_width_tails_and_real_table(qp, ctx, row_len, n_vals):
    """Matrices whose rows all have the same length (circulant band): the SELL-D kernel is
    specialised at compile time on the number of codes in the last code word (TAIL = 2 / 4 / 6,
    or the generic path) for 8-bit (n_vals = 3) and 16-bit (n_vals = 60, longer rows) codes, and on whether
    every coefficient x value product is real (real operator x real coefficient, imaginary operator
    x imaginary coefficient) or not."""
    rng = np.random.default_rng(row_len * 1000 + n_vals)
    n = 4096
    table = rng.standard_normal(n_vals)
    r = np.arange(n)

    def circulant(offsets, phase):
        rows = np.concatenate([r for _ in offsets])
        cols = np.concatenate([(r + d) % n for d in offsets])
        vals = np.concatenate([table[(r * 7 + abs(d)) % n_vals] for d in offsets]) * phase
        A = sp.csr_matrix((vals.astype(complex), (rows, cols)), shape=(n, n))
        A.sort_indices()
        return A

    offs = [0, 1, -1, 5, -5, 33, -33, 64, -64, 200, -200, 777, -777, 1024, -1024, 1500, -1500, 2, -2, 9, -9, 17, -17]
    n_real = row_len // 2
    A_real = circulant(offs[:n_real], 1.0)                     # purely real operator
    A_imag = circulant(offs[n_real:row_len], 1j)               # purely imaginary operator
    gen = qp.DeviceGenerator(ctx, [A_real, A_imag], 1, "selld")
    assert gen.format == "selld" and gen.code_bytes == (2 if gen.n_dict > 256 else 1)
    if n_vals == 60 and row_len >= 6:
        assert gen.code_bytes == 2
    x, y0 = rand_state(rng, n), rand_state(rng, n)
    dx = qp.DeviceState.from_host(ctx, x)
    # imaginary coefficient on the imaginary operator: all products real; complex / real: not
    for c in (-0.7j, 0.3 - 0.4j, 1.5):
        dense = (A_real + c * A_imag).tocsr()
        for alpha, beta in ((1.0, 0.0), (0.5 - 1j, 2.0)):
            dy = qp.DeviceState.from_host(ctx, y0)
            gen.mul(dy, dx, [c], alpha, beta)
            assert rel(dy.to_host(), beta * y0 + alpha * (dense @ x)) < 1e-13
        assert abs(gen.expval(dx, [c]) - np.vdot(x, dense @ x)) < 1e-12 * n


def _bitflip_ops(rng, n_bits, masks, complex_values, complex_diag, few_values=False):
    """Diagonal operators + one operator that couples every row r to r ^ m with the same value."""
    N = 1 << n_bits
    rows = np.arange(N)
    d0 = rng.standard_normal(N) + (1j * rng.standard_normal(N) if complex_diag else 0)
    if few_values:   # both diagonals take a handful of values: stored as one 16-bit code per row
        d0 = rng.choice(np.array([-2.5, -0.0, 0.0, 0.75, 3.0]), N)
    d1 = rng.integers(-3, 4, N).astype(float)
    vals = rng.standard_normal(len(masks)) + (1j * rng.standard_normal(len(masks)) if complex_values else 0)
    X = sum(sp.csr_matrix((np.full(N, v, dtype=complex), (rows, rows ^ m)), shape=(N, N)) for m, v in zip(masks, vals))
    return [sp.diags(d0.astype(complex)).tocsr(), X.tocsr(), sp.diags(d1.astype(complex)).tocsr()]


@pytest.mark.parametrize("n_bits,masks,cv,cd,coeffs", [
    (6, [1, 2, 4, 8, 16, 32], False, False, [0.7, -1.3]),                    # real everywhere: constant-bank products
    (11, [1 << i for i in range(11)], False, False, [0.7, -1.3]),             # transverse field, 11 spins
    (11, [3, 5, 48, 1025, 2047, 7, 640, 96, 31, 33], True, False, [0.7, -1.3]),   # multi-bit flips, complex values
    (10, [1 << i for i in range(10)], False, True, [0.7 - 0.2j, 0.4j]),       # complex diagonal, complex coefficients
    (12, [1 << i for i in range(12)] + [3 << i for i in range(11)], False, False, [1.1, 0.3]),  # 23 terms: tail batches
    (9, [64, 128, 256], False, False, [0.5, 2.0]),                            # no in-warp masks at all
    (11, [1 << i for i in range(11)], False, "few", [0.7, -1.3]),             # coded diagonals, real coefficients
    (10, [1, 2, 512, 77], True, "few", [0.7 - 0.2j, 0.4j]),                   # coded diagonals, complex coefficients
])
def test_operator_mul_bitflip(qp, ctx, n_bits, masks, cv, cd, coeffs):
    """QP_FORMAT_BITFLIP (diagonal vectors + XOR stencil, no matrix stream): 5-argument mul! and the fused
    expectation value against scipy, real and complex coefficient products, masks below and above the warp
    width, term counts around the batch size of 8."""
    rng = np.random.default_rng(n_bits * 100 + len(masks))
    ops = _bitflip_ops(rng, n_bits, masks, cv, cd is True, few_values=cd == "few")
    N = 1 << n_bits
    gen = qp.DeviceGenerator(ctx, ops, 2, "bitflip")
    assert gen.format == "bitflip"
    # real diagonals with few distinct values (<= 256 each, <= 2048 jointly) are stored as one 16-bit code per row
    n0, n2 = (len(np.unique(op.diagonal())) for op in (ops[0], ops[2]))
    coded = cd is not True and max(n0, n2) <= 256 and n0 * n2 <= 2048
    assert gen.stored_bytes == (2 * N if coded else 24 * N if cd is True else 16 * N)
    H = (ops[0] + coeffs[0] * ops[1] + coeffs[1] * ops[2]).tocsr()
    x = rand_state(rng, N)
    y0 = rand_state(rng, N)
    dx = qp.DeviceState.from_host(ctx, x)
    for alpha, beta in [(1.0, 0.0), (0.3 - 0.8j, 0.0), (-0.5j, 1.0), (0.25, -0.6 + 0.1j)]:
        dy = qp.DeviceState.from_host(ctx, y0)
        gen.mul(dy, dx, coeffs, alpha, beta)
        assert rel(dy.to_host(), alpha * (H @ x) + beta * y0) < 1e-13
    assert abs(gen.expval(dx, coeffs) - np.vdot(x, H @ x)) < 1e-11
    # the batched kernels of the same generator still work (tiled / dictionary / CSR forms)
    X = rand_state(rng, N, 5)
    dX = qp.DeviceState.from_host(ctx, X)
    dY = qp.DeviceState(ctx, N, 5).zero()
    gen.mul(dY, dX, coeffs, 1.0, 0.0)
    assert rel(dY.to_host(), H @ X) < 1e-13


def test_bitflip_refuses_other_structures(qp, ctx):
    """Row-dependent values (a Y-type flip: the sign follows the bit), rows that carry an entry without forming a
    sub-cube, or generic sparse matrices are not bit-flip operators."""
    N = 256
    rows = np.arange(N)
    sign = 1 - 2 * ((rows >> 3) & 1)
    Y = sp.csr_matrix((1j * sign.astype(complex), (rows, rows ^ 8)), shape=(N, N))
    with pytest.raises(qp.QPropError):
        qp.DeviceGenerator(ctx, [Y], 0, "bitflip")
    rng = np.random.default_rng(0)
    A = sp.random(N, N, 0.05, random_state=1, format="csr").astype(complex)
    with pytest.raises(qp.QPropError):
        qp.DeviceGenerator(ctx, [A], 0, "bitflip")
    r3 = rows[rows % 3 == 0]                       # the rows carrying the flip are not a sub-cube
    C3 = sp.csr_matrix((np.ones(len(r3), dtype=complex), (r3, r3 ^ 16)), shape=(N, N))
    with pytest.raises(qp.QPropError):
        qp.DeviceGenerator(ctx, [C3], 0, "bitflip")
    # AUTO falls back silently and stays correct
    gen = qp.DeviceGenerator(ctx, [Y], 0)
    assert gen.format != "bitflip"
    x = rand_state(rng, N)
    dy = qp.DeviceState(ctx, N).zero()
    gen.mul(dy, qp.DeviceState.from_host(ctx, x), [], 1.0, 0.0)
    assert rel(dy.to_host(), Y @ x) < 1e-14


@pytest.mark.parametrize("n_bits,coeffs", [(8, [0.7, -1.3]), (11, [0.4 - 0.3j, 1.1j])])
def test_operator_mul_bitflip_conditional_terms(qp, ctx, n_bits, coeffs):
    """Conditional flips (sigma^-/sigma^+ on a bit: present on the rows whose bit is 0 / 1), a two-bit conditional
    flip (the jump term of a decay channel in a Liouvillian), diagonal and flips inside ONE operator."""
    rng = np.random.default_rng(n_bits)
    N = 1 << n_bits
    rows = np.arange(N)

    def lower(k, g):      # |..0..><..1..|: row has bit k = 0, column = row | 2^k
        r = rows[(rows >> k) & 1 == 0]
        return sp.csr_matrix((np.full(len(r), g, dtype=complex), (r, r | (1 << k))), shape=(N, N))

    def flip(m, g):
        return sp.csr_matrix((np.full(N, g, dtype=complex), (rows, rows ^ m)), shape=(N, N))

    def jump(k1, k2, g):  # both bits 1 -> 0
        r = rows[((rows >> k1) & 1 == 0) & ((rows >> k2) & 1 == 0)]
        return sp.csr_matrix((np.full(len(r), g, dtype=complex), (r, r | (1 << k1) | (1 << k2))), shape=(N, N))

    d = sp.diags((rng.standard_normal(N) + 1j * rng.standard_normal(N))).tocsr()
    op0 = (d + jump(1, n_bits - 1, 0.05) + jump(6, 3, 0.07) + lower(2, 0.3 - 0.1j) + lower(n_bits - 2, 0.2).T).tocsr()
    op1 = sum(flip(1 << k, -1j) for k in range(n_bits // 2)) + sum(flip(1 << k, 1j) for k in range(n_bits // 2, n_bits))
    op2 = (sp.diags(rng.integers(-2, 3, N).astype(complex)) + lower(0, 1.5) + lower(0, 1.5).T.conj()).tocsr()
    ops = [op0, op1.tocsr(), op2]
    gen = qp.DeviceGenerator(ctx, ops, 2, "bitflip")
    assert gen.format == "bitflip"
    H = (ops[0] + coeffs[0] * ops[1] + coeffs[1] * ops[2]).tocsr()
    x = rand_state(rng, N)
    y0 = rand_state(rng, N)
    dx = qp.DeviceState.from_host(ctx, x)
    for alpha, beta in [(1.0, 0.0), (0.3 - 0.8j, 1.0)]:
        dy = qp.DeviceState.from_host(ctx, y0)
        gen.mul(dy, dx, coeffs, alpha, beta)
        assert rel(dy.to_host(), alpha * (H @ x) + beta * y0) < 1e-13
    assert abs(gen.expval(dx, coeffs) - np.vdot(x, H @ x)) < 1e-11


@pytest.mark.parametrize("n_spins", [4, 6])
def test_newton_liouvillian_bitflip_vs_csr(qp, ctx, n_spins):
    """BASELINE config 4 shape at reduced size: the Liouvillian (commutator halves = flips, decay channels =
    conditional two-bit flips, complex diagonal) on the bit-flip form against merged CSR and the oracle."""
    w = qp.workloads.config4_liouvillian(n_spins, nt=6, dt=0.05)
    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    ref = O.propagate(w["psi0"], O.hamiltonian(*terms), w["tlist"], "newton")
    outs = {}
    for fmt in ("csr", "bitflip"):
        p = qp.init_prop(w["psi0"], qp.hamiltonian(*terms), w["tlist"], "newton", ctx=ctx, matrix_format=fmt)
        assert p.wrk.krylov.gen.format == fmt
        outs[fmt] = qp.propagate(p)
        assert rel(outs[fmt], ref) < RTOL, fmt
    assert rel(outs["bitflip"], outs["csr"]) < 1e-12


@pytest.mark.parametrize("N,masks", [
    (96, [1, 2, 7, 16, 31]),                                  # N not a power of two: only in-warp partners keep r ^ m < N
    (1 << 12, list(range(1, 41))),                            # 40 terms: nine+ shuffle candidates, more load terms than static slots
    (1 << 13, [1 << k for k in range(13)] + [(1 << k) | 1 for k in range(1, 13)] + [4095, 8191, 5461]),
])
def test_operator_mul_bitflip_term_counts(qp, ctx, N, masks):
    """Edge shapes of the bit-flip form: N not a power of two, more in-warp terms than shuffle slots (the rest
    become loads), more load terms than compile-time constant-bank positions (the dynamic tail loop)."""
    rng = np.random.default_rng(N + len(masks))
    rows = np.arange(N)
    vals = rng.standard_normal(len(masks)) + 1j * rng.standard_normal(len(masks))
    X = sum(sp.csr_matrix((np.full(N, v, dtype=complex), (rows, rows ^ m)), shape=(N, N)) for m, v in zip(masks, vals)).tocsr()
    D = sp.diags(rng.standard_normal(N).astype(complex)).tocsr()
    gen = qp.DeviceGenerator(ctx, [D, X], 1, "bitflip")
    assert gen.format == "bitflip"
    x = rand_state(rng, N)
    dx = qp.DeviceState.from_host(ctx, x)
    for c in (0.8, 0.3 - 1.1j):
        dy = qp.DeviceState(ctx, N).zero()
        gen.mul(dy, dx, [c], 1.0, 0.0)
        assert rel(dy.to_host(), (D + c * X) @ x) < 1e-13


def test_dense_gemv_cta_per_row(qp, ctx):
    """Dense generators on single states with N >= 2048 use the CTA-per-row GEMV (grid-stride over the rows,
    partial sums through shared memory): mul!, fused expectation value (bit-reproducible) and one Chebyshev
    step against NumPy / the oracle."""
    rng = np.random.default_rng(2048)
    n = 2048 + 128   # a multiple of 128: the flat two-pass GEMV; QPROP_GEMV_FLAT=0: CTA per row
    A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n)
    Bm = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n)
    ctx = qp.Context(0)   # a fresh context: the first call below has to grow every scratch buffer
    gen = qp.DeviceGenerator(ctx, [A, Bm], 1)
    assert gen.format == "dense"
    x = rand_state(rng, n)
    y0 = rand_state(rng, n)
    dx = qp.DeviceState.from_host(ctx, x)
    H = A + (0.3 - 0.2j) * Bm
    assert abs(gen.expval(dx, [0.3 - 0.2j]) - np.vdot(x, H @ x)) < 1e-11
    for alpha, beta in [(1.0, 0.0), (0.5j, 1.0)]:
        dy = qp.DeviceState.from_host(ctx, y0)
        gen.mul(dy, dx, [0.3 - 0.2j], alpha, beta)
        assert rel(dy.to_host(), alpha * (H @ x) + beta * y0) < 1e-13
    vals = [gen.expval(dx, [0.3 - 0.2j]) for _ in range(3)]
    assert abs(vals[0] - np.vdot(x, H @ x)) < 1e-11 and vals[0] == vals[1] == vals[2]
    Hh = (A + A.conj().T) / 2
    tlist = np.array([0.0, 0.05])
    out = qp.propagate(x, (Hh,), tlist, "cheby", ctx=ctx)
    ref = O.propagate(x, (Hh,), tlist, "cheby")
    assert rel(out, ref) < RTOL


def test_bitflip_device_coefficient_path(qp, ctx):
    """Launches whose coefficients only exist on the device take the products from shared memory (CK = 0): the
    same bit-flip tests in a child process with that path forced (the switch is read once per process)."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, QPROP_BITFLIP_DEVICE_COEFS="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "test_operator_mul_bitflip or test_cheby_tfim_vs_oracle_and_expm or test_newton_liouvillian_bitflip"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
