"""Pins the CPU oracle against every analytic / known-answer pin the reference's own tests hold
for the Cheby / Newton / Arnoldi / specrange path.  The reference ships no
golden vectors and cannot run here (no Julia), so these pins plus the committed fixtures under
tests/golden/ are what anchor the oracle.  Runs on CPU.
"""

import os

import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import oracle as O
from oracle.controls import IdDict
import qprop_b200.workloads as W

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rand_state(rng, n):
    psi = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    return psi / np.linalg.norm(psi)


def test_tls_rabi_analytic():
    """test/test_propagate.jl:74-150 -- 3π/2 pulse, [1,0] -> [-1/√2, -i/√2]; backward returns."""
    H = np.array([[0, 0.5], [0.5, 0]], dtype=complex)
    tlist = np.linspace(0, 1.5 * np.pi, 101)
    psi0 = np.array([1, 0], dtype=complex)
    expected = np.array([-1 / np.sqrt(2), -1j / np.sqrt(2)])
    out = O.propagate(psi0, (H,), tlist, "cheby", inplace=False)
    assert np.linalg.norm(out - expected) < 1e-12
    back = O.propagate(out, (H,), tlist, "cheby", inplace=False, backward=True)
    assert np.linalg.norm(back - psi0) < 1e-12
    # Newton needs state dimension > 3: embed the TLS in 4 levels
    out_n = O.propagate(np.pad(psi0, (0, 2)), (np.pad(H, (0, 2)),), tlist, "newton")
    assert np.linalg.norm(out_n[:2] - expected) < 1e-12
    store = O.propagate(psi0, (H,), tlist, "cheby", storage=True)
    store_bw = O.propagate(store[:, -1].copy(), (H,), tlist, "cheby", storage=True, backward=True)
    assert abs(abs(store[0, -1]) ** 2 - 0.5) < 1e-12
    assert np.linalg.norm(store - store_bw) < 1e-12  # filled back to front, same trajectory


def test_optomech_newton_equals_cheby():
    """test/test_propagate.jl:153-163 + test/optomech.jl (N=55, 250 steps)."""
    H = W.optomech()
    assert H.shape == (55, 55)
    psi0 = W.optomech_ket(0, 2)
    tlist = np.arange(0, 50 + 1e-9, 0.2)
    assert len(tlist) == 251
    p1 = O.propagate(psi0, (H,), tlist, "newton")
    p2 = O.propagate(psi0, (H,), tlist, "cheby")
    assert (np.linalg.norm(p1) - 1.0) < 1e-12
    assert np.linalg.norm(p1 - p2) < 1e-10
    gold = np.load(os.path.join(GOLDEN, "optomech_final.npz"))
    assert np.linalg.norm(p2 - gold["cheby"]) < 1e-11
    assert np.linalg.norm(p1 - gold["newton"]) < 1e-11
    assert np.linalg.norm(p2 - gold["expm"]) < 1e-9  # independent: 250 dense expm steps


def test_cheby_random_hermitian_vs_exp():
    """test/test_cheby.jl:6-49: Hermitian(rand(ComplexF64,1000,1000)), dt=0.5, range from
    eigvals; 267 or 268 coefficients; ‖Δψ‖ < 1e-10; cheby_coeffs! ≡ cheby_coeffs."""
    rng = np.random.default_rng(0)
    N = 1000
    X = rng.random((N, N)) + 1j * rng.random((N, N))
    H = np.triu(X) + np.triu(X, 1).conj().T  # Julia's Hermitian(X): upper triangle
    H[np.diag_indices(N)] = H[np.diag_indices(N)].real
    dt = 0.5
    psi0 = rng.random(N) + 1j * rng.random(N)
    psi0 /= np.linalg.norm(psi0)
    ev = np.linalg.eigvalsh(H)
    expected = sla.expm(-1j * H * dt) @ psi0
    a = O.cheby_coeffs(ev[-1] - ev[0], dt)
    assert len(a) in (267, 268)
    n, b = O.cheby_coeffs_inplace(np.zeros(20), ev[-1] - ev[0], dt)
    assert n == len(a) and np.allclose(b[:n], a)
    psi = psi0.copy()
    wrk = O.ChebyWrk(psi0, ev[-1] - ev[0], ev[0], dt)
    O.cheby_inplace(psi, H, dt, wrk)
    assert np.linalg.norm(psi - expected) < 1e-10
    assert wrk.n_matvec == len(a) - 1
    assert np.linalg.norm(O.cheby(psi0, H, dt, wrk) - expected) < 1e-10


@pytest.mark.parametrize("hermitian,m_max", [(True, 5), (False, 50)])
def test_newton_random_vs_exp(hermitian, m_max):
    """test/test_newton.jl:7-127: N=1000 is reduced to 400 to keep the CPU suite short."""
    rng = np.random.default_rng(1)
    N = 400
    X = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    if hermitian:
        X = (X + X.conj().T) / 2
    X *= 10 / np.max(np.abs(np.linalg.eigvals(X)))
    psi0 = rand_state(rng, N)
    expected = sla.expm(-1j * X * 0.5) @ psi0
    psi = psi0.copy()
    O.newton_inplace(psi, X, 0.5, O.NewtonWrk(psi0, m_max=m_max), max_restarts=200)
    assert np.linalg.norm(psi - expected) < 1e-10
    if hermitian:
        assert abs(np.linalg.norm(psi) - 1) < 1e-10


def test_newton_sparse_liouvillian():
    """test/test_newton.jl:130-177: N=32 -> 1024, density 0.5 (here N=20 -> 400), func = exp."""
    rng = np.random.default_rng(2)
    N = 20
    L = sp.random(N * N, N * N, density=0.5, random_state=np.random.RandomState(3), format="csr")
    L = L + 1j * sp.random(N * N, N * N, density=0.5, random_state=np.random.RandomState(4), format="csr")
    L = (L * (10 / np.max(np.abs(np.linalg.eigvals(L.toarray()))))).tocsr()
    psi = rand_state(rng, N)
    rho0 = np.outer(psi, psi.conj()).reshape(-1)
    assert abs(np.trace(rho0.reshape(N, N)) - 1) < 1e-14
    expected = sla.expm(L.toarray() * 0.5) @ rho0
    rho = rho0.copy()
    O.newton_inplace(rho, L, 0.5, O.NewtonWrk(rho0, m_max=50), func=np.exp, max_restarts=20)
    assert np.linalg.norm(rho - expected) < 1e-10


def test_ritzvals_and_specrange():
    """test/test_specrad.jl:14-144."""
    rng = np.random.default_rng(5)
    w = W.config1_random(N=500, density=0.1, seed=9)
    H = w["ops"][0]
    ev = np.linalg.eigvalsh(H.toarray())
    D = ev[-1] - ev[0]
    E_min, E_max = O.specrange(H, "arnoldi", prec=1e-4, rng=rng)
    assert ev[0] - 0.05 * D <= E_min <= ev[0]
    assert ev[-1] <= E_max < ev[-1] + 0.05 * D
    lo, hi = O.specrange(H, "diag")
    assert abs(lo - ev[0]) < 1e-12 and abs(hi - ev[-1]) < 1e-12
    assert O.specrange(H, "manual", E_min=-10, E_max=10) == (-10.0, 10.0)
    assert O.specrange(H, E_min=-10, E_max=10) == (-10.0, 10.0)
    with pytest.raises(TypeError):
        O.specrange(H, "manual", E_min=-1.0)
    E_min, E_max = O.specrange(H, rng=rng)  # :auto -> :arnoldi
    assert ev[0] - 0.05 * D <= E_min <= ev[0] and ev[-1] <= E_max < ev[-1] + 0.05 * D
    # Hermitian Ritz values (m 20..60) bracket to 2 % like test_specrad.jl:50-75
    R = O.ritzvals(H, O.random_state(H, rng=rng), 20, 60, prec=1e-3)
    assert 20 <= len(R) <= 60
    assert abs(R[0].real - ev[0]) < 0.02 * D and abs(R[-1].real - ev[-1]) < 0.02 * D
    with pytest.raises(ValueError):
        O.ritzvals(H, O.random_state(H, rng=rng), 10, 10)


def test_cheby_init_prop_spectral_arithmetic():
    """test/test_specrad.jl:147-223: manual ±10 -> E_min = -10.1, Δ = 20.2; buffer 0.1 ->
    -11.0, 22.0; :diag + control_ranges exact."""
    w = W.config1_random(N=120, density=0.1, seed=3, nt=501, T=10.0)
    u = w["controls"][0]
    gen = O.hamiltonian(w["ops"][0], (w["ops"][1], u))
    p = O.init_prop(w["psi0"], gen, w["tlist"], "cheby", E_min=-10, E_max=10)
    assert abs(p.wrk.E_min + 10.1) < 1e-12 and abs(p.wrk.Delta - 20.2) < 1e-12
    assert p.wrk.n_coeffs == 9 and abs(p.wrk.dt - 0.02) < 1e-15
    p = O.init_prop(w["psi0"], gen, w["tlist"], "cheby", E_min=-10, E_max=10,
                    specrange_method="manual", specrange_buffer=0.1)
    assert abs(p.wrk.E_min + 11.0) < 1e-12 and abs(p.wrk.Delta - 22.0) < 1e-12
    p = O.init_prop(w["psi0"], gen, w["tlist"], "cheby", specrange_method="diag", specrange_buffer=0.0,
                    control_ranges=IdDict([(u, (-1, 1))]))
    H0, H1 = w["ops"]
    evs = [np.linalg.eigvalsh((H0 + s * H1).toarray()) for s in (-1, 0, 1)]
    assert abs(p.wrk.E_min - min(e[0] for e in evs)) < 1e-10
    assert abs(p.wrk.Delta - (max(e[-1] for e in evs) - p.wrk.E_min)) < 1e-10
    p = O.init_prop(w["psi0"], gen, w["tlist"], "cheby", rng=np.random.default_rng(0))
    lo, hi = min(e[0] for e in evs), max(e[-1] for e in evs)  # :arnoldi brackets within ~5 % + 1 % buffer
    assert lo - 0.07 * (hi - lo) < p.wrk.E_min <= lo and hi <= p.wrk.E_min + p.wrk.Delta < hi + 0.07 * (hi - lo)


def test_matvec_counts_like_the_reference_docs():
    """docs/src/benchmarks/profiling.md:112 -- N=200 random generator, 100 steps, dt=1:
    Cheby 1200 matrix-vector products (13 coefficients); Newton (m_max=10) 2 restarts/step."""
    rng = np.random.default_rng(7)
    H = W.random_sparse_hermitian(200, 0.2, 1.0, rng)
    ev = np.linalg.eigvalsh(H.toarray())
    assert O.cheby_coeffs(2.0, 1.0).size == 13  # α = 1 -> 13
    psi0 = rand_state(rng, 200)
    tlist = np.linspace(0, 100, 101)
    p = O.init_prop(psi0, (H,), tlist, "cheby", E_min=-1.0, E_max=0.98)
    for _ in range(100):
        O.prop_step(p)
    assert p.wrk.n_matvec == 100 * (p.wrk.n_coeffs - 1) == 1200
    pn = O.init_prop(psi0, (H,), tlist, "newton", m_max=10)
    for _ in range(100):
        O.prop_step(pn)
    assert pn.wrk.n_matvec[0] == 2000  # 2 Arnoldi sweeps of 10 per step
    assert np.linalg.norm(p.state - pn.state) < 1e-9
    assert abs(ev[0]) <= 1.0 and abs(ev[-1]) <= 1.0


def test_discretization_known_answers():
    """test/test_discretization.jl:8-76."""
    assert np.array_equal(O.get_tlist_midpoints([1, 3, 5, 6, 7]), [1, 4, 5.5, 7])
    assert np.array_equal(O.get_tlist_midpoints([1, 3, 5, 6, 7], preserve_start=False), [2, 4, 5.5, 7])
    assert np.array_equal(O.get_tlist_midpoints([1, 3, 5, 6, 7], preserve_end=False), [1, 4, 5.5, 6.5])
    with pytest.raises(ValueError):
        O.get_tlist_midpoints([0, 1])
    tlist = np.linspace(0, 10, 21)
    f = lambda t: np.sin(t) ** 2  # noqa: E731
    on_mid = O.discretize_on_midpoints(f, tlist)
    on_grid = O.discretize(on_mid, tlist)
    assert len(on_mid) == 20 and len(on_grid) == 21
    assert np.allclose(O.discretize_on_midpoints(on_grid, tlist), on_mid, atol=1e-14)
    assert np.allclose(O.discretize(f, tlist), on_grid, atol=1e-14)
    assert on_grid[0] == on_mid[0] and on_grid[-1] == on_mid[-1]
    assert O.t_mid(tlist, 1) == 0.0 and O.t_mid(tlist, 20) == 10.0 and O.t_mid(tlist, 2) == 0.75
    with pytest.raises(ValueError):
        O.discretize(np.zeros(5), tlist)


def test_operator_mul_and_evaluate():
    """test/test_operator_linalg.jl:30-64 (mul! with α,β ∈ {true,false,2.0}; 3-arg dot) and
    test/test_controls.jl:38-80 (evaluate / evaluate! equal H0 + Σ u_l H_l)."""
    rng = np.random.default_rng(11)
    n = 30
    mats = [rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) for _ in range(3)]
    op = O.Operator(mats, [0.3, -2.0j])
    dense = mats[0] + 0.3 * mats[1] - 2.0j * mats[2]
    assert np.allclose(op.toarray(), dense)
    x, y0 = rand_state(rng, n), rand_state(rng, n)
    for alpha in (True, 2.0):
        for beta in (False, True, 2.0):
            y = y0.copy()
            O.op_mul(y, op, x, alpha, beta)
            assert np.linalg.norm(y - (beta * y0 + alpha * dense @ x)) < 1e-12
    assert abs(O.op_dot(y0, op, x) - np.vdot(y0, dense @ x)) < 1e-12
    assert np.linalg.norm((2.0 * 1j) * (dense @ x) - O.ScaledOperator(2.0j, op) @ x) < 1e-12
    assert O.ScaledOperator(1.0, op) is op
    tlist = np.linspace(0, 1, 11)
    e1 = lambda t: 0.5 + t  # noqa: E731
    e2 = np.linspace(1, 2, 10)
    gen = O.hamiltonian(mats[0], (mats[1], e1), (mats[2], e2))
    assert isinstance(gen, O.Generator) and len(gen.ops) == 3
    from oracle.generators import evaluate, evaluate_inplace, get_controls

    assert get_controls(gen) == (e1, e2) or all(a is b for a, b in zip(get_controls(gen), (e1, e2)))
    op3 = evaluate(gen, tlist, 3)
    assert np.allclose(op3.toarray(), mats[0] + e1(0.25) * mats[1] + e2[2] * mats[2], atol=1e-15)
    vals = IdDict([(e1, 1.1), (e2, 2.2)])
    evaluate_inplace(op3, gen, tlist, 5, vals_dict=vals)
    assert np.allclose(op3.toarray(), mats[0] + 1.1 * mats[1] + 2.2 * mats[2], atol=1e-15)
    assert isinstance(O.hamiltonian(mats[0], (mats[1], 2.0)), O.Operator)
    assert O.hamiltonian(mats[0]) is mats[0]
    with pytest.raises(ValueError):
        O.Generator(mats[:1], [e1, e1])


def test_propagator_protocol_and_errors():
    """test/test_prop_interfaces.jl: stepping past the grid returns nothing; Newton is in-place
    only; unknown method; non-uniform grid rejected by Cheby."""
    w = W.config1_random(N=10, density=0.5, seed=21, nt=101, T=5.0)
    gen = O.hamiltonian(w["ops"][0], (w["ops"][1], w["controls"][0]))
    for method in ("cheby", "newton"):
        for backward in (False, True):
            kw = dict(E_min=-10, E_max=10) if method == "cheby" else {}
            p = O.init_prop(w["psi0"], gen, w["tlist"], method, backward=backward, **kw)
            s = O.prop_step(p)
            assert s is p.state and abs(np.linalg.norm(s) - 1) < 1e-12
            O.set_t(p, w["tlist"][0] if backward else w["tlist"][-1])
            assert O.prop_step(p) is None
            O.reinit_prop(p, w["psi0"])
            a = O.prop_step(p).copy()
            O.reinit_prop(p, w["psi0"])
            assert np.array_equal(a, O.prop_step(p))
    with pytest.raises(RuntimeError, match="only implemented in-place"):
        O.init_prop(w["psi0"], gen, w["tlist"], "newton", inplace=False)
    with pytest.raises(ValueError, match="Unknown propagation"):
        O.init_prop(w["psi0"], gen, w["tlist"], "foo")
    tl = w["tlist"].copy()
    tl[50] += 1e-3
    with pytest.warns(UserWarning), pytest.raises(RuntimeError, match="uniform time grid"):
        O.init_prop(w["psi0"], gen, tl, "cheby", E_min=-10, E_max=10)
    pn = O.init_prop(w["psi0"], gen, tl, "newton")  # Newton accepts non-uniform grids
    assert O.prop_step(pn) is not None


def test_tfim_golden_fixture():
    """Committed fixture (tests/golden/make_golden.py): TFIM n=8 with two controls, oracle Cheby
    and Newton vs exact PWC propagation by scipy expm_multiply."""
    gold = np.load(os.path.join(GOLDEN, "tfim8_final.npz"))
    w = W.config2_tfim(n_spins=8, nt=21, dt=0.1)
    assert np.array_equal(w["psi0"], gold["psi0"])
    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    out = O.propagate(w["psi0"], O.hamiltonian(*terms), w["tlist"], "cheby", E_min=w["E_min"], E_max=w["E_max"])
    assert np.linalg.norm(out - gold["cheby"]) < 1e-12
    assert np.linalg.norm(out - gold["exact"]) < 1e-10
    out_n = O.propagate(w["psi0"], O.hamiltonian(*terms), w["tlist"], "newton")
    assert np.linalg.norm(out_n - gold["exact"]) < 1e-10


def test_c_restatement_matches_numpy_oracle():
    """oracle/cheby_ref.c (the timed CPU baseline of bench.py): both the faithful single-thread
    CSC form and the OpenMP CSR variant equal the NumPy oracle to rounding."""
    from oracle import cref

    if not cref.available():
        pytest.skip("oracle/libcheby_ref.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    w = W.config2_tfim(n_spins=10, nt=6, dt=0.1)
    Delta = 1.01 * (w["E_max"] - w["E_min"])
    E_min = w["E_min"] - 0.005 * (w["E_max"] - w["E_min"])
    a = O.cheby_coeffs(Delta, 0.1)
    ref_c = cref.ChebyRef(w["ops"], 2)
    for dt in (0.1, -0.1):
        for coeffs in ([0.3, -0.2], [1.0 + 0.5j, 0.0]):
            expect = w["psi0"].copy()
            O.cheby_inplace(expect, O.Operator(list(w["ops"]), coeffs), dt, O.ChebyWrk(expect, Delta, E_min, 0.1))
            for threads in (0, 2):
                psi = w["psi0"].copy()
                n_mv = ref_c.step(psi, coeffs, a, Delta, E_min, dt, threads=threads)
                assert n_mv == len(a) - 1
                assert np.linalg.norm(psi - expect) < 1e-13
