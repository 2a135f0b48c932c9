"""GPU tests: the device-resident types pass the reference's interface checks
(``check_state`` / ``check_operator`` / ``check_generator`` / ``check_propagator``, mirror of
``test/test_prop_interfaces.jl`` and ``test/test_invalid_interfaces.jl``), and the Liouvillian
generator built by ``liouvillian`` propagates to the reference test's analytic answer."""

import logging

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O

pytestmark = pytest.mark.gpu


def rand_state(rng, n, B=None):
    shape = (n,) if B is None else (n, B)
    psi = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return psi / np.linalg.norm(psi, axis=0)


@pytest.mark.parametrize("n,B", [(2, None), (1000, None), (300, 7)])
def test_device_state_passes_check_state(qp, ctx, n, B):
    rng = np.random.default_rng(n)
    st = qp.DeviceState.from_host(ctx, rand_state(rng, n, B))
    assert qp.supports_inplace(st) and not qp.supports_vector_interface(st)
    assert qp.check_state(st, normalized=True)
    st.lmul(3.0)
    assert qp.check_state(st) and not qp.check_state(st, normalized=True, quiet=True)


@pytest.mark.parametrize("kind", ["csr", "dense", "lazy_sum", "scaled"])
def test_device_operators_pass_check_operator(qp, ctx, kind):
    rng = np.random.default_rng(3)
    n = 200
    A = sp.random(n, n, density=0.05, random_state=np.random.RandomState(1), format="csr") * (1 + 0.5j)
    A = (A + A.conj().T).tocsr()
    Bm = sp.diags(rng.standard_normal(n) + 0j, 0, format="csr")
    op = {
        "csr": A,
        "dense": np.asarray(A.toarray()),
        "lazy_sum": qp.Operator([A, Bm], [0.3 - 0.1j]),
        "scaled": 2.5 * qp.Operator([A, Bm], [0.3]),
    }[kind]
    st = qp.DeviceState.from_host(ctx, rand_state(rng, n))
    assert qp.check_operator(op, state=st, tlist=np.linspace(0, 1, 5))
    # batched states: the operator verbs act on every trajectory
    stB = qp.DeviceState.from_host(ctx, rand_state(rng, n, 20))
    assert qp.check_operator(op, state=stB)


def test_check_operator_reports_dimension_mismatch(qp, ctx, caplog):
    st = qp.DeviceState.from_host(ctx, rand_state(np.random.default_rng(0), 10))
    with caplog.at_level(logging.ERROR, logger="qprop_b200.interfaces"):
        assert not qp.check_operator(sp.identity(11, dtype=complex, format="csr"), state=st)
    assert "`op * state` must be defined" in caplog.text and "dimension" in caplog.text


def test_generator_passes_check_generator(qp, ctx):
    w = qp.workloads.config2_tfim(8, nt=11, dt=0.1)
    G = qp.hamiltonian(w["ops"][0], (w["ops"][1], w["controls"][0]), (w["ops"][2], w["controls"][1]))
    st = qp.DeviceState.from_host(ctx, w["psi0"])
    assert qp.check_generator(G, state=st, tlist=w["tlist"], for_time_continuous=True)
    # tuple generators are canonicalised through hamiltonian()
    assert qp.check_generator((w["ops"][0], (w["ops"][1], w["controls"][0])), state=st, tlist=w["tlist"])
    # an amplitude that is not a number on the grid is reported
    bad = qp.hamiltonian(w["ops"][0], (w["ops"][1], lambda t: "x"))
    assert not qp.check_generator(bad, state=st, tlist=w["tlist"], quiet=True)


@pytest.mark.parametrize("method,backward,inplace", [("cheby", False, True), ("cheby", True, True), ("newton", False, True), ("cheby", False, False)])
def test_propagators_pass_check_propagator(qp, ctx, method, backward, inplace):
    w = qp.workloads.config1_random(N=40, density=0.3, seed=21, nt=21, T=1.0)
    G = qp.hamiltonian(w["ops"][0], (w["ops"][1], w["controls"][0]))
    kw = dict(E_min=w["E_min"], E_max=w["E_max"]) if method == "cheby" else {}
    state = qp.DeviceState.from_host(ctx, w["psi0"]) if inplace else w["psi0"]
    p = qp.init_prop(state, G, w["tlist"], method, ctx=ctx, backward=backward, inplace=inplace, **kw)
    assert qp.check_propagator(p)


def test_liouvillian_tls_dissipation_newton_on_device(qp, ctx):
    """test/test_liouvillian.jl "TLS dissipation" through the device Newton propagator (the
    reference uses expprop): analytic density matrix after T = 1."""
    g1, g2, T = 0.5, 0.2, 1.0
    A1 = np.sqrt(g1) * np.array([[0, 1], [0, 0]], dtype=complex)
    A2 = np.sqrt(2 * g2) * np.array([[0, 0], [0, 1]], dtype=complex)
    psi0 = np.array([1, 1], dtype=complex) / np.sqrt(2)
    rho0 = np.outer(psi0, psi0.conj()).reshape(-1, order="F")
    L = qp.liouvillian(None, [A1, A2], convention="TDSE")
    tlist = np.linspace(0.0, T, 11)
    out = qp.propagate(rho0, L, tlist, "newton", ctx=ctx)
    rho = np.asarray(out).reshape(2, 2, order="F")
    e1, e2 = np.exp(-g1 * T), np.exp(-(g1 / 2 + g2) * T)
    expected = 0.5 * np.array([[2 - e1, e2], [e2, e1]], dtype=complex)
    assert abs(1 - np.trace(rho)) < 1e-12
    assert np.linalg.norm(rho - expected) < 1e-12


def test_liouvillian_generator_vs_oracle(qp, ctx):
    """Driven, dissipative two-spin system: liouvillian(hamiltonian(...), c_ops) on the device
    (Newton) against the oracle on the same super-operators."""
    w = qp.workloads.config2_tfim(2, nt=9, dt=0.1)
    H = qp.hamiltonian(w["ops"][0], (w["ops"][1], w["controls"][0]))
    sm = np.array([[0, 1], [0, 0]], dtype=complex)
    c_ops = [np.sqrt(0.05) * np.kron(sm, np.eye(2)), np.sqrt(0.05) * np.kron(np.eye(2), sm)]
    Lg = qp.liouvillian(H, c_ops, convention="TDSE")
    assert isinstance(Lg, qp.Generator) and qp.get_controls(Lg) == (w["controls"][0],)
    rng = np.random.default_rng(2)
    psi = rand_state(rng, 4)
    rho0 = np.outer(psi, psi.conj()).reshape(-1, order="F")
    out = qp.propagate(rho0, Lg, w["tlist"], "newton", ctx=ctx)
    ref = O.propagate(rho0, O.hamiltonian(Lg.ops[0], (Lg.ops[1], w["controls"][0])), w["tlist"], "newton")
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-10
    assert abs(np.trace(np.asarray(out).reshape(4, 4, order="F")) - 1) < 1e-10


# ---------------------------------------------------------------------------------------
# matrix-free left/right super-operators (qp_op_create_leftright, QP_FORMAT_LR)
# ---------------------------------------------------------------------------------------


def _rand_sparse(rng, n, density):
    A = sp.random(n, n, density=density, random_state=np.random.RandomState(rng.integers(1 << 30)), format="csr")
    B = sp.random(n, n, density=density, random_state=np.random.RandomState(rng.integers(1 << 30)), format="csr")
    return (A + 1j * B).tocsr()


@pytest.mark.parametrize("n", [5, 37, 64])
def test_leftright_operator_mul_matches_explicit_kron(qp, ctx, n):
    """Σ c P ρ Q with left-only, right-only, sandwich and pure-scalar terms, two operators with a
    coefficient, n not a multiple of the warp size: mul! / dot / expval against Σ c Qᵀ ⊗ P."""
    rng = np.random.default_rng(n)
    P1, P2, Q1, Q2 = (_rand_sparse(rng, n, 0.2) for _ in range(4))
    op0 = qp.LeftRightOperator(n, [(P1, None, 0.7), (None, Q1, -0.4j), (P2, Q2, 1.1 - 0.3j), (None, None, 0.25)])
    op1 = qp.LeftRightOperator(n, [(P2, None, 1.0), (None, P2, -1.0), (Q1, P1, 0.5j)])
    coeffs = [0.6 - 0.2j]
    gen = qp.DeviceGenerator(ctx, [op0, op1], 1)
    assert gen.format == "leftright" and gen.n == n * n
    full = (op0.tosparse() + coeffs[0] * op1.tosparse()).tocsr()
    x = rand_state(rng, n * n)
    y0 = rand_state(rng, n * n)
    dx = qp.DeviceState.from_host(ctx, x)
    for alpha, beta in ((1.0, 0.0), (0.5 - 1j, 2.0)):
        dy = qp.DeviceState.from_host(ctx, y0)
        gen.mul(dy, dx, coeffs, alpha, beta)
        ref = beta * y0 + alpha * (full @ x)
        assert np.linalg.norm(dy.to_host() - ref) / np.linalg.norm(ref) < 1e-13
    assert abs(gen.expval(dx, coeffs) - np.vdot(x, full @ x)) < 1e-12 * n
    dy0 = qp.DeviceState.from_host(ctx, y0)
    assert abs(gen.dot(dy0, dx, coeffs) - np.vdot(y0, full @ x)) < 1e-12 * n
    # the lazy-sum Operator of the host mirror passes the operator interface check
    assert qp.check_operator(qp.Operator([op0, op1], coeffs), state=dx)


def test_leftright_errors(qp, ctx):
    n = 8
    rng = np.random.default_rng(0)
    lr = qp.LeftRightOperator(n, [(_rand_sparse(rng, n, 0.3), None, 1.0)])
    with pytest.raises(qp.QPropError, match="mixing"):
        qp.DeviceGenerator(ctx, [lr, sp.identity(n * n, dtype=complex, format="csr")], 0)
    gen = qp.DeviceGenerator(ctx, [lr], 0)
    xB = qp.DeviceState.from_host(ctx, rand_state(rng, n * n, 3))
    with pytest.raises(qp.QPropError, match="single states"):
        gen.mul(xB.similar(), xB, [])
    with pytest.raises(ValueError, match="shape"):
        qp.LeftRightOperator(n, [(sp.identity(n + 1, format="csr"), None, 1.0)])


@pytest.mark.parametrize("n_spins", [3, 5])
def test_liouvillian_matrix_free_newton_vs_explicit_and_oracle(qp, ctx, n_spins):
    """Config 4 shape (driven TFIM + local decay, TDSE convention) propagated with Newton through
    the matrix-free generator, through the explicit sparse super-operators, and by the oracle."""
    H0, H1, _ = qp.workloads.tfim_chain(n_spins)
    nh = 1 << n_spins
    sm = sp.csr_matrix(np.array([[0, 1], [0, 0]], dtype=complex))
    c_ops = []
    for k in range(n_spins):
        left = sp.identity(1 << (n_spins - 1 - k), dtype=complex, format="csr")
        right = sp.identity(1 << k, dtype=complex, format="csr")
        c_ops.append(np.sqrt(0.05) * sp.kron(sp.kron(left, sm), right, format="csr"))
    tlist = np.linspace(0.0, 0.4, 9)

    def u1(t):
        return float(np.sin(np.pi * t / 0.4) ** 2)

    Lf = qp.liouvillian((H0, (H1, u1)), c_ops, convention="TDSE", matrix_free=True)
    Lm = qp.liouvillian((H0, (H1, u1)), c_ops, convention="TDSE")
    rng = np.random.default_rng(n_spins)
    psi = rand_state(rng, nh)
    rho0 = np.outer(psi, psi.conj()).reshape(-1, order="F")
    out_f = qp.propagate(rho0, Lf, tlist, "newton", ctx=ctx)
    out_m = qp.propagate(rho0, Lm, tlist, "newton", ctx=ctx)
    ref = O.propagate(rho0, O.hamiltonian(Lm.ops[0], (Lm.ops[1], u1)), tlist, "newton")
    assert np.linalg.norm(out_f - ref) / np.linalg.norm(ref) < 1e-10
    assert np.linalg.norm(out_f - out_m) / np.linalg.norm(out_m) < 1e-11
    rho = np.asarray(out_f).reshape(nh, nh, order="F")
    assert abs(np.trace(rho) - 1) < 1e-10 and np.linalg.norm(rho - rho.conj().T) < 1e-10


def test_liouvillian_matrix_free_cheby_unitary(qp, ctx):
    """Without dissipation the super-operator [H, ·] is Hermitian: the Chebyshev propagator (all
    fused epilogues of the matrix-free kernel) must reproduce ρ(t) = U ρ U†."""
    import scipy.linalg as sla

    H0, H1, _ = qp.workloads.tfim_chain(4)
    nh = 16
    Hs = (H0 + 0.3 * H1).toarray()
    ev = np.linalg.eigvalsh(Hs)
    Lf = qp.liouvillian(H0 + 0.3 * H1, convention="TDSE", matrix_free=True)
    rng = np.random.default_rng(4)
    psi = rand_state(rng, nh)
    rho0 = np.outer(psi, psi.conj())
    tlist = np.linspace(0.0, 1.0, 6)
    width = ev[-1] - ev[0]
    out = qp.propagate(rho0.reshape(-1, order="F"), Lf, tlist, "cheby", ctx=ctx, E_min=-width, E_max=width)
    U = sla.expm(-1j * Hs * 1.0)
    ref = U @ rho0 @ U.conj().T
    assert np.linalg.norm(np.asarray(out).reshape(nh, nh, order="F") - ref) < 1e-11


def test_propagate_sequence(qp, ctx):
    """test/test_propagate_sequence.jl in miniature: three propagations with different generators,
    instantaneous pre/post transformations, a pre-initialised propagator and storage=True."""
    import scipy.linalg as sla

    rng = np.random.default_rng(12)
    n = 30

    def herm(scale):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        return scale * (A + A.conj().T) / 2

    Ha, Hb, Hc = herm(0.3), herm(0.2), herm(0.4)
    psi0 = rand_state(rng, n)
    t1, t2, t3 = np.linspace(0, 1, 11), np.linspace(1, 1.5, 6), np.linspace(1.5, 3.0, 4)
    phase = np.exp(1j * np.linspace(0, 1, n))
    calls = []

    def to_frame(psi, *args, **kwargs):
        calls.append("pre")
        return phase * np.asarray(psi)

    def from_frame(psi, *args, **kwargs):
        calls.append("post")
        return np.conj(phase) * np.asarray(psi)

    kw = dict(method="newton", ctx=ctx)
    pc = qp.init_prop(psi0, Hc, t3, "newton", ctx=ctx)  # pre-initialised propagator for the last step
    seq = [
        qp.Propagation(Ha, t1),
        qp.Propagation(Hb, t2, pre_propagation=to_frame, post_propagation=from_frame),
        qp.Propagation(pc),
    ]
    states = qp.propagate_sequence(psi0, seq, **kw)
    assert len(states) == 3 and calls == ["pre", "post"]
    s1 = sla.expm(-1j * Ha * 1.0) @ psi0
    s2 = np.conj(phase) * (sla.expm(-1j * Hb * 0.5) @ (phase * s1))
    s3 = sla.expm(-1j * Hc * 1.5) @ s2
    for got, want in zip(states, (s1, s2, s3)):
        assert np.linalg.norm(np.asarray(got) - want) < 1e-10
    # storage=True: one storage array per step, populations as observable
    pops = [lambda psi: np.abs(np.asarray(psi.to_host() if hasattr(psi, "to_host") else psi)) ** 2]
    stor = qp.propagate_sequence(psi0, [qp.Propagation(Ha, t1), qp.Propagation(Hb, t2)], storage=True, observables=pops, **kw)
    assert stor[0].shape[-1] == len(t1) and stor[1].shape[-1] == len(t2)
    assert np.allclose(np.squeeze(stor[0])[..., -1], np.abs(s1) ** 2, atol=1e-10)
    with pytest.raises(TypeError):
        qp.propagate_sequence(psi0, [(Ha, t1)])


def test_storage_module_device_states(qp, ctx):
    """Storage of device-resident states (a list of per-slot copies) and fused expectation values
    of matrix observables on a DeviceState."""
    rng = np.random.default_rng(1)
    n = 50
    psi = rand_state(rng, n)
    st = qp.DeviceState.from_host(ctx, psi)
    tlist = np.linspace(0, 1, 4)
    slots = qp.init_storage(st, tlist)
    assert slots == [None] * 4
    for i in range(1, 5):
        qp.write_to_storage(slots, i, st)
        st.lmul(2.0)
    assert slots[0] is not st and np.allclose(slots[2].to_host(), 4 * psi)
    back = qp.DeviceState(ctx, n)
    assert qp.get_from_storage_(back, slots, 2) is back and np.allclose(back.to_host(), 2 * psi)
    host = np.zeros(n, dtype=complex)
    qp.get_from_storage_(host, slots, 4)
    assert np.allclose(host, 8 * psi)
    O1 = sp.diags(rng.standard_normal(n) + 0j, 0, format="csr")
    O2 = np.diag(rng.standard_normal(n)).astype(complex)
    s0 = qp.DeviceState.from_host(ctx, psi)
    data = qp.map_observables((O1, O2), tlist, 1, s0)
    assert np.allclose(data, [np.vdot(psi, O1 @ psi), np.vdot(psi, O2 @ psi)], atol=1e-13)
    assert qp.init_storage(s0, tlist, (O1, O2)).shape == (2, 4)
    assert abs(qp.map_observable(lambda s: s.norm(), tlist, 1, s0) - 1) < 1e-14


def test_propagate_with_shaped_amplitude(qp, ctx):
    """A Generator whose amplitude is ShapedAmplitude(eps; shape=S) propagates exactly like the plain
    control u(t) = S(t) eps(t) (the PWC value of eps is what `parameters` holds; the shape is
    applied at evaluation time), and `parameters` lists the control, not the amplitude."""
    from qprop_b200.shapes import flattop

    w = qp.workloads.config1_random(N=60, density=0.2, seed=5, nt=41, T=4.0)
    H0, H1 = w["ops"]

    def eps(t):
        return float(np.cos(1.3 * t))

    def S(t):
        return flattop(t, T=4.0, t_rise=1.0)

    kw = dict(E_min=w["E_min"], E_max=w["E_max"], ctx=ctx)
    G_amp = qp.hamiltonian(H0, (H1, qp.ShapedAmplitude(eps, shape=S)))
    G_fun = qp.hamiltonian(H0, (H1, lambda t: S(t) * eps(t)))
    p = qp.init_prop(w["psi0"], G_amp, w["tlist"], "cheby", **kw)
    assert list(p.parameters.keys()) == [eps] and len(p.parameters[eps]) == 40
    out_amp = qp.propagate(p)
    out_fun = qp.propagate(w["psi0"], G_fun, w["tlist"], "cheby", **kw)
    assert np.linalg.norm(out_amp - out_fun) < 1e-13
    st = qp.DeviceState.from_host(ctx, w["psi0"])
    assert qp.check_generator(G_amp, state=st, tlist=w["tlist"], for_time_continuous=True)


@pytest.mark.parametrize("n_spins", [4, 5])
def test_liouvillian_matrix_free_bitflip_form(qp, ctx, n_spins):
    """The factors of the matrix-free Liouvillian (H = diagonal + bit flips on the left or right of rho, the
    decay channels sigma^- rho sigma^+ as conditional flips on both sides, the anticommutators as diagonals) compose
    to the bit-flip form on n^2 rows without the n^2 x n^2 matrices: the library detects it (forced here, AUTO
    from N >= 148 * 1024), and application and Newton propagation agree with the generic matrix-free kernel, with
    the explicit super-operators and with the oracle."""
    H0, H1, _ = qp.workloads.tfim_chain(n_spins)
    nh = 1 << n_spins
    sm = sp.csr_matrix(np.array([[0, 1], [0, 0]], dtype=complex))
    c_ops = []
    for k in range(n_spins):
        left = sp.identity(1 << (n_spins - 1 - k), dtype=complex, format="csr")
        right = sp.identity(1 << k, dtype=complex, format="csr")
        c_ops.append(np.sqrt(0.05) * sp.kron(sp.kron(left, sm), right, format="csr"))
    tlist = np.linspace(0.0, 0.4, 9)

    def u1(t):
        return float(np.sin(np.pi * t / 0.4) ** 2)

    Lf = qp.liouvillian((H0, (H1, u1)), c_ops, convention="TDSE", matrix_free=True)
    Lm = qp.liouvillian((H0, (H1, u1)), c_ops, convention="TDSE")
    rng = np.random.default_rng(n_spins)
    x = rand_state(rng, nh * nh)
    dx = qp.DeviceState.from_host(ctx, x)
    got = {}
    for fmt in ("bitflip", "leftright"):
        gen = qp.DeviceGenerator(ctx, list(Lf.ops), 1, fmt)
        assert gen.format == fmt
        dy = qp.DeviceState(ctx, nh * nh).zero()
        gen.mul(dy, dx, [0.37], 1.0, 0.0)
        got[fmt] = dy.to_host()
    want = (Lm.ops[0] + 0.37 * Lm.ops[1]) @ x
    assert np.linalg.norm(got["bitflip"] - want) / np.linalg.norm(want) < 1e-13
    assert np.linalg.norm(got["bitflip"] - got["leftright"]) / np.linalg.norm(want) < 1e-13
    psi = rand_state(rng, nh)
    rho0 = np.outer(psi, psi.conj()).reshape(-1, order="F")
    out_b = qp.propagate(rho0, Lf, tlist, "newton", ctx=ctx, matrix_format="bitflip")
    ref = O.propagate(rho0, O.hamiltonian(Lm.ops[0], (Lm.ops[1], u1)), tlist, "newton")
    assert np.linalg.norm(out_b - ref) / np.linalg.norm(ref) < 1e-10


def test_leftright_bitflip_refuses_flip_times_varying_diagonal(qp, ctx):
    """A flip on one side multiplied by a NON-constant diagonal on the other side is row dependent: not a
    bit-flip operator (forced format: error; AUTO keeps the generic matrix-free kernel)."""
    n = 16
    rows = np.arange(n)
    X = sp.csr_matrix((np.ones(n, dtype=complex), (rows, rows ^ 2)), shape=(n, n))
    D = sp.diags(np.arange(1, n + 1).astype(complex)).tocsr()
    op = qp.LeftRightOperator(n, [(X, D, 1.0)])
    with pytest.raises(qp.QPropError):
        qp.DeviceGenerator(ctx, [op], 0, "bitflip")
    gen = qp.DeviceGenerator(ctx, [op], 0)
    assert gen.format == "leftright"
    rng = np.random.default_rng(3)
    rho = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    dy = qp.DeviceState(ctx, n * n).zero()
    gen.mul(dy, qp.DeviceState.from_host(ctx, rho.reshape(-1, order="F")), [], 1.0, 0.0)
    want = (X @ rho @ D).reshape(-1, order="F")
    assert np.linalg.norm(dy.to_host() - want) / np.linalg.norm(want) < 1e-13
    # constant diagonal on the other side (a multiple of the identity): fine
    op2 = qp.LeftRightOperator(n, [(X, 2.5 * sp.identity(n, dtype=complex, format="csr"), 1.0), (D, D, 0.5j)])
    gen2 = qp.DeviceGenerator(ctx, [op2], 0, "bitflip")
    assert gen2.format == "bitflip"
    gen2.mul(dy, qp.DeviceState.from_host(ctx, rho.reshape(-1, order="F")), [], 1.0, 0.0)
    want2 = (2.5 * X @ rho + 0.5j * D @ rho @ D).reshape(-1, order="F")
    assert np.linalg.norm(dy.to_host() - want2) / np.linalg.norm(want2) < 1e-13
