"""CPU tests of the host-side interface checks (mirror of the reference's
``test/test_invalid_interfaces.jl`` / ``test/test_prop_interfaces.jl`` on host objects), of
``substitute`` and of ``liouvillian`` (``test/test_liouvillian.jl``)."""

import logging

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

import oracle as O
import qprop_b200 as qp


def _herm(rng, n, rho):
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = (A + A.conj().T) / 2
    return A * (rho / np.max(np.abs(np.linalg.eigvalsh(A))))


def test_check_tlist():
    assert qp.check_tlist(np.linspace(0, 1, 11))
    assert not qp.check_tlist(np.array([0.0]), quiet=True)                # fewer than two points
    assert not qp.check_tlist(np.array([0.0, 1.0, 1.0]), quiet=True)      # not increasing
    assert not qp.check_tlist([0.0, 1.0], quiet=True)                     # not a Vector{Float64}
    assert not qp.check_tlist(np.array([0, 1]), quiet=True)               # integer grid


def test_check_control_and_amplitude(caplog):
    tlist = np.linspace(0, 10, 101)
    assert qp.check_control(lambda t: np.sin(t), tlist)
    assert qp.check_control(np.zeros(101), tlist) and qp.check_control(np.zeros(100), tlist)
    assert qp.check_amplitude(lambda t: 0.5 * t, tlist) and qp.check_amplitude(0.3 + 0.1j, tlist)
    with caplog.at_level(logging.ERROR, logger="qprop_b200.interfaces"):
        assert not qp.check_control(np.zeros(50), tlist)                 # wrong length
        assert not qp.check_control(lambda t: 1, tlist)                  # returns an Int, not a Float64
        assert not qp.check_control(lambda t: float("nan"), tlist)       # not finite
    assert "must return a Float64" in caplog.text and "must be finite" in caplog.text
    assert not qp.check_control(lambda t: "x", tlist, quiet=True)


def test_check_state_host_vectors(caplog):
    psi = np.array([1, 1j, 0], dtype=complex) / np.sqrt(2)
    assert qp.check_state(psi, normalized=True)
    assert qp.supports_inplace(psi) and qp.supports_vector_interface(psi)
    with caplog.at_level(logging.ERROR, logger="qprop_b200.interfaces"):
        assert not qp.check_state(2 * psi, normalized=True)
    assert "`norm(state)` must be 1" in caplog.text
    ro = psi.copy()
    ro.flags.writeable = False
    assert not qp.supports_inplace(ro) and qp.check_state(ro)            # immutable states are valid states

    class Broken:  # no Hilbert-space verbs at all
        pass

    assert not qp.check_state(Broken(), quiet=True)


def test_check_operator_and_generator_host():
    rng = np.random.default_rng(5)
    n = 12
    H0, H1 = _herm(rng, n, 1.0), _herm(rng, n, 0.1)
    psi = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    tlist = np.linspace(0, 1, 11)
    assert qp.check_operator(H0, state=psi) and qp.check_operator(sp.csr_matrix(H1), state=psi, tlist=tlist)
    assert qp.supports_matrix_interface(H0) and not qp.supports_matrix_interface(psi)
    assert not qp.check_operator(np.ones((n + 1, n + 1)), state=psi, quiet=True)   # wrong dimension

    def eps(t):
        return float(np.sin(t))

    G = qp.hamiltonian(H0, (H1, eps))
    assert isinstance(G, qp.Generator)
    # a generator must evaluate to an operator usable on the state: host matrices evaluate into the
    # package's lazy Operator, whose verbs live on the device -- so the generator check on HOST
    # states covers controls / amplitudes / substitute only
    assert qp.check_generator(G, state=psi, tlist=tlist, for_pwc=False)
    assert qp.get_controls(G) == (eps,)
    assert not qp.check_generator(qp.hamiltonian(H0, (H1, lambda t: 1j)), state=psi, tlist=tlist, for_pwc=False, quiet=True)


def test_substitute():
    H0, H1 = np.eye(2, dtype=complex), np.ones((2, 2), dtype=complex)

    def a(t):
        return 1.0

    def b(t):
        return 2.0

    G = qp.hamiltonian(H0, (H1, a))
    assert qp.substitute(G, qp.IdDict([(a, a)])) is G                      # nothing changes: same object
    G2 = qp.substitute(G, qp.IdDict([(a, b)]))
    assert G2 is not G and qp.get_controls(G2) == (b,) and G2.ops[1] is H1
    G3 = qp.substitute(G, qp.IdDict([(H1, H0)]))
    assert G3.ops[1] is H0 and qp.get_controls(G3) == (a,)
    T = qp.substitute((H0, (H1, a)), qp.IdDict([(a, b)]))
    assert T[0] is H0 and T[1][0] is H1 and T[1][1] is b
    assert qp.controls.substitute(a, {}) is a


def test_liouvillian_tls_dissipation_known_answer():
    """test/test_liouvillian.jl "TLS dissipation": analytic density matrix after T = 1."""
    g1, g2, T = 0.5, 0.2, 1.0
    A1 = np.sqrt(g1) * np.array([[0, 1], [0, 0]], dtype=complex)
    A2 = np.sqrt(2 * g2) * np.array([[0, 0], [0, 1]], dtype=complex)
    psi0 = np.array([1, 1], dtype=complex) / np.sqrt(2)
    rho0 = np.outer(psi0, psi0.conj()).reshape(-1, order="F")
    L = qp.liouvillian(None, [A1, A2], convention="TDSE")
    assert sp.issparse(L) and L.shape == (4, 4)
    rho = (sla.expm(-1j * L.toarray() * T) @ rho0).reshape(2, 2, order="F")
    e1, e2 = np.exp(-g1 * T), np.exp(-(g1 / 2 + g2) * T)
    expected = 0.5 * np.array([[2 - e1, e2], [e2, e1]], dtype=complex)
    assert abs(1 - np.trace(rho)) < 1e-15 and abs(np.trace(rho @ rho)) < 1.0
    assert np.linalg.norm(rho - expected) < 1e-15
    # the oracle's Newton propagator on the same generator
    out = O.propagate(rho0, L, np.array([0.0, T]), "newton")
    assert np.linalg.norm(out.reshape(2, 2, order="F") - expected) < 1e-12


def test_liouvillian_matches_lindblad_equation():
    """test/test_liouvillian.jl "LvN": ℒρ⃗ against the Lindblad right-hand side."""
    rng = np.random.default_rng(11)
    n = 20
    H0, H1 = _herm(rng, n, 1.0), _herm(rng, n, 0.1)

    def eps(t):
        return 1.0

    H = H0 + H1
    psi = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    psi /= np.linalg.norm(psi)
    rho = np.outer(psi, psi.conj())
    vec = rho.reshape(-1, order="F")
    unvec = lambda v: v.reshape(n, n, order="F")  # noqa: E731
    L = qp.liouvillian(H, convention="LvN")
    assert sp.issparse(L)
    assert np.linalg.norm(1j * (H @ rho - rho @ H) - unvec(L @ vec)) < 1e-14
    L = qp.liouvillian(H, convention="TDSE")
    assert np.linalg.norm((H @ rho - rho @ H) - unvec(L @ vec)) < 1e-14
    ket = np.eye(n, dtype=complex)
    c_ops = [np.sqrt(0.2) * np.outer(ket[0], ket[i]) for i in range(1, n)]
    c_ops += [np.sqrt(0.1) * np.outer(ket[i], ket[i]) for i in range(n)]

    def rhs(Hm):
        out = 1j * (Hm @ rho - rho @ Hm)
        for A in c_ops:
            AdA = A.conj().T @ A
            out = out + A @ rho @ A.conj().T - AdA @ rho / 2 - rho @ AdA / 2
        return out

    L = qp.liouvillian(H0, c_ops, convention="LvN")
    assert sp.issparse(L) and np.linalg.norm(unvec(L @ vec) - rhs(H0)) < 1e-14
    for Hgen in ((H0, (H1, eps)), qp.hamiltonian(H0, (H1, eps))):
        Lg = qp.liouvillian(Hgen, c_ops, convention="LvN")
        assert isinstance(Lg, qp.Generator) and qp.get_controls(Lg) == (eps,)
        full = Lg.ops[0] + Lg.ops[1] * Lg.amplitudes[0](0.0)
        L0 = qp.evaluate(Lg, 0.0)
        assert np.linalg.norm(full.toarray() - L0.toarray()) < 1e-12
        assert np.linalg.norm(unvec(full @ vec) - rhs(H)) < 1e-14
    # TDSE convention (src/generators.jl:470-508): commutator without the factor i, dissipator times i
    Lt = qp.liouvillian(H0, c_ops, convention="TDSE")
    diss = rhs(H0) - 1j * (H0 @ rho - rho @ H0)
    assert np.linalg.norm(unvec(Lt @ vec) - ((H0 @ rho - rho @ H0) + 1j * diss)) < 1e-14
    import pytest

    with pytest.raises(ValueError, match="convention"):
        qp.liouvillian(H0, convention="SE")
    with pytest.raises(ValueError, match="Empty"):
        qp.liouvillian(None, [], convention="LvN")
    with pytest.raises(TypeError):
        qp.liouvillian(H0)  # the convention is mandatory


def test_liouvillian_matrix_free_equals_explicit_superoperators():
    """``liouvillian(..., matrix_free=True)`` keeps every super-operator as its n x n factors
    (``LeftRightOperator``: Σ c P ρ Q); expanding them with Kronecker products gives exactly the
    matrices of the reference's formulas (src/generators.jl:470-508), in both conventions."""
    rng = np.random.default_rng(1)
    n = 7
    H0, H1 = _herm(rng, n, 1.0), _herm(rng, n, 0.3)
    c_ops = [0.3 * np.diag(np.ones(n - 1), 1).astype(complex), 0.2 * np.diag(np.arange(n)).astype(complex)]

    def u(t):
        return 0.5

    rho = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    for conv in ("TDSE", "LvN"):
        Lm = qp.liouvillian((H0, (H1, u)), c_ops, convention=conv)
        Lf = qp.liouvillian((H0, (H1, u)), c_ops, convention=conv, matrix_free=True)
        assert isinstance(Lf, qp.Generator) and qp.get_controls(Lf) == (u,)
        for a, b in zip(Lm.ops, Lf.ops):
            assert isinstance(b, qp.LeftRightOperator) and b.shape == a.shape
            assert abs(a - b.tosparse()).max() < 1e-15
            assert np.abs(a @ rho - b @ rho).max() < 1e-13
        # left-only and right-only factors are merged: drift = H_eff ρ, ρ H_eff†-like, 2 sandwiches
        assert len(Lf.ops[0].terms) == 4 and len(Lf.ops[1].terms) == 2
    D = qp.liouvillian(None, c_ops, convention="LvN", matrix_free=True)
    assert isinstance(D, qp.LeftRightOperator)
    assert abs(D.tosparse() - qp.liouvillian(None, c_ops, convention="LvN")).max() < 1e-15
    assert abs((2.0 * D).tosparse() - 2.0 * D.tosparse()).max() < 1e-15


def test_storage_module_host():
    """Mirror of the reference's Storage module on host data (src/storage.jl; docs/src/storage.md)."""
    tlist = np.linspace(0, 1, 5)
    psi = np.array([1, 1j, 0], dtype=complex) / np.sqrt(2)
    # states: n x nt matrix, written column by column, read back in place / out of place
    st = qp.init_storage(psi, tlist)
    assert st.shape == (3, 5) and st.dtype == psi.dtype
    for i in range(1, 6):
        qp.write_to_storage(st, i, i * psi)
    assert np.array_equal(qp.get_from_storage(st, 3), 3 * psi)
    buf = np.zeros(3, dtype=complex)
    assert qp.get_from_storage_(buf, st, 5) is buf and np.array_equal(buf, 5 * psi)
    # observables: matrices (expectation values), functions of (state) and of (state, tlist, i)
    Z = np.diag([1.0, -1.0, 0.0]).astype(complex)
    obs = (Z, sp.csr_matrix(Z @ Z))
    data = qp.map_observables(obs, tlist, 1, psi)
    assert isinstance(data, np.ndarray) and np.allclose(data, [0.0, 1.0])
    assert qp.init_storage(psi, tlist, obs).shape == (2, 5)
    assert abs(qp.map_observables((lambda s: float(np.linalg.norm(s)),), tlist, 1, psi) - 1.0) < 1e-15
    assert abs(qp.map_observable(lambda s, tl, i: tl[i - 1] * np.abs(s) ** 2, tlist, 5, psi)[0] - 0.5) < 1e-15
    # scalar data: Vector{T}(undef, nt) (src/storage.jl:44), element-wise writes
    vec = qp.init_storage(psi, tlist, (Z,))
    assert vec.shape == (5,) and vec.dtype == np.complex128
    qp.write_to_storage(vec, 2, 0.25 + 0j)
    assert qp.get_from_storage(vec, 2) == 0.25
    mixed = qp.map_observables((Z, lambda s: "label"), tlist, 1, psi)
    assert isinstance(mixed, tuple) and mixed[1] == "label"
    slots = qp.init_storage(mixed, 4)
    assert slots == [None] * 4
    qp.write_to_storage(slots, 2, mixed)
    assert qp.get_from_storage(slots, 2) is mixed
    import pytest

    with pytest.raises(TypeError, match="must take either"):
        qp.map_observable(lambda a, b: 0, tlist, 1, psi)


def test_shapes():
    """Mirror of the reference's Shapes module (src/shapes.jl; test/test_shapes.jl)."""
    from qprop_b200.shapes import blackman, box, flattop

    assert box(0.5, 0, 1) == 1.0 and box(1.5, 0, 1) == 0.0 and box(0.0, 0, 1) == 1.0
    assert abs(blackman(0.5, 0, 1) - 1.0) < 1e-15 and abs(blackman(0.0, 0, 1)) < 1e-15 and blackman(2.0, 0, 1) == 0.0
    kw = dict(T=10.0, t_rise=2.0)
    assert flattop(5.0, **kw) == 1.0 and flattop(-1.0, **kw) == 0.0 and flattop(11.0, **kw) == 0.0
    assert abs(flattop(0.0, **kw)) < 1e-15 and abs(flattop(10.0, **kw)) < 1e-15
    assert abs(flattop(1.0, **kw) - blackman(1.0, 0.0, 4.0)) < 1e-15          # half a Blackman window
    assert abs(flattop(1.0, func="sinsq", **kw) - 0.5) < 1e-15               # sin^2(pi/4)
    assert abs(flattop(9.5, t_fall=1.0, **kw) - blackman(9.5, 8.0, 10.0)) < 1e-15
    assert flattop(3.0, t0=2.0, **kw) < 1.0 and flattop(4.0, t0=2.0, **kw) == 1.0
    import pytest

    with pytest.raises(ValueError, match="Unknown func"):
        flattop(1.0, func="gauss", **kw)
    # usable as a control: check_control on a shaped pulse
    tlist = np.linspace(0, 10, 101)
    assert qp.check_control(lambda t: flattop(t, **kw), tlist)


def test_amplitudes():
    """Mirror of the reference's Amplitudes module (src/amplitudes.jl; test/test_amplitudes.jl):
    locked, shaped and guided amplitudes inside a Generator."""
    from qprop_b200.shapes import flattop

    tlist = np.linspace(0, 10, 11)

    def eps(t):
        return 0.5 * t

    def S(t):
        return flattop(t, T=10.0, t_rise=2.0)

    H0, H1, H2 = np.eye(2, dtype=complex), np.ones((2, 2), dtype=complex), np.diag([1.0, -1.0]).astype(complex)
    a1 = qp.ShapedAmplitude(eps, shape=S)
    a2 = qp.LockedAmplitude(S)
    G = qp.hamiltonian(H0, (H1, a1), (H2, a2))
    assert qp.get_controls(G) == (eps,)                                   # the locked amplitude has no control
    assert qp.check_amplitude(a1, tlist) and qp.check_amplitude(a2, tlist)
    op = qp.evaluate(G, tlist, 4)
    tm = qp.t_mid(tlist, 4)
    assert op.coeffs == [S(tm) * eps(tm), S(tm)]
    op = qp.evaluate(G, tlist, 4, vals_dict=qp.IdDict([(eps, 2.0)]))       # PWC value of the control
    assert op.coeffs == [S(tm) * 2.0, S(tm)]
    assert a1(3.0) == S(3.0) * eps(3.0) and qp.controls.evaluate(a1, 3.0) == a1(3.0)
    # discretised forms
    ad = qp.ShapedAmplitude(eps, tlist, shape=S)
    assert np.allclose(np.asarray(ad), qp.discretize_on_midpoints(eps, tlist) * qp.discretize_on_midpoints(S, tlist))
    assert qp.controls.evaluate(ad, tlist, 4) == np.asarray(ad)[3]
    assert qp.controls.evaluate(qp.LockedAmplitude(S, tlist), tlist, 2) == qp.discretize_on_midpoints(S, tlist)[1]
    # guided: a = G + S eps
    ag = qp.GuidedAmplitude(eps, guide=lambda t: 1.0 + t, shape=S)
    assert qp.controls.evaluate(ag, tlist, 4) == (1.0 + tm) + S(tm) * eps(tm)
    assert qp.get_controls(qp.hamiltonian(H0, (H1, ag))) == (eps,)
    # substitute replaces the control inside the amplitude
    def eps2(t):
        return 1.0

    G2 = qp.substitute(G, qp.IdDict([(eps, eps2)]))
    assert qp.get_controls(G2) == (eps2,) and G2.amplitudes[1] is a2
    import pytest

    with pytest.raises(ValueError, match="same length"):
        qp.ShapedAmplitude(np.zeros(3), shape=np.zeros(4))
    with pytest.raises(ValueError, match="callable"):
        qp.LockedAmplitude("x")
    with pytest.raises(ValueError, match="only be evaluated"):
        qp.controls.evaluate(qp.LockedAmplitude(np.zeros(10)), 0.5)
