"""GPU parity at BASELINE.json's FULL sizes through size-independent properties (the oracle
cannot run these sizes in seconds): forward followed by backward propagation returns the
initial state, linearity of the propagator, norm conservation to 1e-12 per step for Hermitian
generators, agreement between independent storage formats / kernels, and the first step against
SciPy's `expm_multiply` where that is affordable."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


@pytest.fixture(scope="module")
def tfim20(qp):
    return qp.workloads.config2_tfim(n_spins=20, nt=5, dt=0.1)


def _gen(qp, w):
    return qp.hamiltonian(w["ops"][0], *[(op, u) for op, u in zip(w["ops"][1:], w["controls"])])


def test_config2_full_size_properties(qp, ctx, tfim20):
    """TFIM n = 20 (N = 2^20), H0 + 2 PWC controls, Cheby: config 2 at full size."""
    w = tfim20
    G = _gen(qp, w)
    kw = dict(E_min=w["E_min"], E_max=w["E_max"], ctx=ctx)
    rng = np.random.default_rng(7)
    N = 1 << 20
    psi1 = w["psi0"]
    psi2 = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    psi2 /= np.linalg.norm(psi2)
    # forward then backward over the whole grid is the identity
    p = qp.init_prop(psi1, G, w["tlist"], "cheby", **kw)
    assert p.wrk.gen.format == "selld"
    norms = []
    while qp.prop_step(p) is not None:
        norms.append(p.state.norm())
    assert np.max(np.abs(np.array(norms) - 1)) < 1e-12        # norm conservation per step
    fwd1 = p.state.to_host()
    back = qp.propagate(fwd1, G, w["tlist"], "cheby", backward=True, **kw)
    assert rel(back, psi1) < RTOL
    # linearity: U(a psi1 + b psi2) = a U psi1 + b U psi2
    a, b = 0.6 - 0.3j, -0.2 + 0.7j
    fwd2 = qp.propagate(psi2, G, w["tlist"], "cheby", **kw)
    fwd12 = qp.propagate(a * psi1 + b * psi2, G, w["tlist"], "cheby", **kw)
    assert rel(fwd12, a * fwd1 + b * fwd2) < RTOL
    # unitarity: inner products are preserved
    assert abs(np.vdot(fwd1, fwd2) - np.vdot(psi1, psi2)) < 1e-11
    # an independent storage format and kernel (uncompressed SELL-32, TMA-staged) agrees
    other = qp.propagate(psi1, G, w["tlist"], "cheby", matrix_format="sell", **kw)
    assert rel(other, fwd1) < 1e-12


def test_config2_full_size_first_step_vs_expm_multiply(qp, ctx, tfim20):
    """One interval of config 2 at full size against SciPy's Krylov-free `expm_multiply`."""
    from scipy.sparse.linalg import expm_multiply

    import oracle as O

    w = tfim20
    tl = w["tlist"][:2]
    mid = O.get_tlist_midpoints(w["tlist"])[0]
    H = (w["ops"][0] + w["controls"][0](mid) * w["ops"][1] + w["controls"][1](mid) * w["ops"][2]).tocsc()
    ref = expm_multiply(-1j * (tl[1] - tl[0]) * H, w["psi0"])
    G = _gen(qp, w)
    p = qp.init_prop(w["psi0"], G, w["tlist"], "cheby", E_min=w["E_min"], E_max=w["E_max"], ctx=ctx)
    out = qp.prop_step(p).to_host()
    assert rel(out, ref) < 1e-9  # expm_multiply's own truncation is ~1e-10 here


def test_config3_full_size_properties(qp, ctx):
    """Transmon chain N = 2^16 with per-trajectory amplitudes (config 3 shape, B = 256 of the
    1024 trajectories): batched propagation equals the single-state propagation of each sampled
    trajectory, forward-backward is the identity, norms are conserved."""
    from qprop_b200.ensemble import EnsembleChebyPropagator

    B = 256
    w = qp.workloads.config3_transmon(n_sites=8, levels=4, B=B, nt=4, dt=0.5)
    H0, H1, H2 = w["ops"]
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
    rng = np.random.default_rng(3)
    N = H0.shape[0]
    psi0 = rng.standard_normal((N, B)) + 1j * rng.standard_normal((N, B))  # a different state per trajectory
    psi0 /= np.linalg.norm(psi0, axis=0)
    ens = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], psi0, w["tlist"], -bound, bound, ctx)
    n_steps = len(w["tlist"]) - 1
    for _ in range(n_steps):
        ens.prop_step()
        assert np.max(np.abs(np.asarray(ens.state.norm()) - 1)) < 1e-12
    out = ens.state.to_host()
    # sampled trajectories through the B = 1 path (different kernel: thread per row)
    for b in (0, 100, B - 1):
        s = w["scales"][b]
        Gb = qp.hamiltonian(H0, (H1, lambda t, s=s: s * w["controls"][0](t)), (H2, lambda t, s=s: s * w["controls"][1](t)))
        single = qp.propagate(psi0[:, b].copy(), Gb, w["tlist"], "cheby", E_min=-bound, E_max=bound, ctx=ctx)
        assert rel(out[:, b], single) < RTOL


def test_config5_full_size_dense_properties(qp, ctx):
    """Dense optomechanics generator N = 8192 (config 5): the batched FP64 tensor-core path (B = 16)
    against the B = 1 GEMV path column by column, forward-backward identity, norm conservation."""
    H = qp.workloads.config5_optomech_dense()
    N = H.shape[0]
    ev = float(np.abs(H).sum(axis=1).max())
    gen = qp.DeviceGenerator(ctx, [H], 0)
    rng = np.random.default_rng(9)
    B = 16
    psi = rng.standard_normal((N, B)) + 1j * rng.standard_normal((N, B))
    psi /= np.linalg.norm(psi, axis=0)
    dt = 8.0 / ev
    st = qp.DeviceState.from_host(ctx, psi)
    wrk = qp.ChebyWrk(st, gen, 2 * ev, -ev, dt)
    qp.cheby_(st, None, dt, wrk, coeffs=[])
    assert np.max(np.abs(st.norm() - 1)) < 1e-12
    fwd = st.to_host()
    for b in (0, B - 1):
        s1 = qp.DeviceState.from_host(ctx, psi[:, b].copy())
        w1 = qp.ChebyWrk(s1, gen, 2 * ev, -ev, dt)
        qp.cheby_(s1, None, dt, w1, coeffs=[])
        assert rel(fwd[:, b], s1.to_host()) < RTOL
    qp.cheby_(st, None, -dt, wrk, coeffs=[])
    assert rel(st.to_host(), psi) < RTOL
