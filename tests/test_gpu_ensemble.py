"""GPU tests of the trajectory-sharded ensemble (BASELINE config 3 shape, reduced): the batched
Chebyshev kernel with per-trajectory coefficients against the oracle, and -- when the box has
at least 2 GPUs -- a real 2-rank NCCL run (torchrun) whose gathered states must equal the
single-GPU result."""

import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
from qprop_b200.ensemble import EnsembleChebyPropagator, trajectory_coefficients

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _envelope(ops):
    H0, H1, H2 = ops
    return float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())


@pytest.mark.parametrize("B", [1, 3, 8, 33])
def test_ensemble_single_gpu_vs_oracle(qp, ctx, B):
    w = qp.workloads.config3_transmon(n_sites=4, levels=3, B=B, nt=8, dt=0.5)
    bound = _envelope(w["ops"])
    ens = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], w["psi0"], w["tlist"], -bound, bound, ctx)
    ens.propagate()
    out = ens.gather_states()
    assert out.shape == (81, B)
    coeffs = trajectory_coefficients(w["controls"], w["scales"], w["tlist"])
    wrk = O.ChebyWrk(w["psi0"], ens.wrk.Delta, ens.wrk.E_min, ens.wrk.dt)
    assert wrk.n_coeffs == ens.wrk.n_coeffs
    for b in range(B):
        psi = w["psi0"].copy()
        for n in range(len(w["tlist"]) - 1):
            O.cheby_inplace(psi, O.Operator(list(w["ops"]), list(coeffs[n, :, b])), ens.wrk.dt, wrk)
        assert np.linalg.norm(out[:, b] - psi) / np.linalg.norm(psi) < 1e-10
        assert abs(np.linalg.norm(out[:, b]) - 1) < 1e-12
    assert ens.prop_step() is None  # past the end of the grid


@pytest.mark.parametrize("n_fake,B", [(3, 10), (2, 64), (4, 5), (1, 7)])
def test_library_ensemble_fake_ranks(qp, n_fake, B):
    """The ensemble communicator inside libqprop_b200.so (qp_ens_*) with several "fake ranks" on ONE
    device: ragged trajectory blocks, gather of final states (to the host and into device-resident
    full states) and of per-trajectory numbers, against the whole ensemble on a single rank."""
    from qprop_b200.ensemble import LibraryEnsemble

    ens_comm = LibraryEnsemble.local([0] * n_fake)
    assert ens_comm.world == n_fake and ens_comm.n_local == n_fake and ens_comm.transport == "device copies"
    w = qp.workloads.config3_transmon(n_sites=4, levels=4, B=B, nt=5, dt=0.5)
    bound = _envelope(w["ops"])
    rng = np.random.default_rng(B)
    N = w["ops"][0].shape[0]
    psi0 = rng.standard_normal((N, B)) + 1j * rng.standard_normal((N, B))
    psi0 /= np.linalg.norm(psi0, axis=0)
    members = [EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], psi0, w["tlist"], -bound, bound,
                                       ens_comm.contexts[i], rank=ens_comm.ranks[i], world=n_fake)
               for i in range(n_fake)]
    assert [m.B_local for m in members] == [ens_comm.shard(B, r)[1] - ens_comm.shard(B, r)[0] for r in range(n_fake)]
    for m in members:
        m.propagate()
    states = ens_comm.gather_states([m.state for m in members], B)
    ref = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], psi0, w["tlist"], -bound, bound, ens_comm.contexts[0])
    ref.propagate()
    full = ref.state.to_host().reshape(N, B)
    assert states.shape == (N, B) and np.linalg.norm(states - full) / np.linalg.norm(full) < 1e-13
    # device-resident gather: every member receives the whole ensemble
    fulls = [qp.DeviceState(ens_comm.contexts[i], N, B) for i in range(n_fake)]
    ens_comm.gather_states([m.state for m in members], B, to_host=False, full_states=fulls)
    for f in fulls:
        assert np.array_equal(f.to_host().reshape(N, B), states)
    # per-trajectory numbers: populations of |0> and norms, two values per trajectory
    vals = [np.stack([np.abs(m.state.to_host().reshape(N, -1)[0]) ** 2, np.asarray(m.state.norm()).reshape(-1)]) for m in members]
    got = ens_comm.gather_expvals(vals, B)
    assert got.shape == (2, B)
    assert np.allclose(got[0].real, np.abs(full[0]) ** 2, atol=1e-15) and np.allclose(got[1].real, 1.0, atol=1e-12)
    # the propagator-level route (library_ensemble=...) on a single rank
    solo = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], psi0, w["tlist"], -bound, bound, ens_comm.contexts[0],
                                   library_ensemble=LibraryEnsemble.from_rank(ens_comm.contexts[0], 0, 1))
    solo.propagate()
    assert np.array_equal(solo.gather_states(), full)
    assert np.allclose(solo.gather_expvals(np.asarray(solo.state.norm())), 1.0, atol=1e-12)


def test_ensemble_two_ranks_nccl(qp, ctx):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "ensemble_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "ENSEMBLE_OK" in res.stdout
