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


def test_ensemble_two_ranks_nccl(qp, ctx):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "ensemble_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "ENSEMBLE_OK" in res.stdout
