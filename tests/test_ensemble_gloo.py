"""Host-side logic of the trajectory-sharded ensemble on CPU: world_size-2
gloo run of the shard/coefficient/gather plumbing, with the oracle standing in for the device
stepping.  The GPU path (same plumbing over NCCL) is covered by tests/test_gpu_ensemble.py."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
import qprop_b200 as qp
from qprop_b200.ensemble import gather_blocks, shard_range, trajectory_coefficients


def test_shard_range_covers_everything():
    for n, world in [(1024, 8), (5, 2), (7, 3), (3, 3), (10, 4)]:
        blocks = [shard_range(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_library_shard_matches_host_shard():
    """qp_ens_shard (C ABI, no device needed) cuts the ensemble exactly like the host mirror."""
    import ctypes as C

    from qprop_b200 import _lib

    lib = _lib.load()
    for n, world in [(1024, 8), (5, 2), (7, 3), (3, 3), (10, 4), (0, 2)]:
        for r in range(world):
            b0, b1 = C.c_int64(), C.c_int64()
            assert lib.qp_ens_shard(n, r, world, C.byref(b0), C.byref(b1)) == 0
            if n >= world:
                assert (b0.value, b1.value) == shard_range(n, r, world)
    b0, b1 = C.c_int64(), C.c_int64()
    assert lib.qp_ens_shard(4, 2, 2, C.byref(b0), C.byref(b1)) == _lib.QP_ERR_INVALID_ARG


def test_trajectory_coefficients():
    tlist = np.linspace(0, 1, 6)
    u1 = lambda t: 1.0 + t  # noqa: E731
    u2 = np.arange(5, dtype=float)
    c = trajectory_coefficients([u1, u2], np.array([0.5, 1.0, 2.0]), tlist)
    assert c.shape == (5, 2, 3)
    assert np.allclose(c[:, 0, 2], 2.0 * qp.discretize_on_midpoints(u1, tlist))
    assert np.allclose(c[:, 1, 0], 0.5 * u2)
    c2 = trajectory_coefficients([u1, u2], np.array([[1, 2, 3], [4, 5, 6]]), tlist)
    assert np.allclose(c2[3, 1], u2[3] * np.array([4, 5, 6]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = qp.workloads.config3_transmon(n_sites=2, levels=3, B=B_total, nt=6, dt=0.5)
        b0, b1 = shard_range(B_total, rank, world)
        counts = [shard_range(B_total, r, world)[1] - shard_range(B_total, r, world)[0] for r in range(world)]
        coeffs = trajectory_coefficients(w["controls"], w["scales"][b0:b1], w["tlist"])
        # oracle stands in for the device: propagate this rank's trajectories
        H0, H1, H2 = w["ops"]
        bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
        dt = w["tlist"][1] - w["tlist"][0]
        wrk = O.ChebyWrk(w["psi0"], 2.02 * bound, -1.01 * bound, dt)
        local = np.zeros((H0.shape[0], b1 - b0), dtype=complex)
        for j in range(b1 - b0):
            psi = w["psi0"].copy()
            for n in range(len(w["tlist"]) - 1):
                O.cheby_inplace(psi, O.Operator([H0, H1, H2], list(coeffs[n, :, j])), dt, wrk)
            local[:, j] = psi
        full = gather_blocks(local, counts)                       # complex, ragged (3 + 2)
        pops = gather_blocks(np.abs(local[0]) ** 2, counts)       # real 1-D expectation values
        stacked = gather_blocks(torch.from_numpy(np.stack([local.real, local.imag])), counts)  # 3-D real
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), full=full, pops=pops, stacked=stacked.numpy())
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_matches_single_process(tmp_path):
    B_total, world = 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B_total, str(tmp_path)), nprocs=world, join=True)
    # single-process reference
    w = qp.workloads.config3_transmon(n_sites=2, levels=3, B=B_total, nt=6, dt=0.5)
    coeffs = trajectory_coefficients(w["controls"], w["scales"], w["tlist"])
    H0, H1, H2 = w["ops"]
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
    dt = w["tlist"][1] - w["tlist"][0]
    wrk = O.ChebyWrk(w["psi0"], 2.02 * bound, -1.01 * bound, dt)
    ref = np.zeros((H0.shape[0], B_total), dtype=complex)
    for b in range(B_total):
        psi = w["psi0"].copy()
        for n in range(len(w["tlist"]) - 1):
            O.cheby_inplace(psi, O.Operator([H0, H1, H2], list(coeffs[n, :, b])), dt, wrk)
        ref[:, b] = psi
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert got["full"].shape == ref.shape
        assert np.array_equal(got["full"], ref)          # every rank holds the whole ensemble
        assert np.array_equal(got["pops"], np.abs(ref[0]) ** 2)
        assert np.array_equal(got["stacked"], np.stack([ref.real, ref.imag]))
    # trajectories really differ (different control scales)
    assert np.linalg.norm(ref[:, 0] - ref[:, -1]) > 1e-3
