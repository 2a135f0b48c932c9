#!/usr/bin/env python
"""BASELINE config 3 on 1..8 GPUs: an ensemble of B trajectories (per-trajectory control
amplitudes) of the N = 4^8 transmon chain, trajectory-sharded over the ranks (one process per
GPU, no data-path collective), NCCL only for the barrier, the max-over-ranks timing and the final
gather of one expectation value per trajectory.  Launch with torchrun for N > 1:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_ensemble.py --B 1024 --steps 5

One JSON line on rank 0 ("scaling": "strong": the total number of trajectories is fixed)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qprop_b200 as qp  # noqa: E402
from qprop_b200.ensemble import EnsembleChebyPropagator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--sites", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = qp.Context(local)
    w = qp.workloads.config3_transmon(n_sites=args.sites, levels=4, B=args.B, nt=args.steps + args.warmup + 3, dt=0.5)
    H0, H1, H2 = w["ops"]
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
    ens = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], w["psi0"], w["tlist"], -bound, bound, ctx,
                                  rank=rank, world=world)
    for _ in range(args.warmup):
        ens.prop_step()
    ctx.sync()
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        ens.prop_step()
    e1.record(stream)
    ctx.sync()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # the ensemble's only data collective: per-trajectory populations of |0...0>, gathered at the end
    local_state = ens.state.to_host().reshape(ens.state.n, -1)
    pops = ens.gather_expvals(np.abs(local_state[0]) ** 2)
    norms = ens.gather_expvals(np.linalg.norm(local_state, axis=0))
    if rank == 0:
        N = H0.shape[0]
        n_c = ens.wrk.n_coeffs
        t = float(ms[0]) * 1e-3
        term_bytes_rank = ens.gen.matrix_bytes + 80 * N * ens.B_local
        print(json.dumps({
            "metric": "cheby_trajectory_steps_per_s", "value": args.B * args.steps / t, "unit": "trajectory-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "scaling": "strong", "config": {"workload": f"transmon chain {args.sites}x4 N={N}, B={args.B} trajectories, "
                                            f"B_local={ens.B_local}, n_coeffs={n_c}"},
            "effective_hbm_gbs_per_gpu": term_bytes_rank * (n_c - 1) * args.steps / t / 1e9,
            "gathered": int(pops.shape[-1]), "norm_dev_max": float(np.max(np.abs(norms - 1))),
        }), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
