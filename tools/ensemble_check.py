#!/usr/bin/env python
"""Run under torchrun on N GPUs: propagates a trajectory-sharded ensemble (config-3 shape,
reduced), gathers final states and expectation values over NCCL, and checks them on rank 0
against a single-GPU run of the whole ensemble.  Prints ENSEMBLE_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qprop_b200 as qp  # noqa: E402
from qprop_b200.ensemble import EnsembleChebyPropagator, LibraryEnsemble  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = qp.Context(local)
    B = 13  # ragged over the ranks on purpose
    w = qp.workloads.config3_transmon(n_sites=4, levels=4, B=B, nt=11, dt=0.5)
    H0, H1, H2 = w["ops"]
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
    ens = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], w["psi0"], w["tlist"], -bound, bound, ctx,
                                  rank=rank, world=world)
    ens.propagate()
    states = ens.gather_states()
    pops = ens.gather_expvals(np.abs(ens.state.to_host().reshape(ens.state.n, -1)[0]) ** 2)
    # the same gathers through the communicator INSIDE libqprop_b200.so (qp_ens_*, NCCL): rank 0 obtains
    # the id, the launcher's process group only carries those 128 bytes
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident = torch.tensor(list(LibraryEnsemble.unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(ident, 0)
    comm = LibraryEnsemble.from_rank(ctx, rank, world, bytes(ident.cpu().tolist()))
    assert comm.transport == "nccl" and comm.world == world
    states_lib = comm.gather_states([ens.state], B)
    pops_lib = comm.gather_expvals([np.abs(ens.state.to_host().reshape(ens.state.n, -1)[0]) ** 2], B).real
    ok = np.array_equal(states_lib, states) and np.array_equal(pops_lib, pops)
    if rank == 0:
        ref = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], w["psi0"], w["tlist"], -bound, bound, ctx)
        ref.propagate()
        full = ref.gather_states()
        err = np.linalg.norm(states - full) / np.linalg.norm(full)
        ok = ok and states.shape == (256, B) and err < 1e-13 and np.allclose(pops, np.abs(full[0]) ** 2, atol=1e-15)
        print(f"world={world} gathered {states.shape}, rel.err vs single GPU {err:.2e}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("ENSEMBLE_OK" if flag.item() == 1 else "ENSEMBLE_FAILED")
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
