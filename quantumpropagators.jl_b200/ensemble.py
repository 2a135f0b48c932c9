"""Trajectory-sharded ensembles.

An ensemble is B independent trajectories of ONE system that differ in their control
amplitudes (trajectory b scales control l by ``scales[l][b]``).  Trajectories never interact,
so the ensemble shards by contiguous blocks over the ranks (one process per GPU): every rank
holds the full operators and a ``[N][B_local]`` batched state, steps it with the batched
Chebyshev kernel (per-trajectory coefficients, one shared coefficient table derived from
control ranges that cover the whole ensemble), and NO communication happens inside the time
loop.  The only collectives are the final gathers of expectation values and states
(``torch.distributed``: NCCL on GPU tensors that wrap the state's device memory, gloo on CPU
tensors in the host-logic tests).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .controls import discretize_on_midpoints

__all__ = ["shard_range", "trajectory_coefficients", "gather_blocks", "EnsembleChebyPropagator", "LibraryEnsemble"]


class LibraryEnsemble:
    """Handle of a ``qp_ens_t``: the ensemble communicator INSIDE ``libqprop_b200.so`` (NCCL over
    NVLink, or device copies for fake ranks) -- what a Julia host uses, no torch.distributed.

    ``LibraryEnsemble.local(devices)``: single process; one context per entry of ``devices`` (all
    distinct: NCCL; all equal: "fake ranks" on one device).
    ``LibraryEnsemble.from_rank(ctx, rank, world, id_bytes)``: one process per GPU; ``id_bytes`` are
    the 128 bytes rank 0 got from ``LibraryEnsemble.unique_id()`` and handed to the other processes.
    """

    def __init__(self, handle, lib, contexts, ranks):
        import weakref

        self.handle = handle
        self._lib = lib
        self.contexts = contexts  # Context objects of the local members
        self.ranks = ranks        # their world ranks
        self._finalizer = weakref.finalize(self, lib.qp_ens_destroy, handle)
        n, nl, tr = C.c_int32(), C.c_int32(), C.c_int32()
        L.check(lib.qp_ens_info(handle, C.byref(n), C.byref(nl), C.byref(tr)), None)
        self.world, self.n_local, self.transport = n.value, nl.value, ("nccl" if tr.value else "device copies")

    @staticmethod
    def unique_id() -> bytes:
        lib = L.load()
        buf = (C.c_uint8 * 128)()
        L.check(lib.qp_ens_unique_id(buf), None)
        return bytes(buf)

    @classmethod
    def local(cls, devices):
        from .device import Context

        lib = L.load()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        L.check(lib.qp_ens_create(devs, len(devices), C.byref(h)), None)
        ctxs, ranks = [], []
        for i in range(len(devices)):
            ch, rk = C.c_void_p(), C.c_int32()
            L.check(lib.qp_ens_ctx(h, i, C.byref(ch), C.byref(rk)), None)
            ctxs.append(Context.borrowed(ch, int(devices[i])))
            ranks.append(rk.value)
        return cls(h, lib, ctxs, ranks)

    @classmethod
    def from_rank(cls, ctx, rank, world, id_bytes=None):
        lib = ctx._lib
        h = C.c_void_p()
        buf = (C.c_uint8 * 128)(*id_bytes) if id_bytes is not None else None
        L.check(lib.qp_ens_create_rank(ctx.handle, int(rank), int(world), buf, C.byref(h)), ctx.handle)
        return cls(h, lib, [ctx], [int(rank)])

    def shard(self, n_total, rank):
        b0, b1 = C.c_int64(), C.c_int64()
        L.check(self._lib.qp_ens_shard(int(n_total), int(rank), self.world, C.byref(b0), C.byref(b1)), None)
        return b0.value, b1.value

    def gather_states(self, local_states, n_total, to_host=True, full_states=None):
        """Final states of the whole ensemble.  ``local_states``: one DeviceState per local member.
        Returns the (N, n_total) host array (``to_host``) and / or fills ``full_states`` (one
        (N, n_total) DeviceState per local member)."""
        arr = (C.c_void_p * len(local_states))(*[s.handle for s in local_states])
        full = (C.c_void_p * len(full_states))(*[s.handle for s in full_states]) if full_states else None
        out = np.empty((local_states[0].n, int(n_total)), dtype=np.complex128) if to_host else None
        L.check(self._lib.qp_ens_gather_states(self.handle, arr, int(n_total), full, L.ptr(out) if to_host else None),
                self.contexts[0].handle)
        return out

    def gather_expvals(self, local_values, n_total):
        """``local_values``: one array (..., B_local) per local member; returns (..., n_total)."""
        vals = [np.ascontiguousarray(v, dtype=np.complex128) for v in local_values]
        lead = vals[0].shape[:-1]
        n_values = int(np.prod(lead)) if lead else 1
        flat = [v.reshape(n_values, v.shape[-1]) for v in vals]
        ptrs = (C.c_void_p * len(flat))(*[v.ctypes.data for v in flat])
        out = np.empty((n_values, int(n_total)), dtype=np.complex128)
        L.check(self._lib.qp_ens_gather_expvals(self.handle, ptrs, n_values, int(n_total), L.ptr(out)), self.contexts[0].handle)
        return out.reshape(lead + (int(n_total),))


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block [b0, b1) of trajectories owned by ``rank``; sizes differ by at most 1."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n_total, world)
    b0 = rank * base + min(rank, extra)
    return b0, b0 + base + (1 if rank < extra else 0)


def trajectory_coefficients(controls, scales, tlist):
    """Per-interval, per-trajectory operator coefficients ``u[n, l, b] = scales[l, b] * ε_l(t_n)``
    with ε_l on the interval midpoints (reference ``src/pwc_utils.jl:29-45``).  ``scales`` is
    (B,) (same scale for every control) or (L, B)."""
    mid = np.stack([discretize_on_midpoints(c, tlist) for c in controls])  # (L, nt-1)
    scales = np.asarray(scales, dtype=np.float64)
    if scales.ndim == 1:
        scales = np.broadcast_to(scales, (len(controls), scales.shape[0]))
    return np.ascontiguousarray(np.einsum("ln,lb->nlb", mid, scales))


def gather_blocks(local, counts, group=None):
    """All-gather of per-rank blocks whose LAST axis is the trajectory axis (ragged: rank r
    contributes ``counts[r]`` trajectories).  ``local`` is a torch tensor (CUDA -> NCCL,
    CPU -> gloo) or a NumPy array (converted to a CPU tensor).  Returns the concatenation along
    the last axis, on every rank, in the type it was given."""
    import torch
    import torch.distributed as dist

    as_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if as_numpy else local.contiguous()
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if len(counts) != world:
        raise ValueError("counts must have one entry per rank")
    # complex tensors travel as pairs of reals (NCCL has no complex dtype)
    is_complex = t.is_complex()
    if is_complex:
        t = torch.view_as_real(t)
        lead = t.shape[:-2]
        t = t.movedim(-2, 0).contiguous()  # trajectory axis first: ragged gather = concatenation
    else:
        lead = t.shape[:-1]
        t = t.movedim(-1, 0).contiguous()
    cmax = max(counts)
    pad = torch.zeros((cmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = torch.empty((world * cmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = [out[r * cmax : r * cmax + counts[r]] for r in range(world)]
    full = torch.cat(pieces, dim=0)
    if is_complex:
        full = torch.view_as_complex(full.movedim(0, len(lead)).contiguous())
    else:
        full = full.movedim(0, len(lead)).contiguous()
    return full.numpy() if as_numpy else full


class EnsembleChebyPropagator:
    """Chebyshev propagation of this rank's block of an ensemble on one GPU.

    Parameters: ``ops`` = [H0, H1, ..., HL] (drift first), ``controls`` = L control functions /
    vectors, ``scales`` = (B_total,) or (L, B_total) amplitude scales, ``psi0`` = (N,) shared
    initial state or (N, B_total), spectral range [E_min, E_max] valid for EVERY trajectory
    (the reference's ``control_ranges`` hook, ``src/cheby_propagator.jl:59-66``).
    """

    def __init__(self, ops, controls, scales, psi0, tlist, E_min, E_max, ctx, rank=0, world=1,
                 specrange_buffer=0.01, limit=1e-12, matrix_format="auto", library_ensemble=None):
        from .cheby import ChebyWrk
        from .device import DeviceGenerator, DeviceState

        self.ctx = ctx
        self.tlist = np.asarray(tlist, dtype=np.float64)
        scales = np.asarray(scales, dtype=np.float64)
        self.B_total = scales.shape[-1]
        self.rank, self.world = rank, world
        # gathers through the communicator inside libqprop_b200.so (qp_ens_*) when one is given,
        # else through torch.distributed (kept as a cross-check of the library path)
        self.lib_ens = library_ensemble
        self.gather_backend = (f"libqprop_b200 qp_ens_* ({library_ensemble.transport})" if library_ensemble is not None
                               else "torch.distributed")
        self.counts = [shard_range(self.B_total, r, world)[1] - shard_range(self.B_total, r, world)[0] for r in range(world)]
        self.b0, self.b1 = shard_range(self.B_total, rank, world)
        self.B_local = self.b1 - self.b0
        if self.B_local < 1:
            raise ValueError("every rank needs at least one trajectory")
        self.coeffs = trajectory_coefficients(controls, scales[..., self.b0 : self.b1], self.tlist)
        psi0 = np.asarray(psi0, dtype=np.complex128)
        local = np.repeat(psi0[:, None], self.B_local, axis=1) if psi0.ndim == 1 else psi0[:, self.b0 : self.b1]
        self.state = DeviceState.from_host(ctx, np.ascontiguousarray(local))
        self.gen = DeviceGenerator(ctx, ops, len(controls), matrix_format)
        Delta = float(E_max - E_min)
        delta = specrange_buffer * Delta
        dt = float(self.tlist[1] - self.tlist[0])
        self.wrk = ChebyWrk(self.state, self.gen, Delta + delta, E_min - delta / 2, dt, limit=limit)
        self.n = 1

    def prop_step(self):
        """One ``prop_step!`` of all local trajectories (one batched ``qp_cheby_step``)."""
        from .cheby import cheby_

        if not 0 < self.n < len(self.tlist):
            return None
        cheby_(self.state, None, self.wrk.dt, self.wrk, coeffs=self.coeffs[self.n - 1], per_trajectory=True)
        self.n += 1
        return self.state

    def propagate(self):
        while self.prop_step() is not None:
            pass
        return self.state

    def gather_states(self, group=None):
        """Final states of the WHOLE ensemble, (N, B_total), on every rank."""
        import torch

        if self.lib_ens is not None:
            return self.lib_ens.gather_states([self.state], self.B_total)
        if self.world == 1:
            return self.state.to_host().reshape(self.state.n, -1)
        self.ctx.sync()
        local = torch.as_tensor(self.state, device=f"cuda:{self.ctx.device}").reshape(self.state.n, self.B_local)
        return gather_blocks(local, self.counts, group).cpu().numpy()

    def gather_expvals(self, values, group=None):
        """Gather per-trajectory numbers (last axis = local trajectories) from all ranks."""
        import torch

        values = np.asarray(values)
        if self.lib_ens is not None:
            out = self.lib_ens.gather_expvals([values], self.B_total)
            return out if np.iscomplexobj(values) else out.real
        if self.world == 1:
            return values
        t = torch.from_numpy(np.ascontiguousarray(values)).to(f"cuda:{self.ctx.device}")
        return gather_blocks(t, self.counts, group).cpu().numpy()
