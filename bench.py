#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: Chebyshev ``prop_step!``/s (and the fraction
of the HBM roofline) for the transverse-field Ising chain N = 2^20 with two PWC controls
(BASELINE configs[1]) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--format auto]

N > 1 is launched by torchrun (one rank per GPU).  A single large state stays on one GPU
(north_star), so for config 2 the ranks hold independent trajectories of the same system (each
with its own control scale) -- trajectory sharding, no data-path collective ("scaling": "weak").
NCCL carries the barrier, the max-over-ranks timing and the final gathers.

What one run measures (rank 0 prints ONE JSON line):

  value / ms_per_step   config 2, device-resident: blocks of K prop_step! (CUDA events on the
                        library's stream), repeated until >= --min-seconds of device time; the
                        MEDIAN block is reported (max over ranks), min / max ride along
  roofline              the dominant kernel (one fused Chebyshev term per launch): bytes the timed
                        storage format has to stream per launch / average launch time / measured
                        HBM peak.  The canonical-CSR figure of SURVEY.md 8d is reported separately
                        (`effective_vs_canonical`) -- a compressed format moves fewer bytes than that
  arms.sell_tma         the same workload on the uncompressed SELL-32 format (TMA-staged kernel):
                        matrix 474 MB per term > L2, the HBM-saturating arm
  arms.selld            the same workload on the generic compressed sparse format (what a generator
                        without the diagonal + bit-flip structure gets)
  e2e                   the same metric through the public API with HOST-resident states: every
                        step copies its state from pinned host memory and back (two trajectories
                        per GPU interleaved on two contexts so that the copies of one overlap the
                        step of the other; the strictly sequential single-trajectory figure rides
                        along)
  ensemble              BASELINE configs[2]: 1024 trajectories of the N = 2^16 transmon chain,
                        trajectory-sharded over the N ranks ("strong"), trajectory-steps/s,
                        per-GPU roofline fraction, final gather (NCCL inside libqprop_b200.so)
  cpu_baseline          N = 1 only: oracle/cheby_ref.c on the host cores + `parity_rel_err` of the
                        GPU state against it after the same steps
"""

import argparse
import json
import math
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cheby_prop_steps_per_s"
UNIT = "prop_step!/s"
L2_MB = 126.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-spins", type=int, default=20)
    ap.add_argument("--format", default="auto", choices=["auto", "csr", "sell", "selld", "bitflip"])
    ap.add_argument("--min-seconds", type=float, default=1.0, help="device time to accumulate per timed arm")
    ap.add_argument("--cpu-sample-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sell-arm", action="store_true")
    ap.add_argument("--no-ensemble", action="store_true")
    ap.add_argument("--ensemble-B", type=int, default=1024)
    ap.add_argument("--ensemble-steps", type=int, default=5)
    ap.add_argument("--e2e-lanes", type=int, default=3, help="host-resident trajectories interleaved per GPU in the e2e arm")
    return ap.parse_args()


def workload_name(n_spins, n_c):
    return f"TFIM chain n={n_spins} (N=2^{n_spins}), H0 + 2 PWC controls, Cheby prop_step!, n_coeffs={n_c}, B=1 per GPU"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(fmt):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any
    (profiles/traffic.json; a profiler cannot run inside a timed bench)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        return d.get(fmt), d.get("_source_" + fmt, "profiles/traffic.json") + " -- an ncu capture, not measured in this run"
    except Exception:
        return None, None


def bind_to_gpu_numa(local_rank):
    """Best effort: run this rank (and therefore first-touch / pin its host buffers) on the CPU
    cores of the NUMA node the GPU hangs off.  Returns a description for the JSON line."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "numa node unknown (single node or virtualised)"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"numa node {node}: none of its cpus allowed"
        os.sched_setaffinity(0, allowed)
        return f"numa node {node}, {len(allowed)} cpus"
    except Exception as exc:  # noqa: BLE001
        return f"not bound ({type(exc).__name__})"


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, flag in zip(names, p[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": mx,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def build_workload(n_spins, rank, world):
    import qprop_b200 as qp

    w = qp.workloads.config2_tfim(n_spins=n_spins, nt=101, dt=0.1)
    # trajectory `rank` of the ensemble: its own control scale (varied amplitudes, config 3 style)
    scale = 1.0 if world == 1 else 0.5 + 0.5 * rank / (world - 1)
    u1, u2 = w["controls"]
    w["controls"] = [lambda t, f=u1: scale * f(t), lambda t, f=u2: scale * f(t)]
    return w


# ------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ------------------------------------------------------------------------------------------


def oracle_steps(w, n_steps, n_warm=1, single_steps=1):
    """Time `n_steps` prop_step! of the CPU port on this workload after `n_warm` warm-up steps.

    (a) `omp`: the step with a row-parallel OpenMP CSR SpMV on all host threads
    (cheby_step_csr_omp) -- a stronger baseline than the reference has; (b) `single`: the faithful
    restatement of the reference's CPU path -- per-operator CSC scatter SpMV + level-1 passes,
    Int64 indices, ONE thread, what Julia's SparseArrays `mul!` does (cheby_step_csc), timed over
    `single_steps` steps only (it takes seconds per step).  Returns the state after the
    n_warm + n_steps steps of (a) for the parity check."""
    import oracle as O
    from oracle import cref

    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    p = O.init_prop(w["psi0"], O.hamiltonian(*terms), w["tlist"], "cheby", E_min=w["E_min"], E_max=w["E_max"])
    n_c = p.wrk.n_coeffs
    n_grid = len(w["tlist"]) - 1
    if not cref.available():
        for _ in range(n_warm):
            O.prop_step(p)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            O.prop_step(p)
        sec = (time.perf_counter() - t0) / n_steps
        return {"n_coeffs": n_c, "single": sec, "omp": None, "threads": 1, "final": np.array(p.state), "steps_done": n_warm + n_steps,
                "impl": "numpy/scipy oracle (scipy CSR @, 1 thread)"}
    ref = cref.ChebyRef(w["ops"], len(w["controls"]))
    wrk = p.wrk
    out = {"n_coeffs": n_c, "impl": "oracle/cheby_ref.c"}
    for key, threads, warm, steps in (("omp", cref.max_threads(), n_warm, n_steps), ("single", 0, 1, single_steps)):
        psi = w["psi0"].copy()
        secs = []
        for i in range(warm + steps):
            n = i % n_grid + 1  # interval index (wraps: only the amplitudes depend on it)
            coeffs = [complex(p.parameters[c][n - 1]) for c in p.controls]
            t0 = time.perf_counter()
            ref.step(psi, coeffs, wrk.coeffs, wrk.Delta, wrk.E_min, wrk.dt, threads=threads)
            secs.append(time.perf_counter() - t0)
        out[key] = sum(secs[warm:]) / steps
        if key == "omp":
            out["final"] = psi
            out["steps_done"] = warm + steps
    out["threads"] = cref.max_threads()
    return out


def cpu_numbers(r, n_steps, n_warm):
    best = r["omp"] if r.get("omp") else r["single"]
    cores = r["threads"] if r.get("omp") else 1
    return {
        "value": 1.0 / best, "unit": UNIT, "cores": cores, "kind": "port",
        "single_thread_value": 1.0 / r["single"],
        "sample": f"{n_steps} prop_step! of the same workload after {n_warm} warm-up step(s); {r['impl']}: value = row-parallel "
                  f"OpenMP CSR variant on {cores} threads; single_thread_value = faithful restatement of the reference's "
                  "single-threaded CSC mul! + level-1 passes (1 step after 1 warm-up).  A port, not Julia (absent from this image).",
    }


def find_julia():
    """probe for a Julia install (PATH, then the driver-reserved baseline/_ref)."""
    cands = [shutil.which("julia")]
    for sub in ("bin/julia", "julia/bin/julia"):
        cands.append(os.path.join(ROOT, "baseline", "_ref", sub))
    for c in cands:
        if c and os.path.exists(c) and os.access(c, os.X_OK):
            return c
    return None


def run_julia_reference(julia, args):
    """Time the REAL package (JULIA_NUM_THREADS=1, BLAS threads = nproc) with julia/bench_reference.jl.
    Returns the parsed JSON dict or None (any failure falls back to the port, and says so)."""
    env = dict(os.environ, JULIA_NUM_THREADS="1", JULIA_LOAD_PATH=f"{os.path.join(ROOT, 'baseline', '_ref')}:@:@stdlib",
               JULIA_PROJECT=os.environ.get("JULIA_PROJECT", os.path.join(ROOT, "baseline", "_ref")))
    try:
        r = subprocess.run([julia, os.path.join(ROOT, "julia", "bench_reference.jl"), str(args.n_spins), str(args.steps), str(args.warmup)],
                           capture_output=True, text=True, timeout=900, env=env)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
    except Exception:  # noqa: BLE001
        return None
    return None


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.  The
    real Julia package if a Julia install is found (kind "reference"), else the C restatement
    oracle/cheby_ref.c (kind "port") with all the host threads it can use.  Honours --steps /
    --warmup; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = build_workload(args.n_spins, 0, 1)
    julia = find_julia()
    jl = run_julia_reference(julia, args) if julia else None
    probe = f"julia: {julia}" if julia else "julia: not found on PATH or under baseline/_ref"
    if jl is not None and "prop_steps_per_s" in jl:
        value = float(jl["prop_steps_per_s"])
        base = {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"QuantumPropagators.jl prop_step! x{args.steps} after {args.warmup} warm-up steps, JULIA_NUM_THREADS=1, "
                          f"BLAS threads {jl.get('blas_threads')}; {probe}"}
        n_c = int(jl.get("n_coeffs", 0))
    else:
        r = oracle_steps(w, args.steps, n_warm=max(1, args.warmup))
        base = cpu_numbers(r, args.steps, max(1, args.warmup))
        base["sample"] += f"  ({probe}" + ("; the Julia run failed, fell back to the port)" if julia else ")")
        value = base["value"]
        n_c = r["n_coeffs"]
    out = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (ComplexF64)", "data": "synthetic",
        "config": {"workload": workload_name(args.n_spins, n_c)},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------
# GPU timing helpers
# ------------------------------------------------------------------------------------------


class BlockTimer:
    """Blocks of K steps, each bracketed by CUDA events on the library's stream, repeated until
    `min_seconds` of device time (pilot block decides the count)."""

    def __init__(self, torch, ctx, device):
        self.torch = torch
        self.ctx = ctx
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    def _block(self, step, k):
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(k):
            step()
        e1.record(self.stream)
        return e0, e1

    def run(self, step, k, min_seconds, barrier, before_block=None, max_blocks=2000, counter=None):
        if before_block:
            before_block()
        e0, e1 = self._block(step, k)  # pilot (also a warm-up)
        self.ctx.sync()
        pilot_ms = max(e0.elapsed_time(e1), 1e-3)
        n_blocks = int(min(max_blocks, max(3, math.ceil(min_seconds * 1e3 / pilot_ms))))
        events = []
        barrier()
        c0 = counter() if counter else 0
        t0 = time.perf_counter()
        for _ in range(n_blocks):
            if before_block:
                before_block()
            events.append(self._block(step, k))
        self.ctx.sync()
        wall = time.perf_counter() - t0
        barrier()
        ms = [a.elapsed_time(b) for a, b in events]
        return {"blocks_ms": ms, "median_ms": statistics.median(ms), "min_ms": min(ms), "max_ms": max(ms),
                "n_blocks": n_blocks, "device_s": sum(ms) * 1e-3, "wall_s": wall,
                "launches": (counter() - c0) if counter else None}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import qprop_b200 as qp

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    numa = bind_to_gpu_numa(local_rank)
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    ctx = qp.Context(local_rank)
    w = build_workload(args.n_spins, rank, world)
    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    gen = qp.hamiltonian(*terms)
    N = w["psi0"].shape[0]
    kw = dict(E_min=w["E_min"], E_max=w["E_max"])
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, matrix_format=args.format, **kw)
    n_c = p.wrk.n_coeffs
    n_terms = n_c - 1
    n_grid = len(w["tlist"]) - 1
    K = args.steps
    if K > n_grid:
        raise SystemExit("--steps exceeds the time grid (100 intervals)")
    timer = BlockTimer(torch, ctx, local_rank)

    def make_room(prop, k):
        def f():
            if prop.n + k > n_grid + 1:
                qp.reinit_prop(prop, prop.state)  # t <- tlist[1]; keeps the (already propagated) state
        return f

    # ---------------- device-resident timing: blocks of K prop_step! ------------------------
    for _ in range(args.warmup):
        make_room(p, 1)()
        qp.prop_step(p)
    ctx.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    main_t = timer.run(lambda: qp.prop_step(p), K, args.min_seconds, barrier, before_block=make_room(p, K),
                       counter=lambda: ctx.launch_count)
    clocks = sampler.stop() if rank == 0 else None
    launches = main_t["launches"]  # kernels of libqprop_b200 launched inside the timed region
    norm_dev = abs(p.state.norm() - 1.0)

    # ---------------- the uncompressed SELL-32 / TMA arm (matrix > L2) ----------------------
    sell_t = None
    g_sell = None
    if not args.no_sell_arm and p.wrk.gen.format != "sell":
        p_sell = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, matrix_format="sell", **kw)
        g_sell = p_sell.wrk.gen
        for _ in range(3):
            qp.prop_step(p_sell)
        sell_t = timer.run(lambda: qp.prop_step(p_sell), K, 0.5 * args.min_seconds, barrier, before_block=make_room(p_sell, K))
        del p_sell
    # the generic sparse path (dictionary-compressed SELL-32) when the headline runs on the structural bit-flip form
    dict_t = None
    g_dict = None
    if not args.no_sell_arm and p.wrk.gen.format == "bitflip":
        p_dict = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, matrix_format="selld", **kw)
        g_dict = p_dict.wrk.gen
        for _ in range(3):
            qp.prop_step(p_dict)
        dict_t = timer.run(lambda: qp.prop_step(p_dict), K, 0.5 * args.min_seconds, barrier, before_block=make_room(p_dict, K))
        del p_dict

    # ---------------- end to end through the public API with host buffers -------------------
    # (a) strictly sequential, one trajectory: pinned host state -> device, prop_step!, device -> pinned host
    qp.reinit_prop(p, p.state)
    lib = ctx._lib
    host_a = torch.empty(N, dtype=torch.complex128, pin_memory=True)
    host_a.copy_(torch.from_numpy(w["psi0"]))
    host_b = torch.empty(N, dtype=torch.complex128, pin_memory=True)
    host_b.copy_(host_a)

    def seq_block(k):
        for _ in range(k):
            make_room(p, 1)()
            lib.qp_state_upload(p.state.handle, host_a.data_ptr(), 0, 1)
            qp.prop_step(p)
            lib.qp_state_download(p.state.handle, host_a.data_ptr(), 0, 1)

    seq_block(2)
    barrier()
    t0 = time.perf_counter()
    n_seq = 0
    while n_seq < 3 * K or time.perf_counter() - t0 < 0.3 * args.min_seconds:
        seq_block(K)
        n_seq += K
    ctx.sync()
    seq_s = (time.perf_counter() - t0) / n_seq
    barrier()
    # (b) several host-resident trajectories per GPU, one context (= one stream) each: the copies of
    # one overlap the steps of the others; every step still moves its state in and out
    qp.reinit_prop(p, p.state)
    lanes = [(p, ctx, host_a)]
    for i in range(1, max(2, args.e2e_lanes)):
        ctx_i = qp.Context(local_rank)
        p_i = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx_i, matrix_format=args.format, **kw)
        if i == 1:
            host_i = host_b
        else:
            host_i = torch.empty(N, dtype=torch.complex128, pin_memory=True)
            host_i.copy_(host_a)
        lanes.append((p_i, ctx_i, host_i))
    n_lanes = len(lanes)

    def enqueue(lane):
        prop, c, host = lane
        make_room(prop, 1)()
        c._lib.qp_state_upload_async(prop.state.handle, host.data_ptr(), 0, 1)
        qp.prop_step(prop)
        c._lib.qp_state_download_async(prop.state.handle, host.data_ptr(), 0, 1)

    def pipelined(n_steps):
        for lane in lanes:
            enqueue(lane)
        done = n_lanes
        i = 0
        while done < n_steps:
            lane = lanes[i % n_lanes]
            lane[1].sync()  # the host owns this trajectory's buffer again (result of its last step)
            enqueue(lane)
            done += 1
            i += 1
        for lane in lanes:
            lane[1].sync()
        return done

    pipelined(2 * n_lanes + 2)
    barrier()
    t0 = time.perf_counter()
    n_e2e = 0
    while n_e2e < 3 * K or time.perf_counter() - t0 < 0.5 * args.min_seconds:
        n_e2e += pipelined(max(K, 4))
    e2e_s = (time.perf_counter() - t0) / n_e2e
    barrier()
    del lanes

    # ---------------- ensemble (BASELINE configs[2]) -----------------------------------------
    ens_out = None
    if not args.no_ensemble:
        ens_out = run_ensemble(args, qp, torch, ctx, rank, world, local_rank, barrier, max_over_ranks)

    # ---------------- reduce over ranks ------------------------------------------------------
    red = max_over_ranks([main_t["median_ms"], main_t["min_ms"], main_t["max_ms"], 1e3 * e2e_s, 1e3 * seq_s, norm_dev,
                          sell_t["median_ms"] if sell_t else 0.0, dict_t["median_ms"] if dict_t else 0.0])
    ms_med, ms_min, ms_max, e2e_ms, seq_ms, norm_dev, sell_ms, dict_ms = red

    if rank == 0:
        peak, peak_src = measured_peak()
        g = p.wrk.gen
        fmt = g.format
        canonical_term = g.matrix_bytes + 80 * N          # M + 80 N (canonical CSR)
        stored_term = g.stored_bytes + 80 * N              # what this format has to stream per launch
        launch_us = 1e3 * ms_med / (K * n_terms)
        achieved = stored_term / (launch_us * 1e-6) / 1e9
        kernel = {
            "selld": f"k_spmv_selld* <CHEB_MID> (dictionary-compressed SELL-32, {g.n_dict} table entries, CB={g.code_bytes}, "
                     "programmatic dependent launch)",
            "bitflip": "k_spmv_bitflip<CHEB_MID> (no matrix stream: explicit real diagonals + XOR-stencil terms from the "
                       "constant bank, programmatic dependent launch)",
            "sell": ("k_spmv_sell_tma<CHEB_MID,16,2,8>" if os.environ.get("QPROP_SELL_KERNEL", "tma") != "ldg" else "k_spmv_sell<CHEB_MID>"),
        }.get(fmt, f"k_spmv_{fmt}<CHEB_MID>")
        traffic, traffic_src = profiled_traffic(fmt)
        ws_mb = stored_term / 1e6
        out = {
            "metric": METRIC,
            "value": world * K / (ms_med * 1e-3),
            "unit": UNIT,
            "n_gpus": world,
            "steps": K,
            "warmup": args.warmup,
            "ms_per_step": ms_med / K,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64 (ComplexF64)",
            "data": "synthetic",
            "timing": {
                "protocol": f"blocks of {K} prop_step! bracketed by CUDA events on the library's stream; median block reported, max over ranks",
                "blocks": main_t["n_blocks"], "timed_region_s": main_t["device_s"], "timed_region_wall_s": main_t["wall_s"],
                "ms_per_step_min": ms_min / K, "ms_per_step_max": ms_max / K,
            },
            "config": {
                "workload": workload_name(args.n_spins, n_c),
                "parallelism": f"trajectory-sharded x{world} (independent replicas, no data-path collective)",
                "matrix_format": fmt,
                "matrix_format_note": ("chosen by QP_FORMAT_AUTO from the uploaded sparse matrices (diagonal operators + uniform bit flips, "
                                       "csrc/bitflip.cu); a generator of the same size and sparsity without that structure runs on the "
                                       "kernel of `arms.selld`" if fmt == "bitflip" else "chosen by QP_FORMAT_AUTO / --format"),
                "l2": (f"working set per term {ws_mb:.0f} MB (matrix as stored {g.stored_bytes / 1e6:.1f} MB + {80 * N / 1e6:.1f} MB of vectors) "
                       + (f"> L2 ({L2_MB:.0f} MB): streamed from HBM every term" if ws_mb > 1.5 * L2_MB else
                          f"is NOT larger than L2 ({L2_MB:.0f} MB) and nothing is flushed between terms (a prop_step! is {n_terms} "
                          "back-to-back launches; the vectors written by term k are read by term k+1): part of the traffic is "
                          "served by L2, see roofline.traffic; the arm `arms.sell_tma` streams 474 MB per term from HBM")),
                "norm_deviation_after_run": norm_dev,
                "norm_note": "per-step norm conservation is 1e-12 (tests); over many steps the reference's own coefficient truncation (limit 1e-12 per step) accumulates",
                "host_affinity": numa,
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": kernel + " (one fused Chebyshev term per launch)",
                "bytes_definition": "bytes the timed storage format must stream per launch: matrix as stored + 80 N (read v_k, v_{k-1}, psi; write v_{k+1}, psi)",
                "bytes_per_launch": stored_term, "avg_launch_us": launch_us,
                # canonical-CSR bytes of SURVEY.md 8d (M + 80 N): what an uncompressed format would move
                "canonical_bytes_per_launch": canonical_term,
                "effective_vs_canonical": canonical_term / (launch_us * 1e-6) / 1e9 / peak,
                "vector_floor_us": 80 * N / peak / 1e3,
            },
            "e2e": {
                "value": world * 1e3 / e2e_ms, "unit": UNIT,
                "h2d_bytes_per_step": 16 * N + 16 * p.wrk.gen.n_coeffs, "d2h_bytes_per_step": 16 * N,
                "note": "host-resident states: every step copies its state pinned host -> device, runs prop_step!, copies device -> pinned host; "
                        f"{max(2, args.e2e_lanes)} trajectories per GPU interleaved on their own contexts (streams) so the copies of one overlap the steps of the others",
                "sequential_value": world * 1e3 / seq_ms,
                "sequential_note": "one trajectory, upload -> prop_step! -> download strictly in sequence",
            },
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if sell_t is not None:
            sell_us = 1e3 * sell_ms / (K * n_terms)
            sell_bytes = g_sell.stored_bytes + 80 * N
            out["arms"] = {"sell_tma": {
                "value": world * K / (sell_ms * 1e-3), "unit": UNIT, "ms_per_step": sell_ms / K, "blocks": sell_t["n_blocks"],
                "matrix_format": "sell", "kernel": "k_spmv_sell_tma<CHEB_MID,16,2,8> (uncompressed SELL-32, cp.async.bulk + mbarrier staging)",
                "l2": f"matrix {g_sell.stored_bytes / 1e6:.0f} MB per term > L2: streamed from HBM",
                "roofline": {"bound": "hbm", "achieved": sell_bytes / (sell_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": sell_bytes / (sell_us * 1e-6) / 1e9 / peak, "bytes_per_launch": sell_bytes,
                             "canonical_bytes_per_launch": canonical_term, "avg_launch_us": sell_us},
            }}
        if dict_t is not None:
            dict_us = 1e3 * dict_ms / (K * n_terms)
            dict_bytes = g_dict.stored_bytes + 80 * N
            out.setdefault("arms", {})["selld"] = {
                "value": world * K / (dict_ms * 1e-3), "unit": UNIT, "ms_per_step": dict_ms / K, "blocks": dict_t["n_blocks"],
                "matrix_format": "selld",
                "kernel": f"k_spmv_selld* <CHEB_MID> (generic sparse path: dictionary-compressed SELL-32, {g_dict.n_dict} table entries)",
                "note": "what a generator WITHOUT the diagonal + bit-flip structure of this workload gets at the same size and sparsity",
                "roofline": {"bound": "hbm", "achieved": dict_bytes / (dict_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": dict_bytes / (dict_us * 1e-6) / 1e9 / peak, "bytes_per_launch": dict_bytes,
                             "canonical_bytes_per_launch": canonical_term, "avg_launch_us": dict_us},
            }
        if ens_out is not None:
            out["ensemble"] = ens_out
        if world == 1 and not args.no_cpu_baseline:
            r = oracle_steps(w, args.cpu_sample_steps)
            out["cpu_baseline"] = cpu_numbers(r, args.cpu_sample_steps, 1)
            # parity of the product against the CPU port after the same steps from the same state
            pp = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, matrix_format=args.format, **kw)
            for _ in range(r["steps_done"]):
                qp.prop_step(pp)
            got = pp.state.to_host()
            out["parity_rel_err"] = float(np.linalg.norm(got - r["final"]) / np.linalg.norm(r["final"]))
            out["parity_note"] = f"||psi_gpu - psi_cpu|| / ||psi_cpu|| after {r['steps_done']} prop_step! from the same initial state (bar 1e-10)"
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_ensemble(args, qp, torch, ctx, rank, world, local_rank, barrier, max_over_ranks):
    """BASELINE configs[2]: B trajectories (varied control amplitudes) of the N = 2^16 transmon chain,
    sharded by contiguous trajectory blocks over the ranks; no collective inside the time loop, one
    gather of expectation values and one of the final states at the end."""
    from qprop_b200.ensemble import EnsembleChebyPropagator, LibraryEnsemble

    B, k = args.ensemble_B, args.ensemble_steps
    if B < world:
        return None
    # the communicator INSIDE libqprop_b200.so (qp_ens_*: NCCL over NVLink); the launcher's process
    # group only hands the 128-byte id from rank 0 to the other ranks
    ident = None
    if world > 1:
        import torch.distributed as dist

        t_id = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t_id = torch.tensor(list(LibraryEnsemble.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t_id, 0)
        ident = bytes(t_id.cpu().tolist())
    comm = LibraryEnsemble.from_rank(ctx, rank, world, ident)
    w = qp.workloads.config3_transmon(n_sites=8, levels=4, B=B, nt=201, dt=0.5)
    H0, H1, H2 = w["ops"]
    N = H0.shape[0]
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())
    ens = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], w["psi0"], w["tlist"], -bound, bound, ctx,
                                  rank=rank, world=world, library_ensemble=comm)
    tile = ens.gen.tile_info()
    n_grid = len(w["tlist"]) - 1
    for _ in range(2):
        ens.prop_step()
    timer = BlockTimer(torch, ctx, local_rank)

    def room():
        if ens.n + k > n_grid + 1:
            ens.n = 1

    t = timer.run(ens.prop_step, k, args.min_seconds, barrier, before_block=room, max_blocks=50)
    # the only data collectives: per-trajectory expectation values, then the final states
    ctx.sync()
    barrier()
    t0 = time.perf_counter()
    pops = ens.gather_expvals(np.asarray(ens.state.norm()))
    barrier()
    t_exp = time.perf_counter() - t0
    t0 = time.perf_counter()
    # final states: gathered over NVLink into a device-resident (N, B) state on every rank
    full = qp.DeviceState(ctx, N, B) if world > 1 else None
    if world > 1:
        comm.gather_states([ens.state], B, to_host=False, full_states=[full])
    ctx.sync()
    barrier()
    t_states = time.perf_counter() - t0
    gathered = int(np.asarray(pops).shape[-1])
    n_states = B if world > 1 else ens.B_local
    del full
    med, mn, mx, t_exp, t_states = max_over_ranks([t["median_ms"], t["min_ms"], t["max_ms"], t_exp, t_states])
    peak, _ = measured_peak()
    n_c = ens.wrk.n_coeffs
    term_bytes = ens.gen.stored_bytes + 80 * N * ens.B_local
    per_gpu = term_bytes * (n_c - 1) * k / (med * 1e-3) / 1e9
    out = {
        "metric": "cheby_trajectory_steps_per_s", "value": B * k / (med * 1e-3), "unit": "trajectory-steps/s",
        "scaling": "strong", "n_gpus": world, "steps": k, "blocks": t["n_blocks"], "ms_per_step": med / k,
        "ms_per_step_min": mn / k, "ms_per_step_max": mx / k,
        "workload": f"transmon chain 8x4 levels (N=2^16), B={B} trajectories with their own control amplitudes, B_local={ens.B_local}, "
                    f"n_coeffs={n_c}",
        "kernel": (f"k_spmm_tile<CHEB_MID,3> (two-pass tiled, split {tile['split']} x {tile['blocks']}, {tile['n_table']} table entries; "
                   f"entries per class {tile['entries']})" if tile["available"] and ens.B_local % 32 == 0 else
                   "k_spmm_selld_pairs / k_spmm_selld (one-pass dictionary kernels)"),
        "achieved_gbs_per_gpu": per_gpu, "roofline_frac_per_gpu": per_gpu / peak,
        "bytes_per_term_per_gpu": term_bytes, "ms_per_term": med / k / (n_c - 1),
        "gather": {"expvals_s": t_exp, "states_s": t_states, "trajectories_gathered": gathered, "states_gathered": n_states,
                   "states_bytes": 16 * N * B if world > 1 else 0, "backend": ens.gather_backend,
                   "note": "expvals: host numbers -> every rank; states: device-resident (N, B) on every rank (skipped at 1 GPU)"},
        "norm_dev_max": float(np.max(np.abs(np.asarray(pops) - 1))),
    }
    del ens
    return out


if __name__ == "__main__":
    main()
