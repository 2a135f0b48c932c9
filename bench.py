#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: Chebyshev ``prop_step!``/s (and effective
HBM GB/s against the roofline) for the transverse-field Ising chain N = 2^20 with two PWC
controls (BASELINE configs[1]) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n-spins 20]

N > 1 is launched by torchrun (one rank per GPU).  A single large state stays on one GPU
(north_star), so the ranks hold independent trajectories of the same system (each with its own
control scale) -- trajectory sharding, no data-path collective; NCCL is used for the barrier,
the max-over-ranks timing and one final all_gather of an expectation value.  ``value`` is
whole-job prop_step!/s = N * K / max-rank time ("scaling": "weak").

One JSON line is printed by rank 0.  Keys beyond the base contract: ``roofline``,
``cpu_baseline``, ``e2e``, ``gpu_launches``, ``clocks``.
"""

import argparse
import json
import os
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-spins", type=int, default=20)
    ap.add_argument("--format", default="auto", choices=["auto", "csr", "sell", "selld"])
    ap.add_argument("--cpu-sample-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(fmt):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(fmt)
    except Exception:
        return None


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


def oracle_steps(w, n_steps):
    """Time `n_steps` prop_step! of the CPU port on this workload (after one warm-up step).

    Returns a dict with seconds per step for (a) the faithful restatement of the reference's CPU
    path -- per-operator CSC scatter SpMV + level-1 passes, Int64 indices, ONE thread, exactly
    what Julia's SparseArrays `mul!` does (oracle/cheby_ref.c:cheby_step_csc) -- and (b) a stronger
    baseline the reference does not have: the same step with a row-parallel OpenMP CSR SpMV on
    all host threads (cheby_step_csr_omp).  Falls back to the NumPy/SciPy oracle if the C port is
    not built."""
    import oracle as O
    from oracle import cref

    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    p = O.init_prop(w["psi0"], O.hamiltonian(*terms), w["tlist"], "cheby", E_min=w["E_min"], E_max=w["E_max"])
    n_c = p.wrk.n_coeffs
    if not cref.available():
        O.prop_step(p)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            O.prop_step(p)
        sec = (time.perf_counter() - t0) / n_steps
        return {"n_coeffs": n_c, "single": sec, "omp": None, "threads": 1, "impl": "numpy/scipy oracle (scipy CSR @, 1 thread)"}
    ref = cref.ChebyRef(w["ops"], len(w["controls"]))
    wrk = p.wrk
    out = {"n_coeffs": n_c, "impl": "oracle/cheby_ref.c"}
    for key, threads in (("single", 0), ("omp", cref.max_threads())):
        psi = w["psi0"].copy()
        secs = []
        for n in range(1, n_steps + 2):  # first step = warm-up
            coeffs = [complex(p.parameters[c][n - 1]) for c in p.controls]
            t0 = time.perf_counter()
            ref.step(psi, coeffs, wrk.coeffs, wrk.Delta, wrk.E_min, wrk.dt, threads=threads)
            secs.append(time.perf_counter() - t0)
        out[key] = sum(secs[1:]) / n_steps
    out["threads"] = cref.max_threads()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = build_workload(args.n_spins, 0, 1)
    n_steps = max(1, min(args.steps, args.cpu_sample_steps))
    r = oracle_steps(w, n_steps)
    out = {
        "impl": "reference",
        "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": n_steps, "warmup": 1,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (ComplexF64)", "data": "synthetic",
        "config": {"workload": f"TFIM chain n={args.n_spins} (N=2^{args.n_spins}), H0 + 2 PWC controls, Cheby prop_step!, n_coeffs={r['n_coeffs']}"},
    }
    out.update(cpu_numbers(r, n_steps))
    out["value"] = out["cpu_baseline"]["value"]
    out["ms_per_step"] = 1e3 / out["value"]
    out["e2e"] = {"value": out["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(out), flush=True)


def cpu_numbers(r, n_steps):
    """`cpu_baseline` object: the reported value uses all the host threads the port can use (the
    OpenMP variant); the faithful single-thread figure rides along."""
    best = r["omp"] if r.get("omp") else r["single"]
    cores = r["threads"] if r.get("omp") else 1
    return {
        "cpu_baseline": {
            "value": 1.0 / best, "unit": UNIT, "cores": cores, "kind": "port",
            "single_thread_value": 1.0 / r["single"],
            "sample": f"{n_steps} prop_step! of the same workload after 1 warm-up step; {r['impl']}: value = row-parallel "
                      f"OpenMP CSR variant on {cores} threads, single_thread_value = faithful restatement of the reference's "
                      "single-threaded CSC mul! + level-1 passes.  A port, not Julia (absent from this image).",
        }
    }


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
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = qp.Context(local_rank)
    w = build_workload(args.n_spins, rank, world)
    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    gen = qp.hamiltonian(*terms)
    N = w["psi0"].shape[0]
    p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, E_min=w["E_min"], E_max=w["E_max"],
                     matrix_format=args.format)
    n_c = p.wrk.n_coeffs
    step_bytes = p.wrk.step_bytes
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    def reset():
        qp.reinit_prop(p, p.state)  # t <- tlist[1]; keeps the (already propagated) state

    # ---------------- device-resident timing: K prop_step! ----------------------------------
    for _ in range(args.warmup):
        qp.prop_step(p)
    ctx.sync()
    if args.steps + args.warmup >= len(w["tlist"]):
        raise SystemExit("steps + warmup exceed the time grid")
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        qp.prop_step(p)
    ev1.record(stream)
    ctx.sync()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    norm_dev = abs(p.state.norm() - 1.0)

    # ---------------- end to end through the public API with host buffers -------------------
    # every step: pinned host state -> device, prop_step!, device -> pinned host
    reset()
    host = torch.empty(N, dtype=torch.complex128, pin_memory=True)
    host.copy_(torch.from_numpy(w["psi0"]))
    host_np = host.numpy()
    for _ in range(2):
        p.state.upload(host_np)
        qp.prop_step(p)
        host_np[:] = p.state.to_host()
    reset()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p.state.upload(host_np)
        qp.prop_step(p)
        ctx._lib.qp_state_download(p.state.handle, host.data_ptr(), 0, 1)
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---------------- reduce over ranks ------------------------------------------------------
    times = torch.tensor([ms, 1e3 * e2e_s], dtype=torch.float64, device="cuda")
    gathered = None
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        # the only data collective of the ensemble: gather one expectation value per trajectory
        mine = torch.tensor([norm_dev], dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        norm_dev = float(max(g.item() for g in gathered))
    ms_max, e2e_ms_max = float(times[0]), float(times[1])

    if rank == 0:
        peak, peak_src = measured_peak()
        n_terms = n_c - 1
        term_bytes = step_bytes / n_terms
        launch_us = 1e3 * ms_max / (args.steps * n_terms)
        achieved = term_bytes / (launch_us * 1e-6) / 1e9
        fmt = p.wrk.gen.format
        g = p.wrk.gen
        stored_term = g.stored_bytes + 80 * N  # bytes one launch actually has to move
        kernel = {
            "selld": f"k_spmv_selld<CHEB_MID,CB={g.code_bytes},TAIL,REALT> (dictionary-compressed SELL-32, {g.n_dict} table entries, "
                     "programmatic dependent launch)",
            "sell": ("k_spmv_sell_tma<CHEB_MID,16,2,8>" if os.environ.get("QPROP_SELL_KERNEL", "tma") != "ldg" else "k_spmv_sell<CHEB_MID>"),
        }.get(fmt, f"k_spmv_{fmt}<CHEB_MID>")
        out = {
            "metric": METRIC,
            "value": world * args.steps / (ms_max * 1e-3),
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64 (ComplexF64)",
            "data": "synthetic",
            "config": {
                "workload": f"TFIM chain n={args.n_spins} (N=2^{args.n_spins}), H0 + 2 PWC controls, Cheby prop_step!, n_coeffs={n_c}, B=1 per GPU",
                "parallelism": f"trajectory-sharded x{world} (independent replicas, no data-path collective)",
                "matrix_format": fmt,
                "l2": "inputs larger than L2 (matrix %.0f MB streamed every term)" % (p.wrk.gen.matrix_bytes / 1e6),
                "effective_hbm_gbs_per_gpu": achieved,
                "norm_deviation_after_run": norm_dev,
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": profiled_traffic(fmt), "peak_source": peak_src,
                "kernel": kernel + " (fused Chebyshev term)",
                "algorithmic_bytes_per_launch": term_bytes, "avg_launch_us": launch_us,
                # `achieved` counts the canonical-CSR bytes of SURVEY.md 8d (M + 80 N).  A compressed
                # format moves fewer bytes than that: the second pair is measured against the bytes
                # the chosen format really has to stream (matrix as stored + 80 N of vectors).
                "stored_bytes_per_launch": stored_term,
                "achieved_stored": stored_term / (launch_us * 1e-6) / 1e9,
                "frac_stored": stored_term / (launch_us * 1e-6) / 1e9 / peak,
            },
            "e2e": {
                "value": world * args.steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": 16 * N + 16 * gen_ncoeffs(p), "d2h_bytes_per_step": 16 * N,
                "note": "per step: pinned host state -> device, prop_step!, device -> pinned host (host-resident-state usage)",
            },
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            out.update(cpu_numbers(oracle_steps(w, args.cpu_sample_steps), args.cpu_sample_steps))
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def gen_ncoeffs(p):
    return p.wrk.gen.n_coeffs


if __name__ == "__main__":
    main()
