#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs that are NOT the bench.py headline
(configs[1]) -- config 1 (N=1000, launch-latency regime), config 3 (transmon ensemble, batched
SpMM), config 4 (Liouvillian + Newton/Arnoldi, reduced or full size), config 5 (dense generator,
FP64 tensor-core path).  One JSON line per measurement on stdout.  CUDA-event timing on the
library's stream after warm-up; inputs larger than L2 except config 1 (stated).

    python tools/bench_configs.py --configs 1,3,5 [--B 1024] [--steps 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import qprop_b200 as qp  # noqa: E402

PEAK_HBM = 6545.6
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        PEAK_HBM = float(json.load(f)["hbm_gbs"])
except Exception:
    pass


class Timer:
    def __init__(self, ctx):
        self.ctx = ctx
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=ctx.device)
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        self.ctx.sync()
        self.e0.record(self.stream)
        return self

    def __exit__(self, *a):
        self.e1.record(self.stream)
        self.ctx.sync()
        self.ms = self.e0.elapsed_time(self.e1)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def config1(ctx, args):
    w = qp.workloads.config1_random(N=1000, density=0.1, seed=1000)
    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    p = qp.init_prop(w["psi0"], qp.hamiltonian(*terms), w["tlist"], "cheby", ctx=ctx, E_min=w["E_min"], E_max=w["E_max"])
    for _ in range(20):
        qp.prop_step(p)
    steps = 400
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    with Timer(ctx) as t:
        for _ in range(steps):
            qp.prop_step(p)
    wall = time.perf_counter() - t0
    # the whole 500-step grid as ONE library call (qp_cheby_propagate; what propagate() does without a callback)
    p2 = qp.init_prop(w["psi0"], qp.hamiltonian(*terms), w["tlist"], "cheby", ctx=ctx, E_min=w["E_min"], E_max=w["E_max"])
    qp.propagate(p2)
    p2 = qp.init_prop(w["psi0"], qp.hamiltonian(*terms), w["tlist"], "cheby", ctx=ctx, E_min=w["E_min"], E_max=w["E_max"])
    t1 = time.perf_counter()
    with Timer(ctx) as t2:
        qp.propagate(p2)
    wall2 = time.perf_counter() - t1
    n2 = len(w["tlist"]) - 1
    emit(config=1, workload="random sparse Hermitian N=1000 + 1 control, Cheby, propagate() = one qp_cheby_propagate call",
         steps=n2, prop_steps_per_s=n2 / (t2.ms * 1e-3), us_per_step=1e3 * t2.ms / n2, wall_us_per_step=1e6 * wall2 / n2,
         norm_dev=abs(p2.state.norm() - 1))
    emit(config=1, workload="random sparse Hermitian N=1000 + 1 control, Cheby", n_coeffs=p.wrk.n_coeffs, steps=steps,
         prop_steps_per_s=steps / (t.ms * 1e-3), us_per_step=1e3 * t.ms / steps, wall_us_per_step=1e6 * wall / steps,
         launches_per_step=(ctx.launch_count - l0) / steps, format=p.wrk.gen.format,
         note="L2-resident (4 MB matrix): latency-bound, no roofline", norm_dev=abs(p.state.norm() - 1))


def config3(ctx, args):
    from qprop_b200.ensemble import EnsembleChebyPropagator

    B = args.B
    w = qp.workloads.config3_transmon(n_sites=args.sites, levels=4, B=B, nt=args.steps + 6, dt=0.5)
    H0, H1, H2 = w["ops"]
    bound = float((abs(H0) + 0.1 * abs(H1) + 0.1 * abs(H2)).sum(axis=1).max())  # |u_l s_b| <= 0.075
    ens = EnsembleChebyPropagator(w["ops"], w["controls"], w["scales"], w["psi0"], w["tlist"], -bound, bound, ctx,
                                  matrix_format=args.format)
    N = H0.shape[0]
    n_c = ens.wrk.n_coeffs
    for _ in range(3):
        ens.prop_step()
    l0 = ctx.launch_count
    with Timer(ctx) as t:
        for _ in range(args.steps):
            ens.prop_step()
    term_bytes = ens.gen.matrix_bytes + 80 * N * B
    us_term = 1e3 * t.ms / (args.steps * (n_c - 1))
    norms = ens.state.norm()
    emit(config=3, workload=f"transmon chain {args.sites}x4 levels N={N}, B={B} trajectories (per-trajectory amplitudes), Cheby",
         n_coeffs=n_c, steps=args.steps, trajectory_steps_per_s=B * args.steps / (t.ms * 1e-3), ms_per_step=t.ms / args.steps,
         us_per_term=us_term, algorithmic_bytes_per_term=term_bytes, effective_gbs=term_bytes / us_term / 1e3,
         frac_of_measured_hbm=term_bytes / us_term / 1e3 / PEAK_HBM, format=ens.gen.format, n_dict=ens.gen.n_dict,
         launches=ctx.launch_count - l0, norm_dev_max=float(np.max(np.abs(np.asarray(norms) - 1))))


def config5(ctx, args):
    H = qp.workloads.config5_optomech_dense()
    N = H.shape[0]
    ev_bound = float(np.abs(H).sum(axis=1).max())
    gen = qp.DeviceGenerator(ctx, [H], 0)
    rng = np.random.default_rng(5000)
    for B in [int(b) for b in args.dense_B.split(",")]:
        psi = rng.standard_normal((N, B)) + 1j * rng.standard_normal((N, B))
        psi /= np.linalg.norm(psi, axis=0)
        st = qp.DeviceState.from_host(ctx, psi if B > 1 else psi[:, 0])
        dt = 60.0 / ev_bound  # alpha = Delta dt / 2 ~ 30 -> n_coeffs ~ 60
        wrk = qp.ChebyWrk(st, gen, 2 * ev_bound, -ev_bound, dt)
        for _ in range(2):
            qp.cheby_(st, None, dt, wrk, coeffs=[])
        steps = 3
        with Timer(ctx) as t:
            for _ in range(steps):
                qp.cheby_(st, None, dt, wrk, coeffs=[])
        n_terms = wrk.n_coeffs - 1
        us_term = 1e3 * t.ms / (steps * n_terms)
        flops = 8.0 * N * N * B
        bytes_term = 16.0 * N * N + 80.0 * N * B
        nrm = np.atleast_1d(st.norm())
        emit(config=5, workload=f"dense optomech generator N={N}, B={B} states, Cheby", n_coeffs=wrk.n_coeffs, steps=steps,
             state_steps_per_s=B * steps / (t.ms * 1e-3), us_per_term=us_term, tflops_fp64=flops / us_term / 1e6,
             effective_gbs=bytes_term / us_term / 1e3, frac_of_measured_hbm=bytes_term / us_term / 1e3 / PEAK_HBM,
             kernel="k_gemv_dense" if B == 1 else "k_gemm_dense (DMMA m8n8k4.f64)", norm_dev_max=float(np.max(np.abs(nrm - 1))))


def config4(ctx, args):
    n_spins = args.liou_spins
    t_host = time.perf_counter()
    w = qp.workloads.config4_liouvillian(n_spins=n_spins, nt=args.newton_steps + 3, dt=0.05, matrix_free=args.matrix_free)
    t_host = time.perf_counter() - t_host
    t_build = time.perf_counter()
    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    p = qp.init_prop(w["psi0"], qp.hamiltonian(*terms), w["tlist"], "newton", ctx=ctx, m_max=10, relerr=1e-12)
    t_build = time.perf_counter() - t_build
    qp.prop_step(p)
    ctx.sync()
    l0 = ctx.launch_count
    with Timer(ctx) as t:
        for _ in range(args.newton_steps):
            qp.prop_step(p)
    N = w["psi0"].shape[0]
    free, total = torch.cuda.mem_get_info()
    emit(config=4, workload=f"Liouvillian of {n_spins}-spin TFIM + decay, dim {N}, Newton m_max=10" + (", matrix-free" if args.matrix_free else ""),
         steps=args.newton_steps, host_build_s=t_host, device_memory_used_gb=(total - free) / 2**30,
         prop_steps_per_s=args.newton_steps / (t.ms * 1e-3), ms_per_step=t.ms / args.newton_steps, format=p.wrk.krylov.gen.format,
         n_dict=p.wrk.krylov.gen.n_dict, matrix_bytes=p.wrk.krylov.gen.matrix_bytes, launches_per_step=(ctx.launch_count - l0) / args.newton_steps,
         restarts_per_step=getattr(p.wrk, "restarts", None), setup_s=t_build)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,5")
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--sites", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--format", default="auto")
    ap.add_argument("--dense-B", default="1,16,64")
    ap.add_argument("--liou-spins", type=int, default=10)
    ap.add_argument("--newton-steps", type=int, default=5)
    ap.add_argument("--matrix-free", action="store_true", help="config 4 with liouvillian(..., matrix_free=True)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = qp.Context(0)
    for c in args.configs.split(","):
        {"1": config1, "3": config3, "4": config4, "5": config5}[c](ctx, args)


if __name__ == "__main__":
    main()
