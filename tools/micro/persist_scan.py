#!/usr/bin/env python
"""Per-term cost of the Chebyshev step at small N: n_coeffs is varied through dt so that the
fixed cost per step and the cost per term separate; run with QPROP_PERSISTENT=0 and 1."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import qprop_b200 as qp  # noqa: E402

ctx = qp.default_context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=ctx.device)
for N in (1000, 8000):
    for T in (10.0, 100.0, 400.0):
        w = qp.workloads.config1_random(N=N, density=100.0 / N, seed=1000, nt=501, T=T)
        terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
        p = qp.init_prop(w["psi0"], qp.hamiltonian(*terms), w["tlist"], "cheby", ctx=ctx, E_min=w["E_min"], E_max=w["E_max"])
        for _ in range(20):
            qp.prop_step(p)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.sync()
        e0.record(stream)
        steps = 300
        for _ in range(steps):
            qp.prop_step(p)
        e1.record(stream)
        ctx.sync()
        us = 1e3 * e0.elapsed_time(e1) / steps
        n_a = p.wrk.n_coeffs
        print(json.dumps(dict(persistent=os.environ.get("QPROP_PERSISTENT", "1"), N=N, n_coeffs=n_a, us_per_step=us,
                              us_per_term=us / (n_a - 1), norm_dev=abs(p.state.norm() - 1))), flush=True)
