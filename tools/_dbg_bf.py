import numpy as np, sys, os
sys.path.insert(0, "/root/repo")
import qprop_b200 as qp
import oracle as O
ctx = qp.Context(0)
order = sys.argv[1].split(",")
for n_spins in (6,):
    w = qp.workloads.config2_tfim(n_spins, nt=11, dt=0.1)
    kw = dict(E_min=w["E_min"], E_max=w["E_max"])
    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    ref = O.propagate(w["psi0"], O.hamiltonian(*terms), w["tlist"], "cheby", **kw)
    for fmt in order:
        gen = qp.hamiltonian(*terms)
        p = qp.init_prop(w["psi0"], gen, w["tlist"], "cheby", ctx=ctx, matrix_format=fmt, **kw)
        out = qp.propagate(p)
        print(n_spins, fmt, "propagate err", np.linalg.norm(out - ref) / np.linalg.norm(ref), "n_coeffs", p.wrk.n_coeffs, flush=True)
