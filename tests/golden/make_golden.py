"""Generates the committed golden fixtures in this directory.

The reference (Julia) cannot run in this image, and it ships no golden vectors; these
fixtures therefore hold (a) the oracle's results and (b) an INDEPENDENT ground truth for the
same inputs (dense `scipy.linalg.expm` / `scipy.sparse.linalg.expm_multiply` applied interval
by interval), so that the oracle is pinned against something that is not itself.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.linalg as sla
from scipy.sparse.linalg import expm_multiply

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle as O  # noqa: E402
import qprop_b200.workloads as W  # noqa: E402


def optomech():
    H = W.optomech()
    psi0 = W.optomech_ket(0, 2)
    tlist = np.arange(0, 50 + 1e-9, 0.2)
    cheby = O.propagate(psi0, (H,), tlist, "cheby")
    newton = O.propagate(psi0, (H,), tlist, "newton")
    U = sla.expm(-1j * H.toarray() * 0.2)
    exact = psi0.copy()
    for _ in range(len(tlist) - 1):
        exact = U @ exact
    np.savez(os.path.join(HERE, "optomech_final.npz"), cheby=cheby, newton=newton, expm=exact)
    print("optomech: |cheby-newton| = %.2e, |cheby-expm| = %.2e" % (np.linalg.norm(cheby - newton), np.linalg.norm(cheby - exact)))


def tfim8():
    w = W.config2_tfim(n_spins=8, nt=21, dt=0.1)
    terms = [w["ops"][0]] + list(zip(w["ops"][1:], w["controls"]))
    cheby = O.propagate(w["psi0"], O.hamiltonian(*terms), w["tlist"], "cheby", E_min=w["E_min"], E_max=w["E_max"])
    tl = w["tlist"]
    mids = O.get_tlist_midpoints(tl)
    exact = w["psi0"].copy()
    for k in range(len(tl) - 1):
        Hk = w["ops"][0] + w["controls"][0](mids[k]) * w["ops"][1] + w["controls"][1](mids[k]) * w["ops"][2]
        exact = sla.expm(-1j * (tl[k + 1] - tl[k]) * Hk.toarray()) @ exact
    np.savez(os.path.join(HERE, "tfim8_final.npz"), psi0=w["psi0"], cheby=cheby, exact=exact)
    print("tfim8: |cheby-exact| = %.2e" % np.linalg.norm(cheby - exact))


if __name__ == "__main__":
    optomech()
    tfim8()
