"""Host logic of the two-pass tile format (csrc/tile_format.h) on the CPU: the builder classifies
every entry (same block / same position / other / diagonal), merges operators that share a column
and numbers the distinct entries; `apply_host` walks the result the way the two CUDA kernels do.
Checked against SciPy on the BASELINE config-3 operator family and on unstructured matrices."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "tilefmt_test.cpp")
INC = os.path.join(ROOT, "quantumpropagators.jl_b200", "csrc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tilefmt") / "libtilefmt_test.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", INC, SRC, "-o", out], check=True)
    return C.CDLL(out)


def merged_csr(ops):
    """Merged multi-operator CSR as the library builds it: rows hold the entries of all operators
    back to back, operator index in the top 4 bits of the column word."""
    n = ops[0].shape[0]
    ops = [sp.csr_matrix(A, dtype=np.complex128) for A in ops]
    ptr = np.zeros(n + 1, dtype=np.uint32)
    cols, vals = [], []
    for r in range(n):
        for l, A in enumerate(ops):
            a, b = A.indptr[r], A.indptr[r + 1]
            cols.append(A.indices[a:b].astype(np.uint32) | np.uint32(l << 28))
            vals.append(A.data[a:b])
        ptr[r + 1] = ptr[r] + sum(len(c) for c in cols[-len(ops):])
    colop = np.concatenate(cols).astype(np.uint32) if cols else np.zeros(0, np.uint32)
    val = np.concatenate(vals).astype(np.complex128) if vals else np.zeros(0, np.complex128)
    return ptr, colop, val


def run(lib, ops, B, rng, S=0):
    n = ops[0].shape[0]
    ptr, colop, val = merged_csr(ops)
    u = rng.standard_normal((len(ops), B)) + 1j * rng.standard_normal((len(ops), B))
    x = rng.standard_normal((n, B)) + 1j * rng.standard_normal((n, B))
    y = np.zeros((n, B), dtype=np.complex128)
    stats = (C.c_longlong * 12)()
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib.tilefmt_apply(C.c_longlong(n), len(ops), P(ptr), P(colop), P(val), C.c_longlong(B), P(u), P(x), P(y), stats, S)
    if rc != 0:
        return None, None
    want = sum(u[l][None, :] * (sp.csr_matrix(ops[l]) @ x) for l in range(len(ops)))
    return np.max(np.abs(y - want)) / np.max(np.abs(want)), list(stats)


def test_choose_split(lib):
    assert lib.tilefmt_choose_split(C.c_longlong(65536)) == 256
    assert lib.tilefmt_choose_split(C.c_longlong(4096)) == 64
    assert lib.tilefmt_choose_split(C.c_longlong(256)) == 16
    assert lib.tilefmt_choose_split(C.c_longlong(512)) == 32
    assert lib.tilefmt_choose_split(C.c_longlong(1 << 17)) == 0      # needs tiles of more than 256 rows
    assert lib.tilefmt_choose_split(C.c_longlong(1000)) == 8          # 1000 = 8 * 125
    assert lib.tilefmt_choose_split(C.c_longlong(999)) == 0


def test_transmon_chain_classes(lib):
    """Config-3 operator family (4 sites x 4 levels, N = 256, S = 16): single-site terms and the hops
    inside one half are class A or B, only the hop that straddles the split is class O; the two
    control operators share their columns (merged entries)."""
    import qprop_b200 as qp

    H0, H1, H2 = qp.workloads.transmon_chain(4, 4)
    err, st = run(lib, [H0, H1, H2], 8, np.random.default_rng(1))
    assert err < 1e-14
    S, NH, WA, WB, n_tab, nA, nB, nO, nD, imag, nP, nG = st
    assert (S, NH) == (16, 16) and imag == 0b100          # H2 = i sum(a^+ - a) is purely imaginary
    assert nP == H1.nnz - 0 and nG == 0                   # every control entry is a quadrature pair (|v1| = |v2|)
    assert 0 < nO < 0.2 * (nA + nB)                        # only the (1,2) hop straddles the split
    merged = nA + nB + nO + nD
    assert merged < H0.nnz + H1.nnz + H2.nnz              # H1 / H2 columns are shared
    assert n_tab < 400


def test_split_follows_the_tensor_structure(lib):
    """3 sites x 4 levels (N = 64): the balanced split 8 x 8 cuts a site in half; the builder takes
    16 x 4 (or 4 x 16), where only the hop across the split is class O."""
    import qprop_b200 as qp

    H0, H1, H2 = qp.workloads.transmon_chain(3, 4)
    err, st = run(lib, [H0, H1, H2], 4, np.random.default_rng(2))
    assert err < 1e-14 and st[0] in (4, 16) and st[0] * st[1] == 64
    assert 4 * st[7] <= st[5] + st[6] + st[7]        # qualifies for the tiled path (<= a quarter class O)


@pytest.mark.parametrize("n,B,S", [(64, 3, 0), (256, 5, 0), (256, 2, 64), (1000, 4, 0), (512, 1, 2)])
def test_unstructured_matrices(lib, n, B, S):
    """Random sparse operators (mostly class O), one purely imaginary, duplicates of columns across
    operators, empty rows."""
    rng = np.random.default_rng(n + B)
    few = lambda k: (lambda size: np.random.RandomState(k).choice([0.5, -1.0, 2.0, 0.25, 3.0], size=size))  # noqa: E731
    A0 = sp.random(n, n, density=min(0.05, 1200 / n**2), random_state=np.random.RandomState(1), data_rvs=few(5), format="lil")
    A0[3, :] = 0
    A1 = sp.random(n, n, density=min(0.03, 800 / n**2), random_state=np.random.RandomState(2), data_rvs=few(6), format="csr") * 1j
    A2 = (sp.csr_matrix(A0) != 0).astype(np.float64) * 0.5 + sp.eye(n)   # same columns as A0 + a diagonal
    err, st = run(lib, [sp.csr_matrix(A0), A1, sp.csr_matrix(A2)], B, rng, S)
    assert err is not None and err < 1e-13
    assert st[9] == 0b010 and st[11] > 0                   # A0 / A2 share columns with unequal values: generic kind


def test_refusals(lib):
    rng = np.random.default_rng(0)
    n = 64
    mixed = sp.random(n, n, density=0.1, random_state=np.random.RandomState(3), format="csr") * (1 + 1j)
    assert run(lib, [mixed], 2, rng)[0] is None            # real and imaginary parts in one operator
    real = sp.random(n, n, density=0.1, random_state=np.random.RandomState(4), format="csr")
    assert run(lib, [real] * 4, 2, rng)[0] is None         # more than 3 operators
    assert run(lib, [sp.random(999, 999, density=0.01, format="csr")], 2, rng)[0] is None  # no split
    dense_vals = sp.csr_matrix(rng.standard_normal((128, 128)))
    assert run(lib, [dense_vals], 1, rng)[0] is None       # > 4095 distinct entries
