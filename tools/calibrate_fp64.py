#!/usr/bin/env python
"""FP64 tensor-core (DMMA) calibration: cuBLAS DGEMM / ZGEMM throughput through torch.matmul,
the denominator for the dense-generator path's utilisation."""
import json

import torch

torch.cuda.set_device(0)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
N = 8192
a = torch.randn(N, N, dtype=torch.float64, device="cuda")
b = torch.randn(N, N, dtype=torch.float64, device="cuda")
ms = timeit(lambda: a @ b)
out["dgemm_8192^3_tflops"] = 2 * N**3 / ms / 1e9
for B in (16, 64, 256):
    za = torch.randn(N, N, dtype=torch.complex128, device="cuda")
    zb = torch.randn(N, B, dtype=torch.complex128, device="cuda")
    ms = timeit(lambda: za @ zb)
    out[f"zgemm_8192x8192x{B}_tflops"] = 8 * N * N * B / ms / 1e9
    out[f"zgemm_8192x8192x{B}_us"] = ms * 1e3
print(json.dumps(out))
