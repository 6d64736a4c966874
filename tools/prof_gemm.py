#!/usr/bin/env python3
"""One large DMMA GEMM (both operand layouts) on device-resident data, for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multicharge_b200 as mc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else n  # k < n: the shape of a trailing update (beta = 1)
dev = torch.device("cuda:0")
ctx = mc.default_context(0)
ext = torch.cuda.ExternalStream(ctx.stream)
a = torch.randn(k, n, dtype=torch.float64, device=dev)  # column-major n x k
b = torch.randn(k, n, dtype=torch.float64, device=dev)  # n x k (transb) or k x n, k * n doubles either way
c = torch.zeros(n, n, dtype=torch.float64, device=dev)
for tb in ("t", "n"):
    for it in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        ctx.dgemm(tb, n, n, k, -1.0, a, n, b, n if tb == "t" else k, 1.0 if k < n else 0.0, c, n)
        e1.record(ext)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        print(f"dgemm {tb} n={n} k={k}: {t*1e3:.2f} ms {2*n*n*k/t/1e12:.2f} TFLOP/s (incl. pad copies)", flush=True)
