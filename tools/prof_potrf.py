#!/usr/bin/env python3
"""One blocked Cholesky (and optionally one dq/dr right-side solve) on device-resident data, for ncu launch lists."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multicharge_b200 as mc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
ctx = mc.default_context(0)
ext = torch.cuda.ExternalStream(ctx.stream)
m = torch.randn(n, n, dtype=torch.float64, device=dev)
s = m @ m.T + n * torch.eye(n, dtype=torch.float64, device=dev)
del m
w = s.clone()
for it in range(reps):
    w.copy_(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    info = ctx.dpotrf(w)
    e1.record(ext)
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"potrf n={n} info={info} {t*1e3:.2f} ms  {n**3/3/t/1e12:.2f} TFLOP/s", flush=True)
