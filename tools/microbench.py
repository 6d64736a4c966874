#!/usr/bin/env python3
"""Quick device-side timings of the building blocks (not the bench contract; see bench.py)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import multicharge_b200 as mc  # noqa: E402
from synth import cluster  # noqa: E402


def timed(fn, stream, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    out = {}
    dev = torch.device("cuda:0")
    ctx = mc.default_context(0)
    ext = torch.cuda.ExternalStream(ctx.stream)
    cur = torch.cuda.current_stream()
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.zeros(n, n, dtype=torch.float64, device=dev)
    t = timed(lambda: torch.matmul(a, b, out=c), cur, reps=5)
    out["cublas_dgemm_8192_tflops"] = 2 * n ** 3 / t / 1e12
    for tb in ("t", "n"):
        t = timed(lambda: ctx.dgemm(tb, n, n, n, 1.0, a, n, b, n, 0.0, c, n), ext, reps=3)
        out[f"mchrg_dgemm_{tb}_8192_tflops_incl_padcopy"] = 2 * n ** 3 / t / 1e12
    for n in (4096, 8192, 16384):
        m = torch.randn(n, n, dtype=torch.float64, device=dev)
        s = m @ m.T + n * torch.eye(n, dtype=torch.float64, device=dev)
        w = s.clone()
        def run():
            w.copy_(s)
            ctx.dpotrf(w)
        t = timed(run, ext, reps=2)
        out[f"mchrg_dpotrf_{n}_s_incl_copies"] = t
        out[f"mchrg_dpotrf_{n}_tflops"] = n ** 3 / 3 / t / 1e12
        t = timed(lambda: torch.linalg.cholesky(s), cur, reps=2)
        out[f"cusolver_potrf_{n}_tflops"] = n ** 3 / 3 / t / 1e12
        del m, s, w
    for nat in (1024, 4096, 8192, 16384):
        num, xyz = cluster(nat, 1)
        mol = mc.Structure(num, xyz, 0.0)
        model = mc.new_eeq2019_model(mol)
        for want in (("qvec", "energy", "gradient"), ("qvec", "energy", "gradient", "dqdr")):
            if nat > 8192 and "dqdr" in want and os.environ.get("MB_BIG", "1") != "1":
                continue
            t0 = time.time()
            model.singlepoint(mol, want=want)
            t1 = time.time()
            model.singlepoint(mol, want=want)
            t2 = time.time()
            out[f"singlepoint_{nat}_{'dqdr' if 'dqdr' in want else 'grad'}_wall_s"] = [t1 - t0, t2 - t1]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
