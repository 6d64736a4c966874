#!/bin/bash
# Residency / phase-clock experiments on the batch factorisation kernel (one GPU).
#   1. uniform-size batches, factor kernel time vs molecules resident per SM (MCHRG_B200_FACTOR_SMEM_KB)
#   2. per-warp phase clocks of sampled molecules (-DLL_PROF build, libmchrg_b200_llprof.so)
#   3. MCHRG_B200_FACTOR_GROUP / MCHRG_B200_CHUNKS sweeps on the bench mix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/llexp.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import numpy as np, torch
import multicharge_b200 as mc
from synth import batch
ctx = mc.Context(0)
model = mc.new_eeq2019_batch_model(ctx)
dev = torch.device("cuda", 0)
def run(nmol, nmin, nmax, reps=3):
    offsets, num, xyz, charge = batch(nmol, 5, nmin, nmax)
    d_in = [torch.tensor(a, device=dev) for a in (offsets, num, xyz, charge)]
    ntot = int(offsets[-1])
    out = dict(qvec=torch.empty(ntot, dtype=torch.float64, device=dev), energy=torch.empty(ntot, dtype=torch.float64, device=dev),
               gradient=torch.empty(ntot, 3, dtype=torch.float64, device=dev), status=torch.empty(nmol, dtype=torch.int32, device=dev))
    best = None
    for _ in range(reps):
        mc.singlepoint_batch(model, *d_in, out=out)
        torch.cuda.synchronize()
        ms, nl = ctx.batch_trace()
        best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
    assert int((out["status"] != 0).sum()) == 0
    return best
mode = sys.argv[1]
os.environ["MCHRG_B200_TRACE"] = "1"
if mode == "resid":
    for n in (48, 100, 150, 200):
        nmol = 16000 if n <= 100 else 8000
        for kb in (0, 44, 56, 75, 113, 200):
            os.environ["MCHRG_B200_FACTOR_SMEM_KB"] = str(kb)
            a, f, c = run(nmol, n, n)
            print(f"n {n} nmol {nmol} smem_kb {kb}: pairA {a:.3f} factor {f:.3f} pairC {c:.3f} ms -> factor {1e3*f/nmol:.3f} us/molecule", flush=True)
    os.environ.pop("MCHRG_B200_FACTOR_SMEM_KB")
elif mode == "prof":
    run(20000, 30, 200, reps=1)
elif mode == "sweep":
    for g in (13, 7, 4, 2):
        os.environ["MCHRG_B200_FACTOR_GROUP"] = str(g)
        print("group", g, run(100000, 30, 200), flush=True)
    os.environ.pop("MCHRG_B200_FACTOR_GROUP")
    for ch in (8, 16, 32, 64):
        os.environ["MCHRG_B200_CHUNKS"] = str(ch)
        print("chunks", ch, run(100000, 30, 200), flush=True)
PY
timeout 600 python /tmp/llexp.py resid > gpurun_out/ll_resid.log 2>&1
MCHRG_B200_LIB=$PWD/multicharge_b200/libmchrg_b200_llprof.so timeout 300 python /tmp/llexp.py prof > gpurun_out/ll_prof.log 2>&1
timeout 600 python /tmp/llexp.py sweep > gpurun_out/ll_sweep.log 2>&1
tail -30 gpurun_out/ll_resid.log; tail -5 gpurun_out/ll_sweep.log; grep -c "\[ll\]" gpurun_out/ll_prof.log
