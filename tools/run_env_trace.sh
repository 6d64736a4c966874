#!/bin/bash
# Per-kernel trace times (pair A | factorisation | pair C over the whole 100k-molecule bench mix) under environment
# settings given as arguments ("A=1+B=2" sets two variables; "" = defaults).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/envtrace.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import numpy as np, torch
import multicharge_b200 as mc
from synth import batch
ctx = mc.Context(0)
model = mc.new_eeq2019_batch_model(ctx)
dev = torch.device("cuda", 0)
nmol = 100000
offsets, num, xyz, charge = batch(nmol, 5, 30, 200)
d_in = [torch.tensor(a, device=dev) for a in (offsets, num, xyz, charge)]
ntot = int(offsets[-1])
out = dict(qvec=torch.empty(ntot, dtype=torch.float64, device=dev), energy=torch.empty(ntot, dtype=torch.float64, device=dev),
           gradient=torch.empty(ntot, 3, dtype=torch.float64, device=dev), status=torch.empty(nmol, dtype=torch.int32, device=dev))
os.environ["MCHRG_B200_TRACE"] = "1"
ref = None
for env in sys.argv[1:]:
    kv = dict(x.split("=") for x in env.split("+") if x)
    for k, v in kv.items(): os.environ[k] = v
    best = None
    for _ in range(4):
        mc.singlepoint_batch(model, *d_in, out=out); torch.cuda.synchronize()
        ms, nl = ctx.batch_trace()
        best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
    q = out["qvec"].cpu().numpy().copy()
    if ref is None: ref = q
    print(f"{env or 'default':50s} pairA {best[0]:.3f}  factor {best[1]:.3f}  pairC {best[2]:.3f} ms  fails {int((out['status'] != 0).sum())}  max|dq| {np.abs(q - ref).max():.1e}", flush=True)
    for k in kv: os.environ.pop(k)
PY
timeout 900 python /tmp/envtrace.py "$@" 2>&1 | tee gpurun_out/env_trace.log | tail -30
