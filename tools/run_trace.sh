#!/bin/bash
# Per-kernel times (MCHRG_B200_TRACE=1: pair A | factorisation | pair C, each over the whole batch) and the pipelined
# step time of the 100k-molecule bench mix, plus the batch parity tests.  Optional args: libs to compare (paths).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/trace.py <<'PY'
import sys, os, time
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
def call():
    mc.singlepoint_batch(model, *d_in, out=out); torch.cuda.synchronize()
call()
t0 = time.perf_counter()
for _ in range(4): call()
step = 1e3 * (time.perf_counter() - t0) / 4
os.environ["MCHRG_B200_TRACE"] = "1"
best = None
for _ in range(4):
    call()
    ms, nl = ctx.batch_trace()
    best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
print(f"{os.environ.get('MCHRG_B200_LIB', 'default lib')}: step {step:.3f} ms ({nmol/step/1e3:.3f} M/s)  pairA {best[0]:.3f}  factor {best[1]:.3f}  pairC {best[2]:.3f} ms  fails {int((out['status'] != 0).sum())}", flush=True)
PY
timeout 600 python /tmp/trace.py 2>&1 | tail -2 | tee gpurun_out/trace.log
for lib in "$@"; do MCHRG_B200_LIB=$PWD/$lib timeout 600 python /tmp/trace.py 2>&1 | tail -2 | tee -a gpurun_out/trace.log; done
