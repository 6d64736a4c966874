#!/bin/bash
# Co-residency experiments: the whole pipelined batch step (pair A | factorisation | pair C on three streams) with the
# factorisation kernel's residency capped through its shared-memory request, so that pair CTAs share the SMs with it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/mix.py <<'PY'
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
def step_ms(reps=4):
    mc.singlepoint_batch(model, *d_in, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        mc.singlepoint_batch(model, *d_in, out=out)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps
ref = None
for env in sys.argv[1:]:
    kv = dict(x.split("=") for x in env.split("+") if x)
    for k, v in kv.items(): os.environ[k] = v
    ms = step_ms()
    q = out["qvec"].cpu().numpy().copy()
    if ref is None: ref = q
    print(f"{env or 'default':60s} {ms:8.3f} ms/step  {nmol/ms/1e3:6.3f} M molecules/s  max|dq| {np.abs(q-ref).max():.1e}", flush=True)
    for k in kv: os.environ.pop(k)
PY
timeout 900 python /tmp/mix.py "$@" > gpurun_out/mix.log 2>&1
cat gpurun_out/mix.log
