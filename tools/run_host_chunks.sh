#!/bin/bash
# End-to-end batch step (pinned host buffers in and out) for several MCHRG_B200_HOST_CHUNKS settings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/hostchunks.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import numpy as np, torch
import multicharge_b200 as mc
from synth import batch
ctx = mc.Context(0)
model = mc.new_eeq2019_batch_model(ctx)
nmol = 100000
offsets, num, xyz, charge = batch(nmol, 5, 30, 200)
ntot = int(offsets[-1])
h_in = [torch.from_numpy(a).pin_memory() for a in (offsets, num, xyz, charge)]
h_out = dict(qvec=torch.empty(ntot, dtype=torch.float64).pin_memory(), energy=torch.empty(ntot, dtype=torch.float64).pin_memory(),
             gradient=torch.empty(ntot, 3, dtype=torch.float64).pin_memory(), status=torch.empty(nmol, dtype=torch.int32).pin_memory())
n_in = [t.numpy() for t in h_in]
n_out = {k: v.numpy() for k, v in h_out.items()}
def step():
    mc.singlepoint_batch(model, *n_in, out=n_out)
for hc in sys.argv[1:]:
    if hc != "0": os.environ["MCHRG_B200_HOST_CHUNKS"] = hc
    else: os.environ.pop("MCHRG_B200_HOST_CHUNKS", None)
    step(); step()
    t0 = time.perf_counter()
    for _ in range(4): step()
    ms = 1e3 * (time.perf_counter() - t0) / 4
    print(f"HOST_CHUNKS={hc:>3s} (0 = default): {ms:7.3f} ms/step  {nmol/ms/1e3:6.3f} M molecules/s end to end", flush=True)
PY
timeout 600 python /tmp/hostchunks.py "$@" 2>&1 | tee gpurun_out/host_chunks.log
