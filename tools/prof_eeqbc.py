#!/usr/bin/env python3
"""EEQ-BC single points (configs[4] path) on a synthetic cluster, q + E + gradient, device resident; prints the
CUDA-event time per call.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split.
Usage: prof_eeqbc.py [nat] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import multicharge_b200 as mc  # noqa: E402
from synth import cluster, synthetic_vdw_en  # noqa: E402

nat = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
ctx = mc.default_context(0)
ext = torch.cuda.ExternalStream(ctx.stream)
num, xyz = cluster(nat, 20250105)
mol = mc.Structure(num, xyz, 0.0)
rvdw, en = synthetic_vdw_en(mol.num)
model = mc.new_eeqbc2025_model(mol, rvdw, en * 3.98, ctx)
f64 = dict(dtype=torch.float64, device=dev)
txyz, tid = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
o = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                       gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64))
for it in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    mc.singlepoint_rows(model, mol, 0, nat, out=o, xyz=txyz, id=tid)
    e1.record(ext)
    torch.cuda.synchronize()
    print(f"EEQ-BC N = {nat}: q + E + gradient {e0.elapsed_time(e1):.2f} ms", flush=True)
