#!/usr/bin/env python3
"""Periodic EEQ2019 single points (configs[3]: cubic cell, N = 16^3 = 4096 by default), q + E + gradient + sigma,
device resident; prints the CUDA-event time per call.  Run under `ncu --metrics gpu__time_duration.sum` for the
per-kernel split."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import multicharge_b200 as mc  # noqa: E402
from synth import periodic_cell  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
ctx = mc.default_context(0)
ext = torch.cuda.ExternalStream(ctx.stream)
num, xyz, lat = periodic_cell(ncell, 20250104, shear=0.0)
mol = mc.Structure(num, xyz, 0.0, lattice=lat)
model = mc.new_eeq2019_model(mol)
for it in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    res = model.singlepoint(mol, want=("qvec", "energy", "gradient"))
    e1.record(ext)
    torch.cuda.synchronize()
    print(f"periodic N = {mol.nat}: q + E + gradient {e0.elapsed_time(e1):.2f} ms (host arrays in / out)", flush=True)
