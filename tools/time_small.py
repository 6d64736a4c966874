#!/usr/bin/env python3
"""Latency of get_charges (charges only) for small molecules: fused one-launch route vs the general single-system path."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import multicharge_b200 as mc  # noqa: E402
from synth import cluster  # noqa: E402

out = {}
for nat in (16, 64, 200):
    num, xyz = cluster(nat, 5)
    mol = mc.Structure(num, xyz, 0.0)
    model = mc.new_eeq2019_model(mol)
    for tag, env in (("fused", None), ("general", "0")):
        if env is None:
            os.environ.pop("MCHRG_B200_SMALL_FUSED", None)
        else:
            os.environ["MCHRG_B200_SMALL_FUSED"] = env
        for _ in range(20):
            mc.get_charges(model, mol)
        t0 = time.perf_counter()
        reps = 300
        for _ in range(reps):
            mc.get_charges(model, mol)
        out[f"n{nat}_{tag}_us"] = (time.perf_counter() - t0) / reps * 1e6
print(json.dumps(out, indent=1))
