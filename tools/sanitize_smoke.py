#!/usr/bin/env python3
"""Small run of every device path for compute-sanitizer (memcheck): EEQ2019 0-D / periodic, EEQ-BC, batch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import multicharge_b200 as mc  # noqa: E402
import oracle as orc  # noqa: E402
from synth import batch, cluster, periodic_cell  # noqa: E402

want = ("qvec", "energy", "gradient", "dqdr")
num, xyz = cluster(150, 1)
mol = mc.Structure(num, xyz, 1.0)
r = mc.new_eeq2019_model(mol).singlepoint(mol, want=want)
print("eeq 0-d", float(r.qvec.sum()))
num, xyz, lat = periodic_cell(3, 2, shear=0.1)
mol = mc.Structure(num, xyz, 0.0, lattice=lat)
r = mc.new_eeq2019_model(mol).singlepoint(mol, want=want)
print("eeq pbc", float(r.qvec.sum()))
num, xyz = cluster(70, 3)
mol = mc.Structure(num, xyz, 0.0)
rvdw, en = orc.synthetic_vdw_en(mol.num)
r = mc.new_eeqbc2025_model(mol, rvdw, en * 3.98).singlepoint(mol, want=want)
print("eeqbc", float(r.qvec.sum()))
offsets, num, xyz, charge = batch(40, 4, 5, 208)
out = mc.singlepoint_batch(mc.new_eeq2019_batch_model(), offsets, num, xyz, charge)
print("batch", int((out["status"] != 0).sum()), float(out["qvec"].sum()), float(charge.sum()))
