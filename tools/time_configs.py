#!/usr/bin/env python3
"""Wall/device timing of BASELINE.json configs[3] (periodic N = 4096, charges + dq/dL) and of a non-periodic
EEQ-BC cluster (configs[4] at single-GPU size), device-resident outputs."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import multicharge_b200 as mc  # noqa: E402
from multicharge_b200 import _lib  # noqa: E402
from synth import cluster, periodic_cell  # noqa: E402

dev = torch.device("cuda:0")
ctx = mc.default_context(0)
ext = torch.cuda.ExternalStream(ctx.stream)
out = {}


def timed(fn, reps=2):
    fn()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        fn()
        e1.record(ext)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def bufs(nat, dq):
    f64 = dict(dtype=torch.float64, device=dev)
    b = dict(cn=torch.empty(nat, **f64), qloc=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
             gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64), qvec=torch.empty(nat, **f64))
    if dq:
        b["dqdr"] = torch.empty(nat, nat, 3, **f64)
        b["dqdL"] = torch.empty(nat, 3, 3, **f64)
    return b


def run_sp(model, mol, b):
    s = mol._c(xyz=b["_xyz"], id=b["_id"])
    p = lambda k: b[k].data_ptr() if k in b else None  # noqa: E731
    rc = _lib.load().mchrg_b200_singlepoint(model._h, C.byref(s), p("cn"), p("qloc"), p("energy"), p("gradient"), p("sigma"),
                                            p("qvec"), p("dqdr"), p("dqdL"))
    assert rc == 0, rc


ncell = int(os.environ.get("NCELL", "16"))
for shear in (0.0, 0.1):
    num, xyz, lat = periodic_cell(ncell, 20250104, shear=shear)
    mol = mc.Structure(num, xyz, 0.0, lattice=lat)
    model = mc.new_eeq2019_model(mol)
    for dq in (False, True):
        b = bufs(mol.nat, dq)
        b["_xyz"], b["_id"] = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
        t = timed(lambda: run_sp(model, mol, b))
        out[f"pbc_n{mol.nat}_{'tri' if shear else 'cub'}_{'dqdr' if dq else 'grad'}_s"] = t
        del b
nbc = int(os.environ.get("NBC", "8192"))
import oracle as orc  # noqa: E402
num, xyz = cluster(nbc, 20250105)
mol = mc.Structure(num, xyz, 0.0)
rvdw, en = orc.synthetic_vdw_en(mol.num)
model = mc.new_eeqbc2025_model(mol, rvdw, en * 3.98)
for dq in ((False, True) if os.environ.get("NBC_DQ", "1") == "1" else (False,)):
    b = bufs(mol.nat, dq)
    b["_xyz"], b["_id"] = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
    t = timed(lambda: run_sp(model, mol, b), reps=int(os.environ.get("NBC_REPS", "2")))
    out[f"eeqbc_n{mol.nat}_{'dqdr' if dq else 'grad'}_s"] = t
    del b
print(json.dumps(out, indent=1))
