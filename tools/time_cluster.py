#!/usr/bin/env python3
"""Device-resident timing of the single-system path (configs[2]): N-atom cluster, q+E+gradient and full dq/dr."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import multicharge_b200 as mc  # noqa: E402
from multicharge_b200 import _lib  # noqa: E402
from synth import cluster  # noqa: E402

dev = torch.device("cuda:0")
ctx = mc.default_context(0)
ext = torch.cuda.ExternalStream(ctx.stream)
out = {}
for nat in [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]:
    num, xyz = cluster(nat, 1)
    mol = mc.Structure(num, xyz, 0.0)
    model = mc.new_eeq2019_model(mol)
    txyz = torch.tensor(mol.xyz, device=dev)
    tid = torch.tensor(mol.id, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    bufs = dict(cn=torch.empty(nat, **f64), qloc=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64), qvec=torch.empty(nat, **f64),
                dqdL=torch.empty(nat, 3, 3, **f64))
    if not os.environ.get("NO_DQ"):
        bufs["dqdr"] = torch.empty(nat, nat, 3, **f64)
    s = mol._c(xyz=txyz, id=tid)
    p = lambda k: bufs[k].data_ptr()  # noqa: E731

    def run(dq):
        rc = _lib.load().mchrg_b200_singlepoint(model._h, C.byref(s), p("cn"), p("qloc"), p("energy"), p("gradient"), p("sigma"),
                                                p("qvec"), p("dqdr") if dq else None, p("dqdL") if dq else None)
        assert rc == 0, rc

    for dq in ((False,) if os.environ.get("NO_DQ") else (False, True)):
        run(dq)
        best = 1e30
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            run(dq)
            e1.record(ext)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        key = f"n{nat}_{'dqdr' if dq else 'grad'}"
        out[key + "_s"] = best
        flops = nat ** 3 / 3 + (6.0 * nat ** 3 if dq else 0.0)
        out[key + "_tflops"] = flops / best / 1e12
    del bufs
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
