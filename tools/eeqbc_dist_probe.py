"""EEQ-BC q+E+gradient of one cluster over the ranks of a torchrun launch, with the library's traces on (where the time
goes when the system is shared): torchrun --nproc-per-node N tools/eeqbc_dist_probe.py [nat]"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import multicharge_b200 as mc  # noqa: E402
from synth import cluster, synthetic_vdw_en  # noqa: E402

nat = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
ctx = mc.Context(local)
if world > 1:
    mc.init_comm_from_torch(ctx)
num, xyz = cluster(nat, 20250105)
mol = mc.Structure(num, xyz, 0.0)
rvdw, en = synthetic_vdw_en(mol.num)
model = mc.new_eeqbc2025_model(mol, rvdw, en * 3.98, ctx)
txyz, tid = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
f64 = dict(dtype=torch.float64, device=dev)
o = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                       gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64))
for it in range(3):
    if it == 2:
        os.environ["MCHRG_B200_DIST_TRACE"] = "1"
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    mc.singlepoint_rows(model, mol, 0, nat, out=o, xyz=txyz, id=tid)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"nat {nat} world {world} run {it}: {time.perf_counter() - t0:.3f} s  sum q {float(o['qvec'].sum()):.2e}", flush=True)
ctx.comm_destroy()
dist.barrier()
dist.destroy_process_group()
