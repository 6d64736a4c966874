"""Multi-GPU factorisation check (run under torchrun, one rank per GPU):
   torchrun --nproc-per-node N tools/dist_check.py [nat]
Every rank solves the same cluster twice -- replicated single-GPU factorisation, then the block-cyclic
factorisation and row-sharded passes over the library's native NCCL communicator (mchrg_b200_comm_init_nccl) --
and compares q, E, gradient; then times both."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import multicharge_b200 as mc  # noqa: E402
from synth import cluster  # noqa: E402


def main():
    nat = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    ctx = mc.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    num, xyz = cluster(nat, 7)
    mol = mc.Structure(num, xyz, 0.0)
    model = mc.new_eeq2019_model(mol, ctx)
    txyz, tid = torch.tensor(mol.xyz, device=dev), torch.tensor(mol.id, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)

    def run():
        out = mc.api.SolveResult(cn=torch.empty(nat, **f64), qvec=torch.empty(nat, **f64), energy=torch.zeros(nat, **f64),
                                 gradient=torch.empty(nat, 3, **f64), sigma=torch.zeros(3, 3, **f64))
        mc.singlepoint_rows(model, mol, 0, nat, out=out, xyz=txyz, id=tid)
        ext.synchronize()
        return out

    ref = run()
    mc.init_comm_from_torch(ctx)
    got = run()
    err = {k: float((got[k] - ref[k]).abs().max() / ref[k].abs().max()) for k in ("qvec", "energy", "gradient")}
    worst = torch.tensor([max(err.values())], **f64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    # all ranks must hold the same result
    q0 = got["qvec"].clone()
    dist.broadcast(q0, src=0)
    same = float((q0 - got["qvec"]).abs().max())

    def timed(n=3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            run()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    t_dist = timed()

    # diagnostic: raw broadcast rate of panel-sized messages on an idle GPU
    buf = torch.empty(16384 * 256, **f64)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for b in range(64):
        dist.broadcast(buf, src=b % world)
    torch.cuda.synchronize()
    t_bc = time.perf_counter() - t0
    ctx.comm_destroy()
    t_rep = timed()
    if rank == 0:
        print(f"64 broadcasts of 33.5 MB: {1e3 * t_bc:.1f} ms ({64 * buf.numel() * 8 / t_bc / 1e9:.0f} GB/s)")
    if rank == 0:
        print(f"nat {nat} world {world}: rel err vs replicated {err}  worst over ranks {float(worst):.2e}  "
              f"rank-to-rank |dq| {same:.1e}")
        print(f"q+E+gradient: distributed factorisation {1e3 * t_dist:.1f} ms, replicated {1e3 * t_rep:.1f} ms")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
