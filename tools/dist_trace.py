"""Knob sweep + per-category timing of the multi-GPU factorisation (run under torchrun, one rank per GPU):
   torchrun --nproc-per-node N tools/dist_trace.py [n]
For every MCHRG_B200_DIST_GROUP value: Cholesky of a random SPD matrix of order n, timed with
CUDA events (max over ranks), then once more with MCHRG_B200_DIST_TRACE=1 (per-category sums on stderr)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multicharge_b200 as mc  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    ctx = mc.Context(local)
    mc.init_comm_from_torch(ctx)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    m = torch.randn(n, n, generator=g, dtype=torch.float64, device=dev)
    s = m @ m.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    del m
    w = s.clone()

    def run():
        w.copy_(s)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        ctx.dpotrf(w)
        e1.record(ext)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    configs = [512, 256, 1024, 384]
    if len(sys.argv) > 2:
        configs = [int(c) for c in sys.argv[2:]]
    for grp in configs:
        os.environ["MCHRG_B200_DIST_GROUP"] = str(grp)
        run()
        best = min(run() for _ in range(3))
        if rank == 0:
            print(f"n {n} world {world} group {grp}: {best:.2f} ms (incl. ~2.5 ms pad copies)", flush=True)
        os.environ["MCHRG_B200_DIST_TRACE"] = os.environ.get("TRACE_LEVEL", "1")
        run()
        os.environ.pop("MCHRG_B200_DIST_TRACE")
        dist.barrier()
    ctx.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
