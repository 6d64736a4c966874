"""N > 1 host logic on CPU: world_size-2 gloo run of the molecule partitioning used by bench.py / multi-GPU
batches.  Every rank computes the same partition, takes its shard, 'solves' it with the CPU oracle (standing in
for the device here: there is no GPU in this test) and the shards are all-gathered and checked to be a disjoint
cover that reproduces the single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    import oracle as orc
    from multicharge_b200 import shard
    from synth import batch
    offsets, num, xyz, charge = batch(40, 123, 5, 60)
    parts = shard.partition_by_cost(np.diff(offsets), world)
    mine = parts[rank]
    so, sn, sx, sc = shard.gather_shard(offsets, num, xyz, charge, mine)
    res = orc.eeq2019_singlepoint_batch(so, sn, sx, sc, threads=1)
    # all-gather shard sizes and charges
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(mine)], dtype=torch.int64))
    total = int(sum(int(c.item()) for c in counts))
    qsum = torch.tensor([float(res["qvec"].sum())], dtype=torch.float64)
    dist.all_reduce(qsum)
    obj = [None] * world
    dist.all_gather_object(obj, (mine.tolist(), so.tolist(), res["qvec"].tolist(), res["status"].tolist()))
    if rank == 0:
        out = dict(qvec=np.zeros(len(num)), energy=None, gradient=None, status=np.full(len(charge), -1, dtype=np.int32))
        for idx, soff, q, st in obj:
            shard.scatter_results(len(charge), offsets, idx, soff, dict(qvec=np.array(q), status=np.array(st)), out)
        full = orc.eeq2019_singlepoint_batch(offsets, num, xyz, charge, threads=1)
        ret["total"] = total
        ret["cover"] = sorted(sum((o[0] for o in obj), []))
        ret["maxdiff"] = float(np.abs(out["qvec"] - full["qvec"]).max())
        ret["status_ok"] = bool((out["status"] == 0).all())
        ret["qsum"] = float(qsum.item())
        ret["charge_sum"] = float(charge.sum())
        ret["balance"] = [float(shard.molecule_cost(np.diff(offsets)[p]).sum()) for p in parts]
    dist.barrier()
    dist.destroy_process_group()


def test_partition_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29400 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["total"] == 40
    assert ret["cover"] == list(range(40))
    assert ret["maxdiff"] == 0.0 and ret["status_ok"]
    assert abs(ret["qsum"] - ret["charge_sum"]) < 1e-9
    b = ret["balance"]
    assert abs(b[0] - b[1]) / max(b) < 0.05


def test_partition_properties():
    from multicharge_b200 import shard
    rng = np.random.default_rng(0)
    sizes = rng.integers(30, 201, size=1000)
    for world in (1, 2, 4, 8):
        parts = shard.partition_by_cost(sizes, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(1000))
        loads = [shard.molecule_cost(sizes[p]).sum() for p in parts]
        assert (max(loads) - min(loads)) / max(loads) < 0.01
