"""N > 1 host logic of the multi-GPU single-system path on CPU, world_size 2 and 3 over gloo:
  * the numpy mirror of the device factorisation schedule (multicharge_b200.shard.block_cyclic_cholesky_host ==
    dpotrf_dist in csrc/dense.cu): every rank starts with ONLY the column groups it owns (everything else NaN), the
    broadcasts are real gloo broadcasts of the contiguous sub-panel ranges, and every rank must end with the complete
    Cholesky factor;
  * the unique-id hand-over of multicharge_b200.dist (rank 0's 128 bytes reach every rank);
  * the row partitions (shard.row_bounds) cover all atoms disjointly.
The CUDA side of the same schedule is checked on the GPU by tests/test_gpu_dist.py (local backend, one GPU)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from multicharge_b200.dist import share_unique_id
    from multicharge_b200.shard import block_cyclic_cholesky_host, group_width

    def bcast(view, root):
        dist.broadcast(torch.from_numpy(view), src=root)

    out = {}
    for n, group, sub in ((96, 32, 16), (100, 48, 16), (64, 16, 16), (130, 64, 32)):
        rng = np.random.default_rng(5 + n)
        b = rng.standard_normal((n, n))
        spd = b @ b.T + n * np.eye(n)
        a = spd.copy()
        for g in range((n + group - 1) // group):  # a rank holds only its own column groups on entry
            if g % world != rank:
                a[:, g * group:(g + 1) * group] = np.nan
        w = np.asfortranarray(a).reshape(-1, order="F").copy()
        nb = block_cyclic_cholesky_host(w, n, group, sub, rank, world, bcast)
        low = np.tril(w.reshape(n, n).T)
        ref = np.linalg.cholesky(spd)
        out[(n, group, sub)] = (float(np.abs(low - ref).max()), nb)
    uid = share_unique_id(lambda: bytes(range(128)))
    out["uid_ok"] = uid == bytes(range(128))
    out["gw"] = group_width(16512, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        ret["all"] = gathered
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_block_cyclic_schedule_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 90) + world
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for per_rank in ret["all"]:
        assert per_rank.pop("uid_ok")
        assert per_rank.pop("gw") == 512
        for (n, group, sub), (err, nb) in per_rank.items():
            assert err < 1e-11, (n, group, sub, err)  # complete factor on every rank
            expect = sum(-(-(min(n, (g + 1) * group) - g * group) // sub) for g in range((n + group - 1) // group))
            assert nb == expect                       # one broadcast per sub-panel


def test_row_bounds_cover():
    from multicharge_b200.shard import row_bounds
    for nat, world in ((16384, 8), (1000, 3), (300, 2), (129, 4)):
        b = row_bounds(nat, world)
        assert b[0] == 0 and b[-1] == nat and np.all(np.diff(b) >= 0)
        assert all(x % 128 == 0 for x in b[1:-1])
