"""N > 1 host logic of the multi-GPU single-system path on CPU, world_size 2 and 3 over gloo:
  * the numpy mirror of the device factorisation schedule (multicharge_b200.shard.block_cyclic_cholesky_host ==
    dpotrf_dist in csrc/dense.cu): every rank starts with ONLY the column groups it owns (everything else NaN), the
    broadcasts are real gloo broadcasts of the contiguous sub-panel ranges, and every rank must end with the complete
    Cholesky factor;
  * the unique-id hand-over of multicharge_b200.dist (rank 0's 128 bytes reach every rank);
  * the row partitions (shard.row_bounds) cover all atoms disjointly;
  * the numpy mirror of the shared periodic assembly (shard.periodic_matrix_shared_host == pbc_amat with `shard`): pair
    tiles round robin, row blocks of the reciprocal-space matrix, diagonal by row shares, ONE gloo all-reduce.
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
    # periodic assembly shared between the ranks (pbc_amat): pair tiles round robin, row blocks of the reciprocal-space
    # matrix, the diagonal by row shares, ONE all-reduce -- every rank must end with the one-rank matrix
    from multicharge_b200.shard import periodic_matrix_shared_host

    def allreduce(arr):
        t = torch.from_numpy(arr)
        dist.all_reduce(t)

    for nat in (70, 300):
        rng = np.random.default_rng(nat)
        pos = rng.uniform(0.0, 9.0, size=(nat, 3))
        rec = rng.standard_normal((nat, 16))
        recip = rec @ rec.T
        pair = lambda i, j: 1.0 / (1.0 + np.sum((pos[i] - pos[j]) ** 2))  # noqa: E731
        diag = lambda i: 2.0 + 0.01 * i  # noqa: E731
        got = periodic_matrix_shared_host(nat, pair, recip, diag, rank, world, allreduce)
        full = periodic_matrix_shared_host(nat, pair, recip, diag, 0, 1, lambda arr: None)
        out[("pbc", nat)] = float(np.abs(got - full).max())
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
        for nat in (70, 300):  # shared periodic assembly == one-rank matrix (sums of at most two non-zero terms + zeros)
            assert per_rank.pop(("pbc", nat)) < 1e-13
        for (n, group, sub), (err, nb) in per_rank.items():
            assert err < 1e-11, (n, group, sub, err)  # complete factor on every rank
            expect = sum(-(-(min(n, (g + 1) * group) - g * group) // sub) for g in range((n + group - 1) // group))
            assert nb == expect                       # one broadcast per sub-panel


def test_periodic_shares_cover():
    """Every pair tile and every row block of the reciprocal-space product belongs to exactly one rank."""
    from multicharge_b200.shard import pair_tiles, rowblock_share, tile_share
    for nat, world in ((4096, 8), (343, 3), (1728, 2), (31, 4)):
        nt = len(pair_tiles(nat))
        owned = sorted(b for r in range(world) for b in tile_share(nt, r, world))
        assert owned == list(range(nt))
        nblk = -(-nat // 128)
        rows = sorted(b for r in range(world) for b in rowblock_share(nblk, r, world))
        assert rows == list(range(nblk))


def test_row_bounds_cover():
    from multicharge_b200.shard import row_bounds
    for nat, world in ((16384, 8), (1000, 3), (300, 2), (129, 4)):
        b = row_bounds(nat, world)
        assert b[0] == 0 and b[-1] == nat and np.all(np.diff(b) >= 0)
        assert all(x % 128 == 0 for x in b[1:-1])
