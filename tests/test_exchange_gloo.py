"""N > 1 host logic of the multi-GPU factorisation on CPU: world_size-2 gloo run of the panel-exchange protocol
(multicharge_b200.PanelExchange, the torch.distributed side of mchrg_b200_set_exchange) driving the numpy mirror
of the device schedule (multicharge_b200.shard.block_cyclic_cholesky_host).  Every rank must end with the complete
Cholesky factor although it only ever updates its own block columns."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from multicharge_b200.api import PanelExchange
    from multicharge_b200.shard import block_cyclic_cholesky_host
    out = {}
    for n, panel in ((96, 16), (100, 32), (64, 64)):
        rng = np.random.default_rng(5 + n)  # the same matrix on every rank (replicated assembly)
        b = rng.standard_normal((n, n))
        spd = b @ b.T + n * np.eye(n)
        w = np.asfortranarray(spd).reshape(-1, order="F").copy()
        xch = PanelExchange(ctx=None, as_tensor=lambda view, nbytes: torch.from_numpy(view))
        block_cyclic_cholesky_host(w, n, panel, xch)
        low = np.tril(w.reshape(n, n).T)
        ref = np.linalg.cholesky(spd)
        nblk = (n + panel - 1) // panel
        out[(n, panel)] = (float(np.abs(low - ref).max()), xch.started, nblk, len(xch.pending))
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        ret["all"] = gathered
    dist.barrier()
    dist.destroy_process_group()


def test_panel_exchange_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for per_rank in ret["all"]:
        for (n, panel), (err, started, nblk, pending) in per_rank.items():
            assert err < 1e-11, (n, panel, err)      # complete factor on every rank
            assert started == nblk and pending == 0  # one broadcast per panel, all of them waited for
