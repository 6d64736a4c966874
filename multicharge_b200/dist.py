"""Host plumbing of the multi-GPU single-system path for torch.distributed launches (torchrun, one process per
GPU): the library's communicator is native NCCL (csrc/comm.cu); all it needs from the launcher is the 128-byte
unique id of rank 0 on every rank.  torch.distributed is used for exactly that broadcast (any backend: nccl on the
GPU box, gloo in the CPU tests) -- the data path never goes through Python."""


def share_unique_id(make_id, group=None):
    """Rank 0 of `group` calls make_id() (-> 128 bytes); every rank returns those bytes."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    box = [make_id() if rank == 0 else None]
    src = 0 if group is None else dist.get_global_rank(group, 0)
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def init_comm_from_torch(ctx, group=None):
    """Install the NCCL communicator of the torch.distributed world (or `group`) on the context `ctx`."""
    import torch.distributed as dist
    from .api import Context
    world = dist.get_world_size(group)
    if world == 1:
        return ctx
    uid = share_unique_id(Context.comm_unique_id, group)
    ctx.comm_init_nccl(dist.get_rank(group), world, uid)
    return ctx
