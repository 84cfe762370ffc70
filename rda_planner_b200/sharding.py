"""Multi-GPU plumbing: planning instances are independent, so a batch is split contiguously over
the ranks of one node (one process per GPU) and NO collective runs inside the ADMM loop.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to scatter the inputs of a
solve from rank 0 and to gather trajectories back."""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """Contiguous split of `total` instances; the first `total % world` ranks get one more."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scatter_batch(tensors, total, device, src=0, group=None):
    """Rank `src` holds dict name -> tensor [total, ...]; every rank receives its shard (padded to the
    largest shard so that the collective is uniform).  Returns (dict of shards, n_valid)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(total, world, rank)
    if world == 1:
        return {k: v[lo:hi].to(device) for k, v in tensors.items()}, hi - lo
    cap = -(-total // world)
    meta = [None]
    if rank == src:
        meta[0] = {k: (tuple(v.shape[1:]), v.dtype) for k, v in tensors.items()}
    dist.broadcast_object_list(meta, src=src, group=group)
    out = {}
    for k, (shape, dtype) in meta[0].items():
        recv = torch.empty((cap,) + shape, dtype=dtype, device=device)
        chunks = None
        if rank == src:
            chunks = []
            for r in range(world):
                a, b = shard_bounds(total, world, r)
                c = torch.zeros((cap,) + shape, dtype=dtype, device=device)
                c[:b - a] = tensors[k][a:b].to(device)
                chunks.append(c)
        dist.scatter(recv, chunks, src=src, group=group)
        out[k] = recv[:hi - lo]
    return out, hi - lo


def gather_batch(shard, total, dst=0, group=None):
    """Inverse of scatter_batch for one tensor [n_local, ...]; returns the full tensor on `dst`, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return shard
    cap = -(-total // world)
    pad = torch.zeros((cap,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
    pad[:shard.shape[0]] = shard
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        a, b = shard_bounds(total, world, r)
        parts.append(bufs[r][:b - a])
    return torch.cat(parts, 0)
