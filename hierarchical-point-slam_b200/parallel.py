"""Multi-GPU mapping: one process per GPU, the ray batch is sharded contiguously across ranks,
cloud / feature tables / decoders are replicated, and one gradient all-reduce per iteration (NCCL over
NVLink on the GPU box, gloo in the CPU tests) makes every rank apply the identical optimiser step.

The reference has no distributed code at all; this is the data-parallel form of Mapper.py:783-860.
"""
import torch


def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of n rays for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rank, world, *tensors):
    """Slice every per-ray tensor (first dim = rays) to this rank's shard."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return tuple(t[lo:hi] for t in tensors)


def allreduce_grads(tensors, dist, small_numel=1 << 16, group=None):
    """Sum `.grad` of every tensor across ranks (missing grads count as zero).

    Large gradients (feature tables) are reduced in place; small ones (decoder parameters, pose,
    exposure) are flattened into one bucket so the collective count per iteration stays at O(#tables)."""
    small = []
    for t in tensors:
        if t.grad is None:
            t.grad = torch.zeros_like(t)
        if t.grad.numel() >= small_numel:
            dist.all_reduce(t.grad, op=dist.ReduceOp.SUM, group=group)
        else:
            small.append(t)
    if small:
        flat = torch.cat([t.grad.reshape(-1) for t in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for t in small:
            n = t.grad.numel()
            t.grad.copy_(flat[off:off + n].view_as(t.grad))
            off += n


def allreduce_scalars(values, dist, group=None):
    """Sum a few python/0-d values across ranks (tracker outlier threshold: sum and count of tmp,
    Tracker.py:217-218)."""
    t = torch.stack([torch.as_tensor(v, dtype=torch.float64).reshape(()) for v in values])
    dev = values[0].device if torch.is_tensor(values[0]) else "cpu"
    t = t.to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
