"""Data-parallel plumbing for the dream batch: ranks own disjoint dreams (no communication during `generate`), and the
policy/value head gradients are averaged with ONE flat all-reduce per update — what DDP does for the reference under
`accelerate` (reference dreamer4/trainers.py:1388-1396, 1436, 1446), without the per-bucket launches."""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size():
    return dist.get_world_size() if is_distributed() else 1


def rank():
    return dist.get_rank() if is_distributed() else 0


def shard_batch(global_batch: int):
    """Dreams owned by this rank: a contiguous, near-equal split of range(global_batch)."""
    w, r = world_size(), rank()
    base, rem = divmod(global_batch, w)
    start = r * base + min(r, rem)
    return start, base + (1 if r < rem else 0)


def allreduce_mean_grads_(params):
    """Averages `.grad` of `params` across ranks in place with a single flat-buffer all-reduce.
    Parameters whose grad is None on this rank contribute zeros (and receive the average)."""
    if not is_distributed():
        return 0
    params = [p for p in params if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world_size())
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)
    return flat.numel() * flat.element_size()
