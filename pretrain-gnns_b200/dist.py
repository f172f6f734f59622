"""Data-parallel plumbing: shard graphs across ranks, one all-reduce of a flat fp32 gradient buffer per step.

The reference has no parallelism at all; graphs in a batch are independent, so the
path shards by graph with no activation exchange and the only collective is the gradient all-reduce
(section 8(e)).  Each rank's BatchNorm statistics are local (like DDP without SyncBN).  Backend: NCCL on
GPUs (NVLink 5 / NVSwitch), gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_graphs(num_graphs: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a global batch's graphs owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_graphs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradAllReducer:
    """Packs every parameter's .grad into one flat fp32 buffer, all-reduces it once, unpacks the mean.

    7.4 MB for chem GIN: a single latency-bound collective instead of 42 small ones."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(self.sizes), self.params)]

    def all_reduce_mean(self):
        world = dist.get_world_size(self.group)
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
        torch._foreach_copy_([p.grad for p in self.params], self.views)
        return self.flat
