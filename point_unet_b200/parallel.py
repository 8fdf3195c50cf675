"""Data-parallel plumbing for PointSegment on one NVSwitch box.

* Training: batch-sharded replicas; the ONLY exchange is an all-reduce (average) of the gradients, done on one
  flat fp32 buffer that every parameter's ``.grad`` aliases (4.99 M params = 20 MB, a single NCCL call over
  NVLink 5 -- NVSwitch gives every peer full bandwidth, so there is nothing to gain from bucketing by link).
  Batch-norm statistics stay per replica, exactly like a batch-4 single-GPU run of the reference.
* KNN / inference: clouds or volumes are dealt round-robin to ranks, no communication ("replicas only").
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradBucket:
    """One contiguous gradient buffer; ``p.grad`` of every parameter is a view into it."""

    def __init__(self, params, device=None):
        self.params = list(params)
        device = device if device is not None else self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def all_reduce_mean(self, group=None):
        """Average over ranks.  NCCL: a single AVG all-reduce; other backends (gloo in the CPU tests): SUM then scale."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / world)


def shard_round_robin(n_items: int, rank: int, world: int) -> list[int]:
    """Indices of the clouds / volumes owned by ``rank`` (64 volumes over 8 GPUs -> 8 each; no communication)."""
    return list(range(rank, n_items, world))


def max_over_ranks(value_ms: float, device) -> float:
    """Device-timed durations are reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value_ms
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
