"""Batch data-parallel plumbing (new: the reference is single-device).

One process per GPU; each rank holds a batch shard.  Exactness: ranks all-reduce(SUM) the local
feature column sums before the distance losses, every per-sample loss term is normalised by the GLOBAL batch, and
parameter gradients are all-reduced with SUM (not AVG).  torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests) is the transport."""
from __future__ import annotations

import torch
import torch.distributed as dist


class Comm:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised')
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.calls = 0
        # set by StepRunner while it captures a step piecewise: called instead of the collective (which it then issues
        # itself between two graph segments)
        self.capture_hook = None

    def all_reduce_sum(self, t: torch.Tensor):
        self.calls += 1
        if self.capture_hook is not None:
            self.capture_hook('all_reduce', t)
            return
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def begin_deferred(self, tag):
        """Everything up to end_deferred() (a gradient all-reduce and the optimizer update that consumes it) may run on
        a side stream, overlapped with whatever the caller enqueues next; only honoured under piecewise graph capture."""
        if self.capture_hook is not None:
            self.capture_hook('defer_begin', tag)

    def end_deferred(self):
        if self.capture_hook is not None:
            self.capture_hook('defer_end', None)

    def all_reduce_sum_now(self, t: torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_sum_partial(self, t: torch.Tensor, slots):
        """Sums only the given slots across ranks (per-sample loss terms); the others are already global."""
        idx = torch.tensor(list(slots), device=t.device)
        part = t[idx].clone()
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        t[idx] = part


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows [rank*B/W, (rank+1)*B/W) of a global batch."""
    B = t.shape[0]
    if B % world:
        raise ValueError(f'global batch {B} is not divisible by world size {world}')
    per = B // world
    return t[rank * per:(rank + 1) * per]
