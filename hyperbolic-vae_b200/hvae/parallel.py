"""Data parallelism for the train step: one process per GPU, parameters replicated, batch
rows sharded, ONE all-reduce per step over a flat pre-allocated fp32 gradient bucket.

The reference has no distributed code at all (devices=1 everywhere: training/trainer_mnist.py:19); this is
the part the build adds.  Every hot-path op is row-independent, so the only cross-rank exchange is the
parameter-gradient sum.  Loss reductions decide the scale rule:
  batch-SUM losses (model B: models/vae_hyperbolic.py:216,219)  -> SUM, no rescale reproduces 1-GPU grads;
  batch-MEAN losses (models A/C: ...gyroplane_decoder.py:151)   -> each rank's mean is over its shard, so
                                                                   SUM then divide by world size (AVG).
The payload is 0.2-4 MB: latency-bound on NVLink 5/NVSwitch.  The bucket can be cut into an EARLY segment (the
parameters whose gradients autograd finishes first — the decoder tail) and the rest: the early segment's all-reduce
is issued from a gradient hook on a side stream and runs under the remaining backward; the rest follows at the end
(`TrainStep` picks the cut from the observed completion order of the first warm-up backward).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Owns one contiguous buffer; every parameter's .grad is a view into it, so backward writes straight
    into the bucket and the all-reduce needs no gather/scatter copies."""

    def __init__(self, params: Iterable[torch.nn.Parameter], early: Iterable[torch.nn.Parameter] = ()):
        """early: parameters to place first in the buffer (one contiguous segment [0, split))."""
        ps = [p for p in params if p.requires_grad]
        if not ps:
            raise ValueError("no trainable parameters")
        early_ids = [id(p) for p in early]
        first = [p for i in early_ids for p in ps if id(p) == i]
        self.params: List[torch.nn.Parameter] = first + [p for p in ps if id(p) not in set(early_ids)]
        self.n_early = len(first)
        dev, dt = self.params[0].device, self.params[0].dtype
        total = 0
        self.offsets = []
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("FlatGradBucket needs all parameters on one device/dtype")
            self.offsets.append(total)
            total += (p.numel() + 31) // 32 * 32  # keep every view 128-byte aligned
        self.split = self.offsets[self.n_early] if 0 < self.n_early < len(self.params) else 0
        self.buffer = torch.zeros(total, device=dev, dtype=dt)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.buffer[off:off + p.numel()].view_as(p)

    def zero_(self):
        self.buffer.zero_()

    def rebind(self):
        """Re-point .grad at the bucket (e.g. after an optimizer did set_to_none)."""
        for p, off in zip(self.params, self.offsets):
            p.grad = self.buffer[off:off + p.numel()].view_as(p)

    @property
    def nbytes(self) -> int:
        return self.buffer.numel() * self.buffer.element_size()

    @staticmethod
    def _active(group=None) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def all_reduce_segment(self, which: str, average: bool, group=None):
        """SUM (or AVG) all-reduce of the 'early' ([0, split)) or 'late' ([split, end)) segment on the current stream."""
        if not self._active(group):
            return
        seg = self.buffer[:self.split] if which == "early" else self.buffer[self.split:]
        if seg.numel() == 0:
            return
        dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=group)
        if average:
            seg.div_(dist.get_world_size(group))

    def all_reduce(self, average: bool, group=None, async_op: bool = False):
        if not self._active(group):
            return None
        work = dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average:
            if async_op:
                work.wait()
                work = None
            self.buffer.div_(dist.get_world_size(group))
        return work


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous row split of a global batch: rank r owns [lo, hi)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def philox_offset_for_shard(global_offset: int, lo: int, per_row: int) -> int:
    """Counter offset so a sharded run reproduces the single-GPU noise stream."""
    return global_offset + lo * per_row
