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

import os
from typing import Iterable, List

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Owns one contiguous buffer; every parameter's .grad is a view into it, so backward writes straight
    into the bucket and the all-reduce needs no gather/scatter copies."""

    def __init__(self, params: Iterable[torch.nn.Parameter], early: Iterable[torch.nn.Parameter] = (), symmetric: bool = None):
        """early: parameters to place first in the buffer (one contiguous segment [0, split)).
        symmetric: allocate the buffer in symmetric (peer-addressable) memory and all-reduce it with the repo's own
        NVLink kernel (hvae_allreduce_p2p_f32) instead of NCCL.  Default: on for CUDA buckets of an initialised
        multi-rank NCCL job unless HVAE_DP_P2P=0; construction is then a collective call (every rank, same order)."""
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
        if symmetric is None:
            symmetric = (dev.type == "cuda" and dt == torch.float32 and self._active() and dist.get_backend() == "nccl"
                         and os.environ.get("HVAE_DP_P2P", "1") != "0")
        self._symm = None
        self.p2p_blocks = 64  # grid of the peer-memory kernel; agreed across ranks below (block b meets block b)
        self.nvls = False     # in-switch reduction (multimem) instead of the peer-load loop
        self.nvls_blocks = 0  # 0: the kernel sizes its grid from the element count (identical on every rank)
        if symmetric:
            self.buffer, self._symm = self._alloc_symmetric(total, dev)
            if self._symm is not None:
                nb = torch.tensor([int(os.environ.get("HVAE_AR_BLOCKS", "64"))], device=dev)
                dist.all_reduce(nb, op=dist.ReduceOp.MIN)  # one value for the whole job, whatever each process's env says
                self.p2p_blocks = max(1, min(128, int(nb.item())))
                # NVLS (in-switch reduction) when the handle has a multicast mapping on EVERY rank; HVAE_DP_NVLS=0 turns it off
                mc = torch.tensor([1 if (getattr(self._symm, "multicast_ptr", 0) and os.environ.get("HVAE_DP_NVLS", "1") != "0") else 0],
                                  device=dev)
                dist.all_reduce(mc, op=dist.ReduceOp.MIN)
                self.nvls = bool(int(mc.item()))
        if self._symm is None:
            self.buffer = torch.zeros(total, device=dev, dtype=dt)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.buffer[off:off + p.numel()].view_as(p)

    @staticmethod
    def _alloc_symmetric(total: int, dev):
        """-> (buffer, handle) in torch symmetric memory, or (None, None) when peers cannot map each other."""
        try:
            import torch.distributed._symmetric_memory as symm

            from . import _cabi as C

            need = 4 * (64 + 2 * C.lib().hvae_allreduce_p2p_slots(dist.get_world_size()))
            if symm.get_signal_pad_size() < need:
                symm.set_signal_pad_size(need)
            buf = symm.empty(total, dtype=torch.float32, device=dev)
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            buf.zero_()
            ok = torch.ones(1, device=dev)
        except Exception as ex:  # no P2P mapping on this node: fall back to NCCL on every rank together
            import sys

            sys.stderr.write("hvae.parallel: symmetric memory unavailable (%s); using NCCL\n" % (str(ex).splitlines()[0],))
            buf, hdl, ok = None, None, torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks take the same path
        if float(ok) < 1.0:
            return None, None
        return buf, hdl

    def zero_(self):
        self.buffer.zero_()

    def release(self):
        """Detach .grad from the bucket before a backward: autograd then hands each gradient tensor over as is (no
        `grad += g` kernel per parameter); adopt() copies them into the bucket afterwards in one multi-tensor copy."""
        for p in self.params:
            p.grad = None

    def adopt(self, lo: int = 0, hi: int = None):
        """Copy the gradients autograd produced for params[lo:hi] into their bucket views (one fused multi-tensor copy)
        and point .grad back at the views.  Parameters without a gradient keep their (zeroed) view."""
        hi = len(self.params) if hi is None else hi
        src, dst = [], []
        for p, off in zip(self.params[lo:hi], self.offsets[lo:hi]):
            view = self.buffer[off:off + p.numel()].view_as(p)
            g = p.grad
            if g is not None and g.data_ptr() != view.data_ptr():
                if g.numel() >= (1 << 16):
                    # big tensors: an elementwise KERNEL that fills the GPU (the multi-tensor copy gives each tensor only
                    # numel/64K blocks: 10 us per 1.9 MB weight).  Not Tensor.copy_: into symmetric memory torch issues it as a
                    # DtoD memcpy, and three copy-engine nodes in the captured graph cost the step ~29 us of node hand-overs.
                    torch.mul(g, 1.0, out=view)
                else:
                    src.append(g if g.is_contiguous() else g.contiguous())
                    dst.append(view)
            p.grad = view
        if src:
            torch._foreach_copy_(dst, src)

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
        lo, hi = (0, self.split) if which == "early" else (self.split, self.buffer.numel())
        if hi <= lo:
            return
        if self._symm is not None:
            self._p2p(lo, hi - lo, 0 if which == "early" else 1, average)
            return
        seg = self.buffer[lo:hi]
        dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=group)
        if average:
            seg.div_(dist.get_world_size(group))

    def _p2p(self, offset: int, n: int, site: int, average: bool):
        """The repo's own all-reduce kernel over NVLink peer memory (csrc/allreduce_p2p.cu) on the current stream.
        site: which of the two disjoint signal-pad slot ranges to use (calls that may overlap in time need their own)."""
        from . import _cabi as C

        h = self._symm
        W = h.world_size
        base = 64 + site * C.lib().hvae_allreduce_p2p_slots(W)
        if self.nvls:
            # the NVSwitch reduces: multimem.ld_reduce / multimem.st on the bucket's multicast address
            C.call("hvae_allreduce_nvls_f32", h.multicast_ptr, h.signal_pad_ptrs_dev, h.rank, W, offset, n, base,
                   1.0 / W if average else 1.0, self.nvls_blocks, C.stream())
            return
        C.call("hvae_allreduce_p2p_f32", h.buffer_ptrs_dev, h.signal_pad_ptrs_dev, h.rank, W, offset, n, base,
               1.0 / W if average else 1.0, self.p2p_blocks, C.stream())

    def check(self):
        """Host-side health check of the peer-memory exchange (syncs the device): raises if one of this rank's
        all-reduce barriers timed out, i.e. a peer never arrived (it skipped a step, raised, or died)."""
        if self._symm is None:
            return
        pad = self._symm.get_signal_pad(self._symm.rank, (1,), dtype=torch.int32)
        if int(pad[0].item()) != 0:
            raise RuntimeError("hvae.parallel: a peer-memory all-reduce barrier timed out on rank %d "
                               "(a peer rank did not reach the exchange)" % self._symm.rank)

    def all_reduce(self, average: bool, group=None, async_op: bool = False):
        if not self._active(group):
            return None
        if self._symm is not None:
            self._p2p(0, self.buffer.numel(), 1, average)
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
