"""TrainStep: the user-facing "one training step" call — forward + loss + backward (+ the data-parallel
gradient all-reduce), optionally captured once into a CUDA graph and replayed.

At the reference's own sizes (batch 128-4096, latent 2-64) the hyperbolic path is a few KB-MB of data: the
step is launch-latency bound, so the win is launch count + zero host syncs + graph replay.
The reference's equivalents of this call are LightningModule.training_step -> loss -> backward
(models/vae_hyperbolic.py:250-254 etc.), which sync the host several times per step.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.distributed as dist

from .parallel import FlatGradBucket


class TrainStep:
    def __init__(self, model: torch.nn.Module, example_input: torch.Tensor, average_grads: bool = False,
                 use_graph: bool = True, loss_key: str = "loss_total", noise_shard="auto", optimizer=None, **loss_kwargs):
        """noise_shard: (first_global_row, global_rows) of this rank's shard for the in-kernel samplers, "auto" =
        rank * B_local of a world * B_local batch in a multi-rank job (every rank then draws its own rows' noise, the
        rows a single-GPU run of the global batch would draw), None = leave the process-wide setting alone."""
        self.model, self.loss_key, self.loss_kwargs = model, loss_key, loss_kwargs
        self.average = average_grads
        # optimizer: an hvae.optim.RiemannianAdam (one fused launch over the reduced bucket) stepped at the end of every
        # step, inside the captured graph; None = forward + backward only (the BASELINE metric)
        self.optimizer = optimizer
        self._opt_on = False   # warm-up and capture rehearsal steps do not update the parameters
        if noise_shard == "auto":
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                from . import ops
                from .parallel import philox_offset_for_shard

                b_local = int(example_input.shape[0])
                ops.set_noise_shard(philox_offset_for_shard(0, dist.get_rank() * b_local, 1), dist.get_world_size() * b_local)
        elif noise_shard is not None:
            from . import ops

            ops.set_noise_shard(int(noise_shard[0]), noise_shard[1])
        self.bucket = FlatGradBucket(model.parameters())
        self._hooks, self._order, self._fired = [], [], 0
        self._side = torch.cuda.Stream()
        self._early_launched = False
        # NCCL path: overlap the early gradient segment's collective with the rest of the backward.  With the bucket in
        # symmetric memory the exchange is ONE peer-memory kernel at the end of the step instead (measured at 2 GPUs:
        # 674 us/step against 693 us for overlapped NCCL and 712 us for a single NCCL all-reduce; 636 us without any
        # exchange) - its barriers spin, so it is not run concurrently with compute.
        self.overlap = (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                        and self.bucket._symm is None and os.environ.get("HVAE_DP_OVERLAP", "1") != "0")
        self.x = torch.empty_like(example_input)  # static input buffer (device)
        self.x.copy_(example_input)
        self._stage = torch.empty_like(example_input)  # landing buffer of the asynchronous host->device prefetch
        self._copy_stream = torch.cuda.Stream()
        self._staged = None   # event: the prefetch into _stage has landed
        self._consumed = None  # event: _stage has been copied into the static input
        self.loss = None
        self.graph = None
        if self.overlap:
            self._plan_overlap()
        for _ in range(3):
            self.loss = self._step()
        torch.cuda.synchronize()
        if self.optimizer is not None:
            self.optimizer.prepare()   # state + launch plan allocated now, outside any capture
        if use_graph:
            self._capture()
        self._opt_on = True

    # ---- data-parallel overlap: the early segment's all-reduce runs under the rest of the backward -------------
    def _plan_overlap(self, early_fraction: float = 0.35):
        """Observe one backward: the order in which parameter gradients complete.  The first parameters to finish
        (>= early_fraction of the payload) become the bucket's early segment; a hook on them launches that segment's
        all-reduce on a side stream as soon as the last of them has accumulated."""
        order = []
        hs = [p.register_post_accumulate_grad_hook(lambda p_, order=order: order.append(p_)) for p in self.bucket.params]
        self.bucket.zero_()
        self.model.loss(self.x, **self.loss_kwargs)[self.loss_key].backward()
        torch.cuda.synchronize()
        for h in hs:
            h.remove()
        total = sum(p.numel() for p in order)
        early, acc = [], 0
        for p in order[:-1]:  # keep at least one parameter late
            early.append(p)
            acc += p.numel()
            if acc >= early_fraction * total:
                break
        if not early or acc > 0.9 * total:
            self.overlap = False
            return
        self.bucket = FlatGradBucket(self.bucket.params, early=early)  # (a collective when the bucket is symmetric)
        self._n_early = len(early)

        def hook(_p):
            self._fired += 1
            if self._fired == self._n_early:
                cur = torch.cuda.current_stream()
                self.bucket.adopt(0, self._n_early)
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    self.bucket.all_reduce_segment("early", self.average)
                self._early_launched = True

        self._hooks = [p.register_post_accumulate_grad_hook(hook) for p in early]

    def _step(self):
        # .grad is detached from the bucket during the backward (autograd hands gradients over without an add kernel per
        # parameter) and adopted back with one multi-tensor copy; after the step .grad is the (reduced) bucket view.
        self.bucket.zero_()
        self.bucket.release()
        self._fired, self._early_launched = 0, False
        out = self.model.loss(self.x, **self.loss_kwargs)
        out[self.loss_key].backward()
        if self.overlap:
            if not self._early_launched:  # (a parameter without gradient this step): reduce it here instead
                self.bucket.adopt(0, self.bucket.n_early)
                self.bucket.all_reduce_segment("early", self.average)
            self.bucket.adopt(self.bucket.n_early, None)
            self.bucket.all_reduce_segment("late", self.average)
            torch.cuda.current_stream().wait_stream(self._side)
        else:
            self.bucket.adopt()
            self.bucket.all_reduce(average=self.average)
        if self.optimizer is not None and self._opt_on:
            self.optimizer.step()
        return out[self.loss_key].detach()

    def _capture(self):
        try:
            self._opt_on = False   # rehearsal steps below must not move the parameters
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            self._opt_on = True
            with torch.cuda.graph(g):
                loss = self._step()
            if self.optimizer is None:   # (with an optimizer a validating replay would be a hidden training step)
                g.replay()
            torch.cuda.synchronize()
            self.graph, self.loss = g, loss
        except Exception as ex:
            self.graph = None
            self._opt_on = True
            sys.stderr.write("hvae.TrainStep: CUDA graph capture failed (%s); running eagerly\n" % (str(ex).splitlines()[0],))
            torch.cuda.synchronize()

    def prefetch(self, x_host: torch.Tensor):
        """Start the host->device copy of the NEXT batch on a side stream; it overlaps with the step in flight.
        Pair with run_prefetched()."""
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)  # do not overwrite the landing buffer before the previous hand-over
        with torch.cuda.stream(cs):
            self._stage.copy_(x_host, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(cs)

    def run_prefetched(self) -> torch.Tensor:
        """One step on the batch handed over by prefetch()."""
        if self._staged is None:
            raise RuntimeError("TrainStep.run_prefetched() without a prefetch()")
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        self.x.copy_(self._stage, non_blocking=True)  # device->device hand-over (a few microseconds)
        self._consumed = torch.cuda.Event()
        self._consumed.record(cur)
        self._staged = None
        return self.run()

    def run(self, x: torch.Tensor = None) -> torch.Tensor:
        """One step. x: new batch (host pinned or device) copied into the static buffer, or None to reuse it.
        Returns the (device) loss tensor; gradients are in self.bucket.buffer / each parameter's .grad."""
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
            # a replay writes the bucket, not .grad: re-point .grad at the views in case an optimizer's
            # zero_grad(set_to_none=True) dropped them since the last step (it would silently get no updates)
            if any(p.grad is None for p in self.bucket.params):
                self.bucket.rebind()
        else:
            self.loss = self._step()
        return self.loss
