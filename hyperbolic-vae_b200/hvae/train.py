"""TrainStep: the user-facing "one training step" call — forward + loss + backward (+ the data-parallel
gradient all-reduce), optionally captured once into a CUDA graph and replayed.

At the reference's own sizes (batch 128-4096, latent 2-64) the hyperbolic path is a few KB-MB of data: the
step is launch-latency bound, so the win is launch count + zero host syncs + graph replay.
The reference's equivalents of this call are LightningModule.training_step -> loss -> backward
(models/vae_hyperbolic.py:250-254 etc.), which sync the host several times per step.
"""
from __future__ import annotations

import sys

import torch

from .parallel import FlatGradBucket


class TrainStep:
    def __init__(self, model: torch.nn.Module, example_input: torch.Tensor, average_grads: bool = False,
                 use_graph: bool = True, loss_key: str = "loss_total", **loss_kwargs):
        self.model, self.loss_key, self.loss_kwargs = model, loss_key, loss_kwargs
        self.average = average_grads
        self.bucket = FlatGradBucket(model.parameters())
        self.x = torch.empty_like(example_input)  # static input buffer (device)
        self.x.copy_(example_input)
        self.loss = None
        self.graph = None
        for _ in range(3):
            self.loss = self._step()
        torch.cuda.synchronize()
        if use_graph:
            self._capture()

    def _step(self):
        self.bucket.zero_()
        out = self.model.loss(self.x, **self.loss_kwargs)
        out[self.loss_key].backward()
        self.bucket.all_reduce(average=self.average)
        return out[self.loss_key].detach()

    def _capture(self):
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = self._step()
            g.replay()
            torch.cuda.synchronize()
            self.graph, self.loss = g, loss
        except Exception as ex:
            self.graph = None
            sys.stderr.write("hvae.TrainStep: CUDA graph capture failed (%s); running eagerly\n" % (str(ex).splitlines()[0],))
            torch.cuda.synchronize()

    def run(self, x: torch.Tensor = None) -> torch.Tensor:
        """One step. x: new batch (host pinned or device) copied into the static buffer, or None to reuse it.
        Returns the (device) loss tensor; gradients are in self.bucket.buffer / each parameter's .grad."""
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.loss = self._step()
        return self.loss
