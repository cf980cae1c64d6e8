"""Drop-in for `geoopt.optim.RiemannianAdam` as the reference uses it (hyperbolic_vae/models/vae_hyperbolic.py:235-243,
...gyroplane_decoder.py:173, ...rnaseq.py:139, vae_one_b.py:270): same constructor, same state (`step`, `exp_avg`,
`exp_avg_sq` per parameter), same arithmetic - Euclidean Adam for plain parameters, and for ManifoldParameters on the
Poincare ball egrad2rgrad, the Riemannian second moment, retraction project(x + u) and parallel transport of `exp_avg`.

The whole model is ONE kernel launch per step (csrc/riemannian_adam.cu).  Gradients are read where they are - with
hvae.train.TrainStep that is the flat, already all-reduced bucket.  Hyper-parameters and the step count are device
scalars, so the step can be captured in a CUDA graph and a scheduler's new learning rate takes effect on replay
(call `sync_hyper()` after changing `param_groups[i]["lr"]`; torch schedulers do not know about it).
"""
from __future__ import annotations

import ctypes

import torch

from . import _cabi as C
from .manifolds import ManifoldParameter


class RiemannianAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, stabilize=None):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by the reference and not implemented")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._stabilize = stabilize   # geoopt re-projects every `stabilize` steps; the retraction here always projects
        self._plans = {}

    # ---- one launch plan per param group: descriptor table on the device, hyper-parameters on the device ----
    def _plan(self, gi, group):
        ps = [p for p in group["params"] if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        plan = self._plans.get(gi)
        if plan is not None and plan["key"] == key:
            return plan
        if not ps:
            return None
        dev = ps[0].device
        C.require_cuda(*[p.data for p in ps])
        L = C.lib()
        dsz = L.hvae_riemannian_adam_desc_bytes()
        host = ctypes.create_string_buffer(dsz * len(ps))
        blocks = 0
        for i, p in enumerate(ps):
            st = self.state[p]
            if len(st) == 0:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            if not (p.is_contiguous() and p.grad.is_contiguous()):
                raise RuntimeError("hvae.optim.RiemannianAdam needs contiguous parameters and gradients")
            man = getattr(p, "manifold", None) if isinstance(p, ManifoldParameter) else None
            c = float(man.c_value) if man is not None else 0.0
            cols = p.shape[-1] if man is not None else 1
            nb = L.hvae_riemannian_adam_describe(ctypes.cast(host, ctypes.c_void_p), i, p.data_ptr(), p.grad.data_ptr(),
                                                 st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), cols, c, blocks)
            if nb < 0:
                raise RuntimeError("hvae_riemannian_adam_describe rejected parameter %d" % i)
            blocks += nb
        table = torch.frombuffer(bytearray(host.raw), dtype=torch.uint8).to(dev)
        hyper = torch.zeros(L.hvae_riemannian_adam_hyper_bytes() // 4, dtype=torch.float32, device=dev)
        plan = dict(key=key, params=ps, table=table, hyper=hyper, blocks=blocks, n=len(ps), step=None)
        self._plans[gi] = plan
        self._write_hyper(plan, group)
        return plan

    def _write_hyper(self, plan, group):
        b1, b2 = group["betas"]
        step = max((self.state[p]["step"] for p in plan["params"]), default=0)
        vals = torch.tensor([group["lr"], b1, b2, group["eps"], group["weight_decay"], float(step)], dtype=torch.float32)
        plan["hyper"].copy_(vals.to(plan["hyper"].device), non_blocking=False)
        plan["lr"] = group["lr"]

    def prepare(self):
        """Allocate the moments and build the launch plans now (needs .grad on the parameters) - required before the step
        is captured in a CUDA graph, where allocations and host->device copies are not allowed."""
        for gi, group in enumerate(self.param_groups):
            self._plan(gi, group)

    def sync_hyper(self):
        """Push lr / betas / eps / weight_decay of every param group to the device (after a scheduler changed them)."""
        for gi, group in enumerate(self.param_groups):
            plan = self._plans.get(gi)
            if plan is not None:
                self._write_hyper(plan, group)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        for gi, group in enumerate(self.param_groups):
            plan = self._plan(gi, group)
            if plan is None:
                continue
            if not capturing and plan.get("lr") != group["lr"]:
                self._write_hyper(plan, group)
            plan["hyper"][5:6].add_(1.0)   # the step count of this update, advanced in-stream (graph-replay safe)
            C.call("hvae_riemannian_adam_step_f32", plan["table"].data_ptr(), plan["n"], plan["blocks"], plan["hyper"].data_ptr(),
                   C.stream())
            if not capturing:
                for p in plan["params"]:
                    self.state[p]["step"] += 1
        return loss
