"""ORACLE: geoopt.optim.RiemannianAdam restated (App. A.1 end) — used only as the "next" row
§8(f)-1 checker; call sites /root/reference/hyperbolic_vae/models/vae_hyperbolic.py:236 etc."""
import torch

from .tensor import ManifoldParameter, ManifoldTensor


class RiemannianAdam(torch.optim.Adam):
    def __init__(self, *args, stabilize=None, **kwargs):
        super().__init__(*args, **kwargs)
        self._stabilize = stabilize

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            eps, lr, wd = group["eps"], group["lr"], group["weight_decay"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad
                man = p.manifold if isinstance(p, (ManifoldParameter, ManifoldTensor)) else None
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if wd:
                    grad = grad + wd * p
                if man is not None:
                    grad = man.egrad2rgrad(p, grad)
                    gg = man.inner(p, grad, keepdim=True)
                else:
                    gg = grad * grad
                m.mul_(b1).add_(grad, alpha=1 - b1)
                v.mul_(b2).add_(gg * (1 - b2))
                bc1 = 1 - b1 ** st["step"]
                bc2 = 1 - b2 ** st["step"]
                denom = (v / bc2).sqrt() + eps
                direction = (m / bc1) / denom
                if man is not None:
                    new_p, new_m = man.retr_transp(p, -lr * direction, m)
                    p.copy_(new_p)
                    m.copy_(new_m)
                else:
                    p.add_(direction, alpha=-lr)
        return loss
