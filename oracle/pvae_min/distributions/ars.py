"""ORACLE: adaptive-rejection sampler with a FIXED tangent hull, restating pvae/distributions/ars.py
(SURVEY.md App. A.2): upper hull from tangents at `ns` abscissae, piecewise-exponential proposal,
accept iff U2 < exp(h(x) - offset - u_seg(x)); the hull is never refined."""
import torch

INF = float("inf")


def _diff(x):
    return x[:, 1:] - x[:, :-1]


class ARS:
    def __init__(self, logpdf, grad_logpdf, device, xi, lb=-INF, ub=INF, use_lower=False, ns=50, **fargs):
        self.device, self.lb, self.ub = device, lb, ub
        self.logpdf, self.grad_logpdf, self.fargs = logpdf, grad_logpdf, fargs
        self.ns = ns
        self.xi = xi.to(device)
        self.B, self.K = self.xi.size()
        self.h = torch.zeros(self.B, ns, device=device)
        self.hprime = torch.zeros(self.B, ns, device=device)
        self.x = torch.zeros(self.B, ns, device=device)
        self.h[:, : self.K] = self.logpdf(self.xi, **fargs)
        self.hprime[:, : self.K] = self.grad_logpdf(self.xi, **fargs)
        self.x[:, : self.K] = self.xi
        self.offset = self.h.max(-1)[0].view(-1, 1)
        self.h = self.h - self.offset
        if not (self.hprime[:, 0] > 0).all():
            raise IOError("initial anchor points must span mode of PDF (left)")
        if not (self.hprime[:, self.K - 1] < 0).all():
            raise IOError("initial anchor points must span mode of PDF (right)")
        self._build_hull()

    def _build_hull(self):
        K = self.K
        self.z = torch.zeros(self.B, K + 1, device=self.device)
        self.z[:, 0] = self.lb
        self.z[:, K] = self.ub
        self.z[:, 1:K] = (_diff(self.h[:, :K]) - _diff(self.x[:, :K] * self.hprime[:, :K])) / -_diff(self.hprime[:, :K])
        idx = [0] + list(range(K))
        self.u = self.h[:, idx] + self.hprime[:, idx] * (self.z - self.x[:, idx])
        self.s = _diff(torch.exp(self.u)) / self.hprime[:, :K]
        self.s[self.hprime[:, :K] == 0.0] = 0.0
        self.cs = torch.cat((torch.zeros(self.B, 1, device=self.device), torch.cumsum(self.s, dim=-1)), dim=-1)
        self.cu = self.cs[:, -1]

    def sample_upper(self, shape=torch.Size()):
        u = torch.rand(self.B, *shape, device=self.device)
        i = (self.cs / self.cu.unsqueeze(-1)).unsqueeze(-1) <= u.unsqueeze(1).expand(*self.cs.shape, *shape)
        idx = i.sum(1) - 1
        hp = self.hprime.gather(1, idx)
        xt = self.x.gather(1, idx) + (
            -self.h.gather(1, idx)
            + torch.log(hp * (self.cu.unsqueeze(-1) * u - self.cs.gather(1, idx)) + torch.exp(self.u.gather(1, idx)))
        ) / hp
        return xt, idx

    def sample(self, shape=torch.Size()):
        shape = shape if isinstance(shape, torch.Size) else torch.Size([shape])
        samples = torch.ones(self.B, *shape, device=self.device)
        pending = torch.ones(self.B, *shape, dtype=torch.bool, device=self.device)
        self.rounds = 0
        while pending.sum() != 0:
            self.rounds += 1
            xt, i = self.sample_upper(shape)
            ht = self.logpdf(xt, **self.fargs) - self.offset
            ut = self.h.gather(1, i) + (xt - self.x.gather(1, i)) * self.hprime.gather(1, i)
            u = torch.rand(shape, device=self.device)
            accept = u < torch.exp(ht - ut)
            take = pending & accept
            samples[take] = xt[take]
            pending = pending & ~accept
        return samples.t().unsqueeze(-1)
