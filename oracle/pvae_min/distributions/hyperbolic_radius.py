"""ORACLE: pvae/distributions/hyperbolic_radius.py restated.

Density on r>0:  rho(r) = exp(-r^2/(2 s^2)) (sinh(sqrt(c) r)/sqrt(c))^(dim-1) / Z(s, c, dim).
With n = dim-1, b_k = (n-2k) sqrt(c) and sinh^n(t) = 2^-n sum_k (-1)^k C(n,k) e^{(n-2k) t}:
  Z      = c^{-n/2} 2^{-n} s sqrt(pi/2) * sum_k (-1)^k C(n,k) e^{b_k^2 s^2/2} (1 + erf(b_k s/sqrt2))
  F(r)   = sum_k (-1)^k C(n,k) e^{b_k^2 s^2/2} (erf((r - b_k s^2)/(s sqrt2)) + erf(b_k s/sqrt2)) / (same sum with 1+erf)
All series are evaluated in float64 with a signed log-sum-exp, as pvae does, and returned in fp32.
The first two moments (pvae's `mean`/`variance` closed forms) only seed the ARS abscissae; they are
restated from the same expansion (see _moments).
"""
import math
from numbers import Number

import torch
from torch.autograd import Function

from ..utils import log_sum_exp_signs, logsinh
from .ars import ARS

SQRT2 = math.sqrt(2.0)


def _series(scale: torch.Tensor, c: torch.Tensor, dim: int):
    """Return (k, signs, logbinom, b) broadcast over scale[..., None] in float64."""
    n = dim - 1
    k = torch.arange(dim, dtype=torch.float64, device=scale.device)
    signs = torch.where((torch.arange(dim, device=scale.device) % 2) == 0, 1.0, -1.0).double()
    dimf = torch.tensor(float(dim), dtype=torch.float64, device=scale.device)
    logbinom = torch.lgamma(dimf) - torch.lgamma(k + 1) - torch.lgamma(dimf - k)
    b = (n - 2 * k) * c.double().sqrt()
    return k, signs, logbinom, b


def _log_normalizer_value(scale: torch.Tensor, c: torch.Tensor, dim: int) -> torch.Tensor:
    s = scale.double().unsqueeze(-1)
    cd = c.double()
    _, signs, logbinom, b = _series(scale, c, dim)
    v = logbinom + (b * s).pow(2) / 2 + torch.log1p(torch.erf(b * s / SQRT2))
    lse = log_sum_exp_signs(v, signs, dim=-1)
    n = dim - 1
    return 0.5 * (math.log(math.pi) - math.log(2)) + scale.double().log() - n * (0.5 * cd.log() + math.log(2)) + lse


class _LogNormalizer(Function):
    """pvae wraps logZ in a custom Function with a hand-written sigma-gradient, because autograd through
    log1p(erf(x)) is NaN once erf saturates at -1.  d logZ/d s = 1/s + sum_k (-1)^k [b_k^2 s E_k + C_k b_k sqrt(2/pi)]
    / sum_k (-1)^k E_k with E_k = C_k e^{b_k^2 s^2/2}(1 + erf(b_k s/sqrt2))  (no gradient to c, as in pvae)."""

    @staticmethod
    def forward(ctx, scale, c, dim):
        ctx.save_for_backward(scale.detach())
        ctx.c, ctx.dim = c.detach(), dim
        return _log_normalizer_value(scale.detach(), c.detach(), dim).to(torch.float64)

    @staticmethod
    def backward(ctx, grad):
        (scale,) = ctx.saved_tensors
        s = scale.double().unsqueeze(-1)
        _, signs, logbinom, b = _series(scale, ctx.c, ctx.dim)
        v = logbinom + (b * s).pow(2) / 2 + torch.log1p(torch.erf(b * s / SQRT2))
        m = v.max(dim=-1, keepdim=True)[0]
        E = torch.exp(v - m)
        num = (signs * (b * b * s * E + torch.exp(logbinom - m) * b * math.sqrt(2 / math.pi))).sum(-1)
        den = (signs * E).sum(-1)
        g = 1.0 / scale.double() + num / den
        return (grad * g).to(scale.dtype), None, None


def log_normalizer(scale: torch.Tensor, c: torch.Tensor, dim: int) -> torch.Tensor:
    """float64 logZ(scale) with pvae's analytic sigma-gradient."""
    return _LogNormalizer.apply(scale, c, dim)


def cdf_r(value: torch.Tensor, scale: torch.Tensor, c: torch.Tensor, dim: int) -> torch.Tensor:
    """float64 CDF of the radius; value/scale broadcastable, returns value's broadcast shape."""
    r = value.double().unsqueeze(-1)
    s = scale.double().unsqueeze(-1)
    _, signs, logbinom, b = _series(scale, c, dim)
    base = logbinom + (b * s).pow(2) / 2
    m = base.max(dim=-1, keepdim=True)[0]
    w = signs * torch.exp(base - m)
    num = (w * (torch.erf((r - b * s * s) / (s * SQRT2)) + torch.erf(b * s / SQRT2))).sum(-1)
    den = (w * (1 + torch.erf(b * s / SQRT2))).sum(-1)
    return num / den


def _moments(scale: torch.Tensor, c: torch.Tensor, dim: int):
    """E[r], Var[r] in float64 from the same binomial expansion:
       int r e^{-r^2/2s^2 + b r}   = s^2 + b s^3 sqrt(pi/2) e^{b^2 s^2/2}(1+erf)
       int r^2 e^{-r^2/2s^2 + b r} = b s^4 + s^3 sqrt(pi/2)(1 + b^2 s^2) e^{b^2 s^2/2}(1+erf)"""
    s = scale.double().unsqueeze(-1)
    _, signs, logbinom, b = _series(scale, c, dim)
    base = logbinom + (b * s).pow(2) / 2
    m = base.max(dim=-1, keepdim=True)[0]
    w = signs * torch.exp(base - m)
    w0 = signs * torch.exp(logbinom - m)  # weights of the non-exponential terms
    g = math.sqrt(math.pi / 2) * (1 + torch.erf(b * s / SQRT2))
    z0 = (w * s * g).sum(-1)
    z1 = (w0 * s * s + w * b * s**3 * g).sum(-1)
    z2 = (w0 * b * s**4 + w * s**3 * (1 + (b * s).pow(2)) * g).sum(-1)
    mean = z1 / z0
    var = z2 / z0 - mean * mean
    return mean, var


def grad_cdf_value_scale(value, scale, c, dim):
    """(dF/dr, dF/dscale) in float64; dF/dr = rho(r); dF/dscale by differentiating cdf_r."""
    value = value.detach().double().requires_grad_(True)
    scale = scale.detach().double().requires_grad_(True)
    with torch.enable_grad():
        F = cdf_r(value, scale, c, dim)
        gv, gs = torch.autograd.grad(F.sum(), (value, scale))
    return gv, gs


class impl_rsample(Function):
    """Implicit reparameterisation: dr/dscale = -(dF/dscale)/(dF/dr)."""

    @staticmethod
    def forward(ctx, value, scale, c, dim):
        ctx.save_for_backward(value.detach(), scale.detach())
        ctx.c, ctx.dim = c.detach(), dim
        return value

    @staticmethod
    def backward(ctx, grad_output):
        value, scale = ctx.saved_tensors
        sc = scale.expand(value.shape)
        gv, gs = grad_cdf_value_scale(value, sc, ctx.c, ctx.dim)
        dr_ds = (-gs / gv).to(grad_output.dtype)
        g = grad_output * dr_ds
        # reduce over broadcast (sample) dims back to scale's shape
        while g.dim() > scale.dim():
            g = g.sum(0)
        for i, (a, b) in enumerate(zip(g.shape, scale.shape)):
            if a != b:
                g = g.sum(i, keepdim=True)
        return None, g, None, None


class HyperbolicRadius(torch.distributions.Distribution):
    support = torch.distributions.constraints.positive
    has_rsample = True
    arg_constraints = {}

    def __init__(self, dim, c, scale, ars=True, validate_args=None):
        self.dim = int(dim)
        self.c = c if torch.is_tensor(c) else torch.tensor(float(c))
        self.scale = scale
        self.device = scale.device
        self.ars = ars
        batch_shape = torch.Size() if isinstance(scale, Number) else self.scale.size()
        self.log_normalizer = self._log_normalizer()
        if torch.isnan(self.log_normalizer).any() or torch.isinf(self.log_normalizer).any():
            raise ValueError("nan or inf in log_normalizer")
        super().__init__(batch_shape, validate_args=validate_args)

    def _log_normalizer(self):
        return log_normalizer(self.scale, self.c, self.dim).float()

    @property
    def mean(self):
        return _moments(self.scale, self.c, self.dim)[0].float()

    @property
    def variance(self):
        return _moments(self.scale, self.c, self.dim)[1].float()

    @property
    def stddev(self):
        return self.variance.sqrt()

    def log_prob(self, value):
        sc = self.c.sqrt()
        res = (
            -value.pow(2) / (2 * self.scale.pow(2))
            + (self.dim - 1) * logsinh(sc * value)
            - (self.dim - 1) / 2 * self.c.log()
            - self.log_normalizer
        )
        return res

    def grad_log_prob(self, value):
        sc = self.c.sqrt()
        return -value / self.scale.pow(2) + (self.dim - 1) * sc * torch.cosh(sc * value) / torch.sinh(sc * value)

    def cdf(self, value):
        return cdf_r(value, self.scale, self.c, self.dim).float()

    def sample(self, sample_shape=torch.Size()):
        if sample_shape == torch.Size():
            sample_shape = torch.Size([1])
        with torch.no_grad():
            mean = self.mean
            stddev = self.stddev
            bad = torch.isnan(stddev)
            if bad.any():
                stddev[bad] = self.scale[bad]
            bad = torch.isnan(mean)
            if bad.any():
                mean[bad] = ((self.dim - 1) * self.scale.pow(2) * self.c.sqrt())[bad]
            steps = torch.linspace(0.1, 3, 10).to(self.device)
            steps = torch.cat((-steps.flip(0), steps))
            xi = torch.cat([mean + s * torch.min(stddev, 0.95 * mean / 3) for s in steps], dim=1)
            ars = ARS(self.log_prob, self.grad_log_prob, self.device, xi=xi, ns=20, lb=0)
            value = ars.sample(sample_shape)
        return value

    def rsample(self, sample_shape=torch.Size()):
        value = self.sample(sample_shape)
        return impl_rsample.apply(value, self.scale, self.c, self.dim)
