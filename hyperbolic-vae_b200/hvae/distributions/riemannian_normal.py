"""Drop-in for `hyperbolic_vae.distributions.old_pvae_riemannian_normal.RiemannianNormal` and the pvae
classes under it (HyperbolicRadius, HypersphericalUniform) — reference: hyperbolic_vae/distributions/
old_pvae_riemannian_normal.py:12-52; pvae semantics per SURVEY.md App. A.2.

  rsample:  alpha ~ U(S^{D-1}), r ~ rho(.; sigma) (K6: in-kernel rejection sampling, Philox),
            z = expmap_polar(mu, alpha, r); dr/dsigma by implicit reparameterisation (K6').
  log_prob: -dist(mu, z)^2 / (2 sigma^2) - log|S^{D-1}| - logZ(sigma)   (K7: float64 signed series).

The sampler cannot be bit-identical to pvae's (its ARS consumes a data-dependent number of torch.rand
calls); parity is exact formulas on injected (alpha, r) + a KS test of the sampled radii against cdf_r.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from .. import ops
from ..manifolds import PoincareBall


class HypersphericalUniform(torch.distributions.Distribution):
    """Uniform on S^{dim} (dim = D-1)."""

    support = torch.distributions.constraints.real
    has_rsample = False
    arg_constraints = {}

    def __init__(self, dim, device="cuda", validate_args=None):
        super().__init__(torch.Size([dim]), validate_args=False)
        self._dim, self._device = dim, device

    @property
    def dim(self):
        return self._dim

    def sample(self, shape=torch.Size()):
        """(*shape, dim + 1) unit vectors.  On CUDA the normals are drawn in the kernel (Philox, ops.sphere_sample): one
        launch instead of randn + norm + divide, graph-replay safe, and reproducible per row under data parallelism."""
        shape = torch.Size(shape)
        dev = torch.device(self._device)
        if dev.type == "cuda" and len(shape) >= 1:
            S = int(shape[:-1].numel()) if len(shape) > 1 else 1
            return ops.sphere_sample(S, int(shape[-1]), self._dim + 1, dev).view(*shape, self._dim + 1)
        v = torch.randn(*shape, self._dim + 1, device=self._device)
        return v / v.norm(dim=-1, keepdim=True)

    def _log_normalizer(self) -> float:
        return math.log(2) + (self._dim + 1) / 2 * math.log(math.pi) - math.lgamma((self._dim + 1) / 2)

    def entropy(self):
        return torch.tensor(self._log_normalizer())

    def log_prob(self, x):
        return torch.full(x.shape[:-1], -self._log_normalizer(), device=x.device)


class HyperbolicRadius(torch.distributions.Distribution):
    """pvae HyperbolicRadius(dim, c, scale): scale (..., 1) per-row sigma."""

    support = torch.distributions.constraints.positive
    has_rsample = True
    arg_constraints = {}

    def __init__(self, dim, c, scale: Tensor, ars=True, validate_args=None):
        self.dim = int(dim)
        self.c_value = float(c) if not torch.is_tensor(c) else float(c)
        self.scale = scale
        self.device = scale.device
        # (logZ, dlogZ/dsigma) from one kernel; the derivative is also what the fused head's backward consumes
        lz, dlz = ops.hradius_lognorm_fwd(ops._c(scale).view(-1), self.dim, self.c_value)
        self.log_normalizer, self._dlogz = lz.view(scale.shape), dlz
        super().__init__(self.scale.size(), validate_args=False)

    philox_counter: Tensor = None  # optional override of the per-device noise counter (ops.philox_counter)

    def sample(self, sample_shape=torch.Size(), seed=None, offset=None) -> Tensor:
        """Counters default to the per-device DEVICE-side Philox counter, advanced in-stream: graph replays draw fresh
        noise (a host offset would be frozen into a captured graph)."""
        S = int(torch.Size(sample_shape).numel()) if len(sample_shape) else 1
        r = ops.hradius_sample(self.scale, S, self.dim, self.c_value, seed=seed, offset=offset,
                               offset_dev=HyperbolicRadius.philox_counter)
        return r.view(S, *self.scale.shape)

    def rsample(self, sample_shape=torch.Size(), r: Tensor = None) -> Tensor:
        """r: optional injected radii of shape (S, *scale.shape)."""
        if r is None:
            r = self.sample(sample_shape)
        S = r.shape[0]
        out, _ = ops.hradius_reparam(ops._c(r.detach()).view(S, -1), ops._c(self.scale).view(-1), self.dim, self.c_value)
        return out.view(r.shape)

    def cdf(self, value: Tensor) -> Tensor:
        S = value.shape[0]
        return ops.hradius_cdf(value.reshape(S, -1), self.scale, self.dim, self.c_value).view(value.shape)

    def log_prob(self, value: Tensor) -> Tensor:
        sc = math.sqrt(self.c_value)
        x = sc * value
        logsinh = x + torch.log1p(-torch.exp(-2 * x)) - math.log(2)
        return (-value.pow(2) / (2 * self.scale.pow(2)) + (self.dim - 1) * logsinh
                - (self.dim - 1) / 2 * math.log(self.c_value) - self.log_normalizer)


class RiemannianNormal(torch.distributions.Distribution):
    arg_constraints = {}
    support = torch.distributions.constraints.real
    has_rsample = True
    validate_loc = False  # the reference asserts no-NaN and on-manifold on every construction (host syncs)

    @property
    def mean(self):
        return self.loc

    def __init__(self, loc: Tensor, scale: Tensor, manifold: PoincareBall, validate_args=None, *, scale_is_clamped: bool = False):
        """scale_is_clamped (keyword-only extension): the caller already applied the reference's clamp(0.1, 7), e.g. inside
        the fused scale head (ops.sigma_head) - it is not applied a second time."""
        if validate_args or RiemannianNormal.validate_loc:
            assert not (torch.isnan(loc).any() or torch.isnan(scale).any())
            manifold.assert_check_point_on_manifold(loc)
        self.manifold = manifold
        self.loc = loc
        self.scale = scale if scale_is_clamped else scale.clamp(min=0.1, max=7.0)
        D = loc.shape[-1]
        self.radius = HyperbolicRadius(D, manifold.c_value, self.scale)
        self.direction = HypersphericalUniform(D - 1, device=loc.device)
        super().__init__(loc.shape[:-1], loc.shape[-1:], validate_args=False)

    def sample(self, shape=torch.Size()):
        with torch.no_grad():
            return self.rsample(shape)

    def rsample(self, sample_shape=torch.Size(), alpha: Tensor = None, r: Tensor = None) -> Tensor:
        sample_shape = torch.Size(sample_shape)
        if alpha is None:
            alpha = self.direction.sample(torch.Size([*sample_shape, *self.loc.shape[:-1]]))
        radius = self.radius.rsample(sample_shape, r=r)
        if len(sample_shape) == 0:
            return ops.expmap_polar(self.loc, alpha.reshape(self.loc.shape), radius.reshape(*self.loc.shape[:-1], 1),
                                    self.manifold.c_value)
        return ops.expmap_polar(self.loc, alpha, radius, self.manifold.c_value)

    def rsample_kl(self, prior: "RiemannianNormal", alpha: Tensor = None, r: Tensor = None):
        """One sample z (1, B, D) and its Monte-Carlo KL term log q(z) - log p(z) (1, B) against an origin-centred prior
        with scalar sigma - rsample() followed by kl_mc(), as ONE kernel per direction (ops.rn_head): the backward
        delivers the total gradients of loc and scale (KL terms, the sample's path through expmap_polar, the implicit
        reparameterisation dr/dsigma and the normaliser's dlogZ/dsigma) without any autograd glue kernels."""
        B, D = self.loc.shape[-2], self.loc.shape[-1]
        c = self.manifold.c_value
        mu = ops._c(self.loc).view(B, D)
        sig = ops._c(self.scale).view(B)
        if alpha is None:
            alpha = self.direction.sample(torch.Size([1, B]))
        if r is None:
            r = self.radius.sample(torch.Size([1]))
        alpha = ops._c(alpha.detach()).view(B, D)
        r = ops._c(r.detach()).view(1, B)
        dr = ops.hradius_rgrad(r, sig.detach(), D, c)
        z, kl = ops.rn_head_fwd(mu, sig, self.radius.log_normalizer.detach().view(B), self.radius._dlogz.detach().view(B), alpha,
                                r.view(B), dr.view(B), ops._c(prior.scale).detach().view(1),
                                ops._c(prior.radius.log_normalizer).detach().view(1), c)
        return z.view(1, B, D), kl.view(1, B)

    def kl_mc(self, z: Tensor, prior: "RiemannianNormal") -> Tensor:
        """Monte-Carlo KL term log q(z) - log p(z) for z (S, B, D), q = self with per-row sigma (B,1) and an
        origin-centred prior with scalar sigma — ONE fused kernel (forward) instead of two log_prob graphs.
        Returns (S, B).  (The log|S^{D-1}| normalisers cancel.)"""
        S, B, D = z.shape
        return ops.rn_kl_fwd(ops._c(self.loc).view(B, D), ops._c(self.scale).view(B), ops._c(self.radius.log_normalizer).view(B),
                             ops._c(z), ops._c(prior.scale).view(1), ops._c(prior.radius.log_normalizer).view(1),
                             self.manifold.c_value)

    def log_prob(self, value: Tensor) -> Tensor:
        loc = self.loc.expand(value.shape)
        radius_sq = self.manifold.dist(loc, value, keepdim=True).pow(2)
        return -radius_sq / 2 / self.scale.pow(2) - self.direction._log_normalizer() - self.radius.log_normalizer
