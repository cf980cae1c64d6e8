"""Drop-in for `hyperbolic_vae.distributions.wrapped_normal.WrappedNormal`
(reference: hyperbolic_vae/distributions/wrapped_normal.py:14-89).

rsample and log_prob are each ONE fused kernel (K4 / K5).  The reference validates `loc` with an
`allclose` on every construction (:48-52), which is a host sync; here the check is opt-in
(`WrappedNormal.validate_loc = True` or validate_args=True) so the training step stays async.
"""
from __future__ import annotations

import torch
from torch import Tensor
from torch.nn import functional as F

from .. import ops
from ..manifolds import PoincareBall


class WrappedNormal(torch.distributions.Distribution):
    arg_constraints = {
        "loc": torch.distributions.constraints.real,
        "scale": torch.distributions.constraints.positive,
    }
    support = torch.distributions.constraints.real
    has_rsample = True
    _mean_carrier_measure = 0
    validate_loc = False  # class-level switch for the reference's on-manifold assertion

    @property
    def mean(self):
        return self.loc

    @property
    def stddev(self):
        raise NotImplementedError

    @property
    def scale(self):
        return F.softplus(self._scale) if self.softplus else self._scale

    def __init__(self, loc: Tensor, scale: Tensor, manifold: PoincareBall, validate_args=None, softplus=False):
        self.dtype = loc.dtype
        self.softplus = softplus
        if loc.shape != scale.shape:
            loc, scale = torch.broadcast_tensors(loc, scale)
        self.loc, self._scale = loc, scale
        self.manifold = manifold
        if validate_args or WrappedNormal.validate_loc:
            try:
                self.manifold.assert_check_point_on_manifold(self.loc)
            except Exception as e:
                print(self.loc)
                raise e
        self.device = loc.device
        super().__init__(self.loc.shape[:-1], self.loc.shape[-1:], validate_args=False)

    def sample(self, shape=torch.Size()):
        with torch.no_grad():
            return self.rsample(shape)

    def rsample(self, sample_shape=torch.Size(), eps: Tensor = None) -> Tensor:
        """eps: optional injected standard-normal noise of shape sample_shape + loc.shape."""
        shape = self._extended_shape(sample_shape)
        if eps is None:
            eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        elif eps.shape != shape:
            raise ValueError("eps must have shape %s" % (tuple(shape),))
        D = shape[-1]
        mu = self.loc.reshape(-1, D)
        sig = self.scale.reshape(-1, D)
        B = mu.shape[0]
        z = ops.wrapped_sample_fwd(ops._c(mu), ops._c(sig), ops._c(eps).view(-1, B, D), self.manifold.c_value)
        return z.view(shape)

    def log_prob(self, x: Tensor) -> Tensor:
        """x: (S, *batch, D) (or missing the batch dims -> broadcast like the reference) -> (S, *batch, 1)"""
        D = int(self.event_shape[0])
        if self._is_origin_prior() and x.dim() >= 2:
            # origin prior with a scalar scale: every row of x is scored independently, any leading shape
            xs = ops._c(x).view(1, -1, D)
            lp = ops.wrapped_logprob_prior_fwd(xs, self._prior_sigma, self.manifold.c_value)
            return lp.view(*x.shape[:-1], 1)
        loc_shape = torch.Size([x.shape[0], *self.batch_shape, D])
        if x.dim() < len(loc_shape):
            x = x.unsqueeze(1)
        full = torch.broadcast_shapes(loc_shape, x.shape)  # (S, *batch, D)
        if x.shape != full:
            x = x.expand(full)
        B = 1
        for n in full[1:-1]:
            B *= int(n)
        xs = ops._c(x).view(full[0], B, D)
        if self._is_origin_prior():
            lp = ops.wrapped_logprob_prior_fwd(xs, self._prior_sigma, self.manifold.c_value)
        else:
            mu = ops._c(self.loc.expand(full[1:])).view(B, D)
            sig = ops._c(self.scale.expand(full[1:])).view(B, D)
            lp = ops.wrapped_logprob_fwd(mu, sig, xs, self.manifold.c_value)
        return lp.view(*full[:-1], 1)

    # The reference builds its prior as WrappedNormal(origin(D), const*ones(D)) (vae_hyperbolic.py:194-199):
    # callers can mark that case explicitly to take the cheaper prior kernel.
    _prior_sigma = None

    def _is_origin_prior(self):
        return self._prior_sigma is not None

    @classmethod
    def origin_prior(cls, dim: int, prior_scale: float, manifold: PoincareBall, device=None):
        loc = manifold.origin(dim, device=device)
        d = cls(loc, torch.ones_like(loc) * prior_scale, manifold)
        d._prior_sigma = float(prior_scale)
        return d
