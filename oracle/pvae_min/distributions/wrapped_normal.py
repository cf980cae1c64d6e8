"""ORACLE: pvae/distributions/wrapped_normal.py restated; the base class that
/root/reference/hyperbolic_vae/distributions/old_pvae_wrapped_normal.py:16 subclasses."""
import torch
from torch.distributions import Normal
from torch.distributions.utils import _standard_normal, broadcast_all
from torch.nn import functional as F


class WrappedNormal(torch.distributions.Distribution):
    arg_constraints = {"loc": torch.distributions.constraints.real, "scale": torch.distributions.constraints.positive}
    support = torch.distributions.constraints.real
    has_rsample = True

    @property
    def mean(self):
        return self.loc

    @property
    def scale(self):
        return F.softplus(self._scale) if self.softplus else self._scale

    def __init__(self, loc, scale, manifold, validate_args=None, softplus=False):
        self.dtype = loc.dtype
        self.softplus = softplus
        self.loc, self._scale = broadcast_all(loc, scale)
        self.manifold = manifold
        self.manifold.assert_check_point_on_manifold(self.loc)
        self.device = loc.device
        super().__init__(self.loc.shape[:-1], torch.Size([manifold.coord_dim]), validate_args=validate_args)

    def sample(self, shape=torch.Size()):
        with torch.no_grad():
            return self.rsample(shape)

    def rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(sample_shape)
        v = self.scale * _standard_normal(shape, dtype=self.loc.dtype, device=self.loc.device)
        zero = self.manifold.zero
        v = v / self.manifold.lambda_x(zero, keepdim=True)
        u = self.manifold.transp(zero, self.loc, v)
        return self.manifold.expmap(self.loc, u)

    def log_prob(self, x):
        shape = x.shape
        loc = self.loc.unsqueeze(0).expand(x.shape[0], *self.batch_shape, self.manifold.coord_dim)
        if len(shape) < len(loc.shape):
            x = x.unsqueeze(1)
        zero = self.manifold.zero
        v = self.manifold.logmap(loc, x)
        v = self.manifold.transp(loc, zero, v)
        u = v * self.manifold.lambda_x(zero, keepdim=True)
        norm_pdf = Normal(torch.zeros_like(self.scale), self.scale).log_prob(u).sum(-1, keepdim=True)
        return norm_pdf - self.manifold.logdetexp(loc, x, keepdim=True)
