"""ORACLE: pvae/distributions/riemannian_normal.py restated."""
import torch

from .hyperbolic_radius import HyperbolicRadius
from .hyperspherical_uniform import HypersphericalUniform


class RiemannianNormal(torch.distributions.Distribution):
    arg_constraints = {}
    support = torch.distributions.constraints.real
    has_rsample = True

    @property
    def mean(self):
        return self.loc

    def __init__(self, loc, scale, manifold, validate_args=None):
        assert not (torch.isnan(loc).any() or torch.isnan(scale).any())
        self.manifold = manifold
        self.loc = loc
        self.manifold.assert_check_point_on_manifold(self.loc)
        self.scale = scale.clamp(min=0.1, max=7.0)
        self.radius = HyperbolicRadius(manifold.coord_dim, manifold.c, self.scale)
        self.direction = HypersphericalUniform(manifold.coord_dim - 1, device=loc.device)
        super().__init__(self.loc.shape[:-1], torch.Size([manifold.coord_dim]), validate_args=validate_args)

    def sample(self, shape=torch.Size()):
        with torch.no_grad():
            return self.rsample(shape)

    def rsample(self, sample_shape=torch.Size()):
        alpha = self.direction.sample(torch.Size([*sample_shape, *self.loc.shape[:-1]]))
        radius = self.radius.rsample(sample_shape)
        return self.manifold.expmap_polar(self.loc, alpha, radius)

    def log_prob(self, value):
        loc = self.loc.expand(value.shape)
        radius_sq = self.manifold.dist(loc, value, keepdim=True).pow(2)
        return -radius_sq / 2 / self.scale.pow(2) - self.direction._log_normalizer() - self.radius.log_normalizer
