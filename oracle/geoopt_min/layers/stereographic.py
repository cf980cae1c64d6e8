"""ORACLE: geoopt.layers.stereographic.Distance2StereographicHyperplanes — the gyroplane decoder
the reference's scripts 5/6/7 actually use (call sites /root/reference/hyperbolic_vae/models/
vae_hyperbolic.py:83, vae_hyperbolic_gyroplane_decoder.py:70, vae_hyperbolic_rnaseq.py:49).
Same computation as the reference's local copy (layers.py:193-228) without the bias."""
import torch

from ..tensor import ManifoldParameter
from ..utils import size2shape


class Distance2StereographicHyperplanes(torch.nn.Module):
    n = 0

    def __init__(self, plane_shape, num_planes, signed=True, squared=False, *, ball, std=1.0):
        super().__init__()
        self.signed = signed
        self.squared = squared
        self.ball = ball
        self.plane_shape = size2shape(plane_shape)
        self.num_planes = num_planes
        self.points = ManifoldParameter(torch.empty(num_planes, plane_shape), manifold=self.ball)
        self.std = std
        self.reset_parameters()

    def forward(self, input):
        x = input.unsqueeze(-self.n - 1)
        pts = self.points.permute(1, 0)
        pts = pts.view(pts.shape + (1,) * self.n)
        d = self.ball.dist2plane(x=x, p=pts, a=pts, signed=self.signed, dim=-self.n - 2)
        if self.squared and self.signed:
            d = d**2 * d.sign()
        elif self.squared:
            d = d**2
        return d

    @torch.no_grad()
    def reset_parameters(self):
        direction = torch.randn_like(self.points)
        direction /= direction.norm(dim=-1, keepdim=True)
        distance = torch.empty_like(self.points[..., 0]).normal_(std=self.std)
        self.points.set_(self.ball.expmap0(direction * distance.unsqueeze(-1)))
