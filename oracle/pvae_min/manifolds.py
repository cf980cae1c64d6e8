"""ORACLE: pvae/manifolds/poincareball.py — PoincareBall(dim, c) with zero / logdetexp /
expmap_polar / normdist2plane on top of the geoopt ball."""
import torch

from ..geoopt_min.manifolds.stereographic import PoincareBall as _GeooptBall
from ..geoopt_min.manifolds.stereographic import math as gmath

MIN_NORM = 1e-15


class PoincareBall(_GeooptBall):
    def __init__(self, dim, c=1.0):
        super().__init__(c)
        self.register_buffer("dim", torch.as_tensor(dim, dtype=torch.int))

    @property
    def coord_dim(self):
        return int(self.dim)

    @property
    def device(self):
        return self.c.device

    @property
    def zero(self):
        return torch.zeros(1, int(self.dim)).to(self.device)

    def logdetexp(self, x, y, is_vector=False, keepdim=False):
        d = y.norm(dim=-1, keepdim=keepdim) * self.lambda_x(x, keepdim=keepdim) if is_vector else self.dist(x, y, keepdim=keepdim)
        sc = self.c.sqrt()
        return (int(self.dim) - 1) * (torch.sinh(sc * d) / sc / d).log()

    def expmap_polar(self, x, u, r, dim: int = -1):
        sqrt_c = self.c**0.5
        u_norm = torch.norm(u, dim=-1, p=2, keepdim=True).clamp_min(MIN_NORM)
        second_term = gmath.tanh(sqrt_c / 2 * r) * u / (sqrt_c * u_norm)
        return self.mobius_add(x, second_term, dim=dim)

    def normdist2plane(self, x, a, p, keepdim: bool = False, signed: bool = False, dim: int = -1, norm: bool = False):
        return normdist2plane(self, x, a, p, keepdim=keepdim, signed=signed, dim=dim, norm=norm)


def normdist2plane(ball, x, a, p, keepdim: bool = False, signed: bool = False, dim: int = -1, norm: bool = False):
    """pvae PoincareBall.normdist2plane == /root/reference/hyperbolic_vae/manifolds.py:41-65 (free function there)."""
    c = ball.c
    sqrt_c = c**0.5
    diff = ball.mobius_add(-p, x, dim=dim)
    diff_norm2 = diff.pow(2).sum(dim=dim, keepdim=keepdim).clamp_min(MIN_NORM)
    sc_diff_a = (diff * a).sum(dim=dim, keepdim=keepdim)
    if not signed:
        sc_diff_a = sc_diff_a.abs()
    a_norm = a.norm(dim=dim, keepdim=keepdim, p=2).clamp_min(MIN_NORM)
    num = 2 * sqrt_c * sc_diff_a
    denom = (1 - c * diff_norm2) * a_norm
    res = gmath.arsinh(num / denom.clamp_min(MIN_NORM)) / sqrt_c
    if norm:
        res = res * a_norm
    return res
