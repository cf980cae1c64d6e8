"""ORACLE: the PoincareBall object surface of geoopt (manifolds/stereographic/manifold.py @ a41c09b7)
that /root/reference reaches.  c is stored as isp_c = log(exp(c)-1) (fp32) and read
back as softplus(isp_c); k = -c."""
import torch

from ...tensor import ManifoldTensor
from ...utils import size2shape
from . import math


class Manifold(torch.nn.Module):
    name = "manifold"


class Stereographic(Manifold):
    name = "Stereographic"
    ndim = 1

    def __init__(self, k=0.0, learnable=False):
        super().__init__()
        k = torch.as_tensor(k)
        if not torch.is_floating_point(k):
            k = k.to(torch.get_default_dtype())
        self.k = torch.nn.Parameter(k, requires_grad=learnable)

    # -- checks ---------------------------------------------------------------------------
    def _check_point_on_manifold(self, x, *, atol=1e-5, rtol=1e-5, dim=-1):
        px = math.project(x, k=self.k, dim=dim)
        ok = torch.allclose(x, px, atol=atol, rtol=rtol)
        return ok, (None if ok else "'x' norm lies out of the bounds [-1/sqrt(c)+eps, 1/sqrt(c)-eps]")

    def check_point_on_manifold(self, x, *, explain=False, atol=1e-5, rtol=1e-5):
        ok, reason = self._check_point_on_manifold(x, atol=atol, rtol=rtol)
        return (ok, reason) if explain else ok

    def assert_check_point_on_manifold(self, x, *, atol=1e-5, rtol=1e-5):
        ok, reason = self._check_point_on_manifold(x, atol=atol, rtol=rtol)
        if not ok:
            raise ValueError("`x` seems to be a tensor not lying on {} manifold.\nerror: {}".format(self.name, reason))

    def check_vector_on_tangent(self, x, u, *, explain=False, atol=1e-5, rtol=1e-5):
        return (True, None) if explain else True

    def assert_check_vector_on_tangent(self, x, u, *, ok_point=False, atol=1e-5, rtol=1e-5):
        return None

    # -- math -----------------------------------------------------------------------------
    def projx(self, x, *, dim=-1):
        return math.project(x, k=self.k, dim=dim)

    def lambda_x(self, x, *, dim=-1, keepdim=False):
        return math.lambda_x(x, k=self.k, dim=dim, keepdim=keepdim)

    def dist(self, x, y, *, keepdim=False, dim=-1):
        return math.dist(x, y, k=self.k, keepdim=keepdim, dim=dim)

    def egrad2rgrad(self, x, u, *, dim=-1):
        return math.egrad2rgrad(x, u, k=self.k, dim=dim)

    def inner(self, x, u, v=None, *, keepdim=False, dim=-1):
        if v is None:
            v = u
        return math.inner(x, u, v, k=self.k, keepdim=keepdim, dim=dim)

    def retr(self, x, u, *, dim=-1):
        return math.project(x + u, k=self.k, dim=dim)

    def transp(self, x, y, v, *, dim=-1):
        return math.parallel_transport(x, y, v, k=self.k, dim=dim)

    def retr_transp(self, x, u, v, *, dim=-1):
        y = self.retr(x, u, dim=dim)
        return y, self.transp(x, y, v, dim=dim)

    def transp0(self, y, u, *, dim=-1):
        return math.parallel_transport0(y, u, k=self.k, dim=dim)

    def expmap(self, x, u, *, project=True, dim=-1):
        res = math.expmap(x, u, k=self.k, dim=dim)
        return math.project(res, k=self.k, dim=dim) if project else res

    def expmap0(self, u, *, project=True, dim=-1):
        res = math.expmap0(u, k=self.k, dim=dim)
        return math.project(res, k=self.k, dim=dim) if project else res

    def logmap(self, x, y, *, dim=-1):
        return math.logmap(x, y, k=self.k, dim=dim)

    def logmap0(self, x, *, dim=-1):
        return math.logmap0(x, k=self.k, dim=dim)

    def mobius_add(self, x, y, *, dim=-1, project=True):
        res = math.mobius_add(x, y, k=self.k, dim=dim)
        return math.project(res, k=self.k, dim=dim) if project else res

    def mobius_matvec(self, m, x, *, dim=-1, project=True):
        res = math.mobius_matvec(m, x, k=self.k, dim=dim)
        return math.project(res, k=self.k, dim=dim) if project else res

    def dist2plane(self, x, p, a, *, dim=-1, keepdim=False, signed=False, scaled=False):
        return math.dist2plane(x, p, a, dim=dim, k=self.k, keepdim=keepdim, signed=signed, scaled=scaled)

    def origin(self, *size, dtype=None, device=None, seed=42):
        return ManifoldTensor(torch.zeros(*size2shape(*size), dtype=dtype, device=device), manifold=self)

    def wrapped_normal(self, *size, mean, std=1, dtype=None, device=None):
        size = size2shape(*size)
        v = torch.randn(size, device=mean.device, dtype=mean.dtype) * std
        lambda_x = self.lambda_x(mean).unsqueeze(-1)
        return ManifoldTensor(self.expmap(mean, v / lambda_x), manifold=self)

    def extra_repr(self):
        return "c={}".format(float(-self.k))


class PoincareBall(Stereographic):
    name = "Poincare ball"

    @property
    def k(self):
        return -self.c

    @property
    def c(self):
        return torch.nn.functional.softplus(self.isp_c)

    def __init__(self, c=1.0, learnable=False):
        super().__init__(k=c, learnable=learnable)
        k = self._parameters.pop("k")
        with torch.no_grad():
            self.isp_c = k.exp_().sub_(1).log_()
