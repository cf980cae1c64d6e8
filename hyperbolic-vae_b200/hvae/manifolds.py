"""Drop-in for `hyperbolic_vae.manifolds` (+ the geoopt.PoincareBall surface the reference calls).

reference: hyperbolic_vae/manifolds.py:12-13 (PoincareBallWithExtras), :25-35 (logdetexp),
:41-65 (normdist2plane), :38 (MIN_NORM); geoopt call sites listed in SURVEY.md §8(c).

The heavy methods (expmap0, logmap0, expmap, logmap, mobius_add, mobius_matvec, dist, dist2plane)
launch the sm_100a kernels through `hvae.ops`; the light scalar helpers (lambda_x, transp0, …) are
plain tensor expressions on the caller's device.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import ops

MIN_NORM = 1e-15


class ManifoldTensor(torch.Tensor):
    """Tensor tagged with the manifold it lives on (geoopt.ManifoldTensor surface)."""

    def __new__(cls, data, manifold=None, requires_grad=False):
        inst = torch.Tensor._make_subclass(cls, data.detach() if isinstance(data, torch.Tensor) else torch.as_tensor(data), requires_grad)
        inst.manifold = manifold
        return inst


class ManifoldParameter(ManifoldTensor, nn.Parameter):
    """Parameter tagged with its manifold (geoopt.ManifoldParameter surface; layers.py:53,184)."""

    def __new__(cls, data=None, manifold=None, requires_grad=True):
        if data is None:
            data = torch.empty(0)
        return ManifoldTensor.__new__(cls, data, manifold=manifold, requires_grad=requires_grad)

    def __repr__(self):
        return "ManifoldParameter on {}:\n".format(self.manifold) + torch.Tensor.__repr__(self)

    def __deepcopy__(self, memo):
        out = type(self)(self.data.clone(memory_format=torch.preserve_format), self.manifold, self.requires_grad)
        memo[id(self)] = out
        return out


def _softplus_roundtrip(c: float) -> float:
    """float(geoopt.PoincareBall(c).c): c stored as log(exp(c)-1) in fp32 and read via softplus."""
    t = torch.as_tensor(float(c), dtype=torch.float32)
    return float(torch.nn.functional.softplus(t.exp().sub(1).log()))


class PoincareBall(nn.Module):
    """geoopt.PoincareBall(c) surface: `isp_c` parameter, `c` = softplus(isp_c), `k` = -c."""

    name = "Poincare ball"
    ndim = 1

    def __init__(self, c: float = 1.0, learnable: bool = False):
        super().__init__()
        if learnable:
            raise NotImplementedError("hvae kernels take the curvature by value; learnable c is not supported")
        k = torch.as_tensor(float(c), dtype=torch.float32)
        self.isp_c = nn.Parameter(k.exp().sub(1).log(), requires_grad=False)
        self._c_value = float(torch.nn.functional.softplus(self.isp_c.detach().cpu()))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._c_value = float(torch.nn.functional.softplus(self.isp_c.detach().cpu()))

    # -- curvature ------------------------------------------------------------------------------
    @property
    def c(self) -> Tensor:
        return torch.nn.functional.softplus(self.isp_c)

    @property
    def k(self) -> Tensor:
        return -self.c

    @property
    def c_value(self) -> float:
        """float(manifold.c) without a device sync (what every kernel receives)."""
        return self._c_value

    # -- kernel-backed maps -----------------------------------------------------------------------
    def expmap0(self, u: Tensor, *, project: bool = True, dim: int = -1) -> Tensor:
        _last_dim(u, dim)
        if not project:
            raise NotImplementedError("expmap0(project=False) is not on the reference's path")
        return ops.expmap0(u, self._c_value)

    def logmap0(self, y: Tensor, *, dim: int = -1) -> Tensor:
        _last_dim(y, dim)
        return ops.logmap0(y, self._c_value)

    def expmap(self, x: Tensor, u: Tensor, *, project: bool = True, dim: int = -1) -> Tensor:
        _last_dim(x, dim)
        if not project:
            raise NotImplementedError("expmap(project=False) is not on the reference's path")
        return ops.expmap(x, u, self._c_value)

    def logmap(self, x: Tensor, y: Tensor, *, dim: int = -1) -> Tensor:
        _last_dim(x, dim)
        return ops.logmap(x, y, self._c_value)

    def mobius_add(self, x: Tensor, y: Tensor, *, dim: int = -1, project: bool = True) -> Tensor:
        _last_dim(x, dim)
        return ops.mobius_add(x, y, self._c_value, project)

    def dist(self, x: Tensor, y: Tensor, *, keepdim: bool = False, dim: int = -1) -> Tensor:
        _last_dim(x, dim)
        return ops.dist(x, y, self._c_value, keepdim)

    def mobius_matvec(self, m: Tensor, x: Tensor, *, dim: int = -1, project: bool = True) -> Tensor:
        _last_dim(x, dim)
        if m.dim() != 2 or not project:
            raise NotImplementedError("mobius_matvec supports a 2-D matrix with project=True (the reference's use)")
        return ops.mobius_matvec(x, m, self._c_value)

    def dist2plane(self, x: Tensor, p: Tensor, a: Tensor, *, dim: int = -1, keepdim: bool = False, signed: bool = False,
                   scaled: bool = False) -> Tensor:
        """Supports the layer layout of the reference (layers.py:194-200): x (...,D,1), p/a (D,P), dim=-2."""
        if dim == -2 and x.shape[-1] == 1 and p.dim() == 2 and a.dim() == 2:
            flags = (ops.GYRO_SIGNED if signed else 0) | (ops.GYRO_SCALED if scaled else 0)
            pp = p.t()
            aa = pp if a is p else a.t()
            out = ops.gyroplane(x.squeeze(-1), pp, aa, None, self._c_value, flags)
            return out.unsqueeze(-2) if keepdim else out
        raise NotImplementedError("dist2plane: only the all-pairs layer layout (dim=-2) is kernel-backed")

    # -- light helpers (tensor expressions) -----------------------------------------------------------
    def projx(self, x: Tensor, *, dim: int = -1) -> Tensor:
        maxnorm = 0.996 / (self._c_value ** 0.5) if x.dtype == torch.float32 else (1 - 1e-5) / (self._c_value ** 0.5)
        norm = x.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
        return torch.where(norm > maxnorm, x / norm * maxnorm, x)

    def lambda_x(self, x: Tensor, *, dim: int = -1, keepdim: bool = False) -> Tensor:
        return 2 / (1 - self._c_value * x.pow(2).sum(dim=dim, keepdim=keepdim)).clamp_min(MIN_NORM)

    def transp0(self, y: Tensor, u: Tensor, *, dim: int = -1) -> Tensor:
        return u * (1 - self._c_value * y.pow(2).sum(dim=dim, keepdim=True)).clamp_min(MIN_NORM)

    def gyration(self, u: Tensor, v: Tensor, w: Tensor, *, dim: int = -1) -> Tensor:
        k = -self._c_value
        u2 = u.pow(2).sum(dim=dim, keepdim=True)
        v2 = v.pow(2).sum(dim=dim, keepdim=True)
        uv = (u * v).sum(dim=dim, keepdim=True)
        uw = (u * w).sum(dim=dim, keepdim=True)
        vw = (v * w).sum(dim=dim, keepdim=True)
        k2 = k * k
        a = -k2 * uw * v2 - k * vw + 2 * k2 * uv * vw
        b = -k2 * vw * u2 + k * uw
        d = 1 - 2 * k * uv + k2 * u2 * v2
        return w + 2 * (a * u + b * v) / d.clamp_min(MIN_NORM)

    def transp(self, x: Tensor, y: Tensor, v: Tensor, *, dim: int = -1) -> Tensor:
        return self.gyration(y, -x, v, dim=dim) * self.lambda_x(x, keepdim=True, dim=dim) / self.lambda_x(y, keepdim=True, dim=dim)

    def egrad2rgrad(self, x: Tensor, u: Tensor, *, dim: int = -1) -> Tensor:
        return u / self.lambda_x(x, keepdim=True, dim=dim) ** 2

    def inner(self, x: Tensor, u: Tensor, v: Tensor = None, *, keepdim: bool = False, dim: int = -1) -> Tensor:
        if v is None:
            v = u
        res = self.lambda_x(x, keepdim=True, dim=dim) ** 2 * (u * v).sum(dim=dim, keepdim=True)
        return res if keepdim else res.squeeze(dim)

    def retr(self, x: Tensor, u: Tensor, *, dim: int = -1) -> Tensor:
        return self.projx(x + u, dim=dim)

    def retr_transp(self, x, u, v, *, dim: int = -1):
        y = self.retr(x, u, dim=dim)
        return y, self.transp(x, y, v, dim=dim)

    def origin(self, *size, dtype=None, device=None, seed=42) -> Tensor:
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        if device is None:
            device = self.isp_c.device
        return ManifoldTensor(torch.zeros(*size, dtype=dtype, device=device), manifold=self)

    # -- checks: the reference runs these on every construction (a host sync); see WrappedNormal ----
    def _check_point_on_manifold(self, x: Tensor, *, atol=1e-5, rtol=1e-5, dim=-1):
        px = self.projx(x, dim=dim)
        ok = torch.allclose(x, px, atol=atol, rtol=rtol)
        return ok, (None if ok else "'x' norm lies out of the bounds [-1/sqrt(c)+eps, 1/sqrt(c)-eps]")

    def check_point_on_manifold(self, x: Tensor, *, explain=False, atol=1e-5, rtol=1e-5):
        ok, reason = self._check_point_on_manifold(x, atol=atol, rtol=rtol)
        return (ok, reason) if explain else ok

    def assert_check_point_on_manifold(self, x: Tensor, *, atol=1e-5, rtol=1e-5):
        ok, reason = self._check_point_on_manifold(x, atol=atol, rtol=rtol)
        if not ok:
            raise ValueError("`x` seems to be a tensor not lying on {} manifold.\nerror: {}".format(self.name, reason))

    def check_vector_on_tangent(self, x, u, *, explain=False, **kw):
        return (True, None) if explain else True

    def assert_check_vector_on_tangent(self, x, u, **kw):
        return None

    def extra_repr(self):
        return "c={}".format(self._c_value)


class PoincareBallWithExtras(PoincareBall):
    """hyperbolic_vae/manifolds.py:12-13 — an empty subclass in the reference; adds pvae's `zero`,
    `coord_dim`, `logdetexp`, `expmap_polar`, `normdist2plane` conveniences (App. A.2) when `dim` is given."""

    def __init__(self, c: float = 1.0, dim: int = None):
        super().__init__(c)
        self._dim = dim

    @property
    def coord_dim(self):
        return int(self._dim)

    @property
    def zero(self):
        return torch.zeros(1, int(self._dim), device=self.isp_c.device)

    def logdetexp(self, x, y, is_vector=False, keepdim=False):
        if is_vector:
            raise NotImplementedError
        return logdetexp(self, x, y, keepdim=keepdim)

    def expmap_polar(self, x, u, r, dim: int = -1):
        return ops.expmap_polar(x, u, r, self._c_value)

    def normdist2plane(self, x, a, p, keepdim=False, signed=False, dim=-1, norm=False):
        return normdist2plane(self, x, a, p, keepdim=keepdim, signed=signed, dim=dim, norm=norm)


def _last_dim(x: Tensor, dim: int):
    if dim != -1 and dim != x.dim() - 1:
        raise NotImplementedError("hvae kernels reduce over the last dimension only")


def logdetexp(manifold: PoincareBall, x: Tensor, y: Tensor, keepdim: bool = False) -> Tensor:
    """hyperbolic_vae/manifolds.py:25-35: (D-1) (log sinh(sqrt(c) d) - log sqrt(c) - log d), d = dist(x,y)."""
    d = manifold.dist(x, y, keepdim=keepdim)
    sc = manifold.c_value ** 0.5
    n = x.shape[-1]
    return (n - 1) * ops.log_sinhc(sc * d)


def normdist2plane(manifold_poincare: PoincareBall, x: Tensor, a: Tensor, p: Tensor, keepdim: bool = False,
                   signed: bool = False, dim: int = -1, norm: bool = False) -> Tensor:
    """hyperbolic_vae/manifolds.py:41-65.  Kernel-backed for the all-pairs call GeodesicLayer makes:
    x expanded over the plane axis (stride 0 at dim -2), a and p of shape (P, D)."""
    if dim != -1 or a.dim() != 2 or p.dim() != 2 or x.dim() < 2 or x.shape[-2] != a.shape[0] or x.stride(-2) != 0:
        raise NotImplementedError("normdist2plane: only the expanded all-pairs layout of GeodesicLayer is kernel-backed")
    x0 = x.select(-2, 0)  # undo the expand
    flags = ops.GYRO_PVAE | (ops.GYRO_SIGNED if signed else 0) | (ops.GYRO_SCALED if norm else 0)
    out = ops.gyroplane(x0, p, a, None, manifold_poincare.c_value, flags)
    return out.unsqueeze(-1) if keepdim else out
