"""Drop-in for `hyperbolic_vae.layers` (same class names, ctor signatures, parameter names/shapes so
reference state_dicts load): RiemannianLayer, GeodesicLayer (= GyroplaneLayer), MobiusLayer, ExpMap0,
Distance2PoincareHyperplanes, plus geoopt's Distance2StereographicHyperplanes.

reference: hyperbolic_vae/layers.py:35-76 (RiemannianLayer), :79-121 (GeodesicLayer), :124-130
(ExpMap0), :133-147 (MobiusLayer), :150-228 (Distance2PoincareHyperplanes).
"""
from __future__ import annotations

import math

import torch
from torch import Tensor, nn
from torch.nn import init

from . import ops
from .manifolds import ManifoldParameter, PoincareBall


class ExpMap0(nn.Module):
    """layers.py:124-130"""

    def __init__(self, manifold: PoincareBall):
        super().__init__()
        self.manifold = manifold

    def forward(self, input: Tensor) -> Tensor:
        return self.manifold.expmap0(input)


class RiemannianLayer(nn.Module):
    """layers.py:35-76.  `weight`/`bias` are produced together by one weight-prep kernel (K1b)."""

    def __init__(self, in_features: int, out_features: int, manifold: PoincareBall, over_param: bool, weight_norm: bool):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.manifold = manifold
        self._weight = nn.Parameter(torch.empty(out_features, in_features))
        self.over_param = over_param
        self.weight_norm = weight_norm
        if self.over_param:
            self._bias = ManifoldParameter(torch.empty(out_features, in_features), manifold=manifold)
        else:
            self._bias = nn.Parameter(torch.empty(out_features, 1))
        self.reset_parameters()

    def _prep(self):
        """-> (bias point (P,F), transported weight (P,F))"""
        c = self.manifold.c_value
        if self.over_param:
            return ops.weight_prep_overparam(self._weight, self._bias, c)
        return ops.weight_prep(self._weight, self._bias, c)

    @property
    def weight(self) -> Tensor:
        return self._prep()[1]

    @property
    def bias(self) -> Tensor:
        return self._prep()[0]

    def reset_parameters(self):
        init.kaiming_normal_(self._weight, a=math.sqrt(5))
        fan_in, _ = init._calculate_fan_in_and_fan_out(self._weight)
        bound = 4 / math.sqrt(fan_in)
        init.uniform_(self._bias, -bound, bound)
        if self.over_param:
            # reference does expmap0 here (layers.py:74-76); at init time parameters live on the host,
            # so this one-off uses the tensor expression, not the kernel.
            with torch.no_grad():
                c = self.manifold.c_value
                n = self._bias.norm(dim=-1, keepdim=True).clamp_min(1e-15)
                sc = c ** 0.5
                self._bias.copy_(self.manifold.projx(torch.tanh((sc * n).clamp(-15, 15)) * self._bias / (sc * n)))


class GeodesicLayer(RiemannianLayer):
    """layers.py:79-121 with pvae's batched semantics (App. A.2): input (..., D) -> (..., out_features);
    signed normdist2plane of every row to every gyroplane (a = bias point, p = transported weight)."""

    def __init__(self, in_features, out_features, manifold, over_param=False, weight_norm=False):
        super().__init__(in_features, out_features, manifold, over_param, weight_norm)

    def forward(self, input: Tensor, relu: bool = False) -> Tensor:
        """relu=True (not in the reference's signature): return relu(layer(input)) with the activation fused into the kernels."""
        bpt, w = self._prep()
        flags = ops.GYRO_PVAE | ops.GYRO_SIGNED | (ops.GYRO_SCALED if self.weight_norm else 0)
        return ops.gyroplane(input, w, bpt, None, self.manifold.c_value, flags, relu=relu)


GyroplaneLayer = GeodesicLayer  # the name BASELINE.json uses; a commented-out stub in the reference (layers.py:16-32)


class MobiusLayer(RiemannianLayer):
    """layers.py:133-147: mobius_matvec(weight, input) with projection."""

    def __init__(self, in_features, out_features, manifold, over_param=False, weight_norm=False):
        super().__init__(in_features, out_features, manifold, over_param, weight_norm)

    def forward(self, input: Tensor) -> Tensor:
        return ops.mobius_matvec(input, self._prep()[1], self.manifold.c_value)


class Distance2PoincareHyperplanes(nn.Module):
    """layers.py:150-228 (the reference's `bias=False` raises KeyError at :185-188; here it works)."""

    n = 0

    def __init__(self, plane_shape: int, num_planes: int, bias: bool = True, signed=True, squared=False, *,
                 ball: PoincareBall, std=1.0):
        super().__init__()
        self.signed = signed
        self.squared = squared
        self.ball = ball
        self.plane_shape = (int(plane_shape),)
        self.num_planes = num_planes
        self.points = ManifoldParameter(torch.empty(num_planes, plane_shape), manifold=self.ball)
        if bias:
            self.bias = nn.Parameter(torch.empty(num_planes))
        else:
            self.register_parameter("bias", None)
        self.std = std
        self.reset_parameters()

    def forward(self, input: Tensor) -> Tensor:
        flags = (ops.GYRO_SIGNED if self.signed else 0) | (ops.GYRO_SQUARED if self.squared else 0)
        return ops.gyroplane(input, self.points, self.points, self.bias, self.ball.c_value, flags)

    def extra_repr(self):
        return "plane_shape={}, num_planes={}".format(self.plane_shape, self.num_planes)

    @torch.no_grad()
    def reset_parameters(self):
        direction = torch.randn_like(self.points)
        direction /= direction.norm(dim=-1, keepdim=True)
        distance = torch.empty_like(self.points[..., 0]).normal_(std=self.std)
        u = direction * distance.unsqueeze(-1)
        c = self.ball.c_value
        sc = c ** 0.5
        n = u.norm(dim=-1, keepdim=True).clamp_min(1e-15)
        self.points.copy_(self.ball.projx(torch.tanh((sc * n).clamp(-15, 15)) * u / (sc * n)))
        if self.bias is not None:
            init.uniform_(self.bias, -1.0, 1.0)


class Distance2StereographicHyperplanes(Distance2PoincareHyperplanes):
    """geoopt.layers.stereographic.Distance2StereographicHyperplanes (no bias) — the gyroplane decoder the
    reference's scripts use (models/vae_hyperbolic.py:83, vae_hyperbolic_gyroplane_decoder.py:70, …rnaseq.py:49)."""

    def __init__(self, plane_shape: int, num_planes: int, signed=True, squared=False, *, ball: PoincareBall, std=1.0):
        super().__init__(plane_shape, num_planes, bias=False, signed=signed, squared=squared, ball=ball, std=std)


class Linear(nn.Linear):
    """torch.nn.Linear (same parameters, same state_dict keys) whose GEMM-sized CUDA fp32 calls run on the tensor-core
    fp32-accurate GEMM (ops.linear); everything else defers to torch.  reference: the nn.Linear trunk layers of
    hyperbolic_vae/models/*.py and pvae's Enc/Dec."""

    def forward(self, input, relu: bool = False):
        """relu=True (extension): relu(linear(input)) with the activation fused into the GEMM epilogue / gradient split."""
        return ops.linear(input, self.weight, self.bias, relu=relu)
