"""ORACLE: pvae/ops/manifold_layers.py restated — identical in structure to the reference's
layers.py:35-147 except that GeodesicLayer unsqueezes the input before expanding (App. A.2)."""
import math

import torch
from torch import nn
from torch.nn import init

from ...geoopt_min.tensor import ManifoldParameter
from ..manifolds import normdist2plane


class RiemannianLayer(nn.Module):
    def __init__(self, in_features, out_features, manifold, over_param, weight_norm):
        super().__init__()
        self.in_features, self.out_features, self.manifold = in_features, out_features, manifold
        self._weight = nn.Parameter(torch.Tensor(out_features, in_features))
        self.over_param, self.weight_norm = over_param, weight_norm
        if over_param:
            self._bias = ManifoldParameter(torch.Tensor(out_features, in_features), manifold=manifold)
        else:
            self._bias = nn.Parameter(torch.Tensor(out_features, 1))
        self.reset_parameters()

    @property
    def weight(self):
        return self.manifold.transp0(self.bias, self._weight)

    @property
    def bias(self):
        return self._bias if self.over_param else self.manifold.expmap0(self._weight * self._bias)

    def reset_parameters(self):
        init.kaiming_normal_(self._weight, a=math.sqrt(5))
        fan_in, _ = init._calculate_fan_in_and_fan_out(self._weight)
        bound = 4 / math.sqrt(fan_in)
        init.uniform_(self._bias, -bound, bound)
        if self.over_param:
            with torch.no_grad():
                self._bias.set_(self.manifold.expmap0(self._bias))


class GeodesicLayer(RiemannianLayer):
    def __init__(self, in_features, out_features, manifold, over_param=False, weight_norm=False):
        super().__init__(in_features, out_features, manifold, over_param, weight_norm)

    def forward(self, input):
        # pvae slices the leading dims as input.shape[:-(input.dim()-2)], which is only right for its 3-D
        # (K, B, D) inputs; shape[:-1] is the same thing there and also covers (B, D).
        input = input.unsqueeze(-2).expand(*input.shape[:-1], self.out_features, self.in_features)
        return normdist2plane(self.manifold, input, self.bias, self.weight, signed=True, norm=self.weight_norm)


class MobiusLayer(RiemannianLayer):
    def __init__(self, in_features, out_features, manifold, over_param=False, weight_norm=False):
        super().__init__(in_features, out_features, manifold, over_param, weight_norm)

    def forward(self, input):
        return self.manifold.mobius_matvec(self.weight, input)
