"""ORACLE: pvae/distributions/hyperspherical_uniform.py (uniform on S^{dim}; dim = D-1)."""
import math

import torch


class HypersphericalUniform(torch.distributions.Distribution):
    support = torch.distributions.constraints.real
    has_rsample = False
    arg_constraints = {}

    def __init__(self, dim, device="cpu", validate_args=None):
        super().__init__(torch.Size([dim]), validate_args=validate_args)
        self._dim = dim
        self._device = device

    @property
    def dim(self):
        return self._dim

    def sample(self, shape=torch.Size()):
        shape = torch.Size(shape) if not isinstance(shape, torch.Size) else shape
        v = torch.randn(*shape, self._dim + 1, device=self._device)
        return v / v.norm(dim=-1, keepdim=True)

    def entropy(self):
        return self._log_normalizer()

    def log_prob(self, x):
        return -torch.ones(x.shape[:-1]).to(self._device) * self._log_normalizer()

    def _log_normalizer(self):
        # log surface area of S^{dim}:  log 2 + (dim+1)/2 log pi - lgamma((dim+1)/2)
        return torch.tensor(
            math.log(2) + (self._dim + 1) / 2 * math.log(math.pi) - math.lgamma((self._dim + 1) / 2)
        )
