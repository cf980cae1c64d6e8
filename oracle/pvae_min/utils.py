"""ORACLE: pvae/utils.py surface (Constants, logsinh, signed log-sum-exp, rexpand, misc)."""
import math

import torch


class Constants:
    eta = 1e-5
    log2 = math.log(2)
    logpi = math.log(math.pi)
    log2pi = math.log(2 * math.pi)
    logceilc = 88
    logfloorc = -104
    invsqrt2pi = 1.0 / math.sqrt(2 * math.pi)
    sqrthalfpi = math.sqrt(math.pi / 2)


def logsinh(x: torch.Tensor) -> torch.Tensor:
    # log sinh x = x + log(1 - e^{-2x}) - log 2
    return x + torch.log(1 - torch.exp(-2 * x)) - Constants.log2


def logcosh(x: torch.Tensor) -> torch.Tensor:
    return x + torch.log(1 + torch.exp(-2 * x)) - Constants.log2


def log_sum_exp_signs(value: torch.Tensor, signs: torch.Tensor, dim: int = 0, keepdim: bool = False):
    m, _ = torch.max(value, dim=dim, keepdim=True)
    value0 = value - m
    if keepdim is False:
        m = m.squeeze(dim)
    return m + torch.log(torch.sum(signs * torch.exp(value0), dim=dim, keepdim=keepdim))


def rexpand(A: torch.Tensor, *dimensions):
    """Expand tensor by appending trailing dims."""
    return A.view(A.shape + (1,) * len(dimensions)).expand(A.shape + tuple(dimensions))


def has_analytic_kl(type_p, type_q):
    return (type_p, type_q) in torch.distributions.kl._KL_REGISTRY


def probe_infnan(v, name, extras={}):
    nps = torch.isnan(v)
    s = nps.sum().item()
    if s > 0:
        raise RuntimeError("NaN in %s" % name)
