"""Shared parity helpers for the GPU tests.

Criterion (BASELINE.json north_star: 1e-5 relative in fp32): an element passes if it is within
rtol*|ref64| + atol of the float64 oracle AFTER allowing for the float32 reference's own rounding
deviation from that float64 truth:   |cuda - o64| <= rtol*|o64| + atol + |o32 - o64|.
i.e. the kernel may never be further from the truth than the reference by more than 1e-5 relative.
For gradients that are sums over rows (parameter grads) the scale is the tensor's max-norm.
"""
import torch

RTOL = 1e-5
ATOL = 1e-6


def assert_parity(cuda, o32, o64, rtol=RTOL, atol=ATOL, what="", norm_relative=False):
    cuda = cuda.detach().double().cpu()
    o32 = torch.Tensor(o32.detach()).double().cpu() if o32 is not None else None
    o64 = torch.Tensor(o64.detach()).double().cpu()
    assert cuda.shape == o64.shape, (what, cuda.shape, o64.shape)
    nan_c, nan_o = torch.isnan(cuda), torch.isnan(o64)
    assert torch.equal(nan_c, nan_o), "%s: NaN pattern differs (cuda %d, oracle %d)" % (what, nan_c.sum(), nan_o.sum())
    ok_mask = ~nan_o
    scale = o64[ok_mask].abs().max() if (norm_relative and ok_mask.any()) else o64.abs()
    slack = (o32 - o64).abs() if o32 is not None else 0.0
    err = (cuda - o64).abs()
    bound = rtol * scale + atol + slack
    bad = (err > bound) & ok_mask
    if bad.any():
        i = torch.nonzero(bad)[0].tolist()
        idx = tuple(i)
        raise AssertionError(
            "%s: %d/%d elements out of tolerance; first at %s: cuda=%.9g o64=%.9g o32=%s err=%.3g bound=%.3g"
            % (what, int(bad.sum()), bad.numel(), idx, cuda[idx], o64[idx],
               ("%.9g" % o32[idx]) if o32 is not None else "-", err[idx], bound[idx] if torch.is_tensor(bound) and bound.dim() else float(bound))
        )
