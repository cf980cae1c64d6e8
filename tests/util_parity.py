"""Shared parity helpers for the GPU tests.

Criterion (BASELINE.json north_star: 1e-5 relative in fp32): an element passes if it is within
rtol*|ref64| + atol of the float64 oracle AFTER allowing for the float32 reference's own rounding
deviation from that float64 truth:   |cuda - o64| <= rtol*|o64| + atol + |o32 - o64|.
i.e. the kernel may never be further from the truth than the reference by more than 1e-5 relative.
Vector-valued outputs are judged on their row's max-norm (row_relative), gradients that are sums over
rows (parameter grads) on the tensor's max-norm (norm_relative).
"""
import torch

RTOL = 1e-5
ATOL = 1e-6


def assert_parity(cuda, o32, o64, rtol=RTOL, atol=ATOL, what="", norm_relative=False, row_relative=True, slack_mult=1.0):
    """rtol may be a tensor broadcastable to the output (per-row conditioning)."""
    cuda = cuda.detach().double().cpu()
    o32 = torch.Tensor(o32.detach()).double().cpu() if o32 is not None else None
    o64 = torch.Tensor(o64.detach()).double().cpu()
    assert cuda.shape == o64.shape, (what, cuda.shape, o64.shape)
    nan_c, nan_o = torch.isnan(cuda), torch.isnan(o64)
    assert torch.equal(nan_c, nan_o), "%s: NaN pattern differs (cuda %d, oracle %d)" % (what, nan_c.sum(), nan_o.sum())
    ok_mask = ~nan_o
    if norm_relative and ok_mask.any():
        scale = o64[ok_mask].abs().max()
    elif row_relative and o64.dim() >= 2 and o64.shape[-1] > 1:
        # vector-valued rows: an element that is a cancelling sum of O(row scale) terms cannot carry
        # elementwise-relative accuracy in fp32 (in the reference either); judge it on the row's scale
        scale = torch.nan_to_num(o64, nan=0.0).abs().amax(dim=-1, keepdim=True).expand_as(o64)
    else:
        scale = o64.abs()
    slack = slack_mult * (o32 - o64).abs() if o32 is not None else 0.0
    err = (cuda - o64).abs()
    if torch.is_tensor(rtol):
        rtol = rtol.detach().double().cpu()
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
