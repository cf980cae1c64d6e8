"""Shared parity helpers for the GPU tests.

Criterion (BASELINE.json north_star: 1e-5 relative in fp32): an element passes if it is within
rtol*|ref64| + atol of the float64 oracle AFTER allowing for the float32 reference's own rounding
deviation from that float64 truth:   |cuda - o64| <= rtol*|o64| + atol + |o32 - o64|.
i.e. the kernel may never be further from the truth than the reference by more than 1e-5 relative.
Vector-valued outputs are judged on their row's max-norm (row_relative), gradients that are sums over
rows (parameter grads) on the tensor's max-norm (norm_relative).
"""
import os

import torch

RTOL = 1e-5
ATOL = 1e-6

# ---- strict audit -----------------------------------------------------------------------------------------------
# Besides the conditioned criterion above, every comparison is ALSO judged by the north-star's literal number with no
# float64 slack and no conditioning factor:   |cuda - o32| <= 1e-5 * scale(o32) (+ 1e-7 * scale absolute floor for exact
# zeros), scale = the row's max-norm (vector rows), the tensor's max-norm (batch-summed gradients) or |o32| (scalars).
# The counts are recorded per call (AUDIT; tests/conftest.py writes them to gpurun_out/parity_audit_*.json and prints a
# summary), and on WELL-CONDITIONED elements - those the caller itself grants no more than 2e-5 (kappa < 2) - a strict
# failure fails the test (HVAE_PARITY_STRICT=0 turns the assertion into a report).
STRICT_RTOL = 1e-5
AUDIT = []
STRICT_ASSERT = os.environ.get("HVAE_PARITY_STRICT", "1") != "0"


def _strict_audit(what, cuda, o32, scale_kind, rtol, ok_mask):
    if o32 is None:
        return None
    if scale_kind == "norm":
        scale = o32[ok_mask].abs().max() if ok_mask.any() else torch.tensor(0.0, dtype=torch.float64)
        scale = scale.expand_as(o32)
    elif scale_kind == "row":
        scale = torch.nan_to_num(o32, nan=0.0).abs().amax(dim=-1, keepdim=True).expand_as(o32)
    else:
        scale = o32.abs()
    err = (cuda - o32).abs()
    tiny = 1e-7 * (scale if scale_kind != "elem" else torch.nan_to_num(o32, nan=0.0).abs().max().expand_as(o32))
    fail = (err > STRICT_RTOL * scale + tiny) & ok_mask
    if torch.is_tensor(rtol):
        well = (rtol.expand_as(o32) if rtol.dim() else rtol) <= 2 * RTOL
        well = well & ok_mask if torch.is_tensor(well) and well.dim() else ok_mask & bool(well)
    else:
        well = ok_mask if rtol <= 2 * RTOL else torch.zeros_like(ok_mask)
    ratio = torch.where(ok_mask, err / (STRICT_RTOL * scale + tiny).clamp_min(1e-30), torch.zeros_like(err)).clamp_max(1e9)
    rec = dict(what=what, n=int(ok_mask.sum()), strict_fail=int(fail.sum()), n_well=int(well.sum()),
               strict_fail_well=int((fail & well).sum()),
               worst_well=float(ratio[well].max()) if bool(well.any()) else 0.0,
               worst_all=float(ratio.max()) if ratio.numel() else 0.0)
    AUDIT.append(rec)
    return rec


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
    kind = "norm" if norm_relative else ("row" if (row_relative and o64.dim() >= 2 and o64.shape[-1] > 1) else "elem")
    rec = _strict_audit(what, cuda, o32, kind, rtol, ok_mask)
    bad = (err > bound) & ok_mask
    if bad.any():
        i = torch.nonzero(bad)[0].tolist()
        idx = tuple(i)
        raise AssertionError(
            "%s: %d/%d elements out of tolerance; first at %s: cuda=%.9g o64=%.9g o32=%s err=%.3g bound=%.3g"
            % (what, int(bad.sum()), bad.numel(), idx, cuda[idx], o64[idx],
               ("%.9g" % o32[idx]) if o32 is not None else "-", err[idx], bound[idx] if torch.is_tensor(bound) and bound.dim() else float(bound))
        )
    if STRICT_ASSERT and rec is not None and rec["strict_fail_well"]:
        raise AssertionError("%s: %d of %d well-conditioned elements fail the plain criterion |cuda - ref32| <= 1e-5 * scale "
                             "(worst ratio %.3g)" % (what, rec["strict_fail_well"], rec["n_well"], rec["worst_well"]))


def kappa(c, *points):
    """Per-row condition factor max_i 1/(1 - c|p_i|^2) >= 1 (capped at the fp32 projection radius, 125).
    Every Poincare map divides by (1 - c|.|^2): values carry a relative fp32 error ~ eps*kappa and
    derivatives of two-point maps ~ eps*kappa^2 (both points near the boundary), in ANY evaluation order
    (the reference's included)."""
    k = None
    for p in points:
        p = torch.Tensor(p.detach()).double().cpu()
        p = p.reshape(-1, p.shape[-1])
        ki = 1.0 / (1.0 - c * p.pow(2).sum(-1, keepdim=True)).clamp_min(4e-3)
        k = ki if k is None else torch.maximum(k, ki)
    return k.clamp_min(1.0)


def rtol_val(kap, base=RTOL):
    return base * kap


def rtol_grad(kap, base=RTOL):
    return torch.maximum(torch.full_like(kap, base), 2e-6 * kap * kap)


def pair_kappa(c, x, p):
    """(B,P) condition factor of the gyroplane pair: 1/(1 - c|(-p)(+)x|^2), float64."""
    x = torch.Tensor(x.detach()).double().cpu()
    p = torch.Tensor(p.detach()).double().cpu()
    x2, p2, px = (x * x).sum(-1)[:, None], (p * p).sum(-1)[None, :], x @ p.t()   # O(B P) memory (a condition estimate)
    A = 1 - 2 * c * px + c * x2
    Bc = 1 - c * p2
    den = (1 - 2 * c * px + c * c * p2 * x2).clamp_min(1e-15)
    dn2 = (A * A * p2 - 2 * A * Bc * px + Bc * Bc * x2) / (den * den)
    return 1.0 / (1.0 - c * dn2).abs().clamp_min(1e-7)
