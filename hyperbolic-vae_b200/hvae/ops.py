"""torch custom ops (`torch.ops.hvae.*`) over the C ABI, each with an ANALYTIC backward op — autograd
never differentiates through primitives on this path (BASELINE.json north_star (4)).

All ops take contiguous float32 CUDA tensors flattened to rows; the public helpers below do the
reshaping.  `c` is float(manifold.c) — the fp32 softplus round-trip value, not the ctor argument.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _cabi as C

_op = torch.library.custom_op


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _rows(t: Tensor) -> Tensor:
    return _c(t).view(-1, t.shape[-1])


# ---------------------------------------------------------------------------------------------------
# unary row maps: expmap0 / logmap0
# ---------------------------------------------------------------------------------------------------
def _unary(name):
    fwd_sym, bwd_sym = "hvae_%s_fwd_f32" % name, "hvae_%s_bwd_f32" % name

    @_op("hvae::%s_fwd" % name, mutates_args=())
    def fwd(x: Tensor, c: float) -> Tensor:
        C.require_cuda(x)
        y = torch.empty_like(x)
        C.call(fwd_sym, C.ptr(x), C.ptr(y), x.shape[0], x.shape[1], c, C.stream())
        return y

    @fwd.register_fake
    def _(x, c):
        return torch.empty_like(x)

    @_op("hvae::%s_bwd" % name, mutates_args=())
    def bwd(x: Tensor, g: Tensor, c: float) -> Tensor:
        C.require_cuda(x, g)
        gx = torch.empty_like(x)
        C.call(bwd_sym, C.ptr(x), C.ptr(g), C.ptr(gx), x.shape[0], x.shape[1], c, C.stream())
        return gx

    @bwd.register_fake
    def _(x, g, c):
        return torch.empty_like(x)

    def setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0])
        ctx.c = inputs[1]

    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return bwd(x, _c(g), ctx.c), None

    fwd.register_autograd(backward, setup_context=setup)
    return fwd, bwd


expmap0_fwd, expmap0_bwd = _unary("expmap0")
logmap0_fwd, logmap0_bwd = _unary("logmap0")


def expmap0(u: Tensor, c: float) -> Tensor:
    return expmap0_fwd(_rows(u), c).view(u.shape)


def logmap0(y: Tensor, c: float) -> Tensor:
    return logmap0_fwd(_rows(y), c).view(y.shape)


# ---------------------------------------------------------------------------------------------------
# binary row maps: mobius_add / expmap / logmap / dist
# ---------------------------------------------------------------------------------------------------
@_op("hvae::mobius_add_fwd", mutates_args=())
def mobius_add_fwd(x: Tensor, y: Tensor, c: float, project: bool) -> Tensor:
    C.require_cuda(x, y)
    out = torch.empty_like(x)
    C.call("hvae_mobius_add_fwd_f32", C.ptr(x), C.ptr(y), C.ptr(out), x.shape[0], x.shape[1], c, int(project), C.stream())
    return out


@mobius_add_fwd.register_fake
def _(x, y, c, project):
    return torch.empty_like(x)


@_op("hvae::mobius_add_bwd", mutates_args=())
def mobius_add_bwd(x: Tensor, y: Tensor, g: Tensor, c: float, project: bool) -> Tuple[Tensor, Tensor]:
    C.require_cuda(x, y, g)
    gx, gy = torch.empty_like(x), torch.empty_like(y)
    C.call("hvae_mobius_add_bwd_f32", C.ptr(x), C.ptr(y), C.ptr(g), C.ptr(gx), C.ptr(gy), x.shape[0], x.shape[1], c,
           int(project), C.stream())
    return gx, gy


@mobius_add_bwd.register_fake
def _(x, y, g, c, project):
    return torch.empty_like(x), torch.empty_like(y)


def _madd_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.c, ctx.project = inputs[2], inputs[3]


def _madd_backward(ctx, g):
    x, y = ctx.saved_tensors
    gx, gy = mobius_add_bwd(x, y, _c(g), ctx.c, ctx.project)
    return gx, gy, None, None


mobius_add_fwd.register_autograd(_madd_backward, setup_context=_madd_setup)


def _broadcast_rows(a: Tensor, b: Tensor):
    if a.shape != b.shape:
        a, b = torch.broadcast_tensors(a, b)
    return _rows(a), _rows(b), a.shape


def mobius_add(x: Tensor, y: Tensor, c: float, project: bool = True) -> Tensor:
    xr, yr, shape = _broadcast_rows(x, y)
    return mobius_add_fwd(xr, yr, c, project).view(shape)


def _binary(name, out_is_scalar=False):
    fwd_sym, bwd_sym = "hvae_%s_fwd_f32" % name, "hvae_%s_bwd_f32" % name

    @_op("hvae::%s_fwd" % name, mutates_args=())
    def fwd(x: Tensor, y: Tensor, c: float) -> Tensor:
        C.require_cuda(x, y)
        out = x.new_empty(x.shape[0]) if out_is_scalar else torch.empty_like(x)
        C.call(fwd_sym, C.ptr(x), C.ptr(y), C.ptr(out), x.shape[0], x.shape[1], c, C.stream())
        return out

    @fwd.register_fake
    def _(x, y, c):
        return x.new_empty(x.shape[0]) if out_is_scalar else torch.empty_like(x)

    @_op("hvae::%s_bwd" % name, mutates_args=())
    def bwd(x: Tensor, y: Tensor, g: Tensor, c: float) -> Tuple[Tensor, Tensor]:
        C.require_cuda(x, y, g)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        C.call(bwd_sym, C.ptr(x), C.ptr(y), C.ptr(g), C.ptr(gx), C.ptr(gy), x.shape[0], x.shape[1], c, C.stream())
        return gx, gy

    @bwd.register_fake
    def _(x, y, g, c):
        return torch.empty_like(x), torch.empty_like(y)

    def setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0], inputs[1])
        ctx.c = inputs[2]

    def backward(ctx, g):
        x, y = ctx.saved_tensors
        gx, gy = bwd(x, y, _c(g), ctx.c)
        return gx, gy, None

    fwd.register_autograd(backward, setup_context=setup)
    return fwd, bwd


expmap_fwd, expmap_bwd = _binary("expmap")
logmap_fwd, logmap_bwd = _binary("logmap")
dist_fwd, dist_bwd = _binary("dist", out_is_scalar=True)


def expmap(x: Tensor, u: Tensor, c: float) -> Tensor:
    xr, ur, shape = _broadcast_rows(x, u)
    return expmap_fwd(xr, ur, c).view(shape)


def logmap(x: Tensor, y: Tensor, c: float) -> Tensor:
    xr, yr, shape = _broadcast_rows(x, y)
    return logmap_fwd(xr, yr, c).view(shape)


def dist(x: Tensor, y: Tensor, c: float, keepdim: bool = False) -> Tensor:
    xr, yr, shape = _broadcast_rows(x, y)
    d = dist_fwd(xr, yr, c).view(shape[:-1])
    return d.unsqueeze(-1) if keepdim else d


# ---------------------------------------------------------------------------------------------------
# WrappedNormal: sample / log_prob / fused latent head
# ---------------------------------------------------------------------------------------------------
@_op("hvae::wrapped_sample_fwd", mutates_args=())
def wrapped_sample_fwd(mu: Tensor, sigma: Tensor, eps: Tensor, c: float) -> Tensor:
    """mu, sigma: (B,D); eps: (S,B,D) -> z (S,B,D)"""
    C.require_cuda(mu, sigma, eps)
    S, B, D = eps.shape
    z = torch.empty_like(eps)
    C.call("hvae_wrapped_sample_fwd_f32", C.ptr(mu), C.ptr(sigma), C.ptr(eps), C.ptr(z), S, B, D, c, C.stream())
    return z


@wrapped_sample_fwd.register_fake
def _(mu, sigma, eps, c):
    return torch.empty_like(eps)


@_op("hvae::wrapped_sample_bwd", mutates_args=())
def wrapped_sample_bwd(mu: Tensor, sigma: Tensor, eps: Tensor, gz: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(mu, sigma, eps, gz)
    S, B, D = eps.shape
    gmu, gsig = torch.empty_like(mu), torch.empty_like(sigma)
    C.call("hvae_wrapped_sample_bwd_f32", C.ptr(mu), C.ptr(sigma), C.ptr(eps), C.ptr(gz), C.ptr(gmu), C.ptr(gsig), S, B, D,
           c, C.stream())
    return gmu, gsig


@wrapped_sample_bwd.register_fake
def _(mu, sigma, eps, gz, c):
    return torch.empty_like(mu), torch.empty_like(sigma)


def _ws_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], inputs[2])
    ctx.c = inputs[3]


def _ws_backward(ctx, gz):
    mu, sigma, eps = ctx.saved_tensors
    gmu, gsig = wrapped_sample_bwd(mu, sigma, eps, _c(gz), ctx.c)
    return gmu, gsig, None, None


wrapped_sample_fwd.register_autograd(_ws_backward, setup_context=_ws_setup)


@_op("hvae::wrapped_logprob_fwd", mutates_args=())
def wrapped_logprob_fwd(mu: Tensor, sigma: Tensor, z: Tensor, c: float) -> Tensor:
    """mu, sigma: (B,D); z: (S,B,D) -> logp (S,B)"""
    C.require_cuda(mu, sigma, z)
    S, B, D = z.shape
    out = z.new_empty(S, B)
    C.call("hvae_wrapped_logprob_fwd_f32", C.ptr(mu), C.ptr(sigma), 0.0, C.ptr(z), C.ptr(out), S, B, D, c, C.stream())
    return out


@wrapped_logprob_fwd.register_fake
def _(mu, sigma, z, c):
    return z.new_empty(z.shape[0], z.shape[1])


@_op("hvae::wrapped_logprob_bwd", mutates_args=())
def wrapped_logprob_bwd(mu: Tensor, sigma: Tensor, z: Tensor, g: Tensor, c: float) -> Tuple[Tensor, Tensor, Tensor]:
    C.require_cuda(mu, sigma, z, g)
    S, B, D = z.shape
    gmu, gsig, gz = torch.empty_like(mu), torch.empty_like(sigma), torch.empty_like(z)
    C.call("hvae_wrapped_logprob_bwd_f32", C.ptr(mu), C.ptr(sigma), 0.0, C.ptr(z), C.ptr(g), C.ptr(gmu), C.ptr(gsig),
           C.ptr(gz), S, B, D, c, C.stream())
    return gmu, gsig, gz


@wrapped_logprob_bwd.register_fake
def _(mu, sigma, z, g, c):
    return torch.empty_like(mu), torch.empty_like(sigma), torch.empty_like(z)


def _wl_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], inputs[2])
    ctx.c = inputs[3]


def _wl_backward(ctx, g):
    mu, sigma, z = ctx.saved_tensors
    gmu, gsig, gz = wrapped_logprob_bwd(mu, sigma, z, _c(g), ctx.c)
    return gmu, gsig, gz, None


wrapped_logprob_fwd.register_autograd(_wl_backward, setup_context=_wl_setup)


@_op("hvae::wrapped_logprob_prior_fwd", mutates_args=())
def wrapped_logprob_prior_fwd(z: Tensor, sigma0: float, c: float) -> Tensor:
    """log-density of WrappedNormal(origin, sigma0 * 1) at z (S,B,D) -> (S,B)"""
    C.require_cuda(z)
    S, B, D = z.shape
    out = z.new_empty(S, B)
    C.call("hvae_wrapped_logprob_fwd_f32", None, None, sigma0, C.ptr(z), C.ptr(out), S, B, D, c, C.stream())
    return out


@wrapped_logprob_prior_fwd.register_fake
def _(z, sigma0, c):
    return z.new_empty(z.shape[0], z.shape[1])


@_op("hvae::wrapped_logprob_prior_bwd", mutates_args=())
def wrapped_logprob_prior_bwd(z: Tensor, g: Tensor, sigma0: float, c: float) -> Tensor:
    C.require_cuda(z, g)
    S, B, D = z.shape
    gz = torch.empty_like(z)
    C.call("hvae_wrapped_logprob_bwd_f32", None, None, sigma0, C.ptr(z), C.ptr(g), None, None, C.ptr(gz), S, B, D, c,
           C.stream())
    return gz


@wrapped_logprob_prior_bwd.register_fake
def _(z, g, sigma0, c):
    return torch.empty_like(z)


def _wlp_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])
    ctx.sigma0, ctx.c = inputs[1], inputs[2]


def _wlp_backward(ctx, g):
    (z,) = ctx.saved_tensors
    return wrapped_logprob_prior_bwd(z, _c(g), ctx.sigma0, ctx.c), None, None


wrapped_logprob_prior_fwd.register_autograd(_wlp_backward, setup_context=_wlp_setup)


@_op("hvae::latent_head_fwd", mutates_args=())
def latent_head_fwd(mu: Tensor, sigma: Tensor, eps: Tensor, prior_scale: float, c: float) -> Tuple[Tensor, Tensor]:
    """Fused K4+K5: z = rsample(mu, sigma; eps) (B,D), kl = log q(z|x) - log p(z) (B,)"""
    C.require_cuda(mu, sigma, eps)
    B, D = mu.shape
    z, kl = torch.empty_like(mu), mu.new_empty(B)
    C.call("hvae_latent_head_fwd_f32", C.ptr(mu), C.ptr(sigma), C.ptr(eps), prior_scale, C.ptr(z), C.ptr(kl), B, D, c,
           C.stream())
    return z, kl


@latent_head_fwd.register_fake
def _(mu, sigma, eps, prior_scale, c):
    return torch.empty_like(mu), mu.new_empty(mu.shape[0])


@_op("hvae::latent_head_bwd", mutates_args=())
def latent_head_bwd(mu: Tensor, sigma: Tensor, eps: Tensor, gz: Optional[Tensor], gkl: Optional[Tensor],
                    prior_scale: float, c: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(mu, sigma, eps, gz, gkl)
    B, D = mu.shape
    gmu, gsig = torch.empty_like(mu), torch.empty_like(sigma)
    C.call("hvae_latent_head_bwd_f32", C.ptr(mu), C.ptr(sigma), C.ptr(eps), prior_scale, C.ptr(gz), C.ptr(gkl), C.ptr(gmu),
           C.ptr(gsig), B, D, c, C.stream())
    return gmu, gsig


@latent_head_bwd.register_fake
def _(mu, sigma, eps, gz, gkl, prior_scale, c):
    return torch.empty_like(mu), torch.empty_like(sigma)


def _lh_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)  # an unused output's gradient arrives as None, not as a zero-filled tensor
    ctx.save_for_backward(inputs[0], inputs[1], inputs[2])
    ctx.prior_scale, ctx.c = inputs[3], inputs[4]


def _lh_backward(ctx, gz, gkl):
    mu, sigma, eps = ctx.saved_tensors
    gz = None if gz is None else _c(gz)
    gkl = None if gkl is None else _c(gkl)
    gmu, gsig = latent_head_bwd(mu, sigma, eps, gz, gkl, ctx.prior_scale, ctx.c)
    return gmu, gsig, None, None, None


latent_head_fwd.register_autograd(_lh_backward, setup_context=_lh_setup)


def latent_head(mu: Tensor, sigma: Tensor, eps: Tensor, prior_scale: float, c: float) -> Tuple[Tensor, Tensor]:
    return latent_head_fwd(_c(mu), _c(sigma), _c(eps), float(prior_scale), c)


# ---------------------------------------------------------------------------------------------------
# flags shared with include/hvae_b200.h
# ---------------------------------------------------------------------------------------------------
GYRO_SIGNED, GYRO_SQUARED, GYRO_SCALED, GYRO_PVAE = 1, 2, 4, 8
GYRO_RELU = 16   # SIMT kernels only: the ReLU that follows the layer, fused (forward clamp, backward mask by the recomputed sign)


def log_sinhc(x: Tensor) -> Tensor:
    """log(sinh(x)/x), stable at small x (tensor expression; only used by the free-function logdetexp)."""
    x2 = x * x
    small = x2 * (1.0 / 6.0 + x2 * (-1.0 / 180.0 + x2 * (1.0 / 2835.0)))
    xl = x.clamp_min(0.5)
    large = xl + torch.log1p(-torch.exp(-2.0 * xl)) - 0.6931471805599453 - torch.log(xl)
    return torch.where(x < 0.5, small, large)


# ---------------------------------------------------------------------------------------------------
# K2 gyroplane
# ---------------------------------------------------------------------------------------------------
_WS = {}
_WS_RETIRED = []  # outgrown workspaces stay allocated: a captured CUDA graph may still hold their addresses


def _workspace(nbytes: int, device) -> Tensor:
    """Grow-only per-device scratch owned by torch's allocator (the C library allocates nothing)."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _WS_RETIRED.append(buf)
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


@_op("hvae::gyroplane_fwd", mutates_args=())
def gyroplane_fwd(x: Tensor, p: Tensor, a: Optional[Tensor], bias: Optional[Tensor], c: float, flags: int) -> Tensor:
    """x (B,D); p (P,D); a (P,D) or None (= p); bias (P,) or None -> (B,P)"""
    C.require_cuda(x, p, a, bias)
    B, D = x.shape
    P = p.shape[0]
    out = x.new_empty(B, P)
    C.call("hvae_gyroplane_fwd_f32", C.ptr(x), C.ptr(p), C.ptr(p if a is None else a), C.ptr(bias), C.ptr(out), B, D, P, c,
           flags, C.stream())
    return out


@gyroplane_fwd.register_fake
def _(x, p, a, bias, c, flags):
    return x.new_empty(x.shape[0], p.shape[0])


@_op("hvae::gyroplane_bwd", mutates_args=())
def gyroplane_bwd(x: Tensor, p: Tensor, a: Optional[Tensor], bias: Optional[Tensor], g: Tensor, c: float, flags: int,
                  need_bias: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """bias: the forward's bias, read only with GYRO_RELU (the mask is the sign of the recomputed out + bias)."""
    C.require_cuda(x, p, a, g, bias)
    B, D = x.shape
    P = p.shape[0]
    gx, gp = torch.empty_like(x), torch.empty_like(p)
    ga = torch.empty_like(a) if a is not None else x.new_empty(0)
    gb = x.new_empty(P) if need_bias else x.new_empty(0)
    nbytes = C.lib().hvae_gyroplane_bwd_workspace_bytes(B, D, P)
    ws = _workspace(nbytes, x.device)
    C.call("hvae_gyroplane_relu_bwd_f32", C.ptr(x), C.ptr(p), C.ptr(p if a is None else a), C.ptr(bias), C.ptr(g), C.ptr(gx), C.ptr(gp),
           C.ptr(ga) if a is not None else None, C.ptr(gb) if need_bias else None, B, D, P, c, flags, C.ptr(ws),
           ws.numel(), C.stream())
    C.launch_count += 3   # pair kernel, plane kernel, one slab reduction for gx / gp / ga / gbias
    return gx, gp, ga, gb


@gyroplane_bwd.register_fake
def _(x, p, a, bias, g, c, flags, need_bias):
    return (torch.empty_like(x), torch.empty_like(p), torch.empty_like(a) if a is not None else x.new_empty(0),
            x.new_empty(p.shape[0]) if need_bias else x.new_empty(0))


def _gy_setup(ctx, inputs, output):
    x, p, a, bias, c, flags = inputs
    relu_bias = bias is not None and bool(flags & GYRO_RELU)
    ctx.save_for_backward(x, p, a if a is not None else p, *([bias] if relu_bias else []))
    ctx.has_a, ctx.has_bias, ctx.c, ctx.flags = a is not None, bias is not None, c, flags


def _gy_backward(ctx, g):
    x, p, a = ctx.saved_tensors[:3]
    bias = ctx.saved_tensors[3] if len(ctx.saved_tensors) > 3 else None
    gx, gp, ga, gb = gyroplane_bwd(x, p, a if ctx.has_a else None, bias, _c(g), ctx.c, ctx.flags, ctx.has_bias)
    return gx, gp, (ga if ctx.has_a else None), (gb if ctx.has_bias else None), None, None


gyroplane_fwd.register_autograd(_gy_backward, setup_context=_gy_setup)


@_op("hvae::gyroplane_tc32_fwd", mutates_args=())
def gyroplane_tc32_fwd(x: Tensor, p: Tensor, a: Optional[Tensor], bias: Optional[Tensor], c: float, flags: int) -> Tensor:
    """K2 at fp32 accuracy for any latent dim and any (a, p): tensor-core GEMMs on the three-way split for the inner
    products and the gradient contractions, the pair function elementwise (csrc/gyro_tc32.cu):
    x (B,D); p (P,D); a (P,D) or None (= p); bias (P,) or None -> (B,P)."""
    C.require_cuda(x, p, a, bias)
    B, D = x.shape
    P = p.shape[0]
    out = x.new_empty(B, P)
    ws = _workspace(C.lib().hvae_gyroplane_tc32_fwd_workspace_bytes(B, D, P, int(a is not None)), x.device)
    C.call("hvae_gyroplane_tc32_fwd_f32", C.ptr(x), C.ptr(p), C.ptr(a) if a is not None else None, C.ptr(bias), C.ptr(out), B, D, P, c,
           flags, C.ptr(ws), ws.numel(), C.stream())
    C.launch_count += 7 + (2 if a is not None else 0)
    return out


@gyroplane_tc32_fwd.register_fake
def _(x, p, a, bias, c, flags):
    return x.new_empty(x.shape[0], p.shape[0])


@_op("hvae::gyroplane_tc32_bwd", mutates_args=())
def gyroplane_tc32_bwd(x: Tensor, p: Tensor, a: Optional[Tensor], g: Tensor, c: float, flags: int) -> Tuple[Tensor, Tensor, Tensor]:
    C.require_cuda(x, p, a, g)
    B, D = x.shape
    P = p.shape[0]
    gx, gp = torch.empty_like(x), torch.empty_like(p)
    ga = torch.empty_like(a) if a is not None else x.new_empty(0)
    ws = _workspace(C.lib().hvae_gyroplane_tc32_bwd_workspace_bytes(B, D, P, int(a is not None)), x.device)
    C.call("hvae_gyroplane_tc32_bwd_f32", C.ptr(x), C.ptr(p), C.ptr(a) if a is not None else None, C.ptr(g), C.ptr(gx), C.ptr(gp),
           C.ptr(ga) if a is not None else None, B, D, P, c, flags, C.ptr(ws), ws.numel(), C.stream())
    C.launch_count += 19 + (10 if a is not None else 0)
    return gx, gp, ga


@gyroplane_tc32_bwd.register_fake
def _(x, p, a, g, c, flags):
    return torch.empty_like(x), torch.empty_like(p), torch.empty_like(a) if a is not None else x.new_empty(0)


def _gytc32_setup(ctx, inputs, output):
    x, p, a, bias, c, flags = inputs
    ctx.save_for_backward(x, p, a if a is not None else p)
    ctx.has_a, ctx.has_bias, ctx.c, ctx.flags = a is not None, bias is not None, c, flags


def _gytc32_backward(ctx, g):
    x, p, a = ctx.saved_tensors
    g = _c(g)
    gx, gp, ga = gyroplane_tc32_bwd(x, p, a if ctx.has_a else None, g, ctx.c, ctx.flags)
    gb = colsum(g) if ctx.has_bias else None
    return gx, gp, (ga if ctx.has_a else None), gb, None, None


gyroplane_tc32_fwd.register_autograd(_gytc32_backward, setup_context=_gytc32_setup)

GYRO_SIMT_MAX_D = 64   # the SIMT kernels (csrc/gyroplane.cu) keep a row of x in registers


def gyroplane(x: Tensor, p: Tensor, a: Optional[Tensor], bias: Optional[Tensor], c: float, flags: int, relu: bool = False) -> Tensor:
    """Signed hyperplane distances of every row of x (..., D) to every plane -> (..., P).
    relu=True applies the ReLU that follows the layer: inside the SIMT kernels (GYRO_RELU), as a torch op on the other paths.
    bf16 GEMM mode: the fused tcgen05 kernels (a == p forward + backward; a != p forward under no_grad).  fp32 mode: the
    SIMT kernels up to D = 64, beyond that - and for the a != p backward at any GEMM-sized D - the fp32-accurate
    tensor-core path (split-operand GEMMs + the elementwise pair function)."""
    lead = x.shape[:-1]
    xr = _rows(x)
    a_arg = None if (a is None or a is p) else _c(a)
    b_arg = None if bias is None else _c(bias)
    if a_arg is None and _tc_eligible(xr.shape[0], xr.shape[1], p.shape[0]):
        out = gyroplane_tc_fwd(xr, _c(p), b_arg, c, int(flags))
    elif a_arg is not None and _tc_eligible(xr.shape[0], xr.shape[1], p.shape[0]) and not torch.is_grad_enabled():
        out = geodesic_tc_fwd(xr, _c(p), a_arg, b_arg, c, int(flags))
    elif xr.shape[1] > GYRO_SIMT_MAX_D:
        out = gyroplane_tc32_fwd(xr, _c(p), a_arg, b_arg, c, int(flags))
    else:
        out = gyroplane_fwd(xr, _c(p), a_arg, b_arg, c, int(flags) | (GYRO_RELU if relu else 0))
        relu = False
    if relu:
        out = torch.relu(out)
    return out.view(*lead, p.shape[0])


# ---------------------------------------------------------------------------------------------------
# K1b weight prep + K1 Mobius matvec
# ---------------------------------------------------------------------------------------------------
@_op("hvae::weight_prep_fwd", mutates_args=())
def weight_prep_fwd(W: Tensor, beta: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    """W (P,F) `_weight`, beta (P,) `_bias` -> (bias point expmap0(W*beta) (P,F), transported weight (P,F))"""
    C.require_cuda(W, beta)
    P, F = W.shape
    bpt, M = torch.empty_like(W), torch.empty_like(W)
    C.call("hvae_weight_prep_fwd_f32", C.ptr(W), C.ptr(beta), None, C.ptr(bpt), C.ptr(M), P, F, c, C.stream())
    return bpt, M


@weight_prep_fwd.register_fake
def _(W, beta, c):
    return torch.empty_like(W), torch.empty_like(W)


@_op("hvae::weight_prep_bwd", mutates_args=())
def weight_prep_bwd(W: Tensor, beta: Tensor, gM: Optional[Tensor], gbpt: Optional[Tensor], c: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(W, beta, gM, gbpt)
    P, F = W.shape
    gW, gbeta = torch.empty_like(W), torch.empty_like(beta)
    C.call("hvae_weight_prep_bwd_f32", C.ptr(W), C.ptr(beta), None, C.ptr(gM), C.ptr(gbpt), C.ptr(gW), C.ptr(gbeta), None,
           P, F, c, C.stream())
    return gW, gbeta


@weight_prep_bwd.register_fake
def _(W, beta, gM, gbpt, c):
    return torch.empty_like(W), torch.empty_like(beta)


def _wp_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.c = inputs[2]


def _wp_backward(ctx, gbpt, gM):
    W, beta = ctx.saved_tensors
    gW, gbeta = weight_prep_bwd(W, beta, None if gM is None else _c(gM), None if gbpt is None else _c(gbpt), ctx.c)
    return gW, gbeta, None


weight_prep_fwd.register_autograd(_wp_backward, setup_context=_wp_setup)


def weight_prep(W: Tensor, beta: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    """beta is the (P,1) `_bias` parameter of RiemannianLayer (over_param=False)."""
    return weight_prep_fwd(_c(W), _c(beta).view(-1), c)


@_op("hvae::weight_prep_op_fwd", mutates_args=())
def weight_prep_op_fwd(W: Tensor, bias_pt: Tensor, c: float) -> Tensor:
    """over_param=True: M = W * clamp_min(1 - c|bias_pt|^2)"""
    C.require_cuda(W, bias_pt)
    P, F = W.shape
    M = torch.empty_like(W)
    C.call("hvae_weight_prep_fwd_f32", C.ptr(W), None, C.ptr(bias_pt), None, C.ptr(M), P, F, c, C.stream())
    return M


@weight_prep_op_fwd.register_fake
def _(W, bias_pt, c):
    return torch.empty_like(W)


@_op("hvae::weight_prep_op_bwd", mutates_args=())
def weight_prep_op_bwd(W: Tensor, bias_pt: Tensor, gM: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(W, bias_pt, gM)
    P, F = W.shape
    gW, gb = torch.empty_like(W), torch.empty_like(bias_pt)
    C.call("hvae_weight_prep_bwd_f32", C.ptr(W), None, C.ptr(bias_pt), C.ptr(gM), None, C.ptr(gW), None, C.ptr(gb), P, F, c,
           C.stream())
    return gW, gb


@weight_prep_op_bwd.register_fake
def _(W, bias_pt, gM, c):
    return torch.empty_like(W), torch.empty_like(bias_pt)


def _wpo_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.c = inputs[2]


def _wpo_backward(ctx, gM):
    W, b = ctx.saved_tensors
    gW, gb = weight_prep_op_bwd(W, b, _c(gM), ctx.c)
    return gW, gb, None


weight_prep_op_fwd.register_autograd(_wpo_backward, setup_context=_wpo_setup)


def weight_prep_overparam(W: Tensor, bias_pt: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    return bias_pt, weight_prep_op_fwd(_c(W), _c(bias_pt), c)


@_op("hvae::mobius_matvec_fwd", mutates_args=())
def mobius_matvec_fwd(x: Tensor, M: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    """x (B,F), M (P,F) -> (y (B,P) projected Mobius matvec, mx (B,P) saved for backward)"""
    C.require_cuda(x, M)
    B, F = x.shape
    P = M.shape[0]
    y, mx = x.new_empty(B, P), x.new_empty(B, P)
    C.call("hvae_mobius_matvec_fwd_f32", C.ptr(x), C.ptr(M), C.ptr(y), C.ptr(mx), B, F, P, c, C.stream())
    return y, mx


@mobius_matvec_fwd.register_fake
def _(x, M, c):
    return x.new_empty(x.shape[0], M.shape[0]), x.new_empty(x.shape[0], M.shape[0])


@_op("hvae::mobius_matvec_bwd", mutates_args=())
def mobius_matvec_bwd(x: Tensor, M: Tensor, mx: Tensor, gy: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(x, M, mx, gy)
    B, F = x.shape
    P = M.shape[0]
    gx, gM = torch.empty_like(x), torch.empty_like(M)
    nbytes = C.lib().hvae_mobius_matvec_bwd_workspace_bytes(B, F, P)
    ws = _workspace(nbytes, x.device)
    C.call("hvae_mobius_matvec_bwd_f32", C.ptr(x), C.ptr(M), C.ptr(mx), C.ptr(gy), C.ptr(gx), C.ptr(gM), B, F, P, c,
           C.ptr(ws), ws.numel(), C.stream())
    C.launch_count += 2
    return gx, gM


@mobius_matvec_bwd.register_fake
def _(x, M, mx, gy, c):
    return torch.empty_like(x), torch.empty_like(M)


def _mm_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)  # the saved pre-activation mx never receives a gradient: no (B,P) zero fill
    ctx.save_for_backward(inputs[0], inputs[1], output[1])
    ctx.c = inputs[2]


def _mm_backward(ctx, gy, _gmx):
    if gy is None:
        return None, None, None
    x, M, mx = ctx.saved_tensors
    gx, gM = mobius_matvec_bwd(x, M, mx, _c(gy), ctx.c)
    return gx, gM, None


mobius_matvec_fwd.register_autograd(_mm_backward, setup_context=_mm_setup)


def mobius_matvec(x: Tensor, M: Tensor, c: float) -> Tensor:
    """project(M (x)_c x) for every row of x (..., F) -> (..., P)"""
    lead = x.shape[:-1]
    xr = _rows(x)
    if _tc_eligible(xr.shape[0], xr.shape[1], M.shape[0]):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or M.requires_grad)
        if M.shape[0] % 8 == 0 and (xr.shape[0] % 8 == 0 or not needs_grad):
            y, _ = mobius_matvec_tc(xr, _c(M), c)       # single pass; tensor-core backward
        else:
            y, _ = mobius_matvec_tc_fwd(xr, _c(M), c)   # ragged shapes: two-pass forward, fp32 backward
    else:
        y, _ = mobius_matvec_fwd(xr, _c(M), c)
    return y.view(*lead, M.shape[0])


# ---------------------------------------------------------------------------------------------------
# K6 / K7 HyperbolicRadius + expmap_polar
# ---------------------------------------------------------------------------------------------------
@_op("hvae::hradius_lognorm_fwd", mutates_args=())
def hradius_lognorm_fwd(sigma: Tensor, dim: int, c: float) -> Tuple[Tensor, Tensor]:
    """sigma (B,) -> (logZ (B,), dlogZ/dsigma (B,)); float64 series inside, float32 out (pvae semantics)."""
    C.require_cuda(sigma)
    logz, dlogz = torch.empty_like(sigma), torch.empty_like(sigma)
    C.call("hvae_hradius_lognorm_fwd_f32", C.ptr(sigma), C.ptr(logz), C.ptr(dlogz), sigma.numel(), dim, c, C.stream())
    return logz, dlogz


@hradius_lognorm_fwd.register_fake
def _(sigma, dim, c):
    return torch.empty_like(sigma), torch.empty_like(sigma)


def _hl_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(output[1])


def _hl_backward(ctx, g, _g2):
    if g is None:
        return None, None, None
    (dlogz,) = ctx.saved_tensors
    return g * dlogz, None, None


hradius_lognorm_fwd.register_autograd(_hl_backward, setup_context=_hl_setup)


def hradius_lognorm(sigma: Tensor, dim: int, c: float) -> Tensor:
    shape = sigma.shape
    return hradius_lognorm_fwd(_c(sigma).view(-1), int(dim), c)[0].view(shape)


# Noise stream of the in-kernel samplers.  Sample i of a call draws from Philox counters (seed, base + *counter + i):
#   * `counter` is a per-device int64 DEVICE scalar advanced in-stream after every call, so a captured CUDA graph
#     draws fresh noise on every replay (a host-side offset would be baked into the graph: every replay would repeat
#     the same radii).  It is the default; it must exist before capture starts (TrainStep's eager warm-up steps, or
#     philox_counter(device), create it).
#   * data parallel: set_noise_shard(lo, global_rows) makes a rank that owns rows [lo, lo + B_local) of a
#     global batch draw exactly the numbers the single-GPU run draws for those rows: base = lo and the counter advances by
#     the GLOBAL row count per call.
_philox_counters = {}
_noise_shard = (0, None)  # (first global row of this rank's shard, global rows per call or None = local rows)


def philox_counter(device) -> Tensor:
    """The per-device int64 noise counter (created on first use; creating it inside a graph capture is an error)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    t = _philox_counters.get(idx)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("hvae: the Philox noise counter must exist before CUDA-graph capture starts; run one "
                               "eager step first or call hvae.ops.philox_counter(device)")
        t = torch.zeros(1, dtype=torch.int64, device=torch.device("cuda", idx))
        _philox_counters[idx] = t
    return t


def set_noise_shard(lo: int, global_rows: Optional[int]):
    """Data-parallel noise rule: this rank owns rows [lo, lo + B_local) of a global batch of `global_rows` rows
    (hvae.parallel.philox_offset_for_shard).  (0, None) restores the single-process behaviour."""
    global _noise_shard
    _noise_shard = (int(lo), None if global_rows is None else int(global_rows))


def reset_noise(device=None):
    """Zero the device noise counter(s): two runs with the same torch seed then draw the same stream."""
    for idx, t in _philox_counters.items():
        if device is None or idx == (device.index if device.index is not None else torch.cuda.current_device()):
            t.zero_()


def hradius_sample(sigma: Tensor, S: int, dim: int, c: float, seed: Optional[int] = None, offset: Optional[int] = None,
                   offset_dev: Optional[Tensor] = None) -> Tensor:
    """r (S,B) ~ rho(.; sigma_b) by in-kernel rejection sampling (Philox4x32-10).  No gradient.
    offset: explicit host counter base (deterministic tests) - then no device counter is used unless offset_dev is given.
    Default: the per-device counter + this rank's shard rule (see above)."""
    C.require_cuda(sigma)
    sig = _c(sigma.detach()).view(-1)
    B = sig.numel()
    if seed is None:
        seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    r = sig.new_empty(S, B)
    if offset is not None:
        C.call("hvae_hradius_sample_f32", C.ptr(sig), C.ptr(r), S, B, dim, c, seed, offset, C.ptr(offset_dev), C.stream())
        if offset_dev is not None:
            offset_dev.add_(S * B)
        return r
    if offset_dev is None:
        offset_dev = philox_counter(sig.device)
    lo, grows = _noise_shard
    if grows is None or grows == B:
        C.call("hvae_hradius_sample_f32", C.ptr(sig), C.ptr(r), S, B, dim, c, seed, lo, C.ptr(offset_dev), C.stream())
        offset_dev.add_(S * B)
    else:  # sharded: sample s of global row g uses counter s * global_rows + g, as the single-GPU call would
        for s_ in range(S):
            C.call("hvae_hradius_sample_f32", C.ptr(sig), C.ptr(r[s_]), 1, B, dim, c, seed, s_ * grows + lo,
                   C.ptr(offset_dev), C.stream())
        offset_dev.add_(S * grows)
    return r


def sphere_sample(S: int, B: int, D: int, device, seed: Optional[int] = None, offset: Optional[int] = None) -> Tensor:
    """alpha (S,B,D) ~ U(S^{D-1}) drawn in the kernel (Philox).  Same counter rules as hradius_sample: by default the
    per-device counter (graph-replay safe) and this rank's noise shard - sample s of global row g uses counter
    s * global_rows + g - on a stream disjoint from the radii's.  No gradient."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("hvae ops run only on CUDA (sm_100a) tensors; there is no CPU fallback")
    if seed is None:
        seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    out = torch.empty(S, B, D, dtype=torch.float32, device=dev)
    if offset is not None:
        C.call("hvae_sphere_sample_f32", C.ptr(out), S * B, D, seed, offset, None, C.stream())
        return out
    ctr = philox_counter(dev)
    lo, grows = _noise_shard
    if grows is None or grows == B:
        C.call("hvae_sphere_sample_f32", C.ptr(out), S * B, D, seed, lo, C.ptr(ctr), C.stream())
        ctr.add_(S * B)
    else:
        for s_ in range(S):
            C.call("hvae_sphere_sample_f32", C.ptr(out[s_]), B, D, seed, s_ * grows + lo, C.ptr(ctr), C.stream())
        ctr.add_(S * grows)
    return out


@_op("hvae::hradius_reparam", mutates_args=())
def hradius_reparam(r: Tensor, sigma: Tensor, dim: int, c: float) -> Tuple[Tensor, Tensor]:
    """Identity on r (S,B) carrying the implicit-reparameterisation gradient dr/dsigma (pvae impl_rsample)."""
    C.require_cuda(r, sigma)
    S, B = r.shape
    dr = torch.empty_like(r)
    C.call("hvae_hradius_rgrad_f32", C.ptr(sigma), C.ptr(r), C.ptr(dr), None, S, B, dim, c, C.stream())
    return r.clone(), dr


@hradius_reparam.register_fake
def _(r, sigma, dim, c):
    return torch.empty_like(r), torch.empty_like(r)


def _hr_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(output[1])


def _hr_backward(ctx, g, _g2):
    if g is None:
        return None, None, None, None
    (dr,) = ctx.saved_tensors
    return None, (g * dr).sum(0), None, None


hradius_reparam.register_autograd(_hr_backward, setup_context=_hr_setup)


def hradius_cdf(r: Tensor, sigma: Tensor, dim: int, c: float) -> Tensor:
    C.require_cuda(r, sigma)
    S, B = r.shape
    dr, cdf = torch.empty_like(r), torch.empty_like(r)
    C.call("hvae_hradius_rgrad_f32", C.ptr(_c(sigma).view(-1)), C.ptr(_c(r)), C.ptr(dr), C.ptr(cdf), S, B, dim, c, C.stream())
    return cdf


@_op("hvae::expmap_polar_fwd", mutates_args=())
def expmap_polar_fwd(mu: Tensor, alpha: Tensor, r: Tensor, c: float) -> Tensor:
    """mu (B,D); alpha (S,B,D) directions; r (S,B) radii -> z (S,B,D)"""
    C.require_cuda(mu, alpha, r)
    S, B, D = alpha.shape
    z = torch.empty_like(alpha)
    C.call("hvae_expmap_polar_fwd_f32", C.ptr(mu), C.ptr(alpha), C.ptr(r), C.ptr(z), S, B, D, c, C.stream())
    return z


@expmap_polar_fwd.register_fake
def _(mu, alpha, r, c):
    return torch.empty_like(alpha)


@_op("hvae::expmap_polar_bwd", mutates_args=())
def expmap_polar_bwd(mu: Tensor, alpha: Tensor, r: Tensor, gz: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(mu, alpha, r, gz)
    S, B, D = alpha.shape
    gmu, gr = torch.empty_like(mu), torch.empty_like(r)
    C.call("hvae_expmap_polar_bwd_f32", C.ptr(mu), C.ptr(alpha), C.ptr(r), C.ptr(gz), C.ptr(gmu), C.ptr(gr), S, B, D, c,
           C.stream())
    return gmu, gr


@expmap_polar_bwd.register_fake
def _(mu, alpha, r, gz, c):
    return torch.empty_like(mu), torch.empty_like(r)


def _ep_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], inputs[2])
    ctx.c = inputs[3]


def _ep_backward(ctx, gz):
    mu, alpha, r = ctx.saved_tensors
    gmu, gr = expmap_polar_bwd(mu, alpha, r, _c(gz), ctx.c)
    return gmu, None, gr, None


expmap_polar_fwd.register_autograd(_ep_backward, setup_context=_ep_setup)


def expmap_polar(mu: Tensor, alpha: Tensor, r: Tensor, c: float) -> Tensor:
    """pvae PoincareBall.expmap_polar(x, u, r): mu (..., D) broadcast against alpha (S, ..., D), r (S, ..., 1)."""
    D = alpha.shape[-1]
    if alpha.dim() == mu.dim():
        alpha, r = alpha.unsqueeze(0), r.unsqueeze(0)
        squeeze = True
    else:
        squeeze = False
    S = alpha.shape[0]
    mu_b = mu.expand(alpha.shape[1:]) if mu.shape != alpha.shape[1:] else mu
    B = mu_b.numel() // D
    z = expmap_polar_fwd(_c(mu_b).view(B, D), _c(alpha).view(S, B, D), _c(r.expand(*alpha.shape[:-1], 1)).view(S, B), c)
    z = z.view(alpha.shape)
    return z.squeeze(0) if squeeze else z


# ---------------------------------------------------------------------------------------------------
# K1-TC / K2-TC: tcgen05 bf16 GEMM paths (forward) for GEMM-sized shapes
# ---------------------------------------------------------------------------------------------------
_gemm_mode = "fp32"


def set_gemm_mode(mode: str):
    """'fp32' (default: SIMT fp32 kernels, 1e-5 parity) or 'bf16' (tcgen05 GEMMs with bf16 operands / fp32
    accumulation for GEMM-sized Mobius / gyroplane shapes, 1e-2 parity — BASELINE.json's "bf16 GEMM mode")."""
    global _gemm_mode
    if mode not in ("fp32", "bf16"):
        raise ValueError(mode)
    _gemm_mode = mode


def get_gemm_mode() -> str:
    return _gemm_mode


def _tc_eligible(B: int, K: int, P: int) -> bool:
    return _gemm_mode == "bf16" and B >= 128 and P >= 128 and K >= 64 and K % 8 == 0


@_op("hvae::mobius_matvec_tc_fwd", mutates_args=())
def mobius_matvec_tc_fwd(x: Tensor, M: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    """Two-pass variant (any P): the GEMM materialises mx, a light row pass rescales; backward = the fp32 kernels."""
    C.require_cuda(x, M)
    B, F = x.shape
    P = M.shape[0]
    y, mx = x.new_empty(B, P), x.new_empty(B, P)
    ws = _workspace(C.lib().hvae_tc_workspace_bytes(B, F, P), x.device)
    C.call("hvae_mobius_matvec_tc_fwd_f32", C.ptr(x), C.ptr(M), C.ptr(y), C.ptr(mx), None, B, F, P, c, C.ptr(ws),
           ws.numel(), C.stream())
    C.launch_count += 4
    return y, mx


@mobius_matvec_tc_fwd.register_fake
def _(x, M, c):
    return x.new_empty(x.shape[0], M.shape[0]), x.new_empty(x.shape[0], M.shape[0])


mobius_matvec_tc_fwd.register_autograd(_mm_backward, setup_context=_mm_setup)  # backward: the fp32 kernels on the saved mx


@_op("hvae::mobius_matvec_tc", mutates_args=())
def mobius_matvec_tc(x: Tensor, M: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    """Single-pass variant: |mx_b|^2 = x_b^T (M^T M) x_b from the Gram matrix, rescale + projection fused into the
    main GEMM's epilogue; mx is never materialised.  Returns (y, |mx|^2) — all the tensor-core backward needs."""
    C.require_cuda(x, M)
    B, F = x.shape
    P = M.shape[0]
    y, mxsq = x.new_empty(B, P), x.new_empty(B)
    ws = _workspace(C.lib().hvae_tc_workspace_bytes(B, F, P), x.device)
    C.call("hvae_mobius_matvec_tc_fwd_f32", C.ptr(x), C.ptr(M), C.ptr(y), None, C.ptr(mxsq), B, F, P, c, C.ptr(ws),
           ws.numel(), C.stream())
    C.launch_count += 7
    return y, mxsq


@mobius_matvec_tc.register_fake
def _(x, M, c):
    return x.new_empty(x.shape[0], M.shape[0]), x.new_empty(x.shape[0])


@_op("hvae::mobius_matvec_tc_bwd", mutates_args=())
def mobius_matvec_tc_bwd(x: Tensor, M: Tensor, y: Tensor, mxsq: Tensor, gy: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(x, M, y, mxsq, gy)
    B, F = x.shape
    P = M.shape[0]
    gx, gM = torch.empty_like(x), torch.empty_like(M)
    ws = _workspace(C.lib().hvae_mobius_tc_bwd_workspace_bytes(B, F, P), x.device)
    C.call("hvae_mobius_matvec_tc_bwd_f32", C.ptr(x), C.ptr(M), C.ptr(y), C.ptr(mxsq), C.ptr(gy), C.ptr(gx), C.ptr(gM),
           B, F, P, c, C.ptr(ws), ws.numel(), C.stream())
    C.launch_count += 6
    return gx, gM


@mobius_matvec_tc_bwd.register_fake
def _(x, M, y, mxsq, gy, c):
    return torch.empty_like(x), torch.empty_like(M)


def _mmtc_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(inputs[0], inputs[1], output[0], output[1])
    ctx.c = inputs[2]


def _mmtc_backward(ctx, gy, _g):
    if gy is None:
        return None, None, None
    x, M, y, mxsq = ctx.saved_tensors
    gx, gM = mobius_matvec_tc_bwd(x, M, y, mxsq, _c(gy), ctx.c)
    return gx, gM, None


mobius_matvec_tc.register_autograd(_mmtc_backward, setup_context=_mmtc_setup)


@_op("hvae::gyroplane_tc_fwd", mutates_args=())
def gyroplane_tc_fwd(x: Tensor, p: Tensor, bias: Optional[Tensor], c: float, flags: int) -> Tensor:
    C.require_cuda(x, p, bias)
    B, D = x.shape
    P = p.shape[0]
    out = x.new_empty(B, P)
    ws = _workspace(C.lib().hvae_tc_workspace_bytes(B, D, P), x.device)
    C.call("hvae_gyroplane_tc_fwd_f32", C.ptr(x), C.ptr(p), C.ptr(bias), C.ptr(out), B, D, P, c, flags, C.ptr(ws),
           ws.numel(), C.stream())
    C.launch_count += 2
    return out


@gyroplane_tc_fwd.register_fake
def _(x, p, bias, c, flags):
    return x.new_empty(x.shape[0], p.shape[0])


@_op("hvae::geodesic_tc_fwd", mutates_args=())
def geodesic_tc_fwd(x: Tensor, p: Tensor, a: Tensor, bias: Optional[Tensor], c: float, flags: int) -> Tensor:
    """a != p gyroplane (GeodesicLayer) on the tensor cores: one N-concatenated cta_group::2 GEMM + fused pair epilogue."""
    C.require_cuda(x, p, a, bias)
    B, D = x.shape
    P = p.shape[0]
    out = x.new_empty(B, P)
    ws = _workspace(C.lib().hvae_geodesic_tc_workspace_bytes(B, D, P), x.device)
    C.call("hvae_geodesic_tc_fwd_f32", C.ptr(x), C.ptr(p), C.ptr(a), C.ptr(bias), C.ptr(out), B, D, P, c, flags, C.ptr(ws),
           ws.numel(), C.stream())
    C.launch_count += 2
    return out


@geodesic_tc_fwd.register_fake
def _(x, p, a, bias, c, flags):
    return x.new_empty(x.shape[0], p.shape[0])


def _gtc_setup(ctx, inputs, output):
    x, p, bias, c, flags = inputs
    ctx.save_for_backward(x, p)
    ctx.has_bias, ctx.c, ctx.flags = bias is not None, c, flags


@_op("hvae::gyroplane_tc_bwd", mutates_args=())
def gyroplane_tc_bwd(x: Tensor, p: Tensor, g: Tensor, c: float, flags: int) -> Tuple[Tensor, Tensor]:
    C.require_cuda(x, p, g)
    B, D = x.shape
    P = p.shape[0]
    gx, gp = torch.empty_like(x), torch.empty_like(p)
    ws = _workspace(C.lib().hvae_gyroplane_tc_bwd_workspace_bytes(B, D, P), x.device)
    C.call("hvae_gyroplane_tc_bwd_f32", C.ptr(x), C.ptr(p), C.ptr(g), C.ptr(gx), C.ptr(gp), B, D, P, c, flags, C.ptr(ws),
           ws.numel(), C.stream())
    C.launch_count += 11
    return gx, gp


@gyroplane_tc_bwd.register_fake
def _(x, p, g, c, flags):
    return torch.empty_like(x), torch.empty_like(p)


def _gtc_backward(ctx, g):
    x, p = ctx.saved_tensors
    g = _c(g)
    B, D = x.shape
    P = p.shape[0]
    if B % 8 == 0 and D % 8 == 0 and P % 8 == 0:
        gx, gp = gyroplane_tc_bwd(x, p, g, ctx.c, ctx.flags)          # tensor cores
        gb = colsum(g) if ctx.has_bias else None
        return gx, gp, gb, None, None
    if D > 64:
        raise NotImplementedError("gyroplane backward for D > 64 needs B, D, P to be multiples of 8 (tensor-core path)")
    gx, gp, _, gb = gyroplane_bwd(x, p, None, None, g, ctx.c, ctx.flags, ctx.has_bias)
    return gx, gp, (gb if ctx.has_bias else None), None, None


gyroplane_tc_fwd.register_autograd(_gtc_backward, setup_context=_gtc_setup)


# ---------------------------------------------------------------------------------------------------
# fused Monte-Carlo KL of two Riemannian normals (posterior vs origin prior)
# ---------------------------------------------------------------------------------------------------
@_op("hvae::rn_kl_fwd", mutates_args=())
def rn_kl_fwd(mu: Tensor, sigma_q: Tensor, logz_q: Tensor, z: Tensor, sigma_p: Tensor, logz_p: Tensor, c: float) -> Tensor:
    """mu (B,D); sigma_q, logz_q (B,); z (S,B,D); sigma_p, logz_p 1-element device tensors -> kl (S,B)"""
    C.require_cuda(mu, sigma_q, logz_q, z, sigma_p, logz_p)
    S, B, D = z.shape
    kl = z.new_empty(S, B)
    C.call("hvae_rn_kl_fwd_f32", C.ptr(mu), C.ptr(sigma_q), C.ptr(logz_q), C.ptr(z), C.ptr(sigma_p), C.ptr(logz_p), C.ptr(kl),
           S, B, D, c, C.stream())
    return kl


@rn_kl_fwd.register_fake
def _(mu, sigma_q, logz_q, z, sigma_p, logz_p, c):
    return z.new_empty(z.shape[0], z.shape[1])


@_op("hvae::rn_kl_bwd", mutates_args=())
def rn_kl_bwd(mu: Tensor, sigma_q: Tensor, z: Tensor, sigma_p: Tensor, g: Tensor, c: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    C.require_cuda(mu, sigma_q, z, sigma_p, g)
    S, B, D = z.shape
    gmu, gs, glz, gz = torch.empty_like(mu), torch.empty_like(sigma_q), torch.empty_like(sigma_q), torch.empty_like(z)
    C.call("hvae_rn_kl_bwd_f32", C.ptr(mu), C.ptr(sigma_q), C.ptr(z), C.ptr(sigma_p), C.ptr(g), C.ptr(gmu), C.ptr(gs),
           C.ptr(glz), C.ptr(gz), S, B, D, c, C.stream())
    return gmu, gs, glz, gz


@rn_kl_bwd.register_fake
def _(mu, sigma_q, z, sigma_p, g, c):
    return torch.empty_like(mu), torch.empty_like(sigma_q), torch.empty_like(sigma_q), torch.empty_like(z)


def _rk_setup(ctx, inputs, output):
    mu, sigma_q, logz_q, z, sigma_p, logz_p, c = inputs
    ctx.save_for_backward(mu, sigma_q, z, sigma_p)
    ctx.c = c


def _rk_backward(ctx, g):
    mu, sigma_q, z, sigma_p = ctx.saved_tensors
    gmu, gs, glz, gz = rn_kl_bwd(mu, sigma_q, z, sigma_p, _c(g), ctx.c)
    return gmu, gs, glz, gz, None, None, None


rn_kl_fwd.register_autograd(_rk_backward, setup_context=_rk_setup)


# ---- the RiemannianNormal head fused with the sample (S = 1): csrc/riemannian_kl.cu k_rn_head_{fwd,bwd} ----
@_op("hvae::rn_head", mutates_args=())
def rn_head_fwd(mu: Tensor, sigma_q: Tensor, logz_q: Tensor, dlogz: Tensor, alpha: Tensor, r: Tensor, dr: Tensor,
                sigma_p: Tensor, logz_p: Tensor, c: float) -> Tuple[Tensor, Tensor]:
    """z = expmap_polar(mu, alpha, r) (B,D) and kl = log q(z) - log p(z) (B,) in one kernel.  sigma_q, logz_q, dlogz
    (= dlogZ/dsigma), r, dr (= dr/dsigma, implicit reparameterisation): (B,); sigma_p, logz_p: device scalars.  Only mu
    and sigma_q are differentiable inputs: the backward returns their TOTAL gradients (KL + sample path + normaliser)."""
    C.require_cuda(mu, sigma_q, logz_q, dlogz, alpha, r, dr, sigma_p, logz_p)
    B, D = mu.shape
    z, kl = torch.empty_like(mu), mu.new_empty(B)
    C.call("hvae_rn_head_fwd_f32", C.ptr(mu), C.ptr(alpha), C.ptr(r), C.ptr(sigma_q), C.ptr(logz_q), C.ptr(sigma_p), C.ptr(logz_p),
           C.ptr(z), C.ptr(kl), B, D, c, C.stream())
    return z, kl


@rn_head_fwd.register_fake
def _(mu, sigma_q, logz_q, dlogz, alpha, r, dr, sigma_p, logz_p, c):
    return torch.empty_like(mu), mu.new_empty(mu.shape[0])


@_op("hvae::rn_head_bwd", mutates_args=())
def rn_head_bwd(mu: Tensor, sigma_q: Tensor, dlogz: Tensor, alpha: Tensor, r: Tensor, dr: Tensor, sigma_p: Tensor, z: Tensor,
                gz: Optional[Tensor], gkl: Optional[Tensor], c: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(mu, sigma_q, dlogz, alpha, r, dr, sigma_p, z, gz, gkl)
    B, D = mu.shape
    gmu, gs = torch.empty_like(mu), torch.empty_like(sigma_q)
    C.call("hvae_rn_head_bwd_f32", C.ptr(mu), C.ptr(alpha), C.ptr(r), C.ptr(sigma_q), C.ptr(sigma_p), C.ptr(z), C.ptr(dr), C.ptr(dlogz),
           C.ptr(gz), C.ptr(gkl), C.ptr(gmu), C.ptr(gs), B, D, c, C.stream())
    return gmu, gs


@rn_head_bwd.register_fake
def _(mu, sigma_q, dlogz, alpha, r, dr, sigma_p, z, gz, gkl, c):
    return torch.empty_like(mu), torch.empty_like(sigma_q)


def _rnh_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    mu, sigma_q, logz_q, dlogz, alpha, r, dr, sigma_p, logz_p, c = inputs
    ctx.save_for_backward(mu, sigma_q, dlogz, alpha, r, dr, sigma_p, output[0])
    ctx.c = c


def _rnh_backward(ctx, gz, gkl):
    if gz is None and gkl is None:
        return (None,) * 10
    mu, sigma_q, dlogz, alpha, r, dr, sigma_p, z = ctx.saved_tensors
    gmu, gs = rn_head_bwd(mu, sigma_q, dlogz, alpha, r, dr, sigma_p, z, None if gz is None else _c(gz),
                          None if gkl is None else _c(gkl), ctx.c)
    return gmu, gs, None, None, None, None, None, None, None, None


rn_head_fwd.register_autograd(_rnh_backward, setup_context=_rnh_setup)


def hradius_rgrad(r: Tensor, sigma: Tensor, dim: int, c: float) -> Tensor:
    """dr/dsigma of the implicit reparameterisation for given radii r (S,B), sigma (B,) (no autograd)."""
    C.require_cuda(r, sigma)
    S, B = r.shape
    dr = torch.empty_like(r)
    C.call("hvae_hradius_rgrad_f32", C.ptr(sigma), C.ptr(r), C.ptr(dr), None, S, B, dim, c, C.stream())
    return dr


# ---------------------------------------------------------------------------------------------------
# Trunk dense layers: fp32-accurate GEMM on the tensor cores (three-way bf16 split, six products)
# ---------------------------------------------------------------------------------------------------
_trunk_mode = "x2"


def set_trunk_mode(mode: str):
    """'x2' (default): hvae.layers.Linear runs on the tcgen05 fp16 two-piece GEMM (power-of-two row scales, three piece
    products; fp32 accuracy, fp32 in/out);  'x3': the three-way bf16 split (six products);  'torch': it defers to
    torch.nn.functional.linear (cuBLAS fp32 FMA kernels)."""
    global _trunk_mode
    if mode not in ("x2", "x3", "torch"):
        raise ValueError(mode)
    _trunk_mode = mode


def get_trunk_mode() -> str:
    return _trunk_mode


# Below this many flops (2MNK) a dense layer stays on cuBLAS fp32: the tensor-core path is a pipeline of 5-7 launches per
# direction (operand maxima, splits, GEMM, reductions) - ~15 us of fixed cost that only pays off once the GEMM itself is
# worth that much (measured cross-over ~0.5 GFLOP: config 1's 128 x 784 x 64 layers were 16 % slower on it).
_trunk_min_flops = 2.5e8


def set_trunk_min_flops(flops: float) -> float:
    """Smallest 2MNK routed to the tensor-core trunk path; returns the previous value (tests pass 0 to force the path)."""
    global _trunk_min_flops
    old, _trunk_min_flops = _trunk_min_flops, float(flops)
    return old


def trunk_x3_eligible(x: Tensor, weight: Tensor) -> bool:
    """GEMM-sized fp32 CUDA problem and a tensor-core trunk mode ('x2' or 'x3')."""
    rows = x.numel() // x.shape[-1]
    return (_trunk_mode in ("x2", "x3") and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32
            and rows >= 128 and weight.shape[0] >= 64 and weight.shape[1] >= 64
            and 2.0 * rows * weight.shape[0] * weight.shape[1] >= _trunk_min_flops)


@_op("hvae::gemm_x3", mutates_args=())
def gemm_x3(A: Tensor, a_trans: bool, B: Tensor, b_trans: bool, bias: Optional[Tensor], relu: bool) -> Tensor:
    """C (M,N) = opA (M,K) . opB (N,K)^T (+ bias) (ReLU); *_trans: the operand is stored (K,M) / (K,N)."""
    C.require_cuda(A, B)
    K, M = (A.shape[0], A.shape[1]) if a_trans else (A.shape[1], A.shape[0])
    Kb, N = (B.shape[0], B.shape[1]) if b_trans else (B.shape[1], B.shape[0])
    if K != Kb:
        raise RuntimeError("gemm_x3: contraction mismatch %d vs %d" % (K, Kb))
    out = A.new_empty(M, N)
    ws = _workspace(C.lib().hvae_gemm_x3_workspace_bytes(M, N, K), A.device)
    C.call("hvae_gemm_x3_f32", C.ptr(A), int(a_trans), C.ptr(B), int(b_trans), C.ptr(bias), int(relu), C.ptr(out), M, N, K,
           C.ptr(ws), ws.numel(), C.stream())
    C.launch_count += C.lib().hvae_gemm_x3_num_launches(M, N, K)
    return out


@gemm_x3.register_fake
def _(A, a_trans, B, b_trans, bias, relu):
    M = A.shape[1] if a_trans else A.shape[0]
    N = B.shape[1] if b_trans else B.shape[0]
    return A.new_empty(M, N)


@_op("hvae::split3", mutates_args=())
def split3(x: Tensor) -> Tensor:
    """(rows, cols) fp32 -> (rows, 3*Cp) bf16 [hi | mid | lo], Cp = cols rounded up to 64 (zero padded)."""
    C.require_cuda(x)
    rows, cols = x.shape
    cp = (cols + 63) // 64 * 64
    out = torch.empty(rows, 3 * cp, dtype=torch.bfloat16, device=x.device)
    C.call("hvae_split3_f32", C.ptr(x), C.ptr(out), rows, cols, C.stream())
    C.launch_count += 1
    return out


@split3.register_fake
def _(x):
    return torch.empty(x.shape[0], 3 * ((x.shape[1] + 63) // 64 * 64), dtype=torch.bfloat16, device=x.device)


@_op("hvae::gemm_x3s", mutates_args=())
def gemm_x3s(As: Tensor, a_mn: bool, Bs: Tensor, b_mn: bool, bias: Optional[Tensor], relu: bool, M: int, N: int,
             K: int) -> Tensor:
    """C (M,N) = opA (M,K) . opB (N,K)^T on split3() operands; *_mn: the operand is the split of a (K,M) / (K,N)
    matrix (contraction over its rows), read MN-major by the tensor core — no transpose."""
    C.require_cuda(bias)
    if not (As.is_cuda and Bs.is_cuda and As.dtype == torch.bfloat16 and Bs.dtype == torch.bfloat16):
        raise RuntimeError("gemm_x3s: operands must be CUDA bf16 split3() buffers")
    out = torch.empty(M, N, dtype=torch.float32, device=As.device)
    ws = _workspace(C.lib().hvae_gemm_x3s_workspace_bytes(M, N), As.device)
    C.call("hvae_gemm_x3s_f32", C.ptr(As), int(a_mn), C.ptr(Bs), int(b_mn), C.ptr(bias), int(relu), C.ptr(out), M, N, K,
           C.ptr(ws), ws.numel(), C.stream())
    C.launch_count += C.lib().hvae_gemm_x3s_num_launches(M, N, K)
    return out


@gemm_x3s.register_fake
def _(As, a_mn, Bs, b_mn, bias, relu, M, N, K):
    return torch.empty(M, N, dtype=torch.float32, device=As.device)


@_op("hvae::colsum", mutates_args=())
def colsum(x: Tensor) -> Tensor:
    """(R, C) -> (C,) column sums (a dense layer's bias gradient); deterministic two-pass kernel."""
    C.require_cuda(x)
    R, Cn = x.shape
    out = x.new_empty(Cn)
    ws = _workspace(C.lib().hvae_colsum_workspace_bytes(Cn), x.device)
    C.call("hvae_colsum_f32", C.ptr(x), C.ptr(out), R, Cn, C.ptr(ws), ws.numel(), C.stream())
    C.launch_count += 2
    return out


@colsum.register_fake
def _(x):
    return x.new_empty(x.shape[1])


def _cp64(n: int) -> int:
    return (n + 63) // 64 * 64


@_op("hvae::split3_both", mutates_args=())
def split3_both(x: Tensor, want_rows: bool, want_t: bool) -> Tuple[Tensor, Tensor]:
    """(rows, cols) fp32 -> (rows split (rows, 3*Cp), transposed split (cols, 3*Rp)) from one read of x; a layout that
    is not wanted comes back empty."""
    C.require_cuda(x)
    rows, cols = x.shape
    r = torch.empty((rows, 3 * _cp64(cols)) if want_rows else (0,), dtype=torch.bfloat16, device=x.device)
    t = torch.empty((cols, 3 * _cp64(rows)) if want_t else (0,), dtype=torch.bfloat16, device=x.device)
    C.call("hvae_split3_both_f32", C.ptr(x), C.ptr(r) if want_rows else None, C.ptr(t) if want_t else None, rows, cols,
           C.stream())
    return r, t


@split3_both.register_fake
def _(x, want_rows, want_t):
    rows, cols = x.shape
    return (torch.empty((rows, 3 * _cp64(cols)) if want_rows else (0,), dtype=torch.bfloat16, device=x.device),
            torch.empty((cols, 3 * _cp64(rows)) if want_t else (0,), dtype=torch.bfloat16, device=x.device))


@_op("hvae::linear_x3", mutates_args=())
def linear_x3_fwd(x: Tensor, weight: Tensor, bias: Optional[Tensor], need_gx: bool, need_gw: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (y, transposed split of x, transposed split of W).  Every tensor of a dense layer is needed in both layouts
    (x: forward + weight gradient, W: forward + input gradient), so each is read once and split both ways here; the
    transposed splits are kept for the backward (empty when that gradient is not needed)."""
    xs, xts = split3_both(x, True, need_gw)
    ws, wts = split3_both(weight, True, need_gx)
    y = gemm_x3s(xs, False, ws, False, bias, False, x.shape[0], weight.shape[0], x.shape[1])
    return y, xts, wts


@linear_x3_fwd.register_fake
def _(x, weight, bias, need_gx, need_gw):
    M, K = x.shape
    N = weight.shape[0]
    return (x.new_empty(M, N), torch.empty((K, 3 * _cp64(M)) if need_gw else (0,), dtype=torch.bfloat16, device=x.device),
            torch.empty((K, 3 * _cp64(N)) if need_gx else (0,), dtype=torch.bfloat16, device=x.device))


def _lx3_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)  # the kept splits (19 MB of bf16 at config 2) never receive a gradient
    x, weight, bias, need_gx, need_gw = inputs
    ctx.save_for_backward(output[1], output[2])
    ctx.dims = (x.shape[0], weight.shape[0], x.shape[1])  # rows, out, in
    ctx.has_bias = bias is not None
    ctx.need = (need_gx, need_gw)


def _lx3_backward(ctx, gy, _g1, _g2):
    # gy is read once and split both ways: rows layout (contraction over `out`) for the input gradient, transposed
    # (contraction over the batch) for the weight gradient; x^T and W^T splits come from the forward.
    if gy is None:
        return None, None, None, None, None
    xts, wts = ctx.saved_tensors
    M, n_out, n_in = ctx.dims
    need_gx, need_gw = ctx.need
    gy = _c(gy)
    gx = gw = gb = None
    want_gx = ctx.needs_input_grad[0] and need_gx
    want_gw = ctx.needs_input_grad[1] and need_gw
    if want_gx or want_gw:
        gs, gts = split3_both(gy, want_gx, want_gw)
        if want_gx:
            gx = gemm_x3s(gs, False, wts, False, None, False, M, n_in, n_out)    # gy (M,out) . W (out,in)
        if want_gw:
            gw = gemm_x3s(gts, False, xts, False, None, False, n_out, n_in, M)   # gy^T (out,M) . x (M,in)
    if ctx.has_bias and ctx.needs_input_grad[2]:
        gb = colsum(gy)
    return gx, gw, gb, None, None


linear_x3_fwd.register_autograd(_lx3_backward, setup_context=_lx3_setup)


# ---- fp16 two-piece path (csrc/tc_x2.cu) ----
@_op("hvae::split2h_both", mutates_args=())
def split2h_both(x: Tensor, want_rows: bool, want_t: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(rows, cols) fp32 -> (rows split (rows, 2*Cp) fp16, its inverse row scales (rows,), transposed split (cols, 2*Rp)
    scaled per column of x, its inverse scales (cols,)); a layout that is not wanted comes back empty."""
    C.require_cuda(x)
    rows, cols = x.shape
    dev = x.device
    r = torch.empty((rows, 2 * _cp64(cols)) if want_rows else (0,), dtype=torch.float16, device=dev)
    ri = torch.empty(rows if want_rows else 0, dtype=torch.float32, device=dev)
    t = torch.empty((cols, 2 * _cp64(rows)) if want_t else (0,), dtype=torch.float16, device=dev)
    ti = torch.empty(cols if want_t else 0, dtype=torch.float32, device=dev)
    if want_t:
        ws = _workspace(C.lib().hvae_split2h_workspace_bytes(rows, cols), dev)
        C.call("hvae_split2h_both_f32", C.ptr(x), C.ptr(r) if want_rows else None, C.ptr(ri) if want_rows else None, C.ptr(t),
               C.ptr(ti), rows, cols, C.ptr(ws), ws.numel(), C.stream())
        C.launch_count += 2
    elif want_rows:
        C.call("hvae_split2h_rows_f32", C.ptr(x), C.ptr(r), C.ptr(ri), rows, cols, C.stream())
        C.launch_count += 1
    return r, ri, t, ti


@split2h_both.register_fake
def _(x, want_rows, want_t):
    rows, cols = x.shape
    dev = x.device
    return (torch.empty((rows, 2 * _cp64(cols)) if want_rows else (0,), dtype=torch.float16, device=dev),
            torch.empty(rows if want_rows else 0, dtype=torch.float32, device=dev),
            torch.empty((cols, 2 * _cp64(rows)) if want_t else (0,), dtype=torch.float16, device=dev),
            torch.empty(cols if want_t else 0, dtype=torch.float32, device=dev))


@_op("hvae::split2h_both_ex", mutates_args=())
def split2h_both_ex(x: Tensor, mask: Optional[Tensor], want_rows: bool, want_t: bool, want_colsum: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """split2h_both of x with (optionally) the elements where mask <= 0 zeroed - the ReLU backward folded into the
    gradient's operand split - and (optionally) the column sums of that masked x (the bias gradient) from the same pass:
    -> (rows split, its scales, transposed split, its scales, colsum (cols,)); unwanted outputs come back empty."""
    C.require_cuda(x, mask)
    rows, cols = x.shape
    dev = x.device
    r = torch.empty((rows, 2 * _cp64(cols)) if want_rows else (0,), dtype=torch.float16, device=dev)
    ri = torch.empty(rows if want_rows else 0, dtype=torch.float32, device=dev)
    t = torch.empty((cols, 2 * _cp64(rows)) if want_t else (0,), dtype=torch.float16, device=dev)
    ti = torch.empty(cols if want_t else 0, dtype=torch.float32, device=dev)
    cs = torch.empty(cols if want_colsum else 0, dtype=torch.float32, device=dev)
    if want_rows or want_t or want_colsum:
        ws = _workspace(C.lib().hvae_split2h_ex_workspace_bytes(rows, cols), dev)
        C.call("hvae_split2h_both_ex_f32", C.ptr(x), C.ptr(mask), C.ptr(r) if want_rows else None, C.ptr(ri) if want_rows else None,
               C.ptr(t) if want_t else None, C.ptr(ti) if want_t else None, C.ptr(cs) if want_colsum else None, rows, cols,
               C.ptr(ws), ws.numel(), C.stream())
        general = want_t or want_colsum or mask is not None
        C.launch_count += ((1 + (1 if want_colsum else 0) + (1 if (want_rows or want_t) else 0)) if general else 1) - 1
    return r, ri, t, ti, cs


@split2h_both_ex.register_fake
def _(x, mask, want_rows, want_t, want_colsum):
    rows, cols = x.shape
    dev = x.device
    return (torch.empty((rows, 2 * _cp64(cols)) if want_rows else (0,), dtype=torch.float16, device=dev),
            torch.empty(rows if want_rows else 0, dtype=torch.float32, device=dev),
            torch.empty((cols, 2 * _cp64(rows)) if want_t else (0,), dtype=torch.float16, device=dev),
            torch.empty(cols if want_t else 0, dtype=torch.float32, device=dev),
            torch.empty(cols if want_colsum else 0, dtype=torch.float32, device=dev))


@_op("hvae::gemm_x2s", mutates_args=())
def gemm_x2s(As: Tensor, inv_a: Tensor, Bs: Tensor, inv_b: Tensor, bias: Optional[Tensor], relu: bool, M: int, N: int,
             K: int) -> Tensor:
    """C (M,N) = A (M,K) . B (N,K)^T (+ bias) (ReLU) on split2h operands (rows = output index) and their inverse scales."""
    C.require_cuda(inv_a, inv_b)
    if not (As.is_cuda and Bs.is_cuda and As.dtype == torch.float16 and Bs.dtype == torch.float16):
        raise RuntimeError("gemm_x2s: operands must be CUDA fp16 split2h buffers")
    if As.shape != (M, 2 * _cp64(K)) or Bs.shape != (N, 2 * _cp64(K)) or inv_a.numel() != M or inv_b.numel() != N:
        raise RuntimeError("gemm_x2s: operand shapes %s / %s do not match (M, N, K) = (%d, %d, %d)" % (tuple(As.shape), tuple(Bs.shape), M, N, K))
    out = torch.empty(M, N, dtype=torch.float32, device=As.device)
    ws = _workspace(C.lib().hvae_gemm_x2s_workspace_bytes(M, N), As.device)
    C.call("hvae_gemm_x2s_f32", C.ptr(As), C.ptr(inv_a), C.ptr(Bs), C.ptr(inv_b), C.ptr(bias), int(relu), C.ptr(out), M, N, K,
           C.ptr(ws), ws.numel(), C.stream())
    C.launch_count += C.lib().hvae_gemm_x2s_num_launches(M, N, K)
    return out


@gemm_x2s.register_fake
def _(As, inv_a, Bs, inv_b, bias, relu, M, N, K):
    return torch.empty(M, N, dtype=torch.float32, device=As.device)


@_op("hvae::linear_x2", mutates_args=())
def linear_x2_fwd(x: Tensor, weight: Tensor, bias: Optional[Tensor], need_gx: bool, need_gw: bool, relu: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (y, split of x^T + its scales, split of W^T + its scales): as linear_x3, on the fp16 two-piece path.
    relu: y = relu(x W^T + b) in the GEMM epilogue; the backward masks the upstream gradient by y > 0 inside its operand
    split (no separate activation kernels in either direction)."""
    xs, xi, xts, xti = split2h_both(x, True, need_gw)
    ws, wi, wts, wti = split2h_both(weight, True, need_gx)
    y = gemm_x2s(xs, xi, ws, wi, bias, relu, x.shape[0], weight.shape[0], x.shape[1])
    return y, xts, xti, wts, wti


@linear_x2_fwd.register_fake
def _(x, weight, bias, need_gx, need_gw, relu):
    M, K = x.shape
    N = weight.shape[0]
    dev = x.device
    return (x.new_empty(M, N),
            torch.empty((K, 2 * _cp64(M)) if need_gw else (0,), dtype=torch.float16, device=dev),
            torch.empty(K if need_gw else 0, dtype=torch.float32, device=dev),
            torch.empty((K, 2 * _cp64(N)) if need_gx else (0,), dtype=torch.float16, device=dev),
            torch.empty(K if need_gx else 0, dtype=torch.float32, device=dev))


def _lx2_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    x, weight, bias, need_gx, need_gw, relu = inputs
    ctx.save_for_backward(output[1], output[2], output[3], output[4], *([output[0]] if relu else []))
    ctx.dims = (x.shape[0], weight.shape[0], x.shape[1])  # rows, out, in
    ctx.has_bias = bias is not None
    ctx.need = (need_gx, need_gw)
    ctx.relu = relu


def _lx2_backward(ctx, gy, *_unused):
    if gy is None:
        return None, None, None, None, None, None
    xts, xti, wts, wti = ctx.saved_tensors[:4]
    y_act = ctx.saved_tensors[4] if ctx.relu else None
    M, n_out, n_in = ctx.dims
    need_gx, need_gw = ctx.need
    gy = _c(gy)
    gx = gw = gb = None
    want_gx = ctx.needs_input_grad[0] and need_gx
    want_gw = ctx.needs_input_grad[1] and need_gw
    want_gb = ctx.has_bias and ctx.needs_input_grad[2]
    if want_gx or want_gw or want_gb:
        # one pass for the row / column maxima, the ReLU mask and the bias gradient; one for the split(s)
        gs, gi, gts, gti, cs = split2h_both_ex(gy, y_act, want_gx, want_gw, want_gb)
        if want_gx:
            gx = gemm_x2s(gs, gi, wts, wti, None, False, M, n_in, n_out)      # gy (M,out) . W (out,in)
        if want_gw:
            gw = gemm_x2s(gts, gti, xts, xti, None, False, n_out, n_in, M)    # gy^T (out,M) . x (M,in)
        if want_gb:
            gb = cs
    return gx, gw, gb, None, None, None


linear_x2_fwd.register_autograd(_lx2_backward, setup_context=_lx2_setup)


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], relu: bool = False) -> Tensor:
    """torch.nn.functional.linear semantics (relu=True: followed by ReLU); the tensor-core fp32 path for GEMM-sized CUDA
    inputs - there the activation runs in the GEMM epilogue and its backward inside the gradient's operand split."""
    if trunk_x3_eligible(x, weight) and _trunk_mode == "x2":
        lead = x.shape[:-1]
        grad_on = torch.is_grad_enabled()
        y = linear_x2_fwd(_rows(x), _c(weight), None if bias is None else _c(bias), grad_on and x.requires_grad,
                          grad_on and weight.requires_grad, relu)[0]
        return y.view(*lead, weight.shape[0])
    if relu:
        return torch.relu(linear(x, weight, bias))
    if trunk_x3_eligible(x, weight):
        lead = x.shape[:-1]
        grad_on = torch.is_grad_enabled()
        y, _, _ = linear_x3_fwd(_rows(x), _c(weight), None if bias is None else _c(bias), grad_on and x.requires_grad,
                                grad_on and weight.requires_grad)
        return y.view(*lead, weight.shape[0])
    return torch.nn.functional.linear(x, weight, bias)


# ---------------------------------------------------------------------------------------------------
# Reconstruction-loss head: Bernoulli NLL with logits, one row kernel per direction
# ---------------------------------------------------------------------------------------------------
@_op("hvae::bce_logits_rows_fwd", mutates_args=())
def bce_logits_rows_fwd(logits: Tensor, x: Tensor) -> Tensor:
    """logits (S,B,N), x (B,N) -> nll (S,B) = sum_n BCE-with-logits (x broadcast over S)."""
    C.require_cuda(logits, x)
    S, B, N = logits.shape
    out = logits.new_empty(S, B)
    C.call("hvae_bce_logits_rows_fwd_f32", C.ptr(logits), C.ptr(x), C.ptr(out), S, B, N, C.stream())
    C.launch_count += 1
    return out


@bce_logits_rows_fwd.register_fake
def _(logits, x):
    return logits.new_empty(logits.shape[0], logits.shape[1])


@_op("hvae::bce_logits_rows_bwd", mutates_args=())
def bce_logits_rows_bwd(logits: Tensor, x: Tensor, gnll: Tensor) -> Tensor:
    C.require_cuda(logits, x, gnll)
    S, B, N = logits.shape
    out = torch.empty_like(logits)
    C.call("hvae_bce_logits_rows_bwd_f32", C.ptr(logits), C.ptr(x), C.ptr(gnll), C.ptr(out), S, B, N, C.stream())
    C.launch_count += 1
    return out


@bce_logits_rows_bwd.register_fake
def _(logits, x, gnll):
    return torch.empty_like(logits)


def _bce_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _bce_backward(ctx, g):
    logits, x = ctx.saved_tensors
    return bce_logits_rows_bwd(logits, x, _c(g)), None


bce_logits_rows_fwd.register_autograd(_bce_backward, setup_context=_bce_setup)


def bernoulli_nll_rows(logits: Tensor, x: Tensor) -> Tensor:
    """-Bernoulli(logits=logits).log_prob(x).sum(-1): logits (..., B, N) with x (B, N) broadcast over the leading dims.
    Targets carry no gradient (as in the reference objectives)."""
    lead = logits.shape[:-2]
    B, N = logits.shape[-2:]
    out = bce_logits_rows_fwd(_c(logits).view(-1, B, N), _c(x).view(B, N))
    return out.view(*lead, B)


# ---------------------------------------------------------------------------------------------------
# Generic reconstruction heads: MSE-sum and RelaxedBernoulli NLL, optionally with the decoder's final
# nn.Sigmoid fused; one row kernel per direction
# ---------------------------------------------------------------------------------------------------
RECON_MSE, RECON_SIGMOID_MSE, RECON_RB_LOGITS, RECON_RB_PROBS, RECON_RB_SIGMOID = 1, 2, 3, 4, 5


@_op("hvae::recon_rows_fwd", mutates_args=())
def recon_rows_fwd(inp: Tensor, x: Tensor, kind: int, temperature: float) -> Tensor:
    """inp (S,B,N), x (B,N) -> (S,B) row sums of the reconstruction term `kind` (x broadcast over S)."""
    C.require_cuda(inp, x)
    S, B, N = inp.shape
    out = inp.new_empty(S, B)
    C.call("hvae_recon_rows_fwd_f32", C.ptr(inp), C.ptr(x), C.ptr(out), S, B, N, kind, temperature, C.stream())
    return out


@recon_rows_fwd.register_fake
def _(inp, x, kind, temperature):
    return inp.new_empty(inp.shape[0], inp.shape[1])


@_op("hvae::recon_rows_bwd", mutates_args=())
def recon_rows_bwd(inp: Tensor, x: Tensor, gout: Tensor, kind: int, temperature: float) -> Tensor:
    C.require_cuda(inp, x, gout)
    S, B, N = inp.shape
    out = torch.empty_like(inp)
    C.call("hvae_recon_rows_bwd_f32", C.ptr(inp), C.ptr(x), C.ptr(gout), C.ptr(out), S, B, N, kind, temperature, C.stream())
    return out


@recon_rows_bwd.register_fake
def _(inp, x, gout, kind, temperature):
    return torch.empty_like(inp)


def _rr_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.kind, ctx.temperature = inputs[2], inputs[3]


def _rr_backward(ctx, g):
    inp, x = ctx.saved_tensors
    return recon_rows_bwd(inp, x, _c(g), ctx.kind, ctx.temperature), None, None, None


recon_rows_fwd.register_autograd(_rr_backward, setup_context=_rr_setup)


def recon_rows(inp: Tensor, x: Tensor, kind: int, temperature: float = 1.0) -> Tensor:
    """Per-row reconstruction loss: inp (..., B, N) against the target x (B, N) (no gradient to the target, as in the
    reference objectives) -> (..., B).  kind: RECON_MSE | RECON_SIGMOID_MSE | RECON_RB_LOGITS | RECON_RB_PROBS | RECON_RB_SIGMOID."""
    lead = inp.shape[:-2]
    B, N = inp.shape[-2:]
    out = recon_rows_fwd(_c(inp).view(-1, B, N), _c(x.detach()).view(B, N), int(kind), float(temperature))
    return out.view(*lead, B)


# ---------------------------------------------------------------------------------------------------
# loss tail of the pvae objective: (S,B) nll and KL terms -> {total, recon, kl} in one launch per direction
# ---------------------------------------------------------------------------------------------------
@_op("hvae::pvae_loss_fwd", mutates_args=())
def pvae_loss_fwd(nll: Tensor, kld: Tensor, beta: float) -> Tensor:
    C.require_cuda(nll, kld)
    S, B = nll.shape
    out = nll.new_empty(3)
    C.call("hvae_pvae_loss_fwd_f32", C.ptr(nll), C.ptr(kld), C.ptr(out), S, B, beta, C.stream())
    return out


@pvae_loss_fwd.register_fake
def _(nll, kld, beta):
    return nll.new_empty(3)


@_op("hvae::pvae_loss_bwd", mutates_args=())
def pvae_loss_bwd(gout: Tensor, S: int, B: int, beta: float) -> Tuple[Tensor, Tensor]:
    C.require_cuda(gout)
    gnll, gkld = gout.new_empty(S, B), gout.new_empty(S, B)
    C.call("hvae_pvae_loss_bwd_f32", C.ptr(gout), C.ptr(gnll), C.ptr(gkld), S, B, beta, C.stream())
    return gnll, gkld


@pvae_loss_bwd.register_fake
def _(gout, S, B, beta):
    return gout.new_empty(S, B), gout.new_empty(S, B)


def _pl_setup(ctx, inputs, output):
    ctx.S, ctx.B = inputs[0].shape
    ctx.beta = inputs[2]


def _pl_backward(ctx, g):
    gnll, gkld = pvae_loss_bwd(_c(g), ctx.S, ctx.B, ctx.beta)
    return gnll, gkld, None


pvae_loss_fwd.register_autograd(_pl_backward, setup_context=_pl_setup)


def pvae_loss(nll: Tensor, kld: Tensor, beta: float) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (loss_total, recon, kl) scalars: recon = sum_b mean_s nll, kl = sum_b mean_s kld, total = recon + beta kl."""
    out = pvae_loss_fwd(_c(nll), _c(kld), float(beta))
    return out[0], out[1], out[2]


# ---------------------------------------------------------------------------------------------------
# posterior scale head: clamp(softplus(h) + eps, lo, hi), one launch per direction
# ---------------------------------------------------------------------------------------------------
@_op("hvae::sigma_head_fwd", mutates_args=())
def sigma_head_fwd(h: Tensor, eps: float, lo: float, hi: float) -> Tensor:
    C.require_cuda(h)
    out = torch.empty_like(h)
    C.call("hvae_sigma_head_fwd_f32", C.ptr(h), C.ptr(out), h.numel(), eps, lo, hi, C.stream())
    return out


@sigma_head_fwd.register_fake
def _(h, eps, lo, hi):
    return torch.empty_like(h)


@_op("hvae::sigma_head_bwd", mutates_args=())
def sigma_head_bwd(h: Tensor, g: Tensor, eps: float, lo: float, hi: float) -> Tensor:
    C.require_cuda(h, g)
    gh = torch.empty_like(h)
    C.call("hvae_sigma_head_bwd_f32", C.ptr(h), C.ptr(g), C.ptr(gh), h.numel(), eps, lo, hi, C.stream())
    return gh


@sigma_head_bwd.register_fake
def _(h, g, eps, lo, hi):
    return torch.empty_like(h)


def _sh_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])
    ctx.args = inputs[1:]


def _sh_backward(ctx, g):
    (h,) = ctx.saved_tensors
    return sigma_head_bwd(h, _c(g), *ctx.args), None, None, None


sigma_head_fwd.register_autograd(_sh_backward, setup_context=_sh_setup)


def sigma_head(h: Tensor, eps: float = 1e-5, lo: float = 0.1, hi: float = 7.0) -> Tensor:
    """clamp(softplus(h) + eps, lo, hi): the pvae encoder's scale head and RiemannianNormal's clamp in one kernel."""
    return sigma_head_fwd(_c(h), float(eps), float(lo), float(hi))
