"""ORACLE (test infrastructure, never shipped): CPU restatement of the Poincare-ball arithmetic
of geoopt's ``manifolds/stereographic/math.py`` (github.com/geoopt/geoopt, git rev a41c09b7,
pinned by /root/reference/pyproject.toml:27).  geoopt is NOT vendored under /root/reference and is
not installable here, so this file restates its published algorithm for
the negative-curvature branch (k = -c < 0), which is the only branch the reference reaches through
``geoopt.PoincareBall``.

PARITY STATUS: the third-party arithmetic restated here is UNPINNED (no copy of geoopt exists to
diff against); the reference-owned layer above it IS pinned, because the reference's own files are
executed verbatim over this shim when the golden fixtures are minted (tests/golden/make_golden.py).

Reference call sites this serves: hyperbolic_vae/layers.py:60,67,76,130,146,200,220;
distributions/wrapped_normal.py:49,69-73,79-85; manifolds.py:31,54,62.
"""
from __future__ import annotations

import torch

import contextlib

MIN_NORM = 1e-15
BALL_EPS = {torch.float32: 4e-3, torch.float64: 1e-5}

# "float64 arithmetic, float32 semantics": when the float64 evaluation is used as the tie-breaker truth
# for a float32 result, the dtype-dependent CONSTANTS must stay those of the float32 function
# (projection eps 4e-3, artanh clamp float32(1-1e-7) = 0.99999988), otherwise it is a different function.
_FP32_SEMANTICS = False
_ARTANH_CLAMP_F32 = float(torch.tensor(1 - 1e-7, dtype=torch.float32))


@contextlib.contextmanager
def fp32_semantics(on: bool = True):
    global _FP32_SEMANTICS
    old, _FP32_SEMANTICS = _FP32_SEMANTICS, on
    try:
        yield
    finally:
        _FP32_SEMANTICS = old


# ---- scalar helpers -------------------------------------------------------------------------
def tanh(x: torch.Tensor) -> torch.Tensor:
    return x.clamp(-15, 15).tanh()


def artanh(x: torch.Tensor) -> torch.Tensor:
    if _FP32_SEMANTICS:
        x = x.clamp(-_ARTANH_CLAMP_F32, _ARTANH_CLAMP_F32)
    else:
        x = x.clamp(-1 + 1e-7, 1 - 1e-7)
    return (torch.log(1 + x) - torch.log(1 - x)) * 0.5


def arsinh(x: torch.Tensor) -> torch.Tensor:
    return (x + torch.sqrt(1 + x.pow(2))).clamp_min(MIN_NORM).log().to(x.dtype)


def sign(x: torch.Tensor) -> torch.Tensor:
    # sign(0) == +1
    return torch.sign(x.sign() + 0.5)


def sabs(x: torch.Tensor, eps: float = MIN_NORM) -> torch.Tensor:
    return x.abs() + eps


def clamp_abs(x: torch.Tensor, eps: float = MIN_NORM) -> torch.Tensor:
    return sign(x) * sabs(x, eps=eps)


def _as_k(k, like: torch.Tensor) -> torch.Tensor:
    if not torch.is_tensor(k):
        k = torch.as_tensor(k, dtype=like.dtype)
    return k


def _neg_only(k: torch.Tensor):
    if not bool(torch.all(k < 0)):
        raise NotImplementedError("geoopt_min restates only the k<0 (Poincare ball) branch")


def tan_k(x: torch.Tensor, k) -> torch.Tensor:
    k = _as_k(k, x)
    _neg_only(k)
    k_sqrt = sabs(k).sqrt()
    return k_sqrt.reciprocal() * tanh(x * k_sqrt)


def artan_k(x: torch.Tensor, k) -> torch.Tensor:
    k = _as_k(k, x)
    _neg_only(k)
    k_sqrt = sabs(k).sqrt()
    return k_sqrt.reciprocal() * artanh(x * k_sqrt)


def arsin_k(x: torch.Tensor, k) -> torch.Tensor:
    k = _as_k(k, x)
    _neg_only(k)
    k_sqrt = sabs(k).sqrt()
    return k_sqrt.reciprocal() * arsinh(x * k_sqrt)


# ---- ball operations ------------------------------------------------------------------------
def project(x: torch.Tensor, *, k, dim: int = -1, eps: float = -1.0) -> torch.Tensor:
    k = _as_k(k, x)
    if eps < 0:
        eps = 4e-3 if (x.dtype == torch.float32 or _FP32_SEMANTICS) else 1e-5
    if _FP32_SEMANTICS:
        eps = 1.0 - float(torch.tensor(1 - eps, dtype=torch.float32))  # the fp32 path divides float32(0.996)
    maxnorm = (1 - eps) / (sabs(k) ** 0.5)
    maxnorm = torch.where(k.lt(0), maxnorm, k.new_full((), 1e15))
    norm = x.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
    cond = norm > maxnorm
    projected = x / norm * maxnorm
    return torch.where(cond, projected, x)


def lambda_x(x: torch.Tensor, *, k, keepdim: bool = False, dim: int = -1) -> torch.Tensor:
    k = _as_k(k, x)
    return 2 / (1 + k * x.pow(2).sum(dim=dim, keepdim=keepdim)).clamp_min(MIN_NORM)


def mobius_add(x: torch.Tensor, y: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    k = _as_k(k, x)
    x2 = x.pow(2).sum(dim=dim, keepdim=True)
    y2 = y.pow(2).sum(dim=dim, keepdim=True)
    xy = (x * y).sum(dim=dim, keepdim=True)
    num = (1 - 2 * k * xy - k * y2) * x + (1 + k * x2) * y
    denom = 1 - 2 * k * xy + k**2 * x2 * y2
    return num / denom.clamp_min(MIN_NORM)


def expmap0(u: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    u_norm = u.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return tan_k(u_norm, k) * (u / u_norm)


def logmap0(y: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    y_norm = y.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return (y / y_norm) * artan_k(y_norm, k)


def expmap(x: torch.Tensor, u: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    u_norm = u.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    lam = lambda_x(x, k=k, dim=dim, keepdim=True)
    second_term = tan_k((lam / 2.0) * u_norm, k) * (u / u_norm)
    return mobius_add(x, second_term, k=k, dim=dim)


def logmap(x: torch.Tensor, y: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    sub = mobius_add(-x, y, k=k, dim=dim)
    sub_norm = sub.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    lam = lambda_x(x, k=k, keepdim=True, dim=dim)
    return 2.0 * artan_k(sub_norm, k) * (sub / (lam * sub_norm))


def dist(x: torch.Tensor, y: torch.Tensor, *, k, keepdim: bool = False, dim: int = -1) -> torch.Tensor:
    return 2.0 * artan_k(mobius_add(-x, y, k=k, dim=dim).norm(dim=dim, p=2, keepdim=keepdim), k)


def gyration(u: torch.Tensor, v: torch.Tensor, w: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    k = _as_k(k, u)
    u2 = u.pow(2).sum(dim=dim, keepdim=True)
    v2 = v.pow(2).sum(dim=dim, keepdim=True)
    uv = (u * v).sum(dim=dim, keepdim=True)
    uw = (u * w).sum(dim=dim, keepdim=True)
    vw = (v * w).sum(dim=dim, keepdim=True)
    k2 = k**2
    a = -k2 * uw * v2 - k * vw + 2 * k2 * uv * vw
    b = -k2 * vw * u2 + k * uw
    d = 1 - 2 * k * uv + k2 * u2 * v2
    return w + 2 * (a * u + b * v) / d.clamp_min(MIN_NORM)


def parallel_transport(x, y, v, *, k, dim: int = -1) -> torch.Tensor:
    return (
        gyration(y, -x, v, k=k, dim=dim)
        * lambda_x(x, k=k, keepdim=True, dim=dim)
        / lambda_x(y, k=k, keepdim=True, dim=dim)
    )


def parallel_transport0(y: torch.Tensor, v: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    k = _as_k(k, y)
    return v * (1 + k * y.pow(2).sum(dim=dim, keepdim=True)).clamp_min(MIN_NORM)


def mobius_matvec(m: torch.Tensor, x: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    if m.dim() > 2 and dim != -1:
        raise RuntimeError("broadcasted Mobius matvec is supported for the last dim only")
    x_norm = x.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
    if dim != -1 or m.dim() == 2:
        mx = torch.tensordot(x, m, ([dim], [1]))
    else:
        mx = torch.matmul(m, x.unsqueeze(-1)).squeeze(-1)
    mx_norm = mx.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
    res_c = tan_k(mx_norm / x_norm * artan_k(x_norm, k), k) * (mx / mx_norm)
    cond = (mx == 0).prod(dim=dim, keepdim=True, dtype=torch.bool)
    res_0 = torch.zeros(1, dtype=res_c.dtype, device=res_c.device)
    return torch.where(cond, res_0, res_c)


def dist2plane(x, p, a, *, k, keepdim: bool = False, signed: bool = False, scaled: bool = False, dim: int = -1):
    k = _as_k(k, x)
    diff = mobius_add(-p, x, k=k, dim=dim)
    diff_norm2 = diff.pow(2).sum(dim=dim, keepdim=keepdim).clamp_min(MIN_NORM)
    sc_diff_a = (diff * a).sum(dim=dim, keepdim=keepdim)
    if not signed:
        sc_diff_a = sc_diff_a.abs()
    a_norm = a.norm(dim=dim, keepdim=keepdim, p=2)
    num = 2.0 * sc_diff_a
    denom = clamp_abs((1 + k * diff_norm2) * a_norm)
    distance = arsin_k(num / denom, k)
    if scaled:
        distance = distance * a_norm
    return distance


def egrad2rgrad(x: torch.Tensor, grad: torch.Tensor, *, k, dim: int = -1) -> torch.Tensor:
    return grad / lambda_x(x, k=k, keepdim=True, dim=dim) ** 2


def inner(x, u, v, *, k, keepdim: bool = False, dim: int = -1) -> torch.Tensor:
    res = lambda_x(x, k=k, keepdim=True, dim=dim) ** 2 * (u * v).sum(dim=dim, keepdim=True)
    return res if keepdim else res.squeeze(dim)
