"""CPU check of the closed forms behind the lean pair paths of the SIMT gyroplane kernels (csrc/gyro_pair.cuh):
with e = |x-p|^2, q = <p,p-x>, qa = <a,p-x>, Bc = 1 - c|p|^2, den = 1 - 2c<p,x> + c^2|p|^2|x|^2,
    N1 = -Bc qa - c e <p,a>,      N2 = e (Bc^2 + 2 Bc c q + c^2 e |p|^2)
the reference's normdist2plane (hyperbolic_vae/manifolds.py:41-65 over geoopt's projected mobius_add) is
    unprojected pair :  asinh( 2 sqrt(c) N1 den / (|a| (den^2 - c N2)) ) / sqrt(c)
    projected pair   :  asinh( K N1 / sqrt(N2) ) / sqrt(c),   K = 2 sqrt(c) maxnorm / ((1 - c maxnorm^2) |a|)
and a pair is projected iff N2 > maxnorm^2 den^2.  Evaluated in float64 against the oracle's float64 evaluation of the
reference expression (fp32 semantics: the float32 projection radius).  The kernels themselves are tested on the GPU
(tests/test_gpu_layers.py); this pins the algebra without a device."""
import pytest
import torch

from oracle.geoopt_min.manifolds.stereographic import math as gmath
from oracle.pvae_min.manifolds import PoincareBall, normdist2plane


def _ball(D, c):
    b = PoincareBall(D, c=c)
    b.isp_c.data = torch.log(torch.expm1(torch.tensor(float(b.c), dtype=torch.float64)))  # the fp32 curvature value, in double
    return b


@pytest.mark.parametrize("D,c", [(2, 1.0), (10, 1.0), (16, 0.7), (5, 2.0)])
def test_lean_closed_forms_equal_normdist2plane(D, c):
    torch.manual_seed(D)
    B, P = 300, 40
    ball = _ball(D, c)
    cc = float(ball.c)
    sc = cc ** 0.5
    scale = torch.ones(B, 1, dtype=torch.float64)
    scale[: B // 3] = 6.0           # far out: the projected pairs
    scale[B // 3: 2 * B // 3] = 1.2
    scale[2 * B // 3:] = 0.2
    with gmath.fp32_semantics(True):
        x = ball.expmap0(torch.randn(B, D, dtype=torch.float64) / D ** 0.5 * scale)
        p = ball.expmap0(torch.randn(P, D, dtype=torch.float64) * 0.5 / D ** 0.5)
        a = torch.randn(P, D, dtype=torch.float64)
        ref = normdist2plane(ball, x.unsqueeze(1), a.unsqueeze(0), p.unsqueeze(0), signed=True, dim=-1)   # (B, P)
    with gmath.fp32_semantics(True):  # geoopt's float32 projection radius (eps 4e-3), exactly as the oracle applies it
        far = torch.zeros(1, D, dtype=torch.float64)
        far[0, 0] = 10.0 / sc
        maxnorm = float(gmath.project(far, k=torch.tensor(-cc, dtype=torch.float64), dim=-1).norm())
    assert abs(maxnorm * sc - 0.996) < 1e-6
    e = (x.unsqueeze(1) - p.unsqueeze(0)).pow(2).sum(-1)
    q = (p.unsqueeze(0) * (p.unsqueeze(0) - x.unsqueeze(1))).sum(-1)
    qa = (a.unsqueeze(0) * (p.unsqueeze(0) - x.unsqueeze(1))).sum(-1)
    p2, pa, an = p.pow(2).sum(-1), (p * a).sum(-1), a.norm(dim=-1)
    x2 = x.pow(2).sum(-1, keepdim=True)
    Bc = 1.0 - cc * p2
    den = 1.0 - 2.0 * cc * (p2 - q) + cc * cc * p2 * x2
    N1 = -Bc * qa - cc * e * pa
    N2 = e * (Bc * Bc + 2.0 * Bc * cc * q + cc * cc * e * p2)
    projected = N2 > maxnorm ** 2 * den ** 2
    assert 0.1 < float(projected.double().mean()) < 0.9     # both branches are exercised
    y_un = 2.0 * sc * N1 * den / (an * (den * den - cc * N2))
    K = 2.0 * sc * maxnorm / ((1.0 - cc * maxnorm ** 2) * an)
    y_pr = K * N1 / N2.sqrt()
    out = torch.asinh(torch.where(projected, y_pr, y_un)) / sc
    err = (out - ref).abs() / (1.0 + ref.abs())
    assert float(err.max()) < 1e-9, float(err.max())
    # and the projection test itself: |(-p)(+)x|^2 = N2 / den^2
    with gmath.fp32_semantics(False):
        diff = gmath.mobius_add(-p.unsqueeze(0), x.unsqueeze(1), k=torch.tensor(-cc, dtype=torch.float64), dim=-1)
    assert float(((diff.pow(2).sum(-1) - N2 / den ** 2).abs() / (N2 / den ** 2 + 1e-30)).max()) < 1e-8
