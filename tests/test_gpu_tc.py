"""GPU parity of the tcgen05 (bf16-operand) forward paths against the fp32 oracle: BASELINE.json's
"1e-2 in bf16 GEMM mode".  Inputs are kept away from the projection radius (conditioning, see util_parity)."""
import pytest
import torch

from util_parity import pair_kappa

pytestmark = pytest.mark.gpu


def _oball(c):
    from oracle.geoopt_min import PoincareBall

    return PoincareBall(c=c)


# the last two shapes have >= 148 m-blocks, K <= 512 and >= 4 n-tiles: they run the A-resident schedule - the forward-only
# variant as ONE kernel (Gram tiles + row factor + scaled output tiles per m-block) - with ragged M, N and K
@pytest.mark.parametrize("B,F,P", [(128, 64, 128), (256, 128, 256), (384, 256, 640), (1000, 512, 300), (4096, 512, 1024), (19000, 512, 600),
                                   (37900, 512, 1000), (38100, 200, 1024)])
def test_mobius_tc_forward(B, F, P):
    import hvae
    from hvae import ops
    from oracle.geoopt_min.manifolds.stereographic import math as gm

    torch.manual_seed(B + P)
    c = 1.0
    ob = _oball(c)
    x = ob.expmap0(torch.randn(B, F) * 0.5 / F ** 0.5).detach()
    M = torch.randn(P, F) / F ** 0.5 * 0.7
    ref = gm.project(gm.mobius_matvec(M.double(), x.double(), k=torch.tensor(-c, dtype=torch.float64)), k=torch.tensor(-c, dtype=torch.float64), eps=4e-3)
    ops.set_gemm_mode("bf16")
    try:
        y, mx = ops.mobius_matvec_tc_fwd(x.cuda(), M.cuda(), hvae.PoincareBall(c).c_value)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_mode("fp32")
    mx_ref = x.double() @ M.double().t()
    err_mx = (mx.double().cpu() - mx_ref).abs().max() / mx_ref.abs().max()
    assert err_mx < 1e-2, err_mx
    scale = ref.abs().amax(dim=-1, keepdim=True)
    err = ((y.double().cpu() - ref).abs() / scale).max()
    assert err < 1e-2, err
    # the fp32 SIMT path on the same inputs agrees with the TC path to bf16 accuracy as well
    y32, _ = ops.mobius_matvec_fwd(x.cuda(), M.cuda(), hvae.PoincareBall(c).c_value)
    assert ((y32 - y).abs().cpu() / scale.float()).max() < 1e-2
    # forward-only single-pass variant (Gram-matrix row scale fused into the GEMM epilogue)
    if P % 8 == 0:
        yi, mxsq = ops.mobius_matvec_tc(x.cuda(), M.cuda(), hvae.PoincareBall(c).c_value)
        torch.cuda.synchronize()
        erri = ((yi.double().cpu() - ref).abs() / scale).max()
        assert erri < 1e-2, erri
        sq_ref = mx_ref.pow(2).sum(-1)
        assert ((mxsq.double().cpu() - sq_ref).abs() / sq_ref).max() < 2e-2


# tensor-core backward (row pass + two GEMMs) against float64 autograd of the oracle; includes rows clipped by the
# projection (scale 3.0) and ragged tiles in every dimension
@pytest.mark.parametrize("B,F,P,scale", [(128, 64, 128, 0.5), (256, 128, 256, 0.5), (1000, 512, 296, 0.5), (4096, 256, 1024, 0.5),
                                          (512, 128, 384, 8.0), (19000, 512, 600, 0.5)])
def test_mobius_tc_backward(B, F, P, scale):
    import hvae
    from hvae import ops
    from oracle.geoopt_min.manifolds.stereographic import math as gm

    torch.manual_seed(B + 7 * P)
    c = 1.0
    ob = _oball(c)
    x = ob.expmap0(torch.randn(B, F) * 0.5 / F ** 0.5).detach()
    M = torch.randn(P, F) / F ** 0.5 * 0.7 * scale
    gy = torch.randn(B, P)
    k = torch.tensor(-c, dtype=torch.float64)
    xd, Md = x.double().requires_grad_(True), M.double().requires_grad_(True)
    with gm.fp32_semantics():
        ref = gm.project(gm.mobius_matvec(Md, xd, k=k), k=k)
    ref.backward(gy.double())
    ops.set_gemm_mode("bf16")
    try:
        xc, Mc = x.cuda().requires_grad_(True), M.cuda().requires_grad_(True)
        y = ops.mobius_matvec(xc, Mc, hvae.PoincareBall(c).c_value)
        y.backward(gy.cuda())
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_mode("fp32")
    if scale > 1.0:
        assert float((ref.norm(dim=-1) > 0.995).float().mean()) > 0.2  # the clipped branch is exercised
    # bf16 operands: errors are relative to the row / matrix scale of each gradient (sums over P resp. B of rounded terms)
    gx_ref, gM_ref = xd.grad, Md.grad
    ex = (xc.grad.double().cpu() - gx_ref).norm(dim=-1) / gx_ref.norm(dim=-1).clamp_min(1e-12)
    assert float(ex.max()) < 3e-2, float(ex.max())
    assert float(ex.mean()) < 1e-2, float(ex.mean())
    eM = (Mc.grad.double().cpu() - gM_ref).norm() / gM_ref.norm()
    assert float(eM) < 1e-2, float(eM)
    eMr = (Mc.grad.double().cpu() - gM_ref).norm(dim=-1) / gM_ref.norm(dim=-1).clamp_min(1e-12)
    assert float(eMr.max()) < 3e-2, float(eMr.max())


@pytest.mark.parametrize("B,D,P", [(128, 64, 128), (512, 128, 384), (300, 256, 200), (2048, 512, 1024), (19000, 256, 520),
                                   (37900, 256, 1000)])
def test_gyroplane_tc_forward(B, D, P):
    import hvae
    from hvae import ops
    from oracle.geoopt_min.manifolds.stereographic import math as gm

    torch.manual_seed(B + D)
    c = 1.0
    ob = _oball(c)
    x = ob.expmap0(torch.randn(B, D) * 0.6 / D ** 0.5).detach()
    p = ob.expmap0(torch.randn(P, D) * 0.6 / D ** 0.5).detach()
    bias = torch.randn(P)
    k = torch.tensor(-c, dtype=torch.float64)
    # (the oracle broadcasts to (rows, D, P) float64: in row chunks)
    ref = torch.cat([gm.dist2plane(xc.double().unsqueeze(-1), p.double().t(), p.double().t(), k=k, signed=True, dim=-2)
                     for xc in x.split(4096)]) + bias.double()
    out = ops.gyroplane_tc_fwd(x.cuda(), p.cuda(), bias.cuda(), hvae.PoincareBall(c).c_value, ops.GYRO_SIGNED)
    torch.cuda.synchronize()
    pk = pair_kappa(c, x, p)
    err = (out.double().cpu() - ref).abs()
    bound = 1e-2 * ref.abs() + 1e-2 * pk  # bf16 operands: eps ~ 4e-3 on <x,p>, amplified by the pair conditioning
    assert bool((err <= bound).all()), (err / bound).max()
    assert float(err.max()) < 0.05 * float(ref.abs().max()) + 1e-2


def test_layer_dispatches_to_tc_in_bf16_mode():
    import hvae
    from hvae import _cabi, layers, ops

    ball = hvae.PoincareBall(1.0)
    lay = layers.MobiusLayer(256, 512, ball).cuda()
    x = torch.randn(256, 256, device="cuda") * 0.05
    y32 = lay(x)
    ops.set_gemm_mode("bf16")
    try:
        y16 = lay(x)
    finally:
        ops.set_gemm_mode("fp32")
    assert (y16 - y32).abs().max() < 1e-2 * y32.abs().max() + 1e-4


# gyroplane backward on the tensor cores (recompute GEMM + pair-gradient tile kernel + two GEMMs) against float64
# autograd of the oracle; bf16 operands -> errors relative to each gradient's row / tensor scale
@pytest.mark.parametrize("B,D,P,with_bias", [(256, 128, 256, True), (1024, 512, 384, False), (2048, 256, 1000, True), (4096, 64, 640, False)])
def test_gyroplane_tc_backward(B, D, P, with_bias):
    import hvae
    from hvae import ops
    from oracle.geoopt_min.manifolds.stereographic import math as gm

    torch.manual_seed(B + D + P)
    c = 1.0
    ob = _oball(c)
    x = ob.expmap0(torch.randn(B, D) * 0.6 / D ** 0.5).detach()
    p = ob.expmap0(torch.randn(P, D) * 0.6 / D ** 0.5).detach()
    bias = torch.randn(P) if with_bias else None
    g = torch.randn(B, P)
    k = torch.tensor(-c, dtype=torch.float64)
    xd, pd = x.double().requires_grad_(True), p.double().requires_grad_(True)
    bd = bias.double().requires_grad_(True) if with_bias else None
    with gm.fp32_semantics():
        ref = gm.dist2plane(xd.unsqueeze(-1), pd.t(), pd.t(), k=k, signed=True, dim=-2)
    if with_bias:
        ref = ref + bd
    ref.backward(g.double())
    xc, pc = x.cuda().requires_grad_(True), p.cuda().requires_grad_(True)
    bc = bias.cuda().requires_grad_(True) if with_bias else None
    out = ops.gyroplane_tc_fwd(xc, pc, bc, hvae.PoincareBall(c).c_value, ops.GYRO_SIGNED)
    out.backward(g.cuda())
    torch.cuda.synchronize()
    ex = (xc.grad.double().cpu() - xd.grad).norm(dim=-1) / xd.grad.norm(dim=-1).clamp_min(1e-12)
    assert float(ex.max()) < 3e-2 and float(ex.mean()) < 1e-2, (float(ex.max()), float(ex.mean()))
    ep = (pc.grad.double().cpu() - pd.grad).norm(dim=-1) / pd.grad.norm(dim=-1).clamp_min(1e-12)
    assert float(ep.max()) < 3e-2 and float(ep.mean()) < 1e-2, (float(ep.max()), float(ep.mean()))
    if with_bias:
        assert float((bc.grad.double().cpu() - bd.grad).abs().max() / bd.grad.abs().max()) < 1e-5


def test_layers_train_in_bf16_mode():
    """MobiusLayer -> expmap0 -> gyroplane decoder (geoopt's Distance2PoincareHyperplanes, a == p) at tensor-core sizes: forward AND backward run on the tcgen05 paths in
    bf16 mode and agree with the fp32 SIMT/fp32 paths to bf16 accuracy (gradients relative to their own norm)."""
    import hvae
    from hvae import _cabi, layers, ops

    torch.manual_seed(3)
    ball = hvae.PoincareBall(1.0)
    enc = layers.MobiusLayer(512, 64, ball).cuda()
    dec = layers.Distance2PoincareHyperplanes(64, 256, ball=ball, std=0.3).cuda()
    x = (torch.randn(1024, 512, device="cuda") * 0.03).requires_grad_(True)
    g = torch.randn(1024, 256, device="cuda")
    res = {}
    for mode in ("fp32", "bf16"):
        ops.set_gemm_mode(mode)
        try:
            for m in (enc, dec):
                m.zero_grad(set_to_none=True)
            x.grad = None
            out = dec(ball.expmap0(enc(x)))
            out.backward(g)
            torch.cuda.synchronize()
        finally:
            ops.set_gemm_mode("fp32")
        res[mode] = {"out": out.detach().clone(), "x": x.grad.clone()}
        for tag, m in (("enc", enc), ("dec", dec)):
            for n, p in m.named_parameters():
                if p.grad is not None:
                    res[mode][tag + "." + n] = p.grad.clone()
    assert set(res["bf16"]) == set(res["fp32"]) and "enc._weight" in res["fp32"] and "dec.points" in res["fp32"]
    for n, b in res["fp32"].items():
        a = res["bf16"][n]
        assert float((a - b).norm() / b.norm().clamp_min(1e-30)) < 2e-2, n


# GeodesicLayer (a != p; pvae clamps + projected mobius_add) on the tensor cores: one N-concatenated cta_group::2 GEMM
# (<x,p>, <x,a>) + the pair epilogue, against the float64 oracle of the reference's layer (layers.py:96-121).
@pytest.mark.parametrize("B,D,P,wn", [(1024, 64, 128, False), (4096, 512, 4096, False), (2000, 256, 520, True), (1280, 128, 300, False)])
def test_geodesic_tc_forward(B, D, P, wn):
    import hvae
    from hvae import layers as HL
    from hvae import ops
    from oracle import ref_port as R

    torch.manual_seed(B + D + P)
    c = 1.0
    lay = HL.GeodesicLayer(D, P, hvae.PoincareBall(c), weight_norm=wn)
    ob64 = _oball(c).double()
    ref_l = R.GeodesicLayer(D, P, ob64, weight_norm=wn).double()
    with torch.no_grad():
        ref_l._weight.copy_(lay._weight.double())
        ref_l._bias.copy_(lay._bias.double())
    x = _oball(c).expmap0(torch.randn(B, D) * 0.6 / D ** 0.5).detach()
    with torch.no_grad():
        ref = ref_l(x.double())
        pw = ref_l.weight.float()
    lay = lay.cuda()
    ops.set_gemm_mode("bf16")
    try:
        with torch.no_grad():
            n0 = hvae._cabi.launch_count
            out = lay(x.cuda())
            torch.cuda.synchronize()
    finally:
        ops.set_gemm_mode("fp32")
    pk = pair_kappa(c, x, pw)
    err = (out.double().cpu() - ref).abs()
    bound = 1e-2 * ref.abs() + 1e-2 * pk
    assert bool((err <= bound).all()), float((err / bound).max())
    assert float(err.max()) < 0.05 * float(ref.abs().max()) + 1e-2
    assert hvae._cabi.launch_count - n0 >= 3   # weight prep + the tensor-core entry (not the SIMT gyroplane)
