"""GPU: the rest of the drop-in API surface against the oracle — manifold helper methods, the free
functions logdetexp / normdist2plane, over-parameterised Riemannian layers, sample-dim broadcasting, error behaviour."""
import pytest
import torch

from util_parity import assert_parity, kappa, rtol_grad, rtol_val

pytestmark = pytest.mark.gpu


def _balls(c):
    import hvae
    from oracle.geoopt_min import PoincareBall as OBall

    return hvae.PoincareBall(c), OBall(c)


@pytest.mark.parametrize("c", [0.5, 1.0, 1.4])
def test_manifold_helper_methods(c):
    hb, ob = _balls(c)
    assert hb.c_value == float(ob.c)
    torch.manual_seed(1)
    x = ob.expmap0(torch.randn(40, 6) * 0.4).detach()
    y = ob.expmap0(torch.randn(40, 6) * 0.5).detach()
    v = torch.randn(40, 6)
    xc, yc, vc = x.cuda(), y.cuda(), v.cuda()
    for name, a, b in (
        ("lambda_x", hb.lambda_x(xc, keepdim=True), ob.lambda_x(x, keepdim=True)),
        ("transp0", hb.transp0(yc, vc), ob.transp0(y, v)),
        ("transp", hb.transp(xc, yc, vc), ob.transp(x, y, v)),
        ("projx", hb.projx(xc * 3), ob.projx(x * 3)),
        ("egrad2rgrad", hb.egrad2rgrad(xc, vc), ob.egrad2rgrad(x, v)),
        ("inner", hb.inner(xc, vc, vc.flip(0)), ob.inner(x, v, v.flip(0))),
        ("retr", hb.retr(xc, vc * 0.1), ob.retr(x, v * 0.1)),
        ("dist", hb.dist(xc, yc, keepdim=True), ob.dist(x, y, keepdim=True)),
    ):
        torch.testing.assert_close(a.cpu(), torch.Tensor(b), rtol=2e-5, atol=1e-6, msg=name)
    assert hb.origin(6, device="cuda").shape == (6,) and float(hb.origin(3, 4).abs().sum()) == 0.0
    assert hb.check_point_on_manifold(xc) and not hb.check_point_on_manifold(xc * 10)
    with pytest.raises(ValueError):
        hb.assert_check_point_on_manifold(xc * 10)


def test_free_functions_logdetexp_and_normdist2plane():
    from hvae import manifolds as HM
    from oracle import ref_port as R

    hb, ob = _balls(1.0)
    torch.manual_seed(2)
    x = ob.expmap0(torch.randn(30, 5) * 0.4).detach()
    y = ob.expmap0(torch.randn(30, 5) * 0.5).detach()
    a = HM.logdetexp(hb, x.cuda(), y.cuda(), keepdim=True)
    b = R.logdetexp(ob, x.double(), y.double(), keepdim=True)
    torch.testing.assert_close(a.cpu().double(), b, rtol=2e-5, atol=2e-6)
    # normdist2plane through the expanded layout GeodesicLayer uses (layers.py:98-121)
    P = 7
    pa = ob.expmap0(torch.randn(P, 5) * 0.3).detach()
    pp = torch.randn(P, 5) * 0.4
    xe = x.cuda().unsqueeze(-2).expand(30, P, 5)
    out = HM.normdist2plane(hb, xe, pa.cuda(), pp.cuda(), signed=True, norm=True)
    ref = R.normdist2plane(ob, x.unsqueeze(-2).expand(30, P, 5), pa, pp, signed=True, norm=True)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=5e-6)  # fp32 vs fp32, ill-conditioned pairs included
    with pytest.raises(NotImplementedError):
        HM.normdist2plane(hb, x.cuda(), pa.cuda()[0:1].expand(30, 5), pp.cuda()[0:1].expand(30, 5))


@pytest.mark.parametrize("kind", ["mobius", "geodesic"])
def test_over_param_layers(kind):
    import hvae
    from hvae import layers as HL
    from oracle import ref_port as R
    from oracle.geoopt_min import PoincareBall as OBall

    torch.manual_seed(3)
    c, Fin, Pout, B = 1.0, 12, 9, 50
    o = (R.MobiusLayer if kind == "mobius" else R.GeodesicLayer)(Fin, Pout, OBall(c), over_param=True)
    h = (HL.MobiusLayer if kind == "mobius" else HL.GeodesicLayer)(Fin, Pout, hvae.PoincareBall(c), over_param=True)
    assert h._bias.shape == (Pout, Fin)
    with torch.no_grad():
        h._weight.copy_(o._weight)
        h._bias.copy_(torch.Tensor(o._bias.detach()))
    h = h.cuda()
    ob = OBall(c)
    x = (ob.expmap0(torch.randn(B, Fin) * 0.4).detach() if kind == "geodesic" else torch.randn(B, Fin))
    g = torch.randn(B, Pout)
    xo = x.clone().requires_grad_(True)
    yo = o(xo)
    yo.backward(g)
    xh = x.cuda().requires_grad_(True)
    yh = h(xh)
    yh.backward(g.cuda())
    torch.testing.assert_close(yh.cpu(), yo.detach(), rtol=3e-5, atol=3e-6)
    # fp32 against fp32 (the CPU reference carries its own rounding): 1e-4 relative plus 1e-5 of the tensor's scale
    for got, ref in ((xh.grad.cpu(), xo.grad), (h._weight.grad.cpu(), o._weight.grad), (h._bias.grad.cpu(), torch.Tensor(o._bias.grad))):
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5 * max(1.0, float(ref.abs().max())) + 2e-5)


def test_wrapped_normal_sample_dims_and_softplus():
    import hvae
    from hvae.distributions import WrappedNormal
    from oracle import ref_port as R
    from oracle.geoopt_min import PoincareBall as OBall

    torch.manual_seed(4)
    c, B, D, S = 0.7, 20, 4, 3
    ob = OBall(c)
    mu = ob.expmap0(torch.randn(B, D) * 0.4).detach()
    raw = torch.randn(B, D)
    eps = torch.randn(S, B, D)
    q_o = R.WrappedNormal(mu, raw, ob, softplus=True)
    z_o = q_o.rsample(torch.Size([S]), eps=eps)
    q_h = WrappedNormal(mu.cuda(), raw.cuda(), hvae.PoincareBall(c), softplus=True)
    z_h = q_h.rsample(torch.Size([S]), eps=eps.cuda())
    assert z_h.shape == (S, B, D)
    torch.testing.assert_close(z_h.cpu(), z_o, rtol=3e-5, atol=3e-6)
    lp_h = q_h.log_prob(z_h)
    assert lp_h.shape == (S, B, 1)
    torch.testing.assert_close(lp_h.cpu(), q_o.log_prob(z_o), rtol=3e-5, atol=3e-5)
    assert q_h.rsample().shape == (B, D) and q_h.sample(torch.Size([2])).shape == (2, B, D)
    assert q_h.batch_shape == (B,) and q_h.event_shape == (D,) and q_h.mean is q_h.loc
    with pytest.raises(ValueError):
        WrappedNormal((mu * 50).cuda(), raw.cuda(), hvae.PoincareBall(c), validate_args=True)


def test_unknown_shapes_fail_loudly():
    import hvae
    from hvae import ops

    ball = hvae.PoincareBall(1.0)
    with pytest.raises(RuntimeError, match="shape"):
        ops.expmap0(torch.randn(4, 2000, device="cuda"), ball.c_value)  # D > 1024 is not supported by the row kernels
    with pytest.raises(RuntimeError, match="float32"):
        ball.expmap0(torch.randn(4, 2, device="cuda", dtype=torch.float64))
    with pytest.raises(NotImplementedError):
        ball.expmap0(torch.randn(4, 2, device="cuda"), dim=0)


@pytest.mark.parametrize("S,B,N", [(1, 33, 784), (2, 17, 101), (1, 8, 20000)])
def test_recon_heads_match_torch(S, B, N):
    """MSE-sum and RelaxedBernoulli heads (with and without the fused decoder Sigmoid) against float64 torch:
    F.mse_loss / torch.distributions.RelaxedBernoulli (models/vae_hyperbolic.py:218-225, ...gyroplane_decoder.py:121-122)."""
    from hvae import ops

    torch.manual_seed(S * 100 + N)
    l = (torch.randn(S, B, N) * 3).requires_grad_(True)
    l.data[0, 0, :4] = torch.tensor([40.0, -40.0, 17.0, -17.0])   # saturated sigmoid: probs clamp binds
    x = torch.rand(B, N)
    x[0, :3] = torch.tensor([0.0, 1.0, 1e-30])                    # value clamp binds
    g = torch.randn(S, B)

    def ref(kind, T):
        l64 = l.detach().double().requires_grad_(True)
        x64 = x.double().expand(S, B, N)
        if kind == ops.RECON_MSE:
            out = (l64 - x64).pow(2).sum(-1)
        elif kind == ops.RECON_SIGMOID_MSE:
            out = (torch.sigmoid(l64) - x64).pow(2).sum(-1)
        else:
            eps32, tiny32 = torch.finfo(torch.float32).eps, torch.finfo(torch.float32).tiny
            if kind == ops.RECON_RB_LOGITS:
                lg = l64
            else:
                p = torch.sigmoid(l64) if kind == ops.RECON_RB_SIGMOID else l64
                ps = p.clamp(eps32, 1 - eps32)
                lg = ps.log() - (-ps).log1p()
            v = x64.clamp(tiny32, 1 - eps32)
            y = v.log() - (-v).log1p()
            d = lg - y * T
            sp = torch.nn.functional.softplus
            out = (-torch.log(torch.tensor(T, dtype=torch.float64)) - d + 2 * sp(d) - sp(-y) - sp(y)).sum(-1)
        out.backward(g.double())
        return out.detach(), l64.grad

    for kind, T in ((ops.RECON_MSE, 1.0), (ops.RECON_SIGMOID_MSE, 1.0), (ops.RECON_RB_LOGITS, 0.1), (ops.RECON_RB_SIGMOID, 1.0),
                    (ops.RECON_RB_PROBS, 0.7)):
        if kind == ops.RECON_RB_PROBS:   # probs input: feed probabilities
            l.data = torch.sigmoid(l.data).clamp(1e-6, 1 - 1e-6)
        o_ref, g_ref = ref(kind, T)
        lc = l.detach().cuda().requires_grad_(True)
        out = ops.recon_rows(lc, x.cuda(), kind, T)
        out.backward(g.cuda())
        torch.testing.assert_close(out.double().cpu(), o_ref, rtol=2e-5, atol=2e-5 * float(o_ref.abs().max()))
        torch.testing.assert_close(lc.grad.double().cpu(), g_ref, rtol=2e-5, atol=2e-5 * float(g_ref.abs().max()))


@pytest.mark.parametrize("R,G", [(1024, 20000), (77, 333), (5, 1)])
def test_normalize_rnaseq_matches_reference_recipe(R, G):
    """hvae.data.normalize_rnaseq against the reference's pandas / scipy recipe (datasets/jerby_arnon.py:97-106) in float64."""
    from hvae.data import normalize_rnaseq

    torch.manual_seed(R + G)
    counts = torch.poisson(torch.full((R, G), 100.0)) + torch.rand(R, G)
    xd = counts.double()
    ref_z = (xd - xd.mean(0, keepdim=True)) / xd.std(0, unbiased=False, keepdim=True)     # scipy.stats.zscore per column, ddof=0
    ref_m = xd / xd.sum(1, keepdim=True) * 1e6
    xc = counts.cuda()
    z = normalize_rnaseq(xc, "z_score").double().cpu()
    m = normalize_rnaseq(xc, "sum_to_million").double().cpu()
    o = normalize_rnaseq(xc, "sum_to_one").double().cpu()
    if R > 1 and G > 1:
        assert float((z - ref_z).abs().max()) < 2e-5
    assert float(((m - ref_m).abs() / ref_m.abs().clamp_min(1e-30)).max()) < 1e-6
    assert float((o.sum(1) - 1).abs().max()) < 1e-5
    with pytest.raises(ValueError):
        normalize_rnaseq(xc, "nope")
