"""GPU parity: K3 (expmap0/logmap0), mobius_add, expmap/logmap/dist, K4/K5 (WrappedNormal) and the fused
latent head, against (a) the golden fixtures minted from the reference's own files and (b) the
travelling oracle in float32 and float64 on seeded inputs."""
import pytest
import torch

from util_parity import assert_parity, kappa as _kappa, rtol_grad, rtol_val

pytestmark = pytest.mark.gpu

CURV = [0.1, 0.5, 1.0, 1.4, 2.0]


def _hv():
    import hvae

    return hvae


def _oracle_ball(c, dtype):
    from oracle.geoopt_min import PoincareBall

    b = PoincareBall(c=c)
    if dtype == torch.float64:  # same curvature VALUE as the fp32 ball (softplus round trip), held in double
        b.isp_c.data = torch.log(torch.expm1(torch.tensor(float(b.c), dtype=torch.float64)))
    return b


def _run_oracle(fn, inputs, gout, dtype):
    from oracle.geoopt_min.manifolds.stereographic import math as gmath

    xs = [x.detach().cpu().to(dtype).requires_grad_(True) for x in inputs]
    with gmath.fp32_semantics(dtype == torch.float64):
        out = fn(*xs)
        out.backward(gout.detach().cpu().to(dtype))
    return out.detach(), [x.grad for x in xs]


def _run_cuda(fn, inputs, gout):
    xs = [x.detach().cuda().requires_grad_(True) for x in inputs]
    out = fn(*xs)
    out.backward(gout.cuda())
    return out.detach(), [x.grad for x in xs]


def _compare(name, cuda_fn, oracle_fn, inputs, c, gout=None, seed=0, rtol=1e-5, kap=None):
    hv = _hv()
    ball = hv.PoincareBall(c)
    out_c = None
    xs = [x.detach().cuda().requires_grad_(True) for x in inputs]
    out_c = cuda_fn(ball, *xs)
    if gout is None:
        g = torch.Generator().manual_seed(seed)
        gout = torch.randn(out_c.shape, generator=g)
    out_c.backward(gout.cuda())
    o32, g32 = _run_oracle(lambda *a: oracle_fn(_oracle_ball(c, torch.float32), *a), inputs, gout, torch.float32)
    o64, g64 = _run_oracle(lambda *a: oracle_fn(_oracle_ball(c, torch.float64), *a), inputs, gout, torch.float64)
    rv = rtol if kap is None else (rtol_val(kap, rtol) if out_c.dim() > 1 else rtol_val(kap, rtol).squeeze(-1))
    rg = rtol if kap is None else rtol_grad(kap, rtol)
    assert_parity(out_c, o32, o64, what=name + " fwd", rtol=rv)
    for i, x in enumerate(xs):
        assert_parity(x.grad, g32[i], g64[i], what="%s grad[%d]" % (name, i), rtol=rg)


def test_golden_expmap0_logmap0(golden_ops):
    hv = _hv()
    for rec in golden_ops:
        ball = hv.PoincareBall(rec["c_ctor"])
        assert ball.c_value == rec["c"]
        ob64 = _oracle_ball(rec["c_ctor"], torch.float64)
        for key, g in rec.items():
            if key.startswith("expmap0/"):
                u = g["u"].cuda().requires_grad_(True)
                y = ball.expmap0(u)
                y.backward(g["gout"].cuda())
                o64, (g64,) = _run_oracle(ob64.expmap0, [g["u"]], g["gout"], torch.float64)
                assert_parity(y, g["out"], o64, what="golden %s c=%s D=%d" % (key, rec["c"], rec["D"]))
                assert_parity(u.grad, g["gu"], g64, what="golden %s grad" % key)
            elif key.startswith("logmap0/"):
                yv = g["y"].cuda().requires_grad_(True)
                u = ball.logmap0(yv)
                u.backward(g["gout"].cuda())
                o64, (g64,) = _run_oracle(ob64.logmap0, [g["y"]], g["gout"], torch.float64)
                # rows at the projection radius (sqrt(c)|y| = 0.996): d artanh = 1/(1-c|y|^2) amplifies the fp32
                # rounding of |y| by ~125x, so 1e-5 is not attainable there by ANY fp32 evaluation order.
                kap = _kappa(rec["c"], g["y"])
                assert_parity(u, g["out"], o64, what="golden %s c=%s D=%d" % (key, rec["c"], rec["D"]), rtol=rtol_val(kap))
                assert_parity(yv.grad, g["gy"], g64, what="golden %s grad" % key, rtol=rtol_grad(kap))


@pytest.mark.parametrize("D", [1, 2, 3, 5, 8, 10, 16, 33, 64, 100, 200, 512, 777])
@pytest.mark.parametrize("c", [0.5, 1.0, 2.0])
def test_expmap0_logmap0_seeded(D, c):
    torch.manual_seed(D * 7 + 1)
    B = 257
    for s in (1e-3, 0.3, 3.0, 40.0):
        u = torch.randn(B, D) * s / (D ** 0.5)
        u[0].zero_()
        _compare("expmap0 D=%d s=%g" % (D, s), lambda b, x: b.expmap0(x), lambda b, x: b.expmap0(x), [u], c)
        y = _oracle_ball(c, torch.float32).expmap0(u).detach()
        # sqrt(c)|y| -> 0.996 for s >= 3: d artanh = 1/(1-c|y|^2) ~ 125 amplifies fp32 rounding of |y|
        _compare("logmap0 D=%d s=%g" % (D, s), lambda b, x: b.logmap0(x), lambda b, x: b.logmap0(x), [y], c,
                 rtol=1e-5 if s < 3 else 2e-3)


@pytest.mark.parametrize("D", [2, 5, 10, 64, 130])
@pytest.mark.parametrize("c", [0.1, 1.0, 1.4])
def test_binary_maps_seeded(D, c):
    torch.manual_seed(D + 11)
    B = 129
    ob = _oracle_ball(c, torch.float32)
    x = ob.expmap0(torch.randn(B, D) * 0.6 / D ** 0.5).detach()
    y = ob.expmap0(torch.randn(B, D) * 0.9 / D ** 0.5).detach()
    u = torch.randn(B, D) * 0.5 / D ** 0.5
    cf = float(ob.c)
    kap = _kappa(cf, x, y)  # conditioning of the two-point maps at these points (1/(1-c|.|^2))
    _compare("mobius_add", lambda b, p, q: b.mobius_add(p, q), lambda b, p, q: b.mobius_add(p, q), [x, y], c, kap=kap)
    _compare("mobius_add noproj", lambda b, p, q: b.mobius_add(p, q, project=False),
             lambda b, p, q: b.mobius_add(p, q, project=False), [x, y], c, kap=kap)
    _compare("expmap", lambda b, p, q: b.expmap(p, q), lambda b, p, q: b.expmap(p, q), [x, u], c, kap=_kappa(cf, x))
    _compare("logmap", lambda b, p, q: b.logmap(p, q), lambda b, p, q: b.logmap(p, q), [x, y], c, kap=kap)
    _compare("dist", lambda b, p, q: b.dist(p, q), lambda b, p, q: b.dist(p, q), [x, y], c, kap=kap)


def test_golden_mobius_add(golden_ops):
    hv = _hv()
    for rec in golden_ops:
        ball = hv.PoincareBall(rec["c_ctor"])
        g = rec["mobius_add"]
        x, y = g["x"].cuda().requires_grad_(True), g["y"].cuda().requires_grad_(True)
        out = ball.mobius_add(x, y)
        out.backward(g["gout"].cuda())
        ob64 = _oracle_ball(rec["c_ctor"], torch.float64)
        o64, (gx64, gy64) = _run_oracle(ob64.mobius_add, [g["x"], g["y"]], g["gout"], torch.float64)
        assert_parity(out, g["out"], o64, what="golden mobius_add")
        assert_parity(x.grad, g["gx"], gx64, what="golden mobius_add gx")
        assert_parity(y.grad, g["gy"], gy64, what="golden mobius_add gy")


def _oracle_wn(ball, mu, sc):
    from oracle import ref_port as R

    return R.WrappedNormal(mu, sc, ball)


def test_golden_wrapped_normal(golden_ops):
    hv = _hv()
    from hvae.distributions import WrappedNormal

    for rec in golden_ops:
        c, D = rec["c_ctor"], rec["D"]
        ball = hv.PoincareBall(c)
        ob64 = _oracle_ball(c, torch.float64)
        g = rec["rsample"]
        mu, sc = g["mu"].cuda().requires_grad_(True), g["scale"].cuda().requires_grad_(True)
        z = WrappedNormal(mu, sc, ball).rsample(torch.Size([1]), eps=g["eps"].cuda())
        z.backward(g["gout"].cuda())
        o64, (gm64, gs64) = _run_oracle(lambda m, s: _oracle_wn(ob64, m, s).rsample(torch.Size([1]), eps=g["eps"].double()),
                                        [g["mu"], g["scale"]], g["gout"], torch.float64)
        tag = "golden rsample c=%s D=%d" % (c, D)
        kap = _kappa(rec["c"], g["out"], g["mu"])
        assert_parity(z, g["out"], o64, what=tag, rtol=rtol_val(kap).view(1, -1, 1))
        assert_parity(mu.grad, g["gmu"], gm64, what=tag + " gmu", rtol=rtol_grad(kap), slack_mult=3.0)
        assert_parity(sc.grad, g["gscale"], gs64, what=tag + " gscale", rtol=rtol_grad(kap), slack_mult=3.0)
        for name in ("log_prob", "log_prob_rand"):
            g = rec[name]
            mu, sc, zz = (g[k].cuda().requires_grad_(True) for k in ("mu", "scale", "z"))
            lp = WrappedNormal(mu, sc, ball).log_prob(zz)
            lp.backward(g["gout"].cuda())
            o64, g64 = _run_oracle(lambda m, s, z_: _oracle_wn(ob64, m, s).log_prob(z_), [g["mu"], g["scale"], g["z"]],
                                   g["gout"], torch.float64)
            tag = "golden %s c=%s D=%d" % (name, c, D)
            kap = _kappa(rec["c"], g["z"], g["mu"])
            assert_parity(lp, g["out"], o64, what=tag, rtol=rtol_val(kap).view(1, -1, 1), atol=2e-5, slack_mult=3.0)
            for t_, k32, k64 in ((mu, "gmu", 0), (sc, "gscale", 1), (zz, "gz", 2)):
                kk = kap.view(1, -1, 1) if k32 == "gz" else kap
                assert_parity(t_.grad, g[k32], g64[k64], what=tag + " " + k32, rtol=rtol_grad(kk), atol=2e-5, slack_mult=3.0)
        g = rec["log_prob_prior"]
        zz = g["z"].cuda().requires_grad_(True)
        lp = WrappedNormal.origin_prior(D, g["prior_scale"], ball, device="cuda").log_prob(zz)
        lp.backward(g["gout"].cuda())
        def prior64(z_):
            o = ob64.origin(D, dtype=torch.float64)
            return _oracle_wn(ob64, o, torch.ones_like(o) * g["prior_scale"]).log_prob(z_)
        o64, (gz64,) = _run_oracle(prior64, [g["z"]], g["gout"], torch.float64)
        kap = _kappa(rec["c"], g["z"]).view(1, -1, 1)
        assert_parity(lp, g["out"], o64, what="golden prior log_prob c=%s D=%d" % (c, D), rtol=rtol_val(kap), atol=2e-5)
        assert_parity(zz.grad, g["gz"], gz64, what="golden prior log_prob gz", rtol=rtol_grad(kap), atol=2e-5)


@pytest.mark.parametrize("D", [2, 5, 10, 32, 64])
@pytest.mark.parametrize("c", [0.5, 1.0, 2.0])
def test_latent_head_matches_separate_ops_and_oracle(D, c):
    hv = _hv()
    from hvae import ops
    from oracle import ref_port as R

    torch.manual_seed(100 + D)
    B, prior_scale = 300, 1.3
    ob32, ob64 = _oracle_ball(c, torch.float32), _oracle_ball(c, torch.float64)
    mu0 = ob32.expmap0(torch.randn(B, D) * 0.7 / D ** 0.5).detach()
    sc0 = torch.rand(B, D) * 0.8 + 0.2
    eps = torch.randn(1, B, D)
    gz = torch.randn(B, D)
    gkl = torch.randn(B)

    def oracle(ball, dtype):
        mu = mu0.detach().clone().to(dtype).requires_grad_(True)
        sc = sc0.detach().clone().to(dtype).requires_grad_(True)
        q = R.WrappedNormal(mu, sc, ball)
        z = q.rsample(torch.Size([1]), eps=eps.to(dtype))
        o = ball.origin(D, dtype=dtype)
        kl = (q.log_prob(z) - R.WrappedNormal(o, torch.ones_like(o) * prior_scale, ball).log_prob(z)).view(B)
        ((z.squeeze(0) * gz.to(dtype)).sum() + (kl * gkl.to(dtype)).sum()).backward()
        return z.squeeze(0).detach(), kl.detach(), mu.grad, sc.grad

    from oracle.geoopt_min.manifolds.stereographic import math as gmath

    z32, kl32, gm32, gs32 = oracle(ob32, torch.float32)
    with gmath.fp32_semantics():
        z64, kl64, gm64, gs64 = oracle(ob64, torch.float64)
    ball = hv.PoincareBall(c)
    mu = mu0.cuda().requires_grad_(True)
    sc = sc0.cuda().requires_grad_(True)
    z, kl = ops.latent_head(mu, sc, eps[0].cuda(), prior_scale, ball.c_value)
    ((z * gz.cuda()).sum() + (kl * gkl.cuda()).sum()).backward()
    # conditioning: every map here divides by (1 - c|.|^2); at the fp32 projection radius that is 1/8e-3
    kap = _kappa(c, z64, mu0)
    assert_parity(z, z32, z64, what="head z", rtol=rtol_val(kap))
    assert_parity(kl, kl32, kl64, what="head kl", rtol=rtol_val(kap).squeeze(-1), atol=2e-5, slack_mult=3.0)
    assert_parity(mu.grad, gm32, gm64, what="head gmu", rtol=rtol_grad(kap), atol=2e-5, slack_mult=3.0)
    assert_parity(sc.grad, gs32, gs64, what="head gsigma", rtol=rtol_grad(kap), atol=2e-5, slack_mult=3.0)


def test_ops_reject_cpu_tensors():
    hv = _hv()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hv.PoincareBall(1.0).expmap0(torch.randn(4, 2))


def test_empty_and_single_row():
    hv = _hv()
    ball = hv.PoincareBall(1.0)
    assert ball.expmap0(torch.empty(0, 5, device="cuda")).shape == (0, 5)
    y = ball.expmap0(torch.zeros(1, 3, device="cuda"))
    assert torch.equal(y.cpu(), torch.zeros(1, 3))
