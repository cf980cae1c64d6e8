"""GPU parity: K6/K7 HyperbolicRadius (log-normaliser + sigma-gradient, CDF, implicit reparameterisation
gradient, rejection sampler), expmap_polar, RiemannianNormal, and the config-2 (pvae MNIST) train step."""
import math

import pytest
import torch

from util_parity import assert_parity, kappa, rtol_grad, rtol_val

pytestmark = pytest.mark.gpu

GRID = [(2, 1.0), (3, 0.5), (5, 1.0), (10, 1.0), (10, 2.0), (16, 0.7), (32, 1.0)]


def _sigmas(dim):
    lo = 0.1 if dim <= 10 else 0.3  # below this the alternating series itself loses digits (in the reference too)
    return torch.cat([torch.linspace(lo, 1.0, 40), torch.linspace(1.0, 7.0, 40)])


@pytest.mark.parametrize("dim,c", GRID)
def test_lognorm_and_grad(dim, c):
    from hvae import ops
    from oracle.pvae_min.distributions import hyperbolic_radius as hr

    sig = _sigmas(dim)
    s64 = sig.double()
    ct = torch.tensor(c, dtype=torch.float64)
    o = hr.log_normalizer(s64.unsqueeze(-1), ct, dim).squeeze(-1)
    # pvae differentiates logZ with a hand-written backward; autograd through log1p(erf(.)) is NaN once
    # erf saturates, so the oracle gradient here is a float64 central difference of the oracle's logZ.
    h = 1e-6 * s64
    go = (hr.log_normalizer((s64 + h).unsqueeze(-1), ct, dim) - hr.log_normalizer((s64 - h).unsqueeze(-1), ct, dim)).squeeze(-1) / (2 * h)
    sc = sig.cuda().requires_grad_(True)
    lz = ops.hradius_lognorm(sc, dim, c)
    lz.sum().backward()
    assert_parity(lz, None, o, what="logZ dim=%d" % dim, rtol=1e-5, atol=1e-5, row_relative=False)
    assert_parity(sc.grad, None, go, what="dlogZ dim=%d" % dim, rtol=1e-4, atol=1e-4, row_relative=False)


@pytest.mark.parametrize("dim,c", GRID)
def test_cdf_and_implicit_grad(dim, c):
    from hvae import ops
    from oracle.pvae_min.distributions import hyperbolic_radius as hr

    torch.manual_seed(dim)
    sig = torch.rand(64) * 2.0 + 0.3
    ct = torch.tensor(c, dtype=torch.float64)
    mean, var = hr._moments(sig, ct, dim)
    r = (mean + var.sqrt() * torch.randn(3, 64, dtype=torch.float64) * 0.8).clamp_min(0.05).float()  # (S,B)
    F64 = hr.cdf_r(r.double(), sig.double().expand(3, 64), ct, dim)
    gv, gs = hr.grad_cdf_value_scale(r, sig.expand(3, 64), ct, dim)
    cdf = ops.hradius_cdf(r.cuda(), sig.cuda(), dim, c)
    assert_parity(cdf, None, F64, what="cdf", rtol=1e-5, atol=2e-6, row_relative=False)
    rr = r.cuda()
    s = sig.cuda().requires_grad_(True)
    out, _ = ops.hradius_reparam(rr, s, dim, c)
    g = torch.randn(3, 64)
    out.backward(g.cuda())
    ref = (g.double() * (-gs / gv)).sum(0)
    assert torch.equal(out.cpu(), r)
    assert_parity(s.grad, None, ref, what="implicit reparam grad", rtol=5e-5, atol=1e-5, row_relative=False)


@pytest.mark.parametrize("dim,c,sigma", [(2, 1.0, 0.5), (2, 1.0, 3.0), (5, 1.0, 0.15), (10, 1.0, 1.0), (10, 2.0, 0.3), (32, 0.7, 1.5), (1, 1.0, 0.8)])
def test_sampler_kolmogorov_smirnov(dim, c, sigma):
    """Distributional parity: the rejection sampler's radii against the closed-form CDF (oracle, float64)."""
    from hvae import ops
    from oracle.pvae_min.distributions import hyperbolic_radius as hr

    N = 200_000
    sig = torch.full((N,), sigma, device="cuda")
    r = ops.hradius_sample(sig, 1, dim, c, seed=1234, offset=0).view(-1)
    assert torch.isfinite(r).all() and (r > 0).all()
    rs = r.double().cpu().sort().values
    F = hr.cdf_r(rs, torch.full_like(rs, sigma), torch.tensor(c, dtype=torch.float64), dim)
    i = torch.arange(1, N + 1, dtype=torch.float64)
    D = torch.maximum((i / N - F).abs().max(), (F - (i - 1) / N).abs().max()).item()
    # KS critical value at alpha = 1e-3: 1.95/sqrt(N)
    assert D < 1.95 / math.sqrt(N), "KS statistic %.5f (dim=%d c=%g sigma=%g)" % (D, dim, c, sigma)
    # determinism + stream disjointness of the counter-based RNG
    r2 = ops.hradius_sample(sig, 1, dim, c, seed=1234, offset=0).view(-1)
    assert torch.equal(r, r2)
    r3 = ops.hradius_sample(sig[:1000], 1, dim, c, seed=1234, offset=N).view(-1)
    assert not torch.equal(r3, r[:1000])


def _oracle_pball(dim, c, dtype):
    from oracle.pvae_min.manifolds import PoincareBall

    b = PoincareBall(dim, c)
    if dtype == torch.float64:
        b.isp_c.data = torch.log(torch.expm1(torch.tensor(float(b.c), dtype=torch.float64)))
    return b


@pytest.mark.parametrize("D,c", [(2, 1.0), (5, 0.5), (10, 1.0)])
def test_riemannian_normal_injected_noise(D, c):
    import hvae
    from hvae.distributions import RiemannianNormal
    from oracle import ref_port as R
    from oracle.geoopt_min.manifolds.stereographic import math as gmath

    torch.manual_seed(D)
    B = 200
    ob32 = _oracle_pball(D, c, torch.float32)
    mu0 = ob32.expmap0(torch.randn(B, D) * 0.6 / D ** 0.5).detach()
    sg0 = torch.rand(B, 1) * 1.5 + 0.3
    alpha = torch.randn(1, B, D)
    alpha = alpha / alpha.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        r0 = R.RiemannianNormal(mu0, sg0, ob32).radius.sample(torch.Size([1]))  # (1,B,1) oracle ARS radii
    gz, glp = torch.randn(1, B, D), torch.randn(1, B, 1)

    def oracle(dtype):
        ball = _oracle_pball(D, c, dtype)
        mu = mu0.clone().to(dtype).requires_grad_(True)
        sg = sg0.clone().to(dtype).requires_grad_(True)
        with gmath.fp32_semantics(dtype == torch.float64):
            q = R.RiemannianNormal(mu, sg, ball)
            z = q.rsample(torch.Size([1]), alpha=alpha.to(dtype), r=r0.to(dtype))
            lp = q.log_prob(z)
            ((z * gz.to(dtype)).sum() + (lp * glp.to(dtype)).sum()).backward()
        return z.detach(), lp.detach(), mu.grad, sg.grad

    o32, o64 = oracle(torch.float32), oracle(torch.float64)
    ball = hvae.PoincareBall(c)
    mu = mu0.cuda().requires_grad_(True)
    sg = sg0.cuda().requires_grad_(True)
    q = RiemannianNormal(mu, sg, ball)
    z = q.rsample(torch.Size([1]), alpha=alpha.cuda(), r=r0.cuda())
    lp = q.log_prob(z)
    ((z * gz.cuda()).sum() + (lp * glp.cuda()).sum()).backward()
    kap = kappa(c, o64[0], mu0)  # z lands on the projection radius for large r: conditioning ~1/(1-c|z|^2)
    assert_parity(z, o32[0], o64[0], what="RN z", rtol=rtol_val(kap, 2e-5).view(1, -1, 1), atol=2e-6)
    assert_parity(lp, o32[1], o64[1], what="RN log_prob", rtol=rtol_val(kap, 2e-5).view(1, -1, 1), atol=2e-5, row_relative=False, slack_mult=2.0)
    assert_parity(mu.grad, o32[2], o64[2], what="RN gmu", rtol=rtol_grad(kap, 5e-5), atol=2e-5, slack_mult=2.0)
    assert_parity(sg.grad, o32[3], o64[3], what="RN gsigma", rtol=rtol_grad(kap, 5e-5), atol=2e-5, row_relative=False, slack_mult=2.0)


@pytest.mark.parametrize("D,c,B", [(2, 1.0, 200), (5, 0.5, 129), (10, 1.0, 4096)])
def test_riemannian_head_fused_sample_and_kl(D, c, B):
    """RiemannianNormal.rsample_kl (ops.rn_head: one kernel per direction) against (a) the oracle's rsample +
    log_prob(q) - log_prob(p) with injected (alpha, r) in float32 / float64 and (b) this repo's unfused graph
    (rsample + kl_mc + autograd's adds), which runs the same row arithmetic."""
    import hvae
    from hvae.distributions import RiemannianNormal
    from oracle import ref_port as R
    from oracle.geoopt_min.manifolds.stereographic import math as gmath

    torch.manual_seed(D + B)
    ob32 = _oracle_pball(D, c, torch.float32)
    mu0 = ob32.expmap0(torch.randn(B, D) * 0.6 / D ** 0.5).detach()
    sg0 = torch.rand(B, 1) * 1.5 + 0.3
    sp0 = torch.tensor([[1.3]])
    alpha = torch.randn(1, B, D)
    alpha = alpha / alpha.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        r0 = R.RiemannianNormal(mu0, sg0, ob32).radius.sample(torch.Size([1]))
    gz, gkl = torch.randn(1, B, D), torch.randn(1, B)

    def oracle(dtype):
        ball = _oracle_pball(D, c, dtype)
        mu = mu0.clone().to(dtype).requires_grad_(True)
        sg = sg0.clone().to(dtype).requires_grad_(True)
        with gmath.fp32_semantics(dtype == torch.float64):
            q = R.RiemannianNormal(mu, sg, ball)
            p = R.RiemannianNormal(torch.zeros(1, D, dtype=dtype), sp0.to(dtype), ball)
            z = q.rsample(torch.Size([1]), alpha=alpha.to(dtype), r=r0.to(dtype))
            kl = q.log_prob(z).sum(-1) - p.log_prob(z).sum(-1)
            ((z * gz.to(dtype)).sum() + (kl * gkl.to(dtype)).sum()).backward()
        return z.detach(), kl.detach(), mu.grad, sg.grad

    def ours(fused):
        ball = hvae.PoincareBall(c)
        mu = mu0.cuda().requires_grad_(True)
        sg = sg0.cuda().requires_grad_(True)
        q = RiemannianNormal(mu, sg, ball)
        p = RiemannianNormal(torch.zeros(1, D, device="cuda"), sp0.cuda(), ball)
        if fused:
            z, kl = q.rsample_kl(p, alpha=alpha.cuda(), r=r0.cuda())
        else:
            z = q.rsample(torch.Size([1]), alpha=alpha.cuda(), r=r0.cuda())
            kl = q.kl_mc(z, p)
        ((z * gz.cuda()).sum() + (kl * gkl.cuda()).sum()).backward()
        return z.detach(), kl.detach(), mu.grad, sg.grad

    fu, un = ours(True), ours(False)
    for name, a_, b_ in zip(("z", "kl", "gmu", "gsigma"), fu, un):
        sc = float(b_.abs().max())
        assert float((a_ - b_).abs().max()) <= 2e-6 * sc, (name, float((a_ - b_).abs().max()), sc)
    if B <= 512:   # the oracle's ARS and float64 series are slow; the big case is covered by the config-2 step test
        o32, o64 = oracle(torch.float32), oracle(torch.float64)
        kap = kappa(c, o64[0], mu0)
        assert_parity(fu[0], o32[0], o64[0], what="RN head z", rtol=rtol_val(kap, 2e-5).view(1, -1, 1), atol=2e-6)
        assert_parity(fu[1], o32[1], o64[1], what="RN head kl", rtol=rtol_val(kap, 2e-5).view(1, -1), atol=2e-5, row_relative=False, slack_mult=2.0)
        assert_parity(fu[2], o32[2], o64[2], what="RN head gmu", rtol=rtol_grad(kap, 5e-5), atol=2e-5, slack_mult=2.0)
        assert_parity(fu[3], o32[3], o64[3], what="RN head gsigma", rtol=rtol_grad(kap, 5e-5), atol=2e-5, row_relative=False, slack_mult=2.0)


@pytest.mark.parametrize("fused", [True, False])
def test_pvae_mnist_step_matches_oracle(fused):
    """Config 2 at reduced width: same weights, same data, same injected (alpha, r)."""
    from hvae import models as HM
    from oracle import ref_port as R
    from oracle.geoopt_min.manifolds.stereographic import math as gmath
    from oracle.geoopt_min.manifolds.stereographic.manifold import PoincareBall as OBall

    torch.manual_seed(7)
    B, D, H = 96, 10, 64
    o32 = R.PvaeMnist(latent_dim=D, hidden_dim=H, c=1.0)
    x = torch.rand(B, 1, 28, 28).clamp(1e-5, 1 - 1e-5)
    alpha = torch.randn(1, B, D)
    alpha = alpha / alpha.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        mu_, sg_ = o32.encode(x)
        r = R.RiemannianNormal(mu_, sg_, o32.manifold).radius.sample(torch.Size([1]))
    sd = {k: v.clone() for k, v in o32.state_dict().items()}

    def run_oracle(dtype):
        m = R.PvaeMnist(latent_dim=D, hidden_dim=H, c=1.0)
        m.load_state_dict(sd)
        if dtype == torch.float64:
            c32 = {id(b): float(b.c) for b in m.modules() if isinstance(b, OBall)}
            m = m.double()
            for b in m.modules():
                if isinstance(b, OBall):
                    b.isp_c.data = torch.log(torch.expm1(torch.tensor(c32[id(b)], dtype=torch.float64)))
        with gmath.fp32_semantics(dtype == torch.float64):
            L = m.loss(x.to(dtype), alpha=alpha.to(dtype), r=r.to(dtype))
            L["loss_total"].backward()
        return L, {k: p.grad for k, p in m.named_parameters() if p.grad is not None}

    L32, G32 = run_oracle(torch.float32)
    L64, G64 = run_oracle(torch.float64)
    model = HM.PvaeMnist(latent_dim=D, hidden_dim=H, c=1.0, fused=fused)
    sd_c = {k: v for k, v in sd.items() if not k.endswith("manifold.dim")}  # pvae's PoincareBall(dim, c) buffer
    missing, unexpected = model.load_state_dict(sd_c, strict=False)
    assert not unexpected and all("isp_c" in k for k in missing), (missing, unexpected)
    model = model.cuda()
    L = model.loss(x.cuda(), alpha=alpha.cuda(), r=r.cuda())
    L["loss_total"].backward()
    for k in L32:
        assert_parity(L[k].reshape(1), L32[k].reshape(1), L64[k].reshape(1), what="pvae " + k, rtol=1e-5, atol=1e-4,
                      row_relative=False, slack_mult=2.0)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == set(G32), set(grads) ^ set(G32)
    for k in G32:
        assert_parity(grads[k], G32[k], G64[k], what="pvae grad " + k, rtol=5e-5, atol=1e-6, norm_relative=True, slack_mult=2.0)


def test_riemannian_normal_golden(golden_riemannian):
    """a-7 against the fixture minted by running the reference's OWN old_pvae_riemannian_normal.py (its draws of
    (alpha, r) recorded and injected here): sigma clamp [0.1, 7] incl. its zero gradient, z = expmap_polar, the implicit
    reparameterisation gradient, log_prob (+ gradients) with a clamped-high row, the log-normaliser, the prior form."""
    import hvae
    from hvae.distributions import RiemannianNormal
    from oracle import ref_port as R
    from oracle.geoopt_min.manifolds.stereographic import math as gmath

    for g in golden_riemannian:
        D, c = g["D"], g["c_ctor"]

        def oracle64():
            ball = _oracle_pball(D, c, torch.float64)
            mu = g["mu"].double().requires_grad_(True)
            sg = g["scale"].double().requires_grad_(True)
            mu2, sg2, zz = g["mu"].double().requires_grad_(True), g["scale_lp"].double().requires_grad_(True), g["z"].double().requires_grad_(True)
            with gmath.fp32_semantics(True):
                q = R.RiemannianNormal(mu, sg, ball)
                z = q.rsample(torch.Size([1]), alpha=g["alpha"].double(), r=g["r"].detach().double())
                z.backward(g["gz_up"].double())
                q2 = R.RiemannianNormal(mu2, sg2, ball)
                lp = q2.log_prob(zz)
                lp.backward(g["glp_up"].double())
                lz = q.radius.log_normalizer
            return dict(z=z.detach(), gmu_z=mu.grad, gscale_z=sg.grad, log_prob=lp.detach(), gmu_lp=mu2.grad, gscale_lp=sg2.grad,
                        gz_lp=zz.grad, logZ=lz.detach())

        o64 = oracle64()
        ball = hvae.PoincareBall(c)
        assert ball.c_value == pytest.approx(g["c"], rel=1e-7)
        mu, sg = g["mu"].cuda().requires_grad_(True), g["scale"].cuda().requires_grad_(True)
        q = RiemannianNormal(mu, sg, ball)
        assert torch.equal(q.scale.detach().cpu(), g["scale_clamped"])
        z = q.rsample(torch.Size([1]), alpha=g["alpha"].cuda(), r=g["r"].detach().cuda())
        z.backward(g["gz_up"].cuda())
        tag = "RN golden c=%s D=%d " % (c, D)
        kap = kappa(g["c"], o64["z"], g["mu"])
        assert_parity(q.radius.log_normalizer, g["logZ"], o64["logZ"], what=tag + "logZ", rtol=1e-5, atol=1e-5, row_relative=False)
        assert_parity(z, g["z"], o64["z"], what=tag + "z", rtol=rtol_val(kap, 2e-5).view(1, -1, 1), atol=2e-6)
        assert_parity(mu.grad, g["gmu_z"], o64["gmu_z"], what=tag + "gmu(z)", rtol=rtol_grad(kap, 5e-5), atol=2e-5, slack_mult=2.0)
        assert_parity(sg.grad, g["gscale_z"], o64["gscale_z"], what=tag + "gsigma(z)", rtol=rtol_grad(kap, 5e-5), atol=2e-5,
                      row_relative=False, slack_mult=2.0)
        assert float(sg.grad[0].abs().max()) == 0.0   # sigma below the clamp: no gradient, as in the reference
        mu2, sg2 = g["mu"].cuda().requires_grad_(True), g["scale_lp"].cuda().requires_grad_(True)
        zz = g["z"].cuda().requires_grad_(True)
        q2 = RiemannianNormal(mu2, sg2, ball)
        assert torch.equal(q2.scale.detach().cpu(), g["scale_lp_clamped"])
        lp = q2.log_prob(zz)
        lp.backward(g["glp_up"].cuda())
        assert_parity(lp, g["log_prob"], o64["log_prob"], what=tag + "log_prob", rtol=rtol_val(kap, 2e-5).view(1, -1, 1), atol=2e-5,
                      row_relative=False, slack_mult=2.0)
        assert_parity(mu2.grad, g["gmu_lp"], o64["gmu_lp"], what=tag + "gmu(lp)", rtol=rtol_grad(kap, 5e-5), atol=2e-5, slack_mult=2.0)
        assert_parity(sg2.grad, g["gscale_lp"], o64["gscale_lp"], what=tag + "gsigma(lp)", rtol=rtol_grad(kap, 5e-5), atol=2e-5,
                      row_relative=False, slack_mult=2.0)
        assert float(sg2.grad[1].abs().max()) == 0.0  # sigma above the clamp
        assert_parity(zz.grad, g["gz_lp"], o64["gz_lp"], what=tag + "gz(lp)", rtol=rtol_grad(kap, 5e-5).view(1, -1, 1), atol=2e-5, slack_mult=2.0)
        p0 = RiemannianNormal(torch.zeros(1, D, device="cuda"), torch.full((1, 1), g["prior_sigma"], device="cuda"), ball)
        lp0 = p0.log_prob(g["z"].cuda())
        torch.testing.assert_close(lp0.cpu(), g["log_prob_prior"], rtol=2e-5, atol=2e-4)
