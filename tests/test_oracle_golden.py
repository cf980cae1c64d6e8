"""CPU: pin the travelling oracle (oracle/ref_port.py over geoopt_min/pvae_min) against the golden
fixtures minted from the reference's OWN files (tests/golden/make_golden.py).  Same arithmetic on
the same CPU => tolerance is a few ulp."""
import hashlib
import json
import math
import os

import pytest
import torch

from oracle import ref_port as R
from oracle.geoopt_min import PoincareBall
from oracle.geoopt_min.layers.stereographic import Distance2StereographicHyperplanes

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 2e-6, 1e-7


def close(a, b, rtol=RTOL, atol=ATOL):
    torch.testing.assert_close(torch.Tensor(a), torch.Tensor(b), rtol=rtol, atol=atol, equal_nan=True)


def close_norm(a, b, tol=1e-4):
    """Reductions over rows cancel: compare against the tensor's own scale (max-norm relative)."""
    a, b = torch.Tensor(a), torch.Tensor(b)
    scale = b.abs().max().item()
    assert (a - b).abs().max().item() <= tol * scale + 1e-7, ((a - b).abs().max().item(), scale)


def test_manifest_hashes():
    man = json.load(open(os.path.join(HERE, "golden", "MANIFEST.json")))
    for fn in ("ops_golden.pt", "models_golden.pt", "riemannian_golden.pt"):
        h = hashlib.sha256(open(os.path.join(HERE, "golden", fn), "rb").read()).hexdigest()
        assert h == man[fn]["sha256"]


def _bwd(out, gout, inputs):
    for v in inputs:
        v.grad = None
    out.backward(gout)
    return [v.grad for v in inputs]


def test_ops_against_golden(golden_ops):
    for rec in golden_ops:
        ball = PoincareBall(c=rec["c_ctor"])
        assert float(ball.c) == rec["c"]
        D = rec["D"]
        for key, g in rec.items():
            if key.startswith("expmap0/"):
                u = g["u"].clone().requires_grad_(True)
                out = ball.expmap0(u)
                close(out, g["out"]); close(_bwd(out, g["gout"], [u])[0], g["gu"])
            elif key.startswith("logmap0/"):
                y = g["y"].clone().requires_grad_(True)
                out = ball.logmap0(y)
                close(out, g["out"]); close(_bwd(out, g["gout"], [y])[0], g["gy"])
        g = rec["rsample"]
        mu, sc = g["mu"].clone().requires_grad_(True), g["scale"].clone().requires_grad_(True)
        out = R.WrappedNormal(mu, sc, ball).rsample(torch.Size([1]), eps=g["eps"])
        close(out, g["out"])
        gm, gs = _bwd(out, g["gout"], [mu, sc])
        close(gm, g["gmu"]); close(gs, g["gscale"])
        for name in ("log_prob", "log_prob_rand"):
            g = rec[name]
            mu, sc, z = (g[k].clone().requires_grad_(True) for k in ("mu", "scale", "z"))
            out = R.WrappedNormal(mu, sc, ball).log_prob(z)
            close(out, g["out"], rtol=1e-5, atol=1e-6)
            for a, b in zip(_bwd(out, g["gout"], [mu, sc, z]), (g["gmu"], g["gscale"], g["gz"])):
                close(a, b, rtol=1e-5, atol=1e-6)
        g = rec["gyroplane_bias"]
        lay = R.Distance2PoincareHyperplanes(D, g["points"].shape[0], ball=ball)
        lay.points.data.copy_(g["points"]); lay.bias.data.copy_(g["bias"])
        x = g["x"].clone().requires_grad_(True)
        out = lay(x)
        close(out, g["out"])
        for a, b in zip(_bwd(out, g["gout"], [x, lay.points, lay.bias]), (g["gx"], g["gpoints"], g["gbias"])):
            close(a, b)
        g = rec["geodesic"]
        geo = R.GeodesicLayer(D, g["_weight"].shape[0], ball)
        geo._weight.data.copy_(g["_weight"]); geo._bias.data.copy_(g["_bias"])
        x = g["x"].clone().requires_grad_(True)
        out = geo(x)  # batched pvae semantics == reference looped over B==1 rows
        close(out, g["out"])
        for a, b in zip(_bwd(out, g["gout"], [x, geo._weight, geo._bias]), (g["gx"], g["g_weight"], g["g_bias"])):
            close_norm(a, b)  # batched vs row-looped accumulation order
        for name in ("mobius_layer", "mobius_layer_zero_w"):
            g = rec[name]
            mob = R.MobiusLayer(g["_weight"].shape[1], D, ball)
            mob._weight.data.copy_(g["_weight"]); mob._bias.data.copy_(g["_bias"])
            x = g["x"].clone().requires_grad_(True)
            out = mob(x)
            close(out, g["out"])
            for a, b in zip(_bwd(out, g["gout"], [x, mob._weight, mob._bias]), (g["gx"], g["g_weight"], g["g_bias"])):
                close(a, b, rtol=1e-5, atol=1e-6)


def _build(name):
    if name == "A":
        return R.ModelA(torch.Size([1, 10, 10]), 2, 1.0, 1.0, 1.0)
    if name == "A_c0.5_D5":
        return R.ModelA(torch.Size([1, 10, 10]), 5, 0.5, 2.0, 2.0)
    if name == "B":
        return R.ModelB((1, 16, 16), 2, 1.0, "mobius", "geoopt_gyroplane", 1.0, "mse")
    if name == "B_c1.4_D8":
        return R.ModelB((1, 16, 16), 8, 1.4, "mobius", "geoopt_gyroplane", 2.0, "mse")
    if name == "B_linear_bernoulli":
        return R.ModelB((1, 16, 16), 2, 1.0, "linear", "geoopt_gyroplane", 1.0, "bernoulli")
    if name == "C":
        return R.ModelC(torch.Size([300]), 5, 1.0, 32, 0.5)
    if name == "OneB":
        return R.ModelOneB(torch.Size([300]), 32, 2, 1.0, 2.0, 0.5, "logmap0_analytic")
    if name == "OneB_log_prob":
        return R.ModelOneB(torch.Size([300]), 32, 3, 0.5, 1.0, 1.0, "log_prob")
    raise KeyError(name)


@pytest.mark.parametrize("name", ["A", "A_c0.5_D5", "B", "B_c1.4_D8", "B_linear_bernoulli", "C", "OneB", "OneB_log_prob"])
def test_models_against_golden(golden_models, name):
    g = golden_models[name]
    model = _build(name)
    missing, unexpected = model.load_state_dict(g["state_dict"], strict=False)
    assert not [k for k in missing if "isp_c" not in k], missing
    assert not unexpected, unexpected
    losses = model.loss(g["x"], eps=g["eps"])
    for k, v in g["losses"].items():
        close(losses[k], v, rtol=1e-5, atol=1e-6)
    losses["loss_total"].backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == set(g["grads"]), set(grads) ^ set(g["grads"])
    for k, v in g["grads"].items():
        close_norm(grads[k], v, tol=2e-5)


@pytest.mark.reference
def test_ref_port_matches_reference_live():
    """Authoring container only: run the reference's real MobiusLayer/WrappedNormal beside ref_port."""
    import importlib

    from oracle import reference_loader as rl

    rl.load()
    W = importlib.import_module("hyperbolic_vae.distributions.wrapped_normal")
    import geoopt

    ball = geoopt.PoincareBall(c=0.7)
    torch.manual_seed(3)
    mu = ball.expmap0(torch.randn(9, 4) * 0.5)
    sc = torch.rand(9, 4) + 0.3
    torch.manual_seed(5)
    z_ref = W.WrappedNormal(mu, sc, ball).rsample(torch.Size([1]))
    torch.manual_seed(5)
    eps = torch.randn(1, 9, 4)
    z = R.WrappedNormal(mu, sc, ball).rsample(torch.Size([1]), eps=eps)
    close(z, z_ref)
    close(R.WrappedNormal(mu, sc, ball).log_prob(z), W.WrappedNormal(mu, sc, ball).log_prob(z_ref), rtol=1e-5, atol=1e-6)


# ---- a-7: RiemannianNormal / HyperbolicRadius ---------------------------------------------------------------------------
def test_riemannian_normal_against_golden(golden_riemannian):
    """The travelling restatement (ref_port.RiemannianNormal, injected alpha / r) against what the reference's OWN
    distributions/old_pvae_riemannian_normal.py computed for the draws it made (tests/golden/make_golden.py)."""
    from oracle.pvae_min.manifolds import PoincareBall as PBall

    for g in golden_riemannian:
        ball = PBall(g["D"], g["c_ctor"])
        assert float(ball.c) == g["c"]
        mu, sc = g["mu"].clone().requires_grad_(True), g["scale"].clone().requires_grad_(True)
        q = R.RiemannianNormal(mu, sc, ball)
        close(q.scale, g["scale_clamped"])                      # clamp [0.1, 7]
        assert float(q.scale.min()) == pytest.approx(0.1)
        close(q.radius.log_normalizer, g["logZ"], rtol=1e-6)
        z = q.rsample(torch.Size([1]), alpha=g["alpha"], r=g["r"].detach())
        close(z, g["z"])
        gmu, gsc = _bwd(z, g["gz_up"], [mu, sc])
        close(gmu, g["gmu_z"], rtol=1e-5, atol=1e-6)
        close(gsc, g["gscale_z"], rtol=1e-5, atol=1e-6)
        assert float(gsc[0].abs().max()) == 0.0                # the clamped row carries no gradient
        mu2, sc2, zz = g["mu"].clone().requires_grad_(True), g["scale_lp"].clone().requires_grad_(True), g["z"].clone().requires_grad_(True)
        q2 = R.RiemannianNormal(mu2, sc2, ball)
        close(q2.scale, g["scale_lp_clamped"])
        assert float(q2.scale.max()) == pytest.approx(7.0)
        lp = q2.log_prob(zz)
        close(lp, g["log_prob"], rtol=1e-5, atol=1e-5)
        gmu, gsc, gz = _bwd(lp, g["glp_up"], [mu2, sc2, zz])
        close(gmu, g["gmu_lp"], rtol=1e-4, atol=1e-5)
        close(gsc, g["gscale_lp"], rtol=1e-4, atol=1e-5)
        close(gz, g["gz_lp"], rtol=1e-4, atol=1e-5)
        p0 = R.RiemannianNormal(ball.zero, torch.full((1, 1), g["prior_sigma"]), ball)
        close(p0.log_prob(g["z"]), g["log_prob_prior"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dim,c", [(2, 1.0), (3, 0.5), (5, 1.0), (10, 1.0), (10, 2.0), (16, 0.7)])
def test_radius_normaliser_and_cdf_against_quadrature(dim, c):
    """pvae_min's closed-form log-normaliser (signed log-sum-exp over erf terms) and radial CDF against brute-force
    quadrature of rho(r) ~ exp(-r^2/2s^2) (sinh(sqrt(c) r)/sqrt(c))^(dim-1) in float64 - the third-party arithmetic has
    no copy here to diff against, so its maths is pinned to the integral it claims to evaluate."""
    from oracle.pvae_min.distributions import hyperbolic_radius as hr

    ct = torch.tensor(c, dtype=torch.float64)
    for s in (0.1, 0.35, 1.0, 2.5, 7.0):
        if dim >= 10 and s < 0.3:
            continue  # the alternating series loses digits there (in pvae too); covered by the GPU tests' range
        sc = math.sqrt(c)
        n = dim - 1
        mode = 0.5 * (n * sc * s * s + math.sqrt((n * sc * s * s) ** 2 + 4 * n * s * s)) if n else 0.0
        hi = mode + 14 * s + 1.0
        r = torch.linspace(0, hi, 400001, dtype=torch.float64)
        logf = -r.pow(2) / (2 * s * s) + n * (torch.log(torch.sinh((sc * r).clamp_max(300)).clamp_min(1e-300)) - 0.5 * math.log(c))
        if n:
            big = sc * r > 300
            logf = torch.where(big, -r.pow(2) / (2 * s * s) + n * (sc * r - math.log(2) - 0.5 * math.log(c)), logf)
        m = logf.max()
        f = torch.exp(logf - m)
        Z = torch.trapezoid(f, r)
        logZ_q = float(m + Z.log())
        logZ = float(hr.log_normalizer(torch.tensor([[s]], dtype=torch.float64), ct, dim))
        assert abs(logZ - logZ_q) <= 1e-7 * max(1.0, abs(logZ_q)), (dim, c, s, logZ, logZ_q)
        cum = torch.cumulative_trapezoid(f, r) / Z
        for frac in (0.1, 0.5, 0.9):
            i = int(torch.searchsorted(cum, torch.tensor(frac, dtype=torch.float64)))
            F = float(hr.cdf_r(r[i + 1].view(1, 1), torch.tensor([[s]], dtype=torch.float64), ct, dim))
            assert abs(F - float(cum[i])) <= 2e-6, (dim, c, s, frac, F, float(cum[i]))
