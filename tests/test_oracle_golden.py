"""CPU: pin the travelling oracle (oracle/ref_port.py over geoopt_min/pvae_min) against the golden
fixtures minted from the reference's OWN files (tests/golden/make_golden.py).  Same arithmetic on
the same CPU => tolerance is a few ulp."""
import hashlib
import json
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
    for fn in ("ops_golden.pt", "models_golden.pt"):
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
