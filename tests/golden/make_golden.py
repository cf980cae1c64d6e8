"""Mint the golden fixtures by running the REFERENCE'S OWN FILES (imported unmodified from
/root/reference through oracle/reference_loader.py, over the geoopt_min / pvae_min shims).

Run in the authoring container only:   python tests/golden/make_golden.py
Outputs (committed):  tests/golden/ops_golden.pt, tests/golden/models_golden.pt, MANIFEST.json

Everything stored is a plain tensor; consumers never need /root/reference.
"""
import hashlib
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CURVATURES = [0.1, 0.5, 1.0, 1.4, 2.0]


def t(x):
    return x.detach().clone().contiguous().to(torch.float32) if x.dtype != torch.float64 else x.detach().clone()


def plain(x):
    return torch.Tensor(x.detach()).clone() if isinstance(x, torch.Tensor) else x


def grads_of(out_fn, inputs, gout_seed=7):
    """Run out_fn() with grad, backprop a fixed random upstream gradient, return (out, gout, grads)."""
    for v in inputs:
        v.grad = None
    out = out_fn()
    g = torch.Generator().manual_seed(gout_seed)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout)
    return plain(out), gout, [plain(v.grad) if v.grad is not None else None for v in inputs]


def make_ops():
    rl.load()
    import geoopt

    L = importlib.import_module("hyperbolic_vae.layers")
    M = importlib.import_module("hyperbolic_vae.manifolds")
    W = importlib.import_module("hyperbolic_vae.distributions.wrapped_normal")
    from torch.distributions.utils import _standard_normal

    cases = []
    torch.manual_seed(42)  # the reference's own seed (scripts/_6:63)
    for c in CURVATURES:
        ball = geoopt.PoincareBall(c=c)
        cf = float(ball.c)
        for D in (2, 5, 10, 64):
            B = 12
            rec = {"c_ctor": c, "c": cf, "D": D, "B": B}
            # ---- expmap0 / logmap0 over scales incl. clamp/projection edge rows
            for s in (1e-8, 1e-3, 1.0, 10.0, 1e3):
                u = (torch.randn(B, D) * s).requires_grad_(True)
                u.data[0].zero_()  # ||u|| = 0 row
                out, gout, (gu,) = grads_of(lambda: ball.expmap0(u), [u])
                rec[f"expmap0/{s}"] = dict(u=plain(u), out=out, gout=gout, gu=gu)
                y = plain(out).clone().requires_grad_(True)
                out, gout, (gy,) = grads_of(lambda: ball.logmap0(y), [y])
                rec[f"logmap0/{s}"] = dict(y=plain(y), out=out, gout=gout, gy=gy)
            # ---- mobius_add
            x = ball.expmap0(torch.randn(B, D) * 0.5).detach().requires_grad_(True)
            y = ball.expmap0(torch.randn(B, D) * 0.8).detach().requires_grad_(True)
            out, gout, (gx, gy) = grads_of(lambda: ball.mobius_add(x, y), [x, y])
            rec["mobius_add"] = dict(x=plain(x), y=plain(y), out=out, gout=gout, gx=gx, gy=gy)
            # ---- WrappedNormal rsample (seeded noise recorded) + log_prob + KL against the prior
            mu = ball.expmap0(torch.randn(B, D) * 0.7).detach()
            mu[1] = ball.expmap0(torch.randn(D) * 50.0)  # on the projection boundary
            mu.requires_grad_(True)
            sc = (torch.rand(B, D) + 0.2).requires_grad_(True)
            torch.manual_seed(1000 + D)
            eps = _standard_normal(torch.Size([1, B, D]), dtype=torch.float32, device=torch.device("cpu"))
            torch.manual_seed(1000 + D)
            out, gout, (gmu, gsc) = grads_of(lambda: W.WrappedNormal(mu, sc, ball).rsample(torch.Size([1])), [mu, sc])
            rec["rsample"] = dict(mu=plain(mu), scale=plain(sc), eps=eps, out=out, gout=gout, gmu=gmu, gscale=gsc)
            z = plain(out).clone().requires_grad_(True)  # (1,B,D)
            out, gout, (gmu, gsc, gz) = grads_of(lambda: W.WrappedNormal(mu, sc, ball).log_prob(z), [mu, sc, z])
            rec["log_prob"] = dict(mu=plain(mu), scale=plain(sc), z=plain(z), out=out, gout=gout, gmu=gmu, gscale=gsc, gz=gz)
            zr = ball.expmap0(torch.randn(1, B, D) * 0.6).detach().requires_grad_(True)  # random ball points
            out, gout, (gmu, gsc, gz) = grads_of(lambda: W.WrappedNormal(mu, sc, ball).log_prob(zr), [mu, sc, zr])
            rec["log_prob_rand"] = dict(mu=plain(mu), scale=plain(sc), z=plain(zr), out=out, gout=gout, gmu=gmu, gscale=gsc, gz=gz)
            origin = ball.origin(D)
            out, gout, (gz,) = grads_of(lambda: W.WrappedNormal(origin, torch.ones_like(origin) * 1.5, ball).log_prob(zr), [zr])
            rec["log_prob_prior"] = dict(z=plain(zr), prior_scale=1.5, out=out, gout=gout, gz=gz)
            out, gout, (gx2, gy2) = grads_of(lambda: M.logdetexp(ball, x, y, keepdim=True), [x, y])
            rec["logdetexp"] = dict(x=plain(x), y=plain(y), out=out, gout=gout, gx=gx2, gy=gy2)
            # ---- gyroplane: reference's local layer (with bias), geoopt's layer, squared variant
            P = 17
            torch.manual_seed(2000 + D)
            lay = L.Distance2PoincareHyperplanes(D, P, ball=ball)
            xin = ball.expmap0(torch.randn(B, D) * 0.9).detach()
            xin[2] = plain(lay.points)[3] * (1 + 1e-4)  # x -> p
            xin[3] = ball.expmap0(torch.randn(D) * 50.0)  # ||x|| at the projection boundary
            lay.points.data[5].mul_(1e-9)  # plane through ~origin
            xin.requires_grad_(True)
            out, gout, (gx3, gp, gb) = grads_of(lambda: lay(xin), [xin, lay.points, lay.bias])
            rec["gyroplane_bias"] = dict(x=plain(xin), points=plain(lay.points), bias=plain(lay.bias), out=out, gout=gout, gx=gx3, gpoints=gp, gbias=gb)
            lay2 = geoopt.layers.stereographic.Distance2StereographicHyperplanes(D, P, ball=ball)
            out, gout, (gx3, gp) = grads_of(lambda: lay2(xin), [xin, lay2.points])
            rec["gyroplane_geoopt"] = dict(x=plain(xin), points=plain(lay2.points), out=out, gout=gout, gx=gx3, gpoints=gp)
            # (the reference's own layer raises KeyError for bias=False, layers.py:185-188, so use geoopt's)
            lay3 = geoopt.layers.stereographic.Distance2StereographicHyperplanes(D, P, signed=True, squared=True, ball=ball)
            out, gout, (gx3, gp) = grads_of(lambda: lay3(xin), [xin, lay3.points])
            rec["gyroplane_squared"] = dict(x=plain(xin), points=plain(lay3.points), out=out, gout=gout, gx=gx3, gpoints=gp)
            # ---- GeodesicLayer: the reference's forward only accepts B==1 (layers.py:98-102) -> loop rows
            geo = L.GeodesicLayer(D, P, ball)
            xs = ball.expmap0(torch.randn(6, D) * 0.9).detach().requires_grad_(True)
            for v in (geo._weight, geo._bias):
                v.grad = None
            outs = torch.cat([geo(xs[i : i + 1]) for i in range(xs.shape[0])], 0)
            g = torch.Generator().manual_seed(7)
            gout = torch.randn(outs.shape, generator=g)
            outs.backward(gout)
            rec["geodesic"] = dict(x=plain(xs), _weight=plain(geo._weight), _bias=plain(geo._bias), out=plain(outs), gout=gout,
                                   gx=plain(xs.grad), g_weight=plain(geo._weight.grad), g_bias=plain(geo._bias.grad))
            # ---- MobiusLayer: Euclidean (off-ball) features, zero rows, zero weight rows
            Fdim = 48
            mob = L.MobiusLayer(Fdim, D, ball)
            e = torch.randn(B, Fdim) * 3.0  # off ball: ||e|| >> 1/sqrt(c)
            e[0].zero_()
            e[4] = torch.randn(Fdim) * 1e-3  # inside the ball
            e.requires_grad_(True)
            out, gout, (ge, gw, gb) = grads_of(lambda: mob(e), [e, mob._weight, mob._bias])
            rec["mobius_layer"] = dict(x=plain(e), _weight=plain(mob._weight), _bias=plain(mob._bias), out=out, gout=gout,
                                       gx=ge, g_weight=gw, g_bias=gb, weight=plain(mob.weight))
            mob._weight.data[0].zero_()  # zero weight row
            out, gout, (ge, gw, gb) = grads_of(lambda: mob(e), [e, mob._weight, mob._bias])
            rec["mobius_layer_zero_w"] = dict(x=plain(e), _weight=plain(mob._weight), _bias=plain(mob._bias), out=out, gout=gout,
                                              gx=ge, g_weight=gw, g_bias=gb)
            cases.append(rec)
    return cases


def make_models():
    rl.load()
    mA = importlib.import_module("hyperbolic_vae.models.vae_hyperbolic_gyroplane_decoder")
    mB = importlib.import_module("hyperbolic_vae.models.vae_hyperbolic")
    mC = importlib.import_module("hyperbolic_vae.models.vae_hyperbolic_rnaseq")
    mD = importlib.import_module("hyperbolic_vae.models.vae_one_b")
    from torch.distributions.utils import _standard_normal

    out = {}

    def run(name, model, batch, x, eps_shape, seed):
        torch.manual_seed(seed)
        eps = _standard_normal(torch.Size(eps_shape), dtype=torch.float32, device=torch.device("cpu"))
        torch.manual_seed(seed)
        model.zero_grad()
        losses = model.loss(batch)
        losses["loss_total"].backward()
        out[name] = dict(
            state_dict={k: plain(v) for k, v in model.state_dict().items()},
            x=x, eps=eps,
            losses={k: plain(v) for k, v in losses.items()},
            grads={k: plain(p.grad) for k, p in model.named_parameters() if p.grad is not None},
        )

    torch.manual_seed(42)
    B = 16
    A = mA.VAEHyperbolicGyroplaneDecoder(data_shape=torch.Size([1, 10, 10]), latent_dim=2, manifold_curvature=1.0, prior_scale=1.0)
    x = torch.rand(B, 1, 10, 10)
    run("A", A, (x, None), x, (1, B, 2), 11)
    A2 = mA.VAEHyperbolicGyroplaneDecoder(data_shape=torch.Size([1, 10, 10]), latent_dim=5, manifold_curvature=0.5, beta=2.0, prior_scale=2.0)
    run("A_c0.5_D5", A2, (x, None), x, (1, B, 5), 12)
    Bm = mB.VAEHyperbolicExperiment((1, 16, 16), 2, 1.0, "mobius", "geoopt_gyroplane", loss_recon="mse")
    x32 = torch.rand(B, 1, 16, 16)
    run("B", Bm, (x32, None), x32, (1, B, 2), 13)
    Bm2 = mB.VAEHyperbolicExperiment((1, 16, 16), 8, 1.4, "mobius", "geoopt_gyroplane", beta=2.0, loss_recon="mse")
    run("B_c1.4_D8", Bm2, (x32, None), x32, (1, B, 8), 14)
    Bm3 = mB.VAEHyperbolicExperiment((1, 16, 16), 2, 1.0, "linear", "geoopt_gyroplane", loss_recon="bernoulli")
    run("B_linear_bernoulli", Bm3, (x32.clamp(1e-5, 1 - 1e-5), None), x32.clamp(1e-5, 1 - 1e-5), (1, B, 2), 15)
    C = mC.VAEHyperbolicRNASeq(torch.Size([300]), 5, 1.0, 32, 1e-3, 0.5)
    xr = torch.randn(B, 300)
    run("C", C, {"rnaseq": xr}, xr, (1, B, 5), 16)
    D1 = mD.VAE(torch.Size([300]), 32, 2, 1.0, 2.0, "learned", 1e-3, 0.5, "logmap0_analytic", torch.nn.GELU, "none", "MSE")
    run("OneB", D1, (xr, None), xr, (B, 2), 17)
    D2 = mD.VAE(torch.Size([300]), 32, 3, 0.5, 1.0, "learned", 1e-3, 1.0, "log_prob", torch.nn.GELU, "none", "MSE")
    run("OneB_log_prob", D2, (xr, None), xr, (B, 3), 18)
    return out


def make_riemannian():
    """a-7: run the reference's OWN distributions/old_pvae_riemannian_normal.py:12-52 (over the pvae_min / geoopt_min
    shims) and record what it drew and computed: the clamp of sigma to [0.1, 7], the (alpha, r) of its rsample (captured
    by wrapping the two sampler calls on the instance; the reference code itself runs verbatim), z = expmap_polar(mu,
    alpha, r), the implicit-reparameterisation gradients, log_prob and its gradients, and the log-normaliser."""
    rl.load()
    RN = importlib.import_module("hyperbolic_vae.distributions.old_pvae_riemannian_normal")
    import pvae.manifolds as PM

    cases = []
    for c in (0.5, 1.0, 2.0):
        for D in (2, 5, 10):
            torch.manual_seed(4200 + int(c * 10) + D)
            ball = PM.PoincareBall(D, c)
            B = 24
            mu = ball.expmap0(torch.randn(B, D) * 0.6).detach().requires_grad_(True)
            # (pvae's fixed-hull ARS raises "initial anchor points must span mode" for sigma >~ 2 at D = 10: sampled rows stay
            # below that; the upper clamp is exercised through log_prob, which needs no sampling)
            scale = torch.rand(B, 1) * 1.5 + 0.3
            scale[0] = 0.05   # below the clamp -> 0.1
            scale.requires_grad_(True)
            q = RN.RiemannianNormal(mu, scale, ball)
            drawn = {}
            d_sample, r_rsample = q.direction.sample, q.radius.rsample

            def rec_alpha(shape, _f=d_sample):
                drawn["alpha"] = _f(shape)
                return drawn["alpha"]

            def rec_radius(shape=torch.Size(), _f=r_rsample):
                drawn["r"] = _f(shape)
                return drawn["r"]

            q.direction.sample, q.radius.rsample = rec_alpha, rec_radius
            z = q.rsample(torch.Size([1]))
            g = torch.Generator().manual_seed(7)
            gz_up = torch.randn(z.shape, generator=g)
            z.backward(gz_up)
            rec = dict(c_ctor=c, c=float(ball.c), D=D, B=B, mu=plain(mu), scale=plain(scale), scale_clamped=plain(q.scale),
                       alpha=plain(drawn["alpha"]), r=plain(drawn["r"]), z=plain(z), gz_up=gz_up, gmu_z=plain(mu.grad),
                       gscale_z=plain(scale.grad), logZ=plain(q.radius.log_normalizer))
            mu.grad = None
            scale.grad = None
            zz = plain(z).clone().requires_grad_(True)
            scale2 = plain(scale).clone()
            scale2[1] = 9.0   # above the clamp -> 7.0
            scale2.requires_grad_(True)
            q2 = RN.RiemannianNormal(mu, scale2, ball)
            lp = q2.log_prob(zz)
            glp = torch.randn(lp.shape, generator=g)
            lp.backward(glp)
            rec.update(scale_lp=plain(scale2), scale_lp_clamped=plain(q2.scale), logZ_lp=plain(q2.radius.log_normalizer),
                       log_prob=plain(lp), glp_up=glp, gmu_lp=plain(mu.grad), gscale_lp=plain(scale2.grad), gz_lp=plain(zz.grad))
            # prior-style construction: origin loc, one scalar sigma (pvae prior_iso); log_prob of the same z
            mu.grad = None
            scale.grad = None
            p0 = RN.RiemannianNormal(ball.zero, torch.full((1, 1), 1.3), ball)
            rec["log_prob_prior"] = plain(p0.log_prob(plain(z)))
            rec["prior_sigma"] = 1.3
            cases.append(rec)
    return cases


def make_optim():
    """f-1: the optimizer the reference configures (models/vae_hyperbolic.py:235-243: geoopt.optim.RiemannianAdam over ALL
    parameters, Euclidean and ManifoldParameter alike) stepped 3 times on model B (Mobius encoder + gyroplane decoder:
    the gyroplane `points` are a ManifoldParameter) with seeded noise: initial state_dict, the batch, the noise of every
    step, and the state_dict + optimizer moments after every step."""
    rl.load()
    mB = importlib.import_module("hyperbolic_vae.models.vae_hyperbolic")
    from torch.distributions.utils import _standard_normal

    torch.manual_seed(42)
    B = 16
    model = mB.VAEHyperbolicExperiment((1, 16, 16), 2, 1.0, "mobius", "geoopt_gyroplane", loss_recon="mse")
    opt = model.configure_optimizers()["optimizer"]
    x = torch.rand(B, 1, 16, 16)
    rec = dict(x=x, lr=opt.param_groups[0]["lr"], betas=opt.param_groups[0]["betas"], eps_adam=opt.param_groups[0]["eps"],
               weight_decay=opt.param_groups[0]["weight_decay"],
               state_dict0={k: plain(v) for k, v in model.state_dict().items()}, steps=[])
    names = {id(p): k for k, p in model.named_parameters()}
    for it in range(3):
        seed = 900 + it
        torch.manual_seed(seed)
        eps = _standard_normal(torch.Size((1, B, 2)), dtype=torch.float32, device=torch.device("cpu"))
        torch.manual_seed(seed)
        opt.zero_grad()
        loss = model.loss((x, None))["loss_total"]
        loss.backward()
        grads = {k: plain(p.grad) for k, p in model.named_parameters() if p.grad is not None}
        opt.step()
        rec["steps"].append(dict(eps=eps, loss=plain(loss), grads=grads,
                                 state_dict={k: plain(v) for k, v in model.state_dict().items()},
                                 exp_avg={names[id(p)]: plain(st["exp_avg"]) for p, st in opt.state.items()},
                                 exp_avg_sq={names[id(p)]: plain(st["exp_avg_sq"]) for p, st in opt.state.items()}))
    return rec


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only in ("", "ops"):
        torch.save(make_ops(), os.path.join(HERE, "ops_golden.pt"))
    if only in ("", "models"):
        torch.save(make_models(), os.path.join(HERE, "models_golden.pt"))
    if only in ("", "riemannian"):
        torch.save(make_riemannian(), os.path.join(HERE, "riemannian_golden.pt"))
    if only in ("", "optim"):
        torch.save(make_optim(), os.path.join(HERE, "optim_golden.pt"))
    man = {}
    for fn in ("ops_golden.pt", "models_golden.pt", "riemannian_golden.pt", "optim_golden.pt"):
        if not os.path.exists(os.path.join(HERE, fn)):
            continue
        with open(os.path.join(HERE, fn), "rb") as f:
            man[fn] = {"sha256": hashlib.sha256(f.read()).hexdigest(), "bytes": os.path.getsize(os.path.join(HERE, fn))}
    man["generator"] = "tests/golden/make_golden.py (reference files executed verbatim over oracle shims)"
    man["torch"] = torch.__version__
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1)
    print(json.dumps(man, indent=1))


if __name__ == "__main__":
    main()
