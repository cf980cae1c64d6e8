"""GPU parity of the fused Riemannian Adam (f-1, csrc/riemannian_adam.cu, hvae.optim.RiemannianAdam) against
(a) the oracle restatement of geoopt.optim.RiemannianAdam on injected gradients, Euclidean tensors and Poincare-ball
    rows (incl. rows pushed onto the projection radius), several steps, with and without weight decay, and
(b) the golden fixture minted by stepping the optimizer the reference configures (models/vae_hyperbolic.py:235-243)
    on model B three times (tests/golden/make_golden.py::make_optim)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wd", [0.0, 0.01])
@pytest.mark.parametrize("c", [1.0, 0.5])
def test_riemannian_adam_matches_oracle(c, wd):
    import hvae
    from hvae.optim import RiemannianAdam
    from oracle.geoopt_min import ManifoldParameter as OMP
    from oracle.geoopt_min import PoincareBall as OBall
    from oracle.geoopt_min.optim import RiemannianAdam as ORA

    torch.manual_seed(int(c * 10) + int(wd * 1000))
    ob, hb = OBall(c=c), hvae.PoincareBall(c)
    shapes_e = [(3,), (17, 5), (600, 784), (1,)]
    shapes_m = [(16, 2), (100, 5), (7, 64), (33, 10)]
    init_e = [torch.randn(s) for s in shapes_e]
    init_m = [ob.expmap0(torch.randn(s) * 0.7 / s[-1] ** 0.5).detach() for s in shapes_m]   # |x| ~ 0.6 / sqrt(c)
    init_m[0][1] = ob.expmap0(torch.randn(2) * 50.0)     # a row on the projection radius
    o_params = [torch.nn.Parameter(t.clone()) for t in init_e] + [OMP(t.clone(), manifold=ob) for t in init_m]
    c_params = [torch.nn.Parameter(t.clone().cuda()) for t in init_e] + [hvae.ManifoldParameter(t.clone().cuda(), manifold=hb) for t in init_m]
    oo = ORA(o_params, lr=1e-2, weight_decay=wd)
    co = RiemannianAdam(c_params, lr=1e-2, weight_decay=wd)
    for step in range(5):
        prev = [po.detach().clone() for po in o_params]
        for po, pc in zip(o_params, c_params):
            g = torch.randn(po.shape) * (10.0 if step == 2 else 1.0)   # a violent step: retraction clips rows
            po.grad = g.clone()
            pc.grad = g.clone().cuda()
        oo.step()
        co.step()
        torch.cuda.synchronize()
        for i, (po, pc) in enumerate(zip(o_params, c_params)):
            # rows near the projection radius: lambda = 2 / (1 - c|x|^2) amplifies fp32 rounding of |x|^2 by kappa = 1 / (1 - c|x|^2)
            # (lambda^2 in egrad2rgrad and in the transported moment): per-row tolerance 5e-5 * kappa^2, capped at 5 %
            if i >= len(shapes_e):
                kap = 1.0 / (1.0 - c * torch.maximum(po.detach().pow(2).sum(-1, keepdim=True), prev[i].pow(2).sum(-1, keepdim=True))).clamp_min(4e-3)
                rt = (5e-5 * kap * kap).clamp(max=5e-2)
            else:
                rt = torch.tensor(2e-5)
            d = (pc.detach().cpu() - po.detach()).abs()
            assert bool((d <= rt * po.detach().abs().amax(-1, keepdim=True).clamp_min(1e-3) + 1e-6).all()), (step, i, "param", float(d.max()))
            for key in ("exp_avg", "exp_avg_sq"):
                a, b = co.state[pc][key].cpu(), oo.state[po][key]
                sb = b.abs().amax(-1, keepdim=True).clamp_min(1e-12) if b.dim() > 1 else b.abs().max().clamp_min(1e-12)
                assert bool(((a - b).abs() <= 2.5 * rt * sb + 1e-12).all()), (step, i, key, float((a - b).abs().max()), float(b.abs().max()))
            if i >= len(shapes_e):   # still on the ball
                assert float(pc.detach().norm(dim=-1).max()) <= (1 - 4e-3) / c ** 0.5 * (1 + 1e-6)
    assert co.state[c_params[0]]["step"] == 5


def test_riemannian_adam_golden_model_b():
    """Three optimizer steps of model B from the reference's own initial weights, gradients taken on the GPU path."""
    import os

    from hvae import models as HM
    from hvae.optim import RiemannianAdam

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optim_golden.pt"), weights_only=False)
    model = HM.ModelB((1, 16, 16), 2, 1.0, "mobius", "geoopt_gyroplane", 1.0, "mse")
    missing, unexpected = model.load_state_dict(g["state_dict0"], strict=False)
    assert not unexpected and all("isp_c" in k for k in missing)
    model = model.cuda()
    opt = RiemannianAdam(model.parameters(), lr=g["lr"], betas=g["betas"], eps=g["eps_adam"], weight_decay=g["weight_decay"])
    x = g["x"].cuda()
    for it, st in enumerate(g["steps"]):
        opt.zero_grad(set_to_none=False)
        loss = model.loss(x, eps=st["eps"].cuda())["loss_total"]
        loss.backward()
        assert abs(float(loss) - float(st["loss"])) <= 2e-5 * abs(float(st["loss"])), (it, float(loss), float(st["loss"]))
        opt.step()
        sd = model.state_dict()
        for k, v in st["state_dict"].items():
            if "isp_c" in k:
                continue
            # Adam normalises the update to ~lr per element: a parameter moves by <= lr per step whatever its gradient scale,
            # so the comparison is absolute, at 2 % of one step's movement (first steps divide tiny m by tiny sqrt(v))
            assert float((sd[k].cpu() - v).abs().max()) <= 0.02 * g["lr"] * (it + 1) + 1e-6 * float(v.abs().max()), (it, k)
        names = {id(p): k for k, p in model.named_parameters()}
        for p, s in opt.state.items():
            k = names[id(p)]
            b = st["exp_avg"][k]
            assert float((s["exp_avg"].cpu() - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-9, (it, k, "exp_avg")


def test_train_step_with_fused_optimizer_in_graph():
    """TrainStep(optimizer=...) captures forward + backward + the fused Adam launch in one CUDA graph; replays advance
    the step count and move the parameters like eager steps do."""
    from hvae import models as HM
    from hvae.optim import RiemannianAdam
    from hvae.train import TrainStep

    torch.manual_seed(0)
    x = torch.rand(64, 1, 16, 16, device="cuda")
    eps = torch.randn(1, 64, 2, device="cuda")

    def run(use_graph):
        torch.manual_seed(1)
        m = HM.ModelB((1, 16, 16), 2, 1.0, "mobius", "geoopt_gyroplane", 1.0, "mse").cuda()
        opt = RiemannianAdam(m.parameters(), lr=1e-3)
        ts = TrainStep(m, x, use_graph=use_graph, optimizer=opt, eps=eps)
        for _ in range(4):
            ts.run()
        torch.cuda.synchronize()
        return {k: v.detach().clone() for k, v in m.state_dict().items()}, ts

    sd_e, _ = run(False)
    sd_g, ts = run(True)
    assert ts.graph is not None
    for k in sd_e:
        assert float((sd_e[k] - sd_g[k]).abs().max()) <= 1e-5 * float(sd_e[k].abs().max()) + 1e-7, k
