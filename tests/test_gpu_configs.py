"""GPU parity at the BASELINE.json sizes, through the call a user makes (hvae.train.TrainStep, CUDA-graph replay): the
whole train step - forward + loss + backward - of
  config 1: models A and B ("Mobius encoder + gyroplane decoder + MSE") on 128 MNIST-shaped rows, D = 2, c = 1;
  config 2: the pvae-MNIST graph, batch 4096, hidden 600, D = 10, RiemannianNormal with injected (alpha, r), fp32-accurate
            tensor-core trunk - exactly the step bench.py times;
  config 3: the RNA-seq model on (1024, 20000) z-scored counts, D = 5, hidden 100 (script default) and 512
against the oracle port of the reference's model code (oracle/ref_port.py, pinned to the reference's own files by
tests/test_oracle_golden.py) in float32 and, as the tie-breaker, float64 under float32 clamp semantics.  Same weights
(the oracle's state_dict is loaded as is), same inputs, same injected noise.  Loss terms and EVERY parameter gradient."""
import pytest
import torch

from util_parity import assert_parity

pytestmark = pytest.mark.gpu


def _oracle_run(make, sd, x, dtype, **noise):
    from oracle.geoopt_min.manifolds.stereographic import math as gmath
    from oracle.geoopt_min.manifolds.stereographic.manifold import PoincareBall as OBall

    m = make()
    m.load_state_dict(sd)
    if dtype == torch.float64:
        c32 = {id(b): float(b.c) for b in m.modules() if isinstance(b, OBall)}
        m = m.double()
        for b in m.modules():
            if isinstance(b, OBall):  # keep the fp32 curvature VALUE
                b.isp_c.data = torch.log(torch.expm1(torch.tensor(c32[id(b)], dtype=torch.float64)))
    with gmath.fp32_semantics(dtype == torch.float64):
        L = m.loss(x.to(dtype), **{k: v.to(dtype) for k, v in noise.items()})
        L["loss_total"].backward()
    return {k: v.detach() for k, v in L.items()}, {k: p.grad for k, p in m.named_parameters() if p.grad is not None}


def _cuda_run(model, sd, x, use_graph=True, **noise):
    """Through hvae.train.TrainStep: 3 eager warm-up steps, capture, replay; gradients are read from the flat bucket."""
    from hvae.train import TrainStep

    sd_c = {k: v for k, v in sd.items() if not k.endswith("manifold.dim")}  # pvae's PoincareBall(dim, c) buffer
    missing, unexpected = model.load_state_dict(sd_c, strict=False)
    assert not unexpected and all("isp_c" in k for k in missing), (missing, unexpected)
    model = model.cuda()
    step = TrainStep(model, x.cuda(), use_graph=use_graph, **{k: v.cuda() for k, v in noise.items()})
    assert (step.graph is not None) == use_graph
    loss = step.run()
    loss = step.run()   # a second replay: the step is re-entrant (bucket zeroed, same noise -> same numbers)
    torch.cuda.synchronize()
    with torch.no_grad():
        L = model.loss(x.cuda(), **{k: v.cuda() for k, v in noise.items()})
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None and p.requires_grad}
    return loss, L, grads


def _compare(tag, loss, L, grads, o32, o64, loss_rtol=3e-5, loss_atol=1e-4, grad_rtol=5e-5):
    # (loss terms: cancelling batch sums, fp32 summation order alone moves them by ~1e-5; see tests/test_gpu_models.py)
    L32, G32 = o32
    L64, G64 = o64
    assert_parity(loss.reshape(1), L32["loss_total"].reshape(1), L64["loss_total"].reshape(1), what=tag + " loss_total (TrainStep)",
                  rtol=loss_rtol, atol=loss_atol, row_relative=False, slack_mult=2.0)
    for k in L32:
        assert_parity(L[k].reshape(1), L32[k].reshape(1), L64[k].reshape(1), what="%s %s" % (tag, k), rtol=loss_rtol, atol=loss_atol,
                      row_relative=False, slack_mult=2.0)
    assert set(grads) == set(G32), set(grads) ^ set(G32)
    for k in G32:
        assert_parity(grads[k], G32[k], G64[k], what="%s grad %s" % (tag, k), rtol=grad_rtol, atol=1e-7, norm_relative=True,
                      slack_mult=2.0)


@pytest.mark.parametrize("which", ["A", "B"])
def test_cfg1_step(which):
    from hvae import models as HM
    from oracle import ref_port as R

    torch.manual_seed(42)
    B, D = 128, 2
    if which == "A":
        make_o = lambda: R.ModelA(torch.Size([1, 28, 28]), D, 1.0, 1.0, 1.0)          # noqa: E731
        make_c = lambda: HM.ModelA(torch.Size([1, 28, 28]), D, 1.0, 1.0, 1.0)         # noqa: E731
        x = torch.rand(B, 1, 28, 28)
    else:
        make_o = lambda: R.ModelB((1, 32, 32), D, 1.0, "mobius", "geoopt_gyroplane", 1.0, "mse")    # noqa: E731
        make_c = lambda: HM.ModelB((1, 32, 32), D, 1.0, "mobius", "geoopt_gyroplane", 1.0, "mse")   # noqa: E731
        x = torch.rand(B, 1, 32, 32)
    sd = {k: v.clone() for k, v in make_o().state_dict().items()}
    eps = torch.randn(1, B, D)
    o32 = _oracle_run(make_o, sd, x, torch.float32, eps=eps)
    o64 = _oracle_run(make_o, sd, x, torch.float64, eps=eps)
    loss, L, grads = _cuda_run(make_c(), sd, x, eps=eps)
    # model B's losses are batch SUMS of O(1e4): the absolute floor scales with them
    _compare("cfg1 model " + which, loss, L, grads, o32, o64, loss_atol=1e-4 if which == "A" else 2e-2)


@pytest.mark.parametrize("use_graph", [True, False])
def test_cfg2_step_full_size(use_graph):
    """The step bench.py times: B = 4096, H = 600, D = 10, fp16 two-piece tensor-core trunk, fused heads, CUDA graph."""
    from hvae import models as HM
    from hvae import ops
    from oracle import ref_port as R

    torch.manual_seed(7)
    B, D, H = 4096, 10, 600
    make_o = lambda: R.PvaeMnist(latent_dim=D, hidden_dim=H, c=1.0)   # noqa: E731
    o = make_o()
    sd = {k: v.clone() for k, v in o.state_dict().items()}
    x = torch.rand(B, 1, 28, 28).clamp(1e-5, 1 - 1e-5)
    alpha = torch.randn(1, B, D)
    alpha = alpha / alpha.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        mu_, sg_ = o.encode(x)
        r = R.RiemannianNormal(mu_, sg_, o.manifold).radius.sample(torch.Size([1]))   # the reference's own ARS radii
    o32 = _oracle_run(make_o, sd, x, torch.float32, alpha=alpha, r=r)
    o64 = _oracle_run(make_o, sd, x, torch.float64, alpha=alpha, r=r)
    assert ops.get_trunk_mode() == "x2"   # the fp32-accurate tcgen05 trunk (fp16 two-piece path) is what bench.py runs
    loss, L, grads = _cuda_run(HM.PvaeMnist(latent_dim=D, hidden_dim=H, c=1.0), sd, x, use_graph=use_graph, alpha=alpha, r=r)
    # losses are batch sums of O(2e6)
    _compare("cfg2 B=4096 H=600", loss, L, grads, o32, o64, loss_atol=0.5)


@pytest.mark.parametrize("H", [100, 512])
def test_cfg3_step_full_size(H):
    from hvae import models as HM
    from oracle import ref_port as R

    torch.manual_seed(3)
    B, G, D = 1024, 20000, 5
    counts = torch.poisson(torch.full((B, G), 100.0))
    x = (counts - counts.mean(0, keepdim=True)) / counts.std(0, keepdim=True)   # jerby_arnon.py:102-104 z-score
    make_o = lambda: R.ModelC(torch.Size([G]), D, 1.0, H, 0.5)    # noqa: E731
    sd = {k: v.clone() for k, v in make_o().state_dict().items()}
    eps = torch.randn(1, B, D)
    o32 = _oracle_run(make_o, sd, x, torch.float32, eps=eps)
    o64 = _oracle_run(make_o, sd, x, torch.float64, eps=eps)
    loss, L, grads = _cuda_run(HM.ModelC(torch.Size([G]), D, 1.0, H, 0.5), sd, x, eps=eps)
    _compare("cfg3 G=20000 H=%d" % H, loss, L, grads, o32, o64, loss_atol=1e-2)
