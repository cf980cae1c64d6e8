"""One small, ragged invocation of every kernel family (forward + backward) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import hvae
from hvae import layers, ops
from hvae.distributions import RiemannianNormal, WrappedNormal

dev = "cuda"
torch.manual_seed(0)
ball = hvae.PoincareBall(0.7)
c = ball.c_value
for D in (1, 2, 3, 5, 8, 10, 33, 64, 100, 300, 777):
    B = 37
    u = (torch.randn(B, D, device=dev) * 0.3).requires_grad_(True)
    y = ball.expmap0(u)
    (ball.logmap0(y).sum() + ball.mobius_add(y, y.flip(0)).sum() + ball.dist(y, y.flip(0)).sum()
     + ball.expmap(y, u * 0.1).sum() + ball.logmap(y, y.flip(0)).sum()).backward()
    mu = ball.expmap0(torch.randn(B, D, device=dev) * 0.2).detach().requires_grad_(True)
    sg = (torch.rand(B, D, device=dev) + 0.3).requires_grad_(True)
    q = WrappedNormal(mu, sg, ball)
    z = q.rsample(torch.Size([3]))
    (q.log_prob(z).sum() + WrappedNormal.origin_prior(D, 1.3, ball, device=dev).log_prob(z).sum()).backward()
    mu2 = mu.detach().requires_grad_(True)
    sg2 = sg.detach().requires_grad_(True)
    zz, kl = ops.latent_head(mu2, sg2, torch.randn(B, D, device=dev), 1.0, c)
    (zz.sum() + kl.sum()).backward()
for D, P, B in ((2, 16, 5), (5, 100, 130), (10, 600, 77), (33, 50, 9), (64, 129, 200)):
    for lay in (layers.Distance2PoincareHyperplanes(D, P, ball=ball), layers.GeodesicLayer(D, P, ball)):
        lay = lay.cuda()
        x = ball.expmap0(torch.randn(B, D, device=dev) * 0.3).detach().requires_grad_(True)
        lay(x).sum().backward()
for F, P, B in ((48, 2, 33), (600, 10, 77), (130, 40, 9), (1000, 5, 3), (64, 100, 50)):
    lay = layers.MobiusLayer(F, P, ball).cuda()
    x = (torch.randn(B, F, device=dev) * 0.5).requires_grad_(True)
    lay(x).sum().backward()
for D in (1, 2, 10, 32):
    B = 50
    mu = ball.expmap0(torch.randn(B, D, device=dev) * 0.2).detach().requires_grad_(True)
    sg = (torch.rand(B, 1, device=dev) + 0.4).requires_grad_(True)
    q = RiemannianNormal(mu, sg, ball)
    z = q.rsample(torch.Size([2]))
    p = RiemannianNormal(torch.zeros(1, D, device=dev), torch.ones(1, 1, device=dev), ball)
    (q.log_prob(z).sum() + q.kl_mc(z, p).sum()).backward()
ops.set_gemm_mode("bf16")
x = ball.expmap0(torch.randn(300, 72, device=dev) * 0.1)
M = torch.randn(200, 72, device=dev) * 0.1
ops.mobius_matvec_tc_fwd(x, M, c)
ops.mobius_matvec_tc(x, M, c)
ops.gyroplane_tc_fwd(x, ball.expmap0(M), None, c, ops.GYRO_SIGNED)
ops.set_gemm_mode("fp32")
torch.cuda.synchronize()
print("sanitize smoke ok; launches:", hvae._cabi.launch_count)
