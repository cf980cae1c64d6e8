"""GPU: the in-kernel noise streams (Philox): graph-replay safety (a host-side offset baked into a captured
graph replays the same radii for ever), the data-parallel shard rule (a rank that owns rows [lo, lo + B_local)
of a global batch draws exactly the rows' single-GPU numbers) and the direction sampler's distribution."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graph_replays_draw_fresh_noise():
    from hvae import ops

    sig = torch.full((4096,), 0.8, device="cuda")
    ops.hradius_sample(sig, 1, 10, 1.0)                 # creates the device counter outside any capture
    ops.sphere_sample(1, 4096, 10, "cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.hradius_sample(sig, 1, 10, 1.0)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        r = ops.hradius_sample(sig, 1, 10, 1.0)
        a = ops.sphere_sample(1, 4096, 10, "cuda")
    g.replay()
    torch.cuda.synchronize()
    r1, a1 = r.clone(), a.clone()
    g.replay()
    torch.cuda.synchronize()
    assert not torch.equal(r1, r) and float((r1 - r).abs().mean()) > 0.05
    assert not torch.equal(a1, a)


def test_pvae_step_graph_replays_differ():
    """The default path: TrainStep(use_graph=True) on the config-2 model, no counter installed by the caller."""
    from hvae import models as HM
    from hvae.train import TrainStep

    torch.manual_seed(0)
    m = HM.PvaeMnist(latent_dim=10, hidden_dim=64).cuda()
    x = torch.rand(256, 1, 28, 28, device="cuda").clamp(1e-5, 1 - 1e-5)
    ts = TrainStep(m, x, use_graph=True)
    assert ts.graph is not None
    l1 = float(ts.run())
    l2 = float(ts.run())
    assert l1 != l2 and abs(l1 - l2) > 1e-6 * abs(l1)   # same weights, same batch: only the noise differs


def test_noise_shard_reproduces_the_global_stream():
    from hvae import ops

    torch.manual_seed(123)
    Bg, D, S = 512, 10, 2
    sig = torch.rand(Bg, device="cuda") + 0.4
    try:
        ops.set_noise_shard(0, None)
        ops.philox_counter(sig.device)
        ops.reset_noise()
        a_g = ops.sphere_sample(S, Bg, D, "cuda")
        r_g = ops.hradius_sample(sig, S, D, 1.0)
        a_g2 = ops.sphere_sample(S, Bg, D, "cuda")       # second call of the "step": counters have advanced
        for lo, hi in ((0, 256), (256, 512), (384, 512)):
            ops.reset_noise()
            ops.set_noise_shard(lo, Bg)
            a_s = ops.sphere_sample(S, hi - lo, D, "cuda")
            r_s = ops.hradius_sample(sig[lo:hi], S, D, 1.0)
            a_s2 = ops.sphere_sample(S, hi - lo, D, "cuda")
            assert torch.equal(a_s, a_g[:, lo:hi]) and torch.equal(r_s, r_g[:, lo:hi]) and torch.equal(a_s2, a_g2[:, lo:hi])
    finally:
        ops.set_noise_shard(0, None)


@pytest.mark.parametrize("D", [2, 3, 10, 65])
def test_sphere_sample_distribution(D):
    from hvae import ops

    N = 200_000
    a = ops.sphere_sample(1, N, D, "cuda", seed=7, offset=0)[0].double().cpu()
    assert float((a.norm(dim=-1) - 1).abs().max()) < 1e-5
    # coordinates: mean 0 +- 5 sigma, second moment 1/D
    se = math.sqrt(1.0 / D / N)
    assert float(a.mean(0).abs().max()) < 5 * se
    assert float((a.pow(2).mean(0) - 1.0 / D).abs().max()) < 6 * math.sqrt(2.0 / N) / D * 2
    if D == 2:   # the angle is uniform: Kolmogorov-Smirnov against U(-pi, pi)
        th = torch.atan2(a[:, 1], a[:, 0]).sort().values
        F = (th + math.pi) / (2 * math.pi)
        i = torch.arange(1, N + 1, dtype=torch.float64)
        Dn = torch.maximum((i / N - F).abs().max(), (F - (i - 1) / N).abs().max()).item()
        assert Dn < 1.95 / math.sqrt(N)
    a2 = ops.sphere_sample(1, N, D, "cuda", seed=7, offset=0)[0].double().cpu()
    assert torch.equal(a, a2)
