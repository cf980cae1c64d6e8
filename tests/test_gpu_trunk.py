"""GPU parity of the fp32-accurate tensor-core GEMM (three-way bf16 split, six piece products) that carries the trunk
dense layers: against a float64 product its error must be that of an fp32 FMA GEMM (torch / cuBLAS fp32 with TF32 off),
i.e. well inside the 1e-5 budget of the fp32 mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    return float((a.detach().double() - ref.detach()).abs().max() / ref.detach().abs().max())


# (M, N, K, a_trans, b_trans): trunk shapes of config 2 fwd / bwd (the last two run split-K), ragged and odd sizes
CASES = [
    (4096, 600, 784, False, False),
    (4096, 784, 600, False, False),
    (4096, 600, 784, False, True),
    (600, 784, 4096, True, True),
    (784, 600, 4096, True, True),
    (1000, 300, 100, False, False),
    (257, 601, 77, False, True),
    (130, 65, 1000, True, False),
    (128, 64, 64, False, False),
]


@pytest.mark.parametrize("M,N,K,ta,tb", CASES)
def test_gemm_x3_matches_fp32_accuracy(M, N, K, ta, tb):
    from hvae import ops

    torch.manual_seed(M + N + K)
    torch.backends.cuda.matmul.allow_tf32 = False
    A = torch.randn(M, K, device="cuda") * torch.rand(M, 1, device="cuda").mul(4).exp()  # rows on different scales
    B = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    ref = A.double() @ B.double().t() + bias.double()
    As = A.t().contiguous() if ta else A
    Bs = B.t().contiguous() if tb else B
    out = ops.gemm_x3(As, ta, Bs, tb, bias, False)
    torch.cuda.synchronize()
    e_x3 = _rel(out, ref)
    e_t = _rel(A @ B.t() + bias, ref)
    assert e_x3 < 2e-6, e_x3
    assert e_x3 < 3.0 * e_t + 2e-7, (e_x3, e_t)
    # element-wise: every output within fp32 accumulation error of its own row scale
    row_scale = (A.double().abs() @ B.double().abs().t())
    assert float(((out.double() - ref).abs() / row_scale).max()) < 1e-6
    # ReLU epilogue
    outr = ops.gemm_x3(As, ta, Bs, tb, bias, True)
    assert torch.equal(outr, out.clamp_min(0))


# pre-split operands, every combination of K-major / MN-major reads (the backward GEMMs of a dense layer use the
# forward's splits MN-major instead of transposing), with ragged sizes and split-K shapes
@pytest.mark.parametrize("M,N,K", [(4096, 784, 600), (600, 784, 4096), (257, 130, 77), (128, 64, 64), (1000, 601, 333)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_x3s_operand_majors(M, N, K, a_mn, b_mn):
    from hvae import ops

    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    ref = A.double() @ B.double().t() + bias.double()
    As = ops.split3(A.t().contiguous() if a_mn else A)
    Bs = ops.split3(B.t().contiguous() if b_mn else B)
    out = ops.gemm_x3s(As, a_mn, Bs, b_mn, bias, False, M, N, K)
    torch.cuda.synchronize()
    e = _rel(out, ref)
    assert e < 2e-6, e


def test_split3_is_exact_to_24_bits():
    from hvae import ops

    torch.manual_seed(0)
    x = torch.randn(300, 100, device="cuda") * torch.rand(300, 1, device="cuda").mul(20).sub(10).exp()
    s = ops.split3(x).float().view(300, 3, 128)
    assert torch.equal(s[:, :, 100:], torch.zeros_like(s[:, :, 100:]))
    rec = s[:, 0, :100].double() + s[:, 1, :100].double() + s[:, 2, :100].double()
    assert float(((rec - x.double()).abs() / x.double().abs()).max()) < 2.0 ** -23


def test_linear_layer_forward_backward():
    from hvae import layers, ops

    torch.manual_seed(0)
    lin = layers.Linear(784, 600).cuda()
    x = torch.randn(4096, 784, device="cuda", requires_grad=True)
    gy = torch.randn(4096, 600, device="cuda")
    assert ops.trunk_x3_eligible(x, lin.weight)
    y = lin(x)
    y.backward(gy)
    xd = x.detach().double().requires_grad_(True)
    Wd, bd = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    yd = torch.nn.functional.linear(xd, Wd, bd)
    yd.backward(gy.double())
    for got, ref in ((y, yd), (x.grad, xd.grad), (lin.weight.grad, Wd.grad), (lin.bias.grad, bd.grad)):
        assert _rel(got, ref.detach()) < 2e-6
    # state_dict compatibility with torch.nn.Linear and the torch fallback
    ref_lin = torch.nn.Linear(784, 600).cuda()
    ref_lin.load_state_dict(lin.state_dict())
    ops.set_trunk_mode("torch")
    try:
        y_t = lin(x)
    finally:
        ops.set_trunk_mode("x3")
    assert torch.allclose(y_t, ref_lin(x))
    assert _rel(y, y_t.detach().double()) < 2e-6
    # small shapes defer to torch
    small = layers.Linear(16, 8).cuda()
    xs = torch.randn(4, 16, device="cuda")
    assert torch.equal(small(xs), torch.nn.functional.linear(xs, small.weight, small.bias))


@pytest.mark.parametrize("S,B,N", [(1, 4096, 784), (3, 17, 10), (2, 5, 1)])
def test_bernoulli_nll_rows(S, B, N):
    from hvae import ops

    torch.manual_seed(S + B + N)
    logits = (torch.randn(S, B, N, device="cuda") * 6).requires_grad_(True)  # includes |l| > 15 (saturated sigmoid)
    x = torch.rand(B, N, device="cuda")
    g = torch.randn(S, B, device="cuda")
    out = ops.bernoulli_nll_rows(logits, x)
    out.backward(g)
    ld = logits.detach().double().requires_grad_(True)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(ld, x.double().expand(S, B, N), reduction="none").sum(-1)
    ref.backward(g.double())
    assert float(((out.double() - ref).abs() / ref.abs()).max()) < 1e-5   # 1e-5 relative, fp32 sums over N terms
    gscale = ld.grad.abs().amax(dim=-1, keepdim=True).clamp_min(1e-30)
    assert float(((logits.grad.double() - ld.grad).abs() / gscale).max()) < 1e-5


def test_cfg2_step_fused_heads_match_torch_heads():
    """config-2 shapes: the step with the tensor-core trunk + fused loss head against the same model on torch's
    fp32 Linear / BCE kernels (same noise): loss within 1e-5; parameter gradients within 1e-4 in norm (two fp32
    GEMM implementations differ by ~1e-6, which flips the ReLU mask of the few pre-activations that close to zero)."""
    import hvae
    from hvae import models, ops

    torch.manual_seed(0)
    B = 4096
    x = torch.rand(B, 1, 28, 28, device="cuda")
    model = models.PvaeMnist().cuda()
    alpha = torch.randn(1, B, 10, device="cuda")
    mu, sigma = model.encode(x)
    from hvae.distributions.riemannian_normal import RiemannianNormal

    r = RiemannianNormal(mu, sigma, model.manifold).radius.sample(torch.Size([1])).detach()
    res = {}
    for mode in ("ours", "torch"):
        model.zero_grad(set_to_none=True)
        model.fused = mode == "ours"
        ops.set_trunk_mode("x3" if mode == "ours" else "torch")
        try:
            out = model.loss(x, alpha=alpha, r=r)
            out["loss_total"].backward()
        finally:
            ops.set_trunk_mode("x3")
            model.fused = True
        res[mode] = (out["loss_total"].detach().double(), {n: p.grad.detach().double().clone() for n, p in model.named_parameters() if p.grad is not None})
    la, lb = res["ours"][0], res["torch"][0]
    assert abs(float(la - lb)) / abs(float(lb)) < 1e-5
    for n, gb in res["torch"][1].items():
        ga = res["ours"][1][n]
        assert float((ga - gb).norm() / gb.norm().clamp_min(1e-30)) < 1e-4, n


@pytest.mark.parametrize("R,C", [(4096, 784), (1000, 77), (5, 3), (33, 600)])
def test_colsum(R, C):
    from hvae import ops

    torch.manual_seed(R + C)
    x = torch.randn(R, C, device="cuda")
    ref = x.double().sum(0)
    out = ops.colsum(x)
    assert float((out.double() - ref).abs().max()) < 1e-5 * float(x.double().abs().sum(0).max())
