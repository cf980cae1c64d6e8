"""GPU parity of the fp32-accurate tensor-core GEMM (three-way bf16 split, six piece products) that carries the trunk
dense layers: against a float64 product its error must be that of an fp32 FMA GEMM (torch / cuBLAS fp32 with TF32 off),
i.e. well inside the 1e-5 budget of the fp32 mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max())


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
