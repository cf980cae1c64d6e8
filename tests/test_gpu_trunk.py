"""GPU parity of the fp32-accurate tensor-core GEMMs that carry the trunk dense layers - the fp16 two-piece path with
power-of-two row scales (three piece products, the default) and the three-way bf16 split (six piece products): against a
float64 product the error must be that of an fp32 FMA GEMM (torch / cuBLAS fp32 with TF32 off), i.e. well inside the
1e-5 budget of the fp32 mode."""
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


def _split2h(x, want_rows=True, want_t=False):
    from hvae import ops

    return ops.split2h_both(x, want_rows, want_t)


def test_split2h_scales_and_precision():
    """rows on scales from 1e-30 to 1e30, an all-zero row, elements far below their row's maximum: hi + lo times the
    inverse scale reproduces x to 2^-22 of the element (or 2^-38 of the row maximum), the padding is zero, the scales
    are powers of two that bring the row maximum into [2^14, 2^15)."""
    torch.manual_seed(0)
    R, Cn = 300, 100
    x = torch.randn(R, Cn, device="cuda") * torch.rand(R, 1, device="cuda").mul(138).sub(69).exp()
    x[7] = 0.0
    x[9, 1:] *= 1e-7            # a row dominated by one element
    x[11, 3] = 3.0e38
    r, ri, t, ti = _split2h(x, True, True)
    torch.cuda.synchronize()
    Cp, Rp = 128, 320
    assert r.shape == (R, 2 * Cp) and t.shape == (Cn, 2 * Rp) and ri.shape == (R,) and ti.shape == (Cn,)
    rv = r.float().view(R, 2, Cp)
    assert torch.equal(rv[:, :, Cn:], torch.zeros_like(rv[:, :, Cn:]))
    assert bool(torch.isfinite(rv).all())
    rec = (rv[:, 0, :Cn].double() + rv[:, 1, :Cn].double()) * ri.double()[:, None]
    rowmax = x.double().abs().amax(dim=1, keepdim=True)
    err = (rec - x.double()).abs()
    assert bool((err <= 2.0 ** -22 * x.double().abs() + 2.0 ** -38 * rowmax).all())
    m, e = torch.frexp(ri)
    assert torch.equal(m, torch.full_like(m, 0.5))      # powers of two
    scaled = rowmax[:, 0] / ri.double()
    nz = rowmax[:, 0] > 0
    assert bool(((scaled[nz] >= 2.0 ** 14) & (scaled[nz] < 2.0 ** 15)).all())
    assert float(ri[7]) == 1.0
    # transposed layout: the split of x^T scaled per column of x
    tv = t.float().view(Cn, 2, Rp)
    assert torch.equal(tv[:, :, R:], torch.zeros_like(tv[:, :, R:]))
    rect = (tv[:, 0, :R].double() + tv[:, 1, :R].double()) * ti.double()[:, None]
    colmax = x.double().abs().amax(dim=0)
    errt = (rect - x.double().t()).abs()
    assert bool((errt <= 2.0 ** -22 * x.double().t().abs() + 2.0 ** -38 * colmax[:, None]).all())
    # the rows-only entry point gives the same rows layout
    r2, ri2, _, _ = _split2h(x, True, False)
    assert torch.equal(r2, r) and torch.equal(ri2, ri)


# (M, N, K): trunk shapes of config 2 fwd / dgrad / wgrad (the last two run split-K), config 3's 20000-wide layers, ragged
# and tiny sizes, every tile width the planner can choose
X2_CASES = [(4096, 600, 784), (4096, 784, 600), (600, 784, 4096), (784, 600, 4096), (1024, 100, 20000), (1024, 20000, 100),
            (1000, 300, 100), (257, 601, 77), (130, 65, 1000), (128, 64, 64), (4096, 1000, 256), (300, 250, 513)]


@pytest.mark.parametrize("M,N,K", X2_CASES)
def test_gemm_x2s_matches_fp32_accuracy(M, N, K):
    from hvae import ops

    torch.manual_seed(M + N + K)
    torch.backends.cuda.matmul.allow_tf32 = False
    A = torch.randn(M, K, device="cuda") * torch.rand(M, 1, device="cuda").mul(8).sub(4).exp()   # rows on different scales
    B = torch.randn(N, K, device="cuda") * torch.rand(N, 1, device="cuda").mul(8).sub(4).exp()
    A[:, ::7] *= 1e-4                                                                              # and columns
    bias = torch.randn(N, device="cuda")
    ref = A.double() @ B.double().t() + bias.double()
    As, ai, _, _ = _split2h(A)
    Bs, bi, _, _ = _split2h(B)
    out = ops.gemm_x2s(As, ai, Bs, bi, bias, False, M, N, K)
    torch.cuda.synchronize()
    # element-wise: every output within fp32 accumulation error of its own scale sum_k |a_k b_k| (+ the bias)
    scale = A.double().abs() @ B.double().abs().t() + bias.double().abs()
    e_x2 = float(((out.double() - ref).abs() / scale).max())
    e_t = float((((A @ B.t() + bias).double() - ref).abs() / scale).max())
    assert e_x2 < 1e-6, e_x2
    assert e_x2 < 3.0 * e_t + 2e-7, (e_x2, e_t)
    outr = ops.gemm_x2s(As, ai, Bs, bi, bias, True, M, N, K)
    assert torch.equal(outr, out.clamp_min(0))
    # the contraction over the ROWS of two matrices (weight gradient): transposed splits with per-column scales
    if M * N <= 4096 * 1000:
        X = torch.randn(K, M, device="cuda") * torch.rand(1, M, device="cuda").mul(6).sub(3).exp()
        Y = torch.randn(K, N, device="cuda") * torch.rand(1, N, device="cuda").mul(6).sub(3).exp()
        _, _, Xt, xi = _split2h(X, False, True)
        _, _, Yt, yi = _split2h(Y, False, True)
        outw = ops.gemm_x2s(Xt, xi, Yt, yi, None, False, M, N, K)
        refw = X.double().t() @ Y.double()
        scalew = X.double().abs().t() @ Y.double().abs()
        assert float(((outw.double() - refw).abs() / scalew).max()) < 1e-6


def test_gemm_x2s_plan_fills_one_wave_at_config2():
    import ctypes

    from hvae import _cabi as C

    bn, sp, st = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    for (M, N, K) in ((4096, 600, 784), (4096, 784, 600), (600, 784, 4096)):
        assert C.lib().hvae_gemm_x2s_plan(M, N, K, ctypes.addressof(bn), ctypes.addressof(sp), ctypes.addressof(st)) == 0
        units = -(-M // 128) * -(-N // bn.value) * sp.value
        assert units <= 148 or sp.value > 1, (M, N, K, bn.value, sp.value)
        assert bn.value % 32 == 0 and 128 <= bn.value <= 256 and 2 <= st.value <= 4


@pytest.mark.parametrize("mode", ["x2", "x3"])
def test_linear_layer_forward_backward(mode):
    from hvae import layers, ops

    assert ops.get_trunk_mode() == "x2"   # the default
    ops.set_trunk_mode(mode)
    try:
        _linear_layer_forward_backward(mode)
    finally:
        ops.set_trunk_mode("x2")


def _linear_layer_forward_backward(mode):
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
        ops.set_trunk_mode(mode)
    assert torch.allclose(y_t, ref_lin(x))
    assert _rel(y, y_t.detach().double()) < 2e-6
    # small shapes defer to torch (too narrow, or too few flops to pay for the operand-split pipeline)
    small = layers.Linear(16, 8).cuda()
    xs = torch.randn(4, 16, device="cuda")
    assert torch.equal(small(xs), torch.nn.functional.linear(xs, small.weight, small.bias))
    tiny = layers.Linear(784, 64).cuda()
    xt = torch.randn(128, 784, device="cuda")
    assert not ops.trunk_x3_eligible(xt, tiny.weight)
    assert torch.equal(tiny(xt), torch.nn.functional.linear(xt, tiny.weight, tiny.bias))
    assert torch.equal(tiny(xt, relu=True), torch.relu(torch.nn.functional.linear(xt, tiny.weight, tiny.bias)))


@pytest.mark.parametrize("rows,n_in,n_out,need_gx", [(4096, 784, 600, False), (4096, 600, 784, True), (1000, 333, 601, True), (130, 257, 77, True)])
def test_linear_relu_fused_forward_backward(rows, n_in, n_out, need_gx):
    """Linear(relu=True): ReLU in the GEMM epilogue, its backward mask inside the gradient's operand split, the bias
    gradient out of the same maximum pass - against float64 relu(linear(x)) and its autograd."""
    from hvae import layers, ops

    torch.manual_seed(rows + n_in)
    lin = layers.Linear(n_in, n_out).cuda()
    x = torch.randn(rows, n_in, device="cuda", requires_grad=need_gx)
    gy = torch.randn(rows, n_out, device="cuda")
    old_min = ops.set_trunk_min_flops(0.0)   # the small shapes too: the tensor-core path, not the cuBLAS route for tiny layers
    try:
        assert ops.trunk_x3_eligible(x, lin.weight) and ops.get_trunk_mode() == "x2"
        y = lin(x, relu=True)
        y.backward(gy)
    finally:
        ops.set_trunk_min_flops(old_min)
    xd = x.detach().double().requires_grad_(True)
    Wd, bd = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    zd = torch.nn.functional.linear(xd, Wd, bd)
    # pre-activations within fp32 rounding of zero may land on either side of the kink: take the mask the kernel used
    mask = (y.detach() > 0).double()
    assert float(((zd.detach() > 0).double() - mask).abs().sum()) <= 1e-4 * mask.numel()
    yd = zd * mask
    yd.backward(gy.double())
    pairs = [(y, yd), (lin.weight.grad, Wd.grad), (lin.bias.grad, bd.grad)] + ([(x.grad, xd.grad)] if need_gx else [])
    for got, ref in pairs:
        assert _rel(got, ref.detach()) < 2e-6
    assert float((y < 0).sum()) == 0 and 0.3 < float((y == 0).float().mean()) < 0.7


@pytest.mark.parametrize("R,C", [(4096, 784), (1000, 77), (5, 3), (33, 600), (64, 130)])
@pytest.mark.parametrize("masked", [False, True])
def test_split2h_both_ex_mask_and_colsum(R, C, masked):
    """The general operand split: masked input == the plain split of the masked tensor (bit for bit), column sums == a
    float64 sum of the masked tensor."""
    from hvae import ops

    torch.manual_seed(R * 3 + C)
    x = torch.randn(R, C, device="cuda") * torch.rand(R, 1, device="cuda").mul(3).exp()
    m = torch.randn(R, C, device="cuda") if masked else None
    xm = x * (m > 0) if masked else x
    r, ri, t, ti, cs = ops.split2h_both_ex(x, m, True, True, True)
    r0, ri0, t0, ti0 = ops.split2h_both(xm.contiguous(), True, True)
    assert torch.equal(r, r0) and torch.equal(ri, ri0) and torch.equal(t, t0) and torch.equal(ti, ti0)
    ref = xm.double().sum(0)
    assert float((cs.double() - ref).abs().max()) < 1e-5 * float(xm.double().abs().sum(0).max() + 1e-30)
    # column sums alone / transposed layout alone
    _, _, _, _, cs2 = ops.split2h_both_ex(x, m, False, False, True)
    assert torch.equal(cs, cs2)
    _, _, t3, ti3, _ = ops.split2h_both_ex(x, m, False, True, False)
    assert torch.equal(t3, t0) and torch.equal(ti3, ti0)


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
        ops.set_trunk_mode("x2" if mode == "ours" else "torch")
        try:
            out = model.loss(x, alpha=alpha, r=r)
            out["loss_total"].backward()
        finally:
            ops.set_trunk_mode("x2")
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
