"""CPU check of the arithmetic behind the fp16 two-piece tensor-core GEMM (csrc/tc_x2.cu): a float32 row scaled by a power
of two into [2^14, 2^15) is the sum of two float16 pieces to 2^-22, the three piece products hi*hi + hi*lo + lo*hi
reproduce an fp32 GEMM to fp32 accuracy whatever the rows' magnitudes - and without the scale they do not (fp16's range).
(The kernels are tested on the GPU in tests/test_gpu_trunk.py; this pins the arithmetic without a device.)"""
import torch


def _scale_rows(v):
    m = v.abs().amax(dim=1, keepdim=True)
    e = torch.floor(torch.log2(m.double().clamp_min(1e-300))).clamp_min(-100.0)
    s = torch.where(m > 0, torch.pow(2.0, 14.0 - e), torch.ones_like(e)).float()
    return s


def _split2(v):
    h = v.to(torch.float16)
    l = (v - h.float()).to(torch.float16)
    return h, l


def test_two_fp16_pieces_carry_22_bits_after_row_scaling():
    torch.manual_seed(0)
    v = torch.randn(512, 100) * torch.rand(512, 1).mul(120).sub(60).exp()
    s = _scale_rows(v)
    vs = v * s                                   # exact: power of two
    assert bool(((vs.abs().amax(dim=1) >= 2.0 ** 14) & (vs.abs().amax(dim=1) < 2.0 ** 15)).all())
    h, l = _split2(vs)
    assert bool(torch.isfinite(h.float()).all())
    rec = (h.double() + l.double()) / s.double()
    rowmax = v.double().abs().amax(dim=1, keepdim=True)
    assert bool(((rec - v.double()).abs() <= 2.0 ** -22 * v.double().abs() + 2.0 ** -38 * rowmax).all())


def _emulated(A, B, scaled):
    sa = _scale_rows(A) if scaled else torch.ones(A.shape[0], 1)
    sb = _scale_rows(B) if scaled else torch.ones(B.shape[0], 1)
    ah, al = _split2(A * sa)
    bh, bl = _split2(B * sb)
    acc = al.double() @ bh.double().t() + ah.double() @ bl.double().t() + ah.double() @ bh.double().t()   # smallest first
    return acc / (sa.double() * sb.double().t())


def test_three_products_reach_fp32_accuracy_with_row_scales():
    torch.manual_seed(1)
    A = torch.randn(96, 784) * torch.rand(96, 1).mul(30).sub(15).exp()
    B = torch.randn(64, 784) * torch.rand(64, 1).mul(30).sub(15).exp()
    ref = A.double() @ B.double().t()
    scale = A.double().abs() @ B.double().abs().t()
    e3 = float(((_emulated(A, B, True) - ref).abs() / scale).max())
    e_fp32 = float((((A @ B.t()).double() - ref).abs() / scale).max())
    assert e3 < 1e-7            # per-term 3 * 2^-22, random in sign over 784 terms
    assert e3 < 4 * e_fp32 + 1e-8
    # without the scales fp16 overflows / flushes: rows at e^15 ~ 3e6 and e^-15 ~ 3e-7 are outside its range
    bad = _emulated(A, B, False)
    assert not bool(torch.isfinite(bad).all()) or float(((bad - ref).abs() / scale).max()) > 1e-3
