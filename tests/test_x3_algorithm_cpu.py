"""CPU check of the algorithm behind the fp32-accurate tensor-core GEMM (csrc/tc_gemm.cu, EPI_X3): a float32 value is
the exact sum of three bfloat16 pieces, and the six piece products down to 2^-16 reproduce an fp32 GEMM to fp32
accuracy — while the three largest products alone do not.  (The kernel itself is tested on the GPU in
tests/test_gpu_trunk.py; this pins the arithmetic it relies on without a device.)"""
import torch


def _split3(v):
    h = v.to(torch.bfloat16)
    r1 = v - h.float()
    m = r1.to(torch.bfloat16)
    r2 = r1 - m.float()
    l = r2.to(torch.bfloat16)
    return h, m, l


def test_three_bf16_pieces_carry_24_bits():
    torch.manual_seed(0)
    v = torch.randn(4096) * torch.rand(4096).mul(40).sub(20).exp()
    h, m, l = _split3(v)
    rec = h.double() + m.double() + l.double()
    assert float(((rec - v.double()).abs() / v.double().abs()).max()) < 2.0 ** -23
    # the residual subtractions are exact in fp32 (Sterbenz): re-adding in fp32 in any order gives the value back to 1 ulp
    assert float(((h.float() + (m.float() + l.float())) - v).abs().div(v.abs()).max()) < 2.0 ** -22


def _emulated_gemm(A, B, pairs):
    pa, pb = _split3(A), _split3(B)
    acc = torch.zeros(A.shape[0], B.shape[0], dtype=torch.float64)
    for i, j in pairs:  # the tensor core multiplies bf16 pieces exactly and accumulates in (at least) fp32
        acc += pa[i].double() @ pb[j].double().t()
    return acc


def test_six_products_reach_fp32_accuracy_three_do_not():
    torch.manual_seed(1)
    A, B = torch.randn(96, 784), torch.randn(64, 784)
    ref = A.double() @ B.double().t()
    scale = ref.abs().max()
    six = [(1, 1), (0, 2), (2, 0), (0, 1), (1, 0), (0, 0)]  # the kernel's order: smallest products first
    three = [(0, 1), (1, 0), (0, 0)]
    e6 = float((_emulated_gemm(A, B, six) - ref).abs().max() / scale)
    e3 = float((_emulated_gemm(A, B, three) - ref).abs().max() / scale)
    e_fp32 = float(((A @ B.t()).double() - ref).abs().max() / scale)
    assert e6 < 2e-7            # dropped terms (mid*lo, lo*mid, lo*lo) are ~2^-24 of a product, random in sign
    assert e6 < 4 * e_fp32 + 1e-8
    assert e3 > 20 * e6         # without the 2^-16 terms the product is only ~16-bit accurate
