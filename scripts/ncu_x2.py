"""Tiny driver for the ncu capture of the trunk GEMM (x2::k_x2_gemm): the five GEMMs of one config-2 step on pre-split
operands, twice (the second round is the one to capture: --launch-skip = k_x2_gemm launches of round one)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
from hvae import ops

dev = torch.device("cuda")
B, H, n_in = 4096, 600, 784
g = torch.Generator(device=dev).manual_seed(0)
gemms = [(B, H, n_in), (B, n_in, H), (B, H, n_in), (n_in, H, B), (H, n_in, B)]
tc_ops = []
for m, n, k in gemms:
    a_s, a_i, _, _ = ops.split2h_both(torch.randn(m, k, device=dev, generator=g), True, False)
    b_s, b_i, _, _ = ops.split2h_both(torch.randn(n, k, device=dev, generator=g), True, False)
    tc_ops.append((a_s, a_i, b_s, b_i, m, n, k))
for _ in range(2):
    for a_s, a_i, b_s, b_i, m, n, k in tc_ops:
        out = ops.gemm_x2s(a_s, a_i, b_s, b_i, None, False, m, n, k)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
