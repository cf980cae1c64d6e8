"""Tiny driver for ncu captures of the tcgen05 kernels (one forward of each at B=2^16, 512 -> 4096)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import hvae
from hvae import ops

dev = torch.device("cuda")
c = hvae.PoincareBall(1.0).c_value
B, F, P = 1 << 16, 512, 4096
g = torch.Generator(device=dev).manual_seed(0)
x = ops.expmap0(torch.randn(B, F, device=dev, generator=g) * 0.1, c)
M = torch.randn(P, F, device=dev, generator=g) / F ** 0.5
pts = ops.expmap0(torch.randn(P, F, device=dev, generator=g) * 0.03, c)
for _ in range(2):
    y, _ = ops.mobius_matvec_tc(x, M, c)
    o = ops.gyroplane_tc_fwd(x, pts, None, c, ops.GYRO_SIGNED)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()), float(o.abs().mean()))
