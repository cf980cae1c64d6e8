"""Tiny driver for ncu captures of the SIMT gyroplane kernels at the config-2 decoder shape (B=4096, D=10, P=600, a != p)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import hvae
from hvae import ops

dev = torch.device("cuda")
c = hvae.PoincareBall(1.0).c_value
B, D, H = 4096, 10, 600
g = torch.Generator(device=dev).manual_seed(0)
z = ops.expmap0(torch.randn(B, D, device=dev, generator=g) * 0.3, c)
Wg = torch.randn(H, D, device=dev, generator=g) * 0.3
bg = torch.randn(H, device=dev, generator=g) * 0.1
bpt, Mg = ops.weight_prep_fwd(Wg, bg, c)
FL = ops.GYRO_PVAE | ops.GYRO_SIGNED
for _ in range(3):
    out = ops.gyroplane_fwd(z, Mg, bpt, None, c, FL)
    gout = torch.randn_like(out)
    r = ops.gyroplane_bwd(z, Mg, bpt, None, gout, c, FL, False)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
