#!/usr/bin/env python
"""One call of each tensor-core op at config-5 shape (B = 2^logB, 512 -> 4096), for ncu launch lists / captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import hvae
from hvae import ops

logB = int(sys.argv[1]) if len(sys.argv) > 1 else 18
which = sys.argv[2] if len(sys.argv) > 2 else "all"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda")
c = hvae.PoincareBall(1.0).c_value
B, F, P = 1 << logB, 512, 4096
g = torch.Generator(device=dev).manual_seed(0)
x = ops.expmap0(torch.randn(B, F, device=dev, generator=g) * 0.1, c)
M = torch.randn(P, F, device=dev, generator=g) / F ** 0.5
pts = ops.expmap0(torch.randn(P, F, device=dev, generator=g) * 0.03, c)
ops.set_gemm_mode("bf16")
for _ in range(reps):
    if which in ("all", "mobius"):
        y, mxsq = ops.mobius_matvec_tc(x, M, c)
        gy = torch.randn_like(y)
        torch.cuda.synchronize()
        ops.mobius_matvec_tc_bwd(x, M, y, mxsq, gy, c)
        del y, gy
    if which in ("all", "gyro"):
        out = ops.gyroplane_tc_fwd(x, pts, None, c, ops.GYRO_SIGNED)
        og = torch.randn_like(out)
        torch.cuda.synchronize()
        ops.gyroplane_tc_bwd(x, pts, og, c, ops.GYRO_SIGNED)
torch.cuda.synchronize()
print("done")
