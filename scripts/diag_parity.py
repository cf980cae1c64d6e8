#!/usr/bin/env python
"""Diagnostic (GPU): the worst well-conditioned element of one gyroplane parity case against the fp32 / float64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import test_gpu_layers as T
from util_parity import pair_kappa

kind, D, P, B = "unsigned", 2, 512, 77
c = 1.0
torch.manual_seed(D * 1000 + P)
layer, make_o, names = T._gyro_layers(kind, D, P, c)
ob = T._oball(c)
x = ob.expmap0(torch.randn(B, D) * 0.8 / D ** 0.5).detach()
params = {k: getattr(layer, k).detach().clone() for k in names}
x[2] = params["points"][1] * (1 + 1e-4)
x[3] = ob.expmap0(torch.randn(D) * 50.0)
params["points"][min(3, P - 1)] *= 1e-9
gout = torch.randn(B, P)
cu = T._cuda_layer_run(layer, params, x, gout)
o32 = T._oracle_layer_run(make_o, params, x, gout, torch.float32)
o64 = T._oracle_layer_run(make_o, params, x, gout, torch.float64)
pk = pair_kappa(float(ob.c), x, params["points"])
out, r32, r64 = cu[0].double().cpu(), o32[0].double(), o64[0]
g = r32.abs().max()
b = 1e-5 * r64.abs() + 1e-6 * g
e64, e32 = (out - r64).abs(), (out - r32).abs()
ratio = torch.minimum(e64, e32) / b
ratio[pk >= 2] = 0
for idx in ratio.flatten().argsort(descending=True)[:5].tolist():
    i, j = divmod(idx, P)
    print("(%d,%d) ratio %.2f pk %.3f cuda %.9g o32 %.9g o64 %.9g | x %s p %s" % (i, j, float(ratio[i, j]), float(pk[i, j]), float(out[i, j]), float(r32[i, j]), float(r64[i, j]), x[i].tolist(), params["points"][j].tolist()))
