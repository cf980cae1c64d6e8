#!/usr/bin/env python
"""Diagnostic (GPU): where does logmap0's gradient leave the plain 1e-5 band on well-conditioned rows?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import hvae
from test_gpu_row_ops import _oracle_ball, _run_oracle

for D, c, s in ((512, 1.0, 1e-3), (512, 1.0, 0.3), (64, 0.5, 1e-3), (10, 1.0, 0.3)):
    torch.manual_seed(D * 7 + 1)
    B = 257
    u = torch.randn(B, D) * s / (D ** 0.5)
    u[0].zero_()
    y = _oracle_ball(c, torch.float32).expmap0(u).detach()
    g = torch.Generator().manual_seed(0)
    gout = torch.randn(B, D, generator=g)
    ball = hvae.PoincareBall(c)
    yc = y.cuda().requires_grad_(True)
    out = ball.logmap0(yc)
    out.backward(gout.cuda())
    o32, (g32,) = _run_oracle(lambda a: _oracle_ball(c, torch.float32).logmap0(a), [y], gout, torch.float32)
    o64, (g64,) = _run_oracle(lambda a: _oracle_ball(c, torch.float64).logmap0(a), [y], gout, torch.float64)
    gc = yc.grad.double().cpu()
    sc = g64.abs().amax(-1, keepdim=True)
    e64 = ((gc - g64).abs() / sc)
    e32 = ((gc - g32.double()).abs() / sc)
    r32 = ((g32.double() - g64).abs() / sc)
    worst = e64.amax(-1)
    idx = worst.argsort(descending=True)[:4]
    print("D=%d c=%g s=%g: rows failing both: %d; kernel-vs-64 max %.3g, ref32-vs-64 max %.3g" % (D, c, s, int(((e64 > 1e-5) & (e32 > 1e-5)).any(-1).sum()), float(e64.max()), float(r32.max())))
    for i in idx.tolist():
        j = int(e64[i].argmax())
        print("   row %d |y|=%.4g  worst elem %d: cuda=%.8g o32=%.8g o64=%.8g  rowscale=%.4g  gdot=%.4g" % (i, float(y[i].norm()), j, float(gc[i, j]), float(g32[i, j]), float(g64[i, j]), float(sc[i]), float((gout[i].double() * y[i].double()).sum())))
    fo = ((out.double().cpu() - o64).abs() / o64.abs().amax(-1, keepdim=True).clamp_min(1e-30))
    print("   fwd kernel-vs-64 max %.3g (row %d)" % (float(fo.max()), int(fo.amax(-1).argmax())))
