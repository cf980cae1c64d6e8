#!/usr/bin/env python
"""GPU probe of the CTA-pair tcgen05 kernel (tc_gemm2.cu): plain-GEMM correctness against torch on ragged shapes,
then the config-5 timings (scripts/microbench_kernels.py --only tc).  Prints JSON lines."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    import hvae
    from hvae import ops

    c = hvae.PoincareBall(1.0).c_value
    dev = torch.device("cuda")
    ops.set_gemm_mode("bf16")
    res = {}
    for (B, F, P) in [(1024, 64, 128), (4096, 512, 1024), (19000, 512, 600), (2048, 128, 4096), (3000, 72, 136)]:
        torch.manual_seed(B)
        x = ops.expmap0(torch.randn(B, F, device=dev) * 0.5 / F ** 0.5, c)
        M = torch.randn(P, F, device=dev) / F ** 0.5 * 0.7
        y, mx = ops.mobius_matvec_tc_fwd(x, M, c)          # PLAIN epilogue + rowsq partials
        torch.cuda.synchronize()
        ref = x.double() @ M.double().t()
        res["plain_%d_%d_%d" % (B, F, P)] = float((mx.double() - ref).abs().max() / ref.abs().max())
        if P % 8 == 0:
            y2, mxsq = ops.mobius_matvec_tc(x, M, c)       # ROWDOT + MOBIUS epilogues
            torch.cuda.synchronize()
            res["fused_vs_2pass_%d_%d_%d" % (B, F, P)] = float((y2 - y).abs().max() / y.abs().max())
            res["mxsq_%d_%d_%d" % (B, F, P)] = float(((mxsq.double() - ref.pow(2).sum(-1)).abs() / ref.pow(2).sum(-1)).max())
    print(json.dumps(res))
    sys.stdout.flush()


if __name__ == "__main__":
    main()
