#!/usr/bin/env python
"""Per-kernel microbenchmarks at roofline-relevant sizes (NOT the contract bench; see bench.py).
  - config-5 shapes (B x 512 -> 4096): tcgen05 Mobius / gyroplane forward vs the bf16 tensor peak and HBM write bound
  - large-row HBM kernels: expmap0, wrapped sample, KL, fused latent head (fwd and bwd)
Prints one JSON object; CUDA events around several launches after warm-up; inputs >> L2 (126 MB)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) * 1e-3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logB", type=int, default=18)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from hvae import ops
    import hvae

    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    dev = torch.device("cuda")
    c = hvae.PoincareBall(1.0).c_value
    out = {"peaks": {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops": pk["bf16_tflops"]}}
    g = torch.Generator(device=dev).manual_seed(0)
    if args.only in ("", "tc"):
        B, F, P = 1 << args.logB, 512, 4096
        x = ops.expmap0(torch.randn(B, F, device=dev, generator=g) * 0.1, c)
        M = torch.randn(P, F, device=dev, generator=g) / F ** 0.5
        pts = ops.expmap0(torch.randn(P, F, device=dev, generator=g) * 0.03, c)
        ops.set_gemm_mode("bf16")
        t = timeit(lambda: ops.mobius_matvec_tc_fwd(x, M, c))
        fl = 2.0 * B * F * P
        out["mobius_tc_fwd"] = {"B": B, "F": F, "P": P, "ms": t * 1e3, "tflops": fl / t / 1e12, "frac_tensor": fl / t / 1e12 / pk["bf16_tflops"],
                                "alg_bytes": 4 * (B * F + P * F + B * P), "gbs_alg": 4 * (B * F + P * F + B * P) / t / 1e9,
                                "note": "v1 = bf16 convert + GEMM(+|mx|^2 partials, fp32 mx out) + rescale pass: 3 passes over B*P fp32"}
        t = timeit(lambda: ops.mobius_matvec_tc(x, M, c))
        out["mobius_tc_fwd_fused"] = {"B": B, "F": F, "P": P, "ms": t * 1e3, "tflops": fl / t / 1e12, "frac_tensor": fl / t / 1e12 / pk["bf16_tflops"],
                                      "gbs_alg": 4 * (B * F + P * F + B * P) / t / 1e9, "frac_hbm": 4 * (B * F + P * F + B * P) / t / 1e9 / pk["hbm_gbs"],
                                      "note": "single pass over the output: Gram row scale + fused rescale epilogue (forward-only)"}
        y, mxsq = ops.mobius_matvec_tc(x, M, c)
        gy = torch.randn_like(y)
        t = timeit(lambda: ops.mobius_matvec_tc_bwd(x, M, y, mxsq, gy, c))
        out["mobius_tc_bwd"] = {"B": B, "F": F, "P": P, "ms": t * 1e3, "tflops": 2 * fl / t / 1e12, "frac_tensor": 2 * fl / t / 1e12 / pk["bf16_tflops"],
                                "note": "row pass (bf16 gmx) + bf16 transpose + 2 GEMMs (gx over P with fused axpy, gM over B)"}
        del y, gy
        t = timeit(lambda: ops.gyroplane_tc_fwd(x, pts, None, c, ops.GYRO_SIGNED))
        og = torch.randn(B, P, device=dev, generator=g)
        tb = timeit(lambda: ops.gyroplane_tc_bwd(x, pts, og, c, ops.GYRO_SIGNED))
        out["gyroplane_tc_bwd"] = {"B": B, "D": F, "P": P, "ms": tb * 1e3, "tflops": 3 * fl / tb / 1e12, "frac_tensor": 3 * fl / tb / 1e12 / pk["bf16_tflops"],
                                   "note": "recompute GEMM + pair-gradient tile kernel + bf16 transpose + 2 GEMMs (3 x 2BDP flop)"}
        del og
        out["gyroplane_tc_fwd"] = {"B": B, "D": F, "P": P, "ms": t * 1e3, "tflops": fl / t / 1e12, "frac_tensor": fl / t / 1e12 / pk["bf16_tflops"],
                                   "alg_bytes": 4 * (B * F + 2 * P * F + B * P), "gbs_alg": 4 * (B * F + 2 * P * F + B * P) / t / 1e9,
                                   "frac_hbm": 4 * (B * F + 2 * P * F + B * P) / t / 1e9 / pk["hbm_gbs"]}
        # cuBLAS bf16 GEMM of the same shape for scale (library, not ours)
        xb, Mb = x.bfloat16(), M.bfloat16()
        t = timeit(lambda: torch.matmul(xb, Mb.t()))
        out["cublas_bf16_same_shape"] = {"ms": t * 1e3, "tflops": fl / t / 1e12, "note": "bf16 output (half the write bytes of ours)"}
        ops.set_gemm_mode("fp32")
        del x, M, pts, xb, Mb
        torch.cuda.empty_cache()
    if args.only in ("", "rows"):
        for D in (2, 10, 64):
            Bb = (1 << 28) // (4 * D)
            u = torch.randn(Bb, D, device=dev, generator=g) * 0.3
            mu = ops.expmap0(u, c)
            sg = torch.rand(Bb, D, device=dev, generator=g) + 0.3
            eps = torch.randn(Bb, D, device=dev, generator=g)
            z, kl = ops.latent_head_fwd(mu, sg, eps, 1.0, c)
            gz, gkl = torch.randn_like(z), torch.randn_like(kl)
            cases = [
                ("expmap0_fwd", lambda: ops.expmap0_fwd(u, c), 8 * Bb * D),
                ("expmap0_bwd", lambda: ops.expmap0_bwd(u, gz, c), 12 * Bb * D),
                ("wrapped_sample_fwd", lambda: ops.wrapped_sample_fwd(mu, sg, eps.view(1, Bb, D), c), 16 * Bb * D),
                ("wrapped_logprob_fwd", lambda: ops.wrapped_logprob_fwd(mu, sg, z.view(1, Bb, D), c), 12 * Bb * D + 4 * Bb),
                ("latent_head_fwd", lambda: ops.latent_head_fwd(mu, sg, eps, 1.0, c), 16 * Bb * D + 4 * Bb),
                ("latent_head_bwd", lambda: ops.latent_head_bwd(mu, sg, eps, gz, gkl, 1.0, c), 20 * Bb * D + 4 * Bb),
            ]
            for name, fn, nbytes in cases:
                t = timeit(fn, iters=4, warm=1)
                out["%s_D%d" % (name, D)] = {"rows": Bb, "ms": t * 1e3, "gbs": nbytes / t / 1e9, "frac_hbm": nbytes / t / 1e9 / pk["hbm_gbs"]}
            del u, mu, sg, eps, z, kl, gz, gkl
            torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
