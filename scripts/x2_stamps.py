#!/usr/bin/env python
"""Experiment build only: clock64 stamps of CTA 0 of the trunk GEMM (producer after each empty-wait, MMA issuer after each
full-wait, one epilogue warp after each hand-over and after its stores).  HVAE_LIB_PATH=..._exp.so python scripts/x2_stamps.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    from hvae import _cabi as C
    from hvae import ops

    M, N, K = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096x600x784").split("x"))
    dev = torch.device("cuda")
    A, B = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    a2, ai, _, _ = ops.split2h_both(A, True, False)
    b2, bi, _, _ = ops.split2h_both(B, True, False)
    ts = torch.zeros(192 + 3 * 148, dtype=torch.int64, device=dev)
    f = C.lib().hvae_exp_x2_timestamps
    f.argtypes = [ctypes.c_void_p]
    f.restype = None
    for _ in range(3):
        ops.gemm_x2s(a2, ai, b2, bi, None, False, M, N, K)
    torch.cuda.synchronize()
    f(ts.data_ptr())
    ops.gemm_x2s(a2, ai, b2, bi, None, False, M, N, K)
    torch.cuda.synchronize()
    f(None)
    t = ts.cpu().tolist()
    t0 = min(v for v in t if v)
    rel = lambda xs: [v - t0 for v in xs if v]  # noqa: E731
    cta = [(t[192 + 3 * i], t[192 + 3 * i + 1], t[192 + 3 * i + 2]) for i in range(148) if t[192 + 3 * i]]
    g0 = min(c[0] for c in cta)
    per_cta = sorted((c[1] - g0, c[0] - g0, c[2]) for c in cta)   # (end ns, start ns, smid), relative to the first CTA start
    c0 = [c for c in ((t[192], t[193]),)][0]
    extra = {"cta0_first_instr_globaltimer_to_setup_done_ns": c0[0] - t[190], "cta0_setup_done_to_end_ns": c0[1] - c0[0],
             "cta0_first_instr_clock": t[191], "cta0_end_clock": t[189], "cta0_clock_span": t[189] - t[191]}
    t = t[:189]
    t0 = min(v for v in t if v)
    print(json.dumps({"shape": [M, N, K], "extra": extra, "cta_end_start_smid_ns_sorted_by_end": per_cta[:4] + per_cta[-12:], "producer_after_empty_wait": rel(t[:64]), "mma_after_full_wait": rel(t[64:128]),
                      "epilogue_after_tfull_wait_then_end": rel(t[128:192])}))


if __name__ == "__main__":
    main()
