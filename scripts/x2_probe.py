#!/usr/bin/env python
"""Per-shape timing of the trunk GEMM paths (two-piece fp16 `x2`, three-piece bf16 `x3`, cuBLAS fp32) and of the operand
splits, with and without an L2 flush between launches.  Knobs of the experiment build (HVAE_LIB_PATH=..._exp.so):
HVAE_X2_BN / HVAE_X2_SPLITS / HVAE_X2_NST / HVAE_X2_HAND.  Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def _graph_us(fns, reps=10):
    """Median time of one replay of a CUDA graph holding `reps` repetitions of the calls in fns, per repetition."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns:
                f()
    ts = []
    for _ in range(7):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3 / reps)
    ts.sort()
    return ts[len(ts) // 2]


def timeit(fn, flush=None):
    """Device time of fn per call (CUDA-graph replay: no launch overhead); with `flush`, an L2-sized write precedes every
    call and its own time is subtracted."""
    if flush is None:
        return _graph_us([fn])
    fl = lambda: flush.add_(1.0)  # noqa: E731
    return _graph_us([fl, fn]) - _graph_us([fl])


def main():
    from hvae import ops

    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    shapes = [(4096, 600, 784), (4096, 784, 600), (784, 600, 4096), (600, 784, 4096)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    flush = torch.empty(64 << 20, device=dev)
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("HVAE_X2")}}
    for (M, N, K) in shapes:
        A, B = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
        a2, ai, _, _ = ops.split2h_both(A, True, False)
        b2, bi, _, _ = ops.split2h_both(B, True, False)
        a3, b3 = ops.split3(A), ops.split3(B)
        r = {}
        for tag, fl in (("warm", None), ("flushed", flush)):
            r["x2_" + tag] = timeit(lambda: ops.gemm_x2s(a2, ai, b2, bi, None, False, M, N, K), fl)
            r["x3_" + tag] = timeit(lambda: ops.gemm_x3s(a3, False, b3, False, None, False, M, N, K), fl)
            r["cublas_fp32_" + tag] = timeit(lambda: torch.mm(A, B.t()), fl)
        r["split2h_rows"] = timeit(lambda: ops.split2h_both(A, True, False))
        r["split2h_both"] = timeit(lambda: ops.split2h_both(A, True, True))
        r["split2h_t"] = timeit(lambda: ops.split2h_both(A, False, True))
        r["split3_both"] = timeit(lambda: ops.split3_both(A, True, True))
        out["%dx%dx%d" % (M, N, K)] = {k: round(v, 2) for k, v in r.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
