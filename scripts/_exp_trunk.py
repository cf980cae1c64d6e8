import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import hvae
from hvae import ops
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / it * 1e3
M, I, O = 4096, 784, 600
x = torch.randn(M, I, device=dev); W = torch.randn(O, I, device=dev); gy = torch.randn(M, O, device=dev); b = torch.randn(O, device=dev)
xs, ws, gs = ops.split3(x), ops.split3(W), ops.split3(gy)
print("split3 x   %.1f us" % timeit(lambda: ops.split3(x)))
print("split3 W   %.1f us" % timeit(lambda: ops.split3(W)))
print("split3 gy  %.1f us" % timeit(lambda: ops.split3(gy)))
print("fwd  KK (4096,600,784)  %.1f us" % timeit(lambda: ops.gemm_x3s(xs, False, ws, False, b, False, M, O, I)))
print("dgrad K,MN (4096,784,600) %.1f us" % timeit(lambda: ops.gemm_x3s(gs, False, ws, True, None, False, M, I, O)))
print("wgrad MN,MN (600,784,4096) %.1f us" % timeit(lambda: ops.gemm_x3s(gs, True, xs, True, None, False, O, I, M)))
print("old gemm_x3 fwd  %.1f us" % timeit(lambda: ops.gemm_x3(x, False, W, False, b, False)))
print("old gemm_x3 dgrad %.1f us" % timeit(lambda: ops.gemm_x3(gy, False, W, True, None, False)))
print("old gemm_x3 wgrad %.1f us" % timeit(lambda: ops.gemm_x3(gy, True, x, True, None, False)))
Wt = W.t().contiguous(); wts = ops.split3(Wt)
print("dgrad K,K presplit %.1f us" % timeit(lambda: ops.gemm_x3s(gs, False, wts, False, None, False, M, I, O)))
torch.backends.cuda.matmul.allow_tf32 = False
print("torch fwd  %.1f us" % timeit(lambda: torch.nn.functional.linear(x, W, b)))
print("torch dgrad %.1f us" % timeit(lambda: gy @ W))
print("torch wgrad %.1f us" % timeit(lambda: gy.t() @ x))
lg = torch.randn(1, M, I, device=dev); t = torch.rand(M, I, device=dev); g = torch.randn(1, M, device=dev)
print("bce fwd %.1f us" % timeit(lambda: ops.bce_logits_rows_fwd(lg, t)))
print("bce bwd %.1f us" % timeit(lambda: ops.bce_logits_rows_bwd(lg, t, g)))
print("torch bce fwd %.1f us" % timeit(lambda: torch.nn.functional.binary_cross_entropy_with_logits(lg, t.view(1, M, I), reduction="none").sum(-1)))
