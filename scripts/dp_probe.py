"""torchrun probe (N ranks): all-reduce alone vs. the config-2 step with / without the collective and with the
early-segment overlap.  CUDA events, graph replay, no L2 flush (relative numbers only)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import hvae
from hvae import models, train

dev = torch.device("cuda", local)


def timed(fn, it=200):
    for _ in range(10):
        fn()
    dist.barrier(device_ids=[local]); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / it], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3  # us


out = {"world": world}
buf = torch.zeros(954_000, device=dev)
out["allreduce_3.8MB_us"] = timed(lambda: dist.all_reduce(buf))
half = buf[:477_000]
out["allreduce_1.9MB_us"] = timed(lambda: dist.all_reduce(half))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    dist.all_reduce(buf)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    dist.all_reduce(buf)
out["allreduce_3.8MB_graph_us"] = timed(g.replay)

from hvae.parallel import FlatGradBucket
pp = [torch.nn.Parameter(torch.zeros(954_000, device=dev))]
bk = FlatGradBucket(pp, symmetric=True)
if bk._symm is not None:
    out["p2p_allreduce_3.8MB_us"] = timed(lambda: bk.all_reduce(average=False))
    bk.buffer.fill_(float(rank + 1)); bk.all_reduce(average=False); torch.cuda.synchronize()
    out["p2p_sum_check"] = float(bk.buffer.min()), float(bk.buffer.max()), world * (world + 1) / 2
torch.manual_seed(0)
x = torch.rand(4096, 1, 28, 28, device=dev)
ref_grads = {}
for mode in ("p2p_overlap", "p2p_single", "overlap", "single", "none"):
    os.environ["HVAE_DP_OVERLAP"] = "1" if mode.endswith("overlap") else "0"
    os.environ["HVAE_DP_P2P"] = "1" if mode.startswith("p2p") else "0"
    torch.manual_seed(0)
    model = models.PvaeMnist().to(dev)
    ts = train.TrainStep(model, x)
    if mode == "none":
        ts.graph = None
        ts.bucket.all_reduce = lambda **k: None
        ts.overlap = False
        ts._capture()
    out["step_%s_us" % mode] = timed(ts.run, it=100)
    out["step_%s_graph" % mode] = ts.graph is not None
    out["step_%s_symm" % mode] = ts.bucket._symm is not None
    if mode != "none":
        # same weights, same batch, noise from a reset counter: the reduced gradients of every mode must agree
        gen = torch.Generator(device=dev).manual_seed(77 + rank)
        ts.loss_kwargs = dict(alpha=torch.randn(1, 4096, 10, device=dev, generator=gen),
                              r=torch.rand(1, 4096, 1, device=dev, generator=gen) * 2 + 0.1)
        ts.graph = None
        ts.run(); torch.cuda.synchronize()
        ref_grads[mode] = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    if mode == "overlap":
        out["early_bytes"] = ts.bucket.split * 4 if ts.overlap else 0
        out["total_bytes"] = ts.bucket.nbytes
base = ref_grads.get("single")
for mode, g in ref_grads.items():
    if mode == "single" or base is None:
        continue
    out["maxrel_%s_vs_nccl" % mode] = max(float((g[n] - base[n]).abs().max() / base[n].abs().max().clamp_min(1e-30)) for n in base)
if rank == 0:
    print(json.dumps(out))
dist.barrier(device_ids=[local]); torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
