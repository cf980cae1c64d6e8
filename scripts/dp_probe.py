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

torch.manual_seed(0)
x = torch.rand(4096, 1, 28, 28, device=dev)
for mode in ("overlap", "single", "none"):
    os.environ["HVAE_DP_OVERLAP"] = "1" if mode == "overlap" else "0"
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
    if mode == "overlap":
        out["early_bytes"] = ts.bucket.split * 4 if ts.overlap else 0
        out["total_bytes"] = ts.bucket.nbytes
if rank == 0:
    print(json.dumps(out))
dist.barrier(device_ids=[local]); torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
