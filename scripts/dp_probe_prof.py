#!/usr/bin/env python
"""2-rank probe: per-kernel device time of the graph-replayed config-2 step with an ordinary bucket vs a symmetric-memory
bucket (no exchange in either), torch.profiler (CUPTI), rank 0 prints the kernels whose time differs."""
import collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import hvae.parallel as HP
from hvae import models as HM
from hvae.train import TrainStep

x = torch.rand(4096, 1, 28, 28, generator=torch.Generator().manual_seed(1000 + rank)).clamp(1e-5, 1 - 1e-5).to(dev)
res = {}
EXCH = os.environ.get("PROBE_EXCHANGE", "0") == "1"
if not EXCH:
    HP.FlatGradBucket.all_reduce = lambda self, average, group=None, async_op=False: None
    HP.FlatGradBucket.all_reduce_segment = lambda self, which, average, group=None: None
for mode in (("symm",) if EXCH else ("plain", "symm")):
    os.environ["HVAE_DP_P2P"] = "0" if mode == "plain" else "1"
    os.environ["HVAE_DP_OVERLAP"] = "0"
    torch.manual_seed(42)
    m = HM.PvaeMnist(latent_dim=10, hidden_dim=600).to(dev)
    ts = TrainStep(m, x, use_graph=True, average_grads=False)
    for _ in range(10):
        ts.run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            ts.run()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name[:60]][0] += ev.device_time / 20
            agg[ev.name[:60]][1] += 1
    res[mode] = {k: (round(v[0], 2), v[1] // 20) for k, v in agg.items()}
    res[mode + "_total"] = round(sum(v[0] for v in agg.values()), 1)
    ts.graph = None
    del ts, m
if EXCH:
    ar = {k: v for k, v in res["symm"].items() if "allreduce" in k}
    print("rank", rank, "kernel-time total", res["symm_total"], "all-reduce kernel:", ar, flush=True)
    dist.barrier(device_ids=[local]); torch.cuda.synchronize(); os._exit(0)
if rank == 0:
    print("totals", res["plain_total"], res["symm_total"])
    keys = set(res["plain"]) | set(res["symm"])
    rows = sorted(((res["symm"].get(k, (0, 0))[0] - res["plain"].get(k, (0, 0))[0], k) for k in keys), reverse=True)
    for d, k in rows[:12]:
        print("%+8.2f us  %-60s plain %s symm %s" % (d, k, res["plain"].get(k), res["symm"].get(k)))
dist.barrier(device_ids=[local]); torch.cuda.synchronize(); os._exit(0)
