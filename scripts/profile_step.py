"""Kernel-time breakdown of the config-2 step as it runs in the benchmark (CUDA-graph replay, warm L2), from CUPTI
via torch.profiler — ncu's launch list is cold-cache and serialised, this one is the in-situ share."""
import collections, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
import hvae
from hvae import models, train

dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.rand(4096, 1, 28, 28, device=dev)
model = models.PvaeMnist().to(dev)
ts = train.TrainStep(model, x)
for _ in range(10):
    ts.run()
torch.cuda.synchronize()
N = 20
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        ts.run()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
t0, t1 = None, None
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = re.sub(r"\(.*", "", ev.name)[:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        st = ev.time_range.start; en = ev.time_range.end
        t0 = st if t0 is None else min(t0, st); t1 = en if t1 is None else max(t1, en)
tot = sum(v[1] for v in agg.values())
out = {"steps": N, "graph": ts.graph is not None, "sum_kernel_us_per_step": tot / N, "span_us_per_step": (t1 - t0) / N,
       "kernels": [{"name": k, "launches_per_step": c / N, "us_per_step": t / N, "share": t / tot} for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out))
