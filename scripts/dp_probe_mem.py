#!/usr/bin/env python
"""2-rank probe: local access speed of a torch symmetric-memory buffer vs an ordinary one (memset, copy in, copy out)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 950_000
a = symm.empty(n, dtype=torch.float32, device=dev)
h = symm.rendezvous(a, dist.group.WORLD)
b = torch.empty(n, device=dev)
src = torch.randn(n, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {"multicast_ptr": bool(getattr(h, "multicast_ptr", 0))}


def t(fn, reps=50, fl=True):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        if fl:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 2)

for name, buf in (("symm", a), ("plain", b)):
    out[name + "_memset_us"] = t(lambda: buf.zero_())
    out[name + "_copy_in_us"] = t(lambda: buf.copy_(src))
    out[name + "_copy_out_us"] = t(lambda: src.copy_(buf))
    out[name + "_memset_warm_us"] = t(lambda: buf.zero_(), fl=False)
if rank == 0:
    print(json.dumps(out))
dist.barrier(device_ids=[local]); torch.cuda.synchronize(); os._exit(0)
