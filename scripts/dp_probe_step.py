#!/usr/bin/env python
"""N-rank probe (torchrun): where does the config-2 step's data-parallel overhead come from?  Times the graph-replayed
step (as bench.py does: per-step events, L2 flush) with (a) no exchange + ordinary bucket, (b) no exchange + symmetric
bucket, (c) the real exchange; max over ranks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import hvae.parallel as HP
    from hvae import models as HM
    from hvae.train import TrainStep

    x = torch.rand(4096, 1, 28, 28, generator=torch.Generator().manual_seed(1000 + rank)).clamp(1e-5, 1 - 1e-5).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}

    def timed(ts, steps=40):
        for _ in range(5):
            ts.run()
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ts.run()
            e.record()
            evs.append((s, e))
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        t = torch.tensor([sum(s.elapsed_time(e) for s, e in evs) / steps * 1e3], device=dev, dtype=torch.float64)
        tmin = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        return round(float(t), 1), round(float(tmin), 1)

    for mode in ("plain_noexchange", "symm_noexchange", "exchange_p2p", "exchange_nvls", "exchange_nccl"):
        torch.manual_seed(42)
        m = HM.PvaeMnist(latent_dim=10, hidden_dim=600).to(dev)
        os.environ["HVAE_DP_NVLS"] = "1" if mode == "exchange_nvls" else "0"
        os.environ["HVAE_DP_P2P"] = "0" if mode in ("plain_noexchange", "exchange_nccl") else "1"
        orig = HP.FlatGradBucket.all_reduce
        if mode.endswith("noexchange"):
            HP.FlatGradBucket.all_reduce = lambda self, average, group=None, async_op=False: None
            HP.FlatGradBucket.all_reduce_segment = lambda self, which, average, group=None: None
        ts = TrainStep(m, x, use_graph=True, average_grads=False)
        out[mode] = timed(ts)
        out[mode + "_symm"] = ts.bucket._symm is not None
        HP.FlatGradBucket.all_reduce = orig
        ts.graph = None
        del ts, m
    if rank == 0:
        print(json.dumps(out))
    dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
