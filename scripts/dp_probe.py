#!/usr/bin/env python
"""2-rank probe of the gradient exchange (run under torchrun): the bucket all-reduce alone, back to back, per path
(NVLS multimem / peer loop / NCCL) and grid size, on the config-2 bucket (3.8 MB) and a config-3-sized one (16 MB)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from hvae.parallel import FlatGradBucket

    out = {}
    for nelem in (4096, 950_000, 4_000_000):
        params = [torch.nn.Parameter(torch.zeros(nelem, device=dev))]
        for path in ("nvls", "p2p", "nccl"):
            os.environ["HVAE_DP_NVLS"] = "1" if path == "nvls" else "0"
            b = FlatGradBucket(params, symmetric=None if path != "nccl" else False)
            for blocks in ((8, 32, 128) if path == "nvls" else (32, 128) if path == "p2p" else (0,)):
                if blocks:
                    b.p2p_blocks = blocks
                if path == "nvls":
                    b.nvls_blocks = blocks
                for _ in range(5):
                    b.all_reduce(average=False)
                dist.barrier(device_ids=[local])
                torch.cuda.synchronize()
                # 50 exchanges captured in ONE CUDA graph: the replay is free of per-call host overhead
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    b.all_reduce(average=False)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    for _ in range(50):
                        b.all_reduce(average=False)
                g.replay()
                dist.barrier(device_ids=[local])
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(4):
                    g.replay()
                e.record()
                e.synchronize()
                out["%s_%d_b%d" % (path, nelem, blocks)] = round(s.elapsed_time(e) / 200 * 1e3, 2)
                del g
            del b
    if rank == 0:
        print(json.dumps(out))
    dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
