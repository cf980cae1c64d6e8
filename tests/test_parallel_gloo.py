"""CPU, world_size 2, gloo: the data-parallel host logic (flat gradient bucket + single all-reduce + the
SUM/AVG scale rule + row sharding) reproduces the single-process gradients."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, reduction, out_q):
    sys.path.insert(0, os.path.join(ROOT, "hyperbolic-vae_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hvae.parallel import FlatGradBucket, shard_rows

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x = torch.randn(10, 6)
    lo, hi = shard_rows(10, rank, world)
    params = list(model.parameters())
    if reduction == "segments":
        # the overlap layout: the decoder-tail parameters (first to finish in backward) lead the buffer; the two
        # segment all-reduces together must equal the single one.  Report in the canonical parameter order.
        bucket = FlatGradBucket(params, early=params[2:])
        assert bucket.n_early == 2 and 0 < bucket.split < bucket.buffer.numel()
        assert [id(p) for p in bucket.params] == [id(p) for p in params[2:] + params[:2]]
    else:
        bucket = FlatGradBucket(params)
    bucket.zero_()
    y = model(x[lo:hi]).pow(2).sum(-1)
    (y.mean() if reduction == "mean" else y.sum()).backward()
    if reduction == "segments":
        bucket.all_reduce_segment("early", average=False)
        bucket.all_reduce_segment("late", average=False)
        flat = torch.cat([torch.cat([p.grad.flatten(), torch.zeros((-p.numel()) % 32)]) for p in params])
        out_q.put((rank, flat.tolist()))   # plain floats: a tensor would travel as a shared-memory fd the exiting child may close first
        dist.barrier()
        dist.destroy_process_group()
        return
    if reduction == "sum":
        bucket.all_reduce(average=False)
    else:
        # per-shard means with unequal shard sizes: weight by the shard's share of the global batch
        bucket.buffer.mul_((hi - lo) * world / 10.0)
        bucket.all_reduce(average=True)
    out_q.put((rank, bucket.buffer.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def _single(reduction):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x = torch.randn(10, 6)
    y = model(x).pow(2).sum(-1)
    (y.mean() if reduction == "mean" else y.sum()).backward()
    return torch.cat([(torch.cat([p.grad.flatten(), torch.zeros((-p.numel()) % 32)])) for p in model.parameters()])


def _free_port():
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(reduction, port=None):
    port = port or _free_port()   # a fixed port can still be in TIME_WAIT from an earlier run on the same host
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, reduction, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {k: torch.tensor(v) for k, v in (q.get(timeout=300) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _single(reduction)
    for r in range(2):
        torch.testing.assert_close(res[r], ref, rtol=1e-5, atol=1e-6)


def test_dp_sum_loss_matches_single_process():
    _run("sum")


def test_dp_mean_loss_matches_single_process():
    _run("mean")


def test_dp_early_late_segments_match_single_process():
    _run("segments")


def test_shard_rows_cover_batch():
    sys.path.insert(0, os.path.join(ROOT, "hyperbolic-vae_b200"))
    from hvae.parallel import shard_rows

    for n in (1, 7, 4096, 65536):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
