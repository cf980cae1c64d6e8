#!/usr/bin/env python
"""bench.py — train samples/s (forward + loss + backward) of the Poincare-VAE hot path on N B200s.

Workload at N=1 (BASELINE.json configs[1]): the pvae-replicate MNIST-shape graph — 784 -> 600 ReLU ->
MobiusLayer(600,10)+expmap0, sigma = softplus(Linear(600,1)); RiemannianNormal prior/posterior with the
HyperbolicRadius rejection sampler; GeodesicLayer(10,600) ReLU Linear(600,784); Bernoulli loss; batch 4096
per GPU, fp32, synthetic data, random-init weights.  One "step" = one forward + loss + backward over one
batch; for N>1 each rank owns its own 4096-row shard (weak scaling) and the step ends with ONE NCCL
all-reduce of the flat gradient bucket.

Prints ONE JSON line (rank 0).  `value` = device-timed whole-job samples/s with the batch resident in HBM
(CUDA-graph replay of the step, per-step CUDA events, L2 flushed between steps); `e2e` = the same step through
the public module API with HOST input: pinned H2D copy of the batch + eager step + D2H read of the loss inside
the timed region.  `roofline` = the dominant own kernel of the step against the measured HBM peak;
`cpu_baseline` = the oracle port of the reference's CPU path timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's CPU implementation of the same step (oracle port; /root/reference does
not travel and its geoopt/pvae dependencies are not installable) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "hyperbolic-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    "cfg2": dict(desc="pvae-replicate MNIST-shape, RiemannianNormal + HyperbolicRadius sampler, Bernoulli loss",
                 batch=4096, latent=10, hidden=600, c=1.0, data=(1, 28, 28)),
}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle port of the reference's CPU step
# ---------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(wl, steps, warmup, sample_rows):
    """samples/s of the reference's CPU path (oracle port: reference layers over the geoopt/pvae restatement,
    ARS sampler included) — all host threads, fp32, fwd + loss + bwd, no optimizer."""
    from oracle import ref_port as R

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    m = R.PvaeMnist(latent_dim=wl["latent"], hidden_dim=wl["hidden"], c=wl["c"], data_size=wl["data"])
    x = torch.rand(sample_rows, *wl["data"]).clamp(1e-5, 1 - 1e-5)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        m.zero_grad(set_to_none=True)
        out = m.loss(x)
        out["loss_total"].backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return sample_rows * len(times) / tot, tot / len(times)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.cpu_rows
    rate, per_step = cpu_reference_step_rate(wl, args.steps, max(args.warmup, 1), rows)
    line = {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd)", "value": rate, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: " + wl["desc"], "batch_per_gpu": wl["batch"], "latent_dim": wl["latent"],
                   "hidden": wl["hidden"], "curvature": wl["c"]},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d-row batches of the cfg2 step (oracle/ref_port.PvaeMnist, ARS sampler), %d timed steps"
                                   % (rows, args.steps)},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def build_model(wl, device):
    from hvae import models as HM
    torch.manual_seed(42)
    m = HM.PvaeMnist(latent_dim=wl["latent"], hidden_dim=wl["hidden"], c=wl["c"], data_size=wl["data"]).to(device)
    return m


def x3_traffic():
    """DRAM bytes per launch of the trunk GEMM from the committed ncu --set full capture (mean of the five launches)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r1_x3_traffic.json")))["traffic_bytes_per_launch_mean"]
    except Exception:
        return None


def kernel_roofline(wl, device, pk, pk_kind):
    """Time each own kernel of the cfg2 step alone (CUDA events per launch, L2 flushed between launches) and
    report the dominant one against the HBM roofline.  Algorithmic bytes per SURVEY.md §8(d)."""
    from hvae import _cabi as C
    from hvae import ops

    B, F, D, H = wl["batch"], wl["hidden"], wl["latent"], wl["hidden"]
    c = wl["c"]
    g = torch.Generator(device=device).manual_seed(1)
    x = torch.randn(B, F, device=device, generator=g)
    W = torch.randn(D, F, device=device, generator=g) * 0.05
    beta = torch.randn(D, device=device, generator=g) * 0.1
    _, M = ops.weight_prep_fwd(W, beta, c)
    y, mx = ops.mobius_matvec_fwd(x, M, c)
    gy = torch.randn_like(y)
    z = ops.expmap0(torch.randn(B, D, device=device, generator=g) * 0.3, c)
    Wg = torch.randn(H, D, device=device, generator=g) * 0.3
    bg = torch.randn(H, device=device, generator=g) * 0.1
    bpt, Mg = ops.weight_prep_fwd(Wg, bg, c)
    FL = ops.GYRO_PVAE | ops.GYRO_SIGNED
    out = ops.gyroplane_fwd(z, Mg, bpt, None, c, FL)
    gout = torch.randn_like(out)
    cases = {
        "mobius_matvec_fwd": (lambda: ops.mobius_matvec_fwd(x, M, c), 4 * (B * F + D * F + 2 * B * D)),
        "mobius_matvec_bwd": (lambda: ops.mobius_matvec_bwd(x, M, mx, gy, c), 4 * (2 * B * F + 2 * D * F + 3 * B * D)),
        "gyroplane_fwd": (lambda: ops.gyroplane_fwd(z, Mg, bpt, None, c, FL), 4 * (B * D + 2 * H * D + B * H)),
        "gyroplane_bwd": (lambda: ops.gyroplane_bwd(z, Mg, bpt, gout, c, FL, False), 4 * (B * H + 2 * B * D + 4 * H * D)),
    }
    # trunk dense layers: the five fp32-accurate tensor-core GEMMs of one step (enc fwd, dec fwd, dec dgrad, dec wgrad,
    # enc wgrad), timed as GEMM launches on pre-split operands.  Algorithmic flops = 2 M N K of the fp32 product; the
    # kernel executes 6x that in bf16 (three-way split, six piece products).
    n_in = int(torch.Size(wl["data"]).numel())
    gemms = [(B, H, n_in), (B, n_in, H), (B, H, n_in), (n_in, H, B), (H, n_in, B)]
    tc_ops = []
    for (m_, n_, k_) in gemms:
        a_s = ops.split3(torch.randn(m_, k_, device=device, generator=g))
        b_s = ops.split3(torch.randn(n_, k_, device=device, generator=g))
        tc_ops.append((a_s, b_s, m_, n_, k_))

    def run_gemms():
        for a_s, b_s, m_, n_, k_ in tc_ops:
            ops.gemm_x3s(a_s, False, b_s, False, None, False, m_, n_, k_)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    res = {}
    for name, (fn, nbytes) in cases.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e) * 1e-3)
        ts.sort()
        t = sum(ts[2:-2]) / len(ts[2:-2])
        res[name] = {"seconds": t, "bytes": nbytes, "gbs": nbytes / t / 1e9}
    for _ in range(3):
        run_gemms()
    ts = []
    for _ in range(20):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run_gemms()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    t_g = sum(ts[2:-2]) / len(ts[2:-2])
    fl = sum(2.0 * m_ * n_ * k_ for (m_, n_, k_) in gemms)
    n_g = sum(C.lib().hvae_gemm_x3s_num_launches(m_, n_, k_) for (m_, n_, k_) in gemms)
    top = max(res, key=lambda k: res[k]["seconds"])
    r = res[top]
    hbm_roof = {"bound": "hbm", "kernel": top, "achieved": r["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": r["gbs"] / pk["hbm_gbs"], "launch_us": r["seconds"] * 1e6, "algorithmic_bytes": r["bytes"],
                "note": "cfg2 sizes (~10 MB per kernel) are launch/latency-bound; large-row rooflines in kernel_rooflines"}
    # the dominant kernel of the step is the trunk GEMM (5 launches, ~1/3 of the step): tensor-bound
    # The kernel's work is the six bf16 piece products (6 x 2MNK tensor flops, none redundant): that is what is held
    # against the measured bf16 tensor peak.  The fp32 product it stands for (2MNK) is reported beside it.
    roof = {"bound": "tensor", "kernel": "k_tc_gemm<EPI_X3> (trunk dense layers, fp32-accurate split-bf16 GEMM)",
            "achieved": 6.0 * fl / t_g / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": 6.0 * fl / t_g / 1e12 / pk["bf16_tflops"],
            "traffic": x3_traffic(), "peak_source": pk_kind, "launch_us": t_g / 5 * 1e6, "launches_per_step": n_g,
            "algorithmic_flops_per_step": 6.0 * fl, "fp32_equivalent_tflops": fl / t_g / 1e12,
            "note": "achieved = bf16 tensor flops of the split algorithm (3 pieces per operand, 6 piece products = 6 x 2MNK, "
                    "summed over the step's five GEMMs) / their launch time, L2 flushed between repetitions; "
                    "fp32_equivalent_tflops = 2MNK / time (B200 fp32 FMA peak is ~72 TFLOP/s). cta_group::1 128x128 "
                    "MMAs are bound by shared-memory bandwidth at ~68 % of tensor peak in this schedule (DESIGN.md 4); "
                    "tile quantisation (160 / 224 tiles on 148 SMs) takes the rest. Largest HBM-bound own kernel in hbm_kernel.",
            "hbm_kernel": hbm_roof}
    others = {k: {"us": v["seconds"] * 1e6, "gbs": v["gbs"]} for k, v in res.items()}
    others["trunk_gemm_x3_5launches"] = {"us": t_g * 1e6, "tflops_algorithmic": fl / t_g / 1e12}
    # large-row HBM rooflines of the row kernels (inputs >> L2)
    big = {}
    for D_ in (2, 16, 64):
        Bb = (1 << 28) // (4 * D_)  # 256 MiB per tensor
        u = torch.randn(Bb, D_, device=device, generator=g) * 0.3
        mu = ops.expmap0(u, c)
        sg = torch.rand(Bb, D_, device=device, generator=g) + 0.3
        eps = torch.randn(Bb, D_, device=device, generator=g)
        for name, fn, nbytes in (
            ("expmap0_fwd", lambda: ops.expmap0_fwd(u, c), 8 * Bb * D_),
            ("latent_head_fwd", lambda: ops.latent_head_fwd(mu, sg, eps, 1.0, c), 16 * Bb * D_ + 4 * Bb),
        ):
            for _ in range(2):
                fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                fn()
            e.record()
            e.synchronize()
            t = s.elapsed_time(e) * 1e-3 / 5
            big["%s_D%d" % (name, D_)] = {"gbs": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / pk["hbm_gbs"], "rows": Bb}
        del u, mu, sg, eps
    return roof, others, big


def dbg(msg):
    if os.environ.get("HVAE_BENCH_DEBUG"):
        sys.stderr.write("[bench r%s %.1fs] %s\n" % (os.environ.get("RANK", "0"), time.perf_counter() % 1000, msg))
        sys.stderr.flush()


def run_ours(args, wl):
    import torch.distributed as dist

    from hvae import _cabi as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: the hvae path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    torch.backends.cuda.matmul.allow_tf32 = False  # fp32 trunk: parity contract is 1e-5 against the fp32 reference
    torch.backends.cudnn.allow_tf32 = False
    C.lib()
    dbg('init done')

    from hvae.train import TrainStep

    model = build_model(wl, device)
    B = wl["batch"]
    gen = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.rand(B, *wl["data"], generator=gen).clamp(1e-5, 1 - 1e-5).pin_memory()
    x_dev = x_host.to(device, non_blocking=True)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    # model loss is a batch SUM over the shard (App. A.2 vae_objective) -> SUM all-reduce, no rescale
    n0 = C.launch_count
    ts = TrainStep(model, x_dev, average_grads=False, use_graph=False)
    launches_per_step = (C.launch_count - n0) // 3  # TrainStep runs 3 eager warm-up steps
    dbg('eager warmup done')
    if not args.no_graph:
        ts._capture()
    graph_on = ts.graph is not None
    dbg('capture done graph=%s' % graph_on)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        ts.run()
    barrier()
    do_flush = os.environ.get("HVAE_BENCH_FLUSH", "1") != "0"
    clocks_on = rank == 0 or os.environ.get("HVAE_BENCH_CLOCKS_ALL", "0") == "1"  # one nvidia-smi poller per node

    class _NoClocks:
        samples = []

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def summary(self):
            return None

    with (Clocks(local) if clocks_on else _NoClocks()) as clk:
        evs = []
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            if do_flush:
                flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ts.run()          # batch already resident in HBM
            e.record()
            evs.append((s, e))
        barrier()
        t_wall = time.perf_counter() - t_wall0
        dbg('timed loop done')
        dev_s = sum(s.elapsed_time(e) for s, e in evs) * 1e-3

        # ---- e2e: host batch -> pinned H2D -> TrainStep.run (public API) -> D2H loss, every step ------------
        # Every step: one pinned H2D copy of that step's batch (prefetched on a copy stream while the previous step
        # runs), the step, and a D2H read of its loss that the host waits for.
        for _ in range(3):
            loss_host.copy_(ts.run(x_host), non_blocking=True)
        barrier()
        e2e_steps = args.steps
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ts.prefetch(x_host)                                   # batch of step 0
        for i in range(e2e_steps):
            loss_dev = ts.run_prefetched()
            if i + 1 < e2e_steps:
                ts.prefetch(x_host)                           # batch of step i+1 streams in under step i
            loss_host.copy_(loss_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()         # the caller reads the loss every step
        e.record()
        barrier()
        e2e_s = s.elapsed_time(e) * 1e-3
        dbg('e2e done')
        # nvidia-smi takes ~0.1 s per sample: keep the same step running (untimed, the SAME count on every
        # rank — each step holds a collective) so the clock record under load has a handful of samples
        for _ in range(8):
            for _ in range(50):
                ts.run()
            torch.cuda.synchronize()
    loss = float(loss_host)
    dbg('postroll done')
    t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = t.tolist()

    dbg('allreduce timing done')
    if rank == 0:
        pk, pk_kind = peaks()
        roof, others, big = kernel_roofline(wl, device, pk, pk_kind)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, per = cpu_reference_step_rate(wl, 3, 1, args.cpu_rows)
            cpu = {"value": rate, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "%d-row batches of the cfg2 step (oracle/ref_port.PvaeMnist incl. ARS sampler), 3 timed steps"
                             % args.cpu_rows}
        line = {
            "metric": "train samples/sec (fwd+bwd)", "value": world * B * args.steps / dev_s, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: " + wl["desc"], "batch_per_gpu": B, "latent_dim": wl["latent"], "hidden": wl["hidden"],
                       "curvature": wl["c"], "parallelism": "dp%d" % world, "cuda_graph": graph_on,
                       "l2": "256 MiB buffer written between timed steps (L2 flush)",
                       "trunk": "hvae.layers.Linear -> tcgen05 split-bf16 GEMM (3 pieces, 6 products, chunked fp32 accumulation): fp32-accurate, own kernel"},
            "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "samples/s",
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4, "steps": e2e_steps, "mode": "hvae.train.TrainStep prefetch()/run_prefetched(): H2D of step i+1 overlaps step i"},
            "gpu_launches": launches_per_step * (args.steps + e2e_steps),
            "gpu_launches_per_step": launches_per_step,
            "clocks": clk.summary(),
            "roofline": roof, "step_kernels_us": others, "kernel_rooflines": big,
            "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall,
            "loss": loss,
        }
        print(json.dumps(line))
    if world > 1:
        # Teardown: the captured graph holds NCCL work; destroying the process group under it can hang, so
        # synchronize, drop the graph, and leave without the collective teardown.
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        ts.graph = None
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-rows", type=int, default=1024, help="rows per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
