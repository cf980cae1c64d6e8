#!/usr/bin/env python
"""bench.py — train samples/s (forward + loss + backward) of the Poincare-VAE hot path on N B200s.

Workloads (BASELINE.json `configs`; `--workload`, default cfg2 = the configuration the metric is quoted on):
  cfg1   model A (VAEHyperbolicGyroplaneDecoder, scripts/_6): 784-d MNIST shape, batch 128, D = 2, c = 1
  cfg1b  model B (Mobius encoder + gyroplane decoder + MSE, scripts/_5/_7) on (1,32,32), batch 128, D = 2
  cfg2   pvae-replicate MNIST shape: 784 -> 600 ReLU -> MobiusLayer(600,10)+expmap0, sigma = softplus(Linear(600,1));
         RiemannianNormal prior/posterior with the HyperbolicRadius rejection sampler; GeodesicLayer(10,600) ReLU
         Linear(600,784); Bernoulli loss; batch 4096 per GPU
  cfg3   RNA-seq shape (VAEHyperbolicRNASeq, scripts/_8): 20 000 genes, batch 1024 per GPU, D = 5, hidden 100
  cfg4   model B, 8192 rows per GPU (65 536 over 8), D = 16 (`--latent`), c = 1 (`--curv`)
  cfg5   MobiusLayer + gyroplane microbench: 2^20 rows x 512 -> 4096, forward + backward, bf16 tensor-core mode
fp32 (cfg5: bf16 GEMM operands), synthetic data, random-init weights.  One "step" = one forward + loss + backward over
one batch; for N > 1 each rank owns its own batch (weak scaling) and the step ends with ONE exchange of the flat
gradient bucket: the repo's own all-reduce kernel over symmetric memory when the ranks can map each other - in-switch
reduction with multimem.ld_reduce / multimem.st when the NVSwitch offers a multicast mapping (`config.exchange` = "nvls"),
else peer loads / stores ("p2p") - else NCCL ("nccl").

Prints ONE JSON line (rank 0).  `value` = device-timed whole-job samples/s with the batch resident in HBM (CUDA-graph
replay of the step, per-step CUDA events, L2 flushed between steps); `e2e` = the same step through the public API with
HOST input: pinned H2D copy of every batch + the step + a D2H read of the loss inside the timed region.  `roofline` =
the dominant own kernel of the step; the cfg2 line also carries `tc_rooflines`: the tcgen05 Mobius / gyroplane kernels
at the config-5 shape (B = 2^20) with SURVEY 8(d) algorithmic flops.  `cpu_baseline` = the oracle port of the
reference's CPU path timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's CPU implementation of the same workload (oracle port; /root/reference does
not travel and its geoopt/pvae dependencies are not installable) on the host cores, at the batch it prints.
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
    "cfg1": dict(kind="step", model="A", batch=128, data=(1, 28, 28), latent=2, c=1.0, average=True,
                 desc="model A (gyroplane decoder, RelaxedBernoulli loss), MNIST shape 784-d, batch 128, D=2, c=1"),
    "cfg1b": dict(kind="step", model="B", batch=128, data=(1, 32, 32), latent=2, c=1.0, average=False,
                  desc="model B (Mobius encoder + gyroplane decoder, WrappedNormal, MSE), (1,32,32), batch 128, D=2, c=1"),
    "cfg2": dict(kind="step", model="pvae", batch=4096, data=(1, 28, 28), latent=10, hidden=600, c=1.0, average=False,
                 desc="pvae-replicate MNIST-shape, RiemannianNormal + HyperbolicRadius sampler, Bernoulli loss"),
    "cfg3": dict(kind="step", model="C", batch=1024, genes=20000, latent=5, hidden=100, c=1.0, beta=0.5, average=True,
                 desc="RNA-seq Jerby-Arnon shape, 20000 genes, gyroplane decoder, batch 1024 per GPU, D=5"),
    "cfg4": dict(kind="step", model="B", batch=8192, data=(1, 32, 32), latent=16, c=1.0, average=False,
                 desc="latent-dim / curvature grid point of scripts/_7: model B, 8192 rows per GPU (65536 over 8)"),
    "cfg5": dict(kind="layers", logB=20, F=512, P=4096, c=1.0,
                 desc="MobiusLayer + gyroplane microbench: 2^20 x 512 -> 4096, fwd+bwd, bf16 GEMM mode"),
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


class _NoClocks:
    samples = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def summary(self):
        return None


# ---------------------------------------------------------------------------------------------------
# workloads: the model (ours or the oracle port), its batch
# ---------------------------------------------------------------------------------------------------
def make_model(wl, mod):
    """mod: hvae.models (ours) or oracle.ref_port (the reference's CPU path); same constructors."""
    if wl["model"] == "A":
        return mod.ModelA(torch.Size(wl["data"]), wl["latent"], wl["c"], 1.0, 1.0)
    if wl["model"] == "B":
        return mod.ModelB(tuple(wl["data"]), wl["latent"], wl["c"], "mobius", "geoopt_gyroplane", 1.0, "mse")
    if wl["model"] == "C":
        return mod.ModelC(torch.Size([wl["genes"]]), wl["latent"], wl["c"], wl["hidden"], wl["beta"])
    if wl["model"] == "pvae":
        return mod.PvaeMnist(latent_dim=wl["latent"], hidden_dim=wl["hidden"], c=wl["c"], data_size=wl["data"])
    raise KeyError(wl["model"])


def make_batch(wl, rows, seed):
    gen = torch.Generator().manual_seed(seed)
    if wl["model"] == "C":
        # fake-data recipe of datasets/jerby_arnon.py:199-219 + z-score (:102-104)
        counts = torch.poisson(torch.full((rows, wl["genes"]), 100.0), generator=gen)
        return (counts - counts.mean(0, keepdim=True)) / counts.std(0, keepdim=True).clamp_min(1e-6)
    x = torch.rand(rows, *wl["data"], generator=gen)
    return x.clamp(1e-5, 1 - 1e-5) if wl["model"] == "pvae" else x


def config_of(wl, name, world, extra):
    cfg = {"workload": "%s: %s" % (name, wl["desc"]), "parallelism": "dp%d" % world}
    for k in ("batch", "latent", "hidden", "genes", "c"):
        if k in wl:
            cfg[{"batch": "batch_per_gpu", "latent": "latent_dim", "c": "curvature"}.get(k, k)] = wl[k]
    cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle port of the reference's CPU step
# ---------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(wl, steps, warmup, rows):
    """samples/s of the reference's CPU path (oracle port: reference layers over the geoopt/pvae restatement, ARS sampler
    included) — all host threads, fp32, fwd + loss + bwd, no optimizer — on `rows`-row batches."""
    from oracle import ref_port as R

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    m = make_model(wl, R)
    x = make_batch(wl, rows, 7)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        m.zero_grad(set_to_none=True)
        m.loss(x)["loss_total"].backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return rows * len(times) / tot, tot / len(times)


def cpu_reference_layers_rate(wl, rows_mobius, rows_gyro, reps):
    """cfg5 on the CPU: the reference's MobiusLayer and gyroplane layer forward + backward.  The reference's gyroplane
    broadcasts (B, D, P) intermediates (8.8 PB at B = 2^20), so it is timed on small row chunks and, like the Mobius
    layer, reported per row.  -> (rows/s of the pair, seconds per row of each)."""
    from oracle import ref_port as R
    from oracle.geoopt_min import PoincareBall as OBall
    from oracle.geoopt_min.layers.stereographic import Distance2StereographicHyperplanes as OGeo

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    F, P = wl["F"], wl["P"]
    ball = OBall(c=wl["c"])
    mob, gyr = R.MobiusLayer(F, P, ball), OGeo(F, P, ball=ball)

    def run(layer, rows):
        x = ball.expmap0(torch.randn(rows, F) * 0.1).detach().requires_grad_(True)
        g = torch.randn(rows, P)
        ts = []
        for i in range(reps + 1):
            t0 = time.perf_counter()
            layer.zero_grad(set_to_none=True)
            x.grad = None
            layer(x).backward(g)
            if i:
                ts.append(time.perf_counter() - t0)
        return sum(ts) / len(ts) / rows

    tm, tg = run(mob, rows_mobius), run(gyr, rows_gyro)
    return 1.0 / (tm + tg), tm, tg


def run_reference(args, name, wl):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    warm = max(args.warmup, 1)
    if wl["kind"] == "layers":
        logB = args.logB or wl["logB"]
        rate, tm, tg = cpu_reference_layers_rate(wl, 2048, 32, max(1, min(args.steps, 3)))
        per_step = (1 << logB) / rate
        sample = ("MobiusLayer fwd+bwd on 2048 rows and gyroplane fwd+bwd on 32-row chunks (the reference's (B,D,P) broadcast "
                  "does not fit more), per-row times %.3g s + %.3g s extrapolated to 2^%d rows" % (tm, tg, logB))
        batch_cfg = {"rows_per_gpu": 1 << logB, "F": wl["F"], "P": wl["P"]}
    else:
        rows = wl["batch"]  # the batch the line prints IS the batch that is timed
        rate, per_step = cpu_reference_step_rate(wl, args.steps, warm, rows)
        sample = "%d-row batches of the %s step (oracle/ref_port), %d timed steps" % (rows, name, args.steps)
        batch_cfg = {}
    line = {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd)", "value": rate, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(wl, name, int(os.environ.get("WORLD_SIZE", "1")), batch_cfg),
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# rooflines
# ---------------------------------------------------------------------------------------------------
def _time_flushed(fn, flush, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    k = max(1, len(ts) // 10)
    mid = ts[k:-k] if len(ts) > 2 * k else ts
    return sum(mid) / len(mid)


def trunk_roofline(gemms, device, pk, pk_kind, flush):
    """The fp32-accurate tensor-core GEMMs of one step, timed as GEMM launches on pre-split operands, next to cuBLAS
    fp32 (TF32 off) on the same shapes in the same run.  gemms: [(M, N, K)].  frac is on the ALGORITHMIC fp32 flops
    (2MNK); the kernel executes 3x that as fp16 piece products (executed_*)."""
    import ctypes

    from hvae import _cabi as C
    from hvae import ops

    g = torch.Generator(device=device).manual_seed(2)
    tc_ops, cb_ops, plans = [], [], []
    bn, sp, st = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    for (m_, n_, k_) in gemms:
        a, b = torch.randn(m_, k_, device=device, generator=g), torch.randn(n_, k_, device=device, generator=g)
        a_s, a_i, _, _ = ops.split2h_both(a, True, False)
        b_s, b_i, _, _ = ops.split2h_both(b, True, False)
        tc_ops.append((a_s, a_i, b_s, b_i, m_, n_, k_))
        cb_ops.append((a, b))
        C.lib().hvae_gemm_x2s_plan(m_, n_, k_, ctypes.addressof(bn), ctypes.addressof(sp), ctypes.addressof(st))
        plans.append({"tile": [128, bn.value], "split_k": sp.value, "stages": st.value})

    def run_x2():
        for a_s, a_i, b_s, b_i, m_, n_, k_ in tc_ops:
            ops.gemm_x2s(a_s, a_i, b_s, b_i, None, False, m_, n_, k_)

    def run_cublas():
        for a, b in cb_ops:
            torch.mm(a, b.t())

    t_g, t_c = _time_flushed(run_x2, flush), _time_flushed(run_cublas, flush)
    fl = sum(2.0 * m_ * n_ * k_ for (m_, n_, k_) in gemms)
    n_g = sum(C.lib().hvae_gemm_x2s_num_launches(m_, n_, k_) for (m_, n_, k_) in gemms)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2b_x2_traffic.json")   # ncu --set full capture of these five launches (committed)
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["traffic_bytes_per_launch_mean"], "profiles/r2b_x2_traffic.json (ncu capture, not this run)"
    return {"bound": "tensor", "kernel": "k_x2_gemm (trunk dense layers: fp32-accurate two-piece fp16 tcgen05 GEMM, power-of-two row scales)",
            "achieved": fl / t_g / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": fl / t_g / 1e12 / pk["bf16_tflops"],
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": sum((m_ + n_) * 2 * ((k_ + 63) // 64 * 64) * 2 + 4 * m_ * n_ for (m_, n_, k_) in gemms) / len(gemms),
            "peak_source": pk_kind, "launch_us": t_g / len(gemms) * 1e6, "launches_per_step": n_g,
            "algorithmic_flops_per_step": fl, "executed_tflops": 3.0 * fl / t_g / 1e12,
            "executed_frac": 3.0 * fl / t_g / 1e12 / pk["bf16_tflops"],
            "cublas_fp32_same_shapes": {"us": t_c * 1e6, "tflops": fl / t_c / 1e12, "ours_over_cublas_time": t_g / t_c,
                                        "note": "torch.mm fp32, TF32 off, same run, same shapes, L2 flushed"},
            "gemm_shapes_MNK": gemms, "plans": plans,
            "note": "achieved / frac = ALGORITHMIC fp32 flops (sum of 2MNK over the step's GEMMs) / launch time against the "
                    "measured 16-bit tensor peak; the kernel executes 3x that as fp16 piece products (two-piece operand split, "
                    "executed_*: what the tensor pipe actually does; the B200 fp32 FMA peak is ~72 TFLOP/s)"}


def row_kernel_rooflines(device, pk, c):
    """Large-row HBM fractions of the row kernels (inputs >> L2)."""
    from hvae import ops

    g = torch.Generator(device=device).manual_seed(3)
    big = {}
    for D_ in (2, 16, 64):
        Bb = (1 << 28) // (4 * D_)  # 256 MiB per tensor
        u = torch.randn(Bb, D_, device=device, generator=g) * 0.3
        mu = ops.expmap0(u, c)
        sg = torch.rand(Bb, D_, device=device, generator=g) + 0.3
        eps = torch.randn(Bb, D_, device=device, generator=g)
        z, kl = ops.latent_head_fwd(mu, sg, eps, 1.0, c)
        gz, gkl = torch.randn_like(z), torch.randn_like(kl)
        for name, fn, nbytes in (
            ("expmap0_fwd", lambda: ops.expmap0_fwd(u, c), 8 * Bb * D_),
            ("latent_head_fwd", lambda: ops.latent_head_fwd(mu, sg, eps, 1.0, c), 16 * Bb * D_ + 4 * Bb),
            ("latent_head_bwd", lambda: ops.latent_head_bwd(mu, sg, eps, gz, gkl, 1.0, c), 20 * Bb * D_ + 4 * Bb),
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
        del u, mu, sg, eps, z, kl, gz, gkl
    torch.cuda.empty_cache()
    return big


def cfg2_kernels(wl, device, pk, flush):
    """Each hyperbolic kernel of the cfg2 step alone (CUDA events per launch, L2 flushed), algorithmic bytes per SURVEY 8(d)."""
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
        "gyroplane_bwd": (lambda: ops.gyroplane_bwd(z, Mg, bpt, None, gout, c, FL, False), 4 * (B * H + 2 * B * D + 4 * H * D)),
    }
    res = {}
    for name, (fn, nbytes) in cases.items():
        t = _time_flushed(fn, flush)
        res[name] = {"seconds": t, "bytes": nbytes, "gbs": nbytes / t / 1e9}
    top = max(res, key=lambda k: res[k]["seconds"])
    r = res[top]
    hbm = {"bound": "hbm", "kernel": top, "achieved": r["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": r["gbs"] / pk["hbm_gbs"],
           "launch_us": r["seconds"] * 1e6, "algorithmic_bytes": r["bytes"],
           "note": "cfg2 sizes (~10 MB per kernel) are launch/latency-bound; large-row rooflines in kernel_rooflines"}
    return hbm, {k: {"us": v["seconds"] * 1e6, "gbs": v["gbs"]} for k, v in res.items()}


def tc_rooflines(device, pk, logB, F=512, P=4096, c_ctor=1.0, iters=4):
    """The north-star kernels: Mobius and gyroplane layers as tcgen05 GEMMs (bf16 operands) at the config-5 shape,
    timed per op with CUDA events; achieved = SURVEY 8(d) ALGORITHMIC flops (2BFP forward, 4BFP backward) / time."""
    import hvae
    from hvae import ops

    c = hvae.PoincareBall(c_ctor).c_value
    B = 1 << logB
    g = torch.Generator(device=device).manual_seed(0)
    x = ops.expmap0(torch.randn(B, F, device=device, generator=g) * 0.1, c)
    M = torch.randn(P, F, device=device, generator=g) / F ** 0.5
    pts = ops.expmap0(torch.randn(P, F, device=device, generator=g) * 0.03, c)
    fl = 2.0 * B * F * P

    def tm(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        e.synchronize()
        return s.elapsed_time(e) * 1e-3 / iters

    out = {}
    ops.set_gemm_mode("bf16")
    try:
        t = tm(lambda: ops.mobius_matvec_tc(x, M, c))
        out["mobius_fwd"] = {"ms": t * 1e3, "tflops": fl / t / 1e12, "frac": fl / t / 1e12 / pk["bf16_tflops"], "flops": "2BFP",
                             "hbm_gbs_algorithmic": 4.0 * (B * F + P * F + B * P) / t / 1e9}
        y, mxsq = ops.mobius_matvec_tc(x, M, c)
        gy = torch.randn_like(y)
        t = tm(lambda: ops.mobius_matvec_tc_bwd(x, M, y, mxsq, gy, c))
        out["mobius_bwd"] = {"ms": t * 1e3, "tflops": 2 * fl / t / 1e12, "frac": 2 * fl / t / 1e12 / pk["bf16_tflops"], "flops": "4BFP"}
        del y, gy
        t = tm(lambda: ops.gyroplane_tc_fwd(x, pts, None, c, ops.GYRO_SIGNED))
        out["gyroplane_fwd"] = {"ms": t * 1e3, "tflops": fl / t / 1e12, "frac": fl / t / 1e12 / pk["bf16_tflops"], "flops": "2BDP",
                                "hbm_gbs_algorithmic": 4.0 * (B * F + 2 * P * F + B * P) / t / 1e9}
        og = torch.randn(B, P, device=device, generator=g)
        t = tm(lambda: ops.gyroplane_tc_bwd(x, pts, og, c, ops.GYRO_SIGNED))
        out["gyroplane_bwd"] = {"ms": t * 1e3, "tflops": 2 * fl / t / 1e12, "frac": 2 * fl / t / 1e12 / pk["bf16_tflops"], "flops": "4BDP"}
        del og
        xb, Mb = x.bfloat16(), M.bfloat16()
        t = tm(lambda: torch.matmul(xb, Mb.t()))
        out["cublas_bf16_same_shape"] = {"ms": t * 1e3, "tflops": fl / t / 1e12, "note": "library GEMM, bf16 OUTPUT (half our write bytes)"}
    finally:
        ops.set_gemm_mode("fp32")
    out["shape"] = {"B": B, "F": F, "P": P}
    out["note"] = ("bf16 GEMM mode; algorithmic flops only (no credit for the backward's recompute GEMM or the Gram/row-dot "
                   "passes); each op includes its fp32->bf16 operand conversion and post-passes")
    del x, M, pts
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch.distributed as dist

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
    return dist, world, rank, local, device


def _finish(dist, world, local, graph_holder=None):
    if world > 1:
        # Teardown: the captured graph holds collective work; destroying the process group under it can hang, so
        # synchronize, drop the graph, and leave without the collective teardown.
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        if graph_holder is not None:
            graph_holder.graph = None
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_step(args, name, wl):
    from hvae import _cabi as C

    dist, world, rank, local, device = _dist_setup()
    C.lib()
    from hvae import models as HM
    from hvae.train import TrainStep

    torch.manual_seed(42)
    model = make_model(wl, HM).to(device)
    B = wl["batch"]
    x_host = make_batch(wl, B, 1000 + rank).pin_memory()
    x_dev = x_host.to(device, non_blocking=True)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    # batch-SUM losses (model B, pvae objective) -> SUM all-reduce reproduces the single-GPU gradient; batch-MEAN losses
    # (models A, C) -> each rank's mean is over its shard: SUM then divide by the world size
    n0 = C.launch_count
    opt = None
    if args.optimizer:
        from hvae.optim import RiemannianAdam

        opt = RiemannianAdam(model.parameters(), lr=1e-3)   # the reference's configure_optimizers (lr = 1e-3)
    ts = TrainStep(model, x_dev, average_grads=wl["average"], use_graph=False, optimizer=opt)
    launches_per_step = (C.launch_count - n0) // 3  # TrainStep runs 3 eager warm-up steps
    if not args.no_graph:
        ts._capture()
    graph_on = ts.graph is not None
    exchange = "none" if world == 1 else (("nvls" if ts.bucket.nvls else "p2p") if ts.bucket._symm is not None else "nccl")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        ts.run()
    barrier()
    with (Clocks(local) if rank == 0 else _NoClocks()) as clk:
        evs = []
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ts.run()          # batch already resident in HBM
            e.record()
            evs.append((s, e))
        barrier()
        t_wall = time.perf_counter() - t_wall0
        per_ms = sorted(s.elapsed_time(e) for s, e in evs)
        dev_s = sum(per_ms) * 1e-3
        # median and p10 / p90 of the per-step device times (this rank's)
        step_pct = {"p10": per_ms[int(0.10 * (len(per_ms) - 1))], "p50": per_ms[int(0.50 * (len(per_ms) - 1))],
                    "p90": per_ms[int(0.90 * (len(per_ms) - 1))]}
        # ---- e2e: host batch -> pinned H2D -> TrainStep (public API) -> D2H loss, every step; the H2D of step i+1 is
        # prefetched on a copy stream under step i, the host waits for every step's loss
        for _ in range(3):
            loss_host.copy_(ts.run(x_host), non_blocking=True)
        barrier()
        e2e_steps = args.steps
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ts.prefetch(x_host)
        for i in range(e2e_steps):
            loss_dev = ts.run_prefetched()
            if i + 1 < e2e_steps:
                ts.prefetch(x_host)
            loss_host.copy_(loss_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e.record()
        barrier()
        e2e_s = s.elapsed_time(e) * 1e-3
        # nvidia-smi takes ~0.1 s per sample: keep the same step running (untimed, the SAME count on every rank — each
        # step holds a collective) so the clock record under load has a handful of samples
        for _ in range(400):
            ts.run()
        torch.cuda.synchronize()
    loss = float(loss_host)
    t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = t.tolist()
    if rank == 0:
        pk, pk_kind = peaks()
        extra = {"cuda_graph": graph_on, "exchange": exchange, "grad_bucket_bytes": ts.bucket.nbytes,
                 "optimizer": "fused RiemannianAdam step inside the timed step (hvae.optim, one launch)" if opt is not None else
                              "none (the metric is fwd+bwd; --optimizer adds the fused Riemannian Adam step)",
                 "l2": "256 MiB buffer written between timed steps (L2 flush)",
                 "trunk": "hvae.layers.Linear -> tcgen05 two-piece fp16 GEMM with power-of-two row scales (fp32-accurate, own kernel) for GEMM-sized layers; "
                          "layers narrower than 64 or below 0.25 GFLOP (2MNK) and the conv stack of model B run on cuBLAS / cuDNN with TF32 off"}
        line = {
            "metric": "train samples/sec (fwd+bwd)", "value": world * B * args.steps / dev_s, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, name, world, extra),
            "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "steps": e2e_steps,
                    "mode": "hvae.train.TrainStep prefetch()/run_prefetched(): H2D of step i+1 overlaps step i"},
            "gpu_launches": launches_per_step * (args.steps + e2e_steps), "gpu_launches_per_step": launches_per_step,
            "clocks": clk.summary(), "wall_s_timed_region": t_wall, "loss": loss,
        }
        # ---- rooflines (rank 0, after the timed regions) ------------------------------------------------
        roof = None
        if wl["model"] == "pvae":
            n_in = int(torch.Size(wl["data"]).numel())
            H = wl["hidden"]
            gemms = [(B, H, n_in), (B, n_in, H), (B, H, n_in), (n_in, H, B), (H, n_in, B)]
            roof = trunk_roofline(gemms, device, pk, pk_kind, flush)
            roof["hbm_kernel"], line["step_kernels_us"] = cfg2_kernels(wl, device, pk, flush)
        elif wl["model"] == "C":
            G, H = wl["genes"], wl["hidden"]
            gemms = [(B, H, G), (B, G, H), (B, H, G), (G, H, B), (H, G, B)]
            roof = trunk_roofline(gemms, device, pk, pk_kind, flush)
        big = row_kernel_rooflines(device, pk, wl["c"]) if world == 1 else None
        if roof is None:
            k = "latent_head_fwd_D%d" % (2 if wl["latent"] <= 2 else 16)
            roof = {"bound": "hbm", "kernel": "k_latent_head_fwd (fused WrappedNormal sample + MC-KL) at 2^28/(4D) rows",
                    "achieved": big[k]["gbs"] if big else None, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": big[k]["frac"] if big else None, "traffic": None, "peak_source": pk_kind,
                    "note": "at this workload's sizes every own kernel moves < 1 MB: the step is launch/latency-bound (graph replay); "
                            "the HBM fraction is the same kernel at large row counts"}
        line["roofline"] = roof
        line["ms_per_step_percentiles"] = step_pct
        line["kernel_rooflines"] = big
        if world == 1 and name == "cfg2" and not args.no_tc_rooflines:
            line["tc_rooflines"] = tc_rooflines(device, pk, args.tc_logB)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rows = min(B, args.cpu_rows)
            rate, per = cpu_reference_step_rate(wl, 10, 2, rows)
            cpu = {"value": rate, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "%d-row batches of the %s step (oracle/ref_port), 10 timed steps" % (rows, name)}
        line["cpu_baseline"] = cpu
        print(json.dumps(line))
    _finish(dist, world, local, ts)


def run_layers(args, name, wl):
    """cfg5: MobiusLayer(512 -> 4096) and the gyroplane layer (D = 512, P = 4096) forward + backward on 2^logB rows per
    GPU in bf16 tensor-core mode; the parameter gradients (gM, gp: 8 MB each) are summed across ranks with NCCL."""
    from hvae import _cabi as C

    dist, world, rank, local, device = _dist_setup()
    C.lib()
    import hvae
    from hvae import ops

    c = hvae.PoincareBall(wl["c"]).c_value
    logB = args.logB or wl["logB"]
    B, F, P = 1 << logB, wl["F"], wl["P"]
    x_host = (torch.randn(B, F, generator=torch.Generator().manual_seed(100 + rank)) * 0.02).pin_memory()
    x = ops.expmap0(x_host.to(device) * 5.0, c)
    M = torch.randn(P, F, device=device, generator=torch.Generator(device=device).manual_seed(5)) / F ** 0.5
    pts = ops.expmap0(torch.randn(P, F, device=device, generator=torch.Generator(device=device).manual_seed(6)) * 0.03, c)
    gy = torch.randn(B, P, device=device, generator=torch.Generator(device=device).manual_seed(200 + rank))  # upstream gradient
    ops.set_gemm_mode("bf16")
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def step(xin):
        y, mxsq = ops.mobius_matvec_tc(xin, M, c)
        _, gM = ops.mobius_matvec_tc_bwd(xin, M, y, mxsq, gy, c)
        del y
        ops.gyroplane_tc_fwd(xin, pts, None, c, ops.GYRO_SIGNED)
        _, gp = ops.gyroplane_tc_bwd(xin, pts, gy, c, ops.GYRO_SIGNED)
        if world > 1:
            dist.all_reduce(gM)
            dist.all_reduce(gp)
        return gM, gp

    n0 = C.launch_count
    step(x)
    launches_per_step = C.launch_count - n0

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(x)
    barrier()
    steps = args.steps
    with (Clocks(local) if rank == 0 else _NoClocks()) as clk:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            step(x)
        e.record()
        barrier()
        dev_s = s.elapsed_time(e) * 1e-3
        # e2e: the batch comes from pinned host memory every step (2 GB H2D), a scalar of the result goes back
        xin = torch.empty_like(x)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            xin.copy_(x_host, non_blocking=True)
            gM, gp = step(xin)
            loss_host.copy_(gM[0, 0] + gp[0, 0], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e.record()
        barrier()
        e2e_s = s.elapsed_time(e) * 1e-3
    ops.set_gemm_mode("fp32")
    t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = t.tolist()
    if rank == 0:
        pk, pk_kind = peaks()
        fl = 12.0 * B * F * P   # 2BFP + 4BFP (Mobius) + 2BDP + 4BDP (gyroplane), D = F
        per = dev_s / steps
        del gy, xin
        torch.cuda.empty_cache()
        tcr = tc_rooflines(device, pk, logB) if world == 1 else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, tm, tg = cpu_reference_layers_rate(wl, 1024, 16, 2)
            cpu = {"value": rate, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "MobiusLayer fwd+bwd on 1024 rows + gyroplane fwd+bwd on 16-row chunks (the reference broadcasts "
                             "(B,D,P)), per-row times %.3g s + %.3g s" % (tm, tg)}
        line = {
            "metric": "train samples/sec (fwd+bwd)", "value": world * B * steps / dev_s, "unit": "samples/s", "n_gpus": world,
            "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_of(wl, name, world, {"rows_per_gpu": B, "F": F, "P": P, "exchange": "nccl" if world > 1 else "none",
                                                  "l2": "inputs and outputs (2-17 GB each) exceed the 126 MB L2"}),
            "e2e": {"value": world * B * steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "steps": steps},
            "gpu_launches": launches_per_step * 2 * steps, "gpu_launches_per_step": launches_per_step, "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "kernel": "tc2::k_tc_gemm2 (cta_group::2 tcgen05 GEMMs of the Mobius / gyroplane layers, fwd + bwd)",
                         "achieved": fl / per / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": fl / per / 1e12 / pk["bf16_tflops"],
                         "traffic": None, "peak_source": pk_kind,
                         "note": "ALGORITHMIC flops of the four ops (2BFP + 4BFP + 2BDP + 4BDP) / the whole step, conversions and "
                                 "post-passes included; per-op fractions in tc_rooflines"},
            "tc_rooflines": tcr, "cpu_baseline": cpu, "loss": float(loss_host),
        }
        print(json.dumps(line))
    _finish(dist, world, local)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--latent", type=int, default=None, help="override the workload's latent dim (cfg4 grid)")
    ap.add_argument("--curv", type=float, default=None, help="override the workload's curvature (cfg4 grid)")
    ap.add_argument("--hidden", type=int, default=None, help="override the workload's hidden width (cfg3: 100 or 512)")
    ap.add_argument("--logB", type=int, default=None, help="cfg5: log2 rows per GPU (default 20)")
    ap.add_argument("--tc-logB", type=int, default=20, help="rows (log2) of the tc_rooflines section of the cfg2 line")
    ap.add_argument("--cpu-rows", type=int, default=1024, help="rows per CPU-baseline step of our arm (bounded sample)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--optimizer", action="store_true", help="include the fused Riemannian Adam step in every timed step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tc-rooflines", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    for k, v in (("latent", args.latent), ("c", args.curv), ("hidden", args.hidden)):
        if v is not None and k in wl:
            wl[k] = v
    if wl["kind"] == "layers" and args.steps > 10 and args.impl == "ours":
        args.steps = 10   # a cfg5 step is ~30 ms of kernels on 60 GB of tensors: 10 steps are plenty
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    elif wl["kind"] == "layers":
        run_layers(args, args.workload, wl)
    else:
        run_step(args, args.workload, wl)


if __name__ == "__main__":
    main()
