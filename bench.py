#!/usr/bin/env python
"""bench.py — train samples/s of the fused sequential-fusion step on B200.

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): MIMIC-shaped synthetic batch —
tabular (6) + time-series (99) + image-embedding (1024) MIMIC_MLPEncoder(64, F, (32, 32)) encoders,
2 x MLPDecoder(64, (32, 32), 2), state 64, dropout 0.2, err 1 / state-change 0.3, Adam 1e-3,
B = 65536 rows per GPU (weak scaling), fp32.  One step = MultiModN.train_epoch over one batch:
fused forward + backward kernel, gradient all-reduce (N > 1), fused Adam.

  python bench.py [--gpus N] [--steps K] [--warmup W]          (torchrun launches N ranks)
  python bench.py --impl reference ...                          CPU arm: the vectorised torch port
                                                                of the reference on the host cores

  python bench.py --workload c4_wide ...                        the wide regime (BASELINE configs[3], bf16) as a full line

Prints ONE JSON line (rank 0).  At N = 1 the default (C2) line also carries `wide_regime`: the config-4 train step on the
tcgen05 GEMM path, same metric, with its tensor-pipe roofline (skip with --no-wide).  `value`: inputs resident in HBM.  `e2e`: same metric through the
public API with pinned HOST inputs, H2D copies and a D2H read of the epoch's loss inside the timed
region.  `roofline`: the step kernel alone, timed with CUDA events on its stream.  `cpu_baseline`:
the port (oracle/torch_port.py) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# DRAM bytes of ONE launch of the step kernel at B = 65536 (ncu --set full; profiles/)
DRAM_TRAFFIC_PER_LAUNCH = {"fp32-fma": 709.7e6 + 252.2e6, "tcgen05-3xtf32": 701.6e6 + 238.2e6}

WORKLOADS = {
    # BASELINE.json configs[1] — the configuration the metric is quoted on (default)
    "c2_mimic": dict(name="c2_mimic", S=64, features=[6, 99, 1024], enc_hidden=(32, 32), n_decoders=2,
                     dec_hidden=(32, 32), dropout=0.2, err_penalty=1.0, state_change_penalty=0.3, lr=1e-3,
                     precision="fp32", batch=65536, ref_batch=8192),
    # BASELINE.json configs[3] — the wide regime: layer-wise tcgen05 bf16 GEMMs (--workload c4_wide)
    "c4_wide": dict(name="c4_wide", S=1024, features=[1024, 1024, 768, 768], enc_hidden=(2048, 2048), n_decoders=2,
                    dec_hidden=(2048,), dropout=0.0, err_penalty=1.0, state_change_penalty=0.3, lr=1e-4,
                    precision="bf16", batch=8192, ref_batch=512),
}
WORKLOAD = WORKLOADS["c2_mimic"]


def macs_per_row(w):
    S, D, E = w["S"], w["n_decoders"], len(w["features"])
    enc = 0
    for F in w["features"]:
        dims = [F + S, *w["enc_hidden"], S]
        enc += sum(a * b for a, b in zip(dims, dims[1:]))
    dd = [S, *w["dec_hidden"], 2]
    dec = sum(a * b for a, b in zip(dd, dd[1:])) * D * (E + 1)
    return enc + dec


def make_spec(seed=1):
    from oracle.spec_io import random_spec
    w = WORKLOAD
    return random_spec(np.random.default_rng(seed), w["S"], w["features"], enc_kind="mimic",
                       enc_hidden=w["enc_hidden"], dropout=w["dropout"], n_decoders=w["n_decoders"],
                       dec_hidden=w["dec_hidden"])


def make_batch(rng, B, pin=False, device=None):
    xs = [torch.from_numpy(rng.standard_normal((B, F), dtype=np.float32)) for F in WORKLOAD["features"]]
    y = torch.from_numpy((rng.random((B, WORKLOAD["n_decoders"])) < 0.3).astype(np.int64))
    if pin:
        xs, y = [x.pin_memory() for x in xs], y.pin_memory()
    if device is not None:
        xs, y = [x.to(device) for x in xs], y.to(device)
    return xs, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=5)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx or None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0), "fallback"


def cpu_port_rate(steps, warmup, B):
    """samples/s of the vectorised torch port on the host cores (bounded sample of the workload)."""
    from oracle.torch_port import TorchPort
    w = WORKLOAD
    torch.set_num_threads(os.cpu_count() or 1)
    port = TorchPort(make_spec(), w["err_penalty"], 0.01 * w["state_change_penalty"], lr=w["lr"])
    rng = np.random.default_rng(123)
    xs, y = make_batch(rng, B)
    for _ in range(warmup):
        port.train_step(xs, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        port.train_step(xs, y)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps


def wide_regime_summary(dev, steps=10):
    """Secondary line (N = 1 only): BASELINE.json configs[3], the wide regime — state 1024, hidden 2048, bf16 — whose layers
    are real dense contractions and run on the hand-written TMA + tcgen05 GEMM.  Same metric, measured the same way
    (`python bench.py --workload c4_wide` prints it as a full bench line with e2e and the CPU arm)."""
    from torch.nn import CrossEntropyLoss
    from multimodn_b200 import FusedAdam
    from model_utils import model_from_spec
    global WORKLOAD
    saved = WORKLOAD
    try:
        WORKLOAD = w = WORKLOADS["c4_wide"]
        B = w["batch"]
        model = model_from_spec(make_spec(3), w["err_penalty"], w["state_change_penalty"], dev, "row", precision="bf16")
        opt = FusedAdam(model, lr=w["lr"])
        rt = model.runtime()
        rng = np.random.default_rng(7)
        batches = [make_batch(rng, B, device=dev) for _ in range(2)]
        crit = CrossEntropyLoss()

        def timed(fn, n):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        step = lambda i: model.train_epoch([batches[i % 2]], opt, crit)  # noqa: E731
        for i in range(3):
            step(i)
        l0 = int(rt.lib.dll.mmn_wide_launch_count())
        ms = timed(step, steps)
        launches = (int(rt.lib.dll.mmn_wide_launch_count()) - l0) // steps
        peaks, kind = measured_peaks()
        flops = 6.0 * macs_per_row(w) * B
        tf = flops / (ms * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        return dict(workload="c4_wide", metric="train samples/sec", value=B / (ms * 1e-3), unit="samples/s", ms_per_step=ms,
                    dtype="bf16", batch_per_gpu=B, state_size=w["S"], features=w["features"], enc_hidden=list(w["enc_hidden"]),
                    dec_hidden=list(w["dec_hidden"]), gpu_launches_per_step=launches + 2,
                    roofline=dict(bound="tensor", achieved=tf, peak=peak, unit="TFLOP/s", frac=tf / peak, peak_source=kind,
                                  note="whole train step incl. Adam; algorithmic FLOPs = 6 x MACs"))
    except Exception as exc:  # noqa: BLE001   (the headline line must not depend on the secondary one)
        return dict(workload="c4_wide", error=f"{type(exc).__name__}: {exc}")
    finally:
        WORKLOAD = saved


def workload_config(B, world, **extra):
    w = WORKLOAD
    cfg = dict(workload=w["name"], state_size=w["S"], features=w["features"], enc_hidden=list(w["enc_hidden"]),
               decoders=w["n_decoders"], dec_hidden=list(w["dec_hidden"]), dropout=w["dropout"], precision=w["precision"],
               batch_per_gpu=B, global_batch=B * world, missing_mode="row", parallelism=f"dp{world}")
    cfg.update(extra)
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.ref_batch
    rate, sec = cpu_port_rate(args.steps, args.warmup, B)
    cores = torch.get_num_threads()
    line = dict(impl="reference", metric="train samples/sec", value=rate, unit="samples/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="fp32", data="synthetic",
                config=workload_config(args.batch, max(1, args.gpus), optimizer="torch.optim.Adam",
                                       device="host CPU", rows_per_timed_step=B),
                cpu_baseline=dict(value=rate, unit="samples/s", cores=cores, kind="port",
                                  sample=f"{args.steps} train steps of {B} rows (vectorised torch port of the "
                                         f"reference, oracle/torch_port.py, {cores} threads)"),
                e2e=dict(value=rate, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_mimic", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="rows per GPU per step (default: the workload's)")
    ap.add_argument("--ref-batch", type=int, default=None, help="rows per step of the CPU arm")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-wide", action="store_true", help="skip the secondary wide-regime (config 4) measurement")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = WORKLOADS[args.workload]
    args.batch = args.batch or WORKLOAD["batch"]
    args.ref_batch = args.ref_batch or WORKLOAD["ref_batch"]
    if args.impl == "reference":
        return run_reference(args)

    from torch.nn import CrossEntropyLoss
    from multimodn_b200 import FusedAdam, MultiModNHistory
    from model_utils import model_from_spec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    w = WORKLOAD
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    torch.manual_seed(1)
    model = model_from_spec(make_spec(), w["err_penalty"], w["state_change_penalty"], dev, "row", precision=w["precision"])
    if world > 1:
        model.enable_data_parallel()
    opt = FusedAdam(model, lr=w["lr"])
    crit = CrossEntropyLoss()
    rt = model.runtime()
    rng = np.random.default_rng(100 + rank)
    n_resident = 4                                     # distinct resident batches, each > L2 (297 MB)
    resident = [make_batch(rng, B, device=dev) for _ in range(n_resident)]
    bytes_in = B * (sum(w["features"]) * 4 + w["n_decoders"] * 8)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- value: inputs resident in HBM --------------------------------------------------------
    def step_resident(i):
        xs, y = resident[i % n_resident]
        model.train_epoch([(xs, y)], opt, crit)

    for i in range(W):
        step_resident(i)
    clk = ClockSampler(local)
    clk.__enter__()                                   # sampled over the value and kernel-only timed regions
    ms = timed(step_resident, K)
    ms_per_step = ms / K
    value = B * world * K / (ms * 1e-3)

    # ---- roofline: the step kernel alone (memset + fused fwd/bwd launch), events on its stream --
    seq = [(i, i) for i in range(len(w["features"]))]
    metrics = rt.new_metrics()

    def kernel_only(i):
        xs, y = resident[i % n_resident]
        mb, keep, n = rt.prepare_batch(xs, y, seq, "row", None)
        rt.train_step(mb, n, w["err_penalty"], 0.01 * w["state_change_penalty"], True, metrics)

    for i in range(3):
        kernel_only(i)
    wide_l0 = int(rt.lib.dll.mmn_wide_launch_count())
    kms = timed(kernel_only, K) / K
    wide_launches_per_step = (int(rt.lib.dll.mmn_wide_launch_count()) - wide_l0) // K
    t_load = time.perf_counter()                      # keep the load on until nvidia-smi has a few samples in
    while len(clk.rows) < 6 and time.perf_counter() - t_load < 4.0:
        timed(kernel_only, K)
    clk.__exit__()
    clocks = clk.summary()
    peaks, peak_kind = measured_peaks()
    engine = {0: "fp32-fma", 1: "tcgen05-3xtf32", 2: "tcgen05-3xtf32-tmem", 3: "tcgen05-bf16-layerwise"}[
        int(rt.lib.dll.mmn_plan_engine(rt.plan))]
    macs = macs_per_row(w)
    alg_bytes = B * (2 * 4 * sum(w["features"]) + 8 * w["n_decoders"])   # x read in fwd and again for wgrad
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    flops = 6.0 * macs * B                                               # fwd + dgrad + wgrad
    fma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    roofline = dict(bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"],
                    traffic=DRAM_TRAFFIC_PER_LAUNCH.get(engine) if B == 65536 else None,
                    traffic_source="ncu --set full dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_*_step_kernel_ncu.txt",
                    kernel=f"mmn_step_kernel<{engine}, train>", kernel_ms=kms, peak_source=peak_kind,
                    algorithmic_bytes_per_sample=alg_bytes / B,
                    fp32_fma=dict(achieved_tflops=flops / (kms * 1e-3) / 1e12, peak_tflops_nominal=fma_peak,
                                  frac=flops / (kms * 1e-3) / 1e12 / fma_peak,
                                  note="fp32 parity mode is FP32-FMA-bound, not HBM-bound (SURVEY.md 8d)"))

    if w["precision"] == "bf16":
        # the wide regime is tensor-pipe-bound: achieved = algorithmic FLOPs of one step (6 x MACs) over
        # the duration of the step's launches (GEMMs + their elementwise companions); peak = the measured SUSTAINED
        # dense bf16 throughput, because the launches are timed inside a long step
        tf = flops / (kms * 1e-3) / 1e12
        peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        roofline = dict(bound="tensor", achieved=tf, peak=peak_tf, unit="TFLOP/s", frac=tf / peak_tf, traffic=None,
                        kernel="mmn_wide_gemm_kernel (+ elementwise companions) x %d launches per step" % wide_launches_per_step,
                        kernel_ms=kms, peak_source=peak_kind, algorithmic_flops_per_sample=6.0 * macs,
                        hbm=dict(achieved_gbs=achieved, peak_gbs=peaks["hbm_gbs"], frac=achieved / peaks["hbm_gbs"]))

    # ---- e2e: the public call on a loader of pinned HOST batches: every step copies its inputs H2D (one batch
    #      ahead, on a side stream) and reads its loss metrics back D2H (log_interval=1) ----------------------
    host = [make_batch(rng, B, pin=True) for _ in range(3)]
    hist = MultiModNHistory(["a", "b"])
    Ke = max(3, min(K, 12))
    logged = []

    def epoch_e2e(n):
        model.train_epoch([host[i % 3] for i in range(n)], opt, crit, hist, log_interval=1, logger=logged.append)

    epoch_e2e(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    epoch_e2e(Ke)
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ems], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
    assert len(logged) == 3 + Ke
    e2e = dict(value=B * world * Ke / (ems * 1e-3), unit="samples/s", h2d_bytes_per_step=bytes_in,
               d2h_bytes_per_step=rt.n_metrics * 8, steps=Ke, ms_per_step=ems / Ke,
               call="MultiModN.train_epoch(loader of pinned host batches, FusedAdam, CrossEntropyLoss, history, log_interval=1)")

    # ---- CPU baseline: the port on the host cores, bounded sample ------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = args.ref_batch
        _, sec = cpu_port_rate(1, 1, Bc)
        n = max(2, min(50, int(args.cpu_seconds / max(sec, 1e-3))))
        rate, sec = cpu_port_rate(n, 0, Bc)
        cpu = dict(value=rate, unit="samples/s", cores=torch.get_num_threads(), kind="port",
                   sample=f"{n} train steps of {Bc} rows (vectorised torch port of the reference, "
                          f"oracle/torch_port.py; host has {os.cpu_count()} logical cores)")

    wide = None
    if rank == 0 and world == 1 and w["name"] == "c2_mimic" and not args.no_wide:
        del resident, host
        torch.cuda.empty_cache()
        wide = wide_regime_summary(dev)

    if rank == 0:
        line = dict(metric="train samples/sec", value=value, unit="samples/s", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=w["precision"],
                    data="synthetic",
                    config=workload_config(B, world, optimizer="FusedAdam", engine=engine,
                                           l2_policy=f"{n_resident} resident batches of {bytes_in / 1e6:.0f} MB each (> 126 MB L2), cycled"),
                    clocks=clocks, e2e=e2e,
                    gpu_launches=(wide_launches_per_step + 2) * K if w["precision"] == "bf16" else 3 * K,
                    roofline=roofline, cpu_baseline=cpu)
        if wide is not None:
            line["wide_regime"] = wide
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
