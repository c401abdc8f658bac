#!/usr/bin/env python
"""bench.py — train samples/s of the fused sequential-fusion step on B200, on the five BASELINE.json configurations.

  python bench.py [--gpus N] [--steps K] [--warmup W]          (torchrun launches N ranks; one JSON line from rank 0)
  python bench.py --impl reference ...                          CPU arm: the reference's path on the host cores
  python bench.py --workload c2_mimic|c3_mnar|c4_wide ...       another configuration as the headline line

Headline workload = BASELINE.json configs[3] (`c4_wide`: state 1024, hidden 2048, bf16, 8192 rows per GPU — the largest
single-GPU configuration, and the one BASELINE defines as data-parallel on 8 x B200).  One step = MultiModN.train_epoch over
one batch: forward + backward (layer-wise TMA + tcgen05 GEMMs), gradient all-reduce (N > 1, per encoder block, overlapped),
fused Adam.  The line also carries `configs`: every other BASELINE configuration measured the same way —
  c1_titanic  Titanic MLP model, B = 2^20 (HBM-bound)              fp32, fused per-tile kernel
  c2_mimic    MIMIC-shaped (6 / 99 / 1024 -> state 64), B = 65536   fp32 (FP32-FMA kernel) AND bf16 (per-tile mma kernel)
  c3_mnar     8 encoders, 6 decoders, state 256, 30 % MNAR, B = 65536   fp32
  c5_sweep    predict over 16 permuted encoding sequences, N = 2^20 rows   fp32
each with `value`, `roofline` (binding roof named) and, for c2 / c3 / c4, `e2e` + `cpu_baseline`.

`value`: inputs resident in HBM.  `e2e`: the same metric through the public API with pinned HOST inputs, H2D copies and a
D2H read of the step's loss inside the timed region.  `roofline`: the step's kernel(s) alone, CUDA events on their stream.
`cpu_baseline`: `ref_vec` = the vectorised torch port (oracle/torch_port.py) and `ref_asis` = the UNMODIFIED reference
(oracle/_ref) on a bounded sample, both on the box's host cores.
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

WORKLOADS = {
    # BASELINE.json configs[0]: pipelines/titanic/titanic_mlp_pipeline.py:26-76 — MLPEncoder(1, 6, (5, 5)), LogisticDecoder(1)
    "c1_titanic": dict(name="c1_titanic", S=1, features=[6], enc_kind="mlp", enc_hidden=(5, 5), n_decoders=1, dec_hidden=None,
                       dropout=0.0, err_penalty=0.7, state_change_penalty=0.3, lr=1e-2, precision="fp32", batch=1 << 20,
                       ref_batch=1 << 16, p_pos=0.4, mnar=False),
    # BASELINE.json configs[1]: pipelines/mimic/mimic_multi_task_pipeline.py:63-83,118-120 with datasets/mimic/mimic_dataset.py:19-21
    "c2_mimic": dict(name="c2_mimic", S=64, features=[6, 99, 1024], enc_kind="mimic", enc_hidden=(32, 32), n_decoders=2,
                     dec_hidden=(32, 32), dropout=0.2, err_penalty=1.0, state_change_penalty=0.3, lr=1e-3,
                     precision="fp32", batch=65536, ref_batch=65536, p_pos=0.3, mnar=False),
    # BASELINE.json configs[2]: MNAR stress
    "c3_mnar": dict(name="c3_mnar", S=256, features=[6, 99, 242, 110, 768, 768, 1024, 1024], enc_kind="mimic",
                    enc_hidden=(32, 32), n_decoders=6, dec_hidden=(32, 32), dropout=0.2, err_penalty=1.0,
                    state_change_penalty=0.3, lr=1e-3, precision="fp32", batch=65536, ref_batch=16384, p_pos=0.5, mnar=True),
    # BASELINE.json configs[3]: the wide regime
    "c4_wide": dict(name="c4_wide", S=1024, features=[1024, 1024, 768, 768], enc_kind="mimic", enc_hidden=(2048, 2048),
                    n_decoders=2, dec_hidden=(2048,), dropout=0.0, err_penalty=1.0, state_change_penalty=0.3, lr=1e-4,
                    precision="bf16", batch=8192, ref_batch=8192, p_pos=0.3, mnar=False),
}
HEADLINE = "c4_wide"


# ------------------------------------------------------------------------------------------------------------------
# models and synthetic batches (product classes only: nothing here touches oracle/ or tests/)
# ------------------------------------------------------------------------------------------------------------------
def build_model(w, dev, precision=None, seed=1):
    from multimodn_b200 import MultiModN
    from multimodn_b200.decoders import LogisticDecoder, MLPDecoder
    from multimodn_b200.encoders import MIMIC_MLPEncoder, MLPEncoder
    torch.manual_seed(seed)
    S = w["S"]
    if w["enc_kind"] == "mimic":
        encs = [MIMIC_MLPEncoder(S, F, tuple(w["enc_hidden"]), dropout=w["dropout"]) for F in w["features"]]
    else:
        encs = [MLPEncoder(S, F, tuple(w["enc_hidden"])) for F in w["features"]]
    if w["dec_hidden"] is None:
        decs = [LogisticDecoder(S) for _ in range(w["n_decoders"])]
    else:
        decs = [MLPDecoder(S, tuple(w["dec_hidden"]), 2) for _ in range(w["n_decoders"])]
    return MultiModN(S, encs, decs, w["err_penalty"], w["state_change_penalty"], device=dev, missing_mode="row",
                     precision=precision or w["precision"])


def make_batch(w, B, seed, device=None, pin=False, with_missing=None):
    """features ~ N(0, 1) fp32, targets Bernoulli; c3: whole-modality NaN per row, P(miss | y0 = 1) = 0.5, P(miss | y0 = 0)
    = 0.1 (pipelines/mimic/mimic_single_task_mnar_missingness_pipeline.py:142-149, generalised) -> 30 % of the cells."""
    dev = device if device is not None else torch.device("cpu")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    y = (torch.rand((B, w["n_decoders"]), generator=g, device=dev) < w["p_pos"]).to(torch.int64)
    xs = []
    mnar = w["mnar"] if with_missing is None else with_missing
    for F in w["features"]:
        x = torch.randn((B, F), generator=g, device=dev, dtype=torch.float32)
        if mnar:
            p = torch.where(y[:, 0] == 1, 0.5, 0.1)
            miss = torch.rand((B,), generator=g, device=dev) < p
            x[miss] = float("nan")
        xs.append(x)
    if pin:
        xs, y = [x.pin_memory() for x in xs], y.pin_memory()
    return xs, y


def macs_per_row(w, strict=False):
    """forward multiply-accumulates per sample.  strict: (fwd + dgrad + wgrad) / 3 with the first-layer
    data gradient w.r.t. x (never computed) removed from the train count."""
    S, D, E = w["S"], w["n_decoders"], len(w["features"])
    enc = skip = 0
    for F in w["features"]:
        if w["enc_kind"] == "mimic":
            dims = [F + S, *w["enc_hidden"], S]
            enc += sum(a * b for a, b in zip(dims, dims[1:]))
            skip += F * dims[1]
        else:
            dims = [F, *w["enc_hidden"]]
            enc += sum(a * b for a, b in zip(dims, dims[1:])) + (dims[-1] + S) * S
            skip += F * (dims[1] if len(dims) > 1 else S)
    dd = [S, *(w["dec_hidden"] or ()), 2]
    dec = sum(a * b for a, b in zip(dd, dd[1:])) * D * (E + 1)
    fwd = enc + dec
    return (3 * fwd - skip) / 3.0 if strict else fwd


def train_bytes_per_row(w):
    """compulsory HBM bytes per sample of a train step: x read in the forward and again for the first-layer weight
    gradient, int64 targets"""
    return 2 * 4 * sum(w["features"]) + 8 * w["n_decoders"]


# ------------------------------------------------------------------------------------------------------------------
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
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), "fallback"


def measure_fma_peak(dev):
    """FP32-FMA micro-benchmark (SURVEY.md 8d asks for a measured FMA peak): mmn_selftest_fma_peak runs 8 independent FFMA
    chains per thread on every SM; TFLOP/s from CUDA events."""
    from multimodn_b200 import _lib
    import ctypes as C
    lib = _lib.get_lib()
    out = torch.zeros(4, dtype=torch.float32, device=dev)
    iters = 1 << 15
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    flops = C.c_double(0)
    best = 0.0
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.check(lib.dll.mmn_selftest_fma_peak(iters, out.data_ptr(), C.byref(flops), stream))
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


# ------------------------------------------------------------------------------------------------------------------
# CPU arms (the only place bench.py touches oracle/): ref_vec = vectorised torch port, ref_asis = unmodified reference
# ------------------------------------------------------------------------------------------------------------------
def _spec_of(model):
    from oracle.spec_io import spec_from_modules
    return spec_from_modules(model)


def cpu_port_rate(w, steps, warmup, B, seed=123):
    """samples/s of the vectorised torch port on the host cores"""
    from oracle.torch_port import TorchPort
    torch.set_num_threads(os.cpu_count() or 1)
    spec = _cpu_spec(w)
    port = TorchPort(spec, w["err_penalty"], 0.01 * w["state_change_penalty"], lr=w["lr"])
    xs, y = make_batch(w, B, seed)
    for _ in range(warmup):
        port.train_step(xs, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        port.train_step(xs, y)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps


def _cpu_spec(w, seed=1):
    """random-init weights of the workload's architecture as an oracle spec (built from the reference-shaped torch modules
    on the CPU; the product MultiModN class is not instantiated: it refuses non-CUDA devices when used)"""
    from oracle.spec_io import random_spec
    return random_spec(np.random.default_rng(seed), w["S"], w["features"], enc_kind=w["enc_kind"],
                       enc_hidden=tuple(w["enc_hidden"]), dropout=w["dropout"], n_decoders=w["n_decoders"],
                       dec_hidden=tuple(w["dec_hidden"] or ()))


def cpu_asis_rate(w, B, max_seconds=40.0, seed=123):
    """samples/s of the UNMODIFIED reference (oracle/_ref): MultiModN.train_epoch over one NaN-free batch of B rows with
    torch.optim.Adam, all host threads.  None when the reference copy is not present."""
    from oracle.ref_live import load_reference
    ref = load_reference()
    if ref is None:
        return None
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1)
    S = w["S"]
    if w["enc_kind"] == "mimic":
        encs = [ref.MIMIC_MLPEncoder(S, Fe, tuple(w["enc_hidden"]), dropout=w["dropout"], activation=F.relu) for Fe in w["features"]]
    else:
        encs = [ref.MLPEncoder(S, Fe, tuple(w["enc_hidden"]), F.relu) for Fe in w["features"]]
    if w["dec_hidden"] is None:
        decs = [ref.LogisticDecoder(S) for _ in range(w["n_decoders"])]
    else:
        decs = [ref.MLPDecoder(S, tuple(w["dec_hidden"]), 2) for _ in range(w["n_decoders"])]
    model = ref.MultiModN(S, encs, decs, w["err_penalty"], w["state_change_penalty"], device=torch.device("cpu"))
    opt = torch.optim.Adam(model.parameters(), lr=w["lr"])
    crit = torch.nn.CrossEntropyLoss()
    xs, y = make_batch(w, B, seed, with_missing=False)
    loader = [(xs, y)]
    t0 = time.perf_counter()
    model.train_epoch(loader, opt, crit)                       # warm-up step (also bounds the run)
    first = time.perf_counter() - t0
    n = max(1, min(5, int(max_seconds / max(first, 1e-3)) - 1))
    t0 = time.perf_counter()
    for _ in range(n):
        model.train_epoch(loader, opt, crit)
    dt = (time.perf_counter() - t0) / n
    return dict(value=B / dt, unit="samples/s", cores=torch.get_num_threads(), kind="reference",
                sample=f"{n} train_epoch calls of one {B}-row NaN-free batch after 1 warm-up (unmodified reference, "
                       f"oracle/_ref, torch.optim.Adam; host has {os.cpu_count()} logical cores)", ms_per_step=dt * 1e3)


def cpu_baseline(w, seconds, asis_rows):
    """both CPU timers for one workload: ref_vec at the workload's own batch, ref_asis on a reduced batch (its CPython
    element loops cost ~1 us per feature element)"""
    Bc = w["ref_batch"]
    _, sec = cpu_port_rate(w, 1, 1, Bc)
    n = max(2, min(30, int(seconds / max(sec, 1e-3))))
    rate, sec = cpu_port_rate(w, n, 0, Bc)
    out = dict(value=rate, unit="samples/s", cores=torch.get_num_threads(), kind="port", ms_per_step=sec * 1e3,
               rows_per_step=Bc,
               sample=f"{n} train steps of {Bc} rows (ref_vec: vectorised torch port of the reference, oracle/torch_port.py; "
                      f"host has {os.cpu_count()} logical cores)")
    if asis_rows > 0:
        try:
            out["ref_asis"] = cpu_asis_rate(w, asis_rows)
        except Exception as exc:  # noqa: BLE001
            out["ref_asis"] = dict(error=f"{type(exc).__name__}: {exc}")
    return out


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.ref_batch or w["ref_batch"]
    K = args.steps
    # bound the arm to a few minutes: one probe step, then as many of the K steps as fit the budget
    _, sec = cpu_port_rate(w, 1, 1, B)
    K_run = max(2, min(K, int(args.ref_seconds / max(sec, 1e-3))))
    rate, sec = cpu_port_rate(w, K_run, 0, B)
    cores = torch.get_num_threads()
    line = dict(impl="reference", metric="train samples/sec", value=rate, unit="samples/s", n_gpus=args.gpus,
                steps=K_run, warmup=max(1, args.warmup), ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="fp32", data="synthetic",
                config=workload_config(w, args.batch or w["batch"], max(1, args.gpus),
                                       arm=dict(optimizer="torch.optim.Adam", device="host CPU", rows_per_timed_step=B)),
                cpu_baseline=dict(value=rate, unit="samples/s", cores=cores, kind="port",
                                  sample=f"{K_run} train steps of {B} rows (vectorised torch port of the reference, "
                                         f"oracle/torch_port.py, {cores} threads; a stronger baseline than the reference as "
                                         f"shipped, whose CPython element loops take ~75 % of a step)"),
                e2e=dict(value=rate, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    if not args.no_asis:
        try:
            line["cpu_baseline"]["ref_asis"] = cpu_asis_rate(w, args.asis_rows, max_seconds=30.0)
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"]["ref_asis"] = dict(error=f"{type(exc).__name__}: {exc}")
    emit(line)


def workload_config(w, B, world, **extra):
    cfg = dict(workload=w["name"], state_size=w["S"], features=w["features"], enc_hidden=list(w["enc_hidden"]),
               decoders=w["n_decoders"], dec_hidden=list(w["dec_hidden"] or ()), dropout=w["dropout"],
               precision=w["precision"], batch_per_gpu=B, global_batch=B * world, missing_mode="row",
               missing_cells="30% MNAR" if w["mnar"] else "none", parallelism=f"dp{world}")
    cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------------------------
# GPU measurement of one workload
# ------------------------------------------------------------------------------------------------------------------
ENGINES = {0: "fp32-fma", 1: "tcgen05-3xtf32", 2: "tcgen05-3xtf32-tmem", 3: "tcgen05-bf16-layerwise", 4: "mma-bf16-tile"}


class Bench:
    def __init__(self, dev, dist, world, rank):
        self.dev, self.dist, self.world, self.rank = dev, dist, world, rank
        self.peaks, self.peak_kind = measured_peaks()
        self.fma_peak = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """CUDA events around `steps` calls, barrier + synchronize on both sides, max over ranks -> ms (total)"""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.dist is not None:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def train(self, w, B, K, W, precision=None, e2e=True, e2e_steps=None, clock=None, n_resident=None):
        """-> dict(value, ms_per_step, roofline, e2e?, engine, gpu_launches_per_step)"""
        from torch.nn import CrossEntropyLoss
        from multimodn_b200 import FusedAdam, MultiModNHistory
        dev, world = self.dev, self.world
        prec = precision or w["precision"]
        model = build_model(w, dev, prec)
        if world > 1:
            model.enable_data_parallel()
        opt = FusedAdam(model, lr=w["lr"])
        crit = CrossEntropyLoss()
        rt = model.runtime()
        lib = rt.lib.dll
        bytes_in = B * (sum(w["features"]) * 4 + w["n_decoders"] * 8)
        if n_resident is None:                               # distinct resident batches: together well above the 126 MB L2
            n_resident = int(min(8, max(2, -(-400e6 // bytes_in))))
        resident = [make_batch(w, B, 100 + 17 * self.rank + i, device=dev) for i in range(n_resident)]
        engine = ENGINES[int(lib.mmn_plan_engine(rt.plan))]

        def step_resident(i):
            xs, y = resident[i % n_resident]
            model.train_epoch([(xs, y)], opt, crit)

        for i in range(W):
            step_resident(i)
        if clock is not None:
            clock.__enter__()
        ms = self.timed(step_resident, K)
        value = B * world * K / (ms * 1e-3)

        # roofline: the step's kernels alone (no optimizer, no all-reduce), events on their stream
        seq = [(i, i) for i in range(len(w["features"]))]
        metrics = rt.new_metrics()

        def kernel_only(i):
            xs, y = resident[i % n_resident]
            mb, keep, n = rt.prepare_batch(xs, y, seq, "row", None)
            rt.train_step(mb, n, w["err_penalty"], 0.01 * w["state_change_penalty"], True, metrics)

        for i in range(3):
            kernel_only(i)
        l0 = int(lib.mmn_wide_launch_count())
        kms = self.timed(kernel_only, K) / K
        wide_launches = (int(lib.mmn_wide_launch_count()) - l0) // K
        if clock is not None:
            # keep the load on until nvidia-smi (100 ms period) has a dozen samples: the number of extra rounds is computed
            # from the max-over-ranks kernel time, i.e. it is the same on every rank (no rank-local loop conditions)
            extra = int(min(200, max(2, 1500.0 / max(kms * K, 1e-3))))
            for _ in range(extra):
                self.timed(kernel_only, K)
            clock.__exit__()
        macs, macs_strict = macs_per_row(w), macs_per_row(w, strict=True)
        alg_bytes = B * train_bytes_per_row(w)
        gbs = alg_bytes / (kms * 1e-3) / 1e9
        flops = 6.0 * macs * B
        tf = flops / (kms * 1e-3) / 1e12
        hbm = dict(achieved_gbs=gbs, peak_gbs=self.peaks["hbm_gbs"], frac=gbs / self.peaks["hbm_gbs"],
                   algorithmic_bytes_per_sample=alg_bytes / B)
        if prec == "bf16" and engine == "tcgen05-bf16-layerwise":
            peak = self.peaks.get("bf16_tflops_sustained", self.peaks["bf16_tflops"])
            tf_strict = 6.0 * macs_strict * B / (kms * 1e-3) / 1e12
            roofline = dict(bound="tensor", achieved=tf, peak=peak, unit="TFLOP/s", frac=tf / peak, traffic=None,
                            kernel=f"mmn_wide_gemm_kernel (+ element-wise companions), {wide_launches} launches per step",
                            kernel_ms=kms, peak_source=self.peak_kind + " (sustained: timed inside a long step)",
                            algorithmic_flops_per_sample=6.0 * macs,
                            strict=dict(achieved=tf_strict, frac=tf_strict / peak,
                                        note="without the first-layer dX GEMMs, which are never computed"),
                            hbm=hbm)
            launches = wide_launches + 2
        elif prec == "bf16":
            roofline = dict(bound="hbm", achieved=gbs, peak=self.peaks["hbm_gbs"], unit="GB/s", frac=gbs / self.peaks["hbm_gbs"],
                            traffic=TRAFFIC.get((w["name"], engine)), kernel=f"mmn_nb_step_kernel<train> ({engine})", kernel_ms=kms,
                            peak_source=self.peak_kind, algorithmic_bytes_per_sample=alg_bytes / B,
                            tensor=dict(achieved_tflops=tf, note="bf16 mma.sync tile kernel: HBM-bound by design (SURVEY.md 8d)"))
            launches = 4
        else:
            fma_peak = self.fma_peak or 148 * 128 * 2 * 1.965e9 / 1e12
            roofline = dict(bound="fp32-fma" if w["name"] != "c1_titanic" else "hbm",
                            achieved=tf if w["name"] != "c1_titanic" else gbs,
                            peak=fma_peak if w["name"] != "c1_titanic" else self.peaks["hbm_gbs"],
                            unit="TFLOP/s" if w["name"] != "c1_titanic" else "GB/s",
                            frac=(tf / fma_peak) if w["name"] != "c1_titanic" else gbs / self.peaks["hbm_gbs"],
                            traffic=TRAFFIC.get((w["name"], engine)), kernel=f"mmn_step_kernel<{engine}, train>", kernel_ms=kms,
                            peak_source=("measured (FFMA micro-benchmark, mmn_selftest_fma_peak)" if self.fma_peak else "nominal")
                            if w["name"] != "c1_titanic" else self.peak_kind,
                            hbm=hbm, fp32_fma=dict(achieved_tflops=tf, peak_tflops=fma_peak, frac=tf / fma_peak))
            launches = 3
        out = dict(workload=w["name"], metric="train samples/sec", value=value, unit="samples/s", ms_per_step=ms / K,
                   dtype=prec, engine=engine, batch_per_gpu=B, roofline=roofline, gpu_launches_per_step=launches)

        if e2e:
            # the public call on a loader of pinned HOST batches: every step copies its inputs H2D (one batch ahead, on a
            # side stream) and reads its loss metrics back D2H (log_interval = 1)
            host = [make_batch(w, B, 500 + 17 * self.rank + i, pin=True) for i in range(3)]
            hist = MultiModNHistory([str(d) for d in range(w["n_decoders"])])
            # enough steps that the first batch's un-overlapped H2D copy (one batch ahead is all the staging can hide) is
            # amortised: ~150 ms of steps, at least min(K, 12), at most 40
            Ke = e2e_steps or int(min(40, max(3, min(K, 12), 150.0 / max(ms / K, 1e-3))))
            logged = []

            def epoch_e2e(n):
                model.train_epoch([host[i % 3] for i in range(n)], opt, crit, hist, log_interval=1, logger=logged.append)

            epoch_e2e(3)
            ems = self.timed(lambda i: epoch_e2e(Ke) if i == 0 else None, 1)
            assert len(logged) == 3 + Ke
            # the ceiling the host side sets: the same pinned batches copied by every rank at once with nothing else running
            # (all ranks share the host's DRAM and PCIe root complexes); an e2e step cannot be shorter than this copy
            dst = [[torch.empty_like(t, device=dev) for t in xs] + [torch.empty_like(y, device=dev)] for xs, y in host[:1]][0]

            def copy_only(i):
                xs, y = host[i % 3]
                for d, t in zip(dst, list(xs) + [y]):
                    d.copy_(t, non_blocking=True)

            copy_only(0)
            cms = self.timed(copy_only, 6) / 6
            del dst
            out["e2e"] = dict(value=B * world * Ke / (ems * 1e-3), unit="samples/s", h2d_bytes_per_step=bytes_in,
                              d2h_bytes_per_step=rt.n_metrics * 8, steps=Ke, ms_per_step=ems / Ke,
                              h2d_gbs_per_rank=bytes_in / (ems / Ke * 1e-3) / 1e9,
                              h2d_ceiling=dict(gbs_per_rank=bytes_in / (cms * 1e-3) / 1e9, ms_per_step=cms,
                                               note="copy-only time of one step's inputs, all ranks copying at once"),
                              call="MultiModN.train_epoch(loader of pinned host batches, FusedAdam, CrossEntropyLoss, history, log_interval=1)")
            del host
        del resident, model, opt, rt
        torch.cuda.empty_cache()
        return out

    def predict_sweep(self, w, N, n_perm=16):
        """BASELINE.json configs[4]: per-step predictions over permuted encoding sequences, N rows (multimodn.py:422-458).
        8 cyclic + 8 random permutations, with the workload's missingness; value = rows/s through MultiModN.predict (device
        inputs, predictions read back to the host as the API returns them); kernel = forward launch alone."""
        dev = self.dev
        model = build_model(w, dev, "fp32")
        rt = model.runtime()
        E = len(w["features"])
        xs, _ = make_batch(w, N, 900, device=dev)
        g = torch.Generator().manual_seed(4)
        perms = [[(i + s) % E for i in range(E)] for s in range(n_perm // 2)]
        perms += [torch.randperm(E, generator=g).tolist() for _ in range(n_perm - len(perms))]
        for p in perms[:2]:                                         # warm-up (plan, workspace)
            model.predict([xs[e] for e in p], np.array(p))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for p in perms:
            # a permuted sequence visits encoder p[k] at step k, fed by data[p[k]]
            model.predict([xs[e] for e in p], np.array(p))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        preds = torch.zeros((E + 1, w["n_decoders"], N), dtype=torch.uint8, device=dev)

        def fwd_only(i):
            p = perms[i % len(perms)]
            mb, keep, n = rt.prepare_batch([xs[e] for e in p], None, list(enumerate(p)), "row", None)
            rt.forward(mb, n, predictions=preds)

        for i in range(2):
            fwd_only(i)
        kms = self.timed(fwd_only, len(perms)) / len(perms)
        macs = macs_per_row(w)
        fma_peak = self.fma_peak or 148 * 128 * 2 * 1.965e9 / 1e12
        tf = 2.0 * macs * N / (kms * 1e-3) / 1e12
        alg = N * (4 * sum(w["features"]) + (E + 1) * w["n_decoders"])
        gbs = alg / (kms * 1e-3) / 1e9
        engine = ENGINES[int(rt.lib.dll.mmn_plan_forward_engine(rt.plan))]
        out = dict(workload="c5_sweep", metric="predict rows/sec", value=N * len(perms) / dt, unit="rows/s",
                   ms_per_call=dt / len(perms) * 1e3, rows=N, permutations=len(perms), dtype="fp32", engine=engine,
                   kernel_rows_per_s=N / (kms * 1e-3),
                   note="value: MultiModN.predict incl. the D2H of the (E+1, D, N) class ids and the float64 conversion the API "
                        "returns; kernel_rows_per_s: the forward launch alone",
                   roofline=dict(bound="fp32-fma", achieved=tf, peak=fma_peak, unit="TFLOP/s", frac=tf / fma_peak,
                                 kernel=f"mmn_step_kernel<{engine}, forward>", kernel_ms=kms,
                                 hbm=dict(achieved_gbs=gbs, peak_gbs=self.peaks["hbm_gbs"], frac=gbs / self.peaks["hbm_gbs"],
                                          algorithmic_bytes_per_row=alg / N)))
        del xs, preds, model, rt
        torch.cuda.empty_cache()
        return out

    def dp_parity(self, w, rows_per_rank=1024):
        """first-step self-check under --gpus N: the N-rank all-reduced gradient == the 1-rank gradient on the concatenated
        batch (rank 0 recomputes it without data parallelism).  -> max over tensors of |diff| / max|expected|"""
        from torch.nn import CrossEntropyLoss
        dist, dev, world, rank = self.dist, self.dev, self.world, self.rank
        model = build_model(w, dev)
        model.enable_data_parallel()
        rt = model.runtime()
        xs, y = make_batch(w, rows_per_rank, 4000 + rank, device=dev)
        seq = [(i, i) for i in range(len(w["features"]))]
        rt.dropout_base_seed, rt.step_counter = 7, 0
        mb, keep, n = rt.prepare_batch(xs, y, seq, "row", model._dp)
        metrics = rt.new_metrics()
        rt.train_step(mb, n, w["err_penalty"], 0.01 * w["state_change_penalty"], True, metrics)
        model._allreduce_grads(rt, seq)
        torch.cuda.synchronize()
        got = rt.gflat[:rt.packed.n_params].clone()
        # gather every rank's shard on rank 0
        parts_x = [[torch.empty_like(x) for _ in range(world)] for x in xs]
        parts_y = [torch.empty_like(y) for _ in range(world)]
        for x, px in zip(xs, parts_x):
            dist.all_gather(px, x)
        dist.all_gather(parts_y, y)
        res = None
        if rank == 0:
            solo = build_model(w, dev)
            srt = solo.runtime()
            srt.flat.copy_(rt.flat)
            srt.dropout_base_seed, srt.step_counter = 7, 0
            fx, fy = [torch.cat(px) for px in parts_x], torch.cat(parts_y)
            mb2, keep2, n2 = srt.prepare_batch(fx, fy, seq, "row", None)
            srt.train_step(mb2, n2, w["err_penalty"], 0.01 * w["state_change_penalty"], True, srt.new_metrics())
            torch.cuda.synchronize()
            want = srt.gflat[:srt.packed.n_params]
            worst = 0.0
            for p, off, owner in srt.packed.slots:
                a, b = got[off:off + p.numel()].double(), want[off:off + p.numel()].double()
                scale = float(b.abs().max())
                if scale > 0:
                    worst = max(worst, float((a - b).abs().max()) / scale)
            cos = float(torch.dot(got.double(), want.double()) / (got.double().norm() * want.double().norm()))
            res = dict(rows_per_rank=rows_per_rank, max_rel_err_per_tensor=worst, cosine=cos,
                       tolerance=1e-4 if w["precision"] == "fp32" else 2e-2,
                       ok=bool(worst <= (1e-4 if w["precision"] == "fp32" else 2e-2)),
                       note="N-rank all-reduced gradient vs the 1-rank gradient of the concatenated batch, per parameter tensor")
        del model, rt
        torch.cuda.empty_cache()
        return res


# DRAM bytes of ONE launch of the step kernel (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum; profiles/)
TRAFFIC = {("c2_mimic", "fp32-fma"): 709.7e6 + 252.2e6}


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 at the
    first communicator), so fd 1 is pointed at stderr for the whole run and the line goes to a saved copy of the real one."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="rows per GPU per step (default: the workload's)")
    ap.add_argument("--ref-batch", type=int, default=None, help="rows per step of the CPU arm (default: the workload's batch)")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="budget of the --impl reference arm")
    ap.add_argument("--asis-rows", type=int, default=2048, help="rows per step of the unmodified-reference timer")
    ap.add_argument("--no-asis", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configurations")
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"], help="override the workload's precision")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.precision:
        w = dict(w, precision=args.precision)
    if args.impl == "reference":
        return run_reference(args, w)

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

    B, K, W = args.batch or w["batch"], args.steps, max(args.warmup, 3)
    bench = Bench(dev, dist, world, rank)
    try:
        bench.fma_peak = measure_fma_peak(dev)
    except Exception:  # noqa: BLE001
        bench.fma_peak = None

    clk = ClockSampler(local)
    head = bench.train(w, B, K, W, clock=clk)
    clocks = clk.summary()

    configs, parity = {}, None
    if world > 1:
        # multi-GPU: the data-parallel numerical contract in the driver's record, plus C2 as a second scaling point
        parity = bench.dp_parity(w)
        if not args.no_configs and w["name"] != "c2_mimic":
            try:
                c2 = WORKLOADS["c2_mimic"]
                configs["c2_mimic"] = bench.train(c2, c2["batch"], K, W, e2e=False)
                configs["c2_mimic"]["dp_parity"] = bench.dp_parity(c2, rows_per_rank=4096)
                configs["c2_mimic_bf16"] = bench.train(c2, c2["batch"], K, W, precision="bf16", e2e=False)
            except Exception as exc:  # noqa: BLE001
                configs["c2_mimic"] = dict(error=f"{type(exc).__name__}: {exc}")
    elif not args.no_configs:
        def guarded(name, fn):
            try:
                configs[name] = fn()
            except Exception as exc:  # noqa: BLE001   (the headline line must not depend on a secondary one)
                configs[name] = dict(error=f"{type(exc).__name__}: {exc}")
                torch.cuda.empty_cache()
        Ks = max(5, min(K, 10))
        for name in ("c1_titanic", "c2_mimic", "c3_mnar", "c4_wide"):
            if name == w["name"]:
                continue
            ww = WORKLOADS[name]
            guarded(name, lambda ww=ww: bench.train(ww, ww["batch"], Ks, 3, e2e=name in ("c2_mimic", "c3_mnar", "c4_wide"),
                                                    e2e_steps=12 if name == "c2_mimic" else 6))
        c2 = WORKLOADS["c2_mimic"]
        guarded("c2_mimic_bf16", lambda: bench.train(c2, c2["batch"], Ks, 3, precision="bf16", e2e=True, e2e_steps=12))
        c1 = WORKLOADS["c1_titanic"]
        guarded("c1_titanic_bf16", lambda: bench.train(c1, c1["batch"], Ks, 3, precision="bf16", e2e=False))
        guarded("c5_sweep", lambda: bench.predict_sweep(WORKLOADS["c3_mnar"], 1 << 20))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        asis_rows = 0 if args.no_asis else args.asis_rows
        cpu = cpu_baseline(w, args.cpu_seconds, asis_rows)
        if not args.no_configs:
            for name in ("c2_mimic", "c3_mnar"):
                if name in configs and "error" not in configs[name]:
                    try:
                        configs[name]["cpu_baseline"] = cpu_baseline(WORKLOADS[name], args.cpu_seconds, asis_rows)
                    except Exception as exc:  # noqa: BLE001
                        configs[name]["cpu_baseline"] = dict(error=f"{type(exc).__name__}: {exc}")

    if rank == 0:
        n_res = "resident batches cycled, together > 126 MB L2"
        line = dict(metric="train samples/sec", value=head["value"], unit="samples/s", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=head["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=head["dtype"], data="synthetic",
                    config=workload_config(w, B, world, arm=dict(optimizer="FusedAdam", engine=head["engine"], l2_policy=n_res)),
                    clocks=clocks, e2e=head.get("e2e"), gpu_launches=head["gpu_launches_per_step"] * K,
                    roofline=head["roofline"], cpu_baseline=cpu,
                    fp32_fma_peak_tflops=dict(measured=bench.fma_peak, nominal=148 * 128 * 2 * 1.965e9 / 1e12))
        if parity is not None:
            line["dp_parity"] = parity
        if configs:
            line["configs"] = configs
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
