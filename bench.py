#!/usr/bin/env python
"""Benchmark of the Log-ODE hot path (BASELINE.json metric: train samples/sec, EigenWorms
Log-NCDE, at 1/2/4/8 B200; logsig HBM GB/s vs peak).

One "step" = one pass of the hot path over one batch of synthetic EigenWorms-shaped input:
K1 log-signatures of the raw path batch -> Log-NCDE forward (Heun solve) -> CE loss + Lip(2)
-> exact discrete adjoint -> [all-reduce of the flat gradient, N > 1] -> Adam.

    python bench.py --gpus N --steps K --warmup W          (torchrun for N > 1)
    python bench.py --impl reference ...                    CPU port of the reference path

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "log-neural-cdes_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch


# Default = BASELINE.json configs[1]: EigenWorms Log-NCDE (experiment_configs/repeats/log_ncde/EigenWorms.json),
# time channel included.  The other entries are the remaining BASELINE configs, selectable with --workload
# (their numbers go to BASELINE.md; the driver only runs the default).
WORKLOADS = {
    "eigenworms_logncde": dict(name="EigenWorms", model="log_ncde", B=32, L=17984, C=7, stepsize=12, depth=2, H=128,
                               vf_width=32, vf_depth=2, dt0=0.00066711140760507, scale=1000.0, lambd=1e-3, lr=1e-3,
                               label_dim=5, classification=True,
                               desc="EigenWorms Log-NCDE depth-2 logsig: B=32/GPU, L=17984, C=7 (6+time), stepsize 12, "
                                    "H=128, vf 2x32, Heun dt0=1/1499 (1499 steps), CE+Lip2, Adam"),
    "eigenworms_logncde_c6": dict(name="EigenWorms", model="log_ncde", B=32, L=17984, C=6, stepsize=12, depth=2, H=128,
                                  vf_width=32, vf_depth=2, dt0=0.00066711140760507, scale=1000.0, lambd=1e-3, lr=1e-3,
                                  label_dim=5, classification=True,
                                  desc="EigenWorms Log-NCDE, 6 channels (no time channel), otherwise as the default"),
    # the reference's own comparison baseline on the same input (experiment_configs/repeats/nrde/EigenWorms.json)
    "eigenworms_nrde": dict(name="EigenWorms", model="nrde", B=32, L=17984, C=6, stepsize=4, depth=2, H=64, vf_width=64,
                            vf_depth=3, dt0=0.00022237046920169, scale=1.0, lambd=0.0, lr=1e-3, label_dim=5,
                            classification=True,
                            desc="EigenWorms NRDE (reference config): B=32, L=17984, C=6 (no time channel), stepsize 4, "
                                 "depth 2, H=64, vf 3x64, Heun dt0=1/4497 (4497 steps), CE, Adam"),
    "toy_logncde": dict(name="toy", model="log_ncde", B=32, L=100, C=6, stepsize=4, depth=2, H=64, vf_width=128,
                        vf_depth=3, dt0=0.01, scale=1.0, lambd=0.0, lr=3e-4, label_dim=2, classification=True,
                        desc="toy Log-NCDE depth 2, stepsize 4: B=32, L=100, C=6, H=64, vf 3x128, dt0=0.01 (99 steps)"),
    "ppg_logncde_d1": dict(name="ppg", model="log_ncde", B=32, L=49920, C=7, stepsize=10, depth=1, H=128, vf_width=64,
                           vf_depth=3, dt0=0.0002002803925495694, scale=1000.0, lambd=0.0, lr=1e-3, label_dim=1,
                           classification=False, output_step=128,
                           desc="PPG-DaLiA Log-NCDE depth 1 (reference config): B=32, L=49920, C=7, stepsize 10, H=128, "
                                "vf 3x64, 4994 steps, regression (390 outputs), MSE"),
    "ppg_logncde_d2": dict(name="ppg", model="log_ncde", B=32, L=49920, C=7, stepsize=10, depth=2, H=128, vf_width=64,
                           vf_depth=3, dt0=0.0002002803925495694, scale=1000.0, lambd=0.0, lr=1e-3, label_dim=1,
                           classification=False, output_step=128,
                           desc="PPG-DaLiA Log-NCDE depth 2 (BASELINE variant), otherwise as ppg_logncde_d1"),
    "scp1_bd": dict(name="SelfRegulationSCP1", model="bd_linear_ncde", B=32, L=896, C=6, stepsize=16, depth=2, H=64,
                    block_size=4, lambd=0.0, lr=1e-4, label_dim=2, classification=True,
                    desc="SelfRegulationSCP1 BD-LNCDE: B=32, L=896, C=6, stepsize 16, depth 2, H=64, block 4"),
    "eigenworms_bd": dict(name="EigenWorms", model="bd_linear_ncde", B=32, L=17984, C=7, stepsize=12, depth=2, H=128,
                          block_size=4, lambd=1e-3, lr=1e-3, label_dim=5, classification=True,
                          desc="EigenWorms BD-LNCDE: B=32, L=17984, C=7, stepsize 12, depth 2, H=128, block 4"),
    "eigenworms_d": dict(name="EigenWorms", model="diagonal_linear_ncde", B=32, L=17984, C=7, stepsize=12, depth=1,
                         H=128, block_size=1, lambd=1e-3, lr=1e-3, label_dim=5, classification=True,
                         desc="EigenWorms D-LNCDE: depth 1, H=128, block 1"),
    "eigenworms_de": dict(name="EigenWorms", model="dense_linear_ncde", B=32, L=17984, C=7, stepsize=12, depth=2,
                          H=128, block_size=128, lambd=1e-3, lr=1e-3, label_dim=5, classification=True,
                          desc="EigenWorms DE-LNCDE: depth 2, H=128, block 128 (dense)"),
}
WL = dict(WORKLOADS["eigenworms_logncde"])
N_ROT = 16  # distinct input batches rotated through (16 x 16.1 MB = 258 MB > 126 MB L2)
FFMA_PROBE_SHAPE = (148 * 8, 2048)  # CTAs x loop trips of the FP32 probe launch (39.7 GFLOP: ~0.5 ms)
K1_ROOFLINE_BATCH = 2048  # paths per launch of the K1 roofline leg (1.03 GB read + 0.36 GB written: > L2)
K1_TMA_DRAM_TRAFFIC_BYTES = 1.031656e9 + 0.330286e9  # per launch at that shape: dram__bytes_read.sum + dram__bytes_write.sum (profiles/r02_full_k1_tma.md)


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md).
    Started before the warm-up (nvidia-smi needs ~0.3 s to produce its first sample); only the
    samples that arrive between mark_start() and mark_stop() are summarised."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()

        def summarise(lines):
            sm, mx, pw, reasons = [], [], [], set()
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, pw, reasons

        inside = [x for x in self.lines if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e30) + 0.06]
        scope = "timed region"
        if len(inside) < 2:  # very short timed region: fall back to every sample taken under load
            inside, scope = self.lines, "warm-up + timed region"
        sm, mx, pw, reasons = summarise(inside)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm),
                "scope": scope}


def make_inputs(rank, n_rot):
    """Synthetic EigenWorms-shaped batches: ch 0 = t, ch 1..6 = random walk N(0, 0.05^2)."""
    from lncde.data_dir.datasets import synthetic_paths

    linear = WL["model"].endswith("linear_ncde")
    x = synthetic_paths(WL["name"], WL["B"] * n_rot, seed=2345 + rank, scale_to_unit=linear)
    x = x[:, : WL["L"]]
    if x.shape[2] != WL["C"]:  # e.g. EigenWorms without its time channel
        x = np.ascontiguousarray(x[:, :, x.shape[2] - WL["C"]:])
    x = x.reshape(n_rot, WL["B"], WL["L"], WL["C"])
    rng = np.random.default_rng(99 + rank)
    if WL["classification"]:
        lab = np.eye(WL["label_dim"], dtype=np.float32)[rng.integers(0, WL["label_dim"], size=(n_rot, WL["B"]))]
    else:
        n_out = len(np.arange(np.float32(WL["output_step"] / WL["L"]), 1.0, np.float32(WL["output_step"] / WL["L"]),
                              dtype=np.float32)) + 1
        lab = rng.uniform(-1, 1, size=(n_rot, WL["B"], n_out)).astype(np.float32)
    return x, lab


def run_ours(args):
    from lncde import _build, _lib, train as T
    from lncde.data_dir.datasets import make_intervals, make_ts
    from lncde.data_dir.generate_paths import calc_paths
    from lncde.models.generate_model import create_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # The build runs BEFORE the process group exists: rank 0 compiles (a no-op when the .so is current), the other ranks
    # wait on its stamp file - nobody spins inside NCCL meanwhile.  Which kernels run is decided by the library from the
    # shapes (defaults fixed from the round-2 hardware measurements; there is no tuning state and no autotune any more).
    if rank == 0:
        _build.build_library()
    else:
        t_wait = time.time()
        while not (os.path.exists(_build.LIB_PATH) and os.path.exists(_build.STAMP)):
            if time.time() - t_wait > 900:
                raise SystemExit("bench.py: rank 0 never finished building the library")
            time.sleep(0.2)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    _lib.lib()

    B, L, C = WL["B"], WL["L"], WL["C"]
    x_host, lab_host = make_inputs(rank, N_ROT)
    x_pin = torch.from_numpy(x_host).pin_memory()
    lab_pin = torch.from_numpy(lab_host).pin_memory()
    x_dev = x_pin.to(dev)
    lab_dev = lab_pin.to(dev)
    ts0 = make_ts(L)
    iv = make_intervals(ts0, WL["stepsize"])
    ts = torch.from_numpy(ts0)[None, :].expand(B, -1)  # host tensor: only ts[0], ts[-1] are read (t0, t1)
    from lncde.data_dir.generate_paths import logsig_dim

    D = logsig_dim(C, WL["depth"])
    model, _ = create_model(WL["model"], C, D, WL["depth"], iv, WL["label_dim"], WL["H"],
                            block_size=WL.get("block_size"), vf_depth=WL.get("vf_depth"), vf_width=WL.get("vf_width"),
                            classification=WL["classification"], output_step=WL.get("output_step", 1),
                            dt0=WL.get("dt0", 1), scale=WL.get("scale", 1.0), lambd=WL["lambd"], parallel_steps=128,
                            key=2345, device=dev)
    loss_fn = T.classification_loss if WL["classification"] else T.regression_loss
    if world > 1:
        dist.broadcast(model.params.data, src=0)
    opt = T.Adam(model, WL["lr"])
    K = _lib.lib().lncde_logsig_num_windows(L, WL["stepsize"])
    ls_buf = torch.empty((B, K, D), device=dev)
    stage_buf = torch.empty((B, L, C), device=dev)
    lab_buf = torch.empty(tuple(lab_dev.shape[1:]), device=dev)

    def eager_step(raw, lab):
        ls = calc_paths(raw, WL["stepsize"], WL["depth"], out=ls_buf)
        X = (ts, ls, raw[:, 0, :])
        _, _, _, value = T.make_step(model, None, X, lab, loss_fn, None, opt, None, None, world_size=world)
        return value

    # 1 GPU: the whole step (K1 .. Adam) replays as ONE CUDA graph launch.  N > 1: launched eagerly by default - with the
    # kernel-only step (K5) that is ~15 launches + one ncclAllReduce, no longer ~110; capturing the all-reduce inside the
    # graph hung an 8-GPU run in round 1 and stays opt-in (LNCDE_BENCH_GRAPH_NCCL=1) until it has been seen to work.
    use_graph = (not args.no_graph) and (world == 1 or os.environ.get("LNCDE_BENCH_GRAPH_NCCL") == "1")
    step = T.GraphedStep(eager_step, [x_dev[0], lab_dev[0]]) if use_graph else eager_step

    # workloads whose per-step working set is smaller than L2 get an explicit L2 flush between steps
    flush = None
    if N_ROT * B * L * C * 4 < 200e6 and args.workload != "eigenworms_logncde":
        flush = torch.empty(64 * 1024 * 1024, device=dev)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(x_dev[i % N_ROT], lab_dev[i % N_ROT])
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    sampler.mark_start()
    # LNCDE_BENCH_PROFILE_RANGE=1: cudaProfilerStart / Stop around the timed steps, so that
    # `ncu --profile-from-start off` lists exactly the kernels of the timed region (scripts/r2_ncu_default.sh)
    prof_range = os.environ.get("LNCDE_BENCH_PROFILE_RANGE") == "1"
    if prof_range:
        torch.cuda.profiler.start()
    if flush is None:
        e0.record()
        for i in range(args.steps):
            step(x_dev[i % N_ROT], lab_dev[i % N_ROT])
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if prof_range:
            torch.cuda.profiler.stop()
    else:  # time each step on its own, flushing L2 in between (not timed)
        ms = 0.0
        for i in range(args.steps):
            flush.zero_()
            e0.record()
            step(x_dev[i % N_ROT], lab_dev[i % N_ROT])
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
    sampler.mark_stop()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H loss, every step --------------------
    if use_graph:  # copy straight into the graph's static input buffers
        stage_buf, lab_buf = step.static_in

    def e2e_step(i):
        stage_buf.copy_(x_pin[i % N_ROT], non_blocking=True)
        lab_buf.copy_(lab_pin[i % N_ROT], non_blocking=True)
        return float(step(stage_buf, lab_buf).item())  # D2H read of the loss

    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    sync()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        last_loss = e2e_step(i)
    f1.record()
    sync()
    e2e_ms = f0.elapsed_time(f1)

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    out = None
    in_mb = N_ROT * x_pin[0].numel() * 4 / 1e6
    l2_note = (f"rotating {N_ROT} distinct input batches ({in_mb:.0f} MB)"
               + (" and a 1.5 GB adjoint workspace streamed every step: working set > 126 MB L2"
                  if args.workload == "eigenworms_logncde" else
                  "; L2 additionally flushed by a 256 MB write between timed steps" if flush is not None else ""))
    if rank == 0:
        pk, pk_kind = peaks()
        out = {
            "metric": "train samples/sec EigenWorms Log-NCDE" if args.workload == "eigenworms_logncde"
            else f"train samples/sec {args.workload}",
            "value": world * B * args.steps / (ms * 1e-3),
            "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(world),
            "impl_detail": {"launch": "one CUDA graph per step" if use_graph else "eager", "l2_this_run": l2_note,
                            "kernels": "library defaults (chosen by shape; include/lncde_b200.h)"},
            "e2e": {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": int(x_pin[0].numel() * 4 + lab_pin[0].numel() * 4),
                    "d2h_bytes_per_step": 4, "last_loss": last_loss},
            "gpu_launches": own_launches_per_step(WL["model"], WL) * args.steps,
            "clocks": clocks,
        }
        if WL["model"].endswith("linear_ncde"):
            # K2 workloads: the whole step against its COMPULSORY bytes (logsig rows in, states out; the
            # reverse sweep reads the rows, the states and their cotangents) - flows, partials and the bracket GEMM
            # operands are implementation traffic, not counted.  At B = 32 these steps are launch / chain-latency
            # bound, so the fraction is small by construction; the HBM-sized measurement is scripts/lncde_sweep.py.
            T_rows = K
            alg = world * B * (2 * 4 * T_rows * D + 3 * 4 * T_rows * WL["H"])
            ach = alg / (ms / args.steps * 1e-3) / 1e9
            out["roofline"] = {"kernel": f"K2 train step ({WL['model']}: flow build, scans, bracket gradient)", "bound": "hbm",
                               "achieved": ach, "peak": pk["hbm_gbs"] * world, "unit": "GB/s",
                               "frac": ach / (pk["hbm_gbs"] * world), "traffic": None,
                               "peak_source": f"{pk_kind} MEASURED_PEAKS.json hbm_gbs x {world} GPU(s)",
                               "algorithmic_bytes_per_step": alg}
        if args.workload == "eigenworms_logncde":
            if world > 1:
                args.no_cpu_baseline = True  # rank 0 at N = 1 only: with N ranks busy on the host the figure is meaningless
            out.update(extra_measurements(args, model, opt, x_dev, lab_dev, step, pk, pk_kind, ms / args.steps))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def own_launches_per_step(model_name, wl):
    """Kernels of liblncde_b200.so launched by one train step (counted from the host code of csrc/ and the kernel-only
    step of the host mirrors, for the kernels the library picks by default; the ncu launch lists under profiles/ show the
    same kernels).  Since round 2 a classification step launches nothing else of substance: the timed region of the default
    workload shows these kernels plus one 3.5 us torch copy of the strided x[:, 0, :] slice (profiles/r02b_launches_logncde.md)."""
    adam = 2                                     # adam_dev_kernel + the step-counter bump
    lip = 1 if wl.get("lambd", 0.0) else 0       # lip2_kernel
    k5 = 1 + 2 + 1 + lip                         # affine_fwd, read-out (per-sample + reduce), affine_bwd, Lip(2)
    if model_name == "nrde":                     # autograd glue around the kernels: K1, solve, sweep, 3 GEMM groups, reduce
        return 1 + 1 + 1 + 3 + 1 + adam
    if model_name == "log_ncde":
        if not wl.get("classification", True):   # regression read-out stays in torch
            return 1 + 1 + 1 + 3 + 1 + adam
        return 1 + k5 + 1 + (1 + 3 + 1) + adam   # K1, K5, solve, sweep + 3 gradient-GEMM groups + gradient reduce, Adam
    mean = 2                                     # time_mean_fwd / _bwd
    if model_name == "diagonal_linear_ncde":
        # dlncde_fwd (forms the time mean), dlncde_bwd (takes its cotangent) + its batch reduce (two levels beyond 32 samples)
        return 1 + k5 + 1 + 2 + (1 if wl.get("B", 32) > 32 else 0) + adam
    brackets = 2 * (1 + (wl.get("depth", 2) - 1))  # K2a forward and reverse: one launch per degree >= 2 + the transpose
    if model_name == "dense_linear_ncde":
        # two operand packs + tcgen05 flow build + cluster sweep | lam-only reverse cluster sweep, tcgen05 bracket gradient, split-K reduce
        return 1 + k5 + mean + brackets + (3 + 1) + 3 + adam
    # block-diagonal: flow build + time-parallel scans (products / boundary / apply | products / local / boundary / local) + GEMM, reduce
    return 1 + k5 + mean + brackets + (1 + 3) + (4 + 2) + adam


def extra_measurements(args, model, opt, x_dev, lab_dev, step, pk, pk_kind, ms_per_step):
    """Roofline of the HBM-bound kernel (K1 logsig at a batch larger than L2), per-kernel
    breakdown of the step, and the CPU baseline.  Rank 0 only."""
    from lncde import train as T
    from lncde.data_dir.datasets import synthetic_paths
    from lncde.data_dir.generate_paths import calc_paths

    res = {}
    dev = x_dev.device
    L, C, s = WL["L"], WL["C"], WL["stepsize"]
    # ---- K1 roofline leg: B = 2048 EigenWorms paths = 1.03 GB read + 0.36 GB written per launch
    from lncde import _lib
    from lncde.data_dir.generate_paths import logsig_dim

    Bk = K1_ROOFLINE_BATCH
    K, D = _lib.lib().lncde_logsig_num_windows(L, s), logsig_dim(C, 2)
    n_src = min(64, Bk)
    big = torch.from_numpy(synthetic_paths("EigenWorms", n_src, seed=7)[:, :L]).to(dev).repeat(Bk // n_src, 1, 1)
    big += 0.001 * torch.randn(Bk, 1, 1, device=dev)
    out = torch.empty((Bk, K, D), device=dev)
    for _ in range(3):
        calc_paths(big, s, 2, out=out)
    torch.cuda.synchronize()
    n_rep = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_rep):
        calc_paths(big, s, 2, out=out)
    e1.record()
    torch.cuda.synchronize()
    k1_ms = e0.elapsed_time(e1) / n_rep
    alg_bytes = Bk * (4 * L * C + 4 * K * D)  # 677 436 B per sample (EigenWorms, C = 7)
    ach = alg_bytes / (k1_ms * 1e-3) / 1e9
    res["roofline"] = {"kernel": "logsig_levy_tma_kernel<7,depth2> (K1, TMA bulk-copy staging)",
                       "bound": "hbm", "achieved": ach,
                       "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                       # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed
                       # ncu --set full capture (profiles/r02_full_k1_tma.md); None until that capture exists
                       "traffic": K1_TMA_DRAM_TRAFFIC_BYTES,
                       "peak_source": f"{pk_kind} MEASURED_PEAKS.json hbm_gbs (burst copy)",
                       "launch_ms": k1_ms, "algorithmic_bytes_per_launch": alg_bytes,
                       "shape": f"B={Bk} x EigenWorms (L={L}, C={C}, stepsize {s}, depth 2): inputs "
                                f"{Bk * L * C * 4 / 1e9:.2f} GB > L2"}
    del big, out
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return res  # N > 1: the other ranks are waiting - breakdown, FP32 probe and CPU baseline are N = 1 measurements
    # ---- per-kernel breakdown of the train step (CUDA events around each phase) -----------------
    res["step_breakdown_ms"] = breakdown(model, opt, x_dev, lab_dev)
    res["step_breakdown_ms"]["whole_step"] = ms_per_step
    # The kernel that dominates the STEP is the adjoint sweep (+ its deferred gradient GEMMs).  It is bound
    # by the latency of ~12 dependent phases per Heun stage, not by HBM or the tensor pipe:
    # its FP32 rate is quoted against the NOMINAL FFMA peak (148 SM x 128 lanes x 2 x 1.965 GHz) for scale.
    try:
        ffma_peak, ffma_src = measured_ffma_tflops(dev), "measured on this device by lncde_probe_ffma (nominal 74.4)"
    except Exception as e:  # the probe must never cost the bench its line
        ffma_peak, ffma_src = 74.4, f"nominal FP32 FFMA peak (probe failed: {type(e).__name__})"
    macs_stage = 98304                                  # restructured field, SURVEY.md 8d (C = 7)
    n_stage = 2 * int(model._table[1].shape[0])        # Heun: two stages per step (1 499 steps for EigenWorms)
    flop = 2.0 * (2 * macs_stage) * n_stage * WL["B"]   # transposed mat-vecs + weight-gradient outer products
    t_ms = res["step_breakdown_ms"]["backward_K3b+gemm"]
    ach = flop / (t_ms * 1e-3) / 1e12
    res["roofline_dominant"] = {"kernel": "logncde_bwd_saved_kernel<7,depth2> + outer_gemm_kernel (K3b)",
                                "bound": "latency (dependent phases), fp32 issue", "achieved": ach,
                                "peak": ffma_peak, "unit": "TFLOP/s", "frac": ach / ffma_peak,
                                "peak_source": ffma_src,
                                "share_of_step": t_ms / ms_per_step,
                                "us_per_stage": 1e3 * t_ms / n_stage,
                                "algorithmic_flop_per_step": flop}
    # ---- CPU baseline ------------------------------------------------------------------------------
    if not args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline()
    return res


def measured_ffma_tflops(dev, reps=5):
    """FP32 FFMA rate of this device at its current clocks (csrc/probe.cu: 8 independent chains per thread, 8 CTAs of
    256 threads per SM) - the denominator for the FP32-issue bound solve kernels; MEASURED_PEAKS.json has no FP32 row."""
    from lncde import _lib

    L = _lib.lib()
    blocks, iters = FFMA_PROBE_SHAPE
    out = torch.empty(blocks * 256, device=dev)
    for _ in range(2):
        _lib.check(L.lncde_probe_ffma(_lib.stream_ptr(), _lib.ptr(out), blocks, iters), "lncde_probe_ffma")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(L.lncde_probe_ffma(_lib.stream_ptr(), _lib.ptr(out), blocks, iters), "lncde_probe_ffma")
    e1.record()
    torch.cuda.synchronize()
    return L.lncde_probe_ffma_flops(blocks, iters) * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12


def breakdown(model, opt, x_dev, lab_dev):
    from lncde import train as T
    from lncde.data_dir.generate_paths import calc_paths

    ev = lambda: torch.cuda.Event(enable_timing=True)
    names = ["logsig_K1", "forward_solve_K3+readout", "backward_K3b+gemm", "adam_K4"]
    acc = {n: 0.0 for n in names}
    reps = 3  # timed repetitions, after one untimed pass (first eager pass after graph capture pays one-off costs)
    B, L = WL["B"], WL["L"]
    ts = torch.zeros(B, L)
    ts[:, -1] = float(np.float32(1.0 / L) * np.float32(L - 1))
    for i in range(-1, reps):
        raw, lab = x_dev[i % N_ROT], lab_dev[i % N_ROT]
        es = [ev() for _ in range(5)]
        es[0].record()
        ls = calc_paths(raw, WL["stepsize"], WL["depth"])
        es[1].record()
        pred = model((ts, ls, raw[:, 0, :]))
        loss = torch.mean(-torch.sum(lab * torch.log(pred + 1e-8), dim=1)) + T._lip2(model)
        es[2].record()
        (g,) = torch.autograd.grad(loss, model.params)
        es[3].record()
        opt.update(model, g)
        es[4].record()
        torch.cuda.synchronize()
        if i < 0:
            continue
        for n, a, b in zip(names, es[:-1], es[1:]):
            acc[n] += a.elapsed_time(b) / reps
    return acc


class CpuReferenceStep:
    """The reference's formulation of the WHOLE hot-path step on the host cores (oracle port, kind "port": the
    reference's JAX-CPU code cannot run offline): numpy calc_paths (signax-style Chen / log / t2l) of the raw batch,
    then oracle.train.LogNCDEStep (1 + C full-MLP JVPs per stage via torch.func, all n Heun steps, autograd adjoint,
    Adam) - the same batch shape, path length, step count and hyper-parameters as the GPU arm's step."""

    def __init__(self, n_rot=2):
        from lncde.data_dir.datasets import synthetic_paths
        from oracle import log_ncde as O
        from oracle import paths as OP
        from oracle.train import LogNCDEStep

        self.OP = OP
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        B, s, C = WL["B"], WL["stepsize"], WL["C"]
        x = synthetic_paths(WL["name"], B * n_rot, seed=2345)[:, : WL["L"]]
        if x.shape[2] != C:
            x = np.ascontiguousarray(x[:, :, x.shape[2] - C:])
        self.x = x.reshape(n_rot, B, WL["L"], C)
        ts = OP.make_ts(WL["L"])
        iv = OP.make_intervals(ts, s)
        self.rows, self.sc, _ = O.stage_table(ts, iv, WL["dt0"], len(iv) - 1)
        self.stepper = LogNCDEStep(WL["H"], C, WL["vf_width"], WL["vf_depth"], WL["label_dim"], WL["depth"], WL["scale"],
                                   WL["lambd"], WL["lr"])
        self.y = torch.eye(WL["label_dim"])[torch.arange(B) % WL["label_dim"]]
        self.t_logsig = self.t_train = 0.0

    def step(self, i, n_steps=None):
        """One full step on batch i (n_steps: truncate the solve to its first n Heun steps - warm-up / cross-check only)."""
        x = self.x[i % len(self.x)]
        t0 = time.perf_counter()
        ls = self.OP.calc_paths(x, WL["stepsize"], WL["depth"], np.float32)
        t1 = time.perf_counter()
        rows, sc = self.rows, self.sc
        if n_steps is not None:
            rows, sc = rows[:n_steps], sc[:n_steps]
        loss = self.stepper.step(torch.from_numpy(ls), torch.from_numpy(x[:, 0].copy()), self.y, rows, sc)
        t2 = time.perf_counter()
        self.t_logsig, self.t_train = t1 - t0, t2 - t1
        return t2 - t0, loss


def cpu_baseline(full_steps=1):
    """In-line CPU baseline of the default bench run (rank 0, N = 1 only): `full_steps` complete train steps of the full
    batch - MEASURED, no extrapolation (about 11 s each on the GPU box's 16 cores) after a short truncated warm-up."""
    ref = CpuReferenceStep()
    n_full = len(ref.rows)
    ref.step(0, n_steps=16)  # warm-up of torch's thread pool / allocator: 16 of the Heun steps, not timed
    times = []
    for i in range(full_steps):
        t, loss = ref.step(i)
        times.append(t)
    t_step = float(np.median(times))
    B = WL["B"]
    ls_bytes = B * (4 * WL["L"] * WL["C"] + 4 * (len(ref.rows)) * 29)
    return {"value": B / t_step, "unit": "samples/s", "cores": ref.cores, "kind": "port",
            "logsig": {"value": ls_bytes / ref.t_logsig / 1e9, "unit": "GB/s", "kind": "port",
                       "sample": f"numpy oracle calc_paths on the full B={B} x L={WL['L']} x C={WL['C']} batch: "
                                 f"{ref.t_logsig:.3f} s"},
            "sample": f"oracle port (numpy logsig + torch-CPU fp32 Log-NCDE step in the reference formulation, 1+C JVPs, "
                      f"autograd adjoint, Adam) on the FULL workload: B={B}, all {n_full} Heun steps; {full_steps} "
                      f"measured step(s) of {t_step:.2f} s (logsig {ref.t_logsig:.2f} s + train {ref.t_train:.2f} s), "
                      f"no extrapolation; last loss {loss:.4f}"}


def ref_budget_s():
    """Wall-clock budget of the reference arm: 25 full CPU steps take ~280 s on the GPU box's 16 cores; a slower host
    runs fewer timed steps (reported as such) instead of running long."""
    return float(os.environ.get("LNCDE_REF_BUDGET_S", "480"))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation cannot run offline (jax absent), so this times the
    oracle port (kind "port") on all host cores: `--warmup` untimed and `--steps` timed FULL train steps of the arm's
    workload (B = 32, every Heun step; nothing is scaled up).  Rank 0 only; the other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if WL["model"] != "log_ncde":
        raise SystemExit("--impl reference times the Log-NCDE workloads (the headline metric)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    budget = ref_budget_s()
    t_start = time.perf_counter()
    ref = CpuReferenceStep()
    n_full = len(ref.rows)
    ref.step(0, n_steps=16)  # torch's first pass (thread pool, functorch set-up), 16 Heun steps: not a measurement
    warm, steps, times = 0, 0, []
    for i in range(args.warmup):
        t, _ = ref.step(i)
        warm += 1
        if (time.perf_counter() - t_start) + t * (args.steps + args.warmup - warm) > budget:
            break  # a slow host: spend what is left of the budget on timed steps
    t0 = time.perf_counter()
    for i in range(args.steps):
        t, loss = ref.step(warm + i)
        times.append(t)
        steps += 1
        if (time.perf_counter() - t_start) + t > budget and steps >= 2:
            break
    total = time.perf_counter() - t0
    B = WL["B"]
    value = B * steps / total
    cb = {"value": value, "unit": "samples/s", "cores": ref.cores, "kind": "port",
          "sample": f"oracle port (numpy logsig + torch-CPU fp32 Log-NCDE step, reference formulation 1+C JVPs, autograd "
                    f"adjoint, Adam) on the FULL workload: B={B}, all {n_full} Heun steps per step; {warm} warm-up + "
                    f"{steps} timed steps, {total / steps:.2f} s each, measured (no extrapolation)"
                    + (f"; at N={world} rank 0 runs one rank's batch" if world > 1 else "")}
    out = {"impl": "reference", "metric": "train samples/sec EigenWorms Log-NCDE" if args.workload == "eigenworms_logncde"
           else f"train samples/sec {args.workload}", "value": value,
           "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warm,
           "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": args.scaling,
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": bench_config(world),
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t_start}
    if steps != args.steps or warm != args.warmup:
        out["requested"] = {"steps": args.steps, "warmup": args.warmup, "budget_s": budget}
    print(json.dumps(out))


def bench_config(world):
    """`config` of the JSON line - the WORKLOAD, identical for both arms (what differs between the arms - launch mode,
    kernel tuning - is reported under `impl_detail`)."""
    in_mb = N_ROT * WL["B"] * WL["L"] * WL["C"] * 4 / 1e6
    return {"workload": WL["desc"], "global_batch": world * WL["B"], "parallelism": f"dp{world}",
            "l2": f"inputs rotate through distinct batches ({N_ROT} x {in_mb / N_ROT:.1f} MB on the GPU arm) and the adjoint "
                  f"workspace (> 1 GB for the default workload) is streamed every step: working set > 126 MB L2; "
                  f"workloads smaller than L2 are timed step by step with a 256 MB L2 flush in between"}


def _watchdog(seconds):
    """A hung collective must not hang the caller: hard-exit after `seconds`."""
    def fire():
        sys.stderr.write(f"bench.py watchdog: no result after {seconds} s, aborting\n")
        sys.stderr.flush()
        os._exit(3)

    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="eigenworms_logncde", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the contract): the workload's batch per GPU; strong: that batch split over the GPUs")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of as one CUDA graph")
    ap.add_argument("--no-autotune", action="store_true", help="accepted and ignored (there is no autotune any more)")
    args = ap.parse_args()
    WL.clear()
    WL.update(WORKLOADS[args.workload])
    if args.scaling == "strong":  # fixed global batch, B / N samples per GPU (chain-latency bound: expect << N x)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if WL["B"] % world:
            raise SystemExit(f"--scaling strong needs the batch ({WL['B']}) to divide by the number of GPUs ({world})")
        WL["B"] //= world
        WL["desc"] += f" [strong scaling: {WL['B']} samples per GPU]"
    # the default run finishes in ~40 s; never let a stuck collective hold N GPUs for long
    _watchdog(float(os.environ.get("LNCDE_BENCH_TIMEOUT", "600")) + (ref_budget_s() if args.impl == "reference" else 0.0))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
