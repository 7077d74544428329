#!/usr/bin/env python
"""BASELINE.json configs[4]: DE-LNCDE (dense, block = H) against D-LNCDE (diagonal, block = 1) on depth-3
log-signature rows, swept over batch / length / width / channels.

    python scripts/lncde_sweep.py [--quick] [--steps K] [--warmup W] [--out FILE]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/lncde_sweep.py ...           (weak scaling: B samples per GPU, one all-reduce of the flat gradient)

One "step" = flows / scan forward -> CE loss + Lip(2) -> reverse sweep + bracket gradient -> [all-reduce] -> Adam
on log-signature rows drawn directly (level k ~ N(0, 0.1^2) / k!), device resident.  Timing: CUDA events per step,
L2 flushed by a 256 MB write between timed steps, max over ranks.  Rank 0 prints one JSON line per point and a
closing summary line; nothing here is a bench.py headline number.

Quoted against: D-LNCDE - the compulsory bytes of the fused form (rows read + states written
and read back, 8 T (1 + C) + 12 T H per sample and step) over the measured HBM rate; DE-LNCDE - the flow-build
contraction 2 T H^2 n_basis per sample (forward) over the nominal FP32 FFMA rate, and the flows' own bytes
(4 T H^2 per sample, written once and read by both sweeps) over the measured HBM rate."""
import argparse
import itertools
import json
import math
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.join(REPO, "log-neural-cdes_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch

FLOW_BYTES_CAP = 40e9  # dense flows F[B, T, H, H] above this are skipped (reported), not attempted
FFMA_NOMINAL_TFLOPS = 74.4


def rows_like_logsig(B, T, C, depth, dev, seed):
    """[B, T, D] rows with column 0 = 0 and level k ~ N(0, 0.1^2) / k!."""
    from lncde.data_dir.generate_paths import logsig_dim
    from lncde.data_dir.hall_set import HallSet

    D = logsig_dim(C, depth)
    g = torch.Generator(device=dev).manual_seed(seed)
    ls = torch.randn(B, T, D, device=dev, generator=g) * 0.1
    ls[..., 0] = 0
    degs = [1] * C
    if depth > 1:
        hs = HallSet(C, depth)
        deg = {0: 0}
        for k, (a, b) in enumerate(hs.data):
            deg[k] = 1 if a == 0 else deg[a] + deg[b]
        degs = [deg[k] for k in range(1, D)]
    scale = torch.tensor([1.0] + [1.0 / math.factorial(d) for d in degs], device=dev)
    return ls * scale, D


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def run_point(model_name, B, T, H, C, depth, steps, warmup, world, rank, dev, use_graph):
    from lncde import train as Tr
    from lncde.models.generate_model import create_model

    bs = 1 if model_name == "diagonal_linear_ncde" else H
    ls, D = rows_like_logsig(B, T, C, depth, dev, seed=100 + rank)
    x0 = torch.randn(B, C, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))
    y = torch.eye(5, device=dev)[torch.arange(B, device=dev) % 5]
    model, _ = create_model(model_name, C, D, depth, None, 5, H, block_size=bs, classification=True, lambd=1e-3,
                            parallel_steps=128, key=2345, device=dev)
    if world > 1:
        import torch.distributed as dist

        dist.broadcast(model.params.data, src=0)
    opt = Tr.Adam(model, 1e-3)

    def eager(ls_, x0_, y_):
        return Tr.make_step(model, None, (None, ls_, x0_), y_, Tr.classification_loss, None, opt, None, None,
                            world_size=world)[3]

    step = Tr.GraphedStep(eager, [ls, x0, y]) if use_graph else eager
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for _ in range(warmup):
        loss = step(ls, x0, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    for _ in range(steps):
        flush.zero_()
        if world > 1:
            dist.barrier()
        e0.record()
        loss = step(ls, x0, y)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    t = torch.tensor([ms / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), float(loss), D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="B in {32,128}, T in {128,1499}, H = 128, C = 7 only")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--grid", default=None, help='explicit points "B,T,H,C;B,T,H,C;..." instead of the sweep')
    ap.add_argument("--out", default=None)
    ap.add_argument("--kernels", default="default", choices=["default", "generic", "k2_mma_sync"],
                    help="default: the kernels the library picks by shape; generic / k2_mma_sync: lncde._lib.kernel_flags")
    args = ap.parse_args()

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("lncde_sweep.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from lncde import _build, _lib

    if rank == 0:
        _build.build_library()
    if world > 1:
        dist.barrier()
    _lib.lib()

    if args.grid:
        grid = [tuple(int(v) for v in pt.split(",")) for pt in args.grid.split(";") if pt]
    elif args.quick:
        grid = list(itertools.product((32, 128), (128, 1499), (128,), (7,)))
    else:
        grid = list(itertools.product((32, 128, 512), (128, 1024, 1499, 4096), (64, 128), (6, 7)))
    hbm, hbm_kind = peaks()
    use_graph = (not args.no_graph) and world == 1  # N > 1 launches eagerly, like bench.py
    results = []
    if args.kernels != "default":
        _lib.kernel_flags(args.kernels).__enter__()  # for the whole run (captured CUDA graphs keep the kernels they recorded)
    for (B, T, H, C), name in itertools.product(grid, ("diagonal_linear_ncde", "dense_linear_ncde")):
        rec = {"model": name, "B_per_gpu": B, "T": T, "H": H, "C": C, "depth": args.depth, "n_gpus": world}
        flow_bytes = 4.0 * B * T * H * H if name == "dense_linear_ncde" else 0.0
        if flow_bytes > FLOW_BYTES_CAP:
            rec["skipped"] = f"dense flows {flow_bytes / 1e9:.0f} GB > {FLOW_BYTES_CAP / 1e9:.0f} GB cap"
        else:
            try:
                ms, loss, D = run_point(name, B, T, H, C, args.depth, args.steps, args.warmup, world, rank, dev, use_graph)
                rec.update(ms_per_step=ms, samples_per_s=world * B / (ms * 1e-3), loss=loss, logsig_dim=D)
                if not math.isfinite(loss):
                    rec["error"] = "non-finite loss"
                if name == "diagonal_linear_ncde":
                    alg = B * (8.0 * T * (1 + C) + 12.0 * T * H)
                    rec["hbm"] = {"algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9, "peak_gbs": hbm,
                                  "frac": alg / (ms * 1e-3) / 1e9 / hbm, "peak_source": hbm_kind}
                else:
                    flop = 2.0 * B * T * H * H * (D - 1)
                    rec["flow_build"] = {"algorithmic_flop_fwd": flop, "tflops_if_whole_step": flop / (ms * 1e-3) / 1e12,
                                         "ffma_nominal_tflops": FFMA_NOMINAL_TFLOPS}
                    alg = 3.0 * flow_bytes
                    rec["hbm"] = {"algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9, "peak_gbs": hbm,
                                  "frac": alg / (ms * 1e-3) / 1e9 / hbm, "peak_source": hbm_kind,
                                  "note": "flows written once, read by the forward and the reverse sweep"}
            except Exception as e:  # one point must not lose the sweep (e.g. out of memory at the largest shapes)
                rec["error"] = f"{type(e).__name__}: {e}"[:300]
                torch.cuda.empty_cache()
        results.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
    if rank == 0:
        pairs = {}
        for r in results:
            if "ms_per_step" in r:
                pairs.setdefault((r["B_per_gpu"], r["T"], r["H"], r["C"]), {})[r["model"]] = r["ms_per_step"]
        ratio = {f"B{k[0]}_T{k[1]}_H{k[2]}_C{k[3]}": v["dense_linear_ncde"] / v["diagonal_linear_ncde"]
                 for k, v in pairs.items() if len(v) == 2}
        summary = {"summary": "DE-LNCDE / D-LNCDE step time ratio per point", "depth": args.depth, "n_gpus": world,
                   "kernels": args.kernels, "ratio": ratio}
        print(json.dumps(summary), flush=True)
        if args.out:
            with open(args.out, "w") as fh:
                json.dump({"points": results, **summary}, fh, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
