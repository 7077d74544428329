"""Kernel variants on the GPU, in a process of their own.  Prints ONE JSON line:

  k1: ``calc_paths(kernel="tma")`` (logsig_levy_tma_kernel) - parity against the fp64 oracle and the classic
      kernel (bit-identical), CUDA-event timings of both at the roofline shape (B=2048 x EigenWorms);
  k3: ``set_tuning("k3_fused", 1)`` (specialised Log-NCDE kernels with the 32 x 32 layer phases fused on single
      warps) - loss, gradient and solver states against the default kernels, CUDA-event timings of the
      EigenWorms train step (forward + adjoint) under both settings.

Used by tests/test_zz_variants_gpu.py and by the autotune leg of bench.py, which run it as a subprocess so that
a fault or hang in a variant that has not been on hardware yet cannot take the caller's CUDA context with it.

  k2: ``set_tuning("k2_fused_d", 1)`` (fused diagonal LNCDE kernels, dlncde.cu) - loss and gradients against flow
      build + scan, timings of the EigenWorms D-LNCDE train step at B = 32 and B = 2048.

  k2de: ``set_tuning("k2_de_v2", 1)`` (dense LNCDE: 8 x 8 flow build, lam-only reverse sweep, triple-product bracket
      gradient, linear_de.cuh) - loss and gradients against the default kernels, timings of the EigenWorms DE step.

  k3g: ``set_tuning("k3_stream", 1)`` (generic Log-NCDE kernels with streamed global-weight mat-vecs) - loss and
      gradients against the default generic kernels on PPG / toy dims, timings of both.

  k2p: ``set_tuning("k2_pscan", 1)`` (time-parallel D / BD scans, linear_pscan.cuh) - loss and gradients against the
      sequential default scans, timings of the EigenWorms BD and D train steps.

    python scripts/variant_check.py [--no-timing] [--no-parity] [--no-oracle] [--only k1|k2|k3|k2de|k3g|k2p] [--levels 1,2]

``--levels`` restricts the k2de legs to the given ``k2_de_v2`` levels, so that a caller can give the level that launches a
library kernel it has never seen on hardware (3: tcgen05 / CUTLASS) a process of its own.
"""
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "log-neural-cdes_b200"))

import numpy as np
import torch


def _persistent(fn):
    """Run fn() with the persistent bulk-copy K1 kernel (tuning k1_tma = 2)."""
    from lncde import _lib

    _lib.set_tuning("k1_tma", 2)
    try:
        out = fn()
        torch.cuda.synchronize()
        return out
    finally:
        _lib.set_tuning("k1_tma", 0)


def parity():
    from lncde.data_dir.generate_paths import calc_paths

    # --no-oracle (how bench.py's autotune calls this script): the variant is judged against the DEFAULT CUDA kernel
    # only (bit-identity); the fp64 oracle is test infrastructure and is consulted from tests/ alone
    use_oracle = "--no-oracle" not in sys.argv
    if use_oracle:
        from oracle import paths

    res = {"cases": 0, "bit_identical_to_classic": True, "persistent_bit_identical": True,
           "max_rel_err_vs_fp64_oracle": 0.0, "integer_kat_exact": True}
    rng = np.random.default_rng(3)
    # integer KATs (bit-exact), all widths / odd sizes / every tile phase
    for C in range(1, 9):
        for depth in (1, 2):
            x = np.cumsum(np.round(rng.normal(size=(5, 101, C))), axis=1).astype(np.float32)
            xd = torch.from_numpy(x).cuda()
            for s in (1, 4, 7, 12, 16, 100, 102, 500):
                if s * C > 640:
                    continue
                a = calc_paths(xd, s, depth, kernel="tma").cpu().numpy()
                b = calc_paths(xd, s, depth, kernel="classic").cpu().numpy()
                a2 = _persistent(lambda: calc_paths(xd, s, depth)).cpu().numpy()
                res["cases"] += 1
                if use_oracle:
                    want = paths.calc_paths(x, s, depth, np.float64)
                    res["integer_kat_exact"] &= bool(np.array_equal(a.astype(np.float64), want))
                res["bit_identical_to_classic"] &= bool(np.array_equal(a, b))
                res["persistent_bit_identical"] &= bool(np.array_equal(a2, b))
    # float paths at the BASELINE shapes (short batch)
    for C, s, L in [(7, 12, 17984), (6, 12, 17984), (6, 16, 896), (7, 10, 4992), (4, 10, 4992), (7, 16, 896)]:
        x = np.cumsum(rng.normal(scale=0.05, size=(3, L, C)), axis=1).astype(np.float32)
        xd = torch.from_numpy(x).cuda()
        a = calc_paths(xd, s, 2, kernel="tma").cpu().numpy()
        b = calc_paths(xd, s, 2, kernel="classic").cpu().numpy()
        a2 = _persistent(lambda: calc_paths(xd, s, 2)).cpu().numpy()
        res["cases"] += 1
        res["bit_identical_to_classic"] &= bool(np.array_equal(a, b))
        res["persistent_bit_identical"] &= bool(np.array_equal(a2, b))
        if use_oracle:
            want = paths.calc_paths(x, s, 2, np.float64)
            err = float(np.abs(a - want).max() / np.abs(want).max())
            res["max_rel_err_vs_fp64_oracle"] = max(res["max_rel_err_vs_fp64_oracle"], err)
    res["ok"] = bool(res["integer_kat_exact"] and res["bit_identical_to_classic"]
                     and res["max_rel_err_vs_fp64_oracle"] < 1e-5)
    res["ok_by_variant"] = {"1": res["ok"], "2": bool(res["ok"] and res["persistent_bit_identical"])}
    return res


def timing(B=2048, L=17984, C=7, s=12, reps=10):
    from lncde.data_dir.generate_paths import calc_paths

    x = (torch.randn(64, L, C, device="cuda") * 0.05).cumsum(1).repeat(B // 64, 1, 1)
    x += 0.001 * torch.randn(B, 1, 1, device="cuda")
    out = torch.empty((B, (L + 1 + s - 1) // s, 1 + C + C * (C - 1) // 2), device="cuda")
    byts = x.numel() * 4 + out.numel() * 4
    res = {"shape": f"B={B} x L={L} x C={C}, stepsize {s}, depth 2", "algorithmic_bytes": byts}
    from lncde import _lib

    for k, mode in (("classic", 0), ("tma", 1), ("tma_persistent", 2)):
        _lib.set_tuning("k1_tma", mode)
        for _ in range(3):
            calc_paths(x, s, 2, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            calc_paths(x, s, 2, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[k] = {"ms": ms, "GB/s": byts / ms / 1e6}
    _lib.set_tuning("k1_tma", 0)
    return res


def _eigenworms_step(B, n_windows, seed=0):
    """Model, inputs and a closure running loss + flat gradient of the EigenWorms Log-NCDE config on a path of
    n_windows * 12 - 1 samples (n_windows = 1499 is the full BASELINE length)."""
    from lncde import train as T
    from lncde.data_dir.datasets import make_intervals, make_ts
    from lncde.data_dir.generate_paths import calc_paths, logsig_dim
    from lncde.models.generate_model import create_model

    C, s, H = 7, 12, 128
    L = n_windows * s - 1
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(B, L, C, device="cuda", generator=g) * 0.05).cumsum(1)
    x[..., 0] = torch.arange(L, device="cuda") / L
    ts0 = make_ts(L)
    iv = make_intervals(ts0, s)
    model, _ = create_model("log_ncde", C, logsig_dim(C, 2), 2, iv, 5, H, vf_depth=2, vf_width=32, dt0=1.0 / n_windows,
                            scale=1000.0, lambd=1e-3, key=2345, device=torch.device("cuda"))
    ts = torch.from_numpy(ts0)[None, :].expand(B, -1)
    y = torch.eye(5, device="cuda")[torch.arange(B, device="cuda") % 5]
    ls = calc_paths(x, s, 2)

    def step():
        return T.classification_loss(model, (ts, ls, x[:, 0, :]), y)

    return model, step


def k3_parity():
    from lncde import _lib

    res = {"cases": 0, "max_rel_diff_grad": {"1": 0.0, "2": 0.0}, "max_rel_diff_loss": {"1": 0.0, "2": 0.0}}
    for B, nwin in ((3, 40), (32, 200)):
        model, step = _eigenworms_step(B, nwin, seed=B)
        _lib.set_tuning("k3_fused", 0)
        l0, g0 = step()
        for v in (1, 2):
            _lib.set_tuning("k3_fused", v)
            l1, g1 = step()
            torch.cuda.synchronize()
            dl = abs(float(l1) - float(l0)) / max(abs(float(l0)), 1e-30)
            dg = float((g1 - g0).abs().max() / g0.abs().max().clamp_min(1e-30))
            res["max_rel_diff_loss"][str(v)] = max(res["max_rel_diff_loss"][str(v)], dl)
            res["max_rel_diff_grad"][str(v)] = max(res["max_rel_diff_grad"][str(v)], dg)
        _lib.set_tuning("k3_fused", 0)
        res["cases"] += 1
    res["ok_by_variant"] = {v: bool(res["max_rel_diff_loss"][v] < 1e-6 and res["max_rel_diff_grad"][v] < 1e-5)
                            for v in ("1", "2")}
    res["ok"] = all(res["ok_by_variant"].values())
    return res


def k3_timing(reps=5):
    from lncde import _lib

    model, step = _eigenworms_step(32, 1499)
    res = {"shape": "EigenWorms Log-NCDE train step (loss + adjoint), B=32, 1499 Heun steps, eager launches; "
                    "fused = k3_fused 1 (32 x 32 phases on single warps), fused2 = k3_fused 2 (+ Q4 folded into Q5)"}
    for name, v in (("default", 0), ("fused", 1), ("fused2", 2)):
        _lib.set_tuning("k3_fused", v)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"ms": e0.elapsed_time(e1) / reps}
    _lib.set_tuning("k3_fused", 0)
    return res


def _bd_step(B, n_windows, seed=0, model_name="bd_linear_ncde", bs=4, depth=2):
    """EigenWorms BD-LNCDE config (H = 128, 4 x 4 blocks, depth 2) - or the diagonal model on the same rows: loss + flat
    gradient closure."""
    from lncde import train as T
    from lncde.data_dir.generate_paths import calc_paths, logsig_dim
    from lncde.models.generate_model import create_model

    C, s, H = 7, 12, 128
    L = n_windows * s - 1
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(B, L, C, device="cuda", generator=g) * 0.05).cumsum(1)
    x = 2 * (x - x.amin((0, 1))) / (x.amax((0, 1)) - x.amin((0, 1))) - 1
    model, _ = create_model(model_name, C, logsig_dim(C, depth), depth, None, 5, H, block_size=bs, lambd=1e-3, key=2345,
                            device=torch.device("cuda"))
    ts = torch.zeros(B, L)
    y = torch.eye(5, device="cuda")[torch.arange(B, device="cuda") % 5]
    ls = calc_paths(x, s, depth)

    def step():
        return T.classification_loss(model, (ts, ls, x[:, 0, :]), y)

    return model, step


_K2P_CASES = (("bd", "bd_linear_ncde", 4, 2), ("d", "diagonal_linear_ncde", 1, 1))


def k2p_parity():
    """k2_pscan (time-parallel D / BD scans) against the sequential default kernels: loss and flat gradient."""
    from lncde import _lib

    res = {"cases": 0, "max_rel_diff_grad": 0.0, "max_rel_diff_loss": 0.0}
    for _, name, bs, depth in _K2P_CASES:
        for B, nwin in ((3, 70), (32, 1499), (5, 131)):
            model, step = _bd_step(B, nwin, seed=B, model_name=name, bs=bs, depth=depth)
            _lib.set_tuning("k2_pscan", 0)
            l0, g0 = step()
            _lib.set_tuning("k2_pscan", 1)
            l1, g1 = step()
            _lib.set_tuning("k2_pscan", 0)
            torch.cuda.synchronize()
            res["cases"] += 1
            res["max_rel_diff_loss"] = max(res["max_rel_diff_loss"], abs(float(l1) - float(l0)) / max(abs(float(l0)), 1e-30))
            res["max_rel_diff_grad"] = max(res["max_rel_diff_grad"],
                                           float((g1 - g0).abs().max() / g0.abs().max().clamp_min(1e-30)))
            del model, step
    res["ok"] = bool(res["max_rel_diff_loss"] < 1e-5 and res["max_rel_diff_grad"] < 5e-5)
    return res


def k2p_timing(reps=10):
    from lncde import _lib

    res = {"shape": "EigenWorms BD-LNCDE (4 x 4 blocks, depth 2) and D-LNCDE (depth 1) train steps (loss + gradient), "
                    "B=32, eager launches: sequential warp-shuffle scans vs the two-level time-parallel scans"}
    for tag, name, bs, depth in _K2P_CASES:
        model, step = _bd_step(32, 1499, model_name=name, bs=bs, depth=depth)
        for vname, v in (("default", 0), ("pscan", 1)):
            _lib.set_tuning("k2_pscan", v)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                step()
            e1.record()
            torch.cuda.synchronize()
            res[f"{vname}_{tag}"] = {"ms": e0.elapsed_time(e1) / reps}
        del model, step
        torch.cuda.empty_cache()
    _lib.set_tuning("k2_pscan", 0)
    return res


def _dlncde_step(B, n_windows, seed=0):
    """EigenWorms D-LNCDE config (diagonal_linear_ncde, depth-1 rows, H = 128): loss + flat gradient closure."""
    from lncde import train as T
    from lncde.data_dir.generate_paths import calc_paths, logsig_dim
    from lncde.models.generate_model import create_model

    C, s, H = 7, 12, 128
    L = n_windows * s - 1
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(B, L, C, device="cuda", generator=g) * 0.05).cumsum(1)
    x = 2 * (x - x.amin((0, 1))) / (x.amax((0, 1)) - x.amin((0, 1))) - 1
    model, _ = create_model("diagonal_linear_ncde", C, logsig_dim(C, 1), 1, None, 5, H, block_size=1, lambd=1e-3, key=2345,
                            device=torch.device("cuda"))
    ts = torch.zeros(B, L)
    y = torch.eye(5, device="cuda")[torch.arange(B, device="cuda") % 5]
    ls = calc_paths(x, s, 1)

    def step():
        return T.classification_loss(model, (ts, ls, x[:, 0, :]), y)

    return model, step


def k2_parity():
    from lncde import _lib

    res = {"cases": 0, "max_rel_diff_grad": 0.0, "max_rel_diff_loss": 0.0}
    for B, nwin in ((3, 40), (32, 1499), (5, 130)):
        model, step = _dlncde_step(B, nwin, seed=B)
        _lib.set_tuning("k2_fused_d", 0)
        l0, g0 = step()
        _lib.set_tuning("k2_fused_d", 1)
        l1, g1 = step()
        _lib.set_tuning("k2_fused_d", 0)
        torch.cuda.synchronize()
        res["cases"] += 1
        res["max_rel_diff_loss"] = max(res["max_rel_diff_loss"], abs(float(l1) - float(l0)) / max(abs(float(l0)), 1e-30))
        res["max_rel_diff_grad"] = max(res["max_rel_diff_grad"],
                                       float((g1 - g0).abs().max() / g0.abs().max().clamp_min(1e-30)))
    res["ok"] = bool(res["max_rel_diff_loss"] < 1e-5 and res["max_rel_diff_grad"] < 5e-5)
    return res


def k2_timing(reps=10):
    from lncde import _lib

    res = {"shape": "EigenWorms D-LNCDE train step (loss + gradient), eager launches; B=32 and B=2048 (HBM-sized)"}
    for B in (32, 2048):
        model, step = _dlncde_step(B, 1499)
        for name, v in (("default", 0), ("fused", 1)):
            _lib.set_tuning("k2_fused_d", v)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                step()
            e1.record()
            torch.cuda.synchronize()
            res[f"{name}_B{B}"] = {"ms": e0.elapsed_time(e1) / reps}
        del model, step
        torch.cuda.empty_cache()
    _lib.set_tuning("k2_fused_d", 0)
    # algorithmic bytes of the fused pair per step at B = 2048: ys written once, ys + g_ys read once, logsig twice
    T, H, D = 1499, 128, 8
    res["fused_algorithmic_bytes_B2048"] = 2048 * T * (3 * H * 4 + 2 * D * 4)
    return res


def _de_step(B, n_windows, seed=0, depth=2):
    """EigenWorms DE-LNCDE config (dense_linear_ncde, depth 2 [3: the sweep's depth], H = 128, one 128 x 128 block):
    loss + flat gradient."""
    from lncde import train as T
    from lncde.data_dir.generate_paths import calc_paths, logsig_dim
    from lncde.models.generate_model import create_model

    C, s, H = 7, 12, 128
    L = n_windows * s - 1
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(B, L, C, device="cuda", generator=g) * 0.05).cumsum(1)
    x = 2 * (x - x.amin((0, 1))) / (x.amax((0, 1)) - x.amin((0, 1))) - 1
    model, _ = create_model("dense_linear_ncde", C, logsig_dim(C, depth), depth, None, 5, H, block_size=H, lambd=1e-3,
                            key=2345, device=torch.device("cuda"))
    ts = torch.zeros(B, L)
    y = torch.eye(5, device="cuda")[torch.arange(B, device="cuda") % 5]
    ls = calc_paths(x, s, depth)

    def step():
        return T.classification_loss(model, (ts, ls, x[:, 0, :]), y)

    return model, step


def _k2de_levels():
    if "--levels" in sys.argv:
        return tuple(v for v in sys.argv[sys.argv.index("--levels") + 1].split(",") if v in ("1", "2", "3"))
    return ("1", "2", "3")


def k2de_parity():
    from lncde import _lib

    levels = _k2de_levels()
    res = {"cases": 0, "max_rel_diff_grad": {v: 0.0 for v in levels}, "max_rel_diff_loss": {v: 0.0 for v in levels},
           "errors": {}}
    # the last case has depth-3 rows (n_basis = 140): several K tiles in the flow builds, and the reverse path that
    # materialises dF - at level 3 its bracket gradient is the second tcgen05 GEMM (dlie_umma)
    for B, nwin, depth in ((3, 40, 2), (8, 1499, 2), (5, 130, 2), (2, 60, 3)):
        model, step = _de_step(B, nwin, seed=B, depth=depth)
        _lib.set_tuning("k2_de_v2", 0)
        l0, g0 = step()
        for v in (int(x) for x in levels):
            if str(v) in res["errors"]:
                continue
            _lib.set_tuning("k2_de_v2", v)
            try:
                l1, g1 = step()
                torch.cuda.synchronize()
            except _lib.LncdeError as e:  # a status returned by the library before / at launch (e.g. level 3 built without CUTLASS)
                res["errors"][str(v)] = str(e)[:200]
                continue
            dl = abs(float(l1) - float(l0)) / max(abs(float(l0)), 1e-30)
            dg = float((g1 - g0).abs().max() / g0.abs().max().clamp_min(1e-30))
            res["max_rel_diff_loss"][str(v)] = max(res["max_rel_diff_loss"][str(v)], dl)
            res["max_rel_diff_grad"][str(v)] = max(res["max_rel_diff_grad"][str(v)], dg)
        _lib.set_tuning("k2_de_v2", 0)
        res["cases"] += 1
        del model, step
        torch.cuda.empty_cache()
    res["ok_by_variant"] = {v: bool(v not in res["errors"] and res["max_rel_diff_loss"][v] < 1e-5
                                    and res["max_rel_diff_grad"][v] < 5e-5) for v in levels}
    res["ok"] = all(res["ok_by_variant"].values())
    return res


def k2de_timing(reps=5):
    from lncde import _lib

    res = {"shape": "EigenWorms DE-LNCDE train step (loss + gradient), eager launches, B=32 (flows 3.1 GB); v2 = FFMA 8x8 "
                    "flow build + lam-only reverse sweep + triple-product gradient, v2_tc = the same with the flow build and "
                    "the bracket gradient on the tensor pipe (3xTF32), v2_umma = v2_tc with the flow build on tcgen05 (9xBF16)"}
    model, step = _de_step(32, 1499)
    for name, v in (("default", 0), ("v2", 1), ("v2_tc", 2), ("v2_umma", 3)):
        if v and str(v) not in _k2de_levels():
            continue
        _lib.set_tuning("k2_de_v2", v)
        try:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        except _lib.LncdeError as e:
            res[name] = {"error": str(e)[:200]}
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"ms": e0.elapsed_time(e1) / reps}
    _lib.set_tuning("k2_de_v2", 0)
    return res


def _generic_step(kind, B, n_windows, seed=0):
    """Log-NCDE configurations that run the GENERIC solve kernels with a layer in global memory: "ppg" (H 128,
    vf 3 x 64, C 7, depth 2, regression read-out replaced by classification for the check) and "toy" (H 64, vf 3 x 128,
    C 6, depth 2)."""
    from lncde import train as T
    from lncde.data_dir.datasets import make_intervals, make_ts
    from lncde.data_dir.generate_paths import calc_paths, logsig_dim
    from lncde.models.generate_model import create_model

    C, s, H, width, scale = (7, 10, 128, 64, 1000.0) if kind == "ppg" else (6, 4, 64, 128, 1.0)
    L = n_windows * s - 1
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(B, L, C, device="cuda", generator=g) * 0.05).cumsum(1)
    ts0 = make_ts(L)
    model, _ = create_model("log_ncde", C, logsig_dim(C, 2), 2, make_intervals(ts0, s), 5, H, vf_depth=3, vf_width=width,
                            dt0=1.0 / n_windows, scale=scale, lambd=1e-3, key=2345, device=torch.device("cuda"))
    ts = torch.from_numpy(ts0)[None, :].expand(B, -1)
    y = torch.eye(5, device="cuda")[torch.arange(B, device="cuda") % 5]
    ls = calc_paths(x, s, 2)

    def step():
        return T.classification_loss(model, (ts, ls, x[:, 0, :]), y)

    return model, step


def k3g_parity():
    from lncde import _lib

    res = {"cases": 0, "max_rel_diff_grad": 0.0, "max_rel_diff_loss": 0.0}
    for kind, B, nwin in (("ppg", 3, 40), ("toy", 4, 25), ("ppg", 32, 100)):
        model, step = _generic_step(kind, B, nwin, seed=B)
        _lib.set_tuning("k3_stream", 0)
        l0, g0 = step()
        _lib.set_tuning("k3_stream", 1)
        l1, g1 = step()
        _lib.set_tuning("k3_stream", 0)
        torch.cuda.synchronize()
        res["cases"] += 1
        res["max_rel_diff_loss"] = max(res["max_rel_diff_loss"], abs(float(l1) - float(l0)) / max(abs(float(l0)), 1e-30))
        res["max_rel_diff_grad"] = max(res["max_rel_diff_grad"],
                                       float((g1 - g0).abs().max() / g0.abs().max().clamp_min(1e-30)))
    res["ok"] = bool(res["max_rel_diff_loss"] < 1e-6 and res["max_rel_diff_grad"] < 1e-5)
    return res


def k3g_timing(reps=3):
    from lncde import _lib

    res = {"shape": "generic-kernel Log-NCDE train step (loss + adjoint), B=32, eager launches: PPG dims with 500 Heun "
                    "steps (a tenth of the PPG-DaLiA length) and the toy configuration (25 steps)"}
    for kind, nwin in (("ppg", 500), ("toy", 25)):
        model, step = _generic_step(kind, 32, nwin)
        for name, v in (("default", 0), ("stream", 1)):
            _lib.set_tuning("k3_stream", v)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                step()
            e1.record()
            torch.cuda.synchronize()
            res[f"{name}_{kind}"] = {"ms": e0.elapsed_time(e1) / reps}
    _lib.set_tuning("k3_stream", 0)
    return res


if __name__ == "__main__":
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    skip = sys.argv[sys.argv.index("--skip") + 1].split(",") if "--skip" in sys.argv else []
    out = {}
    for name, par, tim in (("k1", parity, timing), ("k3", k3_parity, k3_timing), ("k2", k2_parity, k2_timing),
                           ("k2de", k2de_parity, k2de_timing), ("k3g", k3g_parity, k3g_timing),
                           ("k2p", k2p_parity, k2p_timing)):
        if only not in (None, name) or name in skip:
            continue
        out[name] = {}
        try:
            if "--no-parity" not in sys.argv:
                out[name]["parity"] = par()
            if "--no-timing" not in sys.argv:
                out[name]["timing"] = tim()
        except Exception as e:  # report, keep going with the other variant
            out[name]["error"] = f"{type(e).__name__}: {e}"[:300]
    print(json.dumps(out))
