"""Kernel families on the GPU: the kernels the library picks by default against its BASELINE kernels
(LNCDE_FLAG_GENERIC) and the remaining opt-in variants, in a process of their own.  Prints ONE JSON line.

  k1    closed-form logsig: TMA bulk-copy kernel vs classic staging - bit-identical, fp64 oracle (without --no-oracle),
        CUDA-event timings at the roofline shape (B = 2048 x EigenWorms) and the library's automatic choice
  k3    Log-NCDE, EigenWorms model shape: default kernels vs the six-barrier kernels (LNCDE_FLAG_K3_V3)
  k2    diagonal LNCDE: fused dlncde kernels (default) vs flow build + scan (generic); B = 32 and B = 2048
  k2de  dense LNCDE: tcgen05 flow build (default) / mma.sync flow build (LNCDE_FLAG_K2_MMA_SYNC) vs generic FFMA kernels
  k3g   Log-NCDE, shapes outside the specialised kernels (PPG / toy): streamed mat-vecs (default) vs generic
  k2p   block-diagonal LNCDE: time-parallel scans (default) vs sequential warp scans (generic)

Every comparison is loss + flat gradient of a train step through the product's host mirrors; `lncde._lib.kernel_flags`
names the kernels of each run (the library keeps no tuning state).  Used by tests/test_zz_variants_gpu.py (a subprocess
per family, so a fault cannot poison the caller) and by hand for the timings quoted in BASELINE.md.

    python scripts/variant_check.py [--no-timing] [--no-parity] [--no-oracle] [--only k1|k3|k2|k2de|k3g|k2p]
"""
import contextlib
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "log-neural-cdes_b200"))

import numpy as np
import torch


def _flags(*names):
    from lncde import _lib

    return _lib.kernel_flags(*names) if names else contextlib.nullcontext()


def _timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _compare(step, variants, base=("generic",)):
    """Loss / gradient of `step` under each variant (tuple of kernel_flags names) against the `base` kernels."""
    with _flags(*base):
        l0, g0 = step()
    out = {}
    for name, names in variants.items():
        with _flags(*names):
            l1, g1 = step()
        torch.cuda.synchronize()
        out[name] = (abs(float(l1) - float(l0)) / max(abs(float(l0)), 1e-30),
                     float((g1 - g0).abs().max() / g0.abs().max().clamp_min(1e-30)))
    return out


def _family_parity(cases, make_step, variants, base=("generic",), tol_loss=1e-5, tol_grad=5e-5):
    res = {"cases": 0, "max_rel_diff_loss": {k: 0.0 for k in variants}, "max_rel_diff_grad": {k: 0.0 for k in variants}}
    for case in cases:
        model, step = make_step(*case)
        for name, (dl, dg) in _compare(step, variants, base).items():
            res["max_rel_diff_loss"][name] = max(res["max_rel_diff_loss"][name], dl)
            res["max_rel_diff_grad"][name] = max(res["max_rel_diff_grad"][name], dg)
        res["cases"] += 1
        del model, step
        torch.cuda.empty_cache()
    res["ok_by_variant"] = {k: bool(res["max_rel_diff_loss"][k] < tol_loss and res["max_rel_diff_grad"][k] < tol_grad)
                            for k in variants}
    res["ok"] = all(res["ok_by_variant"].values())
    return res


# ------------------------------------------------------------------------------------------------------ K1
def k1_parity():
    from lncde.data_dir.generate_paths import calc_paths

    use_oracle = "--no-oracle" not in sys.argv  # the fp64 oracle is test infrastructure: consulted from tests/ only
    if use_oracle:
        from oracle import paths
    res = {"cases": 0, "bit_identical_to_classic": True, "auto_bit_identical": True, "max_rel_err_vs_fp64_oracle": 0.0,
           "integer_kat_exact": True}
    rng = np.random.default_rng(3)
    for C in range(1, 9):  # integer KATs (bit-exact), all widths / odd sizes / every tile phase
        for depth in (1, 2):
            x = np.cumsum(np.round(rng.normal(size=(5, 101, C))), axis=1).astype(np.float32)
            xd = torch.from_numpy(x).cuda()
            for s in (1, 4, 7, 12, 16, 100, 102, 500):
                if s * C > 640:
                    continue
                a = calc_paths(xd, s, depth, kernel="tma").cpu().numpy()
                b = calc_paths(xd, s, depth, kernel="classic").cpu().numpy()
                res["cases"] += 1
                if use_oracle:
                    want = paths.calc_paths(x, s, depth, np.float64)
                    res["integer_kat_exact"] &= bool(np.array_equal(a.astype(np.float64), want))
                res["bit_identical_to_classic"] &= bool(np.array_equal(a, b))
    for C, s, L in [(7, 12, 17984), (6, 12, 17984), (6, 16, 896), (7, 10, 4992), (4, 10, 4992), (7, 16, 896)]:
        x = np.cumsum(rng.normal(scale=0.05, size=(3, L, C)), axis=1).astype(np.float32)
        xd = torch.from_numpy(x).cuda()
        a = calc_paths(xd, s, 2, kernel="tma").cpu().numpy()
        b = calc_paths(xd, s, 2, kernel="classic").cpu().numpy()
        res["cases"] += 1
        res["bit_identical_to_classic"] &= bool(np.array_equal(a, b))
        res["auto_bit_identical"] &= bool(np.array_equal(calc_paths(xd, s, 2).cpu().numpy(), b))
        if use_oracle:
            want = paths.calc_paths(x, s, 2, np.float64)
            res["max_rel_err_vs_fp64_oracle"] = max(res["max_rel_err_vs_fp64_oracle"],
                                                    float(np.abs(a - want).max() / np.abs(want).max()))
    res["ok"] = bool(res["integer_kat_exact"] and res["bit_identical_to_classic"] and res["auto_bit_identical"]
                     and res["max_rel_err_vs_fp64_oracle"] < 1e-5)
    res["ok_by_variant"] = {"tma": res["ok"]}
    return res


def k1_timing(B=2048, L=17984, C=7, s=12, reps=10):
    from lncde.data_dir.generate_paths import calc_paths

    x = (torch.randn(64, L, C, device="cuda") * 0.05).cumsum(1).repeat(B // 64, 1, 1)
    x += 0.001 * torch.randn(B, 1, 1, device="cuda")
    out = torch.empty((B, (L + 1 + s - 1) // s, 1 + C + C * (C - 1) // 2), device="cuda")
    byts = x.numel() * 4 + out.numel() * 4
    res = {"shape": f"B={B} x L={L} x C={C}, stepsize {s}, depth 2", "algorithmic_bytes": byts}
    for k, kern in (("classic", "classic"), ("tma", "tma"), ("auto", None)):
        ms = _timed(lambda: calc_paths(x, s, 2, out=out, kernel=kern), reps)
        res[k] = {"ms": ms, "GB/s": byts / ms / 1e6}
    return res


# ------------------------------------------------------------------------------------------------------ models
def _model_step(model_name, B, n_windows, seed=0, *, C=7, s=12, H=128, depth=2, bs=None, vf=None, dt0=None,
                scale=1000.0, unit_range=False):
    """Loss + flat gradient closure of one model on a synthetic path of n_windows * s - 1 samples."""
    from lncde import train as T
    from lncde.data_dir.datasets import make_intervals, make_ts
    from lncde.data_dir.generate_paths import calc_paths, logsig_dim
    from lncde.models.generate_model import create_model

    L = n_windows * s - 1
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(B, L, C, device="cuda", generator=g) * 0.05).cumsum(1)
    if unit_range:
        x = 2 * (x - x.amin((0, 1))) / (x.amax((0, 1)) - x.amin((0, 1))) - 1
    else:
        x[..., 0] = torch.arange(L, device="cuda") / L
    ts0 = make_ts(L)
    iv = make_intervals(ts0, s) if vf else None
    kw = dict(vf_depth=vf[0], vf_width=vf[1], dt0=dt0 or 1.0 / n_windows, scale=scale) if vf else dict(block_size=bs)
    model, _ = create_model(model_name, C, logsig_dim(C, depth), depth, iv, 5, H, lambd=1e-3, key=2345,
                            device=torch.device("cuda"), **kw)
    ts = torch.from_numpy(ts0)[None, :].expand(B, -1)
    y = torch.eye(5, device="cuda")[torch.arange(B, device="cuda") % 5]
    ls = calc_paths(x, s, depth)

    def step():
        return T.classification_loss(model, (ts, ls, x[:, 0, :]), y)

    step.model_forward = lambda: model((ts, ls, x[:, 0, :]))
    return model, step


def _eigenworms_step(B, n_windows, seed=0):
    """EigenWorms Log-NCDE config (n_windows = 1499 is the full BASELINE length); also used by scripts/phase_timing.py."""
    return _model_step("log_ncde", B, n_windows, seed, vf=(2, 32))


def _generic_step(kind, B, n_windows, seed=0):
    """Log-NCDE configurations outside the specialised kernels: "ppg" (H 128, vf 3 x 64, C 7) and "toy" (H 64, vf 3 x 128,
    C 6), depth 2."""
    if kind == "ppg":
        return _model_step("log_ncde", B, n_windows, seed, C=7, s=10, H=128, vf=(3, 64))
    return _model_step("log_ncde", B, n_windows, seed, C=6, s=4, H=64, vf=(3, 128), scale=1.0)


def _bd_step(B, n_windows, seed=0):
    return _model_step("bd_linear_ncde", B, n_windows, seed, bs=4, unit_range=True)


def _dlncde_step(B, n_windows, seed=0):
    return _model_step("diagonal_linear_ncde", B, n_windows, seed, depth=1, bs=1, unit_range=True)


def _de_step(B, n_windows, seed=0, depth=2):
    return _model_step("dense_linear_ncde", B, n_windows, seed, depth=depth, bs=128, unit_range=True)


def k3_parity():
    return _family_parity(((3, 40, 3), (32, 200, 32)), _eigenworms_step, {"v3": ("k3_v3",)}, base=(), tol_loss=1e-6,
                          tol_grad=1e-5)


def k3_timing(reps=5):
    model, step = _eigenworms_step(32, 1499)
    res = {"shape": "EigenWorms Log-NCDE train step (loss + adjoint) and inference forward, B=32, 1499 Heun steps, eager "
                    "launches; v3 = LNCDE_FLAG_K3_V3 (six-barrier stage kernels, logncde_v3.cuh)"}

    def infer():
        with torch.no_grad():
            step.model_forward()

    for name, names in (("default", ()), ("v3", ("k3_v3",))):
        with _flags(*names):
            res[name] = {"ms": _timed(step, reps), "inference_forward_ms": _timed(infer, reps)}
    return res


def k2_parity():
    return _family_parity(((3, 40, 3), (32, 1499, 32), (5, 130, 5)), _dlncde_step, {"fused": ()})


def k2_timing(reps=10):
    res = {"shape": "EigenWorms D-LNCDE train step (loss + gradient), depth-1 rows, H = 128, eager launches; generic = flow "
                    "build + scan, fused = dlncde kernels (no flows in HBM; the library default)"}
    for B in (32, 2048):
        model, step = _dlncde_step(B, 1499)
        for name, names in (("generic", ("generic",)), ("fused", ())):
            with _flags(*names):
                res[f"{name}_B{B}"] = {"ms": _timed(step, reps)}
        del model, step
        torch.cuda.empty_cache()
    T, H, D = 1499, 128, 8
    # the states written once and read once + the level-1 rows twice (the mean read-out needs no [B, T, H] cotangent)
    res["fused_algorithmic_bytes_B2048"] = 2048 * T * (2 * H * 4 + 2 * D * 4)
    return res


_DE_VARIANTS = {"mma_sync": ("k2_mma_sync",), "one_cta": ("k2_one_cta",), "tcgen05": ()}


def k2de_parity():
    # the last case has depth-3 rows (n_basis = 140): five K stages in the flow builds, dF-materialising reverse path
    return _family_parity(((3, 40, 3, 2), (8, 1499, 8, 2), (5, 130, 5, 2), (2, 60, 2, 3)), _de_step, _DE_VARIANTS)


def k2de_timing(reps=5):
    model, step = _de_step(32, 1499)
    res = {"shape": "EigenWorms DE-LNCDE train step (loss + gradient), eager launches, B=32 (flows 3.1 GB); generic = FFMA flow "
                    "build, dF materialised; mma_sync / tcgen05 = 3xTF32 flow build on the warp-level / 5th-generation tensor "
                    "cores + lam-only reverse sweep + triple-product bracket gradient"}
    for name, names in [("generic", ("generic",))] + list(_DE_VARIANTS.items()):
        with _flags(*names):
            res[name] = {"ms": _timed(step, reps)}
    return res


def k3g_parity():
    return _family_parity((("ppg", 3, 40, 3), ("toy", 4, 25, 4), ("ppg", 32, 100, 32)), _generic_step, {"stream": ()},
                          tol_loss=1e-6, tol_grad=1e-5)


def k3g_timing(reps=3):
    res = {"shape": "generic-kernel Log-NCDE train step (loss + adjoint), B=32, eager launches: PPG dims with 500 Heun "
                    "steps (a tenth of the PPG-DaLiA length) and the toy configuration (25 steps)"}
    for kind, nwin in (("ppg", 500), ("toy", 25)):
        model, step = _generic_step(kind, 32, nwin)
        for name, names in (("generic", ("generic",)), ("stream", ())):
            with _flags(*names):
                res[f"{name}_{kind}"] = {"ms": _timed(step, reps)}
    return res


def k2p_parity():
    return _family_parity(((3, 70, 3), (32, 1499, 32), (5, 131, 5)), _bd_step, {"pscan": ()})


def k2p_timing(reps=10):
    model, step = _bd_step(32, 1499)
    res = {"shape": "EigenWorms BD-LNCDE train step (loss + gradient), H = 128, 4 x 4 blocks, B=32, eager launches; generic = "
                    "sequential warp scans, pscan = scans parallel in time (the library default)"}
    for name, names in (("generic", ("generic",)), ("pscan", ())):
        with _flags(*names):
            res[name] = {"ms": _timed(step, reps)}
    return res


if __name__ == "__main__":
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    skip = sys.argv[sys.argv.index("--skip") + 1].split(",") if "--skip" in sys.argv else []
    out = {}
    for name, par, tim in (("k1", k1_parity, k1_timing), ("k3", k3_parity, k3_timing), ("k2", k2_parity, k2_timing),
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
        except Exception as e:  # report, keep going with the other families
            out[name]["error"] = f"{type(e).__name__}: {e}"[:300]
    print(json.dumps(out))
