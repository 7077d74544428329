"""K1 kernel variants on the GPU, in a process of their own: parity of ``kernel="tma"``
(logsig_levy_tma_kernel) against the oracle and the classic kernel, then CUDA-event timings of both
at the roofline shape.  Prints ONE JSON line.  Used by tests/test_zz_variants_gpu.py and by the
``k1_variants`` leg of bench.py, which run it as a subprocess so that a fault in a variant that has
not been on hardware yet cannot take the caller's CUDA context with it.

    python scripts/k1_variant_check.py [--no-timing] [--no-parity]
"""
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "log-neural-cdes_b200"))

import numpy as np
import torch


def parity():
    from lncde.data_dir.generate_paths import calc_paths
    from oracle import paths

    res = {"cases": 0, "bit_identical_to_classic": True, "max_rel_err_vs_fp64_oracle": 0.0, "integer_kat_exact": True}
    rng = np.random.default_rng(3)
    # integer KATs (bit-exact), all widths / odd sizes / every tile phase
    for C in range(1, 9):
        for depth in (1, 2):
            x = np.cumsum(np.round(rng.normal(size=(5, 101, C))), axis=1).astype(np.float32)
            xd = torch.from_numpy(x).cuda()
            for s in (1, 4, 7, 12, 16, 100, 102, 500):
                if s * C > 640:
                    continue
                want = paths.calc_paths(x, s, depth, np.float64)
                a = calc_paths(xd, s, depth, kernel="tma").cpu().numpy()
                b = calc_paths(xd, s, depth, kernel="classic").cpu().numpy()
                res["cases"] += 1
                res["integer_kat_exact"] &= bool(np.array_equal(a.astype(np.float64), want))
                res["bit_identical_to_classic"] &= bool(np.array_equal(a, b))
    # float paths at the BASELINE shapes (short batch)
    for C, s, L in [(7, 12, 17984), (6, 12, 17984), (6, 16, 896), (7, 10, 4992), (4, 10, 4992), (7, 16, 896)]:
        x = np.cumsum(rng.normal(scale=0.05, size=(3, L, C)), axis=1).astype(np.float32)
        xd = torch.from_numpy(x).cuda()
        want = paths.calc_paths(x, s, 2, np.float64)
        a = calc_paths(xd, s, 2, kernel="tma").cpu().numpy()
        b = calc_paths(xd, s, 2, kernel="classic").cpu().numpy()
        res["cases"] += 1
        res["bit_identical_to_classic"] &= bool(np.array_equal(a, b))
        err = float(np.abs(a - want).max() / np.abs(want).max())
        res["max_rel_err_vs_fp64_oracle"] = max(res["max_rel_err_vs_fp64_oracle"], err)
    res["ok"] = bool(res["integer_kat_exact"] and res["bit_identical_to_classic"]
                     and res["max_rel_err_vs_fp64_oracle"] < 1e-5)
    return res


def timing(B=2048, L=17984, C=7, s=12, reps=10):
    from lncde.data_dir.generate_paths import calc_paths

    x = (torch.randn(64, L, C, device="cuda") * 0.05).cumsum(1).repeat(B // 64, 1, 1)
    x += 0.001 * torch.randn(B, 1, 1, device="cuda")
    out = torch.empty((B, (L + 1 + s - 1) // s, 1 + C + C * (C - 1) // 2), device="cuda")
    byts = x.numel() * 4 + out.numel() * 4
    res = {"shape": f"B={B} x L={L} x C={C}, stepsize {s}, depth 2", "algorithmic_bytes": byts}
    for k in ("classic", "tma"):
        for _ in range(3):
            calc_paths(x, s, 2, out=out, kernel=k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            calc_paths(x, s, 2, out=out, kernel=k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[k] = {"ms": ms, "GB/s": byts / ms / 1e6}
    return res


if __name__ == "__main__":
    out = {}
    if "--no-parity" not in sys.argv:
        out["parity"] = parity()
    if "--no-timing" not in sys.argv:
        out["timing"] = timing()
    print(json.dumps(out))
