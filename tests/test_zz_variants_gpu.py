"""Kernel variants that exist in the library but are not the default yet (tuning keys of include/lncde_b200.h).

Every variant passed on the driver's B200 at the end of round 1 (GPUTEST_r01.json: 23 XPASS), so these are plain
asserts now.  They are exercised through scripts/variant_check.py in a subprocess per kernel family, so that a fault
or a hang cannot poison this process: parity against the DEFAULT CUDA kernels and, without --no-oracle, the fp64
oracle."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_results = {}


def _run(which, levels=None):
    """One subprocess per kernel family (and level group) and session; the levels read their verdicts from its report."""
    key = (which, levels)
    if key not in _results:
        cmd = [sys.executable, os.path.join(REPO, "scripts", "variant_check.py"), "--no-timing", "--only", which]
        if levels:
            cmd += ["--levels", levels]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            _results[key] = (r.returncode, r.stdout, r.stderr)
        except subprocess.TimeoutExpired:
            _results[key] = (-1, "", f"variant_check --only {which} killed after 240 s")
    rc, out, err = _results[key]
    assert rc == 0, err[-2000:]
    res = json.loads(out.strip().splitlines()[-1])[which]
    assert "error" not in res, res
    return res


def _check(which, level=None):
    # the tcgen05 level of the dense flow build runs in a process of its own (a fault there must not cost levels 1, 2)
    res = _run(which, None if which != "k2de" else ("3" if level == 3 else "1,2"))
    if level is None:
        assert res["parity"]["ok"], res
    else:
        assert res["parity"]["ok_by_variant"][str(level)], res
    return res


@pytest.mark.parametrize("level", [1, 2])  # 1: bulk-copy staging, 2: the same kernel persistent and double-buffered
def test_k1_tma_variant_parity(built_lib, level):
    _check("k1", level)


@pytest.mark.parametrize("level", [1, 2])  # fused-phase levels of the specialised Log-NCDE kernels
def test_k3_fused_variant_parity(built_lib, level):
    _check("k3", level)


def test_k2_fused_diagonal_variant_parity(built_lib):
    _check("k2")


@pytest.mark.parametrize("level", [1, 2, 3])  # 1: FFMA 8 x 8, 2: 3xTF32 mma.sync, 3: tcgen05 (CUTLASS FastFP32) flow build
def test_k2_dense_v2_variant_parity(built_lib, level):
    _check("k2de", level)


def test_k3_generic_stream_variant_parity(built_lib):
    _check("k3g")


def test_k2_time_parallel_scan_variant_parity(built_lib):
    _check("k2p")


def test_ffma_probe_runs(built_lib):
    """lncde_probe_ffma (csrc/probe.cu): launches, every chain converges to the fixed point 1 of x -> 0.999 x + 0.001."""
    import torch

    from lncde import _lib

    blocks, iters = 148, 4096
    out = torch.full((blocks * 256,), float("nan"), device="cuda")
    _lib.check(built_lib.lncde_probe_ffma(_lib.stream_ptr(), _lib.ptr(out), blocks, iters), "lncde_probe_ffma")
    torch.cuda.synchronize()
    assert built_lib.lncde_probe_ffma_flops(blocks, iters) == blocks * 256 * iters * 4 * 8 * 2
    assert torch.allclose(out, torch.full_like(out, 8.0), atol=1e-3)  # 8 chains per thread, each at ~1 after 16 384 rounds
    assert built_lib.lncde_probe_ffma(None, None, 1, 1) == -1 and built_lib.lncde_probe_ffma_flops(0, 1) < 0
