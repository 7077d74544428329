"""The kernels the library picks by default against its baseline kernels (LNCDE_FLAG_GENERIC), plus the opt-in variants
(LNCDE_FLAG_K2_MMA_SYNC, LNCDE_FLAG_K3_V3): loss and flat gradient of a train step through the host mirrors.

Exercised through scripts/variant_check.py in a subprocess per kernel family, so that a fault or a hang cannot poison
this process.  Plain asserts: every family has passed on hardware (GPUTEST_r01.json, gpurun calls of round 2)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_results = {}


def _run(which):
    """One subprocess per kernel family and session; the tests read their verdicts from its report."""
    if which not in _results:
        cmd = [sys.executable, os.path.join(REPO, "scripts", "variant_check.py"), "--no-timing", "--only", which]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            _results[which] = (r.returncode, r.stdout, r.stderr)
        except subprocess.TimeoutExpired:
            _results[which] = (-1, "", f"variant_check --only {which} killed after 300 s")
    rc, out, err = _results[which]
    assert rc == 0, err[-2000:]
    res = json.loads(out.strip().splitlines()[-1])[which]
    assert "error" not in res, res
    return res


def _check(which, variant=None):
    res = _run(which)
    if variant is None:
        assert res["parity"]["ok"], res
    else:
        assert res["parity"]["ok_by_variant"][variant], res
    return res


def test_k1_tma_and_classic_kernels_bit_identical(built_lib):
    """Both closed-form logsig kernels and the library's automatic choice: bit-identical, integer KATs exact, 1e-5 of fp64."""
    _check("k1")


def test_k3_six_barrier_kernels_parity(built_lib):
    _check("k3", "v3")


def test_k2_fused_diagonal_kernels_vs_flow_build_and_scan(built_lib):
    _check("k2", "fused")


@pytest.mark.parametrize("variant", ["mma_sync", "tcgen05"])  # 3xTF32 flow build: warp-level MMA / hand-written tcgen05
def test_k2_dense_tensor_core_kernels_vs_generic(built_lib, variant):
    _check("k2de", variant)


def test_k3_generic_streamed_kernels_vs_generic(built_lib):
    _check("k3g", "stream")


def test_k2_time_parallel_scans_vs_sequential(built_lib):
    _check("k2p", "pscan")


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
