"""Kernel variants that exist in the library but are not the default yet.

``logsig_levy_tma_kernel`` (calc_paths(kernel="tma"), tuning "k1_tma") and the fused-phase Log-NCDE kernels
(tuning "k3_fused") and the fused diagonal LNCDE kernels (tuning "k2_fused_d") were written after the GPU budget of round 1 was spent: all are verified against the default kernels
(the first two bit-for-bit) on the host emulator (tests/test_emu_kernels_cpu.py) but have not run on a B200.
They are exercised here in a subprocess (scripts/variant_check.py) so that a fault cannot poison this process,
and the tests are non-strict xfails until a hardware run has been seen: XPASS = correct on the GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REASON = "variant not yet run on hardware (emulator-verified only); the default path does not use it"


def _check(which):
    r = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "variant_check.py"), "--no-timing", "--only", which],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])[which]
    assert "error" not in res, res
    assert res["parity"]["ok"], res
    return res


@pytest.mark.xfail(strict=False, reason=REASON)
def test_k1_tma_variant_parity(built_lib):
    _check("k1")


@pytest.mark.xfail(strict=False, reason=REASON)
def test_k3_fused_variant_parity(built_lib):
    _check("k3")


@pytest.mark.xfail(strict=False, reason=REASON)
def test_k2_fused_diagonal_variant_parity(built_lib):
    _check("k2")


@pytest.mark.xfail(strict=False, reason=REASON)
def test_k2_dense_v2_variant_parity(built_lib):
    _check("k2de")


@pytest.mark.xfail(strict=False, reason=REASON)
def test_k3_generic_stream_variant_parity(built_lib):
    _check("k3g")
