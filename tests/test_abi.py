"""The C-ABI library builds, loads and exports every symbol include/*.h declares.
No compute entry point is called here (no GPU)."""
import ctypes
import glob
import os
import re

from conftest import REPO


def _declared():
    names = set()
    for h in glob.glob(os.path.join(REPO, "include", "*.h")):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(lncde_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_header_declares_entry_points():
    names = _declared()
    for must in ("lncde_logsig_fwd", "lncde_hall_t2l_csr", "lncde_logncde_fwd", "lncde_logncde_bwd",
                 "lncde_linear_scan_fwd", "lncde_linear_scan_bwd", "lncde_adam_step"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    from lncde import _lib

    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(handle, name), f"{name} declared in include/ but not exported"
    # and the ctypes table covers the header exactly
    assert sorted(_lib.EXPORTS) == _declared()


def test_status_strings(built_lib):
    assert built_lib.lncde_strerror(0) == b"ok"
    assert b"NULL" in built_lib.lncde_strerror(-1)
    assert built_lib.lncde_version().startswith(b"lncde_b200")


def test_argument_errors_before_launch(built_lib):
    # argument validation happens on the host before any CUDA call
    assert built_lib.lncde_logsig_fwd(None, None, None, 1, 1, 1, 1, 1, 0, 1, None, None, None) == -1
    assert built_lib.lncde_logsig_num_windows(0, 1) == -2
    assert built_lib.lncde_logsig_num_windows(100, 4) == 26
    assert built_lib.lncde_logsig_num_windows(17984, 12) == 1499
    assert built_lib.lncde_hall_size(7, 2) == 29
    assert built_lib.lncde_hall_size(7, 3) == 141
    assert built_lib.lncde_hall_size(0, 2) == -3


def test_no_cpu_fallback():
    import pytest
    import torch

    from lncde import _lib
    from lncde.data_dir.generate_paths import calc_paths

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.LncdeError):
        calc_paths(torch.zeros(2, 10, 3), 4, 2)


def test_new_entry_points_validate_before_launch(built_lib):
    """K3n / K2d / K2 argument errors are reported by status on the host (no CUDA call is reached)."""
    import numpy as np

    L = built_lib
    buf = np.zeros(1 << 16, np.float32)
    p = ctypes.c_void_p(buf.ctypes.data + (-buf.ctypes.data) % 16)  # a 16-byte aligned non-NULL pointer
    mis = ctypes.c_void_p(p.value + 4)
    # lncde_nrde_fwd(stream, params, logsig, y0, rows, scale, ys, B, K, D, H, width, n_vf, n_steps)
    assert L.lncde_nrde_fwd(None, None, p, p, p, p, p, 1, 4, 4, 8, 8, 2, 3) == -1            # NULL
    assert L.lncde_nrde_fwd(None, p, p, p, p, p, p, 0, 4, 4, 8, 8, 2, 3) == -2               # B < 1
    assert L.lncde_nrde_fwd(None, p, p, p, p, p, p, 1, 4, 1, 8, 8, 2, 3) == -2               # D < 2
    assert L.lncde_nrde_fwd(None, p, p, p, p, p, p, 1, 4, 4, 8, 6, 2, 3) == -3               # width % 4 != 0
    assert L.lncde_nrde_fwd(None, p, p, p, p, p, p, 1, 4, 4, 8, 8, 5, 3) == -3               # n_vf > 4
    assert L.lncde_nrde_fwd(None, p, p, p, p, p, p, 1, 4, 4, 600, 8, 2, 3) == -3             # H > 512
    assert L.lncde_nrde_fwd(None, mis, p, p, p, p, p, 1, 4, 4, 8, 8, 2, 3) == -5             # params not 16-byte aligned
    assert L.lncde_nrde_param_count(64, 21, 64, 3) == 3 * 64 * 65 + 64 * 21 * 65
    assert L.lncde_nrde_param_count(64, 21, 64, 0) == -1
    need = L.lncde_nrde_bwd_workspace_bytes(2, 8, 3, 8, 2, 3)
    assert need > 0 and L.lncde_nrde_bwd_workspace_bytes(2, 8, 3, 6, 2, 3) == 0
    # lncde_nrde_bwd(stream, params, logsig, rows, scale, ys, g_ys, g_params, g_y0, B, K, D, H, width, n_vf, n_steps, ws, bytes)
    assert L.lncde_nrde_bwd(None, p, p, p, p, p, p, p, p, 2, 4, 4, 8, 8, 2, 3, p, need - 1) == -4   # workspace too small
    assert L.lncde_nrde_bwd(None, p, p, p, p, p, p, p, p, 2, 4, 4, 8, 8, 2, 3, None, need) == -1
    assert L.lncde_nrde_bwd(None, p, p, p, p, p, p, p, p, 2, 4, 4, 8, 8, 2, 3, mis, need) == -5
    # K2: shapes, workspace
    assert L.lncde_linear_scan_fwd(None, p, p, p, p, 1, 4, 4, 2, 4, 9, p, 1 << 20, 0) == -2    # n_basis > D - 1
    assert L.lncde_linear_scan_fwd(None, p, p, p, p, 1, 4, 4, 2, 4, 3, p, 16, 0) == -4         # workspace too small
    assert L.lncde_linear_scan_workspace_bytes(2, 5, 1, 128, 28) > L.lncde_linear_scan_workspace_bytes(2, 5, 32, 4, 28)
