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
