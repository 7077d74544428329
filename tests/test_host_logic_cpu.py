"""Host-side logic that needs no GPU: solver step grid / stage table, synthetic datasets,
HallSet mirror API, workload table of bench.py."""
import numpy as np
import pytest

from oracle import log_ncde as O
from oracle import paths


@pytest.mark.parametrize("L,s,dt0,n_steps", [(100, 4, 0.01, 99), (17984, 12, 0.00066711140760507, 1499),
                                             (896, 16, 0.002, 500), (49920, 10, 0.0002002803925495694, 4994)])
def test_stage_table_matches_oracle_and_reference_step_counts(L, s, dt0, n_steps, built_lib):
    from lncde.models.solver_grid import stage_table, step_grid

    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    K = built_lib.lncde_logsig_num_windows(L, s)
    rows, scale, grid = stage_table(ts[0], ts[-1], iv, dt0, K)
    r2, s2, g2 = O.stage_table(ts, iv, dt0, K)
    assert np.array_equal(rows, r2) and np.array_equal(scale, s2) and np.array_equal(grid, g2)
    assert len(grid) - 1 == n_steps                       # 99 / 1499 / 500 / ~4994 steps
    assert grid.dtype == np.float32 and grid[-1] == ts[-1] and np.all(np.diff(grid) > 0)
    # quirk B2: the first stage reads the LAST logsig row and divides by -T
    assert rows[0, 0] == K - 1 and scale[0, 0] < 0
    assert rows.min() >= 0 and rows.max() <= K - 1
    assert np.array_equal(step_grid(ts[0], ts[-1], dt0), grid)


def test_step_grid_max_steps():
    from lncde.models.solver_grid import step_grid

    with pytest.raises(RuntimeError):
        step_grid(0.0, 1.0, 1e-4, max_steps=100)


def test_synthetic_datasets_shapes():
    from lncde.data_dir.datasets import SHAPES, make_intervals, make_ts, synthetic_paths

    x = synthetic_paths("toy", 4)
    assert x.shape == (4, 100, 6) and x.dtype == np.float32 and np.array_equal(x, np.round(x))  # integer walks
    x = synthetic_paths("EigenWorms", 2, L=240)
    assert x.shape == (2, 240, 7) and np.allclose(x[0, :, 0], np.arange(240) / 240)             # time channel 0
    x = synthetic_paths("SelfRegulationSCP1", 3, scale_to_unit=True)
    assert x.shape == (3, 896, 6) and x.min() >= -1 and x.max() <= 1
    assert make_intervals(make_ts(17984), 12).shape == (1500,)
    assert set(SHAPES) >= {"toy", "EigenWorms", "SelfRegulationSCP1", "ppg"}


def test_hallset_mirror_api(built_lib):
    from lncde.data_dir.hall_set import HallSet
    from oracle.hall import Hall

    hs = HallSet(7, 2)
    assert len(hs) == 29 and hs.data[0] == (0, 0) and hs.data[1] == (0, 1) and hs.data[8] == (1, 2)
    assert hs.basis_list()[:9] == [1, 2, 3, 4, 5, 6, 7, [1, 2], [1, 3]]
    assert hs.pairs().shape == (28, 2)
    rowptr, cols, vals = hs.t2l_csr()
    assert rowptr[-1] == len(cols) == len(vals) == 49
    assert HallSet(3, 3).basis_list() == Hall(3, 3).basis_list()
    # t2l_matrix of a lower degree through a higher-degree set, as generate_paths.py:36 may ask
    assert np.array_equal(HallSet(3, 3).t2l_matrix(2), Hall(3, 2).t2l_matrix(2))


def test_bench_workloads_cover_baseline_configs():
    import importlib.util
    import os

    from conftest import REPO

    spec = importlib.util.spec_from_file_location("bench", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = bench.WORKLOADS
    assert wl["eigenworms_logncde"]["L"] == 17984 and wl["eigenworms_logncde"]["stepsize"] == 12
    assert wl["eigenworms_logncde"]["B"] == 32 and wl["eigenworms_logncde"]["depth"] == 2
    for name in ("toy_logncde", "scp1_bd", "ppg_logncde_d2", "eigenworms_de", "eigenworms_d"):
        assert name in wl


def test_lncde_error_without_gpu_paths():
    import torch

    from lncde import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.LncdeError):
        _lib.require_cuda(torch.zeros(1))


def test_nrde_mirror_layout_and_errors(built_lib):
    """create_model("nrde"): same errors as generate_model.py:151-153, reference pytree names, parameter count of
    the C ABI (lncde_nrde_param_count) = [vf | mlp_linear] of the host layout; no CPU fallback."""
    import numpy as np
    import torch

    from lncde import _lib
    from lncde.models.generate_model import create_model

    with pytest.raises(ValueError, match="vf_width and vf_depth for a NRDE"):
        create_model("nrde", 6, 22, 2, None, 5, 64, key=0, device="cpu")
    iv = np.linspace(0, 1, 12).astype(np.float32)
    model, state = create_model("nrde", 6, 22, 2, iv, 5, 64, vf_depth=3, vf_width=64, scale=10.0, key=0, device="cpu")
    assert state is None and not model.lip2 and not model.stateful and not model.nondeterministic
    assert model.logsig_dim == 21 and len(model.vf.mlp.layers) == 3
    assert model.vf.mlp.layers[0].weight.shape == (64, 64) and model.mlp_linear.weight.shape == (64 * 21, 64)
    assert model.linear1.weight.shape == (64, 6) and model.linear2.weight.shape == (5, 64)
    assert built_lib.lncde_nrde_param_count(64, 21, 64, 3) == model.n_field
    assert model.params.numel() == model.n_field + 64 * 7 + 5 * 65
    # only the vector field is divided by `scale` (NeuralCDEs.py:209-219)
    assert float(model.vf.mlp.layers[0].weight.detach().abs().max()) <= 1 / 8 / 10 + 1e-6
    assert float(model.mlp_linear.weight.detach().abs().max()) > 1 / 8 / 2
    assert built_lib.lncde_nrde_param_count(64, 21, 30, 3) < 0          # width % 4 != 0: outside the compiled range
    assert built_lib.lncde_nrde_bwd_workspace_bytes(32, 64, 21, 64, 3, 4497) > 0
    if not torch.cuda.is_available():
        with pytest.raises(_lib.LncdeError):
            model((torch.zeros(2, 44), torch.zeros(2, 11, 22), torch.zeros(2, 6)))


def test_kernel_flags_are_per_call_and_thread_local(built_lib):
    """The C ABI has no tuning state (round 2): variants are named by the `flags` word of each call.  The binding's
    `kernel_flags` context only decides which bits the host mirrors pass from THIS thread."""
    import threading

    from lncde import _lib

    assert _lib.call_flags() == 0
    assert not hasattr(built_lib, "lncde_set_tuning_does_not_exist")
    for gone in ("lncde_set_tuning", "lncde_get_tuning"):
        with pytest.raises(AttributeError):
            getattr(built_lib, gone)
    seen = {}
    with _lib.kernel_flags("generic"):
        assert _lib.call_flags() == _lib.FLAG_GENERIC
        with _lib.kernel_flags("k3_v3", "k2_mma_sync"):
            assert _lib.call_flags() == _lib.FLAG_GENERIC | _lib.FLAG_K3_V3 | _lib.FLAG_K2_MMA_SYNC
        t = threading.Thread(target=lambda: seen.setdefault("other", _lib.call_flags()))
        t.start()
        t.join()
        assert _lib.call_flags() == _lib.FLAG_GENERIC
    assert _lib.call_flags() == 0 and seen["other"] == 0
    with pytest.raises(KeyError):
        _lib.kernel_flags("no_such_variant")
