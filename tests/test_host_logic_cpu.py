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


def test_bench_autotune_decision():
    """bench.py switches a kernel variant on only if its own parity check passed AND it was measurably faster."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ok, bad = {"ok": True}, {"ok": False}
    t = lambda **kw: {k: {"ms": v} for k, v in kw.items()}
    k3ok = {"ok": True, "ok_by_variant": {"1": True, "2": True}}
    k1ok = {"ok": True, "ok_by_variant": {"1": True, "2": True}}
    rep = {"k1": {"parity": k1ok, "timing": t(classic=0.33, tma=0.25, tma_persistent=0.26)},
           "k3": {"parity": k3ok, "timing": t(default=19.0, fused=18.9, fused2=18.8)}}
    assert bench.decide_tuning(rep) == {"k1_tma": 1, "k3_fused": 0}       # 1 % is inside the noise margin
    rep["k1"]["timing"] = t(classic=0.33, tma=0.25, tma_persistent=0.22)
    assert bench.decide_tuning(rep)["k1_tma"] == 2
    rep["k1"]["parity"] = {"ok": False, "ok_by_variant": {"1": False, "2": False}}
    rep["k3"]["timing"] = t(default=19.0, fused=15.0, fused2=14.0)
    assert bench.decide_tuning(rep) == {"k1_tma": 0, "k3_fused": 2}       # faster but wrong is never used; best variant wins
    rep["k3"]["parity"] = {"ok": False, "ok_by_variant": {"1": True, "2": False}}
    assert bench.decide_tuning(rep) == {"k1_tma": 0, "k3_fused": 1}       # level 2 failed its check: fall back to level 1
    assert bench.decide_tuning({"k1": {"error": "x"}, "k3": {"parity": k3ok}}) == {"k1_tma": 0, "k3_fused": 0}
    assert bench.decide_tuning({}) == {"k1_tma": 0, "k3_fused": 0}
    # K2 workloads: one variant per family, same rule
    de = {"k2de": {"parity": {"ok": True, "ok_by_variant": {"1": True, "2": True}}, "timing": t(default=10.9, v2=8.0, v2_tc=6.5)}}
    assert bench.decide_tuning(de)["k2_de_v2"] == 2
    de["k2de"]["parity"] = {"ok": False, "ok_by_variant": {"1": True, "2": False}}
    assert bench.decide_tuning(de)["k2_de_v2"] == 1
    de["k2de"]["parity"] = {"ok": False, "ok_by_variant": {"1": False, "2": False}}
    assert bench.decide_tuning(de)["k2_de_v2"] == 0
    assert bench.decide_tuning({"k2": {"parity": ok, "timing": t(default_B32=1.0, fused_B32=0.99)}})["k2_fused_d"] == 0
    # time-parallel scans: judged on the timing of the workload's own model
    p = {"k2p": {"parity": ok, "timing": t(default_bd=1.5, pscan_bd=1.1, default_d=0.9, pscan_d=0.95)}}
    bench.WL["model"] = "bd_linear_ncde"
    assert bench.decide_tuning(p)["k2_pscan"] == 1
    bench.WL["model"] = "diagonal_linear_ncde"
    assert bench.decide_tuning(p)["k2_pscan"] == 0
    p["k2p"]["parity"] = bad
    bench.WL["model"] = "bd_linear_ncde"
    assert bench.decide_tuning(p)["k2_pscan"] == 0


def test_autotune_runs_the_tcgen05_level_in_its_own_process(monkeypatch):
    """bench.autotune("k2de"): levels 1, 2 in one subprocess, level 3 (CUTLASS kernel never seen on hardware) in another;
    the reports are merged, and a level-3 process that dies costs only level 3."""
    import importlib.util
    import json
    import os
    import types

    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    calls = []
    r12 = {"k2de": {"parity": {"cases": 3, "ok": True, "ok_by_variant": {"1": True, "2": True}, "errors": {},
                               "max_rel_diff_grad": {"1": 1e-7, "2": 2e-6}},
                    "timing": {"shape": "s", "default": {"ms": 10.9}, "v2": {"ms": 8.0}, "v2_tc": {"ms": 6.5}}}}
    r3 = {"k2de": {"parity": {"cases": 3, "ok": True, "ok_by_variant": {"3": True}, "errors": {},
                              "max_rel_diff_grad": {"3": 1e-6}},
                   "timing": {"shape": "s", "default": {"ms": 11.0}, "v2_umma": {"ms": 5.0}}}}
    state = {"level3_dies": False}

    def fake_run(cmd, **kw):
        calls.append(cmd)
        lv = cmd[cmd.index("--levels") + 1]
        if lv == "3" and state["level3_dies"]:
            return types.SimpleNamespace(returncode=-11, stdout="", stderr="Segmentation fault")
        return types.SimpleNamespace(returncode=0, stdout=json.dumps(r3 if lv == "3" else r12) + "\n", stderr="")

    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    tuning, rep = bench.autotune(("k2de",))
    assert [c[c.index("--levels") + 1] for c in calls] == ["1,2", "3"] and all("--no-oracle" in c for c in calls)
    assert rep["k2de"]["parity"]["ok_by_variant"] == {"1": True, "2": True, "3": True}
    assert rep["k2de"]["timing"]["default"] == {"ms": 10.9} and rep["k2de"]["timing"]["v2_umma"] == {"ms": 5.0}
    assert tuning["k2_de_v2"] == 3
    state["level3_dies"] = True
    tuning, rep = bench.autotune(("k2de",))
    assert "level3_error" in rep["k2de"] and "error" not in rep["k2de"]
    assert tuning["k2_de_v2"] == 2


def test_tuning_api_roundtrip(built_lib):
    from lncde import _lib

    for key in ("k1_tma", "k3_fused", "k2_fused_d", "k2_de_v2", "k3_stream", "k2_pscan"):
        old = _lib.get_tuning(key)
        _lib.set_tuning(key, 2)
        assert _lib.get_tuning(key) == 2
        _lib.set_tuning(key, old)
        assert _lib.get_tuning(key) == old
    import pytest

    with pytest.raises(_lib.LncdeError):
        _lib.set_tuning("no_such_key", 1)


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


def test_bench_autotune_cache(monkeypatch, tmp_path):
    """One autotune per (GPU model, library build, families): later runs of a session (N = 2, 4, 8) reuse the decision;
    a decision taken after a failed variant check is never cached."""
    import importlib.util
    import os

    import torch

    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    calls = []
    good = ({"k1_tma": 1, "k3_fused": 0}, {"k1": {"parity": {"ok": True}}, "k3": {"parity": {"ok": True}}})
    monkeypatch.setattr(bench, "autotune", lambda fam: (calls.append(fam) or good))
    monkeypatch.setattr(torch.cuda, "get_device_name", lambda i=0: "TEST GPU")
    monkeypatch.setattr(bench, "TUNE_CACHE", str(tmp_path / "cache.json"))
    t1, r1 = bench.cached_autotune(("k1", "k3"))
    t2, r2 = bench.cached_autotune(("k1", "k3"))
    assert t1 == t2 == good[0] and "cached" not in r1 and r2["cached"] and len(calls) == 1
    bench.cached_autotune(("k1", "k3g"))                      # other families: own entry
    bench.cached_autotune(("k1", "k3"), refresh=True)         # --retune
    assert len(calls) == 3
    os.remove(bench.TUNE_CACHE)
    monkeypatch.setattr(bench, "autotune", lambda fam: ({"k1_tma": 0, "k3_fused": 0}, {"k1": {"error": "boom"}, "k3": {}}))
    bench.cached_autotune(("k1", "k3"))
    assert not os.path.exists(bench.TUNE_CACHE)
