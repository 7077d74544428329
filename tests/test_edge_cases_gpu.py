"""Edge cases and error behaviour of the hot-path operators (GPU)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import paths

pytestmark = pytest.mark.gpu


def test_logsig_edge_shapes(built_lib):
    from lncde.data_dir.generate_paths import calc_paths

    rng = np.random.default_rng(0)
    for (B, L, C, s, d) in [(1, 1, 1, 1, 1), (1, 1, 3, 5, 2), (2, 2, 2, 1, 2), (3, 7, 8, 8, 2), (1, 33, 5, 33, 2),
                            (1, 33, 5, 34, 2), (5, 31, 4, 32, 2), (129, 9, 2, 4, 2), (2, 3, 2, 2, 3), (1, 5, 9, 3, 2)]:
        x = np.cumsum(np.round(rng.normal(size=(B, L, C))), axis=1).astype(np.float32)
        want = paths.batch_calc_paths(x, s, d, np.float64)
        from lncde.data_dir.generate_paths import batch_calc_paths

        got = batch_calc_paths(torch.from_numpy(x).cuda(), s, d).cpu().numpy()
        assert got.shape == want.shape, (B, L, C, s, d)
        assert rel_err(got, want) < 1e-6, (B, L, C, s, d)
    # constant path -> all-zero logsig except the jump from the zero basepoint in window 0
    x = np.ones((2, 20, 3), np.float32)
    got = calc_paths(torch.from_numpy(x).cuda(), 4, 2, compat=False).cpu().numpy()
    assert np.array_equal(got[:, 0, 1:4], np.ones((2, 3)))
    assert np.all(got[:, 1:] == 0) and np.all(got[:, 0, 4:] == 0)


def test_logsig_empty_batch_and_nan_containment(built_lib):
    """Empty input -> empty output of the right shape (no launch); a NaN sample poisons only the windows that
    contain it (windows are independent: no cross-window or cross-sample leakage)."""
    from lncde.data_dir.generate_paths import batch_calc_paths, calc_paths

    for d, D in ((1, 4), (2, 7), (3, 15)):
        got = calc_paths(torch.zeros(0, 10, 3, device="cuda"), 4, d)
        assert tuple(got.shape) == (0, 3, D) and got.device == torch.zeros(0, device="cuda").device
    assert tuple(batch_calc_paths(torch.zeros(0, 10, 3, device="cuda"), 4, 2).shape) == (0, 3, 7)
    x = torch.randn(3, 50, 4, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    clean = calc_paths(x, 5, 2)
    x[1, 22, 2] = float("nan")  # row 22 lies in window 4 only (zero row prepended: windows cover rows 5k-1 .. 5k+4)
    got = calc_paths(x, 5, 2)
    bad = torch.isnan(got).any(-1)
    assert bad[1, 4] and int(bad.sum()) == 1
    assert torch.equal(got[~bad], clean[~bad])


def test_logsig_argument_errors(built_lib):
    from lncde import _lib
    from lncde.data_dir.generate_paths import calc_paths

    x = torch.zeros(2, 10, 3, device="cuda")
    with pytest.raises(_lib.LncdeError):
        calc_paths(x, 4, 5)                      # depth outside [1, 4]
    with pytest.raises(_lib.LncdeError):
        calc_paths(x.double(), 4, 2)             # fp32 only
    with pytest.raises(_lib.LncdeError):
        calc_paths(x.cpu(), 4, 2)                # no CPU fallback
    with pytest.raises(_lib.LncdeError):
        calc_paths(x.transpose(1, 2), 4, 2)      # contiguous only
    with pytest.raises(_lib.LncdeError):
        calc_paths(x, 0, 2)                      # stepsize >= 1


def test_create_model_errors(built_lib):
    from lncde.models.generate_model import create_model

    iv = np.linspace(0, 1, 6).astype(np.float32)
    with pytest.raises(ValueError):
        create_model("log_ncde", 3, 7, 2, iv, 2, 8, key=0)                       # vf_width / vf_depth missing
    with pytest.raises(ValueError):
        create_model("no_such_model", 3, 7, 2, iv, 2, 8, key=0)
    with pytest.raises(ValueError):
        create_model("bd_linear_ncde", 3, 7, 2, iv, 2, 10, block_size=4, key=0)  # hidden % block != 0
    with pytest.raises(ValueError):
        create_model("log_ncde", 3, 16, 3, iv, 2, 8, vf_width=4, vf_depth=1, key=0)  # depth 3 (as the reference)
    with pytest.raises(NotImplementedError):
        create_model("ncde", 3, 7, 2, iv, 2, 8, vf_width=4, vf_depth=1, key=0)   # outside the hot path
    m, st = create_model("log_ncde", 3, 7, 2, iv, 2, 8, vf_width=4, vf_depth=1, key=0)
    assert st is None and m.lip2 and not m.stateful and not m.nondeterministic
    assert len(m.vf.mlp.layers) == 2 and tuple(m.vf.mlp.layers[-1].weight.shape) == (24, 4)
    assert tuple(m.pairs.shape) == (6, 2) and tuple(m.linear1.weight.shape) == (8, 3)


def test_solver_status_codes(built_lib):
    L = built_lib
    assert L.lncde_logncde_param_count(128, 7, 32, 2) == 36421 - (128 * 8 + 5 * 129)  # MLP part of the EW model
    assert L.lncde_logncde_param_count(128, 9, 32, 2) == -3   # C > 8 unsupported
    assert L.lncde_logncde_bwd_workspace_bytes(0, 128, 7, 32, 2, 2, 10) == 0
    assert L.lncde_linear_scan_workspace_bytes(2, 5, 4, 4, 6) > 0


def test_batch_of_one_and_inference_mode(built_lib):
    """B = 1 and the no-grad forward (no workspace) agree with the training-mode forward."""
    from lncde.models.generate_model import create_model

    L, C, s = 240, 7, 12
    rng = np.random.default_rng(1)
    x = np.cumsum(rng.normal(scale=0.2, size=(1, L, C)), axis=1).astype(np.float32)
    ls = torch.from_numpy(paths.calc_paths(x, s, 2, np.float32)).cuda()
    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    model, _ = create_model("log_ncde", C, 29, 2, iv, 5, 128, vf_width=32, vf_depth=2, dt0=0.05, scale=2.0, key=3)
    X = (torch.from_numpy(ts)[None], ls, torch.from_numpy(x[:, 0]).cuda())
    with torch.no_grad():
        a = model(X)
    b = model(X)
    assert torch.equal(a, b.detach())
    assert abs(float(a.sum()) - 1.0) < 1e-5


def test_adam_matches_optax_formula(built_lib):
    from lncde import train as T

    class M:
        pass

    m = M()
    torch.manual_seed(0)
    m.params = torch.randn(1000, device="cuda").requires_grad_(True)
    p0 = m.params.detach().clone().double().cpu()
    opt = T.Adam(m, lr=1e-2)
    mm, vv = torch.zeros(1000, dtype=torch.float64), torch.zeros(1000, dtype=torch.float64)
    p = p0.clone()
    for step in range(1, 4):
        g = torch.randn(1000, device="cuda")
        opt.update(m, g, grad_scale=0.5)
        gd = g.double().cpu() * 0.5
        mm = 0.9 * mm + 0.1 * gd
        vv = 0.999 * vv + 0.001 * gd * gd
        p = p - 1e-2 * (mm / (1 - 0.9**step)) / ((vv / (1 - 0.999**step)).sqrt() + 1e-8)
    assert rel_err(m.params.detach().cpu().numpy(), p.numpy()) < 1e-6
    assert int(opt.count.item()) == 3
