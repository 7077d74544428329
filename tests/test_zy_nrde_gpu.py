"""K3n (Neural RDE, csrc/nrde.cu) on the GPU through the C ABI / host mirror against the fp64 oracle
(oracle/nrde.py, pinned to the reference's NeuralRDE by tests/test_oracle_reference_pinned.py).
Tolerance: BASELINE north-star 1e-5 relative in fp32 for states and gradients.

Hardware-proven since the round-1 driver run (12 XPASS on B200): plain asserts.  Every test body still runs in a
child pytest process (conftest.isolated_gpu) so that a fault or a hang cannot poison or stall the session."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, isolated_gpu, rel_err
from oracle import log_ncde as O
from oracle import nrde as ON
from oracle import paths

pytestmark = pytest.mark.gpu

TOL = 1e-5
NR = np.load(os.path.join(GOLDEN, "ref_nrde.npz"))


def _ts(L):
    return (np.float32(1.0 / L) * np.arange(L, dtype=np.float32)).astype(np.float32)


def _intervals(ts, s):
    return np.concatenate([ts[np.arange(0, len(ts), s)], np.array([1.0], np.float32)]).astype(np.float32)

CASES = [
    # B, L,  C, s, depth, H, width, n_vf, dt0, scale
    (2, 40, 3, 4, 2, 16, 8, 2, 0.05, 1.0),
    (2, 40, 3, 4, 1, 12, 8, 1, 0.05, 1.0),
    (1, 36, 5, 3, 2, 22, 12, 3, 0.031, 2.0),
    (4, 400, 6, 4, 2, 64, 64, 3, 0.0101, 1.0),   # EigenWorms NRDE config dims, Wm resident in shared memory
    (2, 60, 7, 6, 2, 128, 64, 4, 0.07, 1.5),     # Wm mostly streamed from L2
    (1, 24, 3, 4, 3, 32, 16, 2, 0.1, 1.0),       # depth-3 rows
]


@pytest.mark.parametrize("case", CASES)
@isolated_gpu
def test_nrde_solve_and_adjoint_vs_oracle(built_lib, case):
    from lncde.models.NeuralCDEs import _NRDESolve

    B, L, C, s, depth, H, width, n_vf, dt0, scale = case
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.normal(scale=0.3 / np.sqrt(L / 40), size=(B, L, C)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float64)
    ts = paths.make_ts(L)
    rows, sc, grid = O.stage_table(ts, paths.make_intervals(ts, s), dt0, logsig.shape[1])
    Dm = logsig.shape[2] - 1
    layers = ON.init_params(H, Dm, width, n_vf, scale=scale, seed=3, dtype=torch.float64)
    y0 = torch.from_numpy(rng.normal(size=(B, H)) * 0.5)
    g_ys = torch.from_numpy(rng.normal(size=(B, len(grid), H)))
    lay = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in layers]
    y0o = y0.clone().requires_grad_(True)
    ys_o = ON.solve(lay, torch.from_numpy(logsig), y0o, rows, sc, H)
    (ys_o * g_ys).sum().backward()
    flat = ON.flatten_params(layers).float().cuda().requires_grad_(True)
    y0d = y0.float().cuda().requires_grad_(True)
    ys = _NRDESolve.apply(flat, torch.from_numpy(logsig).float().cuda(), y0d, torch.from_numpy(rows).cuda(),
                          torch.from_numpy(sc).cuda(), (H, width, n_vf))
    (ys * g_ys.float().cuda()).sum().backward()
    assert rel_err(ys.detach().cpu().numpy(), ys_o.detach().numpy()) < TOL
    assert rel_err(y0d.grad.cpu().numpy(), y0o.grad.numpy()) < TOL
    gp, off = flat.grad.cpu().numpy(), 0
    for W, b in lay:
        for t in (W, b):
            n = t.numel()
            assert rel_err(gp[off : off + n], t.grad.reshape(-1).numpy()) < TOL, (case, off)
            off += n
    # deterministic: a second backward gives the same bits
    flat.grad = None
    ys2 = _NRDESolve.apply(flat, torch.from_numpy(logsig).float().cuda(), y0d, torch.from_numpy(rows).cuda(),
                           torch.from_numpy(sc).cuda(), (H, width, n_vf))
    (ys2 * g_ys.float().cuda()).sum().backward()
    assert torch.equal(ys2, ys) and np.array_equal(flat.grad.cpu().numpy(), gp)


@isolated_gpu
def test_nrde_train_step_eigenworms_config(built_lib):
    """One optimiser step of the reference's EigenWorms NRDE configuration (experiment_configs/repeats/nrde/
    EigenWorms.json: H 64, vf 3 x 64, depth 2, stepsize 4, no time channel) on a shortened path: the loss falls."""
    from lncde import train as T
    from lncde.data_dir.generate_paths import calc_paths
    from lncde.models.generate_model import create_model

    B, L, C, s = 8, 1200, 6, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(B, L, C, device="cuda", generator=g) * 0.05).cumsum(1)
    ls = calc_paths(x, s, 2)
    ts = torch.from_numpy(np.tile(paths.make_ts(L), (B, 1)))
    iv = paths.make_intervals(paths.make_ts(L), s)
    model, _ = create_model("nrde", C, ls.shape[-1], 2, iv, 5, 64, vf_depth=3, vf_width=64, dt0=1.0 / 301, scale=1.0, key=2345)
    y = torch.eye(5, device="cuda")[torch.arange(B, device="cuda") % 5]
    X = (ts, ls, x[:, 0, :])
    opt = T.Adam(model, lr=1e-3)
    l0, grads = T.classification_loss(model, X, y)
    assert torch.isfinite(grads).all()
    for _ in range(3):
        T.make_step(model, None, X, y, T.classification_loss, None, opt, None, None)
    l1, _ = T.classification_loss(model, X, y)
    assert float(l1) < float(l0)


@pytest.mark.parametrize("i", range(int(NR["n"])))
@isolated_gpu
def test_nrde_vs_reference_code(built_lib, i):
    """create_model("nrde") over the K3n kernels against the reference's own NeuralRDE: predictions, loss, gradients."""
    from lncde import train as T
    from lncde.models.generate_model import create_model

    B, L, C, s, depth, H, width, nvf, cls, ostep, ld = (int(v) for v in NR[f"cfg{i}"])
    dt0, scale = (float(v) for v in NR[f"hyper{i}"])
    ts = _ts(L)
    model, state = create_model("nrde", C, NR[f"logsig{i}"].shape[-1], depth, _intervals(ts, s), ld, H, vf_depth=nvf,
                                vf_width=width, classification=bool(cls), output_step=ostep, dt0=dt0, scale=scale, key=1)
    assert state is None and not model.lip2
    assert NR[f"params{i}"].size == model.params.numel()
    with torch.no_grad():
        model.params.copy_(torch.from_numpy(NR[f"params{i}"]).float())
    X = (torch.from_numpy(np.tile(ts, (B, 1))), torch.from_numpy(NR[f"logsig{i}"]).float().cuda(),
         torch.from_numpy(NR[f"x{i}"][:, 0, :].copy()).float().cuda())
    pred = model(X)
    assert rel_err(pred.detach().cpu().numpy(), NR[f"pred{i}"]) < TOL
    y = torch.from_numpy(NR[f"y{i}"]).float().cuda()
    loss, grads = (T.classification_loss if cls else T.regression_loss)(model, X, y)
    want = float(NR[f"loss{i}"])
    assert abs(float(loss) - want) < TOL * max(1.0, abs(want))
    assert rel_err(grads.cpu().numpy(), NR[f"grad{i}"]) < TOL
