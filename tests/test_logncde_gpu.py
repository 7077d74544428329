"""K3/K3b parity: CUDA Log-NCDE solve + discrete adjoint (through the C ABI) vs the
torch-CPU fp64 oracle (reference formulation: 1 + C full MLP JVPs, autograd)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import log_ncde as O
from oracle import paths

pytestmark = pytest.mark.gpu

TOL = 1e-5  # BASELINE.json north_star: 1e-5 relative in fp32 (norm-relative per tensor, fp64 arbiter)


def _setup(B, L, C, s, depth, H, width, n_hidden, dt0, scale, seed=0, compat=True):
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float64, compat=compat)
    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    rows, sc, grid = O.stage_table(ts, iv, dt0, logsig.shape[1])
    layers = O.init_params(H, C, width, n_hidden, scale=scale, seed=seed, dtype=torch.float64)
    y0 = torch.from_numpy(rng.normal(size=(B, H)) * 0.5)
    return x, logsig, rows, sc, grid, layers, y0


def _run_gpu(layers, logsig, y0, rows, sc, dims, g_ys):
    from lncde.models.LogNeuralCDEs import _LogNCDESolve

    flat = O.flatten_params(layers).float().cuda().requires_grad_(True)
    y0g = y0.float().cuda().requires_grad_(True)
    ys = _LogNCDESolve.apply(flat, torch.from_numpy(logsig).float().cuda().contiguous(), y0g,
                             torch.from_numpy(rows).cuda(), torch.from_numpy(sc).cuda(), dims)
    (ys * g_ys.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    return ys.detach().cpu().numpy(), flat.grad.cpu().numpy(), y0g.grad.cpu().numpy()


def _run_oracle(layers, logsig, y0, rows, sc, dims, g_ys):
    H, C, width, n_hidden, depth = dims
    lay = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in layers]
    y0o = y0.clone().requires_grad_(True)
    ys = O.solve(lay, torch.from_numpy(logsig), y0o, rows, sc, C, H, depth)
    (ys * g_ys).sum().backward()
    gflat = torch.cat([t.grad.reshape(-1) for W, b in lay for t in (W, b)])
    return ys.detach().numpy(), gflat.numpy(), y0o.grad.numpy()


CASES = [
    # B, L,  C, s, depth, H, width, n_hidden, dt0, scale
    (3, 40, 3, 4, 2, 16, 8, 2, 0.05, 1.0),
    (2, 40, 3, 4, 1, 16, 8, 2, 0.05, 1.0),
    (2, 60, 2, 5, 2, 8, 8, 1, 0.07, 1.0),
    (2, 36, 5, 3, 2, 24, 12, 3, 0.031, 2.0),
    (2, 240, 7, 12, 2, 128, 32, 2, 0.05, 3.0),    # EigenWorms dims, short path
    (2, 240, 6, 12, 2, 128, 32, 2, 0.05, 3.0),    # EigenWorms dims without time channel (fast path, C=6)
    (3, 120, 4, 12, 2, 128, 32, 2, 0.07, 3.0),    # fast path, C=4
    (2, 240, 7, 12, 1, 128, 32, 2, 0.05, 3.0),    # fast path, depth 1
    (2, 60, 6, 4, 2, 64, 128, 3, 0.05, 4.0),      # toy dims: weights stream from global/L2
    (2, 100, 7, 10, 2, 128, 64, 3, 0.1, 4.0),     # PPG depth-2 dims
    (2, 100, 7, 10, 1, 128, 64, 3, 0.1, 4.0),     # PPG reference config (depth 1)
    (2, 50, 8, 5, 2, 32, 16, 4, 0.11, 2.0),
]


@pytest.mark.parametrize("case", CASES)
def test_solve_forward_and_adjoint(built_lib, case):
    B, L, C, s, depth, H, width, n_hidden, dt0, scale = case
    x, logsig, rows, sc, grid, layers, y0 = _setup(B, L, C, s, depth, H, width, n_hidden, dt0, scale)
    dims = (H, C, width, n_hidden, depth)
    g_ys = torch.from_numpy(np.random.default_rng(1).normal(size=(B, len(grid), H)))
    ys_g, gp_g, gy_g = _run_gpu(layers, logsig, y0, rows, sc, dims, g_ys)
    ys_o, gp_o, gy_o = _run_oracle(layers, logsig, y0, rows, sc, dims, g_ys)
    assert rel_err(ys_g, ys_o) < TOL
    assert rel_err(gy_g, gy_o) < TOL
    # per-layer gradient blocks
    off = 0
    sizes = O.mlp_sizes(H, C, width, n_hidden)
    for i, o in zip(sizes[:-1], sizes[1:]):
        for n in (o * i, o):
            assert rel_err(gp_g[off : off + n], gp_o[off : off + n]) < TOL, (case, off)
            off += n


def test_depth2_with_zero_level2_equals_depth1(built_lib):
    B, L, C, s, H, width, n_hidden = 2, 40, 3, 4, 16, 8, 2
    x, logsig, rows, sc, grid, layers, y0 = _setup(B, L, C, s, 2, H, width, n_hidden, 0.05, 1.0)
    ls2 = logsig.copy()
    ls2[..., 1 + C :] = 0
    g = torch.zeros(B, len(grid), H, dtype=torch.float64)
    a, _, _ = _run_gpu(layers, ls2, y0, rows, sc, (H, C, width, n_hidden, 2), g)
    b, _, _ = _run_gpu(layers, ls2[..., : 1 + C].copy(), y0, rows, sc, (H, C, width, n_hidden, 1), g)
    assert np.allclose(a, b, rtol=0, atol=1e-6)


@pytest.mark.parametrize("classification", [True, False])
def test_model_loss_and_grads(built_lib, classification):
    """LogNeuralCDE.__call__ + train.py losses vs the oracle, through create_model."""
    from lncde.models.generate_model import create_model
    from lncde import train as T

    B, L, C, s, depth, H, width, n_hidden = 4, 64, 3, 4, 2, 16, 8, 2
    label_dim = 3 if classification else 1
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float32)
    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    model, state = create_model("log_ncde", C, logsig.shape[-1], depth, iv, label_dim, H, vf_depth=n_hidden,
                                vf_width=width, classification=classification, output_step=8, dt0=0.04,
                                scale=1.0, lambd=0.01, key=7)
    assert state is None
    X = (torch.from_numpy(np.tile(ts, (B, 1))).cuda(), torch.from_numpy(logsig).cuda(),
         torch.from_numpy(x[:, 0, :].copy()).cuda())
    if classification:
        y = torch.nn.functional.one_hot(torch.tensor([0, 2, 1, 1]), 3).float().cuda()
        loss_fn = T.classification_loss
    else:
        n_out = len(O.regression_times(L, 8)) + 1
        y = torch.from_numpy(rng.normal(size=(B, n_out)).astype(np.float32)).cuda()
        loss_fn = T.regression_loss
    loss, grads = loss_fn(model, X, y)
    # oracle in fp64 on the same parameters
    p = model.params.detach().cpu().double()
    lay = [(l.weight.detach().cpu().double().requires_grad_(True), l.bias.detach().cpu().double().requires_grad_(True))
           for l in model.vf.mlp.layers]
    W1 = model.linear1.weight.detach().cpu().double().requires_grad_(True)
    b1 = model.linear1.bias.detach().cpu().double().requires_grad_(True)
    W2 = model.linear2.weight.detach().cpu().double().requires_grad_(True)
    b2 = model.linear2.bias.detach().cpu().double().requires_grad_(True)
    rows, sc, grid = O.stage_table(ts, iv, 0.04, logsig.shape[1])
    y0 = torch.from_numpy(x[:, 0, :]).double() @ W1.T + b1
    ys = O.solve(lay, torch.from_numpy(logsig).double(), y0, rows, sc, C, H, depth)
    if classification:
        pred = torch.softmax(ys[:, -1] @ W2.T + b2, -1)
        lo = O.classification_loss(pred, y.cpu().double())
    else:
        out = torch.cat([O.interp_outputs(ys, grid, O.regression_times(L, 8)), ys[:, -1:]], 1)
        pred = torch.tanh(out @ W2.T + b2)
        lo = O.regression_loss(pred, y.cpu().double())
    lo = lo + O.lip2_norm(lay, 0.01)
    lo.backward()
    go = torch.cat([t.grad.reshape(-1) for W, b in lay for t in (W, b)] +
                   [W1.grad.reshape(-1), b1.grad, W2.grad.reshape(-1), b2.grad])
    assert abs(float(loss) - float(lo)) < TOL * max(1.0, abs(float(lo)))
    assert rel_err(grads.cpu().numpy(), go.numpy()) < TOL
    assert len(p) == len(go)


@pytest.mark.parametrize("lambd", [0.0, 1e-3])
def test_fused_classification_step_equals_autograd_step(built_lib, lambd):
    """The kernel-only train step (K5 around K3 / K3b: LogNeuralCDE.fused_classification_loss, the path bench.py times)
    against the torch-autograd formulation of the same loss (`kernel_flags("generic")`) at the EigenWorms model dims."""
    from lncde import _lib, train as T
    from lncde.models.generate_model import create_model

    B, L, C, s, H = 6, 480, 7, 12, 128
    rng = np.random.default_rng(2)
    x = np.cumsum(rng.normal(scale=0.05, size=(B, L, C)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, 2, np.float32)
    ts = paths.make_ts(L)
    model, _ = create_model("log_ncde", C, logsig.shape[-1], 2, paths.make_intervals(ts, s), 5, H, vf_depth=2, vf_width=32,
                            dt0=1.0 / 40, scale=10.0, lambd=lambd, key=7)
    X = (torch.from_numpy(np.tile(ts, (B, 1))).cuda(), torch.from_numpy(logsig).cuda(), torch.from_numpy(x[:, 0, :].copy()).cuda())
    y = torch.eye(5, device="cuda")[torch.arange(B, device="cuda") % 5]
    with _lib.kernel_flags("generic"):
        l0, g0 = T.classification_loss(model, X, y)
    l1, g1 = T.classification_loss(model, X, y)
    assert abs(float(l1) - float(l0)) < 1e-6 * max(1.0, abs(float(l0)))
    n_mlp, o2 = model.n_mlp, model.n_mlp + H * (C + 1)
    g0, g1 = g0.cpu().numpy(), g1.cpu().numpy()
    for a, b in ((0, n_mlp), (n_mlp, n_mlp + H * C), (n_mlp + H * C, o2), (o2, o2 + 5 * H), (o2 + 5 * H, g0.size)):
        assert rel_err(g1[a:b], g0[a:b]) < TOL, (a, b)
