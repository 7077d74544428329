"""K2 parity: CUDA D / BD / DE-LNCDE (through the C ABI) vs the torch-CPU fp64 oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import linear_ncde as OL
from oracle import paths

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _data(B, L, C, s, depth, seed=0):
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.normal(scale=0.1, size=(B, L, C)), axis=1)
    lo, hi = x.min((0, 1), keepdims=True), x.max((0, 1), keepdims=True)
    x = (2 * (x - lo) / (hi - lo) - 1).astype(np.float32)  # datasets.py:237-239
    return x, paths.calc_paths(x, s, depth, np.float32)


CASES = [
    # B, L, C, s, depth, H, bs, classification
    (3, 64, 3, 4, 2, 16, 1, True),
    (3, 64, 3, 4, 2, 16, 4, True),
    (2, 64, 3, 4, 2, 16, 16, True),
    (2, 64, 3, 4, 1, 16, 4, True),
    (2, 48, 2, 4, 3, 8, 4, True),      # depth-3 brackets
    (2, 48, 3, 4, 3, 12, 1, True),     # D-LNCDE ignores levels > 1 (quirk B7)
    (2, 64, 3, 4, 2, 16, 4, False),    # regression read-out
    (2, 896, 6, 16, 2, 64, 4, True),   # SCP1 BD-LNCDE shape
    (2, 240, 7, 12, 2, 128, 4, True),  # EigenWorms BD dims, short
    (2, 120, 7, 12, 2, 128, 128, True),  # DE dims, short
    (2, 120, 7, 12, 2, 32, 32, True),
]


@pytest.mark.parametrize("case", CASES)
def test_linear_ncde_loss_and_grads(built_lib, case):
    from lncde import train as T
    from lncde.models.generate_model import create_model

    B, L, C, s, depth, H, bs, classification = case
    x, logsig = _data(B, L, C, s, depth)
    label_dim = 3 if classification else 1
    name = {1: "diagonal_linear_ncde", H: "dense_linear_ncde"}.get(bs, "bd_linear_ncde")
    model, state = create_model(name, C, logsig.shape[-1], depth, None, label_dim, H, block_size=bs,
                                classification=classification, lambd=0.01, w_init_std=0.25, parallel_steps=128,
                                key=3)
    assert state is None
    X = (None, torch.from_numpy(logsig).cuda(), torch.from_numpy(x[:, 0].copy()).cuda())
    rng = np.random.default_rng(1)
    if classification:
        y = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, 3, B)), 3).float().cuda()
        loss, grads = T.classification_loss(model, X, y)
    else:
        y = torch.from_numpy(rng.normal(size=(B, logsig.shape[1])).astype(np.float32)).cuda()
        loss, grads = T.regression_loss(model, X, y)
    # oracle, fp64, same parameters
    vf_A = model.vf_A.detach().cpu().double().requires_grad_(True)
    Wi = model.init_layer.weight.detach().cpu().double().requires_grad_(True)
    bi = model.init_layer.bias.detach().cpu().double().requires_grad_(True)
    Wo = model.out_layer.weight.detach().cpu().double().requires_grad_(True)
    bo = model.out_layer.bias.detach().cpu().double().requires_grad_(True)
    pred = OL.predict(vf_A, Wi, bi, Wo, bo, torch.from_numpy(logsig).double(), torch.from_numpy(x[:, 0]).double(), C,
                      bs, depth, classification)
    yo = y.cpu().double()
    if classification:
        lo = torch.mean(-torch.sum(yo * torch.log(pred + 1e-8), dim=1))
    else:
        lo = torch.mean(torch.mean((pred[:, :, 0] - yo) ** 2, dim=1))
    lo = lo + OL.lip2_norm(vf_A, 0.01)
    lo.backward()
    go = torch.cat([vf_A.grad.reshape(-1), Wi.grad.reshape(-1), bi.grad, Wo.grad.reshape(-1), bo.grad])
    assert abs(float(loss) - float(lo)) < TOL * max(1.0, abs(float(lo)))
    g = grads.cpu().numpy()
    nA = vf_A.numel()
    assert rel_err(g[:nA], go.numpy()[:nA]) < TOL
    assert rel_err(g[nA:], go.numpy()[nA:]) < TOL


@pytest.mark.parametrize("B,H,bs", [(3, 128, 128), (20, 128, 128), (40, 128, 128), (5, 128, 64), (30, 256, 64)])
def test_dense_cluster_sweeps_equal_one_cta_sweeps(built_lib, B, H, bs):
    """linear_cluster.cuh (8 / 4 / 2 CTAs per series, by batch size) against the one-CTA sweeps (LNCDE_FLAG_K2_ONE_CTA):
    same flows, same row arithmetic forward (bit-equal states), same gradient up to summation order."""
    from lncde import _lib, train as T
    from lncde.models.generate_model import create_model

    C, s, depth, L = 7, 12, 2, 12 * 41
    x, logsig = _data(B, L, C, s, depth)
    name = "dense_linear_ncde" if bs == H else "bd_linear_ncde"
    model, _ = create_model(name, C, logsig.shape[-1], depth, None, 3, H, block_size=bs, classification=True,
                            lambd=0.01, w_init_std=0.25, parallel_steps=128, key=3)
    X = (None, torch.from_numpy(logsig).cuda(), torch.from_numpy(x[:, 0].copy()).cuda())
    y = torch.nn.functional.one_hot(torch.from_numpy(np.random.default_rng(1).integers(0, 3, B)), 3).float().cuda()
    ys_c = model.hidden_states(X).detach().cpu().numpy()
    loss_c, g_c = T.classification_loss(model, X, y)
    loss_c, g_c = float(loss_c), g_c.cpu().numpy().copy()
    with _lib.kernel_flags("k2_one_cta"):
        ys_1 = model.hidden_states(X).detach().cpu().numpy()
        loss_1, g_1 = T.classification_loss(model, X, y)
    assert np.array_equal(ys_c, ys_1)
    assert abs(loss_c - float(loss_1)) < 1e-6 * max(1.0, abs(loss_c))
    assert rel_err(g_c, g_1.cpu().numpy()) < TOL


def test_hidden_states_match_oracle_and_chunked_scan(built_lib):
    from lncde.models.generate_model import create_model

    B, L, C, s, depth, H, bs = 2, 600, 3, 4, 2, 16, 4
    x, logsig = _data(B, L, C, s, depth)
    model, _ = create_model("bd_linear_ncde", C, logsig.shape[-1], depth, None, 2, H, block_size=bs, key=0)
    X = (None, torch.from_numpy(logsig).cuda(), torch.from_numpy(x[:, 0].copy()).cuda())
    ys = model.hidden_states(X).detach().cpu().numpy()
    vf_A = model.vf_A.detach().cpu().double()
    for b in range(B):
        y0 = model.init_layer.weight.detach().cpu().double() @ torch.from_numpy(x[b, 0]).double() \
            + model.init_layer.bias.detach().cpu().double()
        seq = OL.forward(vf_A, torch.from_numpy(logsig[b]).double(), y0, C, bs, depth, 1).numpy()
        par = OL.forward(vf_A, torch.from_numpy(logsig[b]).double(), y0, C, bs, depth, 128).numpy()
        assert rel_err(par, seq) < 1e-12  # chunk/remainder path == sequential path (quirk B6)
        assert rel_err(ys[b], seq) < TOL


def test_diagonal_equals_1x1_blocks(built_lib):
    """D-LNCDE == BD with 1x1 blocks at depth 1 (SURVEY 8c identity 3)."""
    L = built_lib
    B, T, C, H = 2, 50, 3, 8
    rng = np.random.default_rng(0)
    ls = torch.from_numpy(rng.normal(scale=0.1, size=(B, T, 1 + C)).astype(np.float32)).cuda()
    lie = torch.from_numpy(rng.normal(size=(H, 1, 1, C)).astype(np.float32)).cuda()
    y0 = torch.from_numpy(rng.normal(size=(B, H)).astype(np.float32)).cuda()
    from lncde.models.LinearNeuralCDEs import _LinearScan

    ys = _LinearScan.apply(lie, ls, y0, (H, 1, C)).cpu().numpy()
    flows = 1 + ls.cpu().numpy()[:, :, 1:].astype(np.float64) @ lie.cpu().numpy().reshape(H, C).T.astype(np.float64)
    want = np.cumprod(np.concatenate([np.ones((B, 1, H)), flows[:, 1:]], 1), axis=1) * y0.cpu().numpy()[:, None, :]
    assert rel_err(ys, want) < TOL


@pytest.mark.parametrize("name,bs", [("diagonal_linear_ncde", 1), ("bd_linear_ncde", 4), ("dense_linear_ncde", 128)])
def test_fused_lncde_classification_step_equals_autograd_step(built_lib, name, bs):
    """LogLinearCDE.fused_classification_loss (the library's default kernels + K5) against the autograd formulation on the
    baseline kernels (`kernel_flags("generic")`), EigenWorms model dims, T > 65 so the time-parallel scans run."""
    from lncde import _lib, train as T
    from lncde.models.generate_model import create_model

    B, L, C, s, H, nl = 4, 1200, 7, 12, 128, 5
    rng = np.random.default_rng(4)
    x = np.cumsum(rng.normal(scale=0.02, size=(B, L, C)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, 2, np.float32)
    model, _ = create_model(name, C, logsig.shape[-1], 2, None, nl, H, block_size=bs, lambd=1e-3, key=5)
    X = (None, torch.from_numpy(logsig).cuda(), torch.from_numpy(x[:, 0, :].copy()).cuda())
    y = torch.eye(nl, device="cuda")[torch.arange(B, device="cuda") % nl]
    with _lib.kernel_flags("generic"):
        l0, g0 = T.classification_loss(model, X, y)
    l1, g1 = T.classification_loss(model, X, y)
    assert abs(float(l1) - float(l0)) < 2e-6 * max(1.0, abs(float(l0)))
    nA, o2 = model.n_A, model.n_A + H * (C + 1)
    g0, g1 = g0.cpu().numpy(), g1.cpu().numpy()
    for a, b in ((0, nA), (nA, nA + H * C), (nA + H * C, o2), (o2, o2 + nl * H), (o2 + nl * H, g0.size)):
        assert rel_err(g1[a:b], g0[a:b]) < TOL, (name, a, b)


@pytest.mark.parametrize("C,depth,nb,bs", [(7, 2, 1, 128), (7, 3, 1, 128), (7, 2, 32, 4), (6, 3, 8, 16), (3, 4, 1, 64)])
def test_lie_bracket_kernels_equal_torch_log_ode(built_lib, C, depth, nb, bs):
    """K2a behind LogLinearCDE.log_ode (values and gradient) against the batched torch.matmul formulation
    (`kernel_flags("generic")`) at the dense / block-diagonal EigenWorms shapes, depth 2 - 4."""
    from lncde import _lib
    from lncde.models.generate_model import create_model

    model, _ = create_model("bd_linear_ncde", C, 1, depth, None, 2, nb * bs, block_size=bs, lambd=0.0, key=2)
    g = torch.Generator().manual_seed(C * 10 + depth)
    vfs = (torch.randn(C, nb, bs, bs, generator=g) / bs ** 0.5).cuda().requires_grad_(True)
    with _lib.kernel_flags("generic"):
        want = model.log_ode(vfs)
    got = model.log_ode(vfs)
    assert got.shape == want.shape
    assert rel_err(got.detach().cpu().numpy(), want.detach().cpu().numpy()) < 2e-6
    w = torch.randn(want.shape, generator=g).cuda()
    (g_want,) = torch.autograd.grad(want, vfs, w)
    (g_got,) = torch.autograd.grad(got, vfs, w)
    assert rel_err(g_got.cpu().numpy(), g_want.cpu().numpy()) < TOL
