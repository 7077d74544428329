"""Parity and size-independent properties at BASELINE.json's FULL sizes (EigenWorms shape:
L = 17 984, stepsize 12 -> 1 499 logsig rows / Heun steps)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import linear_ncde as OL
from oracle import log_ncde as O
from oracle import paths

pytestmark = pytest.mark.gpu

L_EW, S_EW, C_EW, H_EW = 17984, 12, 7, 128
DT0 = 0.00066711140760507


def _ew_paths(B, seed=11):
    from lncde.data_dir.datasets import synthetic_paths

    return synthetic_paths("EigenWorms", B, seed=seed)


def _ew_model(depth=2, scale=1000.0, key=5):
    from lncde.models.generate_model import create_model

    ts = paths.make_ts(L_EW)
    iv = paths.make_intervals(ts, S_EW)
    D = 29 if depth == 2 else 8
    model, _ = create_model("log_ncde", C_EW, D, depth, iv, 5, H_EW, vf_depth=2, vf_width=32, dt0=DT0, scale=scale,
                            lambd=1e-3, key=key)
    return model, ts, iv


def test_logncde_full_length_matches_oracle(built_lib):
    """One EigenWorms-shaped sample through all 1 499 Heun steps vs the fp64 oracle
    (reference formulation).  Tolerance: the north-star 1e-5, norm-relative on the hidden-state trajectory (measured on
    B200: 1.6e-6, profiles/r02_parity_errors.md)."""
    from lncde.data_dir.generate_paths import calc_paths

    x = _ew_paths(1)
    model, ts, iv = _ew_model(scale=50.0)  # larger field than the default 1000 so the test is not vacuous
    xg = torch.from_numpy(x).cuda()
    ls = calc_paths(xg, S_EW, 2)
    X = (torch.from_numpy(ts)[None], ls, xg[:, 0, :].contiguous())
    with torch.no_grad():
        ys, _ = model.hidden_states(X)
    ys = ys.cpu().numpy()
    assert ys.shape == (1, 1500, H_EW)
    lay = [(l.weight.detach().cpu().double(), l.bias.detach().cpu().double()) for l in model.vf.mlp.layers]
    rows, sc, _ = O.stage_table(ts, iv, DT0, 1499)
    y0 = torch.from_numpy(x[:, 0]).double() @ model.linear1.weight.detach().cpu().double().T \
        + model.linear1.bias.detach().cpu().double()
    with torch.no_grad():
        want = O.solve(lay, ls.cpu().double(), y0, rows, sc, C_EW, H_EW, 2).numpy()
    assert np.abs(want[0, -1] - want[0, 0]).max() > 1e-3  # the state really moves
    assert rel_err(ys, want) < 1e-5


def test_logncde_full_length_loss_and_gradient_match_oracle(built_lib):
    """The bench configuration itself (scale = 1000, lambd = 1e-3, 1 499 Heun steps; B = 4 so that the fp64 oracle -
    reference formulation, torch autograd through all 2 998 stages - finishes in about a minute): loss and the gradient
    of every parameter block through train.classification_loss, at the north-star tolerance 1e-5."""
    from lncde import train as T
    from lncde.data_dir.generate_paths import calc_paths

    B = 4
    x = _ew_paths(B, seed=2345)
    model, ts, iv = _ew_model(scale=1000.0)
    xg = torch.from_numpy(x).cuda()
    ls = calc_paths(xg, S_EW, 2)
    X = (torch.from_numpy(ts)[None].expand(B, -1), ls, xg[:, 0, :].contiguous())
    y = torch.eye(5, device="cuda")[torch.arange(B) % 5]
    loss, grads = T.classification_loss(model, X, y)
    lay = [(l.weight.detach().cpu().double().requires_grad_(True), l.bias.detach().cpu().double().requires_grad_(True))
           for l in model.vf.mlp.layers]
    W1 = model.linear1.weight.detach().cpu().double().requires_grad_(True)
    b1 = model.linear1.bias.detach().cpu().double().requires_grad_(True)
    W2 = model.linear2.weight.detach().cpu().double().requires_grad_(True)
    b2 = model.linear2.bias.detach().cpu().double().requires_grad_(True)
    rows, sc, _ = O.stage_table(ts, iv, DT0, 1499)
    y0 = torch.from_numpy(x[:, 0]).double() @ W1.T + b1
    ys = O.solve(lay, ls.cpu().double(), y0, rows, sc, C_EW, H_EW, 2)
    pred = torch.softmax(ys[:, -1] @ W2.T + b2, -1)
    lo = O.classification_loss(pred, y.cpu().double()) + O.lip2_norm(lay, 1e-3)
    lo.backward()
    assert abs(float(loss) - float(lo.detach())) < 1e-5 * max(1.0, abs(float(lo.detach())))
    g = grads.cpu().numpy()
    blocks = [t.grad.reshape(-1).numpy() for W, b in lay for t in (W, b)] + \
        [W1.grad.reshape(-1).numpy(), b1.grad.numpy(), W2.grad.reshape(-1).numpy(), b2.grad.numpy()]
    off = 0
    for k, want in enumerate(blocks):
        assert rel_err(g[off:off + want.size], want) < 1e-5, ("parameter block", k)
        off += want.size
    assert off == g.size


def test_logncde_zero_control_is_identity(built_lib):
    model, ts, iv = _ew_model()
    B = 4
    ls = torch.zeros(B, 1499, 29, device="cuda")
    x0 = torch.randn(B, C_EW, device="cuda")
    with torch.no_grad():
        ys, _ = model.hidden_states((torch.from_numpy(ts)[None].expand(B, -1), ls, x0))
    assert torch.equal(ys, ys[:, :1].expand_as(ys))  # a zero log-signature drives nothing: exact


def test_logncde_depth2_with_zero_areas_equals_depth1(built_lib):
    from lncde.data_dir.generate_paths import calc_paths

    x = _ew_paths(4)
    xg = torch.from_numpy(x).cuda()
    ls2 = calc_paths(xg, S_EW, 2)
    ls2[..., 1 + C_EW:] = 0
    m2, ts, iv = _ew_model(depth=2, scale=50.0)
    m1, _, _ = _ew_model(depth=1, scale=50.0)
    m1.params.data.copy_(m2.params.data)  # same seed -> same layout and values; make it explicit
    tsb = torch.from_numpy(ts)[None].expand(4, -1)
    x0 = xg[:, 0, :].contiguous()
    with torch.no_grad():
        a, _ = m2.hidden_states((tsb, ls2, x0))
        b, _ = m1.hidden_states((tsb, ls2[..., : 1 + C_EW].contiguous(), x0))
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-6


def test_bd_lncde_full_length_matches_oracle(built_lib):
    """EigenWorms BD-LNCDE (H = 128, 4x4 blocks, T = 1 499): hidden states of one sample vs the fp64
    oracle at the north-star tolerance 1e-5 (the recurrence multiplies 1 498 flows; measured on B200: 2.6e-6)."""
    from lncde.data_dir.datasets import synthetic_paths
    from lncde.data_dir.generate_paths import calc_paths
    from lncde.models.generate_model import create_model

    x = synthetic_paths("EigenWorms", 2, seed=3, scale_to_unit=True)
    xg = torch.from_numpy(x).cuda()
    ls = calc_paths(xg, S_EW, 2)
    model, _ = create_model("bd_linear_ncde", C_EW, 29, 2, None, 5, H_EW, block_size=4, lambd=1e-3, key=9)
    with torch.no_grad():
        ys = model.hidden_states((None, ls, xg[:, 0, :].contiguous())).cpu().numpy()
    assert ys.shape == (2, 1499, H_EW)
    vf_A = model.vf_A.detach().cpu().double()
    y0 = model.init_layer.weight.detach().cpu().double() @ torch.from_numpy(x[0, 0]).double() \
        + model.init_layer.bias.detach().cpu().double()
    with torch.no_grad():
        want = OL.forward(vf_A, ls[0].cpu().double(), y0, C_EW, 4, 2, 1).numpy()
    assert np.isfinite(want).all()
    assert rel_err(ys[0], want) < 1e-5


def test_full_batch_train_step_is_finite_and_deterministic(built_lib):
    """The bench workload once, eagerly: finite loss, gradients reproducible bit-for-bit (the split-K
    reductions are ordered, there are no atomics on the path)."""
    from lncde import train as T
    from lncde.data_dir.generate_paths import calc_paths

    x = _ew_paths(32, seed=2345)
    xg = torch.from_numpy(x).cuda()
    model, ts, iv = _ew_model()
    y = torch.eye(5, device="cuda")[torch.arange(32) % 5]
    ls = calc_paths(xg, S_EW, 2)
    X = (torch.from_numpy(ts)[None].expand(32, -1), ls, xg[:, 0, :].contiguous())
    l1, g1 = T.classification_loss(model, X, y)
    l2, g2 = T.classification_loss(model, X, y)
    assert torch.isfinite(l1) and torch.isfinite(g1).all()
    assert torch.equal(l1, l2) and torch.allclose(g1, g2, rtol=1e-6, atol=1e-9)
    assert float(g1.abs().max()) > 0
