"""CUDA path (through the host mirrors and the C ABI) against outputs of the REFERENCE'S OWN CODE:
tests/golden/ref_*.npz, produced by tests/golden/make_golden_ref.py from /root/reference (see
tests/test_oracle_reference_pinned.py for what those fixtures pin).  Tolerance: BASELINE north-star 1e-5
relative in fp32 (norm-relative per tensor); the fixtures are float64."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5
P = np.load(os.path.join(GOLDEN, "ref_paths.npz"))
N = np.load(os.path.join(GOLDEN, "ref_logncde.npz"))
LN = np.load(os.path.join(GOLDEN, "ref_linear.npz"))


def _ts(L):
    return (np.float32(1.0 / L) * np.arange(L, dtype=np.float32)).astype(np.float32)


def _intervals(ts, s):
    return np.concatenate([ts[np.arange(0, len(ts), s)], np.array([1.0], np.float32)]).astype(np.float32)


@pytest.mark.parametrize("i", range(int(P["n"])))
def test_calc_paths_vs_reference_code(built_lib, i):
    from lncde.data_dir.generate_paths import calc_paths

    B, L, C, s, d = (int(v) for v in P[f"cfg{i}"])
    got = calc_paths(torch.from_numpy(P[f"x{i}"]).float().cuda(), s, d).cpu().numpy()
    assert got.shape == P[f"logsig{i}"].shape
    assert rel_err(got, P[f"logsig{i}"]) < TOL


def test_batch_calc_paths_vs_reference_code(built_lib):
    from lncde.data_dir.generate_paths import batch_calc_paths

    s, d = (int(v) for v in P["cfgb"])
    got = batch_calc_paths(torch.from_numpy(P["xb"]).float().cuda(), s, d).cpu().numpy()
    assert np.array_equal(got.astype(np.float64), P["logsigb"])  # integer path: bit-exact


@pytest.mark.parametrize("i", range(int(N["n"])))
def test_log_ncde_vs_reference_code(built_lib, i):
    from lncde import train as T
    from lncde.models.generate_model import create_model

    B, L, C, s, depth, H, width, nh, cls, ostep, ld = (int(v) for v in N[f"cfg{i}"])
    dt0, scale, lambd = (float(v) for v in N[f"hyper{i}"])
    ts = _ts(L)
    model, state = create_model("log_ncde", C, N[f"logsig{i}"].shape[-1], depth, _intervals(ts, s), ld, H, vf_depth=nh,
                                vf_width=width, classification=bool(cls), output_step=ostep, dt0=dt0, scale=scale,
                                lambd=lambd, key=1)
    assert state is None
    flat = np.concatenate([N[f"mlp{i}"], N[f"lin1_{i}"], N[f"lin2_{i}"]])
    assert flat.size == model.params.numel()
    with torch.no_grad():
        model.params.copy_(torch.from_numpy(flat).float())
    if depth == 2:
        assert np.array_equal(model.pairs.numpy(), N[f"pairs{i}"])
    X = (torch.from_numpy(np.tile(ts, (B, 1))), torch.from_numpy(N[f"logsig{i}"]).float().cuda(),
         torch.from_numpy(N[f"x{i}"][:, 0, :].copy()).float().cuda())
    pred = model(X)
    assert rel_err(pred.detach().cpu().numpy(), N[f"pred{i}"]) < TOL
    y = torch.from_numpy(N[f"y{i}"]).float().cuda()
    loss, grads = (T.classification_loss if cls else T.regression_loss)(model, X, y)
    want = float(N[f"loss{i}"])
    assert abs(float(loss) - want) < TOL * max(1.0, abs(want))
    # gradient of the reference's own loss function, same flat layout [mlp | linear1 | linear2]
    assert rel_err(grads.cpu().numpy(), N[f"grad{i}"]) < TOL


@pytest.mark.parametrize("i", range(int(LN["n"])))
def test_linear_ncde_vs_reference_code(built_lib, i):
    from lncde import train as T
    from lncde.models.generate_model import create_model

    B, L, C, s, depth, H, bs, Psteps, cls, ld = (int(v) for v in LN[f"cfg{i}"])
    name = str(LN[f"name{i}"])
    model, state = create_model(name, C, LN[f"logsig{i}"].shape[-1], depth, None, ld, H, block_size=bs,
                                classification=bool(cls), lambd=float(LN[f"lambd{i}"]), parallel_steps=Psteps, key=1)
    assert state is None
    assert str(model.basis_list) == str(LN[f"basis{i}"])
    flat = np.concatenate([LN[f"vfA{i}"].ravel(), LN[f"init{i}"], LN[f"out{i}"]])
    assert flat.size == model.params.numel()
    with torch.no_grad():
        model.params.copy_(torch.from_numpy(flat).float())
    X = (torch.from_numpy(np.tile(_ts(L), (B, 1))), torch.from_numpy(LN[f"logsig{i}"]).float().cuda(),
         torch.from_numpy(LN[f"x{i}"][:, 0, :].copy()).float().cuda())
    pred = model(X)
    assert rel_err(pred.detach().cpu().numpy(), LN[f"pred{i}"]) < TOL
    y = torch.from_numpy(LN[f"y{i}"]).float().cuda()
    loss, grads = (T.classification_loss if cls else T.regression_loss)(model, X, y)
    want = float(LN[f"loss{i}"])
    assert abs(float(loss) - want) < TOL * max(1.0, abs(want))
    assert rel_err(grads.cpu().numpy(), LN[f"grad{i}"]) < TOL  # layout [vf_A | init_layer | out_layer]


@pytest.mark.parametrize("i", range(7))
def test_slice_variants_vs_reference_code(built_lib, monkeypatch, i):
    """Walsh-Hadamard / DPLR / sparse / diagonal + dense LogLinearCDE (tests/golden/ref_slice.npz) on the K2 kernels."""
    import test_slice_variants_cpu as V

    model, X, y, cls, kind = V.build(i, monkeypatch, device="cuda")
    V.check(i, model, X, y, cls, kind, TOL, grad_tol=TOL)  # measured on B200: 4.8e-7 (profiles/r02_parity_errors.md)


@pytest.mark.parametrize("inmemory", [True, False])
@pytest.mark.parametrize("i", range(2))
def test_data_pipeline_vs_reference_code(built_lib, i, inmemory):
    """dataset_generator + Dataloader on the device (log-signatures from K1, one launch per split) against the
    reference's own dataset_generator (tests/golden/ref_data.npz); inmemory=False keeps pinned host copies."""
    import test_data_pipeline_cpu as DP

    ds = DP.build(i, device="cuda", paths_fn=None, inmemory=inmemory)
    (ts, ls, x0), y = next(ds.path_dataloaders["train"].loop_epoch(4))
    assert ls.is_cuda and x0.is_cuda and y.is_cuda
    DP.check(i, ds)
