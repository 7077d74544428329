"""The oracle against outputs of the REFERENCE'S OWN CODE.

tests/golden/ref_*.npz were produced by tests/golden/make_golden_ref.py, which imports the reference's source
files from /root/reference and executes them unmodified under stand-ins for the libraries that cannot be
installed offline (tests/golden/refshim: jax, equinox, diffrax, signax, roughpy).  These tests pin, to the
reference's code: the windowing / basepoint / remainder quirks of calc_paths and the 128-sample chunking of
batch_calc_paths, `pairs`, the Log-ODE field (searchsorted row and interval indexing, JVP reshape, bracket
pairing and sign, division by the interval length), create_model's argument plumbing, the LNCDE flow
construction, bracket recursion and chunked scanner, the read-outs and both loss functions with their Lip(2)
terms.  Third-party semantics (signax / diffrax / equinox / roughpy) stay restated - see refshim/README.md."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import hall as OH
from oracle import linear_ncde as OL
from oracle import log_ncde as O
from oracle import nrde as ON
from oracle import paths

P = np.load(os.path.join(GOLDEN, "ref_paths.npz"))
N = np.load(os.path.join(GOLDEN, "ref_logncde.npz"))
LN = np.load(os.path.join(GOLDEN, "ref_linear.npz"))
NR = np.load(os.path.join(GOLDEN, "ref_nrde.npz"))


@pytest.mark.parametrize("i", range(int(P["n"])))
def test_calc_paths_matches_reference_code(i):
    B, L, C, s, d = (int(v) for v in P[f"cfg{i}"])
    got = paths.calc_paths(P[f"x{i}"], s, d, np.float64)
    want = P[f"logsig{i}"]
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-12


def test_batch_calc_paths_chunking_matches_reference_code():
    s, d = (int(v) for v in P["cfgb"])
    got = paths.batch_calc_paths(P["xb"], s, d, np.float64)
    assert np.array_equal(got, P["logsigb"])  # integer path: exact


def logncde_oracle(i, dtype=torch.float64):
    """(pred, loss, flat gradient in the layout [mlp | linear1 | linear2])."""
    B, L, C, s, depth, H, width, nh, cls, ostep, ld = (int(v) for v in N[f"cfg{i}"])
    dt0, scale, lambd = (float(v) for v in N[f"hyper{i}"])
    sizes = O.mlp_sizes(H, C, width, nh)
    flat, off, lay = torch.from_numpy(N[f"mlp{i}"]), 0, []
    for a, b in zip(sizes[:-1], sizes[1:]):
        lay.append((flat[off : off + a * b].view(b, a).clone().requires_grad_(True),
                    flat[off + a * b : off + a * b + b].clone().requires_grad_(True)))
        off += b * (a + 1)
    l1, l2 = torch.from_numpy(N[f"lin1_{i}"]), torch.from_numpy(N[f"lin2_{i}"])
    W1, b1, W2, b2 = (t.clone().requires_grad_(True) for t in
                      (l1[: H * C].view(H, C), l1[H * C :], l2[: ld * H].view(ld, H), l2[ld * H :]))
    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    logsig = torch.from_numpy(N[f"logsig{i}"])
    rows, sc, grid = O.stage_table(ts, iv, dt0, logsig.shape[1])
    y0 = torch.from_numpy(N[f"x{i}"][:, 0]) @ W1.T + b1
    ys = O.solve(lay, logsig, y0, rows, sc, C, H, depth)
    y = torch.from_numpy(N[f"y{i}"])
    if cls:
        pred = torch.softmax(ys[:, -1] @ W2.T + b2, -1)
        loss = O.classification_loss(pred, y)
    else:
        out = torch.cat([O.interp_outputs(ys, grid, O.regression_times(L, ostep)), ys[:, -1:]], 1)
        pred = torch.tanh(out @ W2.T + b2)
        loss = O.regression_loss(pred, y)
    loss = loss + O.lip2_norm(lay, lambd)
    grads = torch.autograd.grad(loss, [t for Wb in lay for t in Wb] + [W1, b1, W2, b2])
    return pred.detach(), loss.detach(), torch.cat([g.reshape(-1) for g in grads])


@pytest.mark.parametrize("i", range(int(N["n"])))
def test_logncde_matches_reference_code(i):
    B, L, C, s, depth = (int(v) for v in N[f"cfg{i}"][:5])
    # logsig rows fed to the model came from the reference's calc_paths
    assert rel_err(paths.calc_paths(N[f"x{i}"], s, depth, np.float64), N[f"logsig{i}"]) < 1e-12
    if depth == 2:
        assert np.array_equal(np.asarray(OH.Hall(C, 2).basis[1:]), N[f"pairs{i}"])
    pred, loss, grad = logncde_oracle(i)
    # the oracle folds dt / (interval length) into ONE fp32 factor, the reference divides then multiplies: 1e-7
    assert rel_err(pred.numpy(), N[f"pred{i}"]) < 5e-7
    assert abs(float(loss) - float(N[f"loss{i}"])) < 5e-7 * max(1.0, abs(float(N[f"loss{i}"])))
    # gradient of the reference's own loss function (torch reverse mode THROUGH the reference's code)
    assert rel_err(grad.numpy(), N[f"grad{i}"]) < 2e-6


def nrde_oracle(i, dtype=torch.float64):
    """(pred, loss, flat gradient) in the fixture's layout [vf layers | mlp_linear | linear1 | linear2]."""
    B, L, C, s, depth, H, width, nvf, cls, ostep, ld = (int(v) for v in NR[f"cfg{i}"])
    dt0, scale = (float(v) for v in NR[f"hyper{i}"])
    logsig = torch.from_numpy(NR[f"logsig{i}"])
    Dm = logsig.shape[2] - 1
    flat = torch.from_numpy(NR[f"params{i}"]).to(dtype)
    n_field = sum(o * (a + 1) for a, o in ON.layer_sizes(H, Dm, width, nvf))
    lay = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True))
           for W, b in ON.unflatten_params(flat[:n_field], H, Dm, width, nvf)]
    rest = flat[n_field:]
    W1, b1, W2, b2 = (t.clone().requires_grad_(True) for t in
                      (rest[: H * C].view(H, C), rest[H * C : H * C + H], rest[H * C + H : H * C + H + ld * H].view(ld, H),
                       rest[H * C + H + ld * H :]))
    ts = paths.make_ts(L)
    rows, sc, grid = O.stage_table(ts, paths.make_intervals(ts, s), dt0, logsig.shape[1])
    y0 = torch.from_numpy(NR[f"x{i}"][:, 0]) @ W1.T + b1
    ys = ON.solve(lay, logsig, y0, rows, sc, H)
    y = torch.from_numpy(NR[f"y{i}"])
    if cls:
        pred = torch.softmax(ys[:, -1] @ W2.T + b2, -1)
        loss = O.classification_loss(pred, y)
    else:
        out = torch.cat([O.interp_outputs(ys, grid, O.regression_times(L, ostep)), ys[:, -1:]], 1)
        pred = torch.tanh(out @ W2.T + b2)
        loss = O.regression_loss(pred, y)
    grads = torch.autograd.grad(loss, [t for Wb in lay for t in Wb] + [W1, b1, W2, b2])
    return pred.detach(), loss.detach(), torch.cat([g.reshape(-1) for g in grads])


@pytest.mark.parametrize("i", range(int(NR["n"])))
def test_nrde_matches_reference_code(i):
    """oracle/nrde.py against NeuralRDE.__call__ / create_model("nrde") / the reference's loss functions."""
    pred, loss, grad = nrde_oracle(i)
    assert rel_err(pred.numpy(), NR[f"pred{i}"]) < 5e-7
    assert abs(float(loss) - float(NR[f"loss{i}"])) < 5e-7 * max(1.0, abs(float(NR[f"loss{i}"])))
    assert rel_err(grad.numpy(), NR[f"grad{i}"]) < 2e-6  # fixture gradients are stored as fp32


def linear_oracle(i):
    B, L, C, s, depth, H, bs, Psteps, cls, ld = (int(v) for v in LN[f"cfg{i}"])
    lambd = float(LN[f"lambd{i}"])
    vfA = torch.from_numpy(LN[f"vfA{i}"]).clone().requires_grad_(True)
    ini, out = torch.from_numpy(LN[f"init{i}"]), torch.from_numpy(LN[f"out{i}"])
    Wi, bi, Wo, bo = (t.clone().requires_grad_(True) for t in
                      (ini[: H * C].view(H, C), ini[H * C :], out[: ld * H].view(ld, H), out[ld * H :]))
    pred = OL.predict(vfA, Wi, bi, Wo, bo, torch.from_numpy(LN[f"logsig{i}"]), torch.from_numpy(LN[f"x{i}"][:, 0]), C, bs,
                      depth, bool(cls), Psteps)
    y = torch.from_numpy(LN[f"y{i}"])
    loss = O.classification_loss(pred, y) if cls else O.regression_loss(pred, y)
    loss = loss + OL.lip2_norm(vfA, lambd)
    grads = torch.autograd.grad(loss, [vfA, Wi, bi, Wo, bo])
    return pred.detach(), loss.detach(), torch.cat([g.reshape(-1) for g in grads])


@pytest.mark.parametrize("i", range(int(LN["n"])))
def test_linear_ncde_matches_reference_code(i):
    pred, loss, grad = linear_oracle(i)
    assert rel_err(pred.numpy(), LN[f"pred{i}"]) < 1e-12
    assert abs(float(loss) - float(LN[f"loss{i}"])) < 1e-12 * max(1.0, abs(float(LN[f"loss{i}"])))
    assert rel_err(grad.numpy(), LN[f"grad{i}"]) < 1e-11


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference not mounted")
def test_fixtures_are_reproducible_from_the_reference(tmp_path):
    """Live re-derivation: re-run the generator against /root/reference and compare with the committed files."""
    import shutil
    import subprocess
    import sys

    work = tmp_path / "golden"
    shutil.copytree(GOLDEN, work, ignore=shutil.ignore_patterns("ref_*.npz", "__pycache__"))
    r = subprocess.run([sys.executable, str(work / "make_golden_ref.py")], capture_output=True, text=True, timeout=600,
                       env={**os.environ, "PYTHONPATH": ""})
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("ref_paths.npz", "ref_logncde.npz", "ref_linear.npz", "ref_nrde.npz"):
        a, b = np.load(os.path.join(GOLDEN, name)), np.load(str(work / name))
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            if a[k].dtype.kind in "fc":
                tol = 1e-12 if a[k].dtype == np.float64 else 1e-6
                assert np.allclose(a[k], b[k], rtol=tol, atol=tol * 1e-2), (name, k)
            else:
                assert np.array_equal(a[k], b[k]), (name, k)
