"""Host logic of the remaining SLiCE structures of LogLinearCDE (Walsh-Hadamard, diagonal + low rank, sparse,
diagonal + dense) against outputs of the reference's own code (tests/golden/ref_slice.npz,
made by tests/golden/make_golden_ref.py).  These structures only change how the flows are PARAMETERISED; the
scans are the K2 kernels tested elsewhere, replaced here by a plain-torch statement of the kernel contract so
that the host code runs on the CPU."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err

S = np.load(os.path.join(GOLDEN, "ref_slice.npz"))


class TorchScan:
    """Contract of lncde_linear_scan_fwd/bwd: F_t = I + sum_l logsig[t, 1 + l] lie[..., l]; y_t = F_t y_{t-1}, t >= 1."""

    @staticmethod
    def apply(lie, logsig, y0, dims):
        nb, bs, n_basis = dims
        B, T, _ = logsig.shape
        flows = torch.einsum("nijl,btl->btnij", lie.reshape(nb, bs, bs, n_basis), logsig[..., 1 : 1 + n_basis])
        flows = flows + torch.eye(bs, dtype=lie.dtype)
        ys = [y0]
        for t in range(1, T):
            ys.append(torch.einsum("bnij,bnj->bni", flows[:, t], ys[-1].reshape(B, nb, bs)).reshape(B, nb * bs))
        return torch.stack(ys, 1)


def build(i, monkeypatch, device="cpu"):
    from lncde import _lib
    from lncde.models import LinearNeuralCDEs as M
    from lncde.models.generate_model import create_model

    if device == "cpu":
        monkeypatch.setattr(M, "_LinearScan", TorchScan)
        monkeypatch.setattr(M, "_diag_scan",  # the diagonal part of diagonal + dense: same contract with 1 x 1 blocks
                            lambda vf_At, ls, y0: TorchScan.apply(vf_At, ls, y0, (vf_At.shape[0], 1, vf_At.shape[1])))
        monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
        monkeypatch.setattr(_lib, "call_flags", lambda: _lib.FLAG_GENERIC)  # host logic only: the autograd formulation, no library kernels
    B, L, C, s, depth, H, bs, P, cls, ld = (int(v) for v in S[f"cfg{i}"])
    kind, kw = str(S[f"kind{i}"]), ast.literal_eval(str(S[f"kw{i}"]))
    name = {"wh": "wh", "dplr": "dplr", "sparse": "sparse", "diagdense": "diagonal_dense"}[kind] + "_linear_ncde"
    model, state = create_model(name, C, S[f"logsig{i}"].shape[-1], depth, None, ld, H, block_size=bs,
                                classification=bool(cls), lambd=0.1, parallel_steps=P, key=1, device=device, **kw)
    assert state is None
    vf = S[f"vf{i}"]
    if kind == "sparse":  # the reference's pattern: non-zeros of its dense [C, H, H] tensor
        idx = np.flatnonzero(vf)
        model._sparse_idx = torch.from_numpy(idx).to(device)
        assert len(idx) == model._blocks["_sparse_values"][1][0]
        vf_flat = vf[idx]
    else:
        vf_flat = vf
    flat = np.concatenate([vf_flat, S[f"init{i}"], S[f"out{i}"]])
    assert flat.size == model.params.numel()
    with torch.no_grad():
        model.params.copy_(torch.from_numpy(flat).float())
    ts = (np.float32(1.0 / L) * np.arange(L, dtype=np.float32)).astype(np.float32)
    X = (torch.from_numpy(np.tile(ts, (B, 1))), torch.from_numpy(S[f"logsig{i}"]).float().to(device),
         torch.from_numpy(S[f"x{i}"][:, 0, :].copy()).float().to(device))
    y = torch.from_numpy(S[f"y{i}"]).float().to(device)
    return model, X, y, bool(cls), kind


def check(i, model, X, y, cls, kind, tol, grad_tol=None):
    from lncde import train as T

    pred = model(X)
    assert rel_err(pred.detach().cpu().numpy(), S[f"pred{i}"]) < tol
    loss, grads = (T.classification_loss if cls else T.regression_loss)(model, X, y)
    want = float(S[f"loss{i}"])
    assert abs(float(loss) - want) < tol * max(1.0, abs(want))
    g = grads.detach().cpu().numpy()
    n_vf = model.n_A
    vfgrad = S[f"vfgrad{i}"]
    if kind == "sparse":
        vfgrad = vfgrad[model._sparse_idx.cpu().numpy()]  # gradient of the dense tensor at the pattern
    grad_tol = 5 * tol if grad_tol is None else grad_tol
    assert rel_err(g[:n_vf], vfgrad) < grad_tol
    assert rel_err(g[n_vf:], S[f"iograd{i}"]) < grad_tol


@pytest.mark.parametrize("i", range(int(S["n"])))
def test_slice_variant_host_logic_vs_reference_code(i, monkeypatch):
    model, X, y, cls, kind = build(i, monkeypatch)
    check(i, model, X, y, cls, kind, 2e-5)  # fp32 parameters against the float64 fixture


def test_slice_variant_attributes(monkeypatch):
    """Attributes train.py reads: vf_A is None for diagonal + dense (no Lip(2) term), vf_A_sparse.todense() for sparse."""
    model, *_ = build(5, monkeypatch)
    assert model.vf_A is None and model.vf_A_sparse is None and tuple(model.vf_A_diag.shape) == (3, 8)
    assert tuple(model.vf_A_dense.shape) == (3, 4, 4) and model.dense_size == 4
    model, *_ = build(4, monkeypatch)
    assert model.vf_A is None and tuple(model.vf_A_sparse.todense().shape) == (2, 6, 6)
    model, *_ = build(0, monkeypatch)
    assert tuple(model.vf_A.shape) == (3, 8) and tuple(model.hadamard_matrix.shape) == (8, 8)
    assert model.block_size == 8 and model.num_blocks == 1
    model, *_ = build(2, monkeypatch)
    assert tuple(model.vf_A_u.shape) == (3, 8, 2) and tuple(model.vf_A_v.shape) == (3, 8, 2)
