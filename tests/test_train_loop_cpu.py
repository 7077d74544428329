"""train_model (train.py:139-349 mirror): loop, metrics, early stopping and the files it writes, on the CPU with the
K2 scan replaced by the plain-torch statement of its contract (tests/test_slice_variants_cpu.TorchScan)."""

import numpy as np
import pytest
import torch


def _setup(monkeypatch, classification=True):
    from lncde import _lib, train as T
    from lncde.data_dir.datasets import dataset_generator
    from lncde.models import LinearNeuralCDEs as M
    from lncde.models.generate_model import create_model
    from oracle import paths
    from test_slice_variants_cpu import TorchScan

    monkeypatch.setattr(M, "_LinearScan", TorchScan)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "call_flags", lambda: _lib.FLAG_GENERIC)  # host logic only: the autograd formulation, no library kernels

    class CpuAdam:  # the fused K4 kernel needs a GPU: same update in torch for this host-logic test
        def __init__(self, model, lr, b1=0.9, b2=0.999, eps=1e-8):
            self.lr, self.b1, self.b2, self.eps, self.t = lr, b1, b2, eps, 0
            self.m, self.v = torch.zeros_like(model.params), torch.zeros_like(model.params)

        def update(self, model, g, grad_scale=1.0):
            self.t += 1
            g = g * grad_scale
            self.m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            mh, vh = self.m / (1 - self.b1 ** self.t), self.v / (1 - self.b2 ** self.t)
            with torch.no_grad():
                model.params.sub_(self.lr * mh / (vh.sqrt() + self.eps))

    monkeypatch.setattr(T, "Adam", CpuAdam)
    rng = np.random.default_rng(0)
    n, L, C = 80, 24, 2
    lab = rng.integers(0, 2, size=n)
    x = np.cumsum(rng.normal(scale=0.2, size=(n, L, C)), axis=1) + (lab[:, None, None] * np.linspace(0, 1.5, L)[None, :, None])
    y = np.eye(2, dtype=np.float32)[lab] if classification else rng.normal(size=(n, 7)).astype(np.float32)
    pf = lambda d, s, dep: torch.from_numpy(paths.batch_calc_paths(d.numpy().astype(np.float64), s, dep, np.float64)).float()
    ds = dataset_generator("toy", torch.from_numpy(x).float(), torch.from_numpy(y), 4, 2, False, 1.0, key=3, device="cpu",
                           paths_fn=pf)
    model, _ = create_model("bd_linear_ncde", ds.data_dim, ds.logsig_dim, 2, ds.intervals, ds.label_dim if classification else 1,
                            8, block_size=4, classification=classification, lambd=0.0, key=1, device="cpu")
    return T, ds, model


def test_train_model_classification(monkeypatch, tmp_path):
    T, ds, model = _setup(monkeypatch)
    out = str(tmp_path / "run")
    logs = []
    p0 = model.params.detach().clone()
    T.train_model("bd_linear_ncde", "toy", model, "accuracy", None, None, ds.path_dataloaders, 30, 5, 10, 1e-2,
                  lambda lr: lr, 8, 0, out, log=logs.append)
    assert not torch.equal(p0, model.params.detach())
    steps = np.load(out + "/steps.npy")
    tr, va = np.load(out + "/all_train_metric.npy"), np.load(out + "/all_val_metric.npy")
    assert list(steps) == [0, 5, 10, 15, 20, 25] and len(tr) == len(va) == 7 and tr[0] == va[0] == 0.0
    assert all(0.0 <= v <= 1.0 for v in va) and np.isfinite(np.load(out + "/test_metric.npy"))
    assert sum(l.startswith("Step:") for l in logs) == 6
    with pytest.raises(ValueError):
        T.train_model("bd_linear_ncde", "toy", model, "accuracy", None, None, ds.path_dataloaders, 5, 5, 10, 1e-2,
                      lambda lr: lr, 8, 0, out)                      # directory exists
    with pytest.raises(ValueError):
        T.train_model("bd_linear_ncde", "toy", model, "f1", None, None, ds.path_dataloaders, 5, 5, 10, 1e-2,
                      lambda lr: lr, 8, 0, str(tmp_path / "x"))       # unknown metric


def test_train_model_early_stopping_and_schedule(monkeypatch, tmp_path):
    T, ds, model = _setup(monkeypatch, classification=False)
    logs = []
    # lr = 0: the validation MSE never improves strictly; `>=` counts ties as no improvement -> stops early
    T.train_model("bd_linear_ncde", "toy", model, "mse", None, None, ds.path_dataloaders, 100, 2, 3, 0.0,
                  lambda lr: (lambda step: lr), 8, 0, str(tmp_path / "run"), log=logs.append)
    n_evals = sum(l.startswith("Step:") for l in logs)
    assert n_evals == 5   # 1st evaluation improves on the initial 100; ties at evaluations 2-5 count 1..4; 4 > 3 stops


def test_logsig_rescaling_quirk():
    from lncde.train import _scaled_inputs

    X = (torch.zeros(2, 3), torch.ones(2, 4, 5), torch.zeros(2, 2))
    assert float(_scaled_inputs("dplr_linear_ncde", "Heartbeat", X)[1][0, 0, 0]) == pytest.approx(0.01)
    assert float(_scaled_inputs("bd_linear_ncde", "Heartbeat", X)[1][0, 0, 0]) == pytest.approx(0.1)
    assert _scaled_inputs("bd_linear_ncde", "EigenWorms", X)[1] is X[1]
    assert _scaled_inputs("log_ncde", "Heartbeat", X)[1] is X[1]
