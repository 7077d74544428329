"""Run driver: create_dataset with the reference's signature over its processed-pickle layout
(data_dir/datasets.py:241-446), create_dataset_model_and_train (train.py:365-487) and run_experiments over the
reference's JSON configuration format (run_experiment.py:26-244) - host logic on the CPU, the K1 / K2 kernels
replaced by their oracle / plain-torch statements as in tests/test_train_loop_cpu.py."""
import json
import os
import pickle

import numpy as np
import pytest
import torch


def _paths_fn(d, s, dep):
    from oracle import paths

    return torch.from_numpy(paths.batch_calc_paths(d.numpy().astype(np.float64), s, dep, np.float64)).float()


def _write(path, obj):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(obj, f)


@pytest.fixture
def processed(tmp_path):
    rng = np.random.default_rng(0)
    root = str(tmp_path / "data_dir")
    x = np.cumsum(rng.normal(size=(40, 24, 3)), axis=1).astype(np.float32)
    _write(root + "/processed/UEA/Foo/data.pkl", x)
    _write(root + "/processed/UEA/Foo/labels.pkl", rng.integers(0, 3, size=40))
    for split, n in (("train", 20), ("val", 6), ("test", 6)):
        _write(root + f"/processed/UEA/Bar/X_{split}.pkl", rng.normal(size=(n, 24, 2)).astype(np.float32))
        _write(root + f"/processed/UEA/Bar/y_{split}.pkl", np.eye(2, dtype=np.float32)[rng.integers(0, 2, size=n)])
        _write(root + f"/processed/PPG/ppg/X_{split}.pkl", rng.uniform(-1, 1, size=(n, 64, 2)).astype(np.float32))
        _write(root + f"/processed/PPG/ppg/y_{split}.pkl", rng.uniform(-1, 1, size=(n, 4)).astype(np.float32))
    _write(root + "/processed/toy/signature/data.pkl", np.cumsum(np.round(rng.normal(size=(30, 20, 6))), axis=1).astype(np.float32))
    labs = [rng.normal(size=(30, 6)), rng.normal(size=(30, 6, 6)), rng.normal(size=(30, 6, 6, 6)), rng.normal(size=(30, 6, 6, 6, 6))]
    _write(root + "/processed/toy/signature/labels.pkl", labs)
    return root, x, labs


def test_create_dataset_reference_layout(processed):
    from lncde.data_dir.datasets import create_dataset

    root, x, labs = processed
    ds = create_dataset(root, "Foo", False, False, 4, 2, True, 1.0, scale=True, key=3, device="cpu", paths_fn=_paths_fn)
    assert ds.data_dim == 4 and ds.logsig_dim == 1 + 4 + 6 and ds.label_dim == 3
    X, y = next(iter(ds.path_dataloaders["train"].loop_epoch(28)))
    ts, ls, x0 = X
    assert ts.shape == (28, 24) and ls.shape == (28, 7, 11) and x0.shape == (28, 4) and y.shape == (28, 3)
    raw, _ = next(iter(ds.raw_dataloaders["train"].loop_epoch(28)))
    assert float(raw.min()) >= -1.0 - 1e-6 and float(raw.max()) <= 1.0 + 1e-6        # min-max scaling incl. the time channel
    # as in the reference, the scaling (datasets.py:300-314) also maps the time channel to [-1, 1] and the intervals
    # are read from that channel (:143-145, :161-163); LNCDE models, the only ones scaled, never read them
    t_scaled = 2 * (np.float32(1 / 24) * np.arange(24, dtype=np.float32)) / (np.float32(23 / 24) + 1e-8) - 1
    assert np.allclose(ds.intervals, np.concatenate([t_scaled[::4], [1.0]]), atol=1e-6)
    # pre-split UEA
    ds = create_dataset(root, "Bar", False, True, 5, 1, False, 1.0, key=3, device="cpu", paths_fn=_paths_fn)
    assert ds.data_dim == 2 and ds.logsig_dim == 3 and ds.label_dim == 2
    assert sum(1 for _ in ds.path_dataloaders["test"].loop_epoch(3)) == 2
    # PPG: host-resident loaders, regression labels, presplit or concatenated-and-resplit
    ds = create_dataset(root, "ppg", False, True, 8, 2, True, 1.0, key=3, device="cpu", paths_fn=_paths_fn)
    assert ds.label_dim == 1 and ds.data_dim == 3 and ds.path_dataloaders["train"].size == 20
    ds = create_dataset(root, "ppg", False, False, 8, 2, False, 1.0, key=3, device="cpu", paths_fn=_paths_fn)
    assert ds.data_dim == 2
    # toy: label = sign of one signature coordinate (datasets.py:326-333)
    ds = create_dataset(root, "signature2", False, False, 4, 2, False, 1.0, key=3, device="cpu", paths_fn=_paths_fn)
    assert ds.name == "toy" and ds.label_dim == 2 and ds.data_dim == 6
    n_pos = int((np.sign(labs[1][:, 2, 5]) > 0).sum())
    tot = sum(float(y[:, 1].sum()) for split in ("train", "val", "test") for _, y in ds.raw_dataloaders[split].loop_epoch(1))
    assert tot == n_pos
    with pytest.raises(ValueError, match="not found in UEA folder and not toy dataset"):
        create_dataset(root, "Nope", False, False, 4, 2, False, 1.0, key=0, device="cpu")


CFG = {"seeds": [2345, 3456], "data_dir": "data_dir", "output_parent_dir": "", "lr_scheduler": "lambda lr: lr",
       "num_steps": 12, "print_steps": 4, "early_stopping_steps": 10, "batch_size": 8, "model_name": "bd_linear_ncde",
       "metric": "accuracy", "classification": True, "dataset_name": "signature1", "use_presplit": False, "T": 1,
       "scale": 1, "lr": "0.001", "time": "False", "hidden_dim": "8", "block_size": "4", "depth": "2", "stepsize": "4.00",
       "lambd": "0.0"}


def test_parse_config_and_output_dir():
    from lncde.run_experiment import parse_config
    from lncde.train import experiment_output_dir

    run_args, seeds = parse_config("bd_linear_ncde", "signature1", CFG)
    assert seeds == [2345, 3456] and run_args["stepsize"] == 4 and run_args["logsig_depth"] == 2
    ma = run_args["model_args"]
    assert ma["block_size"] == 4 and ma["hidden_dim"] == 8 and ma["dt0"] is None and ma["vf_depth"] is None
    assert ma["parallel_steps"] == 1 and ma["rank"] == 0 and ma["lambd"] == 0.0 and run_args["lr_scheduler"](0.5) == 0.5
    out = experiment_output_dir("bd_linear_ncde", "signature1", 1, False, 12, 0.001, 4, 2, ma, 2345, "")
    assert out == ("outputs/bd_linear_ncde/signature1/T_1.00_time_False_nsteps_12_lr_0.001_block_size_4_hidden_dim_8"
                   "_solver_Heun_stepsize_controller_ConstantStepSize_scale_1_lambd_0.0_parallel_steps_1"
                   "_walsh_hadamard_False_diagonal_dense_False_sparsity_1.0_rank_0_seed_2345")
    nr = dict(CFG, dt0="0.05", vf_depth="3", vf_width="64", hidden_dim="64")
    run_args, _ = parse_config("nrde", "EigenWorms", nr)
    ma = run_args["model_args"]
    assert ma["vf_depth"] == 3 and ma["dt0"] == 0.05 and ma["lambd"] is None and ma["block_size"] is None
    assert "_stepsize_4.00_depth_2_" in experiment_output_dir("nrde", "EigenWorms", 1, False, 12, 0.001, 4, 2, ma, 1, "x/")
    with pytest.raises(NotImplementedError):
        parse_config("lru", "EigenWorms", CFG)


@pytest.mark.skipif(not os.path.isdir("/root/reference/experiment_configs/repeats"), reason="reference not mounted")
def test_parse_config_reads_every_reference_config():
    from lncde.run_experiment import LOG_ODE_MODELS, parse_config

    root, n = "/root/reference/experiment_configs/repeats", 0
    for model in LOG_ODE_MODELS:
        for f in sorted(os.listdir(os.path.join(root, model))):
            with open(os.path.join(root, model, f)) as fh:
                run_args, seeds = parse_config(model, f[:-5], json.load(fh))
            assert len(seeds) >= 1 and run_args["stepsize"] >= 1 and run_args["model_args"]["hidden_dim"] > 0
            n += 1
    assert n >= 9 * 6


def test_run_experiments_end_to_end(monkeypatch, tmp_path):
    """JSON config -> synthetic dataset -> model -> training loop -> the reference's output files, one run per seed."""
    from lncde import _lib, run_experiment as R, train as T
    from lncde.models import LinearNeuralCDEs as M
    from test_slice_variants_cpu import TorchScan

    monkeypatch.setattr(M, "_LinearScan", TorchScan)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "call_flags", lambda: _lib.FLAG_GENERIC)  # host logic only: the autograd formulation, no library kernels
    monkeypatch.setattr(_lib, "require_cuda_device", lambda *a: None)

    class CpuAdam:
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
    folder = tmp_path / "configs" / "bd_linear_ncde"
    folder.mkdir(parents=True)
    cfg = dict(CFG, output_parent_dir=str(tmp_path) + "/")
    (folder / "signature1.json").write_text(json.dumps(cfg))
    logs = []
    models = R.run_experiments(["bd_linear_ncde"], ["signature1"], str(tmp_path / "configs"), data_dir="synthetic:64",
                               device="cpu", paths_fn=_paths_fn, log=logs.append)
    assert len(models) == 2
    base = tmp_path / "outputs" / "bd_linear_ncde" / "signature1"
    runs = sorted(os.listdir(base))
    assert len(runs) == 2 and runs[0].endswith("_seed_2345") and runs[1].endswith("_seed_3456")
    for r in runs:
        assert list(np.load(base / r / "steps.npy")) == [0, 4, 8]
        assert np.isfinite(np.load(base / r / "all_val_metric.npy")).all()
    assert sum(l.startswith("Step:") for l in logs) == 6
    with pytest.raises(ValueError, match="already exists"):
        R.run_experiments(["bd_linear_ncde"], ["signature1"], str(tmp_path / "configs"), data_dir="synthetic:64",
                          device="cpu", paths_fn=_paths_fn, log=logs.append, seeds=[2345])
