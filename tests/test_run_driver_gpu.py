"""Run driver on the GPU (the CPU twin with stubbed kernels is tests/test_run_driver_cpu.py):
run_experiments over the reference's JSON configuration format -> synthetic dataset -> device data pipeline (K1) ->
model -> train_model with the kernel-only train step (K2 / K3 / K5) and the fused Adam kernel (K4) -> the output files the
reference writes (train.py:139-349, run_experiment.py:26-244)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BASE = {"seeds": [2345], "data_dir": "data_dir", "output_parent_dir": "", "lr_scheduler": "lambda lr: lr", "num_steps": 12,
        "print_steps": 4, "early_stopping_steps": 10, "batch_size": 8, "metric": "accuracy", "classification": True,
        "dataset_name": "signature1", "use_presplit": False, "T": 1, "scale": 1, "lr": "0.001", "time": "False",
        "depth": "2", "stepsize": "4.00"}
CASES = {
    "bd_linear_ncde": dict(hidden_dim="8", block_size="4", lambd="0.001"),
    "diagonal_linear_ncde": dict(hidden_dim="16", block_size="1", lambd="0.0"),
    "log_ncde": dict(hidden_dim="16", vf_depth="2", vf_width="8", dt0="0.05", lambd="0.001"),
    "nrde": dict(hidden_dim="16", vf_depth="2", vf_width="8", dt0="0.05"),
}


@pytest.mark.parametrize("model_name", sorted(CASES))
def test_run_experiments_on_the_device(built_lib, tmp_path, model_name):
    from lncde import run_experiment as R

    folder = tmp_path / "configs" / model_name
    folder.mkdir(parents=True)
    cfg = dict(BASE, model_name=model_name, output_parent_dir=str(tmp_path) + "/", **CASES[model_name])
    (folder / "signature1.json").write_text(json.dumps(cfg))
    logs = []
    models = R.run_experiments([model_name], ["signature1"], str(tmp_path / "configs"), data_dir="synthetic:64", device="cuda",
                               log=logs.append)
    assert len(models) == 1 and torch.isfinite(models[0].params).all()
    base = tmp_path / "outputs" / model_name / "signature1"
    (run,) = os.listdir(base)
    assert run.endswith("_seed_2345")
    assert list(np.load(base / run / "steps.npy")) == [0, 4, 8]
    va = np.load(base / run / "all_val_metric.npy")
    assert np.isfinite(va).all() and all(0.0 <= v <= 1.0 for v in va) and np.isfinite(np.load(base / run / "test_metric.npy"))
    assert sum(l.startswith("Step:") for l in logs) == 3
