"""``run_experiments`` - mirror of run_experiment.py:26-244 for the models on the Log-ODE input: reads the
reference's own JSON configuration files (``experiment_configs/repeats/<model>/<dataset>.json``, same keys and the
same string-typed values) and calls ``create_dataset_model_and_train`` once per seed.

    python -m lncde.run_experiment --experiment_folder /path/to/experiment_configs/repeats \\
           --models log_ncde bd_linear_ncde --datasets EigenWorms [--data_dir synthetic:256]
"""
from __future__ import annotations

import argparse
import json

LOG_ODE_MODELS = ("log_ncde", "nrde", "bd_linear_ncde", "diagonal_linear_ncde", "diagonal_dense_linear_ncde",
                  "dense_linear_ncde", "wh_linear_ncde", "sparse_linear_ncde", "dplr_linear_ncde")


def _scheduler(src):
    """The reference's configs hold the schedule as the source of a lambda (``"lambda lr: lr"``, evaluated at
    run_experiment.py:43); here it is evaluated without builtins, with ``math`` as the only name in scope."""
    import math

    fn = eval(src, {"__builtins__": {}, "math": math}, {})
    if not callable(fn):
        raise ValueError(f"lr_scheduler must be the source of a callable, got {src!r}")
    return fn


def parse_config(model_name, dataset_name, data):
    """(run_args, seeds) from one reference config dict (run_experiment.py:40-150, 199-240)."""
    if model_name not in LOG_ODE_MODELS:
        raise NotImplementedError(f"{model_name} does not take the Log-ODE input (SURVEY.md section 8)")
    linear = model_name.endswith("_linear_ncde")
    dt0 = None if linear else float(data["dt0"])
    model_args = {
        "num_blocks": None,
        "block_size": int(data["block_size"]) if linear else None,
        "hidden_dim": int(data["hidden_dim"]),
        "vf_depth": None if linear else int(data["vf_depth"]),
        "vf_width": None if linear else int(data["vf_width"]),
        "ssm_dim": None,
        "ssm_blocks": None,
        "dt0": dt0,
        "solver": "Heun",
        "stepsize_controller": "ConstantStepSize",
        "scale": data["scale"],
        "lambd": float(data["lambd"]) if (linear or model_name == "log_ncde") else None,
        "parallel_steps": int(data.get("parallel_steps", 1)) if linear else None,
        "walsh_hadamard": data.get("walsh_hadamard", False) if linear else None,
        "diagonal_dense": data.get("diagonal_dense", False) if linear else None,
        "sparsity": data.get("sparsity", 1.0) if linear else None,
        "rank": int(data.get("rank", 0)) if linear else None,
    }
    run_args = {
        "data_dir": data["data_dir"],
        "use_presplit": data["use_presplit"],
        "dataset_name": dataset_name,
        "output_step": int(data["output_step"]) if dataset_name == "ppg" else 1,
        "metric": data["metric"],
        "include_time": data["time"].lower() == "true",
        "T": data["T"],
        "model_name": model_name,
        "stepsize": int(float(data["stepsize"])),
        "logsig_depth": int(data["depth"]),
        "model_args": model_args,
        "num_steps": data["num_steps"],
        "print_steps": data["print_steps"],
        "early_stopping_steps": data["early_stopping_steps"],
        "lr": float(data["lr"]),
        "lr_scheduler": _scheduler(data["lr_scheduler"]),
        "batch_size": data["batch_size"],
        "output_parent_dir": data["output_parent_dir"],
    }
    return run_args, list(data["seeds"])


def run_experiments(model_names, dataset_names, experiment_folder, pytorch_experiments=False, *, data_dir=None,
                    num_steps=None, seeds=None, **train_kwargs):
    if pytorch_experiments:
        raise NotImplementedError("the torch_experiments models (mamba, S6) are outside the Log-ODE hot path")
    from .train import create_dataset_model_and_train

    out = []
    for model_name in model_names:
        for dataset_name in dataset_names:
            with open(experiment_folder + f"/{model_name}/{dataset_name}.json", "r") as fh:
                run_args, cfg_seeds = parse_config(model_name, dataset_name, json.load(fh))
            if data_dir is not None:
                run_args["data_dir"] = data_dir
            if num_steps is not None:
                run_args["num_steps"] = num_steps
            for seed in (seeds or cfg_seeds):
                print(f"Running experiment with seed: {seed}")
                out.append(create_dataset_model_and_train(seed=seed, **run_args, **train_kwargs))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--experiment_folder", default="experiment_configs/repeats")
    ap.add_argument("--models", nargs="+", default=list(LOG_ODE_MODELS))
    ap.add_argument("--datasets", nargs="+", default=["EigenWorms", "EthanolConcentration", "Heartbeat", "MotorImagery",
                                                      "SelfRegulationSCP1", "SelfRegulationSCP2"])
    ap.add_argument("--data_dir", default=None, help='override the config, e.g. "synthetic:256" when the datasets are absent')
    ap.add_argument("--num_steps", type=int, default=None)
    a = ap.parse_args()
    run_experiments(a.models, a.datasets, a.experiment_folder, data_dir=a.data_dir, num_steps=a.num_steps)
