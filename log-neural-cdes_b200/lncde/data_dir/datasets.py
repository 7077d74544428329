"""Synthetic stand-ins for the reference datasets (no datasets offline) + the hot-path
parts of data_dir/datasets.py: ``ts`` (:143-156), ``intervals`` (:161-163), the
``(ts, logsig, x0)`` triples (:185-200) and ``logsig_dim`` (:207).

Shapes and value distributions follow SURVEY.md §8d (mined from the reference's
processing scripts); generator = numpy default_rng(seed).
"""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

from .generate_paths import batch_calc_paths

# name: (L, raw channels, time channel, stepsize, depth, n_classes / None for regression)
SHAPES = {
    "toy": dict(L=100, C=6, time=False, stepsize=4, depth=2, label_dim=2, classification=True),
    "EigenWorms": dict(L=17984, C=6, time=True, stepsize=12, depth=2, label_dim=5, classification=True),
    "SelfRegulationSCP1": dict(L=896, C=6, time=False, stepsize=16, depth=2, label_dim=2, classification=True),
    "ppg": dict(L=49920, C=6, time=True, stepsize=10, depth=1, label_dim=1, classification=False),
}


def synthetic_paths(name: str, n: int, seed: int = 2345, T: float = 1.0, scale_to_unit: bool = False,
                    L: int | None = None) -> np.ndarray:
    """[n, L, C(+1)] fp32.  toy: cumsum(round(N(0,1))) (toy_dataset.py:17-19); others: a
    random walk cumsum(N(0, 0.05^2)); channel 0 = t_i = (T/L) i when the config has time
    (datasets.py:270-277); optional per-channel min-max scaling to [-1,1] (LNCDE, :237-239)."""
    cfg = SHAPES[name]
    L = cfg["L"] if L is None else L
    rng = np.random.default_rng(seed)
    if name == "toy":
        x = np.cumsum(np.round(rng.normal(size=(n, L, cfg["C"]))), axis=1)
    else:
        x = np.cumsum(rng.normal(scale=0.05, size=(n, L, cfg["C"])), axis=1)
    if cfg["time"]:
        t = (T / L) * np.arange(L)
        x = np.concatenate([np.broadcast_to(t[None, :, None], (n, L, 1)), x], axis=2)
    if scale_to_unit:
        lo, hi = x.min(axis=(0, 1), keepdims=True), x.max(axis=(0, 1), keepdims=True)
        x = 2 * (x - lo) / (hi - lo) - 1
    return np.ascontiguousarray(x, dtype=np.float32)


def make_ts(L: int, T: float = 1.0) -> np.ndarray:
    return (np.float32(T / L) * np.arange(L, dtype=np.float32)).astype(np.float32)


def make_intervals(ts0: np.ndarray, stepsize: int, T: float = 1.0) -> np.ndarray:
    idx = np.unique(np.r_[0 : ts0.shape[0] : int(stepsize)])
    return np.concatenate([ts0[idx], np.array([T], np.float32)]).astype(np.float32)


@dataclasses.dataclass
class Dataset:
    name: str
    data: torch.Tensor        # [N, L, C] on device
    labels: torch.Tensor      # one-hot [N, n_classes] or [N, n_out]
    ts: torch.Tensor          # [L]
    paths: torch.Tensor       # logsig [N, K, D]
    data_dim: int
    logsig_dim: int
    intervals: np.ndarray
    label_dim: int

    def batch(self, idx):
        """(X, y) with X = (ts, logsig, x0) as in datasets.py:185-200."""
        ts = self.ts[None, :].expand(len(idx), -1)
        return (ts, self.paths[idx], self.data[idx, 0, :]), self.labels[idx]


def create_dataset(name: str, n: int, *, seed: int = 2345, stepsize=None, depth=None, T: float = 1.0,
                   scale: bool = False, device="cuda", L: int | None = None, n_out: int | None = None) -> Dataset:
    cfg = SHAPES[name]
    stepsize = cfg["stepsize"] if stepsize is None else int(stepsize)
    depth = cfg["depth"] if depth is None else int(depth)
    x = synthetic_paths(name, n, seed, T, scale_to_unit=scale, L=L)
    Ln = x.shape[1]
    rng = np.random.default_rng(seed + 1)
    if cfg["classification"]:
        lab = np.eye(cfg["label_dim"], dtype=np.float32)[rng.integers(0, cfg["label_dim"], size=n)]
    else:
        lab = rng.uniform(-1, 1, size=(n, n_out or 1)).astype(np.float32)
    data = torch.from_numpy(x).to(device)
    paths = batch_calc_paths(data, stepsize, depth)
    ts0 = make_ts(Ln, T)
    return Dataset(name, data, torch.from_numpy(lab).to(device), torch.from_numpy(ts0).to(device), paths,
                   x.shape[2], paths.shape[-1], make_intervals(ts0, stepsize, T), cfg["label_dim"])
