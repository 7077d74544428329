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

from .dataloaders import Dataloader
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
    """data_dir/datasets.py:31-40."""
    name: str
    raw_dataloaders: dict
    coeff_dataloaders: dict | None   # NCDE interpolation coefficients: outside the Log-ODE hot path (always None)
    path_dataloaders: dict
    data_dim: int
    logsig_dim: int
    intervals: np.ndarray
    label_dim: int


def dataset_generator(name, data, labels, stepsize, depth, include_time, T, inmemory=True, idxs=None,
                      use_presplit=False, *, key, device="cuda", paths_fn=None):
    """Mirror of data_dir/datasets.py:105-232 for the raw and path (log-signature) dataloaders: 70 / 15 / 15
    split by a random permutation (or ``idxs`` / pre-split tuples), ``ts`` (:143-156), log-signatures of every
    split through K1 - ONE launch per split instead of the reference's loop over chunks of 128, same chunk
    semantics (``batch_calc_paths``) - ``intervals`` (:161-163) and the ``(ts, logsig, x0)`` triples (:185-200).
    Everything stays on ``device``; with ``inmemory=False`` the loaders keep pinned host copies instead.
    As in the reference, ``idxs`` leaves the test split empty (:138-141)."""
    paths_fn = paths_fn or batch_calc_paths
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
    if idxs is None:
        if use_presplit:
            (train_data, val_data, test_data), (train_labels, val_labels, test_labels) = data, labels
        else:
            N = len(data)
            gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
            perm = torch.randperm(N, generator=gen)
            b1, b2 = int(N * 0.7), int(N * 0.85)
            data, labels = as_t(data), as_t(labels)
            pick = lambda a, i: a.index_select(0, i.to(a.device))
            train_data, train_labels = pick(data, perm[:b1]), pick(labels, perm[:b1])
            val_data, val_labels = pick(data, perm[b1:b2]), pick(labels, perm[b1:b2])
            test_data, test_labels = pick(data, perm[b2:]), pick(labels, perm[b2:])
    else:
        data, labels = as_t(data), as_t(labels)
        i0, i1 = (torch.as_tensor(np.asarray(i)).long() for i in idxs[:2])
        train_data, train_labels = data[i0.to(data.device)], labels[i0.to(labels.device)]
        val_data, val_labels = data[i1.to(data.device)], labels[i1.to(labels.device)]
        test_data, test_labels = None, None
    dev = torch.device(device)
    splits = {}
    for split, d, l in (("train", train_data, train_labels), ("val", val_data, val_labels),
                        ("test", test_data, test_labels)):
        if d is None:
            splits[split] = None
            continue
        d = as_t(d).to(dev, torch.float32).contiguous()
        l = as_t(l).to(dev)
        if include_time:
            ts = d[:, :, 0]
        else:
            ts = torch.from_numpy(make_ts(d.shape[1], T)).to(dev)[None, :].expand(d.shape[0], -1)
        splits[split] = (d, l, ts, paths_fn(d, stepsize, depth))
    tr = splits["train"]
    idx = np.unique(np.r_[0 : tr[0].shape[1] : int(stepsize)])
    intervals = np.concatenate([tr[2][0].detach().cpu().numpy()[idx], np.array([T], np.float32)]).astype(np.float32)
    keep = (lambda t: t) if inmemory else (lambda t: t.cpu().pin_memory() if dev.type == "cuda" else t.cpu())
    raw, path = {}, {}
    for split, v in splits.items():
        if v is None:
            raw[split], path[split] = Dataloader(None, None, inmemory, device), Dataloader(None, None, inmemory, device)
            continue
        d, l, ts, p = v
        raw[split] = Dataloader(keep(d), keep(l), inmemory, device)
        path[split] = Dataloader((keep(ts.contiguous()), keep(p), keep(d[:, 0, :].contiguous())), keep(l), inmemory, device)
    label_dim = 1 if (tr[1].dim() == 1 or name == "ppg") else int(tr[1].shape[-1])
    return Dataset(name, raw, None, path, int(tr[0].shape[-1]), int(tr[3].shape[-1]), intervals, label_dim)


def create_synthetic_dataset(name: str, n: int, *, seed: int = 2345, stepsize=None, depth=None, T: float = 1.0,
                             scale: bool = False, device="cuda", L: int | None = None, n_out: int | None = None,
                             inmemory: bool = True, paths_fn=None) -> Dataset:
    """Synthetic stand-in for data_dir/datasets.py:create_dataset (the real datasets are not available offline):
    synthetic paths of the named dataset's shape, then ``dataset_generator`` exactly as the reference calls it."""
    cfg = SHAPES[name]
    stepsize = cfg["stepsize"] if stepsize is None else int(stepsize)
    depth = cfg["depth"] if depth is None else int(depth)
    x = synthetic_paths(name, n, seed, T, scale_to_unit=scale, L=L)
    rng = np.random.default_rng(seed + 1)
    if cfg["classification"]:
        lab = np.eye(cfg["label_dim"], dtype=np.float32)[rng.integers(0, cfg["label_dim"], size=n)]
    else:
        lab = rng.uniform(-1, 1, size=(n, n_out or 1)).astype(np.float32)
    return dataset_generator(name, torch.from_numpy(x), torch.from_numpy(lab), stepsize, depth, cfg["time"], T,
                             inmemory=inmemory, key=seed, device=device, paths_fn=paths_fn)


def _load(path):
    """One of the reference's processed pickles (data_dir/process_uea.py, process_ppg.py, toy_dataset.py write jax
    arrays; anything ``np.asarray`` understands is accepted - unpickling a jax array needs jax importable)."""
    import pickle

    with open(path, "rb") as f:
        return np.asarray(pickle.load(f))


def _with_time(x, T):
    t = (np.float32(T / x.shape[1]) * np.arange(x.shape[1], dtype=np.float32))[None, :, None]
    return np.concatenate([np.broadcast_to(t, (x.shape[0], x.shape[1], 1)), x.astype(np.float32)], axis=2)


def _onehot(labels):
    labels = np.asarray(labels).astype(np.int64)
    return np.eye(len(np.unique(labels)), dtype=np.float32)[labels]


def _minmax(x, lo, hi, eps=1e-8):  # datasets.py:236-238
    return 2.0 * (x - lo) / (hi - lo + eps) - 1.0


def create_dataset(data_dir, name, use_idxs, use_presplit, stepsize, depth, include_time, T, scale=False, *, key,
                   device="cuda", paths_fn=None):
    """data_dir/datasets.py:406-446 with the reference's signature: the processed UEA / toy / PPG pickles under
    ``data_dir/processed`` (:241-403), optional time channel, LNCDE min-max scaling (:300-314), then
    ``dataset_generator``.  ``data_dir = "synthetic"`` or ``"synthetic:<n>"`` builds the named dataset's shape from
    random walks instead (the real datasets are not available offline)."""
    import os

    if str(data_dir).startswith("synthetic"):
        n = int(str(data_dir).split(":")[1]) if ":" in str(data_dir) else 256
        base = "toy" if name.startswith("signature") else name
        if base not in SHAPES:
            raise ValueError(f"Dataset {name} not found in UEA folder and not toy dataset")
        return create_synthetic_dataset(base, n, seed=int(key) if not isinstance(key, torch.Generator) else 2345,
                                        stepsize=stepsize, depth=depth, T=T, scale=scale, device=device,
                                        inmemory=base != "ppg", paths_fn=paths_fn)
    proc = os.path.join(data_dir, "processed")
    sub = lambda d: [f.name for f in os.scandir(os.path.join(proc, d)) if f.is_dir()] if os.path.isdir(os.path.join(proc, d)) else []
    gen = dict(key=key, device=device, paths_fn=paths_fn)
    if name in sub("UEA"):
        d = os.path.join(proc, "UEA", name)
        idxs = None
        if use_presplit:
            data = [_load(os.path.join(d, f"X_{s}.pkl")) for s in ("train", "val", "test")]
            labels = tuple(_load(os.path.join(d, f"y_{s}.pkl")) for s in ("train", "val", "test"))
            if include_time:
                data = [_with_time(x, T) for x in data]
            if scale:
                allx = np.concatenate(data, axis=0)
                lo, hi = allx.min(axis=(0, 1), keepdims=True), allx.max(axis=(0, 1), keepdims=True)
                data = [_minmax(x, lo, hi) for x in data]
            data = tuple(data)
        else:
            data = _load(os.path.join(d, "data.pkl"))
            labels = _onehot(_load(os.path.join(d, "labels.pkl")))
            if use_idxs:
                import pickle

                with open(os.path.join(d, "original_idxs.pkl"), "rb") as f:
                    idxs = pickle.load(f)
            if include_time:
                data = _with_time(data, T)
            if scale:
                data = _minmax(data, data.min(axis=(0, 1), keepdims=True), data.max(axis=(0, 1), keepdims=True))
        return dataset_generator(name, data, labels, stepsize, depth, include_time, T, idxs=idxs,
                                 use_presplit=use_presplit, **gen)
    if name[:-1] in sub("toy"):
        import pickle

        data = _load(os.path.join(proc, "toy", "signature", "data.pkl"))
        with open(os.path.join(proc, "toy", "signature", "labels.pkl"), "rb") as f:
            lab = pickle.load(f)
        pick = {"signature1": lambda l: l[0][:, 2], "signature2": lambda l: l[1][:, 2, 5],
                "signature3": lambda l: l[2][:, 2, 5, 0], "signature4": lambda l: l[3][:, 2, 5, 0, 3]}
        if name not in pick:
            raise ValueError(f"Dataset {name} not found in UEA folder and not toy dataset")
        labels = ((np.sign(np.asarray(pick[name](lab))) + 1) / 2).astype(int)  # datasets.py:326-333
        if include_time:
            data = _with_time(data, T)
        return dataset_generator("toy", data, _onehot(labels), stepsize, depth, include_time, T, idxs=None, **gen)
    if name == "ppg":
        d = os.path.join(proc, "PPG", "ppg")
        data = [_load(os.path.join(d, f"X_{s}.pkl")) for s in ("train", "val", "test")]
        labels = [_load(os.path.join(d, f"y_{s}.pkl")) for s in ("train", "val", "test")]
        if include_time:
            data = [_with_time(x, T) for x in data]
        if use_presplit:
            data, labels = tuple(data), tuple(labels)
        else:
            data, labels = np.concatenate(data, axis=0), np.concatenate(labels, axis=0)
        return dataset_generator("ppg", data, labels, stepsize, depth, include_time, T, inmemory=False,
                                 use_presplit=use_presplit, **gen)
    raise ValueError(f"Dataset {name} not found in UEA folder and not toy dataset")
