"""Host logic of the data pipeline against the reference's own dataset_generator /
Dataloader, executed from /root/reference under the library stand-ins (tests/golden/ref_data.npz).  The
log-signatures come from K1 in the product; here the oracle's batch_calc_paths is injected so that the host
logic (splits, ts, intervals, triples, batching protocol) runs on the CPU."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import paths

D = np.load(os.path.join(GOLDEN, "ref_data.npz"))


def _oracle_paths(d, stepsize, depth):
    return torch.from_numpy(paths.batch_calc_paths(d.cpu().numpy().astype(np.float64), stepsize, depth, np.float64)).float()


def build(i, device="cpu", paths_fn=_oracle_paths, inmemory=True):
    from lncde.data_dir.datasets import dataset_generator

    ntr, nva, nte, L, C, s, depth, include_time, onehot = (int(v) for v in D[f"cfg{i}"])
    data = tuple(torch.from_numpy(D[f"x{i}_{k}"]).float() for k in range(3))
    labels = tuple(torch.from_numpy(D[f"y{i}_{k}"]).float() for k in range(3))
    return dataset_generator(str(D[f"name{i}"]), data, labels, s, depth, bool(include_time), 1.0, inmemory=inmemory,
                             use_presplit=True, key=0, device=device, paths_fn=paths_fn)


def check(i, ds, tol=1e-5):
    assert [ds.data_dim, ds.logsig_dim, ds.label_dim] == [int(v) for v in D[f"dims{i}"]]
    assert ds.coeff_dataloaders is None
    assert rel_err(ds.intervals, D[f"intervals{i}"]) < 1e-6
    for k, split in enumerate(("train", "val", "test")):
        batches = list(ds.path_dataloaders[split].loop_epoch(4))
        assert [len(b[1]) for b in batches] == [int(v) for v in D[f"nb{i}_{k}"]]
        cat = lambda f: torch.cat([f(b) for b in batches]).cpu().numpy()
        assert rel_err(cat(lambda b: b[0][0]), D[f"ts{i}_{k}"]) < 1e-6
        assert rel_err(cat(lambda b: b[0][1]), D[f"ls{i}_{k}"]) < tol
        assert rel_err(cat(lambda b: b[0][2]), D[f"x0{i}_{k}"]) < 1e-6
        assert rel_err(cat(lambda b: b[1]), D[f"lab{i}_{k}"]) < 1e-6
        raw = list(ds.raw_dataloaders[split].loop_epoch(4))
        assert rel_err(torch.cat([b[0] for b in raw]).cpu().numpy(), D[f"x{i}_{k}"]) < 1e-6


@pytest.mark.parametrize("i", range(int(D["n"])))
def test_dataset_generator_vs_reference_code(i):
    check(i, build(i))


def test_loop_protocol_vs_reference_code(monkeypatch):
    """`loop` draws a new permutation once `end >= size`: the last batch of every permutation is never yielded."""
    from lncde.data_dir.dataloaders import Dataloader

    calls = [0]
    orig = torch.randperm

    def counting(*a, **k):
        calls[0] += 1
        return orig(*a, **k)

    monkeypatch.setattr(torch, "randperm", counting)
    for size, bs, K, n_perm, lo, hi in D["loop_protocol"]:
        dl = Dataloader(torch.randn(int(size), 4, 2), torch.randn(int(size)))
        calls[0] = 0
        it = dl.loop(int(bs), key=5)
        sizes, seen = [], []
        for _ in range(int(K)):
            x, y = next(it)
            sizes.append(len(y))
            seen.append(x)
        assert calls[0] == int(n_perm) and min(sizes) == int(lo) and max(sizes) == int(hi)


def test_dataloader_errors_and_modes():
    from lncde.data_dir.dataloaders import Dataloader

    dl = Dataloader(torch.arange(10.0)[:, None, None], torch.arange(10.0))
    with pytest.raises(ValueError):
        next(dl.loop(11, key=0))
    with pytest.raises(ValueError):
        next(dl.loop_epoch(0))
    with pytest.raises(ValueError):
        next(Dataloader(None, None).loop_epoch(1))
    with pytest.raises(RuntimeError):
        iter(dl)
    assert [len(y) for _, y in dl.loop_epoch(4)] == [4, 4, 2]
    assert [len(y) for _, y in dl.loop_epoch(5)] == [5, 5]     # the remainder batch may be a full one
    x, y = next(dl.loop(10, key=0))                             # batch == dataset: yielded as is
    assert len(y) == 10
    # a shuffled batch holds distinct samples and the (x, y) pairing survives the gather
    x, y = next(dl.loop(3, key=1))
    assert len(set(y.tolist())) == 3 and torch.equal(x[:, 0, 0], y)
    # host-resident mode: data stays on the host, batches are produced by func (here: CPU -> CPU)
    tri = (torch.zeros(6, 5), torch.randn(6, 3, 4), torch.randn(6, 2))
    dl = Dataloader(tri, torch.arange(6.0), inmemory=False, device="cpu")
    assert dl.data_is_logsig and not dl.data_is_coeffs and dl.size == 6
    (ts, ls, x0), y = next(dl.loop_epoch(4))
    assert ts.shape == (4, 5) and ls.shape == (4, 3, 4) and x0.shape == (4, 2)
