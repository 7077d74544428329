"""``Dataloader`` - mirror of data_dir/dataloaders.py:35-160 over device-resident tensors.

Same three data layouts (raw ``[N, L, C]``; NCDE coefficient tuples; ``(ts, logsig, x0)`` triples for
Log-NCDE / LNCDE), same batching rules - including the reference's quirk that ``loop`` never yields the LAST
batch of a permutation, even a full one (``while end < self.size``, :88), while ``loop_epoch`` ends with the
remainder batch (:140-160).  Batches are gathered on the device the data lives on (one ``index_select`` per
tensor); with ``inmemory=False`` the data stays in (pinned) host memory and each batch is copied to ``device``,
the reference's host-resident mode for PPG (:57-62).  ``key`` seeds a ``torch.Generator`` (jax.random is not
available; the permutation values differ from the reference's, the protocol does not)."""
from __future__ import annotations

import torch


def _take(x, idx):
    return x.index_select(0, idx.to(x.device))


class Dataloader:
    data_is_coeffs = False
    data_is_logsig = False

    def __init__(self, data, labels, inmemory=True, device="cuda"):
        self.data, self.labels = data, labels
        if isinstance(data, tuple):
            if isinstance(data[1], (tuple, list)):  # (:46-47): a tuple of coefficient arrays
                self.data_is_coeffs = True
            else:
                self.data_is_logsig = True
        if self.data_is_coeffs:
            self.size = len(data[1][0])
        elif self.data_is_logsig:
            self.size = len(data[1])
        elif data is None:
            self.size = 0
        else:
            self.size = len(data)
        if inmemory:
            self.func = lambda x: x
        else:
            dev = torch.device(device)
            self.func = lambda x: x.to(dev, non_blocking=True)

    def __iter__(self):
        raise RuntimeError("Use .loop(batch_size) instead of __iter__")

    def _check(self, batch_size):
        if self.size == 0:
            raise ValueError("This dataloader is empty")
        if not (isinstance(batch_size, int) and batch_size > 0):
            raise ValueError("Batch size must be a positive integer")
        if batch_size > self.size:
            raise ValueError("Batch size larger than dataset size")

    def _batch(self, idx):
        f = self.func
        if self.data_is_coeffs:
            return (f(_take(self.data[0], idx)), tuple(f(_take(d, idx)) for d in self.data[1]),
                    f(_take(self.data[2], idx))), f(_take(self.labels, idx))
        if self.data_is_logsig:
            return (f(_take(self.data[0], idx)), f(_take(self.data[1], idx)), f(_take(self.data[2], idx))), \
                f(_take(self.labels, idx))
        return f(_take(self.data, idx)), f(_take(self.labels, idx))

    def loop(self, batch_size, *, key):
        """Infinite stream of shuffled batches (:67-108)."""
        self._check(batch_size)
        if batch_size == self.size:
            while True:
                yield self._map(self.data), self.func(self.labels)
        gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
        while True:
            perm = torch.randperm(self.size, generator=gen)
            start, end = 0, batch_size
            while end < self.size:
                yield self._batch(perm[start:end])
                start, end = end, end + batch_size

    def loop_epoch(self, batch_size):
        """One ordered pass; the last batch holds the remainder (:110-160)."""
        self._check(batch_size)
        if batch_size == self.size:
            yield self.data, self.labels
            return
        indices = torch.arange(self.size)
        start, end = 0, batch_size
        while end < self.size:
            yield self._batch(indices[start:end])
            start, end = end, end + batch_size
        yield self._batch(indices[start:])

    def _map(self, data):
        if isinstance(data, tuple):
            return tuple(self._map(d) for d in data)
        return self.func(data)
