"""Oracle: windowed Hall-basis log-signatures (``calc_paths``).  TEST INFRASTRUCTURE ONLY.

Restates ``data_dir/generate_paths.py:14-77`` and ``data_dir/datasets.py:43-71,
143-163`` of the reference, including its quirks:

* B1 – the remainder window's basepoint is the last point of the PREVIOUS
  SAMPLE of the same chunk (generate_paths.py:40-48 applied un-vmapped at :59).
* B3 – the time axis is zero-prepended first (generate_paths.py:30-32), so
  window k covers original points k*s-1 .. k*s+s-2.

PINNED to the reference's own calc_paths / batch_calc_paths, executed from /root/reference under
library stand-ins (tests/golden/ref_paths.npz, tests/test_oracle_reference_pinned.py); the
signax semantics underneath (oracle.tensor_algebra) are restated, see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np

from . import tensor_algebra as ta
from .hall import Hall


def hall_basis_logsig(x: np.ndarray, depth: int, t2l) -> np.ndarray:
    """generate_paths.py:14-19.  x: batch+(n_points, C) -> batch+(D,).

    depth >= 3 extends the reference in the obvious way."""
    flat = ta.flatten(ta.log(ta.signature(x, depth)))
    if depth == 1:
        return np.concatenate([np.zeros(flat.shape[:-1] + (1,), flat.dtype), flat], axis=-1)
    return flat @ t2l[:, 1:].T.astype(flat.dtype)


def calc_paths(data: np.ndarray, stepsize: int, depth: int, dtype=np.float32, compat: bool = True) -> np.ndarray:
    """generate_paths.py:22-77.  data[B,L,C] -> logsigs[B,K,D].

    ``compat=True`` reproduces quirk B1; ``compat=False`` uses each sample's own
    predecessor point as the basepoint of the remainder window."""
    data = np.asarray(data, dtype=dtype)
    B, L, C = data.shape
    data = np.concatenate([np.zeros((B, 1, C), dtype), data], axis=1)  # :30-32
    n = L + 1
    t2l = Hall(C, depth).t2l_matrix(depth) if depth >= 2 else None  # :34-38
    stepsize = min(int(stepsize), n)  # :50-51
    r = n % stepsize
    q = n // stepsize
    full = data[:, : q * stepsize].reshape(B, q, stepsize, C)
    # basepoint of window k = last point of window k-1 (zeros for k = 0)   :40-48
    base = np.concatenate([np.zeros((B, 1, C), dtype), full[:, :-1, -1, :]], axis=1)
    win = np.concatenate([base[:, :, None, :], full], axis=2)  # [B,q,s+1,C]
    out = hall_basis_logsig(win, depth, t2l)
    if r != 0:
        final = data[:, -r - 1 :]  # [B, r+1, C], already holds its true predecessor  :54
        if compat:
            # :59 – prepend() un-vmapped: batch axis plays the window axis
            fbase = np.concatenate([np.zeros((1, C), dtype), final[:-1, -1, :]], axis=0)
            final = np.concatenate([fbase[:, None, :], final], axis=1)
        out = np.concatenate([out, hall_basis_logsig(final, depth, t2l)[:, None, :]], axis=1)
    return out


def batch_calc_paths(data: np.ndarray, stepsize: int, depth: int, dtype=np.float32, compat: bool = True):
    """datasets.py:43-71 – chunks of 128 samples (quirk B1 is per chunk)."""
    N = len(data)
    outs = [calc_paths(data[i : i + 128], stepsize, depth, dtype, compat) for i in range(0, N - N % 128, 128)]
    if N % 128:
        outs.append(calc_paths(data[-(N % 128) :], stepsize, depth, dtype, compat))
    return np.concatenate(outs)


def make_ts(L: int, T: float = 1.0) -> np.ndarray:
    """datasets.py:147-150 (include_time False): fp32 (T/L)*arange(L)."""
    return (np.float32(T / L) * np.arange(L, dtype=np.float32)).astype(np.float32)


def make_intervals(ts0: np.ndarray, stepsize: int, T: float = 1.0) -> np.ndarray:
    """datasets.py:161-163: ts[0, 0:L:stepsize] ++ [T]."""
    L = ts0.shape[0]
    idx = np.unique(np.r_[0 : L : int(stepsize)])
    return np.concatenate([ts0[idx], np.array([T], np.float32)]).astype(np.float32)


def levy_logsig_depth2(win: np.ndarray) -> np.ndarray:
    """Independent closed form used as an identity check: level 1 = end - start,
    Hall level 2 (i<j, lexicographic) = Levy area.  win: batch+(n_points, C)."""
    inc = np.diff(win, axis=-2)
    a = win[..., :-1, :] - win[..., :1, :]
    C = win.shape[-1]
    lvl1 = win[..., -1, :] - win[..., 0, :]
    cols = [np.zeros_like(lvl1[..., :1]), lvl1]
    area = 0.5 * (np.einsum("...ni,...nj->...ij", a, inc) - np.einsum("...ni,...nj->...ji", a, inc))
    iu = np.triu_indices(C, 1)
    cols.append(area[..., iu[0], iu[1]])
    return np.concatenate(cols, axis=-1)
