"""``calc_paths`` – mirror of data_dir/generate_paths.py:22-77 over kernel K1."""
from __future__ import annotations

import functools

import torch

from .. import _lib
from .hall_set import HallSet


@functools.lru_cache(maxsize=None)
def _device_t2l(width: int, depth: int, device_index: int):
    rowptr, cols, vals = HallSet(width, depth).t2l_csr()
    dev = torch.device("cuda", device_index)
    return (torch.from_numpy(rowptr).to(dev), torch.from_numpy(cols).to(dev), torch.from_numpy(vals).to(dev))


def logsig_dim(width: int, depth: int) -> int:
    return len(HallSet(width, depth))


def calc_paths(data: torch.Tensor, stepsize, depth: int, *, compat: bool = True, chunk: int | None = None,
               out: torch.Tensor | None = None, kernel: str | None = None) -> torch.Tensor:
    """data[B, L, C] (CUDA fp32) -> logsigs[B, K, D].

    Same argument meaning as the reference: ``stepsize`` is clamped to L+1,
    column 0 of the result is 0, columns 1..C the increments, then the Hall
    coordinates of levels 2.. in Hall order.  ``compat`` reproduces the
    remainder-window quirk of generate_paths.py:59 for a batch processed in
    chunks of ``chunk`` samples (default: the whole call, as in calc_paths).
    depth >= 3 (<= 4) extends the reference (which supports 1 and 2 only).
    ``kernel``: ``None`` lets the library choose between the two closed-form kernels (C <= 8, depth <= 2) by the
    number of windows per path; ``"tma"`` / ``"classic"`` force the bulk-copy / classic-staging kernel
    (LNCDE_FLAG_K1_TMA / LNCDE_FLAG_GENERIC) - bit-identical results.
    """
    _lib.require_cuda(data)
    if data.dtype != torch.float32 or data.dim() != 3:
        raise _lib.LncdeError("calc_paths expects a float32 tensor [B, L, C]")
    L = _lib.lib()
    B, Ln, Cn = data.shape
    stepsize = int(stepsize)
    K = L.lncde_logsig_num_windows(Ln, stepsize)
    _lib.check(min(K, 0), "lncde_logsig_num_windows")
    D = logsig_dim(Cn, depth)
    if out is None:
        out = torch.empty((B, K, D), device=data.device, dtype=torch.float32)
    elif tuple(out.shape) != (B, K, D) or not out.is_contiguous():
        raise _lib.LncdeError("calc_paths: out has the wrong shape")
    if kernel not in (None, "classic", "tma"):
        raise _lib.LncdeError(f"calc_paths: unknown kernel {kernel!r}")
    if B == 0:  # an empty batch (e.g. the empty shard of preprocess_shard) is an empty result, as under jax.vmap
        return out
    flags = (_lib.FLAG_COMPAT if compat else 0) | _lib.call_flags()
    if kernel == "tma":
        flags = (flags | _lib.FLAG_K1_TMA) & ~_lib.FLAG_GENERIC
    elif kernel == "classic":
        flags = (flags | _lib.FLAG_GENERIC) & ~_lib.FLAG_K1_TMA
    rowptr, cols, vals = _device_t2l(Cn, depth, data.device.index or 0)
    with torch.cuda.device(data.device):
        st = L.lncde_logsig_fwd(_lib.stream_ptr(), _lib.ptr(data), _lib.ptr(out), B, Ln, Cn, stepsize, depth,
                                flags, int(chunk or B), _lib.ptr(rowptr), _lib.ptr(cols),
                                _lib.ptr(vals))
    _lib.check(st, "lncde_logsig_fwd")
    return out


def batch_calc_paths(data: torch.Tensor, stepsize, depth: int, *, compat: bool = True) -> torch.Tensor:
    """datasets.py:43-71: the reference loops over chunks of 128 samples; here the
    whole dataset is ONE launch and the chunk boundary only selects which samples
    have no predecessor for the remainder-window quirk."""
    return calc_paths(data, stepsize, depth, compat=compat, chunk=128)


def preprocess_shard(n_samples: int, rank: int, world_size: int, chunk: int = 128) -> slice:
    """Samples one rank preprocesses when a dataset is split over the GPUs of a node.

    Shards are whole ``chunk``-sample groups of ``batch_calc_paths`` (datasets.py:45): the cross-sample basepoint of
    the remainder window (quirk B1) only reaches back inside such a group, so no shard needs a halo row of its
    neighbour and there is no collective - concatenating the ranks' ``batch_calc_paths(data[shard])`` in rank order
    is bit-for-bit the single-process result.  Ranks beyond the number of groups get an empty slice."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("rank must be in [0, world_size)")
    n_chunks = -(-int(n_samples) // int(chunk))
    lo, hi = (n_chunks * rank) // world_size, (n_chunks * (rank + 1)) // world_size
    return slice(min(n_samples, lo * chunk), min(n_samples, hi * chunk))
