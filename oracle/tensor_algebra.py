"""Oracle: truncated tensor algebra – signature, log, flatten.  TEST INFRASTRUCTURE ONLY.

Restates the published algorithm of ``signax==0.1.1`` (pinned in the reference's
README.md:98; NOT vendored in /root/reference, not installed offline) as used at
``data_dir/generate_paths.py:7-9,15``:

* ``signature(path, depth)``: ``inc = diff(path)``; ``S = restricted_exp(inc[0])``;
  for i = 1..n-1: ``S = mult_fused_restricted_exp(inc[i], S)``, whose level k is
  the Horner form ``A_k + (A_{k-1} + ( ... (A_1 + z/k) (x) z/(k-1) ... )) (x) z``.
* ``log``: truncated series ``log(1+x) = sum_n (-1)^(n+1) x^n / n``.
* ``flatten``: concatenation of the C-order ravelled levels 1..depth (no scalar).

PARITY UNPINNED against signax itself (not installable offline, no reference test vectors);
anchored by identities in tests/test_oracle_logsig.py and by agreement (1e-14) with the independent
implementation in tests/golden/refshim/signax that the reference's calc_paths runs on when the
golden fixtures are generated.

All functions broadcast over leading batch axes; levels are arrays of shape
``batch + (C,)*k``.
"""
from __future__ import annotations

import numpy as np


def _otimes(a: np.ndarray, b: np.ndarray, ka: int, kb: int) -> np.ndarray:
    """a: batch+(C,)*ka, b: batch+(C,)*kb -> batch+(C,)*(ka+kb)."""
    return a.reshape(a.shape + (1,) * kb) * b.reshape(b.shape[: b.ndim - kb] + (1,) * ka + b.shape[b.ndim - kb :])


def restricted_exp(z: np.ndarray, depth: int):
    out = [z]
    for k in range(2, depth + 1):
        out.append(_otimes(out[-1], z, k - 1, 1) / z.dtype.type(k))
    return out


def mult_fused_restricted_exp(z: np.ndarray, A: list):
    depth = len(A)
    out = []
    for k in range(1, depth + 1):
        cur = A[0] + z / z.dtype.type(k)
        for i in range(1, k):
            cur = A[i] + _otimes(cur, z / z.dtype.type(k - i), i, 1)
        out.append(cur)
    return out


def signature(path: np.ndarray, depth: int):
    """path: batch + (n_points, C) -> list of levels (signax.signature.signature)."""
    inc = np.diff(path, axis=-2)
    n = inc.shape[-2]
    S = restricted_exp(inc[..., 0, :], depth)
    for i in range(1, n):
        S = mult_fused_restricted_exp(inc[..., i, :], S)
    return S


def tensor_mul(A: list, B: list):
    """Product of two elements WITHOUT scalar part (scalar 0), truncated."""
    depth = len(A)
    out = [np.zeros_like(a) for a in A]
    for ka in range(1, depth + 1):
        for kb in range(1, depth + 1 - ka):
            out[ka + kb - 1] = out[ka + kb - 1] + _otimes(A[ka - 1], B[kb - 1], ka, kb)
    return out


def log(S: list):
    """log(1 + S), truncated (signax.tensor_ops.log); S has no scalar part."""
    depth = len(S)
    dt = S[0].dtype.type
    # Horner: log(1+x) = x (1 - x (1/2 - x (1/3 - ...)))
    # acc_n = 1/n - x * acc_{n+1}
    acc_scalar = dt(1.0) / dt(depth)
    acc = [np.zeros_like(s) for s in S]  # non-scalar part of acc
    for n in range(depth - 1, 0, -1):
        # new = 1/n - x * (acc_scalar + acc)
        prod = tensor_mul(S, acc)
        acc = [-(acc_scalar * s + p) for s, p in zip(S, prod)]
        acc_scalar = dt(1.0) / dt(n)
    prod = tensor_mul(S, acc)
    return [acc_scalar * s + p for s, p in zip(S, prod)]


def flatten(levels: list) -> np.ndarray:
    batch = levels[0].shape[:-1]
    return np.concatenate([lv.reshape(batch + (-1,)) for lv in levels], axis=-1)
