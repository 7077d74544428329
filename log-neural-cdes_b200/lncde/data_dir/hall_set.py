"""``HallSet`` over the native K0 tables (mirror of data_dir/hall_set.py:60-256).

Only what the hot path reads is exposed: ``.data`` (the (left, right) table,
index 0 = (0, 0), read by generate_paths.py:35 and LogNeuralCDEs.py:93-97) and
``.t2l_matrix(degree)`` (generate_paths.py:36).  Tables are host-side integer /
rational data produced by ``lncde_hall_*`` in csrc/hall_tables.cc.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np

from .. import _lib


@functools.lru_cache(maxsize=None)
def _tables(width: int, degree: int):
    L = _lib.lib()
    n = L.lncde_hall_size(width, degree)
    _lib.check(min(n, 0), "lncde_hall_size")
    pairs = np.zeros((n, 2), np.int32)
    _lib.check(L.lncde_hall_basis(width, degree, pairs.ctypes.data_as(C.c_void_p)), "lncde_hall_basis")
    nnz = L.lncde_hall_t2l_nnz(width, degree)
    rowptr = np.zeros(n + 1, np.int32)
    cols = np.zeros(nnz, np.int32)
    vals = np.zeros(nnz, np.float32)
    _lib.check(
        L.lncde_hall_t2l_csr(width, degree, rowptr.ctypes.data_as(C.c_void_p), cols.ctypes.data_as(C.c_void_p),
                             vals.ctypes.data_as(C.c_void_p)), "lncde_hall_t2l_csr")
    return pairs, rowptr, cols, vals


class HallSet:
    def __init__(self, width: int, degree: int = 1):
        self.width = int(width)
        self.degree = int(degree)
        pairs, self._rowptr, self._cols, self._vals = _tables(self.width, self.degree)
        self.data = [(int(a), int(b)) for a, b in pairs]

    def __len__(self):
        return len(self.data)

    def t2l_csr(self):
        """(rowptr, cols, vals) of t2l_matrix(self.degree); cols include the scalar slot."""
        return self._rowptr, self._cols, self._vals

    def t2l_matrix(self, degree=None, dtype=np.float32) -> np.ndarray:
        degree = self.degree if degree is None else int(degree)
        if degree != self.degree:
            return HallSet(self.width, degree).t2l_matrix(degree, dtype)
        n = len(self.data)
        tdim = int(_lib.lib().lncde_tensor_dim(self.width, degree))
        out = np.zeros((n, tdim), dtype)
        for k in range(n):
            lo, hi = self._rowptr[k], self._rowptr[k + 1]
            out[k, self._cols[lo:hi]] = self._vals[lo:hi]
        return out

    def pairs(self) -> np.ndarray:
        """``jnp.asarray(hs.data[1:])`` of LogNeuralCDEs.py:97."""
        return np.asarray(self.data[1:], np.int32).reshape(-1, 2)

    def basis_list(self):
        """Nested-list keys in Hall order – what LinearNeuralCDEs.py:111-116 builds from roughpy."""

        def nested(k):
            l, r = self.data[k]
            return r if l == 0 else [nested(l), nested(r)]

        return [nested(k) for k in range(1, len(self.data))]
