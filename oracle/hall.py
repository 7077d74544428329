"""Oracle: Hall basis and tensor<->Lie maps.  TEST INFRASTRUCTURE ONLY.

Restates ``data_dir/hall_set.py:60-256`` of the reference (pure Python).
PINNED: ``tests/test_oracle_hall.py`` checks every table below for exact
equality with the reference's own ``HallSet`` through ``tests/golden/hall_*.npz``.

Conventions (reference file:line):
* ``basis[0] = (0, 0)``; letters are ``(0, i)`` for i = 1..width  (hall_set.py:72-80)
* a degree-n element is a pair ``(i, j)`` of earlier indices with ``i < j`` and
  ``basis[j][0] <= i``  (hall_set.py:89-119)
* tensor word (l1..ln), letters 1-based, has flat index ``sum l_k * width**(n-k)``
  which is its position in [scalar | level 1 | level 2 (C order) | ...]
  (hall_set.py:16-25)
* ``t2l[k, idx(word)] = coeff_k(rbracket(word)) / len(word)``  (hall_set.py:234-256)
"""
from __future__ import annotations

import numpy as np


def tensor_algebra_dim(width: int, depth: int) -> int:
    """1 + w + w^2 + ... + w^depth  (hall_set.py:28-33)."""
    return sum(width**k for k in range(depth + 1))


class Hall:
    def __init__(self, width: int, depth: int):
        self.width, self.depth = width, depth
        basis = [(0, 0)] + [(0, i) for i in range(1, width + 1)]
        index = {(0, i): i for i in range(1, width + 1)}
        ranges = [(0, 1), (1, width + 1)]  # ranges[d] = [lo, hi) of degree d
        for deg in range(2, depth + 1):
            lo = len(basis)
            for ldeg in range(1, deg // 2 + 1):
                ilo, ihi = ranges[ldeg]
                jlo, jhi = ranges[deg - ldeg]
                for i in range(ilo, ihi):
                    for j in range(max(jlo, i + 1), jhi):
                        if basis[j][0] <= i:
                            index[(i, j)] = len(basis)
                            basis.append((i, j))
            ranges.append((lo, len(basis)))
        self.basis, self.index, self.ranges = basis, index, ranges
        self.degree_of = [0] * len(basis)
        for d, (lo, hi) in enumerate(ranges):
            for k in range(lo, hi):
                self.degree_of[k] = d
        self._prod = {}
        self._rb = {}
        self._exp = {}

    def __len__(self):
        return len(self.basis)

    # [a, b] in the Hall basis as {key: int coeff}  (hall_set.py:132-161)
    def bracket(self, a: int, b: int) -> dict:
        if a == b:
            return {}
        if b < a:
            return {k: -c for k, c in self.bracket(b, a).items()}
        got = self._prod.get((a, b))
        if got is not None:
            return got
        if (a, b) in self.index:
            out = {self.index[(a, b)]: 1}
        else:
            # Jacobi: [a,[l,r]] = [[a,l],r] - [[a,r],l]
            l, r = self.basis[b]
            out = {}
            for k1, c1 in self.bracket(a, l).items():
                for k, c in self.bracket(k1, r).items():
                    out[k] = out.get(k, 0) + c1 * c
            for k1, c1 in self.bracket(a, r).items():
                for k, c in self.bracket(k1, l).items():
                    out[k] = out.get(k, 0) - c1 * c
            out = {k: c for k, c in out.items() if c != 0}
        self._prod[(a, b)] = out
        return out

    # right-nested bracket [l1,[l2,[..., ln]]]  (hall_set.py:191-208)
    def rbracket(self, word: tuple) -> dict:
        if len(word) == 0:
            return {}
        if len(word) == 1:
            return {word[0]: 1}
        got = self._rb.get(word)
        if got is not None:
            return got
        out = {}
        for k1, c1 in self.rbracket(word[1:]).items():
            for k, c in self.bracket(word[0], k1).items():
                out[k] = out.get(k, 0) + c1 * c
        out = {k: c for k, c in out.items() if c != 0}
        self._rb[word] = out
        return out

    # expansion of a Hall element as a tensor  (hall_set.py:163-189)
    def expand(self, key: int) -> dict:
        l, r = self.basis[key]
        if l == 0:
            return {(r,): 1}
        got = self._exp.get(key)
        if got is not None:
            return got
        le, re = self.expand(l), self.expand(r)
        out = {}
        for w1, c1 in le.items():
            for w2, c2 in re.items():
                out[w1 + w2] = out.get(w1 + w2, 0) + c1 * c2
                out[w2 + w1] = out.get(w2 + w1, 0) - c1 * c2
        out = {k: c for k, c in out.items() if c != 0}
        self._exp[key] = out
        return out

    def word_index(self, word: tuple) -> int:
        idx = 0
        for letter in word:
            idx = idx * self.width + letter
        return idx

    def words(self, depth=None):
        depth = self.depth if depth is None else depth
        yield ()
        level = [()]
        for _ in range(depth):
            level = [w + (i,) for w in level for i in range(1, self.width + 1)]
            yield from level

    def t2l_matrix(self, depth=None, dtype=np.float32) -> np.ndarray:
        depth = self.depth if depth is None else depth
        n = self.ranges[depth][1]
        out = np.zeros((n, tensor_algebra_dim(self.width, depth)), dtype=np.float64)
        for word in self.words(depth):
            for k, c in self.rbracket(word).items():
                out[k, self.word_index(word)] += c / len(word)
        return out.astype(dtype)

    def l2t_matrix(self, depth=None, dtype=np.float32) -> np.ndarray:
        depth = self.depth if depth is None else depth
        n = self.ranges[depth][1]
        out = np.zeros((tensor_algebra_dim(self.width, depth), n), dtype=np.float64)
        for key in range(1, n):
            for word, c in self.expand(key).items():
                out[self.word_index(word), key] += c
        return out.astype(dtype)

    def pairs(self) -> np.ndarray:
        """``HallSet.data[1:]`` as int32 [(D-1), 2]  (LogNeuralCDEs.py:93-97)."""
        return np.asarray(self.basis[1:], dtype=np.int32).reshape(-1, 2)

    def key_to_nested(self, key: int):
        """Nested-list form ``1``, ``[1, 2]``, ``[1, [1, 2]]`` – the roughpy
        ``index_to_key`` printout that LinearNeuralCDEs.py:111-116 eval()s."""
        l, r = self.basis[key]
        if l == 0:
            return r
        return [self.key_to_nested(l), self.key_to_nested(r)]

    def basis_list(self):
        return [self.key_to_nested(k) for k in range(1, len(self.basis))]
