"""random_bcoo stand-in: `nse` non-zeros at random positions of a dense tensor (enough for LinearNeuralCDEs.py:138-150,
:321-324: scalar * and /, .todense(), and train.py's regulariser)."""
import torch

import jax.random as jr


class _Sparse:
    def __init__(self, dense):
        self._dense = dense

    def todense(self):
        return self._dense

    def __mul__(self, s):
        return _Sparse(self._dense * s)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return _Sparse(self._dense / s)


def random_bcoo(key, shape, *, nse, generator=None, dtype=None, **kw):
    k1, k2 = jr.split(key, 2)
    total = 1
    for d in shape:
        total *= d
    pos = jr.permutation(k1, total)[: int(nse)]
    flat = torch.zeros(total, dtype=torch.float64)
    flat[pos] = jr.normal(k2, (int(nse),))
    return _Sparse(flat.reshape(tuple(shape)))
