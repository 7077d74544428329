"""Stand-in for the part of `jax` the reference hot path uses (torch float64 underneath).  See ../README.md."""
import torch

from . import numpy, random, nn, lax, tree_util  # noqa: F401,E402
from . import experimental  # noqa: F401,E402

Array = torch.Tensor


def jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def _slice_arg(a, axis, i):
    if axis is None:
        return a
    if isinstance(a, (tuple, list)):
        return type(a)(_slice_arg(x, axis, i) for x in a)
    return a.select(axis, i)


def _batch(a, axis):
    if isinstance(a, (tuple, list)):
        return _batch(a[0], axis)
    return a.shape[axis]


def _stack(outs):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_stack([o[k] for o in outs]) for k in range(len(first)))
    if first is None:
        return None
    return torch.stack([numpy.asarray(o) for o in outs], 0)


def vmap(fun, in_axes=0, out_axes=0):
    """Loop-based vmap: slices every mapped argument (or every leaf of a tuple argument) along its axis."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_batch(a, ax) for a, ax in zip(args, axes) if ax is not None)
        outs = [fun(*[_slice_arg(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        return _stack(outs)
    return mapped


def jvp(fun, primals, tangents):
    return torch.func.jvp(fun, tuple(primals), tuple(tangents))
