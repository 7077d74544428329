"""jax.lax.scan / associative_scan stand-ins (sequential; the combine functions used are associative)."""
import torch

from . import tree_util


def scan(f, init, xs, length=None):
    leaves = tree_util.tree_leaves(xs)
    n = leaves[0].shape[0] if leaves else length
    carry, ys = init, []
    for i in range(n):
        carry, y = f(carry, tree_util.tree_map(lambda a: a[i], xs))
        ys.append(y)
    if not ys:  # zero-length scan: jax returns empty stacked outputs (here: shaped like the carry)
        return carry, torch.zeros((0,) + tuple(init.shape), dtype=init.dtype)
    return carry, tree_util.tree_map(lambda *a: torch.stack(a, 0), *ys)


def associative_scan(fn, elems):
    """Inclusive prefix combine out[i] = fn(out[i-1], elems[i]) (same values as jax's tree order up to rounding)."""
    leaves = tree_util.tree_leaves(elems)
    n = leaves[0].shape[0]
    outs = [tree_util.tree_map(lambda a: a[0], elems)]
    for i in range(1, n):
        outs.append(fn(outs[-1], tree_util.tree_map(lambda a: a[i], elems)))
    return tree_util.tree_map(lambda *a: torch.stack(a, 0), *outs)
