"""Stand-in for the part of `equinox` the reference hot path uses.  See ../README.md.

eqx.nn.Linear: y = W x + b with W[out, in] (equinox 0.12 convention); eqx.nn.MLP: `depth` hidden layers of
`width_size`, `activation` after each, `final_activation` after the output layer."""
import copy
import math

import torch

from . import nn  # noqa: F401


class Module:
    _is_module = True

    def __init_subclass__(cls, **kw):  # annotated fields with defaults stay plain class attributes
        super().__init_subclass__(**kw)


def field(**kw):
    return None


def filter_jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def is_inexact_array(x):
    return isinstance(x, torch.Tensor) and x.is_floating_point()


def is_array(x):
    return isinstance(x, torch.Tensor)


def combine(*models):
    return next(m for m in models if m is not None)


def partition(model, spec):
    return model, None


def filter(model, spec, **kw):  # noqa: A001
    return model


def tree_inference(model, value=True):
    return model


def tree_at(where, pytree, replace):
    """Functional update of the leaves selected by `where` (matched by identity)."""
    new = copy.deepcopy(pytree)
    old_sel = where(pytree)
    new_sel = where(new)
    single = not isinstance(old_sel, (list, tuple))
    if single:
        old_sel, new_sel, replace = [old_sel], [new_sel], [replace]
    targets = {id(t): r for t, r in zip(new_sel, replace)}

    def walk(obj):
        if getattr(obj, "_is_module", False):
            for k, v in list(vars(obj).items()):
                if id(v) in targets:
                    setattr(obj, k, targets[id(v)])
                else:
                    walk(v)
        elif isinstance(obj, (list, tuple)):
            for v in obj:
                walk(v)

    walk(new)
    return new


def filter_value_and_grad(f=None, has_aux=False):
    """Value only (gradients are taken by torch autograd in the golden script where needed)."""
    def deco(fn):
        def wrapped(*a, **k):
            return fn(*a, **k), None
        return wrapped
    return deco if f is None else deco(f)


def apply_updates(model, updates):
    return model
