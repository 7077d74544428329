"""jax.numpy stand-in on torch (float64 unless the caller passes fp32 arrays, e.g. time grids)."""
import numpy as _np
import torch

ndarray = torch.Tensor
float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
float = torch.float32  # noqa: A001  (jnp.float appears in annotations only)
pi = _np.pi


def asarray(x, dtype=None):
    if dtype is torch.float32:
        dtype = torch.float64  # value arrays run in the arbiter's precision (time grids never pass through here)
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, _np.ndarray):
        t = torch.from_numpy(_np.ascontiguousarray(x))
        if dtype is None and t.dtype == torch.float32:
            t = t.double()  # torch does not promote mixed dtypes in `@`; constants (t2l: +-1/2) are exact either way
        return t if dtype is None else t.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack(list(x))
    t = torch.as_tensor(x)
    if dtype is None and t.dtype == torch.float32:
        t = t.double()  # python floats: keep the arbiter's precision
    return t if dtype is None else t.to(dtype)


array = asarray


def concatenate(seq, axis=0):
    return torch.cat([asarray(s) for s in seq], dim=axis)


def stack(seq, axis=0):
    return torch.stack([asarray(s) for s in seq], dim=axis)


def vstack(seq):
    return torch.cat([torch.atleast_2d(asarray(s)) for s in seq], dim=0)


def zeros(shape, dtype=torch.float64):
    return torch.zeros(shape, dtype=dtype)


def ones(shape, dtype=torch.float64):
    return torch.ones(shape, dtype=dtype)


def eye(n, dtype=torch.float64):
    return torch.eye(n, dtype=dtype)


def arange(*a, dtype=None):
    # jnp.arange on concrete Python floats: numpy computes the range (and its LENGTH) in float64, the result is then
    # canonicalised to float32 (x64 is off in the reference)
    if any(isinstance(v, (__builtins__["float"] if isinstance(__builtins__, dict) else __builtins__.float, _np.floating)) for v in a):
        return torch.from_numpy(_np.arange(*a).astype(_np.float32))
    return torch.arange(*a)


def reshape(x, shape):
    return asarray(x).reshape(shape)


def dot(a, b):
    a, b = asarray(a), asarray(b)
    if a.dtype != b.dtype:
        a, b = a.to(torch.float64), b.to(torch.float64)
    return a @ b


def matmul(a, b):
    return torch.matmul(asarray(a), asarray(b))


def einsum(spec, *ops):
    ops = [asarray(o) for o in ops]
    if len({o.dtype for o in ops}) > 1:
        ops = [o.to(torch.float64) for o in ops]
    return torch.einsum(spec, *ops)


def searchsorted(a, v, side="left"):
    return torch.searchsorted(asarray(a), asarray(v), right=(side == "right"))


def mean(x, axis=None):
    return asarray(x).mean() if axis is None else asarray(x).mean(dim=axis)


def sum(x, axis=None):  # noqa: A001
    return asarray(x).sum() if axis is None else asarray(x).sum(dim=axis)


def log(x):
    return torch.log(asarray(x))


def sqrt(x):
    return torch.sqrt(asarray(x, None).double()) if not isinstance(x, torch.Tensor) else torch.sqrt(x)


def tanh(x):
    return torch.tanh(x)


def round(x):  # noqa: A001
    return torch.round(asarray(x))


def diag(x):
    return torch.diag(x)


def sign(x):
    return torch.sign(x)


def argmax(x, axis=None):
    return torch.argmax(x) if axis is None else torch.argmax(x, dim=axis)


def repeat(x, n, axis=None):
    return torch.repeat_interleave(asarray(x), n, dim=axis)


def unique(x):
    return torch.unique(asarray(x))


class linalg:
    @staticmethod
    def norm(x, axis=None):
        return torch.linalg.norm(x) if axis is None else torch.linalg.norm(x, dim=axis)
