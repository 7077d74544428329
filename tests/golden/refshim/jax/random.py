"""jax.random stand-in: keys are integers, draws come from seeded torch generators (values are arbitrary -
the golden script overwrites every parameter it cares about)."""
import torch


def PRNGKey(seed):
    return int(seed)


key = PRNGKey


def split(key, n=2):
    return [int(key) * 1000003 + 7919 * (i + 1) for i in range(n)]


def _gen(key):
    return torch.Generator().manual_seed(int(key) % (2 ** 62))


def normal(key, shape=(), dtype=None):
    return torch.randn(tuple(shape), generator=_gen(key), dtype=torch.float64)


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return torch.rand(tuple(shape), generator=_gen(key), dtype=torch.float64) * (maxval - minval) + minval


def permutation(key, n):
    if isinstance(n, torch.Tensor) and n.dim() >= 1:  # permute the entries of an array
        return n[torch.randperm(n.shape[0], generator=_gen(key))]
    return torch.randperm(int(n), generator=_gen(key))
