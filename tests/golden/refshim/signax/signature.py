import torch


def _exp(z, depth):
    """Truncated tensor exponential of a level-1 element: levels z^{(x) n} / n!."""
    out = [z]
    for n in range(2, depth + 1):
        out.append(torch.einsum("...i,j->...ij", out[-1], z) / n)
    return out


def _mul(a, b, depth):
    """Truncated product of two group-like elements (scalar parts 1), levels 1..depth."""
    out = []
    for n in range(1, depth + 1):
        acc = a[n - 1] + b[n - 1]
        for k in range(1, n):
            acc = acc + torch.tensordot(a[k - 1], b[n - k - 1], dims=0)
        out.append(acc)
    return out


def signature(path, depth, **kw):
    """path [n_points, C] -> list of levels [(C,), (C,C), ...] of the truncated signature (Chen's identity)."""
    path = torch.as_tensor(path)
    inc = path[1:] - path[:-1]
    sig = _exp(inc[0], depth)
    for i in range(1, inc.shape[0]):
        sig = _mul(sig, _exp(inc[i], depth), depth)
    return sig
