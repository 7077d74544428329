import torch


def log(sig):
    """log(1 + x) = sum_k (-1)^{k+1} x^k / k in the truncated tensor algebra; x = levels 1..depth."""
    depth = len(sig)

    def mul(a, b):  # product of two elements WITHOUT scalar part
        out = []
        for n in range(1, depth + 1):
            acc = torch.zeros_like(sig[n - 1])
            for k in range(1, n):
                acc = acc + torch.tensordot(a[k - 1], b[n - k - 1], dims=0)
            out.append(acc)
        return out

    res = [s.clone() for s in sig]
    power = [s.clone() for s in sig]
    for k in range(2, depth + 1):
        power = mul(power, sig)
        for n in range(depth):
            res[n] = res[n] + ((-1) ** (k + 1) / k) * power[n]
    return res
