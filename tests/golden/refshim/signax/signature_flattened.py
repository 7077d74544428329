import torch


def flatten(levels):
    return torch.cat([l.reshape(-1) for l in levels])
