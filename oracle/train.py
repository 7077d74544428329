"""Oracle: one Log-NCDE train step on CPU (torch fp32).  TEST INFRASTRUCTURE ONLY.

Restates train.py:71-136 (loss, value_and_grad, optax.adam update) around
oracle.log_ncde.solve.  Used by tests and as the timed CPU baseline of bench.py
(kind "port": the reference's JAX-CPU path cannot run offline).
The loss formulas are pinned through oracle.log_ncde / oracle.linear_ncde (tests/golden/ref_*.npz);
the optax.adam update is restated (optax absent offline): unpinned."""
from __future__ import annotations

import numpy as np
import torch

from . import log_ncde as O


def adam_update(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam: scale_by_adam (bias-corrected, eps outside sqrt) then -lr."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    mh = m / (1 - b1**step)
    vh = v / (1 - b2**step)
    p.sub_(lr * mh / (vh.sqrt() + eps))


class LogNCDEStep:
    """Holds parameters (MLP layers, linear1, linear2) and performs train steps."""

    def __init__(self, H, C, width, n_hidden, label_dim, depth, scale, lambd, lr, seed=0, dtype=torch.float32):
        self.dims = (H, C, width, n_hidden, depth)
        self.layers = [(W.requires_grad_(True), b.requires_grad_(True))
                       for W, b in O.init_params(H, C, width, n_hidden, scale, seed, dtype)]
        g = torch.Generator().manual_seed(seed + 1)
        u = lambda o, i: ((torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) / np.sqrt(i)).to(dtype)
        self.W1, self.b1 = u(H, C).requires_grad_(True), u(1, C)[0, :1].repeat(H).requires_grad_(True)
        self.W2, self.b2 = u(label_dim, H).requires_grad_(True), u(1, H)[0, :1].repeat(label_dim).requires_grad_(True)
        self.lambd, self.lr, self.count = lambd, lr, 0
        self.leaves = [t for W, b in self.layers for t in (W, b)] + [self.W1, self.b1, self.W2, self.b2]
        self.m = [torch.zeros_like(t) for t in self.leaves]
        self.v = [torch.zeros_like(t) for t in self.leaves]

    def loss(self, logsig, x0, y_onehot, rows, scale):
        H, C, width, n_hidden, depth = self.dims
        y0 = x0 @ self.W1.T + self.b1
        ys = O.solve(self.layers, logsig, y0, rows, scale, C, H, depth)
        pred = torch.softmax(ys[:, -1] @ self.W2.T + self.b2, -1)
        return O.classification_loss(pred, y_onehot) + O.lip2_norm(self.layers, self.lambd)

    def step(self, logsig, x0, y_onehot, rows, scale):
        loss = self.loss(logsig, x0, y_onehot, rows, scale)
        grads = torch.autograd.grad(loss, self.leaves)
        self.count += 1
        with torch.no_grad():
            for p, g, m, v in zip(self.leaves, grads, self.m, self.v):
                adam_update(p, g, m, v, self.count, self.lr)
        return float(loss.detach())
