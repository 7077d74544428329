"""Oracle: Neural RDE forward / gradients on CPU (torch).  TEST INFRASTRUCTURE ONLY.

Restates ``NeuralRDE`` of ``models/NeuralCDEs.py:168-266``:
  * ``vf = VectorField(hidden_dim, vf_hidden_dim, vf_hidden_dim, vf_num_hidden - 1)`` (:209-216): an
    ``eqx.nn.MLP`` with ``vf_num_hidden`` linear layers H -> w -> ... -> w, the VectorField default
    ``activation=jax.nn.relu`` between them (:47) and ``final_activation=tanh`` (:54); weights and biases
    divided by ``scale`` (:59-79);
  * ``mlp_linear = eqx.nn.Linear(w, H * (logsig_dim - 1))`` (:217-219), NOT scaled;
  * ``func`` (:230-238): ``reshape(mlp_linear(vf(y)), (H, Dm)) @ logsig[idx-1][1:] / (intervals[idx] -
    intervals[idx-1])`` with ``idx = searchsorted(intervals, t)`` - the same row / interval-length selection as
    the Log-NCDE (``oracle.log_ncde.stage_table``);
  * the diffrax Heun / ConstantStepSize loop and the read-outs (:240-266), shared with ``oracle.log_ncde``.

PINNED to the reference's own NeuralRDE.__call__ / create_model("nrde"), executed from /root/reference under the
library stand-ins (tests/golden/ref_nrde.npz, tests/test_oracle_reference_pinned.py); the diffrax / equinox
semantics underneath are restated.
"""
from __future__ import annotations

import numpy as np
import torch


def layer_sizes(H, Dm, width, n_vf):
    """[(in, out)] of the n_vf layers of vf followed by mlp_linear."""
    dims = [H] + [width] * n_vf
    return list(zip(dims[:-1], dims[1:])) + [(width, H * Dm)]


def init_params(H, Dm, width, n_vf, scale=1.0, seed=0, dtype=torch.float32):
    """eqx.nn.Linear init U(-1/sqrt(in), 1/sqrt(in)); the vf layers (not mlp_linear) are divided by `scale`."""
    g = torch.Generator().manual_seed(seed)
    out = []
    sizes = layer_sizes(H, Dm, width, n_vf)
    for li, (i, o) in enumerate(sizes):
        lim = 1.0 / np.sqrt(i)
        s = scale if li < n_vf else 1.0
        W = (torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * lim / s
        b = (torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * lim / s
        out.append((W.to(dtype), b.to(dtype)))
    return out


def flatten_params(layers) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for W, b in layers for t in (W, b)])


def unflatten_params(flat, H, Dm, width, n_vf):
    out, off = [], 0
    for i, o in layer_sizes(H, Dm, width, n_vf):
        W = flat[off : off + o * i].reshape(o, i)
        off += o * i
        b = flat[off : off + o]
        off += o
        out.append((W, b))
    return out


def field(layers, y, ls_row, H):
    """NeuralCDEs.py:230-238 WITHOUT the 1/interval factor, one sample."""
    a = y
    n_vf = len(layers) - 1
    for l in range(n_vf):
        W, b = layers[l]
        z = W @ a + b
        a = torch.relu(z) if l < n_vf - 1 else torch.tanh(z)
    Wm, bm = layers[-1]
    return (Wm @ a + bm).reshape(H, -1) @ ls_row[1:]


def solve(layers, logsig, y0, rows, scale, H):
    """Heun solve, batched.  logsig [B,K,D], y0 [B,H]; returns ys [B, n+1, H]."""
    f = torch.func.vmap(lambda y, r: field(layers, y, r, H))
    ys, y = [y0], y0
    rows_t = torch.from_numpy(np.asarray(rows, np.int64))
    sc = torch.from_numpy(np.asarray(scale)).to(y0.dtype)
    for k in range(rows_t.shape[0]):
        k1 = sc[k, 0] * f(y, logsig[:, rows_t[k, 0]])
        k2 = sc[k, 1] * f(y + k1, logsig[:, rows_t[k, 1]])
        y = y + 0.5 * (k1 + k2)
        ys.append(y)
    return torch.stack(ys, dim=1)
