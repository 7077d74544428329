"""Oracle: Log-NCDE forward / loss / gradients on CPU (torch).  TEST INFRASTRUCTURE ONLY.

Restates ``models/LogNeuralCDEs.py:107-162`` (the Lie-bracket Log-ODE field in the
reference's own formulation: 1 primal + C JVPs of the whole MLP, antisymmetrised
over ``pairs``), ``models/NeuralCDEs.py:43-82`` (VectorField: eqx.nn.MLP, SiLU
hidden, tanh final), the diffrax 0.7.0 Heun / ConstantStepSize loop it drives
(un-vendored third party: semantics restated from its published algorithm –
tableau c=[1], a=[[1]], b=[1/2,1/2]; ``tnext = tprev + dt``, clipped to t1 when
within 1e-6 in fp32; default adjoint = exact gradient of the discrete
recurrences, which torch autograd of this loop reproduces) and the read-outs.

PINNED to the reference's own LogNeuralCDE.__call__ / create_model / losses, executed from
/root/reference under library stand-ins (tests/golden/ref_logncde.npz,
tests/test_oracle_reference_pinned.py); the diffrax / equinox semantics are restated.
"""
from __future__ import annotations

import numpy as np
import torch

from .hall import Hall


# --------------------------------------------------------------------------- stage table
def step_grid(t0: np.float32, t1: np.float32, dt0: float, max_steps: int = 16**4) -> np.ndarray:
    """Times t_0 < t_1 < ... of the constant-step solve, all arithmetic in fp32
    (diffrax ConstantStepSize: next = t + (t - tprev); _clip_to_end tol 1e-6)."""
    f = np.float32
    ts = [f(t0)]
    tprev, tnext = f(t0), f(f(t0) + f(dt0))
    while True:
        if tnext > f(t1) - f(1e-6):
            tnext = f(t1)
        ts.append(tnext)
        if tnext >= f(t1) or len(ts) > max_steps:
            break
        tprev, tnext = tnext, f(tnext + f(tnext - tprev))
    return np.asarray(ts, np.float32)


def stage_table(ts: np.ndarray, intervals: np.ndarray, dt0: float, n_rows: int):
    """Per Heun stage: the logsig row it reads and dt / interval-length.

    LogNeuralCDEs.py:114-115,120,138: ``idx = searchsorted(intervals, t)`` (side
    left), row ``idx - 1`` and length ``intervals[idx] - intervals[idx-1]`` with
    Python/JAX negative-index wrap (quirk B2: at t = ts[0] = 0 this is the LAST
    row and -T).  Stage 1 of step n is evaluated at t_n, stage 2 at t_{n+1}.
    Returns (rows int32 [n_steps, 2], scale fp32 [n_steps, 2], grid fp32 [n_steps+1])."""
    intervals = np.asarray(intervals, np.float32)
    grid = step_grid(ts[0], ts[-1], dt0)
    n = len(grid) - 1
    rows = np.zeros((n, 2), np.int32)
    scale = np.zeros((n, 2), np.float32)
    ni = len(intervals)
    for k in range(n):
        dt = np.float32(grid[k + 1] - grid[k])
        for j, t in enumerate((grid[k], grid[k + 1])):
            idx = int(np.searchsorted(intervals, t, side="left"))
            idx = min(idx, ni - 1)  # JAX clamps out-of-range gathers
            lo = idx - 1
            delta = np.float32(intervals[idx] - intervals[lo if lo >= 0 else ni - 1])
            rows[k, j] = lo if lo >= 0 else n_rows - 1
            scale[k, j] = dt / delta
    return rows, scale, grid


# --------------------------------------------------------------------------- model
def mlp_sizes(H, C, width, n_hidden):
    return [H] + [width] * n_hidden + [C * H]


def init_params(H, C, width, n_hidden, scale=1.0, seed=0, dtype=torch.float32):
    """eqx.nn.Linear init U(-1/sqrt(in), 1/sqrt(in)) for W and b, then / scale
    (NeuralCDEs.py:59-79).  Returns list of (W[out,in], b[out])."""
    g = torch.Generator().manual_seed(seed)
    sizes = mlp_sizes(H, C, width, n_hidden)
    out = []
    for i, o in zip(sizes[:-1], sizes[1:]):
        lim = 1.0 / np.sqrt(i)
        W = (torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * lim / scale
        b = (torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * lim / scale
        out.append((W.to(dtype), b.to(dtype)))
    return out


def flatten_params(layers) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for W, b in layers for t in (W, b)])


def unflatten_params(flat: torch.Tensor, H, C, width, n_hidden):
    sizes = mlp_sizes(H, C, width, n_hidden)
    out, off = [], 0
    for i, o in zip(sizes[:-1], sizes[1:]):
        W = flat[off : off + o * i].reshape(o, i)
        off += o * i
        b = flat[off : off + o]
        off += o
        out.append((W, b))
    return out


def vf_apply(layers, y):
    a = y
    for l, (W, b) in enumerate(layers):
        z = W @ a + b
        a = torch.nn.functional.silu(z) if l < len(layers) - 1 else torch.tanh(z)
    return a


def field(layers, y, ls_row, C, H, depth, pairs):
    """LogNeuralCDEs.py:113-138 WITHOUT the 1/interval factor, one sample."""
    vf_out = vf_apply(layers, y).reshape(C, H)
    if depth == 1:
        return ls_row[1 : C + 1] @ vf_out
    jv = torch.func.vmap(lambda v: torch.func.jvp(lambda yy: vf_apply(layers, yy), (y,), (v,))[1])(vf_out)
    jv = jv.reshape(C, C, H)  # jv[a, b] = D f_b . f_a
    p = pairs[C:] - 1
    lie = jv[p[:, 0], p[:, 1]] - jv[p[:, 1], p[:, 0]]
    return ls_row[1 : C + 1] @ vf_out + ls_row[C + 1 :] @ lie


def solve(layers, logsig, y0, rows, scale, C, H, depth):
    """Heun solve, batched.  logsig [B,K,D], y0 [B,H]; returns ys [B, n+1, H]."""
    pairs = torch.from_numpy(Hall(C, depth).pairs().astype(np.int64)) if depth > 1 else None
    f = torch.func.vmap(lambda y, r: field(layers, y, r, C, H, depth, pairs))
    ys = [y0]
    y = y0
    rows_t = torch.from_numpy(np.asarray(rows, np.int64))
    sc = torch.from_numpy(np.asarray(scale)).to(y0.dtype)
    for k in range(rows_t.shape[0]):
        k1 = sc[k, 0] * f(y, logsig[:, rows_t[k, 0]])
        k2 = sc[k, 1] * f(y + k1, logsig[:, rows_t[k, 1]])
        y = y + 0.5 * (k1 + k2)
        ys.append(y)
    return torch.stack(ys, dim=1)


def interp_outputs(ys, grid, out_times):
    """diffrax SaveAt(ts=...) with Heun's local linear interpolation."""
    grid_t = np.asarray(grid, np.float32)
    idx = np.clip(np.searchsorted(grid_t, out_times, side="left"), 1, len(grid_t) - 1)
    w = ((out_times - grid_t[idx - 1]) / (grid_t[idx] - grid_t[idx - 1])).astype(np.float32)
    wt = torch.from_numpy(w).to(ys.dtype)[None, :, None]
    return ys[:, idx - 1] * (1 - wt) + ys[:, idx] * wt


def regression_times(n_ts: int, output_step: int) -> np.ndarray:
    """LogNeuralCDEs.py:143-145: arange(step, 1, step) with step = output_step/len(ts), plus t1."""
    step = output_step / n_ts  # Python float: jnp.arange computes the length in float64, then casts the values
    return np.arange(step, 1.0, step).astype(np.float32)


def lip2_norm(layers, lambd):
    """train.py:79-86: lambd * sum_layers mean(||W||_row + ||b||)."""
    norm = 0.0
    for W, b in layers:
        norm = norm + torch.mean(torch.linalg.norm(W, dim=-1) + torch.linalg.norm(b, dim=-1))
    return lambd * norm


def classification_loss(pred, y_onehot):
    """train.py:95."""
    return torch.mean(-torch.sum(y_onehot * torch.log(pred + 1e-8), dim=1))


def regression_loss(pred, y):
    """train.py:107,125."""
    return torch.mean(torch.mean((pred[:, :, 0] - y) ** 2, dim=1))
