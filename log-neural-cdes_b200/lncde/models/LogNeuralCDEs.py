"""``LogNeuralCDE`` – host-side mirror of models/LogNeuralCDEs.py:37-162 over kernels
K3 (forward Heun solve) and K3b (exact discrete adjoint).

Attribute names follow the reference pytree (``vf.mlp.layers[i].weight/.bias``,
``linear1``, ``linear2``, ``pairs``, ``intervals``, ``classification``, ``lip2``,
``lambd`` ...) so that train.py:79-93,441-458 semantics carry over.  The model is
called on the whole batch ``X = (ts[B,L], logsig[B,K,D], x0[B,C])`` – the
reference's per-sample ``__call__`` under ``jax.vmap`` (train.py:66).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib
from ..data_dir.hall_set import HallSet
from .solver_grid import stage_table


class _Linear:
    """View of W[out,in], b[out] inside a flat parameter vector (eqx.nn.Linear layout)."""

    def __init__(self, flat, off, n_in, n_out):
        self._flat, self._off, self.in_features, self.out_features = flat, off, n_in, n_out

    @property
    def weight(self):
        o = self._off
        return self._flat[o : o + self.out_features * self.in_features].view(self.out_features, self.in_features)

    @property
    def bias(self):
        o = self._off + self.out_features * self.in_features
        return self._flat[o : o + self.out_features]

    def __call__(self, x):
        return torch.addmm(self.bias, x, self.weight.t())

    @property
    def numel(self):
        return self.out_features * (self.in_features + 1)


class _MLP:
    def __init__(self, layers):
        self.layers = layers


class VectorField:
    """models/NeuralCDEs.py:43-82 – eqx.nn.MLP(H -> width x depth -> H*C), SiLU hidden,
    tanh final; evaluated inside the solve kernels, never on the host."""

    def __init__(self, layers):
        self.mlp = _MLP(layers)


def solve_fwd(mlp_flat, logsig, y0, rows, scale, dims, train):
    """lncde_logncde_fwd on the whole batch.  Returns (ys, token); `token` carries what solve_bwd needs to use the
    per-stage intermediates the training-mode forward left in the shared scratch."""
    H, C, width, n_hidden, depth = dims
    _lib.require_cuda(mlp_flat, logsig, y0, rows, scale)
    L = _lib.lib()
    B, K, D = logsig.shape
    n_steps = rows.shape[0]
    ys = torch.empty((B, n_steps + 1, H), device=y0.device, dtype=torch.float32)
    # training mode: hand the adjoint workspace to the forward so it can keep its intermediates
    nbytes = L.lncde_logncde_bwd_workspace_bytes(B, H, C, width, n_hidden, depth, n_steps) if train else 0
    ws = _workspace(nbytes, ys.device) if train else None  # bumps the generation: earlier saved forwards are void
    flags = _lib.call_flags()
    st = L.lncde_logncde_fwd(_lib.stream_ptr(), _lib.ptr(mlp_flat), _lib.ptr(logsig), _lib.ptr(y0),
                             _lib.ptr(rows), _lib.ptr(scale), _lib.ptr(ys), B, K, D, H, C, width, n_hidden,
                             depth, n_steps, _lib.ptr(ws), nbytes, flags)
    _lib.check(st, "lncde_logncde_fwd")
    # shared grow-only scratch: its contents are valid until the next hand-out of the buffer on this device
    return ys, (ws, _WS_GEN.get(ys.device, 0), flags)


def solve_bwd(mlp_flat, logsig, rows, scale, ys, g_ys, dims, token, g_params=None):
    """lncde_logncde_bwd: (g_params of the vector-field MLP [summed over the batch; written into `g_params` if given],
    g_y0)."""
    H, C, width, n_hidden, depth = dims
    L = _lib.lib()
    B, K, D = logsig.shape
    n_steps = rows.shape[0]
    nbytes = L.lncde_logncde_bwd_workspace_bytes(B, H, C, width, n_hidden, depth, n_steps)
    ws, ws_gen, flags = token
    # LNCDE_FLAG_SAVED_FWD only if nobody has been handed the scratch since our forward wrote it
    saved = _lib.FLAG_SAVED_FWD if (ws is not None and _WS.get(ys.device) is ws and _WS_GEN.get(ys.device) == ws_gen) else 0
    if ws is None or saved == 0:
        ws = _workspace(nbytes, ys.device)  # recompute mode overwrites the scratch: bumps the generation too
    if g_params is None:
        g_params = torch.empty_like(mlp_flat)
    g_y0 = torch.empty((B, H), device=ys.device, dtype=torch.float32)
    st = L.lncde_logncde_bwd(_lib.stream_ptr(), _lib.ptr(mlp_flat), _lib.ptr(logsig), _lib.ptr(rows),
                             _lib.ptr(scale), _lib.ptr(ys), _lib.ptr(g_ys), _lib.ptr(g_params), _lib.ptr(g_y0),
                             B, K, D, H, C, width, n_hidden, depth, n_steps, _lib.ptr(ws), nbytes,
                             saved | (flags & ~_lib.FLAG_SAVED_FWD))
    _lib.check(st, "lncde_logncde_bwd")
    if saved:  # the sweep wrote its (left, right) gradient pairs over the buffer: a second backward must recompute
        _WS_GEN[ys.device] = _WS_GEN.get(ys.device, 0) + 1
    return g_params, g_y0


class _LogNCDESolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mlp_flat, logsig, y0, rows, scale, dims):
        mlp_flat = mlp_flat.contiguous()
        ys, ctx.token = solve_fwd(mlp_flat, logsig, y0.contiguous(), rows, scale, dims, any(ctx.needs_input_grad))
        ctx.save_for_backward(mlp_flat, logsig, rows, scale, ys)
        ctx.dims = dims
        return ys

    @staticmethod
    def backward(ctx, g_ys):
        mlp_flat, logsig, rows, scale, ys = ctx.saved_tensors
        g_params, g_y0 = solve_bwd(mlp_flat, logsig, rows, scale, ys, g_ys.contiguous(), ctx.dims, ctx.token)
        return g_params, None, g_y0, None, None, None


_WS = {}
_WS_GEN = {}
_RETIRED = []


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per device (the C ABI never allocates).  Every hand-out is a WRITE claim and bumps the
    device's generation counter: a saved-forward adjoint (LNCDE_FLAG_SAVED_FWD) is only taken when no other solve -
    a later training forward, a recompute-mode adjoint, an NRDE adjoint - has been given the buffer since."""
    _WS_GEN[device] = _WS_GEN.get(device, 0) + 1
    cur = _WS.get(device)
    if cur is None or cur.numel() < nbytes:
        if cur is not None:
            _RETIRED.append(cur)  # a captured CUDA graph may still hold its address: never hand it back
        cur = torch.empty(max(nbytes, 1), device=device, dtype=torch.uint8)
        _WS[device] = cur
    return cur


class LogNeuralCDE:
    stateful = False
    nondeterministic = False
    lip2 = True

    def __init__(self, vf_hidden_dim, vf_num_hidden, hidden_dim, data_dim, depth, label_dim, classification,
                 output_step, intervals, solver, stepsize_controller, dt0, max_steps, scale, lambd, *, key,
                 device="cuda"):
        if depth not in (1, 2):
            raise ValueError("LogNeuralCDE supports logsig depth 1 and 2 (as the reference, LogNeuralCDEs.py:9)")
        self.data_dim, self.depth, self.hidden_dim = int(data_dim), int(depth), int(hidden_dim)
        self.vf_width, self.vf_depth = int(vf_hidden_dim), int(vf_num_hidden)
        self.label_dim = int(label_dim)
        self.classification, self.output_step = bool(classification), int(output_step)
        self.solver, self.stepsize_controller = solver, stepsize_controller
        self.dt0, self.max_steps, self.lambd = float(dt0), int(max_steps), float(lambd)
        self.device = torch.device(device)
        H, C, W = self.hidden_dim, self.data_dim, self.vf_width
        sizes = [H] + [W] * self.vf_depth + [C * H]
        n_mlp = sum(o * (i + 1) for i, o in zip(sizes[:-1], sizes[1:]))
        n_tot = n_mlp + H * (C + 1) + self.label_dim * (H + 1)
        gen = torch.Generator().manual_seed(int(key))
        flat = torch.empty(n_tot, dtype=torch.float32)
        off = 0
        # eqx.nn.Linear: W, b ~ U(-1/sqrt(in), 1/sqrt(in)); the vector field is then divided by `scale`
        for li, (i, o) in enumerate(list(zip(sizes[:-1], sizes[1:])) + [(C, H), (H, self.label_dim)]):
            lim = 1.0 / math.sqrt(i)
            n = o * (i + 1)
            block = (torch.rand(n, generator=gen, dtype=torch.float64) * 2 - 1) * lim
            if li < len(sizes) - 1:
                block = block / scale
            flat[off : off + n] = block.float()
            off += n
        self.params = flat.to(self.device).requires_grad_(True)
        self.n_mlp = n_mlp
        layers, off = [], 0
        for i, o in zip(sizes[:-1], sizes[1:]):
            layers.append(_Linear(self.params, off, i, o))
            off += o * (i + 1)
        self.vf = VectorField(layers)
        self.linear1 = _Linear(self.params, off, C, H)
        off += H * (C + 1)
        self.linear2 = _Linear(self.params, off, H, self.label_dim)
        hs = HallSet(C, self.depth)
        self.pairs = None if self.depth == 1 else torch.from_numpy(hs.pairs())
        self.logsig_dim = len(hs)
        self.intervals = torch.as_tensor(np.asarray(intervals, np.float32))
        self._table = None
        self._tables = {}
        self._table_ts = None

    # -- parameters -----------------------------------------------------------------
    def parameters(self):
        return [self.params]

    def mlp_flat(self):
        return self.params[: self.n_mlp]

    # -- stage table ----------------------------------------------------------------
    def _stage_table(self, ts, n_rows):
        # Two-level cache.  Level 1: the very tensor seen last time (pointer + shape; the tensor is kept referenced so its
        # allocation cannot be recycled for another grid) - no device read-back on the steady-state path.  Level 2: the
        # VALUES (t0, t1, L, n_rows) the table is a function of - a fresh `ts` tensor per batch costs one two-float
        # read-back, not a rebuild.
        key = (ts.data_ptr(), tuple(ts.shape), n_rows)
        if self._table is not None and self._table[0] == key:
            return self._table[1:]
        t0, t1 = (float(v) for v in ts[0, [0, -1]].tolist())
        vkey = (t0, t1, int(ts.shape[1]), n_rows)
        if vkey not in self._tables:
            rows, scale, grid = stage_table(t0, t1, self.intervals.numpy(), self.dt0, n_rows, self.max_steps)
            assert 0 <= int(rows.min()) and int(rows.max()) < n_rows, "stage table reads a logsig row that does not exist"
            interp = None
            if not self.classification:
                # jnp.arange(step, 1.0, step) with a Python-float step: the LENGTH is computed in float64, the values are
                # then cast to float32 (LogNeuralCDEs.py:143-145)
                step = self.output_step / ts.shape[1]
                times = np.arange(step, 1.0, step).astype(np.float32)
                idx = np.clip(np.searchsorted(grid, times, side="left"), 1, len(grid) - 1)
                w = ((times - grid[idx - 1]) / (grid[idx] - grid[idx - 1])).astype(np.float32)
                interp = (torch.from_numpy(idx).to(self.device), torch.from_numpy(w).to(self.device)[None, :, None])
            self._tables[vkey] = (torch.from_numpy(rows.reshape(-1)).to(self.device).view(-1, 2),
                                  torch.from_numpy(scale.reshape(-1)).to(self.device).view(-1, 2), interp)
        self._table = (key, *self._tables[vkey])
        self._table_ts = ts
        return self._table[1:]

    # -- forward ---------------------------------------------------------------------
    def hidden_states(self, X):
        ts, logsig, x0 = X
        _lib.require_cuda(logsig)
        _lib.require_cuda_device(x0)
        rows, scale, interp = self._stage_table(ts, logsig.shape[1])
        y0 = self.linear1(x0)
        dims = (self.hidden_dim, self.data_dim, self.vf_width, self.vf_depth, self.depth)
        ys = _LogNCDESolve.apply(self.mlp_flat(), logsig.contiguous(), y0, rows, scale, dims)
        return ys, interp

    def fused_classification_loss(self, X, y):
        """One classification train step WITHOUT torch autograd / eager glue: initial state, solve, read-out +
        cross-entropy, adjoint sweep, Lip(2) term and the gradients of linear1 / linear2 are all library kernels
        (K3, K3b, K5) writing into one flat gradient in the layout of ``self.params``.  Same value and gradient as
        ``train.classification_loss`` through autograd (tests/test_logncde_gpu.py, tests over the emulator)."""
        ts, logsig, x0 = X
        _lib.require_cuda(logsig)
        _lib.require_cuda_device(x0, y)
        L = _lib.lib()
        rows, scale, _ = self._stage_table(ts, logsig.shape[1])
        H, C, nl = self.hidden_dim, self.data_dim, self.label_dim
        B, n_steps = logsig.shape[0], rows.shape[0]
        dev = logsig.device
        p = self.params.detach()
        x0c, yc, ls = x0.contiguous(), y.contiguous(), logsig.contiguous()
        o1, o2 = self.n_mlp, self.n_mlp + H * (C + 1)  # [mlp | linear1 W, b | linear2 W, b]
        W1, b1, W2, b2 = p[o1 : o1 + H * C], p[o1 + H * C : o2], p[o2 : o2 + nl * H], p[o2 + nl * H :]
        n_par = p.numel()
        out = torch.empty(n_par + 1, device=dev)  # flat gradient, then the loss (where the read-out kernel puts it)
        grads = out[:n_par]
        key = (B, n_steps, dev)
        if getattr(self, "_fused_buf", (None,))[0] != key:
            nb = max(L.lncde_readout_ce_workspace_bytes(B, H, nl), L.lncde_affine_bwd_workspace_bytes(B, C, H), 16)
            # g_ys: zero except its last row, which every step overwrites - never cleared again
            self._fused_buf = (key, torch.zeros((B, n_steps + 1, H), device=dev), torch.empty(nb, dtype=torch.uint8, device=dev),
                               nb)
        _, g_ys, ws5, nb5 = self._fused_buf
        y0 = torch.empty((B, H), device=dev)
        s = _lib.stream_ptr()
        _lib.check(L.lncde_affine_fwd(s, _lib.ptr(x0c), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(y0), B, C, H), "lncde_affine_fwd")
        dims = (H, C, self.vf_width, self.vf_depth, self.depth)
        mlp = p[: self.n_mlp]
        ys, token = solve_fwd(mlp, ls, y0, rows, scale, dims, True)
        last = n_steps * H  # offset of y(t1) inside a sample's [n_steps + 1, H] block
        stride = (n_steps + 1) * H
        _lib.check(L.lncde_readout_ce(s, ys.data_ptr() + 4 * last, stride, _lib.ptr(W2), _lib.ptr(b2), _lib.ptr(yc), None,
                                      g_ys.data_ptr() + 4 * last, stride, out.data_ptr() + 4 * o2, B, H, nl, _lib.ptr(ws5),
                                      nb5), "lncde_readout_ce")
        _, g_y0 = solve_bwd(mlp, ls, rows, scale, ys, g_ys, dims, token, g_params=grads[: self.n_mlp])
        _lib.check(L.lncde_affine_bwd(s, _lib.ptr(x0c), _lib.ptr(g_y0), out.data_ptr() + 4 * o1, B, C, H, _lib.ptr(ws5), nb5),
                   "lncde_affine_bwd")
        if self.lip2 and self.lambd != 0.0:
            import ctypes

            lay = self.vf.mlp.layers
            arr = lambda v: (ctypes.c_int32 * len(v))(*v)
            w_off, b_off, off = [], [], 0
            for l in lay:
                w_off.append(off)
                b_off.append(off + l.out_features * l.in_features)
                off += l.numel
            _lib.check(L.lncde_lip2(s, _lib.ptr(p), out.data_ptr() + 4 * n_par, _lib.ptr(grads), len(lay), arr(w_off),
                                    arr(b_off), arr([l.out_features for l in lay]), arr([l.in_features for l in lay]),
                                    float(self.lambd)), "lncde_lip2")
        return out[n_par], grads

    def __call__(self, X):
        ys, interp = self.hidden_states(X)
        if self.classification:
            return torch.softmax(self.linear2(ys[:, -1]), dim=-1)  # LogNeuralCDEs.py:159-160
        idx, w = interp
        mid = ys[:, idx - 1] * (1 - w) + ys[:, idx] * w  # SaveAt(ts=...) local linear interpolation
        out = torch.cat([mid, ys[:, -1:]], dim=1)        # ... plus t1
        B, n, H = out.shape
        return torch.tanh(self.linear2(out.reshape(B * n, H)).view(B, n, -1))  # :161-162
