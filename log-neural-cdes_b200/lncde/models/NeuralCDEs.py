"""``NeuralRDE`` – host-side mirror of models/NeuralCDEs.py:168-266 over kernel K3n (csrc/nrde.cu):
the reference's comparison baseline on the Log-ODE input ``(ts, logsig, x0)``.

Attribute names follow the reference pytree (``vf.mlp.layers[i].weight/.bias``, ``mlp_linear``, ``linear1``,
``linear2``, ``intervals``, ``logsig_dim`` (= D - 1, NeuralCDEs.py:208), ``classification``, ``lip2 = False`` ...).
``NeuralCDE`` (:85-166, cubic-interpolation control) is a different input pipeline and is not built.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib
from .LogNeuralCDEs import VectorField, _Linear, _workspace
from .solver_grid import stage_table


class _NRDESolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, logsig, y0, rows, scale, dims):
        H, width, n_vf = dims
        _lib.require_cuda(flat, logsig, y0, rows, scale)
        L = _lib.lib()
        B, K, D = logsig.shape
        n_steps = rows.shape[0]
        ys = torch.empty((B, n_steps + 1, H), device=y0.device, dtype=torch.float32)
        flat, y0c = flat.contiguous(), y0.contiguous()
        st = L.lncde_nrde_fwd(_lib.stream_ptr(), _lib.ptr(flat), _lib.ptr(logsig), _lib.ptr(y0c), _lib.ptr(rows),
                              _lib.ptr(scale), _lib.ptr(ys), B, K, D, H, width, n_vf, n_steps)
        _lib.check(st, "lncde_nrde_fwd")
        ctx.save_for_backward(flat, logsig, rows, scale, ys)
        ctx.dims = dims
        return ys

    @staticmethod
    def backward(ctx, g_ys):
        flat, logsig, rows, scale, ys = ctx.saved_tensors
        H, width, n_vf = ctx.dims
        L = _lib.lib()
        B, K, D = logsig.shape
        n_steps = rows.shape[0]
        g_ys = g_ys.contiguous()
        nbytes = L.lncde_nrde_bwd_workspace_bytes(B, H, D - 1, width, n_vf, n_steps)
        if nbytes == 0:
            raise _lib.LncdeError("lncde_nrde_bwd_workspace_bytes: configuration outside the compiled range")
        ws = _workspace(nbytes, ys.device)
        g_params = torch.empty_like(flat)
        g_y0 = torch.empty((B, H), device=ys.device, dtype=torch.float32)
        st = L.lncde_nrde_bwd(_lib.stream_ptr(), _lib.ptr(flat), _lib.ptr(logsig), _lib.ptr(rows), _lib.ptr(scale),
                              _lib.ptr(ys), _lib.ptr(g_ys), _lib.ptr(g_params), _lib.ptr(g_y0), B, K, D, H, width, n_vf,
                              n_steps, _lib.ptr(ws), nbytes)
        _lib.check(st, "lncde_nrde_bwd")
        return g_params, None, g_y0, None, None, None


class NeuralRDE:
    stateful = False
    nondeterministic = False
    lip2 = False

    def __init__(self, vf_hidden_dim, vf_num_hidden, hidden_dim, data_dim, logsig_dim, label_dim, classification,
                 output_step, intervals, solver, stepsize_controller, dt0, max_steps, scale, *, key, device="cuda"):
        if int(vf_num_hidden) < 1:
            raise ValueError("NeuralRDE needs vf_num_hidden >= 1 (eqx.nn.MLP depth = vf_num_hidden - 1)")
        self.data_dim, self.hidden_dim = int(data_dim), int(hidden_dim)
        self.logsig_dim = int(logsig_dim) - 1  # first element is always zero (NeuralCDEs.py:207-208)
        self.vf_width, self.vf_depth = int(vf_hidden_dim), int(vf_num_hidden)
        self.label_dim = int(label_dim)
        self.classification, self.output_step = bool(classification), int(output_step)
        self.solver, self.stepsize_controller = solver, stepsize_controller
        self.dt0, self.max_steps = float(dt0), int(max_steps)
        self.device = torch.device(device)
        H, C, W, Dm = self.hidden_dim, self.data_dim, self.vf_width, self.logsig_dim
        vf_sizes = list(zip([H] + [W] * (self.vf_depth - 1), [W] * self.vf_depth))
        sizes = vf_sizes + [(W, H * Dm), (C, H), (H, self.label_dim)]
        n_tot = sum(o * (i + 1) for i, o in sizes)
        gen = torch.Generator().manual_seed(int(key))
        flat = torch.empty(n_tot, dtype=torch.float32)
        off = 0
        # eqx.nn.Linear: W, b ~ U(-1/sqrt(in), 1/sqrt(in)); only the vector field `vf` is divided by `scale`
        for li, (i, o) in enumerate(sizes):
            lim = 1.0 / math.sqrt(i)
            n = o * (i + 1)
            block = (torch.rand(n, generator=gen, dtype=torch.float64) * 2 - 1) * lim
            if li < len(vf_sizes):
                block = block / scale
            flat[off : off + n] = block.float()
            off += n
        self.params = flat.to(self.device).requires_grad_(True)
        layers, off = [], 0
        for i, o in vf_sizes:
            layers.append(_Linear(self.params, off, i, o))
            off += o * (i + 1)
        self.vf = VectorField(layers)
        self.mlp_linear = _Linear(self.params, off, W, H * Dm)
        off += H * Dm * (W + 1)
        self.n_field = off  # [vf | mlp_linear]: the flat buffer the K3n kernels take
        self.linear1 = _Linear(self.params, off, C, H)
        off += H * (C + 1)
        self.linear2 = _Linear(self.params, off, H, self.label_dim)
        self.intervals = torch.as_tensor(np.asarray(intervals, np.float32))
        self._table = None
        self._tables = {}
        self._table_ts = None

    def parameters(self):
        return [self.params]

    def field_flat(self):
        return self.params[: self.n_field]

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
                # then cast to float32 (NeuralCDEs.py:244-246)
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

    def hidden_states(self, X):
        ts, logsig, x0 = X
        _lib.require_cuda(logsig)
        _lib.require_cuda_device(x0)
        if logsig.shape[2] != self.logsig_dim + 1:
            raise ValueError(f"logsig has {logsig.shape[2]} columns, model was built for {self.logsig_dim + 1}")
        rows, scale, interp = self._stage_table(ts, logsig.shape[1])
        y0 = self.linear1(x0)
        dims = (self.hidden_dim, self.vf_width, self.vf_depth)
        ys = _NRDESolve.apply(self.field_flat(), logsig.contiguous(), y0, rows, scale, dims)
        return ys, interp

    def __call__(self, X):
        ys, interp = self.hidden_states(X)
        if self.classification:
            return torch.softmax(self.linear2(ys[:, -1]), dim=-1)  # NeuralCDEs.py:263-264
        idx, w = interp
        mid = ys[:, idx - 1] * (1 - w) + ys[:, idx] * w  # SaveAt(ts=...) local linear interpolation, plus t1
        out = torch.cat([mid, ys[:, -1:]], dim=1)
        B, n, H = out.shape
        return torch.tanh(self.linear2(out.reshape(B * n, H)).view(B, n, -1))  # :265-266
