"""``LogLinearCDE`` (D / BD / DE-LNCDE) – host-side mirror of
models/LinearNeuralCDEs.py:55-396 over kernels K2 (flow build, scan, reverse scan, bracket
gradient GEMM).  ``log_ode`` (:192-232) is a polynomial in the parameters only and stays in
the host autograd; the kernel boundary is (lie, logsig, y0) -> ys.

The diagonal+dense, Walsh-Hadamard, DPLR and sparse variants are outside the hot path
(SURVEY.md §8 row f3)."""
from __future__ import annotations

import math

import torch

from .. import _lib
from ..data_dir.hall_set import HallSet
from .LogNeuralCDEs import _Linear


class _LinearScan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lie, logsig, y0, dims):
        nb, bs, n_basis = dims
        _lib.require_cuda(logsig)
        L = _lib.lib()
        B, T, D = logsig.shape
        H = nb * bs
        lie = lie.contiguous()
        y0 = y0.contiguous()
        nbytes = L.lncde_linear_scan_workspace_bytes(B, T, nb, bs, n_basis)
        ws = torch.empty(nbytes, device=logsig.device, dtype=torch.uint8)  # kept for the backward pass
        ys = torch.empty((B, T, H), device=logsig.device, dtype=torch.float32)
        st = L.lncde_linear_scan_fwd(_lib.stream_ptr(), _lib.ptr(lie), _lib.ptr(logsig), _lib.ptr(y0), _lib.ptr(ys),
                                     B, T, D, nb, bs, n_basis, _lib.ptr(ws), nbytes)
        _lib.check(st, "lncde_linear_scan_fwd")
        ctx.save_for_backward(lie, logsig, ys, ws)
        ctx.dims = dims
        return ys

    @staticmethod
    def backward(ctx, g_ys):
        lie, logsig, ys, ws = ctx.saved_tensors
        nb, bs, n_basis = ctx.dims
        L = _lib.lib()
        B, T, D = logsig.shape
        g_ys = g_ys.contiguous()
        g_lie = torch.empty_like(lie)
        g_y0 = torch.empty((B, nb * bs), device=ys.device, dtype=torch.float32)
        st = L.lncde_linear_scan_bwd(_lib.stream_ptr(), _lib.ptr(lie), _lib.ptr(logsig), _lib.ptr(ys), _lib.ptr(g_ys),
                                     _lib.ptr(g_lie), _lib.ptr(g_y0), B, T, D, nb, bs, n_basis, _lib.ptr(ws),
                                     ws.numel())
        _lib.check(st, "lncde_linear_scan_bwd")
        return g_lie, None, g_y0, None


class LogLinearCDE:
    lip2 = True
    nondeterministic = False
    stateful = False
    vf_A_sparse = None
    hadamard_matrix = None

    def __init__(self, *, data_dim, hidden_dim, label_dim, block_size, logsig_depth, lambd=1.0,
                 walsh_hadamard=False, w_init_std=0.25, parallel_steps=128, classification=True,
                 diagonal_dense=False, rank=0, sparsity=1.0, key, device="cuda"):
        if hidden_dim % block_size != 0:
            raise ValueError("hidden_dim must be divisible by block_size.")
        if walsh_hadamard or diagonal_dense or rank > 0 or sparsity < 1.0:
            raise NotImplementedError("only the D / BD / DE structures are on the B200 hot path")
        self.hidden_dim, self.block_size = int(hidden_dim), int(block_size)
        self.num_blocks = self.hidden_dim // self.block_size
        self.parallel_steps, self.logsig_depth = int(parallel_steps), int(logsig_depth)
        self.data_dim, self.label_dim = int(data_dim), int(label_dim)
        self.lambd, self.w_init_std, self.classification = float(lambd), float(w_init_std), bool(classification)
        self.device = torch.device(device)
        hs = HallSet(self.data_dim, self.logsig_depth)
        self.basis_list = hs.basis_list()  # what roughpy's lie_basis enumerates (:111-116)
        self._basis_pairs = hs.data
        self._levels = None
        C, H, bs, nb = self.data_dim, self.hidden_dim, self.block_size, self.num_blocks
        n_A = C * nb * bs * bs
        n_tot = n_A + H * (C + 1) + self.label_dim * (H + 1)
        gen = torch.Generator().manual_seed(int(key))
        flat = torch.empty(n_tot, dtype=torch.float32)
        flat[:n_A] = torch.randn(n_A, generator=gen) * self.w_init_std / math.sqrt(bs)  # (:156-160)
        off = n_A
        for i, o in ((C, H), (H, self.label_dim)):
            n = o * (i + 1)
            flat[off : off + n] = ((torch.rand(n, generator=gen, dtype=torch.float64) * 2 - 1) / math.sqrt(i)).float()
            off += n
        self.params = flat.to(self.device).requires_grad_(True)
        self.n_A = n_A
        self.init_layer = _Linear(self.params, n_A, C, H)
        self.out_layer = _Linear(self.params, n_A + H * (C + 1), H, self.label_dim)

    @property
    def vf_A(self):
        return self.params[: self.n_A].view(self.data_dim, self.num_blocks * self.block_size * self.block_size)

    def parameters(self):
        return [self.params]

    def log_ode(self, vfs):
        """vfs [C, nb, bs, bs] -> bracket matrices [nb, bs, bs, n_basis] (:192-232), batched over blocks.
        Level by level, every bracket of a level in ONE batched matmul pair:
        A_[u,v] = A_v A_u - A_u A_v  (:225-227)."""
        if self._levels is None:
            n = len(self._basis_pairs)
            deg = [0] * n
            for k in range(1, n):
                l, r = self._basis_pairs[k]
                deg[k] = 1 if l == 0 else deg[l] + deg[r]
            levels = []
            for d in range(2, max(deg[1:]) + 1):
                ks = [k for k in range(1, n) if deg[k] == d]
                li = torch.tensor([self._basis_pairs[k][0] - 1 for k in ks], device=self.device)
                ri = torch.tensor([self._basis_pairs[k][1] - 1 for k in ks], device=self.device)
                levels.append((li, ri))
            self._levels = levels
        A = vfs  # index k-1 <-> basis element k; the first C entries are the letters in order
        for li, ri in self._levels:
            Al, Ar = A.index_select(0, li), A.index_select(0, ri)
            A = torch.cat([A, torch.matmul(Ar, Al) - torch.matmul(Al, Ar)], dim=0)
        return A.permute(1, 2, 3, 0)

    def hidden_states(self, X):
        ts, logsigs, x0 = X
        _lib.require_cuda(logsigs)
        y0 = self.init_layer(x0)
        C, nb, bs = self.data_dim, self.num_blocks, self.block_size
        if bs == 1:
            lie = self.vf_A.t()                 # [H, C]: flows = 1 + logsig[:, 1:C+1] @ vf_A  (:303-305)
            n_basis = C
        else:
            lie = self.log_ode(self.vf_A.view(C, nb, bs, bs))
            n_basis = lie.shape[-1]
        return _LinearScan.apply(lie, logsigs.contiguous(), y0, (nb, bs, n_basis))

    def __call__(self, X):
        ys = self.hidden_states(X)
        if self.classification:
            return torch.softmax(self.out_layer(ys.mean(dim=1)), dim=-1)   # (:389-391)
        B, T, H = ys.shape
        return torch.tanh(self.out_layer(ys.reshape(B * T, H)).view(B, T, -1))  # (:392-394)
