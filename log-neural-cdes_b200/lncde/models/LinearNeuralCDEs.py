"""``LogLinearCDE`` (D / BD / DE-LNCDE) – host-side mirror of
models/LinearNeuralCDEs.py:55-396 over kernels K2 (flow build, scan, reverse scan, bracket
gradient GEMM).  ``log_ode`` (:192-232) is a polynomial in the parameters only and stays in
the host autograd; the kernel boundary is (lie, logsig, y0) -> ys.

The remaining SLiCE structures of the reference are parameterisations of the same
two scans and reuse the same kernels: Walsh-Hadamard (:319-329), diagonal-plus-low-rank (:322-325) and sparse
(:319-321) build dense H x H vector fields and run the DE path; diagonal+dense (:273-300) runs a diagonal scan
on the first H - bs state components and one dense bs x bs block on the rest.  The reference's two
special-cased recurrent steps (:243-268, depth 1 and parallel_steps 1) are the same linear maps
y -> (I + sum_c a_c A_c) y written without forming A_c; they take the DE path here."""
from __future__ import annotations

import math

import numpy as np
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
        ctx.flags = _lib.call_flags()  # the backward call must name the same kernels
        st = L.lncde_linear_scan_fwd(_lib.stream_ptr(), _lib.ptr(lie), _lib.ptr(logsig), _lib.ptr(y0), _lib.ptr(ys),
                                     B, T, D, nb, bs, n_basis, _lib.ptr(ws), nbytes, ctx.flags)
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
                                     ws.numel(), ctx.flags)
        _lib.check(st, "lncde_linear_scan_bwd")
        return g_lie, None, g_y0, None


class _DiagScan(torch.autograd.Function):
    """Fused diagonal LNCDE (K2d, lncde_dlncde_*): no flows in HBM.  vf_At is [H, C]."""

    @staticmethod
    def forward(ctx, vf_At, logsig, y0):
        _lib.require_cuda(logsig)
        L = _lib.lib()
        B, T, D = logsig.shape
        H, C = vf_At.shape
        vf_At, y0 = vf_At.contiguous(), y0.contiguous()
        ys = torch.empty((B, T, H), device=logsig.device, dtype=torch.float32)
        st = L.lncde_dlncde_fwd(_lib.stream_ptr(), _lib.ptr(vf_At), _lib.ptr(logsig), _lib.ptr(y0), _lib.ptr(ys), B, T, D,
                                H, C)
        _lib.check(st, "lncde_dlncde_fwd")
        ctx.save_for_backward(vf_At, logsig, ys)
        return ys

    @staticmethod
    def backward(ctx, g_ys):
        vf_At, logsig, ys = ctx.saved_tensors
        L = _lib.lib()
        B, T, D = logsig.shape
        H, C = vf_At.shape
        g_ys = g_ys.contiguous()
        nbytes = L.lncde_dlncde_bwd_workspace_bytes(B, H, C)
        ws = torch.empty(nbytes, device=ys.device, dtype=torch.uint8)
        g_A = torch.empty_like(vf_At)
        g_y0 = torch.empty((B, H), device=ys.device, dtype=torch.float32)
        st = L.lncde_dlncde_bwd(_lib.stream_ptr(), _lib.ptr(vf_At), _lib.ptr(logsig), _lib.ptr(ys), _lib.ptr(g_ys),
                                _lib.ptr(g_A), _lib.ptr(g_y0), B, T, D, H, C, _lib.ptr(ws), nbytes)
        _lib.check(st, "lncde_dlncde_bwd")
        return g_A, None, g_y0


class _LieBrackets(torch.autograd.Function):
    """log_ode as kernels (K2a, lncde_lie_brackets_*): vfs [C, nb, bs, bs] -> lie [nb, bs, bs, n_basis]."""

    @staticmethod
    def forward(ctx, vfs, pairs):
        _lib.require_cuda(vfs)
        L = _lib.lib()
        C, nb, bs, _ = vfs.shape
        n_basis = pairs.shape[0]
        vfs = vfs.contiguous()
        nbytes = L.lncde_lie_brackets_workspace_bytes(n_basis, nb, bs)
        ws = torch.empty(nbytes, device=vfs.device, dtype=torch.uint8)  # every A_k: kept for the backward pass
        lie = torch.empty((nb, bs, bs, n_basis), device=vfs.device, dtype=torch.float32)
        st = L.lncde_lie_brackets_fwd(_lib.stream_ptr(), _lib.ptr(vfs), pairs.ctypes.data, _lib.ptr(lie), C, n_basis, nb, bs,
                                      _lib.ptr(ws), nbytes)
        _lib.check(st, "lncde_lie_brackets_fwd")
        ctx.save_for_backward(ws)
        ctx.meta = (pairs, C, n_basis, nb, bs, nbytes)
        return lie

    @staticmethod
    def backward(ctx, g_lie):
        (ws,) = ctx.saved_tensors
        pairs, C, n_basis, nb, bs, nbytes = ctx.meta
        L = _lib.lib()
        g_lie = g_lie.contiguous()
        wsg = torch.empty(nbytes, device=g_lie.device, dtype=torch.uint8)
        g_vf = torch.empty((C, nb, bs, bs), device=g_lie.device, dtype=torch.float32)
        st = L.lncde_lie_brackets_bwd(_lib.stream_ptr(), _lib.ptr(g_lie), pairs.ctypes.data, _lib.ptr(g_vf), C, n_basis, nb, bs,
                                      _lib.ptr(ws), _lib.ptr(wsg), nbytes)
        _lib.check(st, "lncde_lie_brackets_bwd")
        return g_vf, None


def _diag_scan(vf_At, logsigs, y0):
    """Diagonal scan: the fused K2d kernels (no flows in HBM; measured on B200: as fast as flow build + scan at B = 32,
    2.2x faster at B = 2048), or flow build + scan under ``kernel_flags("generic")`` / for more than 8 channels."""
    H, C = vf_At.shape
    if C <= 8 and not (_lib.call_flags() & _lib.FLAG_GENERIC):
        return _DiagScan.apply(vf_At, logsigs, y0)
    return _LinearScan.apply(vf_At, logsigs, y0, (H, 1, C))


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
        self.hidden_dim, self.block_size = int(hidden_dim), int(block_size)
        self.num_blocks = self.hidden_dim // self.block_size
        self.parallel_steps, self.logsig_depth = int(parallel_steps), int(logsig_depth)
        self.data_dim, self.label_dim = int(data_dim), int(label_dim)
        self.lambd, self.w_init_std, self.classification = float(lambd), float(w_init_std), bool(classification)
        self.walsh_hadamard, self.diagonal_dense = bool(walsh_hadamard), bool(diagonal_dense)
        self.rank, self.sparsity = int(rank), float(sparsity)
        self.dense_size = 0
        self.device = torch.device(device)
        hs = HallSet(self.data_dim, self.logsig_depth)
        self.basis_list = hs.basis_list()  # what roughpy's lie_basis enumerates (:111-116)
        self._basis_pairs = hs.data
        self._levels = None
        self._pairs_np = None
        C, H, bs, nb = self.data_dim, self.hidden_dim, self.block_size, self.num_blocks
        gen = torch.Generator().manual_seed(int(key))
        randn = lambda n: torch.randn(n, generator=gen)
        # flat parameter vector: [structure-specific vector-field blocks | init_layer | out_layer]
        self._sparse_idx = None
        if self.diagonal_dense:  # (:124-136): vf_A_diag [C, H - bs], vf_A_dense [C, bs, bs]; vf_A stays None
            blocks = [("vf_A_diag", (C, H - bs), randn(C * (H - bs)) * self.w_init_std),
                      ("vf_A_dense", (C, bs, bs), randn(C * bs * bs) * self.w_init_std / math.sqrt(bs))]
            self.dense_size = bs
        elif self.sparsity < 1.0:  # (:137-153): nnz values at fixed random positions of [C, H, H]
            total = C * H * H
            nnz = int(round(total * self.sparsity))
            if nnz == 0:
                raise ValueError("sparsity too low, no weights left!")
            self._sparse_idx = torch.randperm(total, generator=gen)[:nnz].sort().values.to(self.device)
            blocks = [("_sparse_values", (nnz,), randn(nnz) * self.w_init_std / math.sqrt(H))]
            self.block_size, self.num_blocks = H, 1
        else:  # (:154-177)
            blocks = [("vf_A", (C, nb * bs * bs), randn(C * nb * bs * bs) * self.w_init_std / math.sqrt(bs))]
            if self.rank > 0:
                if nb * bs * bs != H:
                    raise ValueError("the low-rank structure needs block_size 1 (vf_A holds the diagonals, :324)")
                r = self.rank
                blocks += [("vf_A_u", (C, H, r), randn(C * H * r) * self.w_init_std / math.sqrt(r)),
                           ("vf_A_v", (C, H, r), randn(C * H * r) * self.w_init_std / math.sqrt(r))]
                self.block_size, self.num_blocks = H, 1
        self.hadamard_matrix = None
        if self.walsh_hadamard:  # (:181-188)
            if self.diagonal_dense or self.rank > 0 or self.sparsity < 1.0 or nb * bs * bs != H:
                raise ValueError("walsh_hadamard needs block_size 1 and no other structure (:326-328)")
            from scipy.linalg import hadamard

            self.hadamard_matrix = (torch.from_numpy(hadamard(H)).float() / math.sqrt(H)).to(self.device)
            self.block_size, self.num_blocks = H, 1
        n_A = sum(v.numel() for _, _, v in blocks)
        n_tot = n_A + H * (C + 1) + self.label_dim * (H + 1)
        flat = torch.empty(n_tot, dtype=torch.float32)
        self._blocks, off = {}, 0
        for name, shape, v in blocks:
            flat[off : off + v.numel()] = v
            self._blocks[name] = (off, shape)
            off += v.numel()
        for i, o in ((C, H), (H, self.label_dim)):
            n = o * (i + 1)
            flat[off : off + n] = ((torch.rand(n, generator=gen, dtype=torch.float64) * 2 - 1) / math.sqrt(i)).float()
            off += n
        self.params = flat.to(self.device).requires_grad_(True)
        self.n_A = n_A
        self.init_layer = _Linear(self.params, n_A, C, H)
        self.out_layer = _Linear(self.params, n_A + H * (C + 1), H, self.label_dim)

    def _block(self, name):
        if name not in self._blocks:
            return None
        off, shape = self._blocks[name]
        n = 1
        for d in shape:
            n *= d
        return self.params[off : off + n].view(*shape)

    vf_A = property(lambda self: self._block("vf_A"))
    vf_A_diag = property(lambda self: self._block("vf_A_diag"))
    vf_A_dense = property(lambda self: self._block("vf_A_dense"))
    vf_A_u = property(lambda self: self._block("vf_A_u"))
    vf_A_v = property(lambda self: self._block("vf_A_v"))

    @property
    def vf_A_sparse(self):
        """Object with ``.todense()`` -> [C, H, H] (what train.py:89-90 reads), or None."""
        if self._sparse_idx is None:
            return None
        model = self

        class _View:
            @staticmethod
            def todense():
                C, H = model.data_dim, model.hidden_dim
                dense = torch.zeros(C * H * H, device=model.device, dtype=torch.float32)
                return dense.index_put((model._sparse_idx,), model._block("_sparse_values")).view(C, H, H)

        return _View()

    def parameters(self):
        return [self.params]

    def log_ode(self, vfs):
        """vfs [C, nb, bs, bs] -> bracket matrices [nb, bs, bs, n_basis] (:192-232), batched over blocks.
        Level by level, every bracket of a level in ONE batched matmul pair:
        A_[u,v] = A_v A_u - A_u A_v  (:225-227)."""
        if not (_lib.call_flags() & _lib.FLAG_GENERIC) and len(self._basis_pairs) - 1 <= 160:
            if self._pairs_np is None:  # HallSet.data[1:], 1-based (left, right), left = 0 for the letters
                self._pairs_np = np.ascontiguousarray(np.asarray(self._basis_pairs[1:], np.int32).reshape(-1, 2))
            return _LieBrackets.apply(vfs, self._pairs_np)
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
        C, H, nb, bs = self.data_dim, self.hidden_dim, self.num_blocks, self.block_size
        logsigs = logsigs.contiguous()
        if self.diagonal_dense:  # (:273-300): diagonal scan on the first H - bs components, one dense block on the rest
            nd = H - self.dense_size
            ys_d = _diag_scan(self.vf_A_diag.t(), logsigs, y0[:, :nd])
            lie = self.log_ode(self.vf_A_dense.unsqueeze(1))
            ys_e = _LinearScan.apply(lie, logsigs, y0[:, nd:], (1, self.dense_size, lie.shape[-1]))
            return torch.cat([ys_d, ys_e], dim=-1)
        if self._sparse_idx is not None:    # (:319-321)
            vfs = self.vf_A_sparse.todense().view(C, 1, H, H)
        elif self.rank > 0:                 # (:322-325): U_c V_c^T + diag(vf_A[c])
            vfs = (torch.einsum("cir,cjr->cij", self.vf_A_u, self.vf_A_v) + torch.diag_embed(self.vf_A)).unsqueeze(1)
        elif self.walsh_hadamard:           # (:326-328): vf_A[c, j] * Hm[i, j]
            vfs = (self.vf_A[:, None, :] * self.hadamard_matrix[None, :, :]).view(C, 1, H, H)
        elif bs == 1:
            # [H, C]: flows = 1 + logsig[:, 1:C+1] @ vf_A  (:303-305)
            return _diag_scan(self.vf_A.t(), logsigs, y0)
        else:
            vfs = self.vf_A.view(C, nb, bs, bs)
        lie = self.log_ode(vfs)
        return _LinearScan.apply(lie, logsigs, y0, (nb, bs, lie.shape[-1]))

    def fused_classification_loss(self, X, y):
        """One classification train step with the kernels of K2 / K5 only around the scans (initial state, time mean,
        read-out + cross-entropy, Lip(2) term, init_layer / out_layer gradients, bracket matrices K2a) - no torch
        autograd on the step (beyond 160 basis elements `log_ode` falls back to it).  Plain D / BD / DE structures;
        the other SLiCE parameterisations return None and take the autograd formulation."""
        if self.diagonal_dense or self._sparse_idx is not None or self.rank > 0 or self.walsh_hadamard:
            return None
        ts, logsigs, x0 = X
        _lib.require_cuda(logsigs)
        _lib.require_cuda_device(x0, y)
        L = _lib.lib()
        C, H, nb, bs, nl = self.data_dim, self.hidden_dim, self.num_blocks, self.block_size, self.label_dim
        ls, x0c, yc = logsigs.contiguous(), x0.contiguous(), y.contiguous()
        B, T, D = ls.shape
        dev = ls.device
        p = self.params.detach()
        nA, o2 = self.n_A, self.n_A + H * (C + 1)  # [vf_A | init_layer W, b | out_layer W, b]
        n_par = p.numel()
        out = torch.empty(n_par + 1, device=dev)  # flat gradient, then the loss (where the read-out kernel puts it)
        grads = out[:n_par]
        s = _lib.stream_ptr()
        y0 = torch.empty((B, H), device=dev)
        _lib.check(L.lncde_affine_fwd(s, _lib.ptr(x0c), p.data_ptr() + 4 * nA, p.data_ptr() + 4 * (nA + H * C), _lib.ptr(y0),
                                      B, C, H), "lncde_affine_fwd")
        ys = torch.empty((B, T, H), device=dev)
        diag = bs == 1 and C <= 8
        flags = _lib.call_flags()
        hmean, g_mean = torch.empty((B, H), device=dev), torch.empty((B, H), device=dev)
        if diag:  # the read-out only sees the time mean: the fused kernels form it / take its cotangent directly
            vf_At = p[:nA].view(C, H).t().contiguous()
            _lib.check(L.lncde_dlncde_fwd_mean(s, _lib.ptr(vf_At), _lib.ptr(ls), _lib.ptr(y0), _lib.ptr(ys), _lib.ptr(hmean),
                                               B, T, D, H, C), "lncde_dlncde_fwd_mean")
        else:
            n_basis = len(self._basis_pairs) - 1
            br_kernels = n_basis <= 160
            if br_kernels:  # K2a: bracket matrices straight from the flat parameters, every A_k kept for the gradient
                if self._pairs_np is None:
                    self._pairs_np = np.ascontiguousarray(np.asarray(self._basis_pairs[1:], np.int32).reshape(-1, 2))
                br_nb = L.lncde_lie_brackets_workspace_bytes(n_basis, nb, bs)
                br_ws = torch.empty(2 * br_nb, device=dev, dtype=torch.uint8)
                lie = torch.empty((nb, bs, bs, n_basis), device=dev)
                _lib.check(L.lncde_lie_brackets_fwd(s, _lib.ptr(p), self._pairs_np.ctypes.data, _lib.ptr(lie), C, n_basis, nb, bs,
                                                    _lib.ptr(br_ws), br_nb), "lncde_lie_brackets_fwd")
            else:
                with torch.enable_grad():
                    lie_t = self.log_ode(self.vf_A.view(C, nb, bs, bs))
                lie = lie_t.detach().contiguous()
            nbytes = L.lncde_linear_scan_workspace_bytes(B, T, nb, bs, n_basis)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            _lib.check(L.lncde_linear_scan_fwd(s, _lib.ptr(lie), _lib.ptr(ls), _lib.ptr(y0), _lib.ptr(ys), B, T, D, nb, bs,
                                               n_basis, _lib.ptr(ws), nbytes, flags), "lncde_linear_scan_fwd")
        if not diag:
            _lib.check(L.lncde_time_mean_fwd(s, _lib.ptr(ys), _lib.ptr(hmean), B, T, H), "lncde_time_mean_fwd")
        nb5 = max(L.lncde_readout_ce_workspace_bytes(B, H, nl), L.lncde_affine_bwd_workspace_bytes(B, C, H), 16)
        ws5 = torch.empty(nb5, device=dev, dtype=torch.uint8)
        _lib.check(L.lncde_readout_ce(s, _lib.ptr(hmean), H, p.data_ptr() + 4 * o2, p.data_ptr() + 4 * (o2 + nl * H),
                                      _lib.ptr(yc), None, _lib.ptr(g_mean), H, out.data_ptr() + 4 * o2, B, H, nl,
                                      _lib.ptr(ws5), nb5), "lncde_readout_ce")
        g_y0 = torch.empty((B, H), device=dev)
        if diag:
            nbytes = L.lncde_dlncde_bwd_workspace_bytes(B, H, C)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            g_At = torch.empty_like(vf_At)
            _lib.check(L.lncde_dlncde_bwd_mean(s, _lib.ptr(vf_At), _lib.ptr(ls), _lib.ptr(ys), _lib.ptr(g_mean),
                                               _lib.ptr(g_At), _lib.ptr(g_y0), B, T, D, H, C, _lib.ptr(ws), nbytes),
                       "lncde_dlncde_bwd_mean")
            grads[:nA].view(C, H).copy_(g_At.t())
        else:
            g_ys = torch.empty((B, T, H), device=dev)
            _lib.check(L.lncde_time_mean_bwd(s, _lib.ptr(g_mean), _lib.ptr(g_ys), B, T, H), "lncde_time_mean_bwd")
            g_lie = torch.empty_like(lie)
            _lib.check(L.lncde_linear_scan_bwd(s, _lib.ptr(lie), _lib.ptr(ls), _lib.ptr(ys), _lib.ptr(g_ys), _lib.ptr(g_lie),
                                               _lib.ptr(g_y0), B, T, D, nb, bs, n_basis, _lib.ptr(ws), nbytes, flags),
                       "lncde_linear_scan_bwd")
            if br_kernels:
                _lib.check(L.lncde_lie_brackets_bwd(s, _lib.ptr(g_lie), self._pairs_np.ctypes.data, _lib.ptr(grads), C, n_basis, nb,
                                                    bs, _lib.ptr(br_ws), br_ws.data_ptr() + br_nb, br_nb), "lncde_lie_brackets_bwd")
            else:
                (g_p,) = torch.autograd.grad(lie_t, self.params, g_lie)
                grads[:nA].copy_(g_p[:nA])
        _lib.check(L.lncde_affine_bwd(s, _lib.ptr(x0c), _lib.ptr(g_y0), out.data_ptr() + 4 * nA, B, C, H, _lib.ptr(ws5), nb5),
                   "lncde_affine_bwd")
        if self.lip2 and self.lambd != 0.0:  # lambd * mean_rows ||vf_A[row, :]||  (train.py:87-88)
            import ctypes

            one = lambda v: (ctypes.c_int32 * 1)(v)
            _lib.check(L.lncde_lip2(s, _lib.ptr(p), out.data_ptr() + 4 * n_par, _lib.ptr(grads), 1, one(0), one(-1), one(C),
                                    one(nA // C), float(self.lambd)), "lncde_lip2")
        return out[n_par], grads

    def __call__(self, X):
        ys = self.hidden_states(X)
        if self.classification:
            return torch.softmax(self.out_layer(ys.mean(dim=1)), dim=-1)   # (:389-391)
        B, T, H = ys.shape
        return torch.tanh(self.out_layer(ys.reshape(B * T, H)).view(B, T, -1))  # (:392-394)
