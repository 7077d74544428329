"""numpy/ctypes binding of the HOST-EMULATED kernel library (tests only, see cuda_emu.h).

"Device pointers" are plain host pointers here; every wrapper mirrors the call the product
binding (lncde/*.py over liblncde_b200.so) makes, so the same kernel source is exercised."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_emu  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_emu.build())
        i, u, vp, sz = C.c_int, C.c_uint, C.c_void_p, C.c_size_t
        _lib.lncde_logsig_fwd.argtypes = [vp, vp, vp, i, i, i, i, i, u, i, vp, vp, vp]
        _lib.lncde_hall_size.argtypes = [i, i]
        _lib.lncde_hall_t2l_nnz.argtypes = [i, i]
        _lib.lncde_hall_t2l_csr.argtypes = [i, i, vp, vp, vp]
        _lib.lncde_logsig_num_windows.argtypes = [i, i]
        _lib.lncde_logncde_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp, sz, u]
        _lib.lncde_logncde_bwd_workspace_bytes.argtypes = [i, i, i, i, i, i, i]
        _lib.lncde_logncde_bwd_workspace_bytes.restype = sz
        _lib.lncde_logncde_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp, sz, u]
        _lib.lncde_linear_scan_workspace_bytes.argtypes = [i, i, i, i, i]
        _lib.lncde_linear_scan_workspace_bytes.restype = sz
        _lib.lncde_linear_scan_fwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, sz, u]
        _lib.lncde_linear_scan_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, sz, u]
        _lib.lncde_adam_step.argtypes = [vp, vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, i,
                                         C.c_float]
        _lib.lncde_dlncde_bwd_workspace_bytes.argtypes = [i, i, i]
        _lib.lncde_dlncde_bwd_workspace_bytes.restype = sz
        _lib.lncde_dlncde_fwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i]
        _lib.lncde_dlncde_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp, sz]
        _lib.lncde_dlncde_fwd_mean.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i]
        _lib.lncde_dlncde_bwd_mean.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp, sz]
        _lib.lncde_nrde_param_count.argtypes = [i, i, i, i]
        _lib.lncde_nrde_fwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i]
        _lib.lncde_nrde_bwd_workspace_bytes.argtypes = [i, i, i, i, i, i]
        _lib.lncde_nrde_bwd_workspace_bytes.restype = sz
        _lib.lncde_nrde_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp, sz]
        _lib.emu_set_schedule.argtypes = [i, u, i]
        _lib.emu_bulk_copies.restype = C.c_long
        _lib.emu_peer_stores.restype = C.c_long
        _lib.emu_guarded_alloc.restype = C.c_void_p
        _lib.emu_guarded_alloc.argtypes = [C.c_size_t, C.c_int, C.c_int]
    return _lib


# Kernel selection of the emulated calls.  The library takes a per-call flags word (include/lncde_b200.h); the tests
# were written as "baseline kernels vs one variant", so this module keeps a small selection state whose ZERO values
# mean the BASELINE kernels (LNCDE_FLAG_GENERIC) and translates it into the flags of each call:
#   "k3_v3"     1: six-barrier Log-NCDE kernels (LNCDE_FLAG_K3_V3)
#   "k3_stream" 1: streamed global-weight mat-vecs of the generic Log-NCDE kernels (the library default)
#   "k2_pscan"  1: time-parallel D / BD scans (the library default for block sizes <= 8)
#   "k2_de_v2"  2: dense model on mma.sync tensor cores (LNCDE_FLAG_K2_MMA_SYNC), 3: tcgen05 flow build (library default)
#   "k2_cluster" 1: sweeps of the dense model on the cluster kernels (library default; a cluster of one CTA here)
FLAG_COMPAT, FLAG_SAVED_FWD, FLAG_K1_TMA, FLAG_GENERIC, FLAG_K2_MMA_SYNC, FLAG_K3_V3, FLAG_K2_ONE_CTA = 1, 2, 4, 8, 16, 32, 64
_sel = {"k3_v3": 0, "k3_stream": 0, "k2_pscan": 0, "k2_de_v2": 0, "k2_cluster": 0}


def set_tuning(key, value):
    assert key in _sel, key
    _sel[key] = int(value)


def _k3_flags():
    return (FLAG_K3_V3 if _sel["k3_v3"] else 0) | (0 if _sel["k3_stream"] else FLAG_GENERIC)


def _k2_flags(bs):
    if bs >= 64:
        return {0: FLAG_GENERIC, 2: FLAG_K2_MMA_SYNC, 3: 0}[_sel["k2_de_v2"]] | (0 if _sel["k2_cluster"] else FLAG_K2_ONE_CTA)
    return 0 if _sel["k2_pscan"] else FLAG_GENERIC


def set_schedule(mode="ascending", seed=1, blocks_reversed=False):
    """Order in which the emulator runs the threads of a CTA between synchronisation points ("ascending",
    "descending", "random", "warp_greedy", "warp_greedy_desc" - one warp at a time runs as far as it can) and the CTAs of a grid.  Race-free kernels give identical results under every schedule."""
    lib().emu_set_schedule({"ascending": 0, "descending": 1, "random": 2, "warp_greedy": 3, "warp_greedy_desc": 4}[mode], int(seed), 1 if blocks_reversed else 0)


def set_async_mode(lazy: bool):
    """Asynchronous copies deliver at issue (False, default) or at the wait the programming model requires (True)."""
    lib().emu_set_async_mode(1 if lazy else 0)


def bulk_copies():
    """Bulk (TMA) copies issued by the emulated kernels since the last call."""
    return lib().emu_bulk_copies()


def peer_stores():
    """One-sided cluster stores (st.async) issued by the emulated kernels since the last call."""
    return lib().emu_peer_stores()


def guarded(shape, dtype=np.float32, tight_end=True, start_mod16=0):
    """numpy array over a "device" allocation bounded by guard pages (never freed: tests are short)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib().emu_guarded_alloc(max(n, 1), 1 if tight_end else 0, start_mod16)
    assert ptr
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def _p(a):
    return C.c_void_p(0 if a is None else a.ctypes.data)


def _aligned(nbytes, align=256):
    """Workspace as torch.empty hands it out on the GPU: aligned, contents undefined - filled with the NaN pattern so
    that a kernel reading workspace it has not written shows up in the result."""
    raw = np.full(nbytes + align, 0xFF, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off : off + nbytes]


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _dev(a, dtype=np.float32):
    """Copy of `a` in a "device" allocation that ENDS at a guard page (reads or writes past the end fault)."""
    a = np.ascontiguousarray(a, dtype=dtype)
    g = guarded(a.shape, dtype)
    g[...] = a
    return g


def _dev_out(shape, dtype=np.float32):
    g = guarded(shape, dtype)
    g.view(np.uint8)[...] = 0xFF  # NaN pattern: an element the kernel does not write is visible
    return g


def calc_paths(x, stepsize, depth, compat=True, chunk=None, tma=False, misalign=0, tight_end=True):
    """tma: False (classic kernel), True (bulk-copy kernel)."""
    L = lib()
    x = f32(x)
    B, Ln, Cn = x.shape
    K = L.lncde_logsig_num_windows(Ln, int(stepsize))
    D = L.lncde_hall_size(Cn, depth)
    nnz = L.lncde_hall_t2l_nnz(Cn, depth)
    rowptr = np.zeros(D + 1, np.int32)
    cols = np.zeros(nnz, np.int32)
    vals = np.zeros(nnz, np.float32)
    assert L.lncde_hall_t2l_csr(Cn, depth, _p(rowptr), _p(cols), _p(vals)) == 0
    # input and output live in guard-page bounded allocations: out-of-bounds global accesses fault
    out = guarded((B, K, D), tight_end=tight_end, start_mod16=0 if tight_end else 4 * (misalign % 4))
    out[...] = np.nan
    xin = guarded(x.shape, tight_end=tight_end, start_mod16=4 * (misalign % 4))
    xin[...] = x
    st = L.lncde_logsig_fwd(None, _p(xin), _p(out), B, Ln, Cn, int(stepsize), depth,
                            (FLAG_COMPAT if compat else 0) | (FLAG_K1_TMA if tma else FLAG_GENERIC), int(chunk or B),
                            _p(rowptr), _p(cols), _p(vals))
    assert st == 0, st
    return out.copy()


def logncde_fwd_bwd(flat, logsig, y0, rows, scale, dims, g_ys, train=True):
    """lncde_logncde_fwd (+ bwd with the saved-forward workspace when train)."""
    L = lib()
    H, Cn, width, n_hidden, depth = dims
    flat, logsig, y0, scale = _dev(flat), _dev(logsig), _dev(y0), _dev(scale)
    rows = _dev(rows, np.int32)
    B, K, D = logsig.shape
    n_steps = rows.shape[0]  # rows / scale are [n_steps, 2]
    ys = _dev_out((B, n_steps + 1, H))
    nbytes = L.lncde_logncde_bwd_workspace_bytes(B, H, Cn, width, n_hidden, depth, n_steps) if train else 0
    ws = None
    if train:
        ws = guarded((nbytes,), np.uint8)
        ws[...] = 0xFF
    st = L.lncde_logncde_fwd(None, _p(flat), _p(logsig), _p(y0), _p(rows), _p(scale), _p(ys), B, K, D, H, Cn, width,
                             n_hidden, depth, n_steps, _p(ws), nbytes, _k3_flags())
    assert st == 0, st
    if not train:
        return ys.copy(), None, None
    g_ys = _dev(g_ys)
    g_params = _dev_out(flat.shape)
    g_y0 = _dev_out((B, H))
    st = L.lncde_logncde_bwd(None, _p(flat), _p(logsig), _p(rows), _p(scale), _p(ys), _p(g_ys), _p(g_params),
                             _p(g_y0), B, K, D, H, Cn, width, n_hidden, depth, n_steps, _p(ws), nbytes,
                             FLAG_SAVED_FWD | _k3_flags())
    assert st == 0, st
    return ys.copy(), g_params.copy(), g_y0.copy()


def logncde_bwd_recompute(flat, logsig, ys, rows, scale, dims, g_ys):
    """lncde_logncde_bwd without LNCDE_FLAG_SAVED_FWD (recomputing adjoint)."""
    L = lib()
    H, Cn, width, n_hidden, depth = dims
    flat, logsig, ys, scale, g_ys = _dev(flat), _dev(logsig), _dev(ys), _dev(scale), _dev(g_ys)
    rows = _dev(rows, np.int32)
    B, K, D = logsig.shape
    n_steps = rows.shape[0]
    nbytes = L.lncde_logncde_bwd_workspace_bytes(B, H, Cn, width, n_hidden, depth, n_steps)
    ws = guarded((nbytes,), np.uint8)
    ws[...] = 0xFF
    g_params = _dev_out(flat.shape)
    g_y0 = _dev_out((B, H))
    st = L.lncde_logncde_bwd(None, _p(flat), _p(logsig), _p(rows), _p(scale), _p(ys), _p(g_ys), _p(g_params),
                             _p(g_y0), B, K, D, H, Cn, width, n_hidden, depth, n_steps, _p(ws), nbytes, _k3_flags())
    assert st == 0, st
    return g_params.copy(), g_y0.copy()


def nrde_fwd_bwd(flat, logsig, y0, rows, scale, dims, g_ys=None):
    """lncde_nrde_fwd (+ bwd).  dims = (H, width, n_vf); parameter, output and gradient buffers end at guard pages."""
    L = lib()
    H, width, n_vf = dims
    logsig, scale = f32(logsig), f32(scale)
    rows = np.ascontiguousarray(rows, np.int32)
    B, K, D = logsig.shape
    n_steps = rows.shape[0]
    assert L.lncde_nrde_param_count(H, D - 1, width, n_vf) == np.asarray(flat).size
    fl = guarded(np.asarray(flat).shape)
    fl[...] = flat
    y0g = guarded((B, H))
    y0g[...] = y0
    ys = guarded((B, n_steps + 1, H))
    ys[...] = np.nan
    st = L.lncde_nrde_fwd(None, _p(fl), _p(logsig), _p(y0g), _p(rows), _p(scale), _p(ys), B, K, D, H, width, n_vf, n_steps)
    assert st == 0, st
    if g_ys is None:
        return ys.copy(), None, None
    gy = guarded((B, n_steps + 1, H))
    gy[...] = g_ys
    nbytes = L.lncde_nrde_bwd_workspace_bytes(B, H, D - 1, width, n_vf, n_steps)
    assert nbytes > 0
    ws = guarded((nbytes,), np.uint8)
    ws[...] = 0xFF  # NaN pattern: anything read before it is written shows up
    g_params = guarded(fl.shape)
    g_params[...] = np.nan
    g_y0 = guarded((B, H))
    g_y0[...] = np.nan
    st = L.lncde_nrde_bwd(None, _p(fl), _p(logsig), _p(rows), _p(scale), _p(ys), _p(gy), _p(g_params), _p(g_y0), B, K, D, H,
                          width, n_vf, n_steps, _p(ws), nbytes)
    assert st == 0, st
    return ys.copy(), g_params.copy(), g_y0.copy()


def linear_scan(lie, logsig, y0, nb, bs, g_ys=None):
    """lncde_linear_scan_fwd (+ bwd).  lie [nb,bs,bs,n_basis]."""
    L = lib()
    lie, logsig, y0 = _dev(lie), _dev(logsig), _dev(y0)
    B, T, D = logsig.shape
    n_basis = lie.shape[-1]
    H = nb * bs
    nbytes = L.lncde_linear_scan_workspace_bytes(B, T, nb, bs, n_basis)
    ws = guarded((nbytes,), np.uint8)
    ws[...] = 0xFF
    ys = _dev_out((B, T, H))
    st = L.lncde_linear_scan_fwd(None, _p(lie), _p(logsig), _p(y0), _p(ys), B, T, D, nb, bs, n_basis, _p(ws), nbytes,
                                 _k2_flags(bs))
    assert st == 0, st
    if g_ys is None:
        return ys.copy(), None, None
    g_ys = _dev(g_ys)
    g_lie = _dev_out(lie.shape)
    g_y0 = _dev_out((B, H))
    st = L.lncde_linear_scan_bwd(None, _p(lie), _p(logsig), _p(ys), _p(g_ys), _p(g_lie), _p(g_y0), B, T, D, nb, bs,
                                 n_basis, _p(ws), nbytes, _k2_flags(bs))
    assert st == 0, st
    return ys.copy(), g_lie.copy(), g_y0.copy()


def dlncde_mean(vf_At, logsig, y0, g_mean):
    """lncde_dlncde_fwd_mean + lncde_dlncde_bwd_mean: the fused diagonal LNCDE behind a mean-over-time read-out."""
    L = lib()
    H, Cn = vf_At.shape
    B, T, D = logsig.shape

    def dev(a):
        g = guarded(a.shape)
        g[...] = a
        return g

    A, ls, y0d, gm = dev(f32(vf_At)), dev(f32(logsig)), dev(f32(y0)), dev(f32(g_mean))
    ys, ym = guarded((B, T, H)), guarded((B, H))
    ys[...] = np.nan
    ym[...] = np.nan
    st = L.lncde_dlncde_fwd_mean(None, _p(A), _p(ls), _p(y0d), _p(ys), _p(ym), B, T, D, H, Cn)
    assert st == 0, st
    nbytes = L.lncde_dlncde_bwd_workspace_bytes(B, H, Cn)
    ws = guarded((nbytes // 4,))
    g_A, g_y0 = guarded((H, Cn)), guarded((B, H))
    g_A[...] = np.nan
    g_y0[...] = np.nan
    st = L.lncde_dlncde_bwd_mean(None, _p(A), _p(ls), _p(ys), _p(gm), _p(g_A), _p(g_y0), B, T, D, H, Cn, _p(ws), nbytes)
    assert st == 0, st
    return ys.copy(), ym.copy(), g_A.copy(), g_y0.copy()


def dlncde(vf_At, logsig, y0, g_ys=None):
    """lncde_dlncde_fwd (+ bwd): fused diagonal LNCDE.  vf_At [H, C]; buffers end at guard pages."""
    L = lib()
    H, Cn = vf_At.shape
    B, T, D = logsig.shape

    def dev(a):
        g = guarded(a.shape)
        g[...] = a
        return g

    A, ls, y0d = dev(f32(vf_At)), dev(f32(logsig)), dev(f32(y0))
    ys = guarded((B, T, H))
    ys[...] = np.nan
    st = L.lncde_dlncde_fwd(None, _p(A), _p(ls), _p(y0d), _p(ys), B, T, D, H, Cn)
    assert st == 0, st
    if g_ys is None:
        return ys.copy(), None, None
    g = dev(f32(g_ys))
    nbytes = L.lncde_dlncde_bwd_workspace_bytes(B, H, Cn)
    ws = guarded((nbytes // 4,))
    g_A, g_y0 = guarded((H, Cn)), guarded((B, H))
    g_A[...] = np.nan
    g_y0[...] = np.nan
    st = L.lncde_dlncde_bwd(None, _p(A), _p(ls), _p(ys), _p(g), _p(g_A), _p(g_y0), B, T, D, H, Cn, _p(ws), nbytes)
    assert st == 0, st
    return ys.copy(), g_A.copy(), g_y0.copy()
