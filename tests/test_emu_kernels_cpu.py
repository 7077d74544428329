"""The CUDA kernel SOURCES executed on the CPU by the host emulator (tests/emu) and compared with the
oracle - so that `-m "not gpu"` checks the real kernel code, not only the host logic.  The GPU
parity tests (`-m gpu`, through liblncde_b200.so) remain the gate for the compiled product; this
file catches indexing / barrier / alignment mistakes before GPU time is spent on them."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import linear_ncde as OL
from oracle import log_ncde as O
from oracle import paths

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_lib as E  # noqa: E402

TOL = 1e-5


def _toy(B, L, C, seed=1234):
    rng = np.random.default_rng(seed)
    return np.cumsum(np.round(rng.normal(size=(B, L, C))), axis=1).astype(np.float32)


def _walk(B, L, C, seed=2345):
    rng = np.random.default_rng(seed)
    return np.cumsum(rng.normal(scale=0.05, size=(B, L, C)), axis=1).astype(np.float32)


# ------------------------------------------------------------------------------------------ K1
# tma=False: logsig_levy_kernel (classic staging); tma=True: logsig_levy_tma_kernel (bulk copies, one CTA per tile).
# Inputs and outputs sit in guard-page bounded allocations (tight at the end or at the start), so an
# out-of-bounds global access of either kernel faults instead of passing silently.

@pytest.mark.parametrize("tma", [False, True])
@pytest.mark.parametrize("C", [1, 2, 3, 6, 7, 8])
@pytest.mark.parametrize("depth", [1, 2])
def test_emu_logsig_integer_kat_bit_exact(C, depth, tma):
    x = _toy(5, 100, C)
    for s in (1, 4, 7, 12, 100, 101, 500):
        want = paths.calc_paths(x, s, depth, np.float64)
        for tight_end in (True, False):
            got = E.calc_paths(x, s, depth, tma=tma, tight_end=tight_end)
            assert got.shape == want.shape
            assert np.array_equal(got.astype(np.float64), want), (C, depth, s)


@pytest.mark.parametrize("tma", [False, True])
@pytest.mark.parametrize("compat", [True, False])
def test_emu_logsig_compat_and_chunk(compat, tma):
    x = _toy(140, 30, 3)
    want = paths.batch_calc_paths(x, 4, 2, np.float64, compat=compat)
    got = E.calc_paths(x, 4, 2, compat=compat, chunk=128, tma=tma)
    assert np.array_equal(got.astype(np.float64), want)


@pytest.mark.parametrize("tma", [False, True])
@pytest.mark.parametrize("C,s,L", [(7, 12, 1200), (6, 12, 1190), (6, 16, 896), (7, 10, 999), (4, 10, 503), (5, 3, 1000),
                                   (3, 7, 333), (8, 16, 700)])
def test_emu_logsig_float_paths_depth2(C, s, L, tma):
    x = _walk(3, L, C)
    want = paths.calc_paths(x, s, 2, np.float64)
    E.bulk_copies()
    got = E.calc_paths(x, s, 2, tma=tma)
    assert (E.bulk_copies() > 0) == bool(tma)  # the requested kernel is the one that ran
    assert got.shape == want.shape
    assert rel_err(got, want) < TOL


def test_emu_logsig_tma_equals_classic_every_alignment():
    """Same LevyState arithmetic: bit-identical results for every 16-byte phase of the tiles (odd L*C moves
    the phase from sample to sample), every remainder length and tile counts around 32 windows."""
    rng = np.random.default_rng(11)
    for C, L, s in [(7, 12 * 33 + 5, 12), (3, 97, 4), (5, 131, 4), (1, 70, 4), (7, 389, 12), (6, 150, 2), (8, 64, 1)]:
        x = np.cumsum(rng.normal(size=(5, L, C)), axis=1).astype(np.float32)
        for depth in (1, 2):
            a = E.calc_paths(x, s, depth, tma=False)
            b = E.calc_paths(x, s, depth, tma=True)
            assert np.array_equal(a, b), (C, L, s, depth)


def test_emu_logsig_tma_eigenworms_length():
    """One EigenWorms-length sample (47 tiles of 32 windows) through the bulk-copy kernel."""
    x = _walk(1, 17984, 7, seed=5)
    x[..., 0] = np.arange(17984) / 17984
    want = paths.calc_paths(x, 12, 2, np.float64)
    got = E.calc_paths(x, 12, 2, tma=True)
    assert got.shape == (1, 1499, 29)
    assert rel_err(got, want) < TOL
    assert np.array_equal(got, E.calc_paths(x, 12, 2, tma=False))


@pytest.mark.parametrize("C,depth", [(2, 3), (3, 3), (7, 3), (2, 4), (3, 4), (10, 2), (9, 1)])
def test_emu_logsig_tensor_kernels(C, depth):
    x = _toy(3, 50, C)
    for s in (4, 12, 51):
        want = paths.calc_paths(x, s, depth, np.float64)
        assert rel_err(E.calc_paths(x, s, depth), want) < TOL, (C, depth, s)
    xf = _walk(2, 200, C)
    assert rel_err(E.calc_paths(xf, 12, depth), paths.calc_paths(xf, 12, depth, np.float64)) < TOL


# ------------------------------------------------------------------------------------------ K3 / K3b
def _setup(B, L, C, s, depth, H, width, n_hidden, dt0, scale, seed=0):
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float64)
    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    rows, sc, grid = O.stage_table(ts, iv, dt0, logsig.shape[1])
    layers = O.init_params(H, C, width, n_hidden, scale=scale, seed=seed, dtype=torch.float64)
    y0 = torch.from_numpy(rng.normal(size=(B, H)) * 0.5)
    return logsig, rows, sc, grid, layers, y0


def _oracle(layers, logsig, y0, rows, sc, dims, g_ys):
    H, C, width, n_hidden, depth = dims
    lay = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in layers]
    y0o = y0.clone().requires_grad_(True)
    ys = O.solve(lay, torch.from_numpy(logsig), y0o, rows, sc, C, H, depth)
    (ys * g_ys).sum().backward()
    gflat = torch.cat([t.grad.reshape(-1) for W, b in lay for t in (W, b)])
    return ys.detach().numpy(), gflat.numpy(), y0o.grad.numpy()


SOLVE_CASES = [
    # B, L,  C, s, depth, H, width, n_hidden, dt0, scale
    (2, 40, 3, 4, 2, 16, 8, 2, 0.05, 1.0),        # generic kernels
    (2, 40, 3, 4, 1, 16, 8, 2, 0.05, 1.0),
    (1, 36, 5, 3, 2, 24, 12, 3, 0.031, 2.0),
    (2, 96, 7, 12, 2, 128, 32, 2, 0.11, 3.0),     # EigenWorms dims: specialised register-tile kernels
    (1, 96, 6, 12, 2, 128, 32, 2, 0.11, 3.0),
    (1, 96, 4, 12, 2, 128, 32, 2, 0.11, 3.0),
    (1, 96, 7, 12, 1, 128, 32, 2, 0.11, 3.0),
    (1, 40, 6, 4, 2, 64, 128, 3, 0.1, 4.0),       # toy dims: weights partly streamed from global
    (1, 50, 7, 10, 2, 128, 64, 3, 0.2, 4.0),      # PPG dims
]


@pytest.mark.parametrize("case", SOLVE_CASES)
def test_emu_logncde_forward_and_adjoint(case):
    B, L, C, s, depth, H, width, n_hidden, dt0, scale = case
    logsig, rows, sc, grid, layers, y0 = _setup(B, L, C, s, depth, H, width, n_hidden, dt0, scale)
    dims = (H, C, width, n_hidden, depth)
    g_ys = torch.from_numpy(np.random.default_rng(1).normal(size=(B, len(grid), H)))
    flat = O.flatten_params(layers).numpy()
    ys_o, gp_o, gy_o = _oracle(layers, logsig, y0, rows, sc, dims, g_ys)
    # training-mode forward (saves its intermediates) + saved-mode adjoint
    ys_e, gp_e, gy_e = E.logncde_fwd_bwd(flat, logsig, y0.numpy(), rows, sc, dims, g_ys.numpy())
    assert rel_err(ys_e, ys_o) < TOL
    assert rel_err(gy_e, gy_o) < 5 * TOL
    off = 0
    sizes = O.mlp_sizes(H, C, width, n_hidden)
    for i, o in zip(sizes[:-1], sizes[1:]):
        for n in (o * i, o):
            assert rel_err(gp_e[off : off + n], gp_o[off : off + n]) < 5 * TOL, (case, off)
            off += n
    # inference-mode forward and the recomputing adjoint agree with the saved-mode pair
    ys_i, _, _ = E.logncde_fwd_bwd(flat, logsig, y0.numpy(), rows, sc, dims, None, train=False)
    assert np.array_equal(ys_i, ys_e)
    gp_r, gy_r = E.logncde_bwd_recompute(flat, logsig, ys_e, rows, sc, dims, g_ys.numpy())
    assert rel_err(gy_r, gy_o) < 5 * TOL
    assert rel_err(gp_r, gp_o) < 5 * TOL


STREAM_CASES = [
    (1, 40, 6, 4, 2, 64, 128, 3, 0.1, 4.0),       # toy dims: last layer in global memory (in_dim 128: all lanes busy)
    (1, 50, 7, 10, 2, 128, 64, 3, 0.2, 4.0),      # PPG dims, depth 2 (in_dim 64: two row groups per warp)
    (2, 50, 7, 10, 1, 128, 64, 3, 0.2, 4.0),      # PPG reference config (depth 1)
    (1, 30, 4, 5, 2, 128, 64, 3, 0.2, 2.0),       # BASELINE's 4-channel PPG variant
    (1, 24, 3, 4, 2, 36, 200, 2, 0.2, 1.0),       # in_dim 200 > 128: two column chunks per row; H not a power of two
    (1, 24, 2, 4, 2, 40, 176, 4, 0.3, 1.0),       # several layers in global memory (hidden layers too)
]


@pytest.mark.parametrize("case", STREAM_CASES)
def test_emu_logncde_stream_variant(case):
    """lncde_set_tuning("k3_stream", 1): the streamed global-weight mat-vecs of the generic kernels.  The forward
    sums are formed in the same order (bit-identical states); the transposed products associate differently."""
    B, L, C, s, depth, H, width, n_hidden, dt0, scale = case
    logsig, rows, sc, grid, layers, y0 = _setup(B, L, C, s, depth, H, width, n_hidden, dt0, scale)
    dims = (H, C, width, n_hidden, depth)
    g_ys = torch.from_numpy(np.random.default_rng(1).normal(size=(B, len(grid), H)))
    flat = O.flatten_params(layers).numpy()
    ys_o, gp_o, gy_o = _oracle(layers, logsig, y0, rows, sc, dims, g_ys)
    ys_d, _, _ = E.logncde_fwd_bwd(flat, logsig, y0.numpy(), rows, sc, dims, None, train=False)
    E.set_tuning("k3_stream", 1)
    try:
        ys_s, _, _ = E.logncde_fwd_bwd(flat, logsig, y0.numpy(), rows, sc, dims, None, train=False)
        gp_s, gy_s = E.logncde_bwd_recompute(flat, logsig, ys_s, rows, sc, dims, g_ys.numpy())
    finally:
        E.set_tuning("k3_stream", 0)
    assert np.array_equal(ys_s, ys_d)
    assert rel_err(ys_s, ys_o) < TOL
    assert rel_err(gy_s, gy_o) < 5 * TOL
    off = 0
    sizes = O.mlp_sizes(H, C, width, n_hidden)
    for i, o in zip(sizes[:-1], sizes[1:]):
        for n in (o * i, o):
            assert rel_err(gp_s[off : off + n], gp_o[off : off + n]) < 5 * TOL, (case, off)
            off += n


@pytest.mark.parametrize("case", [c for c in SOLVE_CASES if c[5:8] == (128, 32, 2)])
def test_emu_logncde_v3_six_barrier_kernels(case):
    """logncde_v3.cuh (LNCDE_FLAG_K3_V3): six block barriers per stage - partial sums finished by every warp that
    needs them, Lambda-mix applied to the K-slice partials, Heun update by the owners of the first-layer K-slices.
    Another association of three sums: states and gradients agree with the default kernels to rounding and with the
    fp64 oracle inside the parity tolerance; the saved-mode adjoint consumes what the v3 forward saved."""
    B, L, C, s, depth, H, width, n_hidden, dt0, scale = case
    logsig, rows, sc, grid, layers, y0 = _setup(B, L, C, s, depth, H, width, n_hidden, dt0, scale)
    dims = (H, C, width, n_hidden, depth)
    g_ys = np.random.default_rng(1).normal(size=(B, len(grid), H))
    flat = O.flatten_params(layers).numpy()
    ref = E.logncde_fwd_bwd(flat, logsig, y0.numpy(), rows, sc, dims, g_ys)
    E.set_tuning("k3_v3", 1)
    try:
        got = E.logncde_fwd_bwd(flat, logsig, y0.numpy(), rows, sc, dims, g_ys)
        got_i = E.logncde_fwd_bwd(flat, logsig, y0.numpy(), rows, sc, dims, None, train=False)
        got_r = E.logncde_bwd_recompute(flat, logsig, got[0], rows, sc, dims, g_ys)
    finally:
        E.set_tuning("k3_v3", 0)
    assert np.array_equal(got_i[0], got[0])  # saving does not change the arithmetic
    assert not np.array_equal(got[0], ref[0]) or depth == 1  # the new kernels really ran
    assert rel_err(got[0], ref[0]) < 2e-6 and rel_err(got[1], ref[1]) < 1e-5 and rel_err(got[2], ref[2]) < 1e-5
    ys_o, gp_o, gy_o = _oracle(layers, logsig, y0, rows, sc, dims, torch.from_numpy(g_ys))
    assert rel_err(got[0], ys_o) < TOL
    for gp, gy in ((got[1], got[2]), got_r):
        assert rel_err(gp, gp_o) < TOL and rel_err(gy, gy_o) < TOL


# ------------------------------------------------------------------------------------------ K3n (NRDE)
NRDE_CASES = [
    # B, L,  C, s, depth, H, width, n_vf, dt0, scale
    (2, 40, 3, 4, 2, 16, 8, 2, 0.05, 1.0),
    (2, 40, 3, 4, 1, 12, 8, 1, 0.05, 1.0),     # single vf layer, H not a multiple of 4 rows per warp group
    (1, 36, 5, 3, 2, 22, 12, 3, 0.031, 2.0),   # H % 4 != 0
    (1, 48, 6, 4, 2, 64, 64, 3, 0.07, 1.0),    # EigenWorms NRDE config dims (Dm = 21): Wm fully resident
    (1, 30, 7, 6, 2, 128, 64, 4, 0.15, 1.5),   # Wm (917 KB) mostly streamed from global memory
    (1, 24, 3, 4, 3, 32, 16, 2, 0.1, 1.0),     # depth-3 rows (Dm = 13)
]


@pytest.mark.parametrize("case", NRDE_CASES)
def test_emu_nrde_forward_and_adjoint(case):
    from oracle import nrde as ON

    B, L, C, s, depth, H, width, n_vf, dt0, scale = case
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float64)
    ts = paths.make_ts(L)
    rows, sc, grid = O.stage_table(ts, paths.make_intervals(ts, s), dt0, logsig.shape[1])
    Dm = logsig.shape[2] - 1
    layers = ON.init_params(H, Dm, width, n_vf, scale=scale, seed=3, dtype=torch.float64)
    y0 = torch.from_numpy(rng.normal(size=(B, H)) * 0.5)
    g_ys = torch.from_numpy(rng.normal(size=(B, len(grid), H)))
    lay = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in layers]
    y0o = y0.clone().requires_grad_(True)
    ys_o = ON.solve(lay, torch.from_numpy(logsig), y0o, rows, sc, H)
    (ys_o * g_ys).sum().backward()
    flat = ON.flatten_params(layers).numpy()
    ys_e, gp_e, gy_e = E.nrde_fwd_bwd(flat, logsig, y0.numpy(), rows, sc, (H, width, n_vf), g_ys.numpy())
    assert rel_err(ys_e, ys_o.detach().numpy()) < TOL
    assert rel_err(gy_e, y0o.grad.numpy()) < 5 * TOL
    off = 0
    for (W, b) in lay:
        for t in (W, b):
            n = t.numel()
            assert rel_err(gp_e[off : off + n], t.grad.reshape(-1).numpy()) < 5 * TOL, (case, off)
            off += n
    assert off == gp_e.size


# ------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("nb,bs,T,n_basis", [(16, 1, 40, 3), (8, 4, 33, 6), (4, 8, 20, 6), (2, 32, 12, 3),
                                             (1, 64, 9, 6), (3, 5, 14, 3)])
def test_emu_linear_scan_forward_and_adjoint(nb, bs, T, n_basis):
    rng = np.random.default_rng(3)
    B, H = 2, nb * bs
    D = n_basis + 1
    lie = rng.normal(size=(nb, bs, bs, n_basis)) * 0.3
    ls = rng.normal(size=(B, T, D)) * 0.2
    ls[..., 0] = 0
    y0 = rng.normal(size=(B, H))
    g_ys = rng.normal(size=(B, T, H))
    lie_t = torch.from_numpy(lie).requires_grad_(True)
    y0_t = torch.from_numpy(y0).requires_grad_(True)
    flows = torch.einsum("nijl,btl->btnij", lie_t, torch.from_numpy(ls[..., 1:])) + torch.eye(bs, dtype=torch.float64)
    ys = [y0_t]
    for t in range(1, T):  # flow 0 is never applied (quirk B4)
        ys.append(torch.einsum("bnij,bnj->bni", flows[:, t], ys[-1].view(B, nb, bs)).reshape(B, H))
    ys_o = torch.stack(ys, 1)
    (ys_o * torch.from_numpy(g_ys)).sum().backward()
    ys_e, gl_e, gy_e = E.linear_scan(lie, ls, y0, nb, bs, g_ys)
    assert rel_err(ys_e, ys_o.detach().numpy()) < TOL
    assert rel_err(gy_e, y0_t.grad.numpy()) < 5 * TOL
    assert rel_err(gl_e, lie_t.grad.numpy()) < 5 * TOL


@pytest.mark.parametrize("bs,T,B", [(64, 1, 1), (64, 2, 2), (64, 3, 1), (64, 27, 2), (128, 4, 1), (128, 11, 2)])
def test_emu_dense_cluster_sweeps(bs, T, B):
    """linear_cluster.cuh: the forward and the lam-only reverse sweep of the dense model with the series spread over a
    thread-block cluster.  The emulator runs a cluster of ONE CTA: row slicing, the ring of flow stages (deeper than
    T, equal to, and wrapped several times), the st.async + byte-counted mbarrier exchange with its arming order and
    parities, and the exit protocol are the code under test; the peer addressing (mapa) is covered on the GPU."""
    rng = np.random.default_rng(5)
    nb, n_basis = 1, 6
    lie = rng.normal(size=(nb, bs, bs, n_basis)) * (0.5 / np.sqrt(bs))
    ls = rng.normal(size=(B, T, n_basis + 1)) * 0.2
    y0, g_ys = rng.normal(size=(B, bs)), rng.normal(size=(B, T, bs))
    ys0, gl0, gy0 = E.linear_scan(lie, ls, y0, nb, bs, g_ys)
    E.set_tuning("k2_de_v2", 2)
    try:
        ys1, gl1, gy1 = E.linear_scan(lie, ls, y0, nb, bs, g_ys)
        E.set_tuning("k2_cluster", 1)
        E.peer_stores()
        ys2, gl2, gy2 = E.linear_scan(lie, ls, y0, nb, bs, g_ys)
        sent = E.peer_stores()
    finally:
        E.set_tuning("k2_de_v2", 0)
        E.set_tuning("k2_cluster", 0)
    assert sent == B * (T - 1) * (bs + 256)  # forward: one store per row and step; reverse: one per thread and step
    assert rel_err(ys2, ys0) < 2e-6 and rel_err(gy2, gy0) < 2e-6 and rel_err(gl2, gl0) < 1e-5
    assert np.array_equal(ys2, ys1)  # same row arithmetic as the one-CTA sweep
    assert rel_err(gy2, gy1) < 1e-6


@pytest.mark.parametrize("level", [2, 3])
@pytest.mark.parametrize("nb,bs,T,n_basis,B", [(1, 64, 9, 6, 2), (1, 128, 7, 28, 2), (2, 64, 5, 21, 3), (1, 64, 2, 3, 1),
                                               (1, 68, 6, 10, 2), (1, 128, 12, 36, 1), (1, 64, 5, 40, 2), (1, 72, 4, 70, 1),
                                               (1, 128, 14, 5, 3), (1, 128, 2, 4, 5), (1, 128, 6, 70, 2)])
def test_emu_linear_scan_dense_v2_variant(nb, bs, T, n_basis, B, level):
    """Dense-model kernels (linear_de.cuh) against the baseline kernels (LNCDE_FLAG_GENERIC).  Level 2
    (LNCDE_FLAG_K2_MMA_SYNC): lam-only reverse sweep, triple-product bracket gradient, the flow build on the tensor pipe (3xTF32 split precision, mma.sync modelled by the emulator
    with the documented fragment layout) - fp32-grade flows, everything inside the parity tolerance; n_basis > 36
    (depth-3 sizes, several K tiles) keeps the default reverse path.  Level 3: the flow build through
    flow_tcgen05.cu (block sizes that are multiples of 128; others keep level 2) - on the emulator the operand PACKING
    kernel (the K-major / no-swizzle shared-memory image of tcgen05.mma, hi | lo TF32 planes, K padded to 32), the
    identity and the workspace layout are the code under test, the MMA unit itself is replaced by a plain fp32
    evaluation of the same three partial products from the packed operands; with n_basis <= 32 the bracket gradient
    runs through dlie_tcgen05.cu the same way (any n_basis at level 3, in tiles of 32 basis elements: 36 and 70 have a
    partial last tile, 70 takes the scalar stores) (operand-forming code shared with the tcgen05 kernel: pair -> row
    bookkeeping across samples, including T = 2 where every pair starts a new sample, K splits, zero fill, hi | lo
    image; the MMA unit replaced by plain fp32); the rest of the step is level 2."""
    rng = np.random.default_rng(11)
    H, D = nb * bs, n_basis + 2  # one unused trailing logsig column
    lie = rng.normal(size=(nb, bs, bs, n_basis)) * (0.5 / np.sqrt(bs))
    ls = rng.normal(size=(B, T, D)) * 0.2
    ls[..., 0] = 0
    y0 = rng.normal(size=(B, H))
    g_ys = rng.normal(size=(B, T, H))
    ys0, gl0, gy0 = E.linear_scan(lie, ls, y0, nb, bs, g_ys)
    E.set_tuning("k2_de_v2", level)
    try:
        ys1, gl1, gy1 = E.linear_scan(lie, ls, y0, nb, bs, g_ys)
    finally:
        E.set_tuning("k2_de_v2", 0)
    assert rel_err(ys1, ys0) < 2e-6 and rel_err(gy1, gy0) < 2e-6
    assert rel_err(gl1, gl0) < 1e-5
    # and against fp64 autograd
    lie_t = torch.from_numpy(lie).requires_grad_(True)
    y0_t = torch.from_numpy(y0).requires_grad_(True)
    flows = torch.einsum("nijl,btl->btnij", lie_t, torch.from_numpy(ls[..., 1:1 + n_basis])) + torch.eye(bs, dtype=torch.float64)
    y = y0_t
    out = [y]
    for t in range(1, T):
        y = torch.einsum("bnij,bnj->bni", flows[:, t], y.view(B, nb, bs)).reshape(B, H)
        out.append(y)
    ys_o = torch.stack(out, 1)
    (ys_o * torch.from_numpy(g_ys)).sum().backward()
    assert rel_err(ys1, ys_o.detach().numpy()) < TOL
    assert rel_err(gy1, y0_t.grad.numpy()) < 5 * TOL
    assert rel_err(gl1, lie_t.grad.numpy()) < 5 * TOL


@pytest.mark.parametrize("nb,bs,T,n_basis,B", [(32, 4, 200, 6, 2), (5, 4, 66, 3, 1), (128, 1, 131, 5, 2), (8, 2, 97, 4, 3),
                                               (3, 8, 70, 6, 2), (16, 4, 65, 28, 1), (7, 1, 260, 2, 2)])
def test_emu_linear_scan_time_parallel(nb, bs, T, n_basis, B):
    """tuning "k2_pscan" (linear_pscan.cuh): chunk products, boundary recurrence over the chunks, per-chunk sweeps -
    forward states, g_y0 and the bracket gradient against the sequential default kernels (another association of the
    same products: agreement to rounding) and against fp64 autograd of the recurrence.  Shapes: H not a multiple of 32
    (warps straddling chunks), T - 1 not a multiple of the chunk, the shortest T the variant accepts (66), bs 1-8;
    T = 65 (T - 1 = 2 chunks exactly) stays on the default kernels."""
    rng = np.random.default_rng(nb * 100 + T)
    H, D = nb * bs, n_basis + 1
    lie = rng.normal(size=(nb, bs, bs, n_basis)) * (0.3 / np.sqrt(bs))
    ls = rng.normal(size=(B, T, D)) * 0.1
    ls[..., 0] = 0
    y0 = rng.normal(size=(B, H))
    g_ys = rng.normal(size=(B, T, H))
    ys0, gl0, gy0 = E.linear_scan(lie, ls, y0, nb, bs, g_ys)
    E.set_tuning("k2_pscan", 1)
    try:
        ys1, gl1, gy1 = E.linear_scan(lie, ls, y0, nb, bs, g_ys)
    finally:
        E.set_tuning("k2_pscan", 0)
    if T - 1 <= 64:
        assert np.array_equal(ys0, ys1) and np.array_equal(gl0, gl1) and np.array_equal(gy0, gy1)
    else:
        assert not np.array_equal(ys0, ys1)  # the variant really ran
    assert rel_err(ys1, ys0) < 3e-6 and rel_err(gy1, gy0) < 3e-6 and rel_err(gl1, gl0) < 1e-5
    lie_t = torch.from_numpy(lie).requires_grad_(True)
    y0_t = torch.from_numpy(y0).requires_grad_(True)
    flows = torch.einsum("nijl,btl->btnij", lie_t, torch.from_numpy(ls[..., 1:1 + n_basis])) + torch.eye(bs, dtype=torch.float64)
    y = y0_t
    out = [y]
    for t in range(1, T):
        y = torch.einsum("bnij,bnj->bni", flows[:, t], y.view(B, nb, bs)).reshape(B, H)
        out.append(y)
    ys_o = torch.stack(out, 1)
    (ys_o * torch.from_numpy(g_ys)).sum().backward()
    assert rel_err(ys1, ys_o.detach().numpy()) < TOL
    assert rel_err(gy1, y0_t.grad.numpy()) < 5 * TOL
    assert rel_err(gl1, lie_t.grad.numpy()) < 5 * TOL


def test_emu_tf32_split_flow_build_accuracy():
    """3xTF32 flows against fp64: errors at the fp32 rounding level (a single-TF32 product would be ~5e-4 relative)."""
    rng = np.random.default_rng(2)
    nb, bs, T, n_basis, B = 1, 128, 3, 28, 2
    lie = rng.normal(size=(nb, bs, bs, n_basis)) * 0.1
    ls = rng.normal(size=(B, T, n_basis + 1)) * 0.3
    y0 = np.zeros((B, bs))
    y0[:, 5] = 1.0  # ys[:, 1] = column 5 of F_1: reads the flows back through the scan
    want = np.einsum("ijl,bl->bij", lie[0], ls[:, 1, 1:]) + np.eye(bs)
    E.set_tuning("k2_de_v2", 2)
    try:
        ys, _, _ = E.linear_scan(lie, ls, y0, nb, bs)
    finally:
        E.set_tuning("k2_de_v2", 0)
    err = np.abs(ys[:, 1] - want[:, :, 5]).max()
    assert err < 3e-7, err


# ------------------------------------------------------------------------------------------ K2d
@pytest.mark.parametrize("B,T,H,C,D", [(2, 40, 16, 3, 4), (3, 131, 128, 7, 29), (1, 2, 5, 1, 2), (2, 1, 8, 2, 4),
                                       (2, 65, 200, 8, 37), (1, 64, 130, 6, 22), (2, 129, 7, 4, 11)])
def test_emu_fused_diagonal_scan(B, T, H, C, D):
    """lncde_dlncde_fwd/bwd (no flows in HBM) against torch autograd of the recurrence and against the generic
    flow-build + scan kernels with bs = 1."""
    rng = np.random.default_rng(B * 1000 + T)
    A = rng.normal(size=(H, C)) * 0.3
    ls = rng.normal(size=(B, T, D)) * 0.2
    ls[..., 0] = 0
    y0 = rng.normal(size=(B, H))
    g_ys = rng.normal(size=(B, T, H))
    A_t = torch.from_numpy(A).requires_grad_(True)
    y0_t = torch.from_numpy(y0).requires_grad_(True)
    flows = 1 + torch.from_numpy(ls[..., 1 : 1 + C]) @ A_t.T
    ys = [y0_t]
    for t in range(1, T):
        ys.append(flows[:, t] * ys[-1])
    ys_o = torch.stack(ys, 1)
    (ys_o * torch.from_numpy(g_ys)).sum().backward()
    ys_e, gA_e, gy_e = E.dlncde(A, ls, y0, g_ys)
    assert rel_err(ys_e, ys_o.detach().numpy()) < TOL
    assert rel_err(gy_e, y0_t.grad.numpy()) < 5 * TOL
    gA_o = np.zeros((H, C)) if A_t.grad is None else A_t.grad.numpy()  # T = 1: no flow is ever applied
    assert rel_err(gA_e, gA_o) < 5 * TOL
    ys_g, gl_g, gy_g = E.linear_scan(A.reshape(H, 1, 1, C), ls, y0, H, 1, g_ys)
    assert rel_err(ys_e, ys_g) < 1e-5 and rel_err(gy_e, gy_g) < 5e-5 and rel_err(gA_e, gl_g.reshape(H, C)) < 5e-5


@pytest.mark.parametrize("B,T,H,C,D", [(2, 40, 16, 3, 4), (3, 131, 128, 7, 29), (2, 1, 8, 2, 4), (70, 66, 40, 5, 9),
                                       (33, 3, 130, 8, 12)])
def test_emu_fused_diagonal_scan_mean_readout(B, T, H, C, D):
    """lncde_dlncde_fwd_mean / _bwd_mean (the time mean formed by the scan kernel, its cotangent taken as one [B, H]
    array) against the plain pair fed with the broadcast cotangent g_mean / T: same states, the mean within fp32
    rounding of numpy's, bit-equal gradients per sample; B > 32 exercises the two-level batch reduction."""
    rng = np.random.default_rng(B * 77 + T)
    A = (rng.normal(size=(H, C)) * 0.3).astype(np.float32)
    ls = (rng.normal(size=(B, T, D)) * 0.2).astype(np.float32)
    ls[..., 0] = 0
    y0 = rng.normal(size=(B, H)).astype(np.float32)
    gm = rng.normal(size=(B, H)).astype(np.float32)
    ys_m, ym, gA_m, gy_m = E.dlncde_mean(A, ls, y0, gm)
    g_ys = np.broadcast_to((gm / np.float32(T))[:, None, :], (B, T, H)).copy()
    ys_p, gA_p, gy_p = E.dlncde(A, ls, y0, g_ys)
    assert np.array_equal(ys_m, ys_p)
    assert rel_err(ym, ys_p.astype(np.float64).mean(1)) < 2e-6
    assert np.array_equal(gy_m, gy_p)
    assert rel_err(gA_m, gA_p) < 1e-6  # same per-sample partials; B > 32 adds them in chunks of 32
    # fp64 recurrence (checks the batch reduction itself)
    f = 1.0 + ls[..., 1:1 + C].astype(np.float64) @ A.astype(np.float64).T  # [B, T, H]
    g = gm.astype(np.float64) / T
    lam, gA = g.copy(), np.zeros((H, C))
    ys64 = ys_p.astype(np.float64)
    for t in range(T - 1, 0, -1):
        gA += np.einsum("bh,bc->hc", lam * ys64[:, t - 1], ls[:, t, 1:1 + C].astype(np.float64))
        lam = f[:, t] * lam + g
    assert rel_err(gA_m, gA) < 5 * TOL and rel_err(gy_m, lam) < 5 * TOL


# ------------------------------------------------------------------------------------------ reference-executed fixtures
# tests/golden/ref_*.npz hold outputs of the reference's OWN code (tests/golden/make_golden_ref.py): the emulated
# kernels are compared with them directly, the oracle only supplies the glue the reference does in its host code.
from conftest import GOLDEN  # noqa: E402

_P = np.load(os.path.join(GOLDEN, "ref_paths.npz"))
_N = np.load(os.path.join(GOLDEN, "ref_logncde.npz"))
_LN = np.load(os.path.join(GOLDEN, "ref_linear.npz"))
_NR = np.load(os.path.join(GOLDEN, "ref_nrde.npz"))


@pytest.mark.parametrize("tma", [False, True])
@pytest.mark.parametrize("i", range(int(_P["n"])))
def test_emu_logsig_vs_reference_code(i, tma):
    B, L, C, s, d = (int(v) for v in _P[f"cfg{i}"])
    got = E.calc_paths(_P[f"x{i}"], s, d, tma=tma)
    assert got.shape == _P[f"logsig{i}"].shape
    assert rel_err(got, _P[f"logsig{i}"]) < TOL


@pytest.mark.parametrize("fused", [0, 1])  # 1: the six-barrier kernels (where the shape has them)
@pytest.mark.parametrize("i", range(int(_N["n"])))
def test_emu_logncde_vs_reference_code(i, fused):
    B, L, C, s, depth, H, width, nh, cls, ostep, ld = (int(v) for v in _N[f"cfg{i}"])
    dt0 = float(_N[f"hyper{i}"][0])
    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    rows, sc, grid = O.stage_table(ts, iv, dt0, _N[f"logsig{i}"].shape[1])
    l1, l2 = _N[f"lin1_{i}"], _N[f"lin2_{i}"]
    W1, b1, W2, b2 = l1[: H * C].reshape(H, C), l1[H * C :], l2[: ld * H].reshape(ld, H), l2[ld * H :]
    y0 = _N[f"x{i}"][:, 0] @ W1.T + b1
    E.set_tuning("k3_v3", fused)
    try:
        ys, _, _ = E.logncde_fwd_bwd(_N[f"mlp{i}"], _N[f"logsig{i}"], y0, rows, sc, (H, C, width, nh, depth), None,
                                     train=False)
    finally:
        E.set_tuning("k3_v3", 0)
    ys = torch.from_numpy(ys).double()
    if cls:
        pred = torch.softmax(ys[:, -1] @ torch.from_numpy(W2).T + torch.from_numpy(b2), -1)
    else:
        out = torch.cat([O.interp_outputs(ys, grid, O.regression_times(L, ostep)), ys[:, -1:]], 1)
        pred = torch.tanh(out @ torch.from_numpy(W2).T + torch.from_numpy(b2))
    assert rel_err(pred.numpy(), _N[f"pred{i}"]) < TOL


@pytest.mark.parametrize("i", range(int(_NR["n"])))
def test_emu_nrde_vs_reference_code(i):
    """K3n kernels against predictions and parameter gradients of the reference's own NeuralRDE / loss code."""
    from oracle import nrde as ON

    B, L, C, s, depth, H, width, nvf, cls, ostep, ld = (int(v) for v in _NR[f"cfg{i}"])
    dt0 = float(_NR[f"hyper{i}"][0])
    logsig = _NR[f"logsig{i}"]
    Dm = logsig.shape[2] - 1
    flat = _NR[f"params{i}"].astype(np.float64)
    n_field = sum(o * (a + 1) for a, o in ON.layer_sizes(H, Dm, width, nvf))
    rest = torch.from_numpy(flat[n_field:])
    W1, b1, W2, b2 = (t.clone().requires_grad_(True) for t in
                      (rest[: H * C].view(H, C), rest[H * C : H * C + H], rest[H * C + H : H * C + H + ld * H].view(ld, H),
                       rest[H * C + H + ld * H :]))
    ts = paths.make_ts(L)
    rows, sc, grid = O.stage_table(ts, paths.make_intervals(ts, s), dt0, logsig.shape[1])
    y0 = torch.from_numpy(_NR[f"x{i}"][:, 0]) @ W1.T + b1
    ys_np, _, _ = E.nrde_fwd_bwd(flat[:n_field], logsig, y0.detach().numpy(), rows, sc, (H, width, nvf))
    ys = torch.from_numpy(ys_np).double().requires_grad_(True)
    y = torch.from_numpy(_NR[f"y{i}"])
    if cls:
        pred = torch.softmax(ys[:, -1] @ W2.T + b2, -1)
        loss = O.classification_loss(pred, y)
    else:
        out = torch.cat([O.interp_outputs(ys, grid, O.regression_times(L, ostep)), ys[:, -1:]], 1)
        pred = torch.tanh(out @ W2.T + b2)
        loss = O.regression_loss(pred, y)
    assert rel_err(pred.detach().numpy(), _NR[f"pred{i}"]) < TOL
    assert abs(float(loss.detach()) - float(_NR[f"loss{i}"])) < TOL * max(1.0, abs(float(_NR[f"loss{i}"])))
    loss.backward()  # read-out in host autograd, as the product does; the solve's cotangent goes through the kernels
    _, gp, gy0 = E.nrde_fwd_bwd(flat[:n_field], logsig, y0.detach().numpy(), rows, sc, (H, width, nvf), ys.grad.numpy())
    assert rel_err(gp, _NR[f"grad{i}"][:n_field]) < 5 * TOL
    gW1 = gy0.astype(np.float64).T @ _NR[f"x{i}"][:, 0]
    assert rel_err(gW1.ravel(), _NR[f"grad{i}"][n_field : n_field + H * C]) < 5 * TOL


@pytest.mark.parametrize("i", range(int(_LN["n"])))
def test_emu_linear_scan_vs_reference_code(i):
    B, L, C, s, depth, H, bs, Psteps, cls, ld = (int(v) for v in _LN[f"cfg{i}"])
    vfA = torch.from_numpy(_LN[f"vfA{i}"])
    ini, out = _LN[f"init{i}"], _LN[f"out{i}"]
    Wi, bi, Wo, bo = ini[: H * C].reshape(H, C), ini[H * C :], out[: ld * H].reshape(ld, H), out[ld * H :]
    nb = H // bs
    if bs == 1:
        lie = vfA.T.reshape(nb, 1, 1, C).numpy()
    else:
        from oracle.hall import Hall

        hall = Hall(C, depth)
        vfs = vfA.reshape(C, nb, bs, bs)
        lie = torch.stack([OL.log_ode(vfs[:, b], hall) for b in range(nb)]).numpy()
    y0 = _LN[f"x{i}"][:, 0] @ Wi.T + bi
    ys, _, _ = E.linear_scan(lie, _LN[f"logsig{i}"], y0, nb, bs)
    if bs == 1:  # the fused diagonal kernels give the same states
        ys_f, _, _ = E.dlncde(vfA.T.numpy(), _LN[f"logsig{i}"], y0)
        assert rel_err(ys_f, ys) < 1e-5
    ys = torch.from_numpy(ys).double()
    Wo_t, bo_t = torch.from_numpy(Wo), torch.from_numpy(bo)
    pred = torch.softmax(ys.mean(1) @ Wo_t.T + bo_t, -1) if cls else torch.tanh(ys @ Wo_t.T + bo_t)
    assert rel_err(pred.numpy(), _LN[f"pred{i}"]) < TOL
