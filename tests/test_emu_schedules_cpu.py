"""Race probe for the kernel sources on the host emulator: between two synchronisation points the emulator runs the
threads of a CTA one after the other, so a missing barrier (a thread reading what another thread of the same phase
writes) is invisible under one fixed order.  Every kernel family is therefore executed under several schedules -
ascending, descending, a random permutation per sweep and "warp-greedy" (one warp at a time runs as far as it can, so it
reaches the next phase while the other warps are still in the previous one: write-after-read hazards of a missing block
barrier), plus the CTAs of the grid in reverse - and must return
bit-identical results: all kernels are deterministic (no atomics, fixed-order reductions).  The emulator cannot see
races that hardware timing alone creates inside one phase of ONE thread order it never produces; it does see every
read-after-write / write-after-read pair between different threads of a phase that changes the result."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import emu_lib as E  # noqa: E402
from oracle import log_ncde as O  # noqa: E402
from oracle import nrde as ON  # noqa: E402
from oracle import paths  # noqa: E402

SCHEDULES = [("descending", 1, False), ("random", 7, True), ("warp_greedy", 1, False), ("warp_greedy_desc", 1, True)]


def _under_schedules(fn):
    """fn() -> tuple of arrays; asserts bit-identity under every schedule, returns the reference result."""
    E.set_schedule("ascending")
    ref = fn()
    try:
        for mode, seed, rev in SCHEDULES:
            E.set_schedule(mode, seed, rev)
            got = fn()
            for a, b in zip(ref, got):
                if a is None:
                    assert b is None
                    continue
                assert np.array_equal(a, b, equal_nan=True), (mode, seed, rev)
    finally:
        E.set_schedule("ascending")
    return ref


def test_schedule_hook_detects_a_missing_barrier():
    L = E.lib()
    L.emu_selftest_race.argtypes = [C.c_void_p, C.c_int, C.c_int]
    n = 64

    def run(with_barrier):
        out = np.zeros(n, np.float32)
        L.emu_selftest_race(out.ctypes.data, n, with_barrier)
        return out

    try:
        res = {}
        for mode in ("ascending", "descending", "random"):
            E.set_schedule(mode, 3)
            res[mode] = (run(1), run(0))
    finally:
        E.set_schedule("ascending")
    want = np.roll(np.arange(1, n + 1, dtype=np.float32), -1)
    for mode in res:
        assert np.array_equal(res[mode][0], want)                      # with the barrier: right under every schedule
    assert not np.array_equal(res["ascending"][1], res["descending"][1], equal_nan=True)   # without: order-dependent
    assert not np.array_equal(res["ascending"][1], want, equal_nan=True)


@pytest.mark.parametrize("tma", [False, True])
def test_logsig_kernels_schedule_independent(tma):
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.normal(scale=0.3, size=(3, 250, 7)), axis=1).astype(np.float32)
    _under_schedules(lambda: (E.calc_paths(x, 12, 2, tma=tma), E.calc_paths(x[:, :, :5], 7, 1, tma=tma)))


def test_logsig_tensor_kernels_schedule_independent():
    rng = np.random.default_rng(1)
    x = np.cumsum(rng.normal(scale=0.3, size=(2, 60, 3)), axis=1).astype(np.float32)
    _under_schedules(lambda: (E.calc_paths(x, 5, 3), E.calc_paths(x[:, :, :2], 4, 4)))


def _solve_setup(B, L, Cn, s, depth, H, width, n_hidden, dt0, scale):
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.normal(scale=0.3, size=(B, L, Cn)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float64)
    ts = paths.make_ts(L)
    rows, sc, grid = O.stage_table(ts, paths.make_intervals(ts, s), dt0, logsig.shape[1])
    layers = O.init_params(H, Cn, width, n_hidden, scale=scale, seed=0, dtype=torch.float64)
    y0 = rng.normal(size=(B, H)) * 0.5
    g_ys = rng.normal(size=(B, len(grid), H))
    return O.flatten_params(layers).numpy(), logsig, y0, rows, sc, g_ys


@pytest.mark.parametrize("fused", [0, 1])  # 1: the six-barrier kernels (logncde_v3.cuh): shuffles + per-warp slots
def test_specialised_logncde_kernels_schedule_independent(fused):
    flat, logsig, y0, rows, sc, g_ys = _solve_setup(1, 48, 7, 12, 2, 128, 32, 2, 0.26, 3.0)
    dims = (128, 7, 32, 2, 2)
    E.set_tuning("k3_v3", fused)
    try:
        def run():
            ys, gp, gy = E.logncde_fwd_bwd(flat, logsig, y0, rows, sc, dims, g_ys)          # saved-forward adjoint
            gp2, gy2 = E.logncde_bwd_recompute(flat, logsig, ys, rows, sc, dims, g_ys)      # recomputing adjoint
            return ys, gp, gy, gp2, gy2
        _under_schedules(run)
    finally:
        E.set_tuning("k3_v3", 0)


@pytest.mark.parametrize("stream", [0, 1])
@pytest.mark.parametrize("case", [(1, 24, 3, 4, 2, 16, 8, 2, 0.13, 1.0), (1, 20, 4, 5, 2, 128, 64, 3, 0.3, 2.0)])
def test_generic_logncde_kernels_schedule_independent(case, stream):
    B, L, Cn, s, depth, H, width, n_hidden, dt0, scale = case
    flat, logsig, y0, rows, sc, g_ys = _solve_setup(*case)
    dims = (H, Cn, width, n_hidden, depth)
    E.set_tuning("k3_stream", stream)
    try:
        def run():
            ys, _, _ = E.logncde_fwd_bwd(flat, logsig, y0, rows, sc, dims, None, train=False)
            gp, gy = E.logncde_bwd_recompute(flat, logsig, ys, rows, sc, dims, g_ys)
            return ys, gp, gy
        _under_schedules(run)
    finally:
        E.set_tuning("k3_stream", 0)


@pytest.mark.parametrize("nb,bs,T,n_basis,level", [(8, 4, 20, 6, 0), (2, 32, 8, 3, 0), (1, 64, 7, 6, 0), (1, 64, 7, 6, 2),
                                                   (1, 64, 3, 8, 2), (1, 128, 4, 6, 3), (3, 5, 9, 3, 0)])
def test_linear_scan_kernels_schedule_independent(nb, bs, T, n_basis, level):
    rng = np.random.default_rng(3)
    B, H = 2, nb * bs
    lie = rng.normal(size=(nb, bs, bs, n_basis)) * (0.5 / np.sqrt(bs))
    ls = rng.normal(size=(B, T, n_basis + 1)) * 0.2
    y0, g_ys = rng.normal(size=(B, H)), rng.normal(size=(B, T, H))
    E.set_tuning("k2_de_v2", level)
    try:
        _under_schedules(lambda: E.linear_scan(lie, ls, y0, nb, bs, g_ys))
    finally:
        E.set_tuning("k2_de_v2", 0)


@pytest.mark.parametrize("bs,T", [(64, 12), (128, 5)])
def test_dense_cluster_sweeps_schedule_and_delivery_independent(bs, T):
    """linear_cluster.cuh has no block barrier in the forward loop: warps may be a whole step apart."""
    rng = np.random.default_rng(4)
    B, n_basis = 2, 5
    lie = rng.normal(size=(1, bs, bs, n_basis)) * (0.5 / np.sqrt(bs))
    ls = rng.normal(size=(B, T, n_basis + 1)) * 0.2
    y0, g_ys = rng.normal(size=(B, bs)), rng.normal(size=(B, T, bs))
    E.set_tuning("k2_de_v2", 2)
    E.set_tuning("k2_cluster", 1)
    try:
        _under_schedules(lambda: E.linear_scan(lie, ls, y0, 1, bs, g_ys))
        _under_async_modes(lambda: E.linear_scan(lie, ls, y0, 1, bs, g_ys))
    finally:
        E.set_tuning("k2_de_v2", 0)
        E.set_tuning("k2_cluster", 0)


@pytest.mark.parametrize("nb,bs,T,n_basis", [(5, 4, 70, 3), (24, 1, 100, 2), (3, 8, 66, 4)])
def test_time_parallel_scan_schedule_independent(nb, bs, T, n_basis):
    """k2_pscan: every kernel of the two-level scan is deterministic and free of cross-thread hazards between its
    collectives (thread / warp / block orders, reversed grid)."""
    rng = np.random.default_rng(6)
    B, H = 2, nb * bs
    lie = rng.normal(size=(nb, bs, bs, n_basis)) * (0.3 / np.sqrt(bs))
    ls = rng.normal(size=(B, T, n_basis + 1)) * 0.1
    y0, g_ys = rng.normal(size=(B, H)), rng.normal(size=(B, T, H))
    E.set_tuning("k2_pscan", 1)
    try:
        _under_schedules(lambda: E.linear_scan(lie, ls, y0, nb, bs, g_ys))
    finally:
        E.set_tuning("k2_pscan", 0)


def test_fused_diagonal_kernels_schedule_independent():
    rng = np.random.default_rng(4)
    B, T, H, Cn, D = 2, 150, 128, 7, 29
    A, ls = rng.normal(size=(H, Cn)) * 0.3, rng.normal(size=(B, T, D)) * 0.2
    y0, g_ys = rng.normal(size=(B, H)), rng.normal(size=(B, T, H))
    _under_schedules(lambda: E.dlncde(A, ls, y0, g_ys))


@pytest.mark.parametrize("case", [(2, 24, 3, 4, 2, 16, 8, 2, 0.13), (1, 16, 6, 4, 2, 64, 64, 3, 0.26), (1, 12, 7, 6, 2, 128, 64, 4, 0.4)])
def test_nrde_kernels_schedule_independent(case):
    B, L, Cn, s, depth, H, width, n_vf, dt0 = case
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.normal(scale=0.3, size=(B, L, Cn)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float64)
    ts = paths.make_ts(L)
    rows, sc, grid = O.stage_table(ts, paths.make_intervals(ts, s), dt0, logsig.shape[1])
    flat = ON.flatten_params(ON.init_params(H, logsig.shape[2] - 1, width, n_vf, seed=3, dtype=torch.float64)).numpy()
    y0, g_ys = rng.normal(size=(B, H)) * 0.5, rng.normal(size=(B, len(grid), H))
    _under_schedules(lambda: E.nrde_fwd_bwd(flat, logsig, y0, rows, sc, (H, width, n_vf), g_ys))


# ---------------------------------------------------------------------------------------------- asynchronous copies
def test_lazy_async_mode_detects_a_missing_wait():
    L = E.lib()
    L.emu_selftest_async.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    src = np.zeros(8, np.float32)
    off = (-src.ctypes.data) % 16 // 4
    src[off : off + 4] = [1.0, 2.0, 3.0, 4.0]
    out = np.zeros(1, np.float32)
    res = {}
    try:
        for lazy in (False, True):
            E.set_async_mode(lazy)
            for wait in (1, 0):
                L.emu_selftest_async(src.ctypes.data + 4 * off, out.ctypes.data, wait)
                res[(lazy, wait)] = float(out[0])
    finally:
        E.set_async_mode(False)
    assert res[(False, 1)] == res[(True, 1)] == 3.0      # with the wait: right in both modes
    assert res[(False, 0)] == 3.0 and np.isnan(res[(True, 0)])   # without it: only the lazy mode shows the bug


def _under_async_modes(fn):
    E.set_async_mode(False)
    ref = fn()
    try:
        E.set_async_mode(True)
        for sched in (("ascending", 1, False), ("descending", 1, False), ("warp_greedy_desc", 1, True)):
            E.set_schedule(*sched)
            got = fn()
            for a, b in zip(ref, got):
                assert (a is None and b is None) or np.array_equal(a, b, equal_nan=True), sched
    finally:
        E.set_async_mode(False)
        E.set_schedule("ascending")


def test_logsig_tma_kernels_late_delivery():
    """K1 with bulk-copy staging: data delivered only at the mbarrier wait, result rows leaving shared memory only at
    wait_group.read."""
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.normal(scale=0.3, size=(3, 400, 7)), axis=1).astype(np.float32)
    _under_async_modes(lambda: (E.calc_paths(x, 12, 2, tma=True), E.calc_paths(x[:, :, :6], 4, 2, tma=True)))


@pytest.mark.parametrize("level", [0, 2])
def test_dense_scan_tma_rings_late_delivery(level):
    rng = np.random.default_rng(3)
    nb, bs, T, n_basis, B = 1, 64, 11, 6, 2
    lie = rng.normal(size=(nb, bs, bs, n_basis)) * (0.5 / np.sqrt(bs))
    ls = rng.normal(size=(B, T, n_basis + 1)) * 0.2
    y0, g_ys = rng.normal(size=(B, bs)), rng.normal(size=(B, T, bs))
    E.set_tuning("k2_de_v2", level)
    try:
        _under_async_modes(lambda: E.linear_scan(lie, ls, y0, nb, bs, g_ys))
    finally:
        E.set_tuning("k2_de_v2", 0)


@pytest.mark.parametrize("fused", [0, 1])
def test_saved_adjoint_cp_async_prefetch_late_delivery(fused):
    flat, logsig, y0, rows, sc, g_ys = _solve_setup(1, 48, 7, 12, 2, 128, 32, 2, 0.26, 3.0)
    E.set_tuning("k3_v3", fused)
    try:
        _under_async_modes(lambda: E.logncde_fwd_bwd(flat, logsig, y0, rows, sc, (128, 7, 32, 2, 2), g_ys))
    finally:
        E.set_tuning("k3_v3", 0)


# ---------------------------------------------------------------------------------------------- alignment
def test_emulator_traps_misaligned_vector_access():
    """The emulated kernels are compiled with -fsanitize=alignment (tests/emu/build_emu.py): a float4 / float2 / 64-bit
    dereference of a misaligned address - a fault on the GPU, silent on x86 - ends the process.  Checked in a
    subprocess; every other emulator test in the suite runs under the same check."""
    import subprocess
    import textwrap

    code = textwrap.dedent(f'''
        import sys, ctypes as C, numpy as np
        sys.path.insert(0, {os.path.join(HERE, "emu")!r})
        import emu_lib as E
        L = E.lib()
        L.emu_selftest_align.argtypes = [C.c_void_p, C.c_int]
        L.emu_selftest_align.restype = C.c_float
        a = np.arange(16, dtype=np.float32)
        off = (-a.ctypes.data) % 16 // 4
        print("aligned", L.emu_selftest_align(a.ctypes.data, off)); sys.stdout.flush()
        print("misaligned", L.emu_selftest_align(a.ctypes.data, off + int(sys.argv[1])))
    ''')
    ok = subprocess.run([sys.executable, "-c", code, "0"], capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and "misaligned 3.0" in ok.stdout
    bad = subprocess.run([sys.executable, "-c", code, "1"], capture_output=True, text=True, timeout=120)
    assert bad.returncode == 135 and "misaligned vector access" in bad.stderr and "aligned 3.0" in bad.stdout
