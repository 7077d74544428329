"""The PRODUCT host mirrors (lncde.models.*, lncde.train) executed on the CPU over the host-emulated kernel library
(tests/emu): same Python code, same ctypes signatures and argument order as on the GPU box, with the three
device-specific hooks of lncde._lib swapped for the test (library handle, stream, CUDA check).  Test
infrastructure only - the product has no CPU path.  Results are compared with the reference-executed fixtures."""
import contextlib
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

TOL = 1e-5
NR = np.load(os.path.join(GOLDEN, "ref_nrde.npz"))
N = np.load(os.path.join(GOLDEN, "ref_logncde.npz"))


@pytest.fixture
def emulated(monkeypatch):
    import build_emu
    from lncde import _lib

    handle = C.CDLL(build_emu.build())
    _lib._declare(handle)  # the product's own signature table: a symbol or arity mismatch fails here
    monkeypatch.setattr(_lib, "lib", lambda: handle)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "require_cuda_device", lambda *a: None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: C.c_void_p(0))
    monkeypatch.setattr(torch.cuda, "device", lambda *a, **k: contextlib.nullcontext())  # calc_paths' device guard
    from lncde.data_dir import generate_paths as GP
    from lncde.data_dir.hall_set import HallSet

    monkeypatch.setattr(GP, "_device_t2l",  # the K0 tables stay on the host for the emulated library
                        lambda w, d, idx: tuple(torch.from_numpy(a) for a in HallSet(w, d).t2l_csr()))
    return handle


def _ts(L):
    return (np.float32(1.0 / L) * np.arange(L, dtype=np.float32)).astype(np.float32)


def _intervals(ts, s):
    return np.concatenate([ts[np.arange(0, len(ts), s)], np.array([1.0], np.float32)]).astype(np.float32)


@pytest.mark.parametrize("i", range(int(NR["n"])))
def test_nrde_mirror_vs_reference_code(emulated, i):
    from lncde import train as T
    from lncde.models.generate_model import create_model

    B, L, C_, s, depth, H, width, nvf, cls, ostep, ld = (int(v) for v in NR[f"cfg{i}"])
    dt0, scale = (float(v) for v in NR[f"hyper{i}"])
    ts = _ts(L)
    model, state = create_model("nrde", C_, NR[f"logsig{i}"].shape[-1], depth, _intervals(ts, s), ld, H, vf_depth=nvf,
                                vf_width=width, classification=bool(cls), output_step=ostep, dt0=dt0, scale=scale, key=1,
                                device="cpu")
    assert NR[f"params{i}"].size == model.params.numel()
    with torch.no_grad():
        model.params.copy_(torch.from_numpy(NR[f"params{i}"]).float())
    X = (torch.from_numpy(np.tile(ts, (B, 1))), torch.from_numpy(NR[f"logsig{i}"]).float(),
         torch.from_numpy(NR[f"x{i}"][:, 0, :].copy()).float())
    pred = model(X)
    assert rel_err(pred.detach().numpy(), NR[f"pred{i}"]) < TOL
    y = torch.from_numpy(NR[f"y{i}"]).float()
    loss, grads = (T.classification_loss if cls else T.regression_loss)(model, X, y)
    want = float(NR[f"loss{i}"])
    assert abs(float(loss) - want) < TOL * max(1.0, abs(want))
    assert rel_err(grads.numpy(), NR[f"grad{i}"]) < 5 * TOL


@pytest.mark.parametrize("i", [0, 2, 5])
def test_log_ncde_mirror_vs_reference_code(emulated, i):
    """Same check for the hot-path model (the GPU twin is tests/test_reference_golden_gpu.py)."""
    from lncde import train as T
    from lncde.models.generate_model import create_model

    B, L, C_, s, depth, H, width, nh, cls, ostep, ld = (int(v) for v in N[f"cfg{i}"])
    dt0, scale, lambd = (float(v) for v in N[f"hyper{i}"])
    ts = _ts(L)
    model, _ = create_model("log_ncde", C_, N[f"logsig{i}"].shape[-1], depth, _intervals(ts, s), ld, H, vf_depth=nh,
                            vf_width=width, classification=bool(cls), output_step=ostep, dt0=dt0, scale=scale, lambd=lambd,
                            key=1, device="cpu")
    flat = np.concatenate([N[f"mlp{i}"], N[f"lin1_{i}"], N[f"lin2_{i}"]])
    with torch.no_grad():
        model.params.copy_(torch.from_numpy(flat).float())
    X = (torch.from_numpy(np.tile(ts, (B, 1))), torch.from_numpy(N[f"logsig{i}"]).float(),
         torch.from_numpy(N[f"x{i}"][:, 0, :].copy()).float())
    y = torch.from_numpy(N[f"y{i}"]).float()
    loss, grads = (T.classification_loss if cls else T.regression_loss)(model, X, y)
    want = float(N[f"loss{i}"])
    assert abs(float(loss) - want) < TOL * max(1.0, abs(want))
    assert rel_err(grads.numpy(), N[f"grad{i}"]) < 5 * TOL


LN = np.load(os.path.join(GOLDEN, "ref_linear.npz"))


@pytest.mark.parametrize("fused_d", [0, 1])
@pytest.mark.parametrize("i", range(int(LN["n"])))
def test_linear_ncde_mirror_vs_reference_code(emulated, i, fused_d):
    """LogLinearCDE (D / BD / DE, chunked scanner configs) through the product mirror on the emulated K2 kernels;
    fused_d = 1: the library's default kernels (the diagonal configs run the fused K2d kernels), 0: the baseline kernels
    (`kernel_flags("generic")`: flow build + sequential scans)."""
    from lncde import _lib, train as T
    from lncde.models.generate_model import create_model

    B, L, C_, s, depth, H, bs, Psteps, cls, ld = (int(v) for v in LN[f"cfg{i}"])
    model, _ = create_model(str(LN[f"name{i}"]), C_, LN[f"logsig{i}"].shape[-1], depth, None, ld, H, block_size=bs,
                            classification=bool(cls), lambd=float(LN[f"lambd{i}"]), parallel_steps=Psteps, key=1,
                            device="cpu")
    flat = np.concatenate([LN[f"vfA{i}"].ravel(), LN[f"init{i}"], LN[f"out{i}"]])
    with torch.no_grad():
        model.params.copy_(torch.from_numpy(flat).float())
    X = (torch.from_numpy(np.tile(_ts(L), (B, 1))), torch.from_numpy(LN[f"logsig{i}"]).float(),
         torch.from_numpy(LN[f"x{i}"][:, 0, :].copy()).float())
    y = torch.from_numpy(LN[f"y{i}"]).float()
    import contextlib

    with (contextlib.nullcontext() if fused_d else _lib.kernel_flags("generic")):
        loss, grads = (T.classification_loss if cls else T.regression_loss)(model, X, y)
    want = float(LN[f"loss{i}"])
    assert abs(float(loss) - want) < TOL * max(1.0, abs(want))
    assert rel_err(grads.numpy(), LN[f"grad{i}"]) < 5 * TOL


def test_make_step_with_emulated_adam_kernel(emulated):
    """train.make_step end to end (loss, adjoint, fused Adam kernel K4 with its device-side step counter) against
    the oracle's optax-style update."""
    from lncde import train as T
    from lncde.models.generate_model import create_model

    i = 1
    B, L, C_, s, depth, H, width, nh, cls, ostep, ld = (int(v) for v in N[f"cfg{i}"])
    dt0, scale, lambd = (float(v) for v in N[f"hyper{i}"])
    ts = _ts(L)
    model, _ = create_model("log_ncde", C_, N[f"logsig{i}"].shape[-1], depth, _intervals(ts, s), ld, H, vf_depth=nh,
                            vf_width=width, classification=True, dt0=dt0, scale=scale, lambd=lambd, key=1, device="cpu")
    flat = np.concatenate([N[f"mlp{i}"], N[f"lin1_{i}"], N[f"lin2_{i}"]])
    with torch.no_grad():
        model.params.copy_(torch.from_numpy(flat).float())
    X = (torch.from_numpy(np.tile(ts, (B, 1))), torch.from_numpy(N[f"logsig{i}"]).float(),
         torch.from_numpy(N[f"x{i}"][:, 0, :].copy()).float())
    y = torch.from_numpy(N[f"y{i}"]).float()
    opt = T.Adam(model, 1e-2)
    p0 = model.params.detach().clone().double()
    _, _, _, value = T.make_step(model, None, X, y, T.classification_loss, None, opt, None, None)
    assert abs(float(value) - float(N[f"loss{i}"])) < TOL
    g = torch.from_numpy(N[f"grad{i}"])
    # first Adam step (optax.adam, bias-corrected): p - lr * g / (|g| + eps)
    want = p0 - 1e-2 * g / (g.abs() + 1e-8)
    big = g.abs() > 1e-6  # elements whose update is not dominated by eps / fp32 rounding of g
    assert rel_err(model.params.detach().double()[big].numpy(), want[big].numpy()) < 1e-4
    assert int(opt.count) == 1


def test_calc_paths_empty_batch_and_nan_containment(emulated):
    """Host twin of tests/test_edge_cases_gpu.py::test_logsig_empty_batch_and_nan_containment."""
    from lncde.data_dir.generate_paths import batch_calc_paths, calc_paths

    for d, D in ((1, 4), (2, 7), (3, 15)):
        assert tuple(calc_paths(torch.zeros(0, 10, 3), 4, d).shape) == (0, 3, D)
    assert tuple(batch_calc_paths(torch.zeros(0, 10, 3), 4, 2).shape) == (0, 3, 7)
    x = torch.randn(3, 50, 4, generator=torch.Generator().manual_seed(0))
    clean = calc_paths(x, 5, 2)
    x[1, 22, 2] = float("nan")
    got = calc_paths(x, 5, 2)
    bad = torch.isnan(got).any(-1)
    assert bad[1, 4] and int(bad.sum()) == 1
    assert torch.equal(got[~bad], clean[~bad])


def test_saved_forward_scratch_is_not_reused_after_another_solve(emulated):
    """Two training-mode forwards, then the backwards in FORWARD order, and an NRDE adjoint in between: the
    shared grow-only scratch holds the saved intermediates of the LAST forward only, every other adjoint must fall back
    to recompute mode.  Compared with gradients taken one solve at a time."""
    from lncde.models import LogNeuralCDEs as M
    from lncde.models.NeuralCDEs import _NRDESolve
    from oracle import log_ncde as O
    from oracle import nrde as ON
    from oracle import paths

    B, L, C_, s, depth, H, width, nh = 2, 240, 7, 12, 2, 128, 32, 2  # specialised kernels: saved-forward mode exists
    rng = np.random.default_rng(3)
    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    dims = (H, C_, width, nh, depth)

    def problem(seed):
        r = np.random.default_rng(seed)
        x = np.cumsum(r.normal(scale=0.3, size=(B, L, C_)), axis=1).astype(np.float32)
        ls = torch.from_numpy(paths.calc_paths(x, s, depth, np.float32))
        rows, sc, grid = O.stage_table(ts, iv, 0.05, ls.shape[1])
        flat = O.flatten_params(O.init_params(H, C_, width, nh, scale=3.0, seed=seed, dtype=torch.float32))
        y0 = torch.from_numpy(r.normal(size=(B, H)).astype(np.float32) * 0.5)
        g = torch.from_numpy(r.normal(size=(B, len(grid), H)).astype(np.float32))
        return flat, ls, y0, torch.from_numpy(rows), torch.from_numpy(sc), g

    def grads_alone(p):
        flat, ls, y0, rows, sc, g = p
        f, y = flat.clone().requires_grad_(True), y0.clone().requires_grad_(True)
        (M._LogNCDESolve.apply(f, ls, y, rows, sc, dims) * g).sum().backward()
        return f.grad.clone(), y.grad.clone()

    pa, pb = problem(1), problem(2)
    want_a, want_b = grads_alone(pa), grads_alone(pb)
    fa, ya = pa[0].clone().requires_grad_(True), pa[2].clone().requires_grad_(True)
    fb, yb = pb[0].clone().requires_grad_(True), pb[2].clone().requires_grad_(True)
    out_a = M._LogNCDESolve.apply(fa, pa[1], ya, pa[3], pa[4], dims)
    out_b = M._LogNCDESolve.apply(fb, pb[1], yb, pb[3], pb[4], dims)  # overwrites the intermediates saved for A
    (out_a * pa[5]).sum().backward()  # must recompute (and thereby clobbers what B saved)
    (out_b * pb[5]).sum().backward()  # must recompute too
    for got, want in ((fa.grad, want_a[0]), (ya.grad, want_a[1]), (fb.grad, want_b[0]), (yb.grad, want_b[1])):
        assert rel_err(got.numpy(), want.numpy()) < 1e-6

    # an NRDE adjoint between a Log-NCDE forward and its backward uses the same scratch
    fa.grad = ya.grad = None
    out_a = M._LogNCDESolve.apply(fa, pa[1], ya, pa[3], pa[4], dims)
    Dm = pa[1].shape[2] - 1
    nflat = ON.flatten_params(ON.init_params(16, Dm, 8, 2, scale=1.0, seed=4, dtype=torch.float32)).requires_grad_(True)
    ny0 = torch.from_numpy(rng.normal(size=(B, 16)).astype(np.float32)).requires_grad_(True)
    _NRDESolve.apply(nflat, pa[1], ny0, pa[3], pa[4], (16, 8, 2)).sum().backward()
    (out_a * pa[5]).sum().backward()
    assert rel_err(fa.grad.numpy(), want_a[0].numpy()) < 1e-6 and rel_err(ya.grad.numpy(), want_a[1].numpy()) < 1e-6


@pytest.mark.parametrize("lambd", [0.0, 0.01])
@pytest.mark.parametrize("dims", [(3, 64, 3, 4, 2, 16, 8, 2, 3), (2, 96, 7, 12, 2, 128, 32, 2, 5), (5, 40, 2, 4, 1, 8, 8, 1, 2),
                                  (70, 24, 2, 4, 1, 8, 8, 1, 2)])  # the last: batch sums in three chunks of 32 samples
def test_fused_classification_step_equals_autograd_step(emulated, dims, lambd):
    """LogNeuralCDE.fused_classification_loss (K5 kernels: initial state, read-out + cross-entropy, Lip(2), linear1 /
    linear2 gradients around K3 / K3b, no torch autograd) against the autograd formulation of train.classification_loss
    (`kernel_flags("generic")`) and against the fp64 oracle loss."""
    from lncde import _lib, train as T
    from lncde.models.generate_model import create_model
    from oracle import log_ncde as O
    from oracle import paths

    B, L, C_, s, depth, H, width, nh, nl = dims
    rng = np.random.default_rng(B)
    x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C_)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float32)
    ts = paths.make_ts(L)
    iv = paths.make_intervals(ts, s)
    model, _ = create_model("log_ncde", C_, logsig.shape[-1], depth, iv, nl, H, vf_depth=nh, vf_width=width, dt0=0.05,
                            scale=2.0, lambd=lambd, key=3, device="cpu")
    X = (torch.from_numpy(np.tile(ts, (B, 1))), torch.from_numpy(logsig), torch.from_numpy(x[:, 0, :].copy()))
    y = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, nl, size=B)), nl).float()
    with _lib.kernel_flags("generic"):
        l0, g0 = T.classification_loss(model, X, y)
    l1, g1 = T.classification_loss(model, X, y)
    l2, g2 = T.classification_loss(model, X, y)  # second call: cached buffers (zero g_ys rows) are reused
    assert abs(float(l1) - float(l0)) < 1e-6 * max(1.0, abs(float(l0)))
    assert rel_err(g1.numpy(), g0.numpy()) < 2e-6
    assert torch.equal(l1, l2) and torch.equal(g1, g2)
    # every parameter block on its own (a wrong small block would hide behind the largest one)
    n_mlp, o2 = model.n_mlp, model.n_mlp + H * (C_ + 1)
    for a, b in ((0, n_mlp), (n_mlp, n_mlp + H * C_), (n_mlp + H * C_, o2), (o2, o2 + nl * H), (o2 + nl * H, g0.numel())):
        assert rel_err(g1[a:b].numpy(), g0[a:b].numpy()) < 1e-5, (a, b)


@pytest.mark.parametrize("name,bs,lambd", [("diagonal_linear_ncde", 1, 0.1), ("bd_linear_ncde", 4, 0.1),
                                            ("dense_linear_ncde", 8, 0.0), ("bd_linear_ncde", 2, 0.0),
                                            ("dense_linear_ncde", 32, 0.1)])  # rows of 1024 entries: CTA-wide Lip(2) path
def test_fused_lncde_classification_step_equals_autograd_step(emulated, name, bs, lambd):
    """LogLinearCDE.fused_classification_loss (K5 kernels around the K2 scans; autograd only through log_ode) against the
    autograd formulation under `kernel_flags("generic")` - the generic run also uses the baseline scans, so this covers
    the default kernels (fused diagonal scan, time-parallel scans need T > 65) against the baseline ones as well."""
    from lncde import _lib, train as T
    from lncde.models.generate_model import create_model
    from oracle import paths

    B, L, C_, s, depth, H, nl = 3, 280, 3, 4, 2, max(8, bs), 3
    rng = np.random.default_rng(7)
    x = np.cumsum(rng.normal(scale=0.1, size=(B, L, C_)), axis=1).astype(np.float32)
    logsig = paths.calc_paths(x, s, depth, np.float32)
    model, _ = create_model(name, C_, logsig.shape[-1], depth, None, nl, H, block_size=bs, lambd=lambd, key=5, device="cpu")
    X = (None, torch.from_numpy(logsig), torch.from_numpy(x[:, 0, :].copy()))
    y = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, nl, size=B)), nl).float()
    with _lib.kernel_flags("generic"):
        l0, g0 = T.classification_loss(model, X, y)
    l1, g1 = T.classification_loss(model, X, y)
    assert abs(float(l1) - float(l0)) < 2e-6 * max(1.0, abs(float(l0)))
    nA, o2 = model.n_A, model.n_A + H * (C_ + 1)
    for a, b in ((0, nA), (nA, nA + H * C_), (nA + H * C_, o2), (o2, o2 + nl * H), (o2 + nl * H, g0.numel())):
        assert rel_err(g1[a:b].numpy(), g0[a:b].numpy()) < 1e-5, (name, a, b)


@pytest.mark.parametrize("C_,depth,nb,bs", [(3, 2, 2, 4), (3, 3, 1, 8), (2, 4, 3, 5), (7, 2, 1, 40), (2, 3, 1, 33)])
def test_lie_bracket_kernels_equal_torch_log_ode(emulated, C_, depth, nb, bs):
    """K2a (lncde_lie_brackets_fwd / bwd) behind LogLinearCDE.log_ode against the batched-matmul formulation it replaces
    (`kernel_flags("generic")`) and torch autograd through it: all degrees up to 4, several blocks, block sizes below,
    at and above the 32 x 32 tile, parents with children on several levels (gather order of the reverse pass)."""
    from lncde import _lib
    from lncde.models.generate_model import create_model

    H = nb * bs
    model, _ = create_model("bd_linear_ncde", C_, 1, depth, None, 2, H, block_size=bs, lambd=0.0, key=2, device="cpu")
    g = torch.Generator().manual_seed(C_ * 10 + depth)
    vfs = (torch.randn(C_, nb, bs, bs, generator=g) / bs ** 0.5).requires_grad_(True)
    with _lib.kernel_flags("generic"):
        want = model.log_ode(vfs)
    got = model.log_ode(vfs)
    assert got.shape == want.shape == (nb, bs, bs, len(model._basis_pairs) - 1)
    assert rel_err(got.detach().numpy(), want.detach().numpy()) < 2e-6
    w = torch.randn(want.shape, generator=g)
    (g_want,) = torch.autograd.grad(want, vfs, w)
    (g_got,) = torch.autograd.grad(got, vfs, w)
    assert rel_err(g_got.numpy(), g_want.numpy()) < 5e-6
