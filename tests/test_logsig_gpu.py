"""K1 parity: CUDA calc_paths (through the C ABI) vs the oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import paths

pytestmark = pytest.mark.gpu


def _toy(B, L, C, seed=1234):
    rng = np.random.default_rng(seed)
    return np.cumsum(np.round(rng.normal(size=(B, L, C))), axis=1).astype(np.float32)


def _walk(B, L, C, seed=2345, time=True):
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.normal(scale=0.05, size=(B, L, C)), axis=1)
    if time:
        x[..., 0] = np.arange(L)[None, :] / L
    return x.astype(np.float32)


def _gpu(x, s, d, **kw):
    from lncde.data_dir.generate_paths import calc_paths

    out = calc_paths(torch.from_numpy(x).cuda(), s, d, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("C", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("depth", [1, 2])
def test_integer_kat_bit_exact(built_lib, C, depth):
    # integer random walks: every partial sum is exact in fp32 -> bit-exact parity
    x = _toy(9, 100, C)
    for s in (1, 4, 7, 100, 101, 500):
        want = paths.calc_paths(x, s, depth, np.float64)
        got = _gpu(x, s, depth)
        assert got.shape == want.shape
        assert np.array_equal(got.astype(np.float64), want), (C, depth, s)


def test_toy_config_shape(built_lib):
    x = _toy(32, 100, 6)
    got = _gpu(x, 4, 2)
    assert got.shape == (32, 26, 22)
    assert np.array_equal(got, paths.calc_paths(x, 4, 2, np.float32))


@pytest.mark.parametrize("compat", [True, False])
def test_compat_and_chunk(built_lib, compat):
    x = _toy(300, 30, 3)
    want = paths.batch_calc_paths(x, 4, 2, np.float64, compat=compat)
    from lncde.data_dir.generate_paths import batch_calc_paths

    got = batch_calc_paths(torch.from_numpy(x).cuda(), 4, 2, compat=compat).cpu().numpy()
    assert np.array_equal(got.astype(np.float64), want)


@pytest.mark.parametrize("C,s,L", [(7, 12, 17984), (6, 12, 17984), (6, 16, 896), (7, 16, 896), (7, 10, 4992), (4, 10, 4992),
                                   (7, 10, 49920), (4, 10, 49920)])  # the last two: PPG-DaLiA at its full length
def test_float_paths_depth2(built_lib, C, s, L):
    x = _walk(4, L, C)
    want = paths.calc_paths(x, s, 2, np.float64)
    got = _gpu(x, s, 2)
    assert got.shape == want.shape
    # tolerance: BASELINE north_star 1e-5 relative (norm-relative per tensor), fp64 oracle as arbiter
    assert rel_err(got, want) < 1e-5
    # and not worse than the fp32 oracle (reference formulation) by more than a small factor
    o32 = paths.calc_paths(x, s, 2, np.float32)
    assert rel_err(got, want) <= max(4 * rel_err(o32, want), 1e-6)


@pytest.mark.parametrize("C,depth", [(2, 3), (3, 3), (6, 3), (7, 3), (2, 4), (3, 4), (10, 2), (16, 2), (9, 1)])
def test_generic_tensor_kernel(built_lib, C, depth):
    x = _toy(5, 50, C)
    for s in (4, 12, 51):
        want = paths.calc_paths(x, s, depth, np.float64)
        got = _gpu(x, s, depth)
        assert got.shape == want.shape
        assert rel_err(got, want) < 1e-5, (C, depth, s)
    xf = _walk(3, 200, C, time=False)
    want = paths.calc_paths(xf, 12, depth, np.float64)
    assert rel_err(_gpu(xf, 12, depth), want) < 1e-5


def test_large_step_falls_back_to_generic(built_lib):
    x = _toy(3, 2000, 7)
    want = paths.calc_paths(x, 700, 2, np.float64)
    assert rel_err(_gpu(x, 700, 2), want) < 1e-6


def test_full_size_properties(built_lib):
    """BASELINE full size (EigenWorms shape, B=32): size-independent properties."""
    x = _walk(32, 17984, 7)
    xt = torch.from_numpy(x).cuda()
    from lncde.data_dir.generate_paths import calc_paths

    ls = calc_paths(xt, 12, 2, compat=False)
    assert ls.shape == (32, 1499, 29)
    assert torch.all(ls[..., 0] == 0)
    # level 1 telescopes: sum over windows = x[L-1] - 0
    tot = ls[..., 1:8].double().sum(1)
    assert torch.allclose(tot, xt[:, -1].double(), atol=1e-4)
    # time-reversal antisymmetry of Levy areas is covered by linearity: scaling the path by a
    # scales level 1 by a and level 2 by a^2
    ls2 = calc_paths(2 * xt, 12, 2, compat=False)
    assert torch.allclose(ls2[..., 1:8], 2 * ls[..., 1:8], rtol=0, atol=1e-6)
    assert torch.allclose(ls2[..., 8:], 4 * ls[..., 8:], rtol=0, atol=1e-6)
