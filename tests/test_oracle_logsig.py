"""Identities that anchor the signature/log restatement (signax itself is absent offline) and calc_paths."""
import numpy as np
import pytest

from oracle import paths, tensor_algebra as ta
from oracle.hall import Hall


def _toy(B, L, C, seed=1234):
    rng = np.random.default_rng(seed)
    return np.cumsum(np.round(rng.normal(size=(B, L, C))), axis=1)  # toy_dataset.py:17-19


def test_single_increment_log_is_increment():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 2, 4))
    lg = ta.log(ta.signature(x, 4))
    assert np.allclose(lg[0], x[:, 1] - x[:, 0])
    for lv in lg[1:]:
        assert np.abs(lv).max() < 1e-12


def test_chen_identity():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(9, 3))
    S = ta.signature(x, 3)
    A, Bs = ta.signature(x[:5], 3), ta.signature(x[4:], 3)
    # (1+A)(1+B) - 1 = A + B + AB
    AB = ta.tensor_mul(A, Bs)
    for s, a, b, ab in zip(S, A, Bs, AB):
        assert np.allclose(s, a + b + ab, atol=1e-12)


@pytest.mark.parametrize("C,s", [(6, 4), (7, 12), (3, 5)])
def test_depth2_equals_levy_area(C, s):
    x = _toy(4, 50, C)
    o = paths.calc_paths(x, s, 2, np.float64, compat=False)
    xt = np.concatenate([np.zeros((4, 1, C)), x], 1)
    for k in range(1, (51) // s):
        win = xt[:, k * s - 1 : k * s + s]
        assert np.array_equal(paths.levy_logsig_depth2(win), o[:, k])


def test_integer_paths_are_exact_in_fp32():
    # depth-2 logsigs of integer random walks are exact multiples of 1/2
    x = _toy(8, 100, 6)
    o32 = paths.calc_paths(x, 4, 2, np.float32)
    o64 = paths.calc_paths(x, 4, 2, np.float64)
    assert o32.shape == (8, 26, 22)
    assert np.array_equal(o32.astype(np.float64), o64)
    assert np.array_equal(o64 * 2, np.round(o64 * 2))
    assert np.all(o64[..., 0] == 0)


def test_compat_quirk_b1():
    x = _toy(5, 100, 6)
    a = paths.calc_paths(x, 4, 2, np.float64, compat=True)
    b = paths.calc_paths(x, 4, 2, np.float64, compat=False)
    assert np.array_equal(a[:, :-1], b[:, :-1])
    # level 1 of the compat remainder row = x[n, L-1] - x[n-1, L-1] (0 for n = 0)
    prev = np.concatenate([np.zeros((1, 6)), x[:-1, -1]], 0)
    assert np.array_equal(a[:, -1, 1:7], x[:, -1] - prev)
    assert np.array_equal(b[:, -1, 1:7], x[:, -1] - x[:, -2])


def test_depth3_roundtrip_through_l2t():
    rng = np.random.default_rng(2)
    w = rng.normal(size=(4, 7, 3))
    h = Hall(3, 3)
    flat = ta.flatten(ta.log(ta.signature(w, 3)))
    ls = flat @ h.t2l_matrix(3, np.float64)[:, 1:].T
    back = ls @ h.l2t_matrix(3, np.float64).T
    assert np.allclose(back[:, 1:], flat, atol=1e-12)


def test_shapes_and_clamp():
    x = _toy(2, 10, 3)
    assert paths.calc_paths(x, 100, 2).shape == (2, 1, 7)  # stepsize clamped to L+1
    assert paths.calc_paths(x, 11, 1).shape == (2, 1, 4)
    assert paths.calc_paths(x, 1, 2).shape == (2, 11, 7)
    assert paths.batch_calc_paths(_toy(130, 9, 2), 4, 2).shape == (130, 3, 4)


def test_intervals():
    ts = paths.make_ts(100)
    iv = paths.make_intervals(ts, 4)
    assert iv.shape == (26,) and iv[-1] == 1.0 and iv[0] == 0.0
    assert paths.make_intervals(paths.make_ts(17984), 12).shape == (1500,)
