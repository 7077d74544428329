"""Independent cross-checks of the third-party semantics the oracles restate (signax, diffrax; see oracle/__init__.py): every restatement of a
third-party algorithm is compared with a second, structurally different derivation of the
same mathematics, so an error in the restatement cannot hide."""
import itertools

import numpy as np
import torch

from oracle import linear_ncde as OL
from oracle import log_ncde as O
from oracle import tensor_algebra as ta
from oracle.hall import Hall


# ----------------------------------------------------------------------------- signature
def _brute_force_signature(path, depth):
    """Iterated integrals of a piecewise-linear path by explicit ordered sums (no Chen, no Horner):
    level n = sum over k1 <= k2 <= ... <= kn of  (1/m1! m2! ...) D_k1 (x) ... (x) D_kn  where equal
    indices are grouped (m_i = multiplicities)."""
    inc = np.diff(path, axis=0)
    n, C = inc.shape
    levels = []
    for d in range(1, depth + 1):
        out = np.zeros((C,) * d)
        for ks in itertools.combinations_with_replacement(range(n), d):
            mult = 1.0
            for _, grp in itertools.groupby(ks):
                m = len(list(grp))
                mult *= float(np.prod(np.arange(1, m + 1)))
            term = inc[ks[0]]
            for k in ks[1:]:
                term = np.multiply.outer(term, inc[k])
            out += term / mult
        levels.append(out)
    return levels


def test_signature_matches_brute_force_iterated_integrals():
    rng = np.random.default_rng(0)
    path = rng.normal(size=(6, 3))
    got = ta.signature(path, 3)
    want = _brute_force_signature(path, 3)
    for g, w in zip(got, want):
        assert np.allclose(g, w, atol=1e-12)


def test_log_is_inverse_of_exp():
    """exp(log(1+S)) - 1 == S with exp computed by its own power series."""
    rng = np.random.default_rng(1)
    S = ta.signature(rng.normal(size=(5, 3)), 4)
    lg = ta.log(S)
    # exp(L) - 1 = L + L^2/2 + L^3/6 + L^4/24
    acc = [np.zeros_like(x) for x in lg]
    power = [x.copy() for x in lg]
    fact = 1.0
    for n in range(1, 5):
        fact *= n
        acc = [a + p / fact for a, p in zip(acc, power)]
        power = ta.tensor_mul(power, lg)
    for a, s in zip(acc, S):
        assert np.allclose(a, s, atol=1e-12)


def test_log_signature_is_a_lie_element():
    """log S lies in the free Lie algebra: l2t(t2l(flat)) == flat (projection then expansion)."""
    rng = np.random.default_rng(2)
    for C, d in [(2, 4), (3, 3), (4, 2)]:
        flat = ta.flatten(ta.log(ta.signature(rng.normal(size=(7, C)), d)))
        h = Hall(C, d)
        coords = h.t2l_matrix(d, np.float64)[:, 1:] @ flat
        back = h.l2t_matrix(d, np.float64) @ coords
        assert np.allclose(back[1:], flat, atol=1e-11), (C, d)


def test_bch_depth2_for_two_segments():
    """log(exp(a) exp(b)) = a + b + [a,b]/2 at depth 2 (Baker-Campbell-Hausdorff)."""
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=3), rng.normal(size=3)
    path = np.stack([np.zeros(3), a, a + b])
    lg = ta.log(ta.signature(path, 2))
    assert np.allclose(lg[0], a + b)
    assert np.allclose(lg[1], 0.5 * (np.outer(a, b) - np.outer(b, a)))


# ----------------------------------------------------------------------------- Log-NCDE field
def _setup_field(C=3, H=5, width=4, n_hidden=2, seed=0):
    layers = O.init_params(H, C, width, n_hidden, scale=1.0, seed=seed, dtype=torch.float64)
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(H, generator=g, dtype=torch.float64)
    D = len(Hall(C, 2))
    ls = torch.randn(D, generator=g, dtype=torch.float64)
    ls[0] = 0
    return layers, y, ls, C, H


def test_field_equals_lie_brackets_from_full_jacobians():
    """models/LogNeuralCDEs.py:123-138 (C JVPs, antisymmetrised) == sum_i l_i f_i + sum_{i<j} l_ij [f_i, f_j]
    with [f_i, f_j] = Df_j f_i - Df_i f_j from the explicit Jacobian."""
    layers, y, ls, C, H = _setup_field()
    hall = Hall(C, 2)
    pairs = torch.from_numpy(hall.pairs().astype(np.int64))
    got = O.field(layers, y, ls, C, H, 2, pairs)
    f = lambda yy: O.vf_apply(layers, yy).reshape(C, H)
    J = torch.autograd.functional.jacobian(f, y)  # [C, H, H] : J[i] = D f_i
    fv = f(y)
    want = ls[1 : C + 1] @ fv
    for p, (i, j) in enumerate(hall.pairs()[C:]):
        i, j = int(i) - 1, int(j) - 1
        want = want + ls[1 + C + p] * (J[j] @ fv[i] - J[i] @ fv[j])
    assert torch.allclose(got, want, atol=1e-12)


def test_restructured_field_used_by_the_kernels_is_identical():
    """G = sum_i l_i f_i + sum_j [Df . w_j]_{block j}, w = Lam^T f  (csrc/logncde_common.cuh)
    and its 'Lam through the first layer' variant (csrc/logncde_fast.cuh) equal the reference formula."""
    layers, y, ls, C, H = _setup_field(C=4, H=6, width=5, n_hidden=2, seed=3)
    pairs = torch.from_numpy(Hall(C, 2).pairs().astype(np.int64))
    ref = O.field(layers, y, ls, C, H, 2, pairs)
    Lam = torch.zeros(C, C, dtype=torch.float64)
    p = 0
    for i in range(C):
        for j in range(i + 1, C):
            Lam[i, j] = ls[1 + C + p]
            Lam[j, i] = -ls[1 + C + p]
            p += 1
    vf = lambda yy: O.vf_apply(layers, yy)
    fv = vf(y).reshape(C, H)
    w = Lam.T @ fv  # w_j = sum_i Lam_ij f_i
    G = ls[1 : C + 1] @ fv
    for j in range(C):
        jv = torch.func.jvp(vf, (y,), (w[j],))[1].reshape(C, H)
        G = G + jv[j]
    assert torch.allclose(G, ref, atol=1e-12)
    # variant: first layer applied to f_i, then mixed (linearity of W1)
    (W1, b1), (W2, b2), (W3, b3) = layers
    z1 = W1 @ y + b1
    s1 = torch.sigmoid(z1) * (1 + z1 * (1 - torch.sigmoid(z1)))
    a1 = z1 * torch.sigmoid(z1)
    z2 = W2 @ a1 + b2
    s2 = torch.sigmoid(z2) * (1 + z2 * (1 - torch.sigmoid(z2)))
    a2 = z2 * torch.sigmoid(z2)
    o = torch.tanh(W3 @ a2 + b3).reshape(C, H)
    q = o @ W1.T                       # q_i = W1 f_i
    u1 = Lam.T @ q                     # u1_j = sum_i Lam_ij q_i
    u2 = (u1 * s1) @ W2.T
    p2 = u2 * s2
    G2 = ls[1 : C + 1] @ o
    W3b = W3.reshape(C, H, -1)
    for j in range(C):
        G2 = G2 + (1 - o[j] ** 2) * (W3b[j] @ p2[j])
    assert torch.allclose(G2, ref, atol=1e-12)


def test_heun_solve_against_direct_recurrence():
    """O.solve == hand-written Heun steps with the stage table (k1 at t_n, k2 at t_{n+1}, y + (k1+k2)/2)."""
    layers, y, ls, C, H = _setup_field()
    rng = np.random.default_rng(0)
    K = 6
    logsig = torch.from_numpy(rng.normal(scale=0.3, size=(1, K, len(Hall(C, 2)))))
    logsig[..., 0] = 0
    rows = np.array([[5, 0], [0, 1], [1, 3]], np.int32)
    sc = np.array([[-0.1, 0.4], [0.4, 0.4], [0.4, 0.2]], np.float32)
    ys = O.solve(layers, logsig, y[None], rows, sc, C, H, 2)[0]
    pairs = torch.from_numpy(Hall(C, 2).pairs().astype(np.int64))
    cur = y
    for n in range(3):
        k1 = float(sc[n, 0]) * O.field(layers, cur, logsig[0, rows[n, 0]], C, H, 2, pairs)
        k2 = float(sc[n, 1]) * O.field(layers, cur + k1, logsig[0, rows[n, 1]], C, H, 2, pairs)
        cur = cur + 0.5 * (k1 + k2)
        assert torch.allclose(ys[n + 1], cur, atol=1e-12)


# ----------------------------------------------------------------------------- LNCDE
def test_log_ode_brackets_and_structured_cases():
    rng = torch.Generator().manual_seed(0)
    C, bs = 3, 4
    vf = torch.randn(C, bs, bs, generator=rng, dtype=torch.float64)
    h = Hall(C, 3)
    A = OL.log_ode(vf, h)
    assert A.shape == (bs, bs, len(h) - 1)
    for k in range(1, len(h)):
        l, r = h.basis[k]
        if l == 0:
            assert torch.equal(A[:, :, k - 1], vf[r - 1])
        else:  # A_[u,v] = A_v A_u - A_u A_v  (LinearNeuralCDEs.py:225-227)
            Au, Av = A[:, :, l - 1], A[:, :, r - 1]
            assert torch.allclose(A[:, :, k - 1], Av @ Au - Au @ Av, atol=1e-13)
    # commuting (diagonal) fields have vanishing brackets: D-LNCDE needs level 1 only (quirk B7)
    vd = torch.diag_embed(torch.randn(C, bs, generator=rng, dtype=torch.float64))
    Ad = OL.log_ode(vd, h)
    assert torch.allclose(Ad[:, :, C:], torch.zeros_like(Ad[:, :, C:]), atol=1e-14)


def test_diagonal_branch_equals_block_branch_with_diagonal_blocks():
    """The bs = 1 branch (:302-316) == the generic block branch (:318-353) fed diagonal matrices."""
    g = torch.Generator().manual_seed(1)
    C, H, T, depth = 3, 4, 9, 2
    D = len(Hall(C, depth))
    ls = torch.randn(T, D, generator=g, dtype=torch.float64) * 0.2
    ls[:, 0] = 0
    vfd = torch.randn(C, H, generator=g, dtype=torch.float64)
    y0 = torch.randn(H, generator=g, dtype=torch.float64)
    a = OL.forward(vfd, ls, y0, C, 1, depth, 1)
    vfb = torch.diag_embed(vfd).reshape(C, H * H)  # one H x H block per channel, diagonal
    b = OL.forward(vfb, ls, y0, C, H, depth, 1)
    assert torch.allclose(a, b, atol=1e-12)
    # chunked (parallel_steps) scanner == sequential scanner, incl. the remainder path
    for P in (2, 3, 4, 128):
        assert torch.allclose(OL.forward(vfb, ls, y0, C, H, depth, P), b, atol=1e-12)


def test_adam_update_against_closed_form_first_step():
    from oracle.train import adam_update

    p = torch.tensor([1.0, -2.0], dtype=torch.float64)
    g = torch.tensor([0.5, -0.25], dtype=torch.float64)
    m, v = torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    adam_update(p, g, m, v, 1, lr=0.1)
    # first step: m_hat = g, v_hat = g^2  ->  p -= lr * g / (|g| + eps)
    assert torch.allclose(p, torch.tensor([1.0 - 0.1, -2.0 + 0.1], dtype=torch.float64), atol=1e-7)


def test_nrde_field_solve_and_gradient_independent():
    """oracle/nrde.py against a second derivation: the field as an explicit 3-index contraction, the Heun recurrence
    written out per sample, a depth-1 NRDE with a linear vector field solved in closed form for one step, and the
    reverse-mode gradient against central finite differences (fp64)."""
    from oracle import nrde as ON

    rng = np.random.default_rng(9)
    H, Dm, w, n_vf, B, K = 5, 4, 8, 2, 2, 3
    layers = ON.init_params(H, Dm, w, n_vf, scale=2.0, seed=4, dtype=torch.float64)
    ls = torch.from_numpy(rng.normal(size=(B, K, Dm + 1)) * 0.3)
    ls[..., 0] = 0
    y0 = torch.from_numpy(rng.normal(size=(B, H)))
    rows = np.array([[2, 0], [0, 1], [1, 2]], np.int32)
    sc = np.array([[0.3, 0.5], [0.5, 0.4], [0.4, 0.2]], np.float32)
    # field: reshape(W a + b, (H, Dm)) @ l  ==  sum_{d,k} W[h,d,k] l_d a_k + sum_d b[h,d] l_d
    (W0, b0), (W1, b1), (Wm, bm) = layers
    y = y0[0]
    a = torch.tanh(W1 @ torch.relu(W0 @ y + b0) + b1)
    want = torch.einsum("hdk,d,k->h", Wm.view(H, Dm, w), ls[0, 1, 1:], a) + bm.view(H, Dm) @ ls[0, 1, 1:]
    assert torch.allclose(ON.field(layers, y, ls[0, 1], H), want, rtol=1e-13, atol=1e-14)
    # Heun recurrence, one sample at a time
    ys = ON.solve(layers, ls, y0, rows, sc, H)
    for b in range(B):
        yb = y0[b]
        for n in range(rows.shape[0]):
            k1 = float(sc[n, 0]) * ON.field(layers, yb, ls[b, rows[n, 0]], H)
            k2 = float(sc[n, 1]) * ON.field(layers, yb + k1, ls[b, rows[n, 1]], H)
            yb = yb + 0.5 * (k1 + k2)
            assert torch.allclose(ys[b, n + 1], yb, rtol=1e-12, atol=1e-13)
    # gradient vs central differences on a few coordinates of every leaf
    leaves = [t.clone().requires_grad_(True) for Wb in layers for t in Wb]
    lay = list(zip(leaves[0::2], leaves[1::2]))
    g_ys = torch.from_numpy(rng.normal(size=tuple(ys.shape)))
    loss = lambda L: (ON.solve(L, ls, y0, rows, sc, H) * g_ys).sum()
    grads = torch.autograd.grad(loss(lay), leaves)
    eps = 1e-6
    for li, (t, g) in enumerate(zip(leaves, grads)):
        flat = t.detach().reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            def at(delta):
                mod = [x.detach().clone() for x in leaves]
                mod[li].reshape(-1)[idx] += delta
                return float(loss(list(zip(mod[0::2], mod[1::2]))))
            fd = (at(eps) - at(-eps)) / (2 * eps)
            assert abs(fd - float(g.reshape(-1)[idx])) < 1e-6 * max(1.0, abs(fd)), (li, idx)
