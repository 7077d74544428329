"""Oracle: structured linear CDE (D / BD / DE-LNCDE) on CPU (torch).  TEST INFRASTRUCTURE ONLY.

Restates ``models/LinearNeuralCDEs.py``: ``log_ode`` (:192-232), the diagonal branch
(:302-316), the block branch (:318-353), the scanner incl. the parallel_steps chunk /
remainder logic (:355-386) and the read-outs (:389-394).  The basis list comes from the
Hall set (roughpy 0.2.0, un-vendored, enumerates the same Hall basis – equal by inspection
at depth <= 2, assumed at depth 3: that ORDER is unpinned).
PINNED to the reference's own LogLinearCDE.__call__ / log_ode / losses, executed from /root/reference
under library stand-ins (tests/golden/ref_linear.npz, tests/test_oracle_reference_pinned.py)."""
from __future__ import annotations

import torch

from .hall import Hall


def log_ode(vf: torch.Tensor, hall: Hall) -> torch.Tensor:
    """vf [C, bs, bs] -> [bs, bs, n_basis]; A_[u,v] = A_v A_u - A_u A_v (:225-227)."""
    A = [None] * len(hall.basis)
    for k in range(1, len(hall.basis)):
        l, r = hall.basis[k]
        if l == 0:
            A[k] = vf[r - 1]
        else:
            A[k] = A[r] @ A[l] - A[l] @ A[r]
    return torch.stack(A[1:], dim=2)


def _assoc_scan(comp, xs):
    """Sequential emulation of jax.lax.associative_scan (same result in exact arithmetic)."""
    out = [xs[0]]
    for x in xs[1:]:
        out.append(comp(out[-1], x))
    return out


def forward(vf_A, logsigs, y0, data_dim, block_size, depth, parallel_steps=1):
    """One sample.  vf_A [C, nb*bs*bs]; logsigs [T, D]; y0 [H] -> ys [T, H]."""
    C, bs = data_dim, block_size
    T = logsigs.shape[0]
    H = y0.shape[0]
    if bs == 1:
        flows = 1 + logsigs[:, 1 : C + 1] @ vf_A  # (:303-305), quirk B7
        step = lambda y, f: y * f
        comp = lambda f1, f2: f1 * f2
        apply_total = lambda tot, y: y * tot
    else:
        nb = H // bs
        hall = Hall(C, depth)
        vfs = vf_A.reshape(C, nb, bs, bs)
        lie = torch.stack([log_ode(vfs[:, b], hall) for b in range(nb)])  # (nb, bs, bs, n_basis)
        flows = torch.einsum("ijkl,ml->mijk", lie, logsigs[:, 1:]) + torch.eye(bs, dtype=vf_A.dtype)
        step = lambda y, f: (f @ y.reshape(nb, bs, 1)).reshape(H)
        comp = lambda x, y: y @ x
        apply_total = lambda tot, y: (tot @ y.reshape(nb, bs, 1)).reshape(H)
    ys = [y0]
    if parallel_steps == 1:
        for t in range(1, T):
            ys.append(step(ys[-1], flows[t]))
    else:
        P = parallel_steps
        rem = (T - 1) % P
        core_end = T - rem
        for c0 in range(1, core_end, P):
            tot = _assoc_scan(comp, [flows[t] for t in range(c0, c0 + P)])
            y_in = ys[-1]
            ys.extend(apply_total(tt, y_in) for tt in tot)
        for t in range(core_end, T):
            ys.append(step(ys[-1], flows[t]))
    return torch.stack(ys)


def predict(vf_A, W_init, b_init, W_out, b_out, logsigs, x0, data_dim, block_size, depth, classification=True,
            parallel_steps=1):
    """Batched LogLinearCDE.__call__ (:234-396).  logsigs [B,T,D], x0 [B,C]."""
    outs = []
    for b in range(logsigs.shape[0]):
        y0 = W_init @ x0[b] + b_init
        ys = forward(vf_A, logsigs[b], y0, data_dim, block_size, depth, parallel_steps)
        if classification:
            outs.append(torch.softmax(W_out @ ys.mean(0) + b_out, -1))
        else:
            outs.append(torch.tanh(ys @ W_out.T + b_out))
    return torch.stack(outs)


def lip2_norm(vf_A, lambd):
    """train.py:87-88."""
    return lambd * torch.mean(torch.linalg.norm(vf_A, dim=-1))
