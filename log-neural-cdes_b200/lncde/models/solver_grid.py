"""Host tables for the fixed-step Heun solve (replaces the diffrax ConstantStepSize
loop driven from models/LogNeuralCDEs.py:147-157 and the per-stage
``searchsorted`` of :114-115,120,138).

All arithmetic is fp32, exactly as the reference's traced program does it, because
several stage times land on interval boundaries where ``searchsorted(side='left')``
flips the selected logsig row.
"""
from __future__ import annotations

import numpy as np


def step_grid(t0, t1, dt0, max_steps: int = 16**4) -> np.ndarray:
    """t_0 .. t_n of the solve: ``tnext = tprev + dt`` with the step repeated
    (ConstantStepSize) and clipped to t1 once within 1e-6 of it."""
    f = np.float32
    t0, t1 = f(t0), f(t1)
    grid = [t0]
    tprev, tnext = t0, f(t0 + f(dt0))
    while True:
        if tnext > t1 - f(1e-6):
            tnext = t1
        grid.append(tnext)
        if tnext >= t1:
            break
        if len(grid) > max_steps:
            raise RuntimeError("max_steps reached (diffrax would raise here)")
        tprev, tnext = tnext, f(tnext + f(tnext - tprev))
    return np.asarray(grid, np.float32)


def stage_table(t0, t1, intervals, dt0, n_rows: int, max_steps: int = 16**4):
    """(rows int32 [n,2], scale fp32 [n,2], grid fp32 [n+1]).

    Stage 1 of step k is evaluated at t_k, stage 2 at t_{k+1}.  For a stage time t:
    idx = searchsorted(intervals, t, 'left'); logsig row idx-1 and interval length
    intervals[idx] - intervals[idx-1], negative indices wrapping (quirk B2: t = 0
    selects the LAST row and the length -T).  scale = (t_{k+1} - t_k) / length."""
    intervals = np.asarray(intervals, np.float32)
    grid = step_grid(t0, t1, dt0, max_steps)
    n, ni = len(grid) - 1, len(intervals)
    rows = np.zeros((n, 2), np.int32)
    scale = np.zeros((n, 2), np.float32)
    idx_all = np.minimum(np.searchsorted(intervals, grid, side="left"), ni - 1)
    lo_all = idx_all - 1
    delta = (intervals[idx_all] - intervals[np.where(lo_all >= 0, lo_all, ni - 1)]).astype(np.float32)
    row_all = np.where(lo_all >= 0, lo_all, n_rows - 1).astype(np.int32)
    dt = (grid[1:] - grid[:-1]).astype(np.float32)
    rows[:, 0], rows[:, 1] = row_all[:-1], row_all[1:]
    scale[:, 0], scale[:, 1] = dt / delta[:-1], dt / delta[1:]
    return rows, scale, grid
