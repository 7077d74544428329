"""Stand-in for the part of `diffrax` 0.7 the reference hot path uses.  See README.md.

diffeqsolve with Heun + ConstantStepSize, times in float32 (x64 is off in the reference): the first step ends at
t0 + dt0, every later one at t + (t - tprev) (ConstantStepSize.adapt_step_size), and a step end within 1e-6 of t1
is moved onto t1 (_clip_to_end, float32 tolerance); k1 = f(t_n, y_n), k2 = f(t_{n+1}, y_n + dt k1),
y_{n+1} = y_n + dt/2 (k1 + k2) with dt = t_{n+1} - t_n in float32; SaveAt(t1=True) keeps y(t1); SaveAt(ts=...) evaluates the solver's dense output,
which for Heun is the linear interpolant between step ends, at each requested time, then appends y(t1)."""
import numpy as np
import torch


class AbstractSolver:
    pass


class AbstractStepSizeController:
    pass


class Heun(AbstractSolver):
    pass


class ConstantStepSize(AbstractStepSizeController):
    pass


class SaveAt:
    def __init__(self, t1=False, ts=None):
        self.t1, self.ts = t1, ts


class ODETerm:
    def __init__(self, f):
        self.f = f


class Solution:
    def __init__(self, ys, ts):
        self.ys, self.ts = ys, ts


def _f32(x):
    return np.float32(x.item() if hasattr(x, "item") else x)


def diffeqsolve(term, solver, t0, t1, dt0, y0, args=None, *, saveat, stepsize_controller, max_steps=4096, **kw):
    assert isinstance(solver, Heun) and isinstance(stepsize_controller, ConstantStepSize)
    t0, t1, dt0 = _f32(t0), _f32(t1), _f32(dt0)
    f = lambda t, y: term.f(torch.tensor(t, dtype=torch.float32), y, args)
    grid, ys = [t0], [y0]
    t, y, n = t0, y0, 0
    tn = np.float32(t0 + dt0)
    while t < t1:
        assert n < max_steps, "max_steps reached"
        if tn > np.float32(t1 - np.float32(1e-6)):
            tn = t1
        dt = np.float32(tn - t)
        k1 = f(t, y)
        k2 = f(tn, y + float(dt) * k1)
        y = y + (0.5 * float(dt)) * (k1 + k2)
        t, tn = tn, np.float32(tn + np.float32(tn - t))
        n += 1
        grid.append(t)
        ys.append(y)
    out = []
    if saveat.ts is not None:
        g = np.asarray(grid, np.float32)
        for tq in np.asarray(saveat.ts, np.float32):
            i = int(np.clip(np.searchsorted(g, tq, side="left"), 1, len(g) - 1))
            w = float((tq - g[i - 1]) / (g[i] - g[i - 1]))
            out.append(ys[i - 1] + w * (ys[i] - ys[i - 1]))
    if saveat.t1:
        out.append(y)
    return Solution(torch.stack(out, 0), None)
