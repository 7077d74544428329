"""Dev tool: cycles per dependent phase of a Heun stage in the specialised Log-NCDE kernels.

Builds a -DLNCDE_PHASE_TIMING copy of the library (every warp of CTA 0 records its arrival clock at every block
barrier of stages 200..203, keyed by source line), runs the EigenWorms train step with it and prints, for the
default and the `k3_fused` kernels, the time between consecutive barriers (max over warps of the arrival clocks)
next to the source comment of the phase.  Numbers include the clock reads (a few cycles per barrier).

    python scripts/phase_timing.py            (needs a GPU)
"""
import ctypes
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "log-neural-cdes_b200"))
from lncde import _build  # noqa: E402

LIB = os.path.join(R, "log-neural-cdes_b200", "liblncde_b200_timing.so")
_build.build_library(extra_flags=["-DLNCDE_PHASE_TIMING"], lib_path=LIB)
os.environ["LNCDE_LIB"] = LIB

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lncde import _lib  # noqa: E402
from variant_check import _eigenworms_step  # noqa: E402  (scripts/ is on sys.path[0] when run as a script)

STAGES, LINES, WARPS = 4, 1280, 16
SRC = open(os.path.join(R, "log-neural-cdes_b200", "csrc", "logncde_fast.cuh")).read().splitlines()


def label(line):
    """Nearest preceding '// P..' / '// Q..' / '// F..' phase comment above the barrier at `line` (1-based)."""
    for k in range(line - 1, max(line - 40, 0), -1):
        t = SRC[k - 1].strip()
        if t.startswith("//") and len(t) > 3 and t[3:4] in "PQF":
            return t[2:70].strip()
    return ""


def read(reset=True):
    lib = _lib.lib()
    buf = np.zeros(STAGES * LINES * WARPS, np.int64)
    st = lib.lncde_debug_phase_clocks(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.size), 1 if reset else 0)
    assert st == 0, st
    return buf.reshape(STAGES, LINES, WARPS)


def report(title, clocks):
    print(f"== {title}")
    for s in range(1, STAGES):  # slot 0 may hold a partial stage
        marks = [(int(clocks[s, l].max()), l) for l in range(LINES) if clocks[s, l].max() > 0]
        marks.sort()
        if len(marks) < 2:
            continue
        total = marks[-1][0] - marks[0][0]
        print(f"  stage slot {s}: {len(marks)} barriers, {total} cycles from first to last barrier")
        if s == 1:
            for (c0, l0), (c1, l1) in zip(marks[:-1], marks[1:]):
                print(f"    line {l0:4d} -> {l1:4d}: {c1 - c0:6d} cyc   {label(l1)}")


def main():
    torch.cuda.set_device(0)
    for name, v in (("default", 0), ("k3_fused", 1)):
        _lib.set_tuning("k3_fused", v)
        model, step = _eigenworms_step(32, 400)
        read(reset=True)
        loss, grads = step()   # forward (records), then the saved-mode adjoint (records into the same table)
        torch.cuda.synchronize()
        report(f"{name}: forward + adjoint kernels share the table (lines of stage_forward < stage_reverse)", read())


if __name__ == "__main__":
    main()
