"""Dev tool: where a Heun stage of the six-barrier Log-NCDE kernels (logncde_v3.cuh) spends its cycles.

Builds a -DLNCDE_PHASE_TIMING copy of the library: lane 0 of every warp of CTA 0 records clock64() at the top of
stage 200 and before / after each of its block barriers.  Printed per trace point: earliest and latest warp, relative
to the top of the stage - "latest arrival before barrier k" minus "latest departure after barrier k - 1" is the critical
path of phase k, "departure - last arrival" the cost of the barrier itself.

    python scripts/phase_timing.py            (needs a GPU)
"""
import ctypes
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(R, "log-neural-cdes_b200", "liblncde_b200_timing.so")
os.environ["LNCDE_LIB"] = LIB  # before lncde is imported: lncde._lib reads it at import time
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "log-neural-cdes_b200"))
from lncde import _build  # noqa: E402

_build.build_library(extra_flags=["-DLNCDE_PHASE_TIMING"], lib_path=LIB)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lncde import _lib  # noqa: E402
from variant_check import _eigenworms_step  # noqa: E402  (scripts/ is on sys.path[0] when run as a script)

POINTS, WARPS = 32, 16
NAMES = {0: ["A   Heun update + z1 partials", "B1  z1 / a1 / s1, z2 partials", "B2  z2 / a2 / s2, last layer + tanh",
             "C   q_i partials, Lambda-mixed", "D   tangent warps u1 .. p2", "E   last-layer tangents, field rows"],
         1: ["R1  field cotangents, W3^T", "R2  tangent warps", "R3  W1^T + mix", "R4  z3-bar, W3^T",
             "R5  z2-bar, W2^T partials", "R6  z1-bar, W1^T partials"]}


def read(reset):
    buf = np.zeros(2 * POINTS * WARPS, np.int64)
    st = _lib.lib().lncde_debug_v3_trace(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.size), 1 if reset else 0)
    assert st == 0, st
    return buf.reshape(2, POINTS, WARPS)


def report(kern, tr):
    top = tr[12][tr[12] > 0]
    if top.size == 0:
        print("  (no trace recorded)")
        return
    t0 = int(top.min())
    print(f"  top of stage: warps spread over {int(top.max()) - t0} cycles")
    for k, what in ((13, "update of the stage input done"), (14, "cotangents formed")):
        v = tr[k][tr[k] > 0]
        if v.size:
            print(f"  point {k} ({what}): {int(v.min()) - t0} .. {int(v.max()) - t0}")
    prev_dep = int(top.max())
    for k in range(6):
        arr, dep = tr[2 * k], tr[2 * k + 1]
        if arr.max() == 0:
            continue
        a0, a1, d0, d1 = int(arr[arr > 0].min()), int(arr.max()), int(dep[dep > 0].min()), int(dep.max())
        print(f"  {NAMES[kern][k]:40s} arrive {a0 - t0:6d} .. {a1 - t0:6d}   depart {d0 - t0:6d} .. {d1 - t0:6d}   "
              f"phase {a1 - prev_dep:5d}   barrier {d1 - a1:4d}   slowest warp {int(arr.argmax()):2d}")
        prev_dep = d1
    print(f"  stage total (top -> last departure): {prev_dep - t0} cycles")


def main():
    torch.cuda.set_device(0)
    _lib.kernel_flags("k3_v3").__enter__()
    model, step = _eigenworms_step(32, 400)
    step()
    torch.cuda.synchronize()
    read(reset=True)
    step()
    torch.cuda.synchronize()
    tr = read(reset=False)
    print("== forward stage 200 (logncde_fwd_v3_kernel, training mode)")
    report(0, tr[0])
    print("== reverse stage 200 (logncde_bwd_v3_kernel)")
    report(1, tr[1])
    np.save(os.path.join(R, "gpurun_out", "r2_v3_trace.npy"), tr)


if __name__ == "__main__":
    main()
