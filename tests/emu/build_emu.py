"""Build tests/emu/_build/liblncde_emu.so: the kernel sources of log-neural-cdes_b200/csrc compiled for the
HOST against the CUDA emulator (cuda_emu.h).  TEST INFRASTRUCTURE ONLY - see cuda_emu.h.

The sources are used as they are; the only textual rewrites are the two pieces of CUDA syntax a
C++ compiler cannot parse:
  * ``extern __shared__ T name[];``      -> a pointer to the CTA's emulated dynamic shared memory
  * ``kernel<<<grid, block, smem, st>>>(args);`` -> ``emu::launch_k(kernel, body, grid, block, smem, st)``
and ``ptx.cuh`` (inline PTX) is replaced by tests/emu/ptx.cuh.
"""
from __future__ import annotations

import glob
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "log-neural-cdes_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liblncde_emu.so")

_SHARED = re.compile(r"extern\s+__shared__\s+(?:__align__\(\s*\d+\s*\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\s*\[\s*\]\s*;")
_LAUNCH = re.compile(
    r"(?P<k>[A-Za-z_][A-Za-z0-9_:]*(?:<[^<>()]*>)?)\s*<<<(?P<cfg>.*?)>>>\s*\((?P<args>.*?)\)\s*;", re.S)


def transform(text: str) -> str:
    text = _SHARED.sub(lambda m: f"{m.group(1)}* {m.group(2)} = reinterpret_cast<{m.group(1)}*>(emu::dyn_smem());", text)

    def launch(m):
        k, cfg, args = m.group("k"), m.group("cfg"), m.group("args")
        return f"emu::launch_k(emu::fptr({k}), [=]() {{ {k}({args}); }}, {cfg});"

    text = _LAUNCH.sub(launch, text)
    if "<<<" in text or "extern __shared__" in text:
        raise RuntimeError("untransformed CUDA syntax left in source")
    return text


def _digest(files):
    h = hashlib.sha256()
    for f in sorted(files):
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cc")))
    srcs = [s for s in srcs if not s.endswith("xla_ffi_shim.cc")]
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")))
    own = sorted(glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(HERE, "*.cc"))
                 + glob.glob(os.path.join(HERE, "*.cuh")) + [os.path.abspath(__file__)])
    inc = sorted(glob.glob(os.path.join(REPO, "include", "*.h")))
    digest = _digest(srcs + hdrs + own + inc)
    stamp = LIB + ".stamp"
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    src_dir = os.path.join(OUT, "src")
    os.makedirs(src_dir, exist_ok=True)
    for f in glob.glob(os.path.join(src_dir, "*")):
        os.remove(f)
    units = []
    for f in srcs + hdrs:
        base = os.path.basename(f)
        if base == "ptx.cuh":
            continue  # tests/emu/ptx.cuh is found through -I instead
        with open(f) as fh:
            text = transform(fh.read())
        dst = os.path.join(src_dir, base + (".cc" if base.endswith(".cu") else ""))
        with open(dst, "w") as fh:
            fh.write(text)
        if f in srcs:
            units.append(dst)
    units.append(os.path.join(HERE, "emu_runtime.cc"))
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    # -fsanitize=alignment: every dereference of a float4 / float2 / 64-bit pointer the kernels form by reinterpret_cast
    # is checked against the type's alignment (a misaligned 128-bit access is a fault on the GPU, silent on x86);
    # trap-on-error keeps the library free of a sanitizer runtime - emu_runtime.cc turns the trap into a message
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-w", "-fsanitize=alignment", "-fsanitize-undefined-trap-on-error",
             "-DLNCDE_EMU=1", "-include", os.path.join(HERE, "cuda_emu.h"),
             "-I", HERE, "-I", src_dir, "-I", os.path.join(REPO, "include"), "-I", cuda_inc]
    procs, objs = [], []
    for u in units:
        obj = os.path.join(OUT, os.path.basename(u) + ".o")
        objs.append(obj)
        procs.append((u, subprocess.Popen(["g++", *flags, "-x", "c++", "-c", u, "-o", obj], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    for u, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"emulator build failed on {u}:\n{out[-6000:]}")
        if verbose and out:
            print(out)
    r = subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", "-o", LIB, *objs], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"emulator link failed:\n{r.stdout[-6000:]}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
