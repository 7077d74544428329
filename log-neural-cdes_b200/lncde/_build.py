"""In-tree build of the C-ABI shared library ``liblncde_b200.so`` with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "liblncde_b200.so")
STAMP = LIB_PATH + ".stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O3",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cc")))


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(
        glob.glob(os.path.join(REPO, "include", "*.h")))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), lib_path: str | None = None) -> str:
    """Compile every source under csrc/ into one shared library (idempotent).  ``extra_flags`` / ``lib_path``
    build a differently configured copy (dev tools, e.g. -DLNCDE_PHASE_TIMING) next to the product library."""
    if extra_flags or lib_path:
        return _build_variant(list(extra_flags), lib_path or LIB_PATH.replace(".so", "_variant.so"), verbose)
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src) + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(REPO, "include"), "-I", CSRC, "-x", "cu", "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB_PATH


def _build_variant(extra_flags, lib_path, verbose=False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    build_dir = os.path.join(PKG_DIR, "build", os.path.basename(lib_path) + ".d")
    os.makedirs(build_dir, exist_ok=True)
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src) + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-I", os.path.join(REPO, "include"), "-I", CSRC, "-x", "cu", "-c", src,
               "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    r = subprocess.run([nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return lib_path


if __name__ == "__main__":
    import sys

    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
