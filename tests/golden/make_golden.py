"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

What it pins
------------
``hall_w{W}_d{D}.npz``: output of the reference's own ``data_dir/hall_set.py``
(the only file of the hot path that runs offline; it needs jax.numpy only for
``jnp.array`` so a 3-line numpy shim stands in for jax):
``data`` (the (left,right) table), ``t2l`` and ``l2t`` dense matrices.

signax / diffrax / equinox / roughpy are not installed and not vendored, so no
reference OUTPUT exists for the signature, the ODE solve or the scans; those
oracles are anchored on identities and exact integer KATs instead (see
oracle/__init__.py).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CASES = [(2, 2), (3, 2), (4, 2), (6, 2), (7, 2), (2, 3), (3, 3), (6, 3), (7, 3), (2, 4), (3, 4), (2, 5), (10, 2)]


def load_reference_hallset():
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    jnp.array = np.array
    jnp.ndarray = np.ndarray
    jax.numpy = jnp
    sys.modules.setdefault("jax", jax)
    sys.modules.setdefault("jax.numpy", jnp)
    sys.path.insert(0, REF)
    from data_dir.hall_set import HallSet  # noqa: E402

    return HallSet


def main():
    HallSet = load_reference_hallset()
    for w, d in CASES:
        hs = HallSet(w, d)
        np.savez_compressed(
            os.path.join(HERE, f"hall_w{w}_d{d}.npz"),
            data=np.asarray(hs.data, dtype=np.int32),
            t2l=np.asarray(hs.t2l_matrix(d), dtype=np.float32),
            l2t=np.asarray(hs.l2t_matrix(d), dtype=np.float32),
        )
        print(f"hall w={w} d={d}: {len(hs.data)} elements")


if __name__ == "__main__":
    main()
