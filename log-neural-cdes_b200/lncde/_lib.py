"""ctypes binding of include/lncde_b200.h (the tested host binding; the jax.ffi
binding a JAX installation would use is shown in INTEGRATION.md)."""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# LNCDE_LIB: load a differently configured build of the same library (dev tools such as scripts/phase_timing.py)
LIB_PATH = os.environ.get("LNCDE_LIB") or os.path.join(_PKG, "liblncde_b200.so")

_lock = threading.Lock()
_lib = None

c_fp = C.c_void_p  # device pointers travel as void*


class LncdeError(RuntimeError):
    pass


def _declare(lib):
    i, i64, u, f, vp, sz = C.c_int, C.c_int64, C.c_uint, C.c_float, C.c_void_p, C.c_size_t
    sig = {
        "lncde_version": (C.c_char_p, []),
        "lncde_strerror": (C.c_char_p, [i]),
        "lncde_hall_size": (i, [i, i]),
        "lncde_hall_basis": (i, [i, i, vp]),
        "lncde_tensor_dim": (i64, [i, i]),
        "lncde_hall_t2l_nnz": (i, [i, i]),
        "lncde_hall_t2l_csr": (i, [i, i, vp, vp, vp]),
        "lncde_logsig_num_windows": (i, [i, i]),
        "lncde_logsig_fwd": (i, [vp, vp, vp, i, i, i, i, i, u, i, vp, vp, vp]),
        "lncde_logncde_param_count": (i, [i, i, i, i]),
        "lncde_logncde_fwd_smem_bytes": (sz, [i, i, i, i]),
        "lncde_logncde_fwd": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp, sz, u]),
        "lncde_logncde_bwd_workspace_bytes": (sz, [i, i, i, i, i, i, i]),
        "lncde_logncde_bwd": (i, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp, sz, u]),
        "lncde_linear_scan_workspace_bytes": (sz, [i, i, i, i, i]),
        "lncde_linear_scan_fwd": (i, [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, sz, u]),
        "lncde_linear_scan_bwd": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp, sz, u]),
        "lncde_lie_brackets_workspace_bytes": (sz, [i, i, i]),
        "lncde_lie_brackets_fwd": (i, [vp, vp, vp, vp, i, i, i, i, vp, sz]),
        "lncde_lie_brackets_bwd": (i, [vp, vp, vp, vp, i, i, i, i, vp, vp, sz]),
        "lncde_dlncde_bwd_workspace_bytes": (sz, [i, i, i]),
        "lncde_dlncde_fwd": (i, [vp, vp, vp, vp, vp, i, i, i, i, i]),
        "lncde_dlncde_bwd": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp, sz]),
        "lncde_dlncde_fwd_mean": (i, [vp, vp, vp, vp, vp, vp, i, i, i, i, i]),
        "lncde_dlncde_bwd_mean": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp, sz]),
        "lncde_nrde_param_count": (i, [i, i, i, i]),
        "lncde_nrde_fwd": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i]),
        "lncde_nrde_bwd_workspace_bytes": (sz, [i, i, i, i, i, i]),
        "lncde_nrde_bwd": (i, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp, sz]),
        "lncde_adam_step": (i, [vp, vp, vp, vp, vp, i64, f, f, f, f, i, f]),
        "lncde_adam_step_dev": (i, [vp, vp, vp, vp, vp, i64, f, f, f, f, vp, f]),
        "lncde_affine_fwd": (i, [vp, vp, vp, vp, vp, i, i, i]),
        "lncde_affine_bwd_workspace_bytes": (sz, [i, i, i]),
        "lncde_affine_bwd": (i, [vp, vp, vp, vp, i, i, i, vp, sz]),
        "lncde_readout_ce_workspace_bytes": (sz, [i, i, i]),
        "lncde_readout_ce": (i, [vp, vp, C.c_longlong, vp, vp, vp, vp, vp, C.c_longlong, vp, i, i, i, vp, sz]),
        "lncde_lip2": (i, [vp, vp, vp, vp, i, vp, vp, vp, vp, f]),
        "lncde_time_mean_fwd": (i, [vp, vp, vp, i, i, i]),
        "lncde_time_mean_bwd": (i, [vp, vp, vp, i, i, i]),
        "lncde_probe_ffma_flops": (i64, [i, i]),
        "lncde_probe_ffma": (i, [vp, vp, i, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTS = None


def lib():
    """The loaded C-ABI library; raises LncdeError if it has not been built."""
    global _lib, EXPORTS
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise LncdeError(
                    f"{LIB_PATH} is missing – run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU fallback)")
            handle = C.CDLL(LIB_PATH)
            EXPORTS = _declare(handle)
            _lib = handle
    return _lib


# ---- per-call kernel flags (include/lncde_b200.h) --------------------------------------------------------------
FLAG_COMPAT, FLAG_SAVED_FWD, FLAG_K1_TMA, FLAG_GENERIC, FLAG_K2_MMA_SYNC, FLAG_K3_V3, FLAG_K2_ONE_CTA = 1, 2, 4, 8, 16, 32, 64
_VARIANT_BITS = {"k1_tma": FLAG_K1_TMA, "generic": FLAG_GENERIC, "k2_mma_sync": FLAG_K2_MMA_SYNC, "k3_v3": FLAG_K3_V3,
                 "k2_one_cta": FLAG_K2_ONE_CTA}
_call_flags = threading.local()


def call_flags() -> int:
    """Variant bits the host mirrors OR into the `flags` argument of every C-ABI call made from this thread (0 unless
    inside a `kernel_flags(...)` block).  The library itself keeps no such state: the word travels with each call."""
    return getattr(_call_flags, "value", 0)


class kernel_flags:
    """``with kernel_flags("generic"):`` - run the enclosed calls on the named kernel variants (tests, benchmarks):
    "generic" (baseline kernels), "k1_tma", "k2_mma_sync", "k2_one_cta", "k3_v3"; see the LNCDE_FLAG_* comments in the header."""

    def __init__(self, *names):
        self.bits = 0
        for n in names:
            self.bits |= _VARIANT_BITS[n]

    def __enter__(self):
        self.prev = call_flags()
        _call_flags.value = self.prev | self.bits
        return self

    def __exit__(self, *exc):
        _call_flags.value = self.prev
        return False


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().lncde_strerror(int(status)).decode()
        raise LncdeError(f"{what or 'lncde call'} failed with status {status}: {msg}")


def require_cuda(*tensors):
    import torch

    if not torch.cuda.is_available():
        raise LncdeError("lncde needs a CUDA device (sm_100a); there is no CPU fallback")
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise LncdeError("lncde operators take CUDA tensors; got a CPU tensor (no CPU fallback)")
        if not t.is_contiguous():
            raise LncdeError("lncde operators take contiguous tensors")


def require_cuda_device(*tensors):
    """Like require_cuda without the contiguity requirement (inputs of host-side torch ops, e.g. x0 = data[:, 0, :])."""
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise LncdeError("lncde operators take CUDA tensors; got a CPU tensor (no CPU fallback)")


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())
