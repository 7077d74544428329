"""Dry run of the `-m gpu` test files WITHOUT a GPU (dev tool, not part of either test tier):

    python -m pytest -p gpu_on_emu_plugin -m gpu tests/test_full_size_gpu.py        (PYTHONPATH=tests/emu)

The product library is replaced by the host-emulated build of the same kernel sources (tests/emu), "cuda" device
requests are redirected to the CPU and the CUDA-only torch entry points the tests touch become no-ops.  What it
checks before a GPU box is spent: the test code itself (shapes, argument order, fixtures, device plumbing) and the
fp32 tolerances up to the error of the GPU's fast-math units.  It proves nothing about the hardware run."""
import contextlib
import ctypes as C
import os
import sys
import time

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def _is_cuda_dev(d):
    return d is not None and str(d).startswith("cuda")


def _redirect(fn):
    """fn with every "cuda" device argument replaced by the CPU."""
    def wrapped(*args, **kwargs):
        if _is_cuda_dev(kwargs.get("device")):
            kwargs["device"] = "cpu"
        args = tuple("cpu" if isinstance(a, (str, torch.device)) and _is_cuda_dev(a) else a for a in args)
        return fn(*args, **kwargs)

    return wrapped


_FACTORIES = ("empty", "zeros", "ones", "full", "eye", "randn", "rand", "randint", "arange", "linspace", "tensor",
              "as_tensor", "empty_like", "zeros_like", "ones_like", "full_like", "randn_like", "rand_like", "randperm")


class _CpuGenerator(torch.Generator):  # stays a class: torch modules use `torch.Generator | None` in annotations
    def __new__(cls, device=None):
        return super().__new__(cls, "cpu")


class _FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


def pytest_configure(config):
    import build_emu
    from lncde import _build, _lib

    import torch._decomp.decompositions  # noqa: F401  (scripts its helpers at import: before torch.* is wrapped)
    import torch._functorch.eager_transforms  # noqa: F401

    x = torch.ones(3)
    torch.func.jvp(torch.tanh, (x,), (x,))
    handle = C.CDLL(build_emu.build())
    _lib._declare(handle)
    _lib.lib = lambda: handle
    _lib.require_cuda = lambda *a: None
    _lib.require_cuda_device = lambda *a: None
    _lib.stream_ptr = lambda: C.c_void_p(0)
    _build.build_library = lambda *a, **k: "emulated"
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.Event = _FakeEvent
    torch.cuda.empty_cache = lambda: None
    torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    torch.Generator = _CpuGenerator
    for name in _FACTORIES:
        setattr(torch, name, _redirect(getattr(torch, name)))
    torch.Tensor.to = _redirect(torch.Tensor.to)
    for name in ("new_empty", "new_zeros", "new_ones", "new_full", "new_tensor"):
        setattr(torch.Tensor, name, _redirect(getattr(torch.Tensor, name)))


@pytest.fixture(scope="session", autouse=True)
def _announce():
    print("\n[gpu_on_emu_plugin] -m gpu tests running on the CPU over the emulated kernel library")
    yield
