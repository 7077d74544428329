"""bench.py end to end WITHOUT a GPU: the real run_ours() / extra_measurements() / cpu_baseline() code over the
host-emulated kernel library on a shrunken EigenWorms workload (2 samples, 10 Heun steps).  Test infrastructure:
it guards the bench's own Python (JSON contract, autotune decision plumbing, roofline / breakdown legs) between
hardware runs; every number it prints is meaningless.  Run by tests/test_bench_dryrun_cpu.py in a subprocess
because it replaces torch.cuda entry points process-wide."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, REPO, os.path.join(REPO, "log-neural-cdes_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

import gpu_on_emu_plugin

gpu_on_emu_plugin.pytest_configure(None)
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.cuda.get_device_name = lambda *a, **k: "emulated"

import bench
from lncde import train as T


class FakeGraphedStep:
    """Same surface as lncde.train.GraphedStep (static input buffers, replay returns the static output)."""

    def __init__(self, step_fn, example_inputs, warmup=3):
        self.static_in = [t.clone() for t in example_inputs]
        self.step_fn = step_fn

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
        return self.step_fn(*self.static_in)


T.GraphedStep = FakeGraphedStep
bench.WORKLOADS["eigenworms_logncde"].update(B=2, L=120, dt0=0.1)
bench.N_ROT = 2
bench.K1_ROOFLINE_BATCH = 4
bench.FFMA_PROBE_SHAPE = (2, 3)
sys.argv = ["bench.py", "--steps", "2", "--warmup", "3"] + sys.argv[1:]
bench.main()
