"""lncde – host-side mirror of the log-neural-cdes Log-ODE hot path over the
B200 C-ABI library ``liblncde_b200.so`` (see include/lncde_b200.h).

Module names follow the reference (data_dir.generate_paths.calc_paths,
data_dir.hall_set.HallSet, models.generate_model.create_model, train.make_step).
There is no CPU fallback: every compute entry point raises if the CUDA library
or a CUDA device is missing.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
