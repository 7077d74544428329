"""Training step – mirror of train.py:52-136 (calc_output, classification_loss,
regression_loss, make_step) for the hot-path models, plus the batch-sharded
data-parallel step (one process per GPU, one NCCL all-reduce of the flat gradient).

Loss functions return ``(value, flat_grad)`` like ``eqx.filter_value_and_grad``;
gradients are taken w.r.t. every inexact leaf except ``intervals`` / ``pairs``
(train.py:448-452) – here: the model's flat parameter vector.
"""
from __future__ import annotations

import torch

from . import _lib


def calc_output(model, X, state=None, key=None, stateful=False, nondeterministic=False):
    """train.py:52-68; the hot-path models are neither stateful nor nondeterministic."""
    if stateful or nondeterministic:
        raise NotImplementedError("stateful / nondeterministic models are outside the hot path")
    return model(X), state


def _lip2(model):
    """train.py:79-93."""
    if not model.lip2:
        return 0.0
    norm = 0.0
    if hasattr(model, "vf"):
        for layer in model.vf.mlp.layers:
            norm = norm + torch.mean(torch.linalg.norm(layer.weight, dim=-1) + torch.linalg.norm(layer.bias, dim=-1))
    elif getattr(model, "vf_A", None) is not None:
        norm = norm + torch.mean(torch.linalg.norm(model.vf_A, dim=-1))
    elif getattr(model, "vf_A_sparse", None) is not None:  # train.py:89-91
        norm = norm + torch.mean(torch.linalg.norm(model.vf_A_sparse.todense(), dim=-1))
    return norm * model.lambd


def _value_and_grad(model, loss):
    (g,) = torch.autograd.grad(loss, model.params)
    return loss.detach(), g


def classification_loss(model, X, y, state=None, key=None):
    """train.py:71-97: mean(-sum(y * log(pred + 1e-8))) + Lip(2) term."""
    pred, _ = calc_output(model, X, state, key, model.stateful, model.nondeterministic)
    loss = torch.mean(-torch.sum(y * torch.log(pred + 1e-8), dim=1)) + _lip2(model)
    return _value_and_grad(model, loss)


def regression_loss(model, X, y, state=None, key=None):
    """train.py:100-127: mean(mean((pred[:,:,0] - y)^2, axis=1)) + Lip(2) term."""
    pred, _ = calc_output(model, X, state, key, model.stateful, model.nondeterministic)
    loss = torch.mean(torch.mean((pred[:, :, 0] - y) ** 2, dim=1)) + _lip2(model)
    return _value_and_grad(model, loss)


class Adam:
    """optax.adam(lr) (train.py:183) on the flat parameter vector, fused kernel K4."""

    def __init__(self, model, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = float(lr), b1, b2, eps
        self.m = torch.zeros_like(model.params)
        self.v = torch.zeros_like(model.params)
        # the step count lives on the device so that an update captured in a CUDA graph stays valid
        self.count = torch.zeros(1, dtype=torch.int32, device=model.params.device)

    def update(self, model, grads, grad_scale=1.0):
        _lib.require_cuda(model.params, grads)
        p = model.params.detach()
        st = _lib.lib().lncde_adam_step_dev(_lib.stream_ptr(), _lib.ptr(p), _lib.ptr(grads.contiguous()),
                                            _lib.ptr(self.m), _lib.ptr(self.v), p.numel(), self.lr, self.b1, self.b2,
                                            self.eps, _lib.ptr(self.count), float(grad_scale))
        _lib.check(st, "lncde_adam_step_dev")


def shard_slice(global_batch: int, rank: int, world_size: int) -> slice:
    """Contiguous batch shard of one rank (samples are independent through K1/K2/K3, so the
    batch axis is the only axis that is split; parameters are replicated)."""
    if global_batch % world_size:
        raise ValueError("global batch must be divisible by the number of ranks")
    per = global_batch // world_size
    return slice(rank * per, (rank + 1) * per)


def allreduce_sum_(flat_grads: torch.Tensor, world_size: int) -> torch.Tensor:
    """The ONE collective of a data-parallel step: in-place sum of the flat gradient over ranks
    (NCCL over NVLink on GPUs, gloo in the CPU tests).  The 1/world factor that turns the sum of
    per-shard mean-loss gradients into the global-batch gradient is applied by the consumer
    (Adam kernel ``grad_scale``)."""
    if world_size > 1:
        import torch.distributed as dist

        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def make_step(model, filter_spec, X, y, loss_fn, state, opt, opt_state, key, *, world_size=1):
    """train.py:130-136.  With world_size > 1 every rank holds a batch shard; the flat
    gradient is summed with ONE all-reduce and averaged inside the Adam kernel."""
    value, grads = loss_fn(model, X, y, state, key)
    allreduce_sum_(grads, world_size)
    opt.update(model, grads, grad_scale=1.0 / world_size)
    return model, state, opt_state, value


class GraphedStep:
    """One training step captured in a CUDA graph (streams and graphs instead of a tracing compiler):
    the ~100 small launches of a step (K1, solve / scan kernels, read-out + loss autograd, gradient GEMMs,
    [NCCL all-reduce], Adam) replay as ONE graph launch.  Inputs are copied into static buffers."""

    def __init__(self, step_fn, example_inputs, warmup=3):
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = step_fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
