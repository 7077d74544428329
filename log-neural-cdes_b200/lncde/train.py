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
    """train.py:71-97: mean(-sum(y * log(pred + 1e-8))) + Lip(2) term.

    Models that bring a kernel-only step (``fused_classification_loss``: LogNeuralCDE) take it - no torch autograd, no
    eager glue between the solve kernels; ``lncde._lib.kernel_flags("generic")`` keeps the autograd formulation below
    reachable (it is what the fused step is tested against)."""
    if model.classification and hasattr(model, "fused_classification_loss") and not (_lib.call_flags() & _lib.FLAG_GENERIC):
        fused = model.fused_classification_loss(X, y)
        if fused is not None:  # None: a parameterisation without a kernel-only step (the rarer SLiCE structures)
            return fused
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


def _scaled_inputs(model_name, dataset_name, X):
    """train.py:212-216: log-signature rescaling the reference applies to two datasets."""
    if model_name == "dplr_linear_ncde":
        if dataset_name in ("Heartbeat", "MotorImagery"):
            return (X[0], X[1] / 100, X[2])
    elif model_name.endswith("linear_ncde") and dataset_name == "Heartbeat":
        return (X[0], X[1] / 10, X[2])
    return X


def evaluate(model_name, dataset_name, model, dataloader, batch_size, state=None):
    """Metric of one ordered pass (train.py:224-250): accuracy, or mean over samples of the per-sample MSE."""
    preds, labels = [], []
    with torch.no_grad():
        for X, y in dataloader.loop_epoch(batch_size):
            pred, _ = calc_output(model, _scaled_inputs(model_name, dataset_name, X), state, None, model.stateful,
                                  model.nondeterministic)
            preds.append(pred)
            labels.append(y)
    pred, y = torch.cat(preds), torch.cat(labels)
    if model.classification:
        return float((pred.argmax(dim=1) == y.argmax(dim=1)).float().mean())
    return float(((pred[:, :, 0] - y) ** 2).mean(dim=1).mean(dim=0))


def train_model(model_name, dataset_name, model, metric, filter_spec, state, dataloaders, num_steps, print_steps,
                early_stopping_steps, lr, lr_scheduler, batch_size, key, output_dir, *, world_size=1, overwrite=False,
                log=print):
    """train.py:139-349: the training loop with periodic train / validation metrics, early stopping on the
    validation metric, a test pass whenever validation improves, and the .npy curves the reference writes.
    Differences: an existing ``output_dir`` raises unless ``overwrite`` (no interactive prompt); ``lr_scheduler(lr)``
    may return a float or a step -> lr callable; ``key`` seeds the batch shuffling."""
    import os
    import shutil
    import time

    import numpy as np

    if metric == "accuracy":
        best_val, improv, no_improv, first = max, (lambda x, y: x >= y), (lambda x, y: x <= y), 0.0
    elif metric == "mse":
        best_val, improv, no_improv, first = min, (lambda x, y: x <= y), (lambda x, y: x >= y), 100.0
    else:
        raise ValueError(f"Unknown metric: {metric}")
    if os.path.isdir(output_dir):
        if not overwrite:
            raise ValueError(f"Directory {output_dir} already exists. Exiting.")
        shutil.rmtree(output_dir)
    os.makedirs(output_dir)
    sched = lr_scheduler(lr) if lr_scheduler is not None else lr
    opt = Adam(model, sched(0) if callable(sched) else sched)
    loss_fn = classification_loss if model.classification else regression_loss
    running_loss = 0.0
    all_val, all_train, val_for_best, all_time = [first], [first], [first], []
    no_val_improvement, test_metric = 0, None
    start = time.time()
    for step, (X, y) in zip(range(num_steps), dataloaders["train"].loop(batch_size, key=key)):
        if callable(sched):
            opt.lr = float(sched(step))
        X = _scaled_inputs(model_name, dataset_name, X)
        model, state, _, value = make_step(model, filter_spec, X, y, loss_fn, state, opt, None, None,
                                           world_size=world_size)
        running_loss += float(value)
        if (step + 1) % print_steps == 0:
            end = time.time()
            train_metric = evaluate(model_name, dataset_name, model, dataloaders["train"], batch_size, state)
            val_metric = evaluate(model_name, dataset_name, model, dataloaders["val"], batch_size, state)
            total_time = end - start
            log(f"Step: {step + 1}, Loss: {running_loss / print_steps}, Train metric: {train_metric}, "
                f"Validation metric: {val_metric}, Time: {total_time}")
            start = time.time()
            if step > 0:
                if no_improv(val_metric, best_val(val_for_best)):
                    no_val_improvement += 1
                    if no_val_improvement > early_stopping_steps:
                        break
                else:
                    no_val_improvement = 0
                if improv(val_metric, best_val(val_for_best)):
                    val_for_best.append(val_metric)
                    test_metric = evaluate(model_name, dataset_name, model, dataloaders["test"], batch_size, state)
                    log(f"Test metric: {test_metric}")
                running_loss = 0.0
                all_train.append(train_metric)
                all_val.append(val_metric)
                all_time.append(total_time)
                np.save(output_dir + "/steps.npy", np.arange(0, step + 1, print_steps))
                np.save(output_dir + "/all_train_metric.npy", np.array(all_train))
                np.save(output_dir + "/all_val_metric.npy", np.array(all_val))
                np.save(output_dir + "/all_time.npy", np.array(all_time))
                np.save(output_dir + "/test_metric.npy", np.array(test_metric if test_metric is not None else np.nan))
    log(f"Test metric: {test_metric}")
    return model


def experiment_output_dir(model_name, dataset_name, T, include_time, num_steps, lr, stepsize, logsig_depth, model_args,
                          seed, output_parent_dir=""):
    """Directory name of one run, exactly as train.py:386-403 composes it."""
    parent = output_parent_dir + "outputs/" + model_name + "/" + dataset_name
    out = f"T_{T:.2f}_time_{include_time}_nsteps_{num_steps}_lr_{lr}"
    if model_name in ("log_ncde", "nrde"):
        out += f"_stepsize_{stepsize:.2f}_depth_{logsig_depth}"
    for k, v in model_args.items():
        name = str(v)
        if v is not None:
            if "(" in name:
                name = name.split("(", 1)[0]
            if name == "dt0":
                out += f"_{k}_" + f"{v:.2f}"
            else:
                out += f"_{k}_" + name
            if name == "PIDController":
                out += f"_rtol_{v.rtol}_atol_{v.atol}"
    return parent + "/" + out + f"_seed_{seed}"


def create_dataset_model_and_train(seed, data_dir, use_presplit, dataset_name, output_step, metric, include_time, T,
                                   model_name, stepsize, logsig_depth, model_args, num_steps, print_steps,
                                   early_stopping_steps, lr, lr_scheduler, batch_size, output_parent_dir="", *,
                                   device="cuda", world_size=1, overwrite=False, log=print, paths_fn=None):
    """train.py:365-487 with the reference's signature, for the models on the Log-ODE input (``log_ncde``, ``nrde``,
    ``*_linear_ncde``): dataset -> model -> ``train_model`` on the path dataloaders.  No ``filter_spec`` is needed:
    ``intervals`` / ``pairs`` / ``hadamard_matrix`` are not part of the flat parameter vector of the mirrors."""
    from .data_dir.datasets import create_dataset
    from .models.generate_model import create_model

    if not (model_name in ("log_ncde", "nrde") or model_name.endswith("linear_ncde")):
        raise NotImplementedError(f"{model_name} does not take the Log-ODE input (SURVEY.md section 8)")
    output_dir = experiment_output_dir(model_name, dataset_name, T, include_time, num_steps, lr, stepsize, logsig_depth,
                                       model_args, seed, output_parent_dir)
    log(f"Creating dataset {dataset_name}")
    dataset = create_dataset(data_dir, dataset_name, stepsize=stepsize, depth=logsig_depth, include_time=include_time, T=T,
                             use_idxs=False, use_presplit=use_presplit, scale=model_name.endswith("linear_ncde"),
                             key=seed, device=device, paths_fn=paths_fn)
    log(f"Creating model {model_name}")
    model, state = create_model(model_name, dataset.data_dim, dataset.logsig_dim, logsig_depth, dataset.intervals,
                                dataset.label_dim, classification=metric == "accuracy", output_step=output_step,
                                **model_args, key=seed + 1, device=device)
    return train_model(model_name, dataset_name, model, metric, None, state, dataset.path_dataloaders, num_steps,
                       print_steps, early_stopping_steps, lr, lr_scheduler, batch_size, seed + 2, output_dir,
                       world_size=world_size, overwrite=overwrite, log=log)


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
