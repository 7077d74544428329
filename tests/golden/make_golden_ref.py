"""Generate tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES from /root/reference under the
library stand-ins of tests/golden/refshim (jax / equinox / diffrax / signax / roughpy are not installable offline;
see refshim/README.md for what this pins and what it cannot).

    python tests/golden/make_golden_ref.py          (needs /root/reference; fixtures are committed)

Files executed unmodified: data_dir/generate_paths.py (calc_paths), data_dir/datasets.py (batch_calc_paths),
data_dir/hall_set.py, models/generate_model.py (create_model), models/LogNeuralCDEs.py, models/NeuralCDEs.py
(VectorField, NeuralRDE), models/LinearNeuralCDEs.py, train.py (calc_output, classification_loss, regression_loss).
All arithmetic is float64 except what the reference keeps in float32 by construction (time grids).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LNCDE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


# model families / preprocessing outside the hot path: imported by generate_model.py and
# datasets.py at module level, never called here
_out = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("outside the hot path"))  # noqa: E731
_stub("models.LRU", LRU=_out)
_stub("models.S5", S5=_out)
_stub("models.RNN", GRUCell=_out, LinearCell=_out, LSTMCell=_out, MLPCell=_out, RNN=_out)
# NCDE interpolation coefficients (outside the hot path): dataset_generator only needs something tuple-shaped
_stub("data_dir.generate_coeffs",
      calc_coeffs=lambda data, include_time, T: (torch.zeros(data.shape[0], data.shape[1] - 1, data.shape[2]),))

import diffrax  # noqa: E402  (refshim)
import train as ref_train  # noqa: E402
from data_dir.datasets import batch_calc_paths as ref_batch_calc_paths  # noqa: E402
from data_dir.generate_paths import calc_paths as ref_calc_paths  # noqa: E402
from models.generate_model import create_model as ref_create_model  # noqa: E402


def T(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def make_ts(L):  # datasets.py:143-156 with T = 1: fp32 (T / L) * arange(L)
    return (np.float32(1.0 / L) * np.arange(L, dtype=np.float32)).astype(np.float32)


def make_intervals(ts, s):  # datasets.py:161-163
    idx = np.arange(0, len(ts), s)
    return np.concatenate([ts[idx], np.array([1.0], np.float32)]).astype(np.float32)


def paths_cases():
    rng = np.random.default_rng(20261020)
    out = {}
    cases = [(3, 30, 3, 4, 2), (5, 25, 2, 7, 2), (4, 17, 3, 4, 1), (2, 12, 4, 12, 2), (2, 13, 3, 20, 2), (3, 11, 2, 1, 2),
             (3, 8, 2, 9, 2), (2, 47, 7, 12, 2), (2, 36, 6, 12, 2), (3, 21, 1, 5, 2)]
    for i, (B, L, C, s, d) in enumerate(cases):
        x = np.cumsum(rng.normal(size=(B, L, C)), axis=1)
        out[f"x{i}"] = x
        out[f"cfg{i}"] = np.array([B, L, C, s, d])
        out[f"logsig{i}"] = ref_calc_paths(T(x), s, d).numpy()
    out["n"] = np.array(len(cases))
    # datasets.py:43-71: chunks of 128 samples (the remainder-window quirk restarts at every chunk)
    x = np.cumsum(np.round(rng.normal(size=(131, 10, 2))), axis=1)
    out["xb"] = x
    out["cfgb"] = np.array([4, 2])
    out["logsigb"] = ref_batch_calc_paths(T(x), 4, 2).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_paths.npz"), **out)
    print("ref_paths.npz:", len(cases) + 1, "cases")


def _flat_mlp(model):
    return np.concatenate([np.concatenate([l.weight.detach().numpy().ravel(), l.bias.detach().numpy().ravel()])
                           for l in model.vf.mlp.layers])


def logncde_cases():
    rng = np.random.default_rng(7)
    out = {}
    #        B, L,  C, s, depth, H, width, n_hidden, dt0, scale, lambd, classification, output_step, label_dim
    cases = [(3, 40, 3, 4, 2, 16, 8, 2, 0.05, 1.0, 0.01, True, 1, 3),
             (2, 40, 3, 4, 1, 16, 8, 2, 0.05, 1.0, 0.0, True, 1, 2),
             (2, 64, 3, 4, 2, 16, 8, 2, 0.04, 1.0, 0.01, False, 8, 1),
             (2, 36, 5, 3, 2, 24, 12, 3, 0.031, 2.0, 0.001, True, 1, 4),
             (2, 96, 7, 12, 2, 128, 32, 2, 0.11, 3.0, 0.001, True, 1, 5),
             (2, 50, 4, 5, 1, 12, 8, 1, 0.013, 1.0, 0.0, False, 7, 1)]
    for i, (B, L, C, s, depth, H, width, nh, dt0, scale, lambd, cls, ostep, ld) in enumerate(cases):
        x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C)), axis=1)
        ts = make_ts(L)
        iv = make_intervals(ts, s)
        logsig = ref_calc_paths(T(x), s, depth)
        model, state = ref_create_model("log_ncde", C, logsig.shape[-1], depth, T(iv, torch.float32), ld, H,
                                        vf_depth=nh, vf_width=width, classification=cls, output_step=ostep,
                                        solver=diffrax.Heun(), stepsize_controller=diffrax.ConstantStepSize(), dt0=dt0,
                                        scale=scale, lambd=lambd, key=100 + i)
        assert state is None
        leaves = [t for l in model.vf.mlp.layers for t in (l.weight, l.bias)] + [
            model.linear1.weight, model.linear1.bias, model.linear2.weight, model.linear2.bias]
        for t in leaves:
            t.requires_grad_(True)
        X = (T(np.tile(ts, (B, 1)), torch.float32), logsig, T(x[:, 0, :]))
        pred, _ = ref_train.calc_output(model, X, None, None, model.stateful, model.nondeterministic)
        pred = pred.detach()
        if cls:
            y = np.eye(ld)[rng.integers(0, ld, size=B)]
            (loss, _), _ = ref_train.classification_loss(model, None, X, T(y), None, None)
        else:
            y = rng.normal(size=(B, pred.shape[1]))
            (loss, _), _ = ref_train.regression_loss(model, None, X, T(y), None, None)
        # gradient of the reference's own loss function (reverse mode through its code, incl. its jax.jvp calls)
        grads = torch.autograd.grad(loss, leaves)
        out[f"grad{i}"] = np.concatenate([g.numpy().ravel() for g in grads])
        loss = loss.detach()
        for t in leaves:
            t.requires_grad_(False)
        out[f"cfg{i}"] = np.array([B, L, C, s, depth, H, width, nh, int(cls), ostep, ld])
        out[f"hyper{i}"] = np.array([dt0, scale, lambd])
        out[f"x{i}"], out[f"logsig{i}"], out[f"y{i}"] = x, logsig.numpy(), y
        out[f"mlp{i}"] = _flat_mlp(model)
        out[f"lin1_{i}"] = np.concatenate([model.linear1.weight.numpy().ravel(), model.linear1.bias.numpy()])
        out[f"lin2_{i}"] = np.concatenate([model.linear2.weight.numpy().ravel(), model.linear2.bias.numpy()])
        out[f"pairs{i}"] = np.zeros((0, 2), np.int64) if model.pairs is None else np.asarray(model.pairs)
        out[f"pred{i}"], out[f"loss{i}"] = pred.numpy(), np.array(float(loss))
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "ref_logncde.npz"), **out)
    print("ref_logncde.npz:", len(cases), "cases")


def nrde_cases():
    """create_model("nrde") -> NeuralRDE (models/NeuralCDEs.py:168-266) run unmodified: predictions, the reference's
    own losses and their gradients w.r.t. every parameter leaf."""
    rng = np.random.default_rng(17)
    out = {}
    #        B, L,  C, s, depth, H, width, n_vf, dt0, scale, classification, output_step, label_dim
    cases = [(3, 40, 3, 4, 2, 16, 8, 2, 0.05, 1.0, True, 1, 3),
             (2, 40, 3, 4, 1, 12, 8, 1, 0.05, 1.0, True, 1, 2),
             (2, 64, 3, 4, 2, 16, 8, 3, 0.04, 2.0, False, 8, 1),
             (2, 48, 6, 4, 2, 64, 32, 3, 0.07, 1.0, True, 1, 5),
             (1, 30, 7, 6, 2, 128, 16, 4, 0.15, 1.5, True, 1, 5)]
    for i, (B, L, C, s, depth, H, width, nvf, dt0, scale, cls, ostep, ld) in enumerate(cases):
        x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C)), axis=1)
        ts = make_ts(L)
        iv = make_intervals(ts, s)
        logsig = ref_calc_paths(T(x), s, depth)
        model, state = ref_create_model("nrde", C, logsig.shape[-1], depth, T(iv, torch.float32), ld, H,
                                        vf_depth=nvf, vf_width=width, classification=cls, output_step=ostep,
                                        solver=diffrax.Heun(), stepsize_controller=diffrax.ConstantStepSize(), dt0=dt0,
                                        scale=scale, key=300 + i)
        assert state is None and len(model.vf.mlp.layers) == nvf and model.logsig_dim == logsig.shape[-1] - 1
        lin = list(model.vf.mlp.layers) + [model.mlp_linear, model.linear1, model.linear2]
        leaves = [t for l in lin for t in (l.weight, l.bias)]
        for t in leaves:
            t.data = t.data.float().double()  # fp32-representable parameters: the fixture stores them losslessly as fp32
            t.requires_grad_(True)
        X = (T(np.tile(ts, (B, 1)), torch.float32), logsig, T(x[:, 0, :]))
        pred, _ = ref_train.calc_output(model, X, None, None, model.stateful, model.nondeterministic)
        pred = pred.detach()
        if cls:
            y = np.eye(ld)[rng.integers(0, ld, size=B)]
            (loss, _), _ = ref_train.classification_loss(model, None, X, T(y), None, None)
        else:
            y = rng.normal(size=(B, pred.shape[1]))
            (loss, _), _ = ref_train.regression_loss(model, None, X, T(y), None, None)
        grads = torch.autograd.grad(loss, leaves)
        out[f"grad{i}"] = np.concatenate([g.numpy().ravel() for g in grads]).astype(np.float32)
        loss = loss.detach()
        for t in leaves:
            t.requires_grad_(False)
        out[f"cfg{i}"] = np.array([B, L, C, s, depth, H, width, nvf, int(cls), ostep, ld])
        out[f"hyper{i}"] = np.array([dt0, scale])
        out[f"x{i}"], out[f"logsig{i}"], out[f"y{i}"] = x, logsig.numpy(), y
        # flat parameters in the order [vf layers | mlp_linear | linear1 | linear2], each W then b
        out[f"params{i}"] = np.concatenate([t.detach().numpy().ravel() for t in leaves]).astype(np.float32)
        out[f"pred{i}"], out[f"loss{i}"] = pred.numpy(), np.array(float(loss))
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "ref_nrde.npz"), **out)
    print("ref_nrde.npz:", len(cases), "cases")


def linear_cases():
    rng = np.random.default_rng(11)
    out = {}
    #        name, B, L, C, s, depth, H, bs, P, lambd, classification, label_dim
    cases = [("diagonal_linear_ncde", 3, 40, 3, 4, 1, 8, 1, 1, 0.1, True, 3),
             ("diagonal_linear_ncde", 2, 45, 3, 4, 2, 8, 1, 4, 0.1, True, 2),     # depth-2 rows, level 1 used (quirk B7)
             ("bd_linear_ncde", 2, 61, 3, 4, 2, 8, 4, 5, 0.05, True, 3),          # 3 chunks of 5 + remainder
             ("bd_linear_ncde", 2, 41, 2, 4, 2, 6, 2, 128, 0.0, True, 2),         # T - 1 < parallel_steps
             ("dense_linear_ncde", 2, 40, 3, 4, 2, 8, 8, 1, 0.1, False, 2),
             ("bd_linear_ncde", 1, 40, 3, 4, 3, 8, 4, 3, 0.1, True, 3)]           # depth 3 (basis order: Hall set)
    from data_dir.hall_set import HallSet

    for i, (name, B, L, C, s, depth, H, bs, P, lambd, cls, ld) in enumerate(cases):
        x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C)), axis=1)
        ts = make_ts(L)
        if depth <= 2:
            logsig = ref_calc_paths(T(x), s, depth)
        else:  # the reference's calc_paths stops at depth 2: any Hall-ordered rows serve as INPUT of the LNCDE forward
            D = len(HallSet(C, depth).data)
            logsig = T(rng.normal(scale=0.1, size=(B, (L + 1 + s - 1) // s, D)))
            logsig[..., 0] = 0
        model, state = ref_create_model(name, C, logsig.shape[-1], depth, None, ld, H, block_size=bs, classification=cls,
                                        lambd=lambd, parallel_steps=P, key=200 + i)
        assert state is None
        leaves = [model.vf_A, model.init_layer.weight, model.init_layer.bias, model.out_layer.weight, model.out_layer.bias]
        for t in leaves:
            t.requires_grad_(True)
        X = (T(np.tile(ts, (B, 1)), torch.float32), logsig, T(x[:, 0, :]))
        pred, _ = ref_train.calc_output(model, X, None, None, model.stateful, model.nondeterministic)
        pred = pred.detach()
        if cls:
            y = np.eye(ld)[rng.integers(0, ld, size=B)]
            (loss, _), _ = ref_train.classification_loss(model, None, X, T(y), None, None)
        else:
            y = rng.normal(size=(B, pred.shape[1]))
            (loss, _), _ = ref_train.regression_loss(model, None, X, T(y), None, None)
        grads = torch.autograd.grad(loss, leaves)  # gradient of the reference's own loss function
        out[f"grad{i}"] = np.concatenate([g.numpy().ravel() for g in grads])
        loss = loss.detach()
        for t in leaves:
            t.requires_grad_(False)
        out[f"name{i}"] = np.array(name)
        out[f"cfg{i}"] = np.array([B, L, C, s, depth, H, bs, P, int(cls), ld])
        out[f"lambd{i}"] = np.array(lambd)
        out[f"x{i}"], out[f"logsig{i}"], out[f"y{i}"] = x, logsig.numpy(), y
        out[f"vfA{i}"] = model.vf_A.numpy()
        out[f"init{i}"] = np.concatenate([model.init_layer.weight.numpy().ravel(), model.init_layer.bias.numpy()])
        out[f"out{i}"] = np.concatenate([model.out_layer.weight.numpy().ravel(), model.out_layer.bias.numpy()])
        out[f"basis{i}"] = np.array(str(model.basis_list))
        out[f"pred{i}"], out[f"loss{i}"] = pred.numpy(), np.array(float(loss))
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "ref_linear.npz"), **out)
    print("ref_linear.npz:", len(cases), "cases")


def slice_cases():
    """The remaining SLiCE structures of LogLinearCDE (LinearNeuralCDEs.py:124-177, :243-300, :319-329)."""
    rng = np.random.default_rng(13)
    out = {}
    #        kind, B, L, C, s, depth, H, bs, P, extra kwargs, classification, label_dim
    cases = [("wh", 2, 40, 3, 4, 2, 8, 1, 4, dict(walsh_hadamard=True), True, 3),
             ("wh", 2, 40, 3, 4, 1, 8, 1, 1, dict(walsh_hadamard=True), True, 2),      # special recurrent step :243-253
             ("dplr", 2, 41, 3, 4, 2, 8, 1, 3, dict(rank=2), True, 3),
             ("dplr", 2, 40, 3, 4, 1, 8, 1, 1, dict(rank=2), False, 1),                # special recurrent step :255-268
             ("sparse", 2, 40, 2, 4, 2, 6, 1, 4, dict(sparsity=0.3), True, 2),
             ("diagdense", 2, 45, 3, 4, 2, 12, 4, 5, dict(diagonal_dense=True), True, 3),
             ("diagdense", 2, 40, 3, 4, 1, 12, 4, 1, dict(diagonal_dense=True), False, 2)]
    for i, (kind, B, L, C, s, depth, H, bs, P, kw, cls, ld) in enumerate(cases):
        x = np.cumsum(rng.normal(scale=0.3, size=(B, L, C)), axis=1)
        ts = make_ts(L)
        logsig = ref_calc_paths(T(x), s, depth)
        model_name = {"wh": "wh", "dplr": "dplr", "sparse": "sparse", "diagdense": "diagonal_dense"}[kind] + "_linear_ncde"
        model, state = ref_create_model(model_name, C, logsig.shape[-1], depth, None, ld, H, block_size=bs,
                                        classification=cls, lambd=0.1, parallel_steps=P, key=300 + i, **kw)
        names = {"wh": ["vf_A"], "dplr": ["vf_A", "vf_A_u", "vf_A_v"], "sparse": [], "diagdense": ["vf_A_diag", "vf_A_dense"]}[kind]
        leaves = [getattr(model, n) for n in names]
        if kind == "sparse":
            dense = model.vf_A_sparse.todense()
            leaves = [dense]
        leaves += [model.init_layer.weight, model.init_layer.bias, model.out_layer.weight, model.out_layer.bias]
        for t in leaves:
            t.requires_grad_(True)
        X = (T(np.tile(ts, (B, 1)), torch.float32), logsig, T(x[:, 0, :]))
        pred, _ = ref_train.calc_output(model, X, None, None, model.stateful, model.nondeterministic)
        pred = pred.detach()
        if cls:
            y = np.eye(ld)[rng.integers(0, ld, size=B)]
            (loss, _), _ = ref_train.classification_loss(model, None, X, T(y), None, None)
        else:
            y = rng.normal(size=(B, pred.shape[1]))
            (loss, _), _ = ref_train.regression_loss(model, None, X, T(y), None, None)
        grads = torch.autograd.grad(loss, leaves)
        for t in leaves:
            t.requires_grad_(False)
        out[f"kind{i}"] = np.array(kind)
        out[f"cfg{i}"] = np.array([B, L, C, s, depth, H, bs, P, int(cls), ld])
        out[f"kw{i}"] = np.array(str(kw))
        out[f"x{i}"], out[f"logsig{i}"], out[f"y{i}"] = x, logsig.numpy(), y
        out[f"vf{i}"] = np.concatenate([t.detach().numpy().ravel() for t in leaves[: len(leaves) - 4]])
        out[f"vfgrad{i}"] = np.concatenate([g.numpy().ravel() for g in grads[: len(leaves) - 4]])
        out[f"init{i}"] = np.concatenate([model.init_layer.weight.numpy().ravel(), model.init_layer.bias.numpy()])
        out[f"out{i}"] = np.concatenate([model.out_layer.weight.numpy().ravel(), model.out_layer.bias.numpy()])
        out[f"iograd{i}"] = np.concatenate([g.numpy().ravel() for g in grads[len(leaves) - 4 :]])
        out[f"pred{i}"], out[f"loss{i}"] = pred.numpy(), np.array(float(loss.detach()))
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "ref_slice.npz"), **out)
    print("ref_slice.npz:", len(cases), "cases")


def data_cases():
    """datasets.py:dataset_generator (pre-split, so no jax.random permutation is involved) and the batching
    protocol of dataloaders.py:Dataloader."""
    import jax.random as jr
    from data_dir.dataloaders import Dataloader as RefDataloader
    from data_dir.datasets import dataset_generator as ref_dataset_generator

    rng = np.random.default_rng(17)
    out = {}
    cases = [("toy", 11, 5, 4, 18, 3, 4, 2, False, 1.0, True), ("ppg", 9, 4, 5, 23, 3, 5, 1, True, 1.0, False)]
    for i, (name, ntr, nva, nte, L, C, s, depth, include_time, Tend, onehot) in enumerate(cases):
        mk = lambda n: np.cumsum(rng.normal(scale=0.3, size=(n, L, C)), axis=1)
        data = [mk(ntr), mk(nva), mk(nte)]
        if include_time:
            for d in data:
                d[:, :, 0] = (Tend / L) * np.arange(L)
        if onehot:
            labels = [np.eye(3)[rng.integers(0, 3, size=n)] for n in (ntr, nva, nte)]
        else:
            labels = [rng.normal(size=(n, 6)) for n in (ntr, nva, nte)]
        ds = ref_dataset_generator(name, tuple(T(d) for d in data), tuple(T(l) for l in labels), s, depth, include_time,
                                   Tend, use_presplit=True, key=0)
        out[f"cfg{i}"] = np.array([ntr, nva, nte, L, C, s, depth, int(include_time), int(onehot)])
        out[f"name{i}"] = np.array(name)
        out[f"dims{i}"] = np.array([ds.data_dim, ds.logsig_dim, ds.label_dim])
        out[f"intervals{i}"] = np.asarray(ds.intervals, dtype=np.float64)
        for k, split in enumerate(("train", "val", "test")):
            out[f"x{i}_{k}"], out[f"y{i}_{k}"] = data[k], labels[k]
            batches = list(ds.path_dataloaders[split].loop_epoch(4))
            out[f"nb{i}_{k}"] = np.array([len(b[1]) for b in batches])
            out[f"ts{i}_{k}"] = torch.cat([b[0][0] for b in batches]).double().numpy()
            out[f"ls{i}_{k}"] = torch.cat([b[0][1] for b in batches]).numpy()
            out[f"x0{i}_{k}"] = torch.cat([b[0][2] for b in batches]).numpy()
            out[f"lab{i}_{k}"] = torch.cat([b[1] for b in batches]).numpy()
    out["n"] = np.array(len(cases))
    # Dataloader.loop: how many permutations are drawn while K batches are yielded (the last batch of every
    # permutation is dropped, even a full one: `while end < self.size`)
    calls = [0]
    orig = jr.permutation

    def counting(key, n):
        calls[0] += 1
        return orig(key, n)

    jr.permutation = counting
    rows = []
    for size, bs, K in ((10, 3, 7), (10, 5, 4), (12, 4, 5), (7, 6, 3)):
        dl = RefDataloader(T(rng.normal(size=(size, 4, 2))), T(rng.normal(size=(size,))))
        calls[0] = 0
        it = dl.loop(bs, key=5)
        sizes = [len(next(it)[1]) for _ in range(K)]
        rows.append([size, bs, K, calls[0], min(sizes), max(sizes)])
    jr.permutation = orig
    out["loop_protocol"] = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "ref_data.npz"), **out)
    print("ref_data.npz:", len(cases), "datasets +", len(rows), "loop protocols")


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    makers = {"data": data_cases, "paths": paths_cases, "logncde": logncde_cases, "nrde": nrde_cases,
              "linear": linear_cases, "slice": slice_cases}
    for name in (sys.argv[1:] or list(makers)):  # e.g. `make_golden_ref.py nrde` regenerates one fixture file
        makers[name]()
