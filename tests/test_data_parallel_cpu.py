"""N > 1 host logic on CPU: world_size-2 gloo ranks, batch sharding + the single gradient
all-reduce reproduce the full-batch gradient of the oracle model."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    for p in (REPO, os.path.join(REPO, "log-neural-cdes_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lncde import train as T
    from oracle import linear_ncde as OL
    from oracle import paths

    torch.manual_seed(0)
    B, L, C, s, depth, H, bs = 4, 40, 3, 4, 2, 8, 4
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.normal(scale=0.1, size=(B, L, C)), axis=1).astype(np.float32)
    ls = torch.from_numpy(paths.calc_paths(x, s, depth, np.float64))
    x0 = torch.from_numpy(x[:, 0]).double()
    y = torch.nn.functional.one_hot(torch.tensor([0, 1, 1, 0]), 2).double()
    # rank 0 owns the parameters, broadcast like bench.py does
    g = torch.Generator().manual_seed(1)
    params = [torch.randn(C, (H // bs) * bs * bs, generator=g, dtype=torch.float64) * 0.2,
              torch.randn(H, C, generator=g, dtype=torch.float64), torch.randn(H, generator=g, dtype=torch.float64),
              torch.randn(2, H, generator=g, dtype=torch.float64), torch.randn(2, generator=g, dtype=torch.float64)]
    if rank != 0:
        params = [torch.zeros_like(p) for p in params]
    for p in params:
        dist.broadcast(p, src=0)

    def grad_on(sl):
        leaves = [p.clone().requires_grad_(True) for p in params]
        pred = OL.predict(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], ls[sl], x0[sl], C, bs, depth, True)
        loss = torch.mean(-torch.sum(y[sl] * torch.log(pred + 1e-8), dim=1))
        gr = torch.autograd.grad(loss, leaves)
        return torch.cat([t.reshape(-1) for t in gr])

    sl = T.shard_slice(B, rank, world)
    flat = grad_on(sl)
    T.allreduce_sum_(flat, world)
    flat = flat / world
    full = grad_on(slice(0, B))
    ok = torch.allclose(flat, full, rtol=1e-10, atol=1e-12)
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
        fh.write(f"{int(ok)} {sl.start} {sl.stop}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [open(tmp_path / f"rank{r}.txt").read().split() for r in range(world)]
    assert got[0] == ["1", "0", "2"] and got[1] == ["1", "2", "4"]


def test_shard_slice_errors():
    import pytest

    from lncde import train as T

    assert T.shard_slice(32, 3, 8) == slice(12, 16)
    with pytest.raises(ValueError):
        T.shard_slice(30, 0, 8)


def _prep_worker(rank, world, port, out_dir):
    import sys

    for p in (REPO, os.path.join(REPO, "log-neural-cdes_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lncde.data_dir.generate_paths import preprocess_shard
    from oracle import paths

    n, L, C, s = 300, 10, 2, 3  # 3 groups of 128 / 128 / 44 samples; stepsize 3 leaves a remainder window (quirk B1)
    x = np.cumsum(np.round(np.random.default_rng(4).normal(size=(n, L, C))), axis=1)
    full = paths.batch_calc_paths(x, s, 2, np.float64)
    sl = preprocess_shard(n, rank, world)
    mine = paths.batch_calc_paths(x[sl], s, 2, np.float64)
    gathered = [None] * world
    dist.all_gather_object(gathered, (sl.start, sl.stop, mine))  # test only: the product path needs no collective
    cat = np.concatenate([g[2] for g in gathered])
    naive = paths.batch_calc_paths(x[rank * n // world : (rank + 1) * n // world], s, 2, np.float64)
    naive_ok = np.array_equal(naive, full[rank * n // world : (rank + 1) * n // world])
    with open(os.path.join(out_dir, f"prep{rank}.txt"), "w") as fh:
        fh.write(f"{int(np.array_equal(cat, full))} {sl.start} {sl.stop} {int(naive_ok)}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_preprocessing_shards_need_no_halo(tmp_path):
    """Whole-group shards reproduce the single-process log-signatures exactly; an even split that cuts a 128-sample
    group does not (rank 1 would start at sample 150 and lose its predecessor's basepoint)."""
    world = 2
    mp.spawn(_prep_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [open(tmp_path / f"prep{r}.txt").read().split() for r in range(world)]
    assert got[0] == ["1", "0", "128", "1"]
    assert got[1] == ["1", "128", "300", "0"]


def test_preprocess_shard_covers_everything():
    from lncde.data_dir.generate_paths import preprocess_shard

    for n in (1, 127, 128, 129, 1000, 4096):
        for world in (1, 2, 4, 8):
            sl = [preprocess_shard(n, r, world) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(sl[:-1], sl[1:]))
            assert all(s.start % 128 == 0 for s in sl)
