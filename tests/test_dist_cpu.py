"""Host logic of the multi-GPU path on CPU (gloo, world_size 2): row sharding, the single flat
all-reduce of actor gradients, and the per-iteration CEM candidate gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from big_dreamer_b200 import dist as D_


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = D_.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shard_range_partitions_rows():
    for n in (0, 1, 7, 2500, 2 ** 14 + 3):
        for world in (1, 2, 4, 8):
            rs = [D_.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1


def _grads(rank, world):
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
    for i, p in enumerate(lin.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    list(lin.parameters())[1].grad = None            # a parameter without grad is skipped
    D_.allreduce_grads(lin.parameters())
    return [None if p.grad is None else p.grad.clone() for p in lin.parameters()]


def test_allreduce_grads_sums_over_ranks():
    out = _run(_grads)
    for r in range(2):
        for i, g in enumerate(out[r]):
            if i == 1:
                assert g is None
            else:
                assert torch.all(g == 3.0 * (i + 1))     # (1 + 2) * (i+1)


def _gather(rank, world):
    B, H, A, C = 3, 4, 2, 11
    g = torch.Generator().manual_seed(1)
    returns = torch.randn(B, C, generator=g)
    actions = torch.randn(H, B, C, A, generator=g)
    ranges = [D_.shard_range(C, r, world) for r in range(world)]
    c0, c1 = ranges[rank]
    gr, ga = D_.gather_candidates(returns[:, c0:c1].contiguous(), actions[:, :, c0:c1].contiguous(),
                                  [e - b for b, e in ranges])
    return bool(torch.equal(gr, returns) and torch.equal(ga, actions))


def test_gather_candidates_restores_global_order():
    assert _run(_gather) == [True, True]


def _sharded_loss(rank, world):
    """Row-sharded mean loss: summing per-rank gradients of (local sum / global count) equals the
    single-process gradient -- the scaling rule the Dreamer path uses."""
    torch.manual_seed(3)
    w = torch.nn.Parameter(torch.randn(6))
    x = torch.randn(10, 6)
    full = (x @ w).mean()
    gfull, = torch.autograd.grad(full, w)
    b, e = D_.shard_range(10, rank, world)
    w.grad = None
    ((x[b:e] @ w).sum() / 10).backward()
    D_.allreduce_grads([w])
    return bool(torch.allclose(w.grad, gfull, atol=1e-6))


def test_sharded_mean_loss_matches_single_process():
    assert _run(_sharded_loss) == [True, True]
