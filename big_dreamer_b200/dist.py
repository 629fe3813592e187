"""Multi-GPU plumbing (one process per GPU, torch.distributed): the path shards by rows.

* Dreamer imagine + actor loss: start states are independent (src/dreamer.py:193-237), so each
  rank rolls out its own slice with replicated weights; the only exchange is ONE all-reduce of
  the actor (and critic) gradients per train step.
* PlaNet CEM: candidates are independent within an iteration; the top-K + refit is global per
  batch row (src/planner.py:74-87), so each rank evaluates C/G candidates and the ranks
  all-gather (returns, sampled actions) once per iteration.

Backend is NCCL on GPUs; the same code runs on gloo for the CPU tests of the host logic.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from RANK/WORLD_SIZE/LOCAL_RANK/MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of n rows for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_grads(params: Iterable[torch.nn.Parameter], average: bool = False) -> None:
    """Sum (or average) .grad over ranks.  The library writes all parameter gradients of one backward
    call into ONE flat fp32 buffer whose views become the ``.grad`` tensors (functions._zero_grads), so
    the actor's (and the critic's) gradients are reduced IN PLACE with one collective per buffer -- no
    concatenation, no copies back (~0.67 MB each: launch-latency-bound).  Gradients that do not share
    a storage (e.g. produced by plain autograd) fall back to one packed all-reduce."""
    if world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    groups = {}
    for g in grads:
        key = g.untyped_storage().data_ptr() if g.is_contiguous() else None
        groups.setdefault(key, []).append(g)
    loose = groups.pop(None, [])
    for gs in list(groups.values()):
        if len(gs) == 1:
            loose += gs
            continue
        lo = min(g.storage_offset() for g in gs)
        hi = max(g.storage_offset() + g.numel() for g in gs)
        flat = torch.empty(0, dtype=gs[0].dtype, device=gs[0].device).set_(gs[0].untyped_storage(), lo, (hi - lo,))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)      # padding between the views is zeros
        if average:
            flat /= world_size()
    if loose:
        flat = torch.cat([g.reshape(-1) for g in loose])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= world_size()
        off = 0
        for g in loose:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


def gather_candidates(local_returns: torch.Tensor, local_actions: torch.Tensor,
                      sizes: List[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather one CEM iteration's local results.

    local_returns (B, Cl), local_actions (H, B, Cl, A) -> returns (B, C), actions (H, B, C, A)
    with candidates in global order (rank-major), so every rank refits identically."""
    if world_size() == 1:
        return local_returns, local_actions
    world = world_size()
    cmax = max(sizes)
    B, Cl = local_returns.shape
    H, _, _, A = local_actions.shape
    pack = torch.zeros(cmax, B * (1 + H * A), device=local_returns.device, dtype=torch.float32)
    pack[:Cl, :B] = local_returns.t()
    pack[:Cl, B:] = local_actions.permute(2, 1, 0, 3).reshape(Cl, B * H * A)
    out = torch.empty(world * cmax, B * (1 + H * A), device=pack.device, dtype=torch.float32)
    dist.all_gather_into_tensor(out, pack)      # concatenated along dim 0 (works on nccl and gloo)
    out = out.view(world, cmax, B * (1 + H * A))
    rets, acts = [], []
    for r in range(world):
        blk = out[r, :sizes[r]]
        rets.append(blk[:, :B].t())
        acts.append(blk[:, B:].reshape(sizes[r], B, H, A).permute(2, 1, 0, 3))
    return torch.cat(rets, dim=1).contiguous(), torch.cat(acts, dim=2).contiguous()
