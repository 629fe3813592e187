"""Multi-rank parity check on real GPUs (NCCL), launched by tests/test_gpu_round2.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 \
        --master-port 29611 scripts/dist_check.py

(1) PlaNet CEM with the candidates sharded over the ranks (one all-gather per iteration) gives the SAME
    plan as the single-rank planner on the same global noise: elite sets identical in every iteration,
    final action bit-identical in fp32 check mode (fp16: within the 1e-2 contract).
(2) Dreamer actor loss with the start states sharded over the ranks: after dist.allreduce_grads the actor
    gradients equal the single-GPU gradients of the whole batch (fp32: <= 1e-5, summation order only).
Prints DIST_CHECK_OK on rank 0 and writes a log line per check."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd                      # noqa: E402
from big_dreamer_b200 import dist as D_            # noqa: E402
from oracle import rssm_oracle as orc              # noqa: E402  (weights / latents recipe)
from tests import parity_utils as pu               # noqa: E402


def main():
    rank, world, local = D_.init_from_env()
    assert world >= 2, "run under torchrun with >= 2 ranks"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ok = True
    # ---------------- (1) CEM
    d = dict(Be=200, Hi=200, S=30, A=1, E=8, B=2, C=1000, K=100, H=12, iters=4, act="ELU")
    trans, actor, reward, value = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    mods = pu.build_gpu_models(d, trans, reward_sd=reward, device=dev)
    pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    s0, b0 = orc.make_latents(0, d["B"], d["Be"], d["S"])
    gen = torch.Generator(device=dev).manual_seed(77)
    noise = pl.draw_noise(d["B"], dev, generator=gen)          # identical on every rank
    for prec in ("fp32", "fp16"):
        bd.set_precision(prec)
        pl.shard_candidates = True
        out_sh = pl(b0.to(dev), s0.to(dev), noise=noise, trace=True)
        tr_sh = {k: v.clone() for k, v in pl.last_trace.items()}
        pl.shard_candidates = False
        out_1 = pl(b0.to(dev), s0.to(dev), noise=noise, trace=True)
        same_sets = torch.equal(torch.sort(tr_sh["topk"], dim=2)[0], torch.sort(pl.last_trace["topk"], dim=2)[0])
        err = float((out_sh - out_1).abs().max())
        if rank == 0:
            print(f"cem {prec}: sharded over {world} ranks vs single rank: elite sets identical={same_sets}, "
                  f"final action max abs diff {err:.3e}")
        # fp32 check mode: bit-identical.  fp16: the per-row arithmetic is the same but the cluster shape
        # (columns per CTA) follows the local row count; held to the fp16 contract (final action <= 1e-2)
        ok &= (same_sets and err == 0.0) if prec == "fp32" else err <= 1e-2
    # ---------------- (2) row-sharded actor gradients
    d = dict(Be=200, Hi=200, S=30, A=1, E=8, N=600, H=15, act="ELU")
    T = d["H"] - 1
    mods = pu.build_gpu_models(d, trans, actor, reward, value, device=dev)
    pu.freeze(mods.transition, mods.reward, mods.critic)
    agent = pu.agent_ns(mods, d["H"])
    s0, b0 = orc.make_latents(3, d["N"], d["Be"], d["S"])
    ea, ee, es = (t.to(dev) for t in orc.make_imagine_noise(3, T, d["N"], d["S"], d["A"]))
    s0, b0 = s0.to(dev), b0.to(dev)
    params = list(mods.actor.parameters())

    def grads_of(lo, hi):
        for p in params:
            p.grad = None
        nz = dict(eps_a=ea[:, lo:hi].contiguous(), eps_e=ee[:, :, lo:hi].contiguous(), eps_s=es[:, lo:hi].contiguous())
        out = bd.imagine_and_returns(agent, s0[None, lo:hi], b0[None, lo:hi], mods.reward, mods.critic, 0.995, 0.95, nz)
        loss = -(out[6] + 1e-5 * out[3].unsqueeze(-1)).sum() / (T * d["N"])     # slice of the global mean
        loss.backward()
    for prec, tol in (("fp32", 1e-5), ("fp16", 2e-3)):
        bd.set_precision(prec)
        grads_of(0, d["N"])
        full = [p.grad.clone() for p in params]
        lo, hi = D_.shard_range(d["N"], rank, world)
        grads_of(lo, hi)
        D_.allreduce_grads(params)
        err = max(pu.relerr(p.grad, f) for p, f in zip(params, full))
        one_buffer = len({p.grad.untyped_storage().data_ptr() for p in params}) == 1
        if rank == 0:
            print(f"actor grads {prec}: rows sharded over {world} ranks + all-reduce vs single GPU: max rel err "
                  f"{err:.3e}; gradients in one flat buffer (single collective): {one_buffer}")
        ok &= err < tol
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    if rank == 0 and flag.item() == 1.0:
        print("DIST_CHECK_OK")
    torch.distributed.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
