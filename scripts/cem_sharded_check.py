"""torchrun -n 2: sharded CEM must give the same elites/action as the single-rank plan (same noise)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from big_dreamer_b200 import dist as D_
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
rank, world, local = D_.init_from_env()
torch.cuda.set_device(local)
d = dict(Be=200, Hi=200, S=30, A=2, E=8, B=2, C=1000, K=100, H=12, iters=4, act="ELU")
trans, _, reward, _ = orc.make_models(1, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
g = torch.Generator().manual_seed(8)
s0, b0 = orc.make_latents(1, d["B"], d["Be"], d["S"])
ea = torch.randn(d["iters"], d["H"], d["B"], d["C"], d["A"], generator=g)
es = torch.randn(d["iters"], d["H"], d["B"] * d["C"], d["S"], generator=g)
mods = pu.build_gpu_models(d, trans, reward_sd=reward)
pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
noise = dict(eps_act=ea.cuda(), eps_s=es.cuda())
out_sh = pl(b0.cuda(), s0.cuda(), noise=noise, trace=True)
tk_sh = pl.last_trace["topk"].clone()
pl.shard_candidates = False
out_1 = pl(b0.cuda(), s0.cuda(), noise=noise, trace=True)
tk_1 = torch.sort(pl.last_trace["topk"], dim=2)[0]
ok = torch.equal(torch.sort(tk_sh, dim=2)[0], tk_1) and float((out_sh - out_1).abs().max()) < 1e-5
print(f"rank {rank}/{world}: sharded CEM elites equal={torch.equal(torch.sort(tk_sh, dim=2)[0], tk_1)} "
      f"action diff={float((out_sh - out_1).abs().max()):.2e} OK={ok}", flush=True)
import torch.distributed as tdist
tdist.barrier(); tdist.destroy_process_group()
sys.exit(0 if ok else 1)
