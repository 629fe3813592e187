"""A few acting steps (bd.ActPath, actor policy, eager) for an ncu launch list."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
d = dict(Be=200, Hi=200, S=30, A=1, E=1024, act="ELU")
bd.set_precision(os.environ.get("PREC", "fp16"))
trans, actor, reward, _ = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
mods = pu.build_gpu_models(d, trans, actor, reward_sd=reward)
act = bd.ActPath(mods.transition, mods.actor, batch=1, action_noise=0.3)
z = lambda n: torch.zeros(1, n, device="cuda")
emb = torch.randn(1, d["E"], device="cuda")
b, s, a = z(d["Be"]), z(d["S"]), z(d["A"])
for _ in range(int(os.environ.get("REPS", 3))):
    b, s, a = act(b, s, a, emb, explore=True)
torch.cuda.synchronize()
print("ok")
