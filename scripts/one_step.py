"""One bench step (c2 by default) for ncu launch lists."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
d = dict(Be=200, Hi=200, S=30, A=1, E=8, H=15, act="ELU")
N = int(os.environ.get("N", 2500))
bd.set_precision(os.environ.get("PREC", "fp16"))
trans, actor, reward, value = orc.make_models(0, 200, 30, 1, 200, 8)
mods = pu.build_gpu_models(d, trans, actor, reward, value)
s0, b0 = orc.make_latents(0, N, 200, 30)
noise = bd.draw_imagine_noise(14, N, 30, 1, "cuda")
for i in range(int(os.environ.get("REPS", 3))):
    pu.gpu_actor_loss(mods, 15, s0.cuda(), b0.cuda(), noise)
torch.cuda.synchronize()
print("ok")
