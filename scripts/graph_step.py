"""One captured actor-loss step (c2 sizes) replayed a few times: for ncu launch lists of the graph path."""
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
pu.freeze(mods.transition, mods.reward, mods.critic)
agent = pu.agent_ns(mods, 15)
s0, b0 = orc.make_latents(0, N, 200, 30)
s0, b0 = s0.cuda(), b0.cuda()
noise = bd.draw_imagine_noise(14, N, 30, 1, "cuda")
params = list(mods.actor.parameters())
def fn():
    for p in params:
        p.grad = None
    b, s, _, ent = bd.imagine_ahead(agent, s0[None], b0[None], noise)
    rew, val = mods.reward(b, s), mods.critic(b, s)
    ret = bd.lambda_return(rew, val, val[-1], 0.995, 0.95)
    loss = -(ret + 1e-5 * ent.unsqueeze(-1)).mean()
    loss.backward()
    return loss.detach()
step = bd.CapturedStep(fn, warmup=int(os.environ.get("WARM", 3)))
for _ in range(int(os.environ.get("REPS", 3))):
    step.replay()
torch.cuda.synchronize()
print("ok", float(step.outputs))
