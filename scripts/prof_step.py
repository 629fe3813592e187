"""Eager actor-loss steps through bd.imagine_and_returns (FUSED=1/0 forces the fused / piecewise path): timing,
and with BD_TC_PROF=1 the in-kernel cycle counters of the BPTT kernel (printed by the library on stderr)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
d = dict(Be=200, Hi=200, S=30, A=1, E=8, H=15, act="ELU")
N = int(os.environ.get("N", 18944))
fused = {"1": True, "0": False}.get(os.environ.get("FUSED", ""), None)
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
    out = bd.imagine_and_returns(agent, s0[None], b0[None], mods.reward, mods.critic, 0.995, 0.95, noise, fused=fused)
    loss = -(out[6] + 1e-5 * out[3].unsqueeze(-1)).mean()
    loss.backward()
    return loss.detach()
reps = int(os.environ.get("REPS", 3))
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
print("ok: %.3f ms per step (N=%d, fused=%s)" % (e0.elapsed_time(e1) / reps, N, fused))
