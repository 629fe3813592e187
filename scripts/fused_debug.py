"""Debug driver for the fused imagine_and_returns entry (forward, then forward + backward), one config."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
N = int(os.environ.get("N", 300))
d = dict(Be=200, Hi=200, S=30, A=1, E=8, N=N, H=15, act="ELU")
bd.set_precision(os.environ.get("PREC", "fp16"))
trans, actor, reward, value = orc.make_models(7, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
actor["model.8.bias"][d["A"]:] -= 6.0
s0, b0 = orc.make_latents(7, d["N"], d["Be"], d["S"])
ea, ee, es = orc.make_imagine_noise(7, d["H"] - 1, d["N"], d["S"], d["A"])
mods = pu.build_gpu_models(d, trans, actor, reward, value)
pu.freeze(mods.transition, mods.reward, mods.critic)
agent = pu.agent_ns(mods, d["H"])
noise = dict(eps_a=ea.cuda(), eps_e=ee.cuda(), eps_s=es.cuda())
ref = pu.oracle_actor_loss(d, trans, actor, reward, value, s0, b0, ea, ee, es, dtype=torch.float64)
with torch.no_grad():
    out = bd.imagine_and_returns(agent, s0[None].cuda(), b0[None].cuda(), mods.reward, mods.critic, 0.995, 0.95, noise)
torch.cuda.synchronize()
names = ("beliefs", "states", None, "entropy", "reward", "value", "returns")
for n, o in zip(names, out):
    if n:
        print("fwd", n, "%.3e" % pu.relerr(o, ref[1][n]), flush=True)
if os.environ.get("BWD", "1") == "1":
    out = bd.imagine_and_returns(agent, s0[None].cuda(), b0[None].cuda(), mods.reward, mods.critic, 0.995, 0.95, noise)
    loss = -(out[6] + 1e-5 * out[3].unsqueeze(-1)).mean()
    torch.cuda.synchronize()
    print("fwd(grad) ok", flush=True)
    loss.backward()
    torch.cuda.synchronize()
    errs = {k: pu.relerr(p.grad, ref[2][k]) for k, p in mods.actor.named_parameters()}
    print("bwd actor grads max err %.3e" % max(errs.values()), {k: "%.1e" % v for k, v in errs.items()}, flush=True)
print("done")
