"""Weight-share clusters (BD_TC_WS): bit-identity of the rollout outputs against ws = 1, and timing."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
d = dict(Be=200, Hi=200, S=30, A=1, E=8, H=15, act="ELU")
trans, actor, reward, value = orc.make_models(0, 200, 30, 1, 200, 8)
mods = pu.build_gpu_models(d, trans, actor, reward, value)
pu.freeze(mods.transition, mods.reward, mods.critic)
agent = pu.agent_ns(mods, 15)
bd.set_precision("fp16")
os.environ["BD_TC_CLUSTER"] = "1"
for N in [int(x) for x in os.environ.get("NS", "1024,18944,37888").split(",")]:
    s0, b0 = orc.make_latents(0, N, 200, 30)
    s0, b0 = s0.cuda(), b0.cuda()
    noise = bd.draw_imagine_noise(14, N, 30, 1, "cuda")
    ref = None
    for ws in ("1", "2", "4"):
        os.environ["BD_TC_WS"] = ws
        with torch.no_grad():
            out = bd.imagine_ahead(agent, s0[None], b0[None], noise)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                bd.imagine_ahead(agent, s0[None], b0[None], noise)
            e1.record()
            torch.cuda.synchronize()
        flat = [out[0], out[1], out[2][0], out[2][1], out[3]]
        if ref is None:
            ref = [t.clone() for t in flat]
        same = all(torch.equal(a, b) for a, b in zip(flat, ref))
        print(f"N={N} ws={ws}: {e0.elapsed_time(e1) * 200:.1f} us per forward, identical to ws=1: {same}", flush=True)
