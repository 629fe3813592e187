import torch, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
d = dict(Be=200, Hi=200, S=30, A=1, E=8, H=15, act="ELU")
trans, actor, reward, value = orc.make_models(0, 200, 30, 1, 200, 8)
mods = pu.build_gpu_models(d, trans, actor, reward, value)
pu.freeze(mods.transition, mods.reward, mods.critic)
agent = pu.agent_ns(mods, 15)
sizes = [int(x) for x in os.environ.get("SIZES", "2500,18944,131072").split(",")]
for N in sizes:
    s0, b0 = orc.make_latents(0, N, 200, 30)
    s0, b0 = s0.cuda(), b0.cuda()
    noise = bd.draw_imagine_noise(14, N, 30, 1, "cuda")
    for prec in os.environ.get("PRECS", "fp32,fp16").split(","):
        bd.set_precision(prec)
        with torch.no_grad():
            for _ in range(3): bd.imagine_ahead(agent, s0[None], b0[None], noise)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): bd.imagine_ahead(agent, s0[None], b0[None], noise)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2 * (31*200 + 6*200*200 + 200*200 + 60*200 + 230*200 + 3*200*200 + 2*200) * N * 14
        print(f"imagine fwd N={N} {prec}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s  {N*14/ms/1e3:.2f} Msteps/s", flush=True)
