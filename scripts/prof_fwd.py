import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
d = dict(Be=200, Hi=200, S=30, A=1, E=8, H=15, act=os.environ.get("ACT", "ELU"))
trans, actor, reward, value = orc.make_models(0, 200, 30, 1, 200, 8)
mods = pu.build_gpu_models(d, trans, actor, reward, value)
pu.freeze(mods.transition, mods.reward, mods.critic)
agent = pu.agent_ns(mods, 15)
N = int(os.environ.get("N", 18944))
s0, b0 = orc.make_latents(0, N, 200, 30)
s0, b0 = s0.cuda(), b0.cuda()
noise = bd.draw_imagine_noise(14, N, 30, 1, "cuda")
bd.set_precision(os.environ.get("PREC", "fp16"))
with torch.no_grad():
    for _ in range(2):
        bd.imagine_ahead(agent, s0[None], b0[None], noise)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    for _ in range(5):
        bd.imagine_ahead(agent, s0[None], b0[None], noise)
    e1.record()
torch.cuda.synchronize()
print("ok: imagine_ahead forward %.1f us per call (N=%d, CUDA events, 5 calls incl. packing + entropy)" % (e0.elapsed_time(e1) * 200, N))

if os.environ.get("BD_TC_PROF"):
    from big_dreamer_b200 import _lib
    import numpy as np
    ws = list(_lib._ws_cache.values())[0]
    # find the counters: they sit at the 4096-aligned offset after the packed weights; scan for them
    raw = ws.cpu().numpy()
    pack_bytes = int(os.environ["PACK_BYTES"]) if "PACK_BYTES" in os.environ else None
    import struct
    for off in range(0, 4 << 20, 4096):
        v = np.frombuffer(raw[off:off + 40 * 64].tobytes(), dtype=np.int64).reshape(40, 8)
        if 0 < v[0, 2] < 10**9 and 0 < v[0, 4] < 10**9 and v[39, 0] > 0 and v[13:20].sum() == 0 and 0 < v[39, 1] < 10**10:
            names = ["actorL0", "actorL1", "actorL2", "actorL3", "actorOut", "embed", "gru0", "gru1", "gru2", "gru3", "prior1", "priorOut"]
            nph = int((v[:20, 2] > 0).sum())
            if nph != 12:
                names = ["actorL0", "actorL1", "actorL2", "actorL3", "actorOut", "embed"] + ["gru%d" % i for i in range(nph - 8)] + ["prior1", "priorOut"]
            print("phase      iss_dep  iss_wwait iss_issue | epi0_wait epi0_work | epi1_wait epi1_work | mma_loop  (cycles per step, CTA 0, last launch, 14 steps)")
            for i, n in enumerate(names):
                print(f"{n:9s}", " ".join(f"{int(x)//14:9d}" for x in v[i, :8]))
            tot = v[:20, :8].sum(0) // 14
            print("total    ", " ".join(f"{int(x):9d}" for x in tot))
            print("per-stage commit cycles per phase:", [int(x) // 14 for x in v[20:20 + len(names), 0]])
            w = np.frombuffer(raw[off + 40 * 64: off + 40 * 64 + 160 * 24].tobytes(), dtype=np.int64).reshape(160, 3)
            w = w[w[:, 0] > 0]
            t0 = w[:, 0].min()
            print("CTAs:", len(w), "start spread us", (w[:, 0].max() - t0) / 1e3, "end min/max us", (w[:, 1].min() - t0) / 1e3, (w[:, 1].max() - t0) / 1e3)
            dur = (w[:, 1] - w[:, 0]) / 1e3
            print("per-CTA duration us: min %.0f median %.0f max %.0f" % (dur.min(), np.median(dur), dur.max()), "distinct SMs", len(set(w[:, 2].tolist())))
            print("first 12 durations", dur[:12].round().tolist())
            print("CTA0 kernel cycles per launch", int(v[39, 0]), "ns", int(v[39, 1]), "-> GHz", v[39, 0] / max(1, v[39, 1]))
            break
