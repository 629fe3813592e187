import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc

d = dict(Be=200, Hi=200, S=30, A=1, E=8, N=int(os.environ.get("N", 130)), H=4, act="ELU")
prec = os.environ.get("PREC", "fp16")
trans, actor, reward, value = orc.make_models(3, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
actor["model.8.bias"][d["A"]:] -= 6.0
s0, b0 = orc.make_latents(3, d["N"], d["Be"], d["S"])
ea, ee, es = orc.make_imagine_noise(3, d["H"] - 1, d["N"], d["S"], d["A"])
mods = pu.build_gpu_models(d, trans, actor)
pu.freeze(mods.transition)
noise = dict(eps_a=ea.cuda(), eps_e=ee.cuda(), eps_s=es.cuda())
with torch.no_grad():
    ob, os_, (om, osd), oe, oa = orc.imagine_ahead(trans, actor, "ELU", 0.1, d["H"], s0[None], b0[None], ea, ee, es)
    res = {}
    for p in ("fp32", prec):
        bd.set_precision(p)
        res[p] = bd.imagine_ahead(pu.agent_ns(mods, d["H"]), s0.cuda()[None], b0.cuda()[None], noise, return_actions=True)
        from big_dreamer_b200 import functions as F_
        res[p + "_raw"] = F_._debug_last["actor_raw"].clone()
torch.cuda.synchronize()
gb, gs, (gm, gsd), ge, ga = res[prec]
def e(a, b):
    a = a.float().cpu(); return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
for t in range(d["H"] - 1):
    print(f"t={t}: action {e(ga[t], oa[t]):.2e} entropy {e(ge[t], oe[t]):.2e} "
          f"belief {e(gb[t], ob[t]):.2e} [slices "
          + " ".join(f"{e(gb[t][:, c:c+64], ob[t][:, c:c+64]):.1e}" for c in range(0, d['Be'], 64))
          + f"] mean {e(gm[t], om[t]):.2e} std {e(gsd[t], osd[t]):.2e} state {e(gs[t], os_[t]):.2e}")
    if t == 0:
        rows = (gb[0].cpu() - ob[0]).abs().max(dim=1)[0]
        print("  belief err by row block:", [f"{float(rows[i:i+32].max()):.1e}" for i in range(0, d["N"], 32)])

r32, r16_ = res["fp32_raw"][0].cpu(), res[prec + "_raw"][0].cpu()
print("actor_raw t=0 fp32 :", r32[:4].flatten().tolist())
print("actor_raw t=0 tc   :", r16_[:4].flatten().tolist())
print("actor_raw err", e(res[prec + "_raw"][0], r32))
db = (gb[0].cpu() - ob[0]).abs()
print("belief t=0 err per 16-col chunk:", [f"{float(db[:, c:c+16].max()):.1e}" for c in range(0, 200, 16)])
print("tc belief[0][0, 64:72]", gb[0][0, 64:72].tolist())
print("ref belief[0][0, 64:72]", ob[0][0, 64:72].tolist())
print("b0[0,64:72]", b0[0, 64:72].tolist())
