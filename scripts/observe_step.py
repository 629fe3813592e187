"""One observe pass (BASELINE configs[3]) fwd+bwd for ncu launch lists."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from tests import parity_utils as pu
from oracle import rssm_oracle as orc
d = dict(Be=200, Hi=200, S=30, A=1, E=1024, act="ELU")
bd.set_precision(os.environ.get("PREC", "fp16"))
trans, _, _, _ = orc.make_models(0, 200, 30, 1, 200, 1024)
tm = pu.build_gpu_models(d, trans).transition
g = torch.Generator().manual_seed(0)
L, B = 49, 50
s0, b0 = orc.make_latents(0, B, 200, 30)
s0, b0 = s0.cuda(), b0.cuda()
actions = (torch.rand(L, B, 1, generator=g) * 2 - 1).cuda()
emb = torch.randn(L, B, 1024, generator=g).cuda()
nt = torch.ones(L, B, 1, device="cuda")
for i in range(int(os.environ.get("REPS", 3))):
    for p in tm.parameters():
        p.grad = None
    o = tm(s0, actions, b0, emb, nt)
    (o[0].mean() + o[3].mean() + o[4][0].mean() + o[4][1].mean() + o[2][0].mean()).backward()
torch.cuda.synchronize()
print("ok")
import time
def one():
    for p in tm.parameters():
        p.grad = None
    o = tm(s0, actions, b0, emb, nt)
    (o[0].mean() + o[3].mean() + o[4][0].mean() + o[4][1].mean() + o[2][0].mean()).backward()
for _ in range(2):
    one()
torch.cuda.synchronize()
t0 = time.perf_counter(); one(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms")
