import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import big_dreamer_b200 as bd
from oracle import rssm_oracle as orc
bd.set_precision("fp16")
g = torch.Generator().manual_seed(0)
sd = orc.make_mlp_sd(g, [230, 200, 200, 200, 200, 1])
dm = bd.DenseModel(230, 200).cuda(); dm.load_state_dict(sd)
for p in dm.parameters(): p.requires_grad_(os.environ.get("WGRAD", "0") == "1")
rows = int(os.environ.get("ROWS", 148 * 128 * 4))
b = torch.randn(rows, 200, device="cuda", requires_grad=True); s = torch.randn(rows, 30, device="cuda", requires_grad=True)
for _ in range(2):
    y = dm(b, s); y.sum().backward()
torch.cuda.synchronize(); print("ok")
