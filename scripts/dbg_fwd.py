import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity_utils as pu
for d in (dict(Be=200, Hi=200, S=30, A=1, E=8, N=int(os.environ.get("N", 300)), H=int(os.environ.get("H", 2)), act="ELU"),
          dict(Be=32, Hi=32, S=30, A=1, E=8, N=130, H=3, act="ELU")):
    res = pu.run_imagine_case(d, seed=3, precision="fp16", oracle_dtype=torch.float64)
    print(d, {k: f"{v:.2e}" for k, v in res["errors"].items()}, flush=True)
