"""GPU parity, tensor-core modes (tcgen05, 16-bit operands, fp32 accumulation and state).
Tolerance (north_star): max rel err <= 1e-2 after the full horizon."""
import pytest
import torch

import big_dreamer_b200 as bd
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu
TOL = 1e-2

CASES = [
    dict(Be=200, Hi=200, S=30, A=1, E=8, N=300, H=15, act="ELU"),
    dict(Be=32, Hi=32, S=30, A=1, E=8, N=130, H=15, act="ELU"),
    dict(Be=200, Hi=200, S=30, A=6, E=8, N=77, H=6, act="ELU"),
    dict(Be=48, Hi=40, S=10, A=3, E=8, N=129, H=7, act="ReLU"),
    dict(Be=24, Hi=56, S=12, A=2, E=8, N=1, H=4, act="Tanh"),
]


@pytest.fixture(autouse=True)
def _restore():
    yield
    bd.set_precision("fp32")


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("d", CASES)
def test_imagine_actor_loss_tc(d, prec):
    res = pu.run_imagine_case(d, seed=3, precision=prec, oracle_dtype=torch.float64)
    errs = res["errors"]
    print(prec, d, {k: f"{v:.2e}" for k, v in errs.items()})
    tol = TOL if prec == "fp16" else 3e-2
    for k, e in errs.items():
        assert e < tol, errs
