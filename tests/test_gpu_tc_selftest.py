"""tcgen05 building block in isolation: descriptors, KM8 layout, bulk-TMA weight delivery, TMEM
read-back.  Reference = fp32 matmul of the 16-bit-rounded operands (exact up to fp32 summation)."""
import pytest
import torch

import big_dreamer_b200 as bd
from big_dreamer_b200 import _lib

pytestmark = pytest.mark.gpu


def run(K, N, fmt, swap=0, seed=0):
    lib = _lib.load_test()          # libbd_b200_test.so: the debug entry points are not in the product library
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(128, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    y = torch.full((128, N), float("nan"), device="cuda")
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    _lib.check_test(lib.bd_tc_selftest(x.data_ptr(), w.data_ptr(), b.data_ptr(), K, N, fmt, swap,
                                  ws.data_ptr(), ws.numel(), y.data_ptr(), _lib.stream_ptr()),
               "bd_tc_selftest")
    torch.cuda.synchronize()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    ref = x.to(dt).float() @ w.to(dt).float().t() + b.to(dt).float()
    return y, ref


@pytest.mark.parametrize("K,N", [(15, 16), (31, 208), (200, 200), (230, 200), (200, 64), (47, 256)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_tc_linear_tile(K, N, fmt):
    y, ref = run(K, N, fmt)
    err = float((y - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err


@pytest.mark.parametrize("d_col,N", [(256, 208), (448, 64), (304, 208), (496, 16), (192, 64)])
def test_tc_accumulator_column_offsets(d_col, N):
    y, ref = run(200, N, 0, swap=d_col << 8)
    err = float((y - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, (d_col, N, err)


@pytest.mark.parametrize("K,N", [(200, 200), (47, 64), (230, 208)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_tc_a_operand_from_tmem(K, N, fmt):
    """TS mode: the A operand is written to tensor memory with tcgen05.st and read by the MMA from
    there (the round-2 plan for chain layers); same numerics as the shared-memory operand."""
    y, ref = run(K, N, fmt, swap=2)
    err = float((y - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err


@pytest.mark.parametrize("K,N", [(200, 200), (47, 64), (230, 208), (31, 16)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_tc_warp_wide_elect_issue(K, N, fmt):
    """The engine's issuer structure: the whole warp runs the loop on warp-uniform descriptors and
    each tcgen05.mma / commit is one instruction predicated on the elected lane."""
    y, ref = run(K, N, fmt, swap=4)
    err = float((y - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err
