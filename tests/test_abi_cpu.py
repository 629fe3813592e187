"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/bd_b200.h declares, the ctypes structures match the header's layout, the drop-in
modules keep the reference's state_dict keys, and the product path refuses CPU tensors."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

import big_dreamer_b200 as bd
from big_dreamer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bd_b200.h")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.isfile(_lib.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "big_dreamer_b200", "csrc"), "-j8"],
                       check=True, capture_output=True)


def header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"\b(bd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = bd.load_library()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in bd_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.bd_version() == 1
    assert lib.bd_precision_supported(0) == 1


def test_debug_entry_points_live_in_the_test_library_only():
    """bd_tc_selftest / bd_tc_mmabench* / bd_tc_dsmembench (include/bd_b200_test.h) are built into
    libbd_b200_test.so; the product library and its public header do not carry them."""
    lib = bd.load_library()
    tsrc = open(os.path.join(ROOT, "include", "bd_b200_test.h")).read()
    tsyms = sorted(set(re.findall(r"\b(bd_[a-z0-9_]+)\s*\(", tsrc)))
    assert "bd_tc_selftest" in tsyms and len(tsyms) >= 5
    tlib = _lib.load_test()
    for s in tsyms:
        assert hasattr(tlib, s), f"{s} declared in bd_b200_test.h but not exported by the test library"
        assert s in _lib.TEST_SIGNATURES
        assert not hasattr(lib, s), f"{s} leaked into the product library"
        assert s not in header_symbols()


def test_ctypes_struct_sizes_match_header():
    """Compile a tiny C program against the header and compare sizeof() of every struct."""
    names = {"bd_linear": _lib.Linear, "bd_mlp": _lib.Mlp, "bd_rssm": _lib.Rssm,
             "bd_rssm_grads": _lib.RssmGrads, "bd_actor_cfg": _lib.ActorCfg,
             "bd_mlp_bwd_args": _lib.MlpBwdArgs, "bd_transition_args": _lib.TransitionArgs,
             "bd_transition_bwd_args": _lib.TransitionBwdArgs, "bd_imagine_args": _lib.ImagineArgs,
             "bd_imagine_bwd_args": _lib.ImagineBwdArgs, "bd_cem_eval_args": _lib.CemEvalArgs,
             "bd_cem_plan_args": _lib.CemPlanArgs}
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in names)
    prog = f'#include <stdio.h>\n#include "bd_b200.h"\nint main(void){{{body}return 0;}}'
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "s.c"), os.path.join(td, "s")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert C.sizeof(names[n]) == int(sz), f"{n}: ctypes {C.sizeof(names[n])} != C {sz}"


def test_state_dict_keys_match_reference_layout():
    tm = bd.TransitionModel(200, 30, 1, 200, 1024)
    keys = set(tm.state_dict().keys())
    assert keys == {"rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh",
                    "fc_embed_state_action.0.weight", "fc_embed_state_action.0.bias",
                    "belief_prior.model.0.weight", "belief_prior.model.0.bias",
                    "belief_prior.model.2.weight", "belief_prior.model.2.bias",
                    "belief_posterior.model.0.weight", "belief_posterior.model.0.bias",
                    "belief_posterior.model.2.weight", "belief_posterior.model.2.bias"}
    assert tm.rnn.weight_ih.shape == (600, 200)
    assert tm.belief_posterior.model[0].weight.shape == (200, 1224)
    assert isinstance(tm.modules, list) and len(tm.modules) == 3      # reference quirk kept
    dm = bd.DenseModel(230, 200)
    assert list(dm.state_dict().keys()) == [f"model.{i}.{p}" for i in (0, 2, 4, 6, 8)
                                            for p in ("weight", "bias")]
    assert dm.model[8].weight.shape == (1, 200)


def test_reference_state_dict_loads(tmp_path):
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    mods = rh.build_modules(0, 32, 30, 2, 32, 64)
    tm = bd.TransitionModel(32, 30, 2, 32, 64)
    tm.load_state_dict(mods.transition.state_dict())          # strict
    dm = bd.DenseModel(62, 32)
    dm.load_state_dict(mods.reward.state_dict())


def test_no_cpu_fallback():
    dm = bd.DenseModel(10, 8)
    with pytest.raises(bd.BdError):
        dm(torch.zeros(4, 10))
    tm = bd.TransitionModel(8, 4, 2, 8, 16)
    with pytest.raises(bd.BdError):
        tm(torch.zeros(3, 4), torch.zeros(5, 3, 2), torch.zeros(3, 8),
           noise=dict(eps_prior=torch.zeros(5, 3, 4)))
    with pytest.raises(bd.BdError):
        bd.lambda_return(torch.zeros(3, 4, 1), torch.zeros(3, 4, 1), torch.zeros(4, 1))


def test_unsupported_configs_raise():
    with pytest.raises(NotImplementedError):
        bd.TransitionModel(8, 4, 2, 8, 16, latent_distribution="Categorical")
    with pytest.raises(bd.BdError):
        bd.DenseModel(10, 8, activation="GELU")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "big_dreamer_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, f"{fn} references the oracle"


def test_imagine_noise_is_one_buffer_with_aligned_entropy_noise():
    """bd.draw_imagine_noise: the reference's three draws (src/dreamer.py:443-444, src/models.py:72) as ONE generator
    launch -- views of one buffer with the reference's shapes, contiguous, the entropy noise first (the large-batch
    entropy kernel reads it with 16-byte loads)."""
    import torch
    import big_dreamer_b200 as bd
    T, N, S, A = 14, 2500, 30, 1
    nz = bd.draw_imagine_noise(T, N, S, A, "cpu", generator=torch.Generator().manual_seed(0))
    assert nz["eps_a"].shape == (T, N, A) and nz["eps_s"].shape == (T, N, S) and nz["eps_e"].shape == (T, 100, N, A)
    assert all(v.is_contiguous() and v.dtype == torch.float32 for v in nz.values())
    base = nz["eps_e"].untyped_storage().data_ptr()
    assert all(v.untyped_storage().data_ptr() == base for v in nz.values())
    assert nz["eps_e"].storage_offset() == 0 and nz["eps_e"].data_ptr() % 16 == 0
    flat = torch.cat([nz["eps_e"].reshape(-1), nz["eps_s"].reshape(-1), nz["eps_a"].reshape(-1)])
    assert abs(float(flat.mean())) < 5e-3 and abs(float(flat.std()) - 1.0) < 5e-3


def test_heads_pair_falls_back_to_the_two_calls_for_foreign_modules():
    """bd.heads_pair only pairs two bd.DenseModel heads on CUDA tensors; anything else is exactly
    reward_model(b, s), value_model(b, s) (src/dreamer.py:321-322)."""
    import torch
    import big_dreamer_b200 as bd

    class Head(torch.nn.Module):
        def __init__(self, scale):
            super().__init__()
            self.scale = scale

        def forward(self, b, s):
            return (b.sum(-1, keepdim=True) + s.sum(-1, keepdim=True)) * self.scale

    b, s = torch.randn(3, 5, 8), torch.randn(3, 5, 2)
    r, v = bd.heads_pair(Head(2.0), Head(-1.0), b, s)
    assert torch.equal(r, Head(2.0)(b, s)) and torch.equal(v, Head(-1.0)(b, s))
