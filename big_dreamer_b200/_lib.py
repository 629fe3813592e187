"""ctypes binding of libbd_b200.so (C ABI: include/bd_b200.h).

The structures below mirror the header field by field.  The library is built
in-tree by ``big_dreamer_b200/csrc/Makefile`` (``__graft_entry__.build()``); if it
is missing the import of any compute entry point raises -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BD_B200_LIB: another build of the same library (A/B timing of kernel changes); default = the in-tree build
LIB_PATH = os.environ.get("BD_B200_LIB") or os.path.join(_HERE, "libbd_b200.so")
TEST_LIB_PATH = os.path.join(_HERE, "libbd_b200_test.so")

BD_MAX_LAYERS = 8
ACTIVATIONS = {"Identity": 0, "ELU": 1, "ReLU": 2, "Tanh": 3, "Sigmoid": 4}
PRECISIONS = {"fp32": 0, "bf16": 1, "tf32": 2, "fp16": 3}

f32p = C.POINTER(C.c_float)
i64p = C.POINTER(C.c_int64)


class Linear(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("in_features", C.c_int),
                ("out_features", C.c_int)]


class Mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("activation", C.c_int), ("layer", Linear * BD_MAX_LAYERS)]


class Rssm(C.Structure):
    _fields_ = [("belief_size", C.c_int), ("state_size", C.c_int), ("action_size", C.c_int),
                ("hidden_size", C.c_int), ("embedding_size", C.c_int), ("activation", C.c_int),
                ("min_std_dev", C.c_float), ("embed", Linear),
                ("w_ih", C.c_void_p), ("w_hh", C.c_void_p), ("b_ih", C.c_void_p),
                ("b_hh", C.c_void_p), ("prior1", Linear), ("prior2", Linear), ("post1", Linear),
                ("post2", Linear)]


class RssmGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("embed_w", "embed_b", "w_ih", "w_hh", "b_ih", "b_hh", "prior1_w", "prior1_b",
                 "prior2_w", "prior2_b", "post1_w", "post1_b", "post2_w", "post2_b")]


class ActorCfg(C.Structure):
    _fields_ = [("mean_scale", C.c_float), ("raw_init_std", C.c_float), ("min_std", C.c_float),
                ("entropy_samples", C.c_int)]


class MlpBwdArgs(C.Structure):
    _fields_ = [("x1", C.c_void_p), ("k1", C.c_int), ("x2", C.c_void_p), ("k2", C.c_int),
                ("rows", C.c_int64), ("dy", C.c_void_p), ("dx1", C.c_void_p), ("dx2", C.c_void_p),
                ("dw", C.c_void_p * BD_MAX_LAYERS), ("db", C.c_void_p * BD_MAX_LAYERS),
                ("saved", C.c_void_p)]


class TransitionArgs(C.Structure):
    _fields_ = [("rssm", Rssm), ("L", C.c_int), ("B", C.c_int64)] + \
               [(n, C.c_void_p) for n in
                ("init_state", "init_belief", "actions", "embeddings", "nonterminals", "eps_prior",
                 "eps_post", "beliefs", "prior_states", "prior_means", "prior_stds", "post_states",
                 "post_means", "post_stds")]


class TransitionBwdArgs(C.Structure):
    _fields_ = [("fwd", TransitionArgs)] + \
               [(n, C.c_void_p) for n in
                ("g_beliefs", "g_prior_states", "g_prior_means", "g_prior_stds", "g_post_states",
                 "g_post_means", "g_post_stds", "d_init_state", "d_init_belief", "d_actions",
                 "d_embeddings")] + [("grads", RssmGrads)]


class ImagineArgs(C.Structure):
    _fields_ = [("rssm", Rssm), ("actor", Mlp), ("actor_cfg", ActorCfg), ("T", C.c_int),
                ("N", C.c_int64)] + \
               [(n, C.c_void_p) for n in
                ("prev_state", "prev_belief", "eps_a", "eps_e", "eps_s", "beliefs", "states",
                 "means", "stds", "entropy", "actions", "actor_raw", "dent", "tc_saved")]


class ImagineBwdArgs(C.Structure):
    _fields_ = [("fwd", ImagineArgs)] + \
               [(n, C.c_void_p) for n in
                ("g_beliefs", "g_states", "g_means", "g_stds", "g_entropy", "d_prev_state",
                 "d_prev_belief")] + \
               [("actor_dw", C.c_void_p * BD_MAX_LAYERS), ("actor_db", C.c_void_p * BD_MAX_LAYERS)]


class ImagineReturnsArgs(C.Structure):
    _fields_ = [("img", ImagineArgs), ("reward", Mlp), ("value", Mlp), ("discount", C.c_double),
                ("lambda_", C.c_double)] + \
               [(n, C.c_void_p) for n in ("reward_out", "value_out", "returns", "heads_saved")]


class ImagineReturnsBwdArgs(C.Structure):
    _fields_ = [("fwd", ImagineReturnsArgs)] + \
               [(n, C.c_void_p) for n in
                ("g_beliefs", "g_states", "g_means", "g_stds", "g_entropy", "g_reward", "g_value",
                 "g_returns", "d_prev_state", "d_prev_belief")] + \
               [("actor_dw", C.c_void_p * BD_MAX_LAYERS), ("actor_db", C.c_void_p * BD_MAX_LAYERS)]


class CemEvalArgs(C.Structure):
    _fields_ = [("rssm", Rssm), ("reward", Mlp), ("B", C.c_int), ("C", C.c_int), ("H", C.c_int),
                ("c_begin", C.c_int), ("c_end", C.c_int)] + \
               [(n, C.c_void_p) for n in
                ("belief", "state", "action_mean", "action_std", "eps_act", "eps_s", "actions",
                 "returns")]


class CemPlanArgs(C.Structure):
    _fields_ = [("rssm", Rssm), ("reward", Mlp), ("B", C.c_int), ("C", C.c_int), ("K", C.c_int),
                ("H", C.c_int), ("iters", C.c_int)] + \
               [(n, C.c_void_p) for n in
                ("belief", "state", "eps_act", "eps_s", "action_out", "returns_trace",
                 "topk_trace")]


# name -> (restype, argtypes); also the list of symbols the header declares
SIGNATURES = {
    "bd_version": (C.c_int, []),
    "bd_last_error": (C.c_char_p, []),
    "bd_launch_count": (C.c_ulonglong, []),
    "bd_precision_supported": (C.c_int, [C.c_int]),
    "bd_prof_enable": (None, [C.c_int]),
    "bd_prof_read": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "bd_mlp_workspace_bytes": (C.c_size_t, [C.POINTER(Mlp), C.c_int64, C.c_int]),
    "bd_mlp_forward": (C.c_int, [C.POINTER(Mlp), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                 C.c_void_p]),
    "bd_mlp_saved_bytes": (C.c_size_t, [C.POINTER(Mlp), C.c_int, C.c_int, C.c_int64, C.c_int]),
    "bd_mlp_forward_save": (C.c_int, [C.POINTER(Mlp), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_int, C.c_void_p]),
    "bd_mlp_backward": (C.c_int, [C.POINTER(Mlp), C.POINTER(MlpBwdArgs), C.c_void_p, C.c_size_t,
                                  C.c_int, C.c_void_p]),
    "bd_heads_forward_supported": (C.c_int, [C.POINTER(Mlp), C.POINTER(Mlp), C.c_int, C.c_int, C.c_int]),
    "bd_heads_forward_workspace_bytes": (C.c_size_t, [C.POINTER(Mlp), C.POINTER(Mlp)]),
    "bd_heads_backward_workspace_bytes": (C.c_size_t, [C.POINTER(Mlp), C.POINTER(Mlp), C.c_int, C.c_int]),
    "bd_heads_backward": (C.c_int, [C.POINTER(Mlp), C.POINTER(Mlp), C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_int, C.c_void_p]),
    "bd_heads_forward": (C.c_int, [C.POINTER(Mlp), C.POINTER(Mlp), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_size_t, C.c_int, C.c_void_p]),
    "bd_kl_loss_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                     C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_kl_loss_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                      C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_value_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
    "bd_actor_act": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(ActorCfg), C.c_int64, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p]),
    "bd_lambda_return_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                           C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "bd_lambda_return_backward": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_transition_workspace_bytes": (C.c_size_t, [C.POINTER(Rssm), C.c_int, C.c_int64, C.c_int,
                                                   C.c_int]),
    "bd_transition_forward": (C.c_int, [C.POINTER(TransitionArgs), C.c_void_p, C.c_size_t, C.c_int,
                                        C.c_void_p]),
    "bd_transition_backward": (C.c_int, [C.POINTER(TransitionBwdArgs), C.c_void_p, C.c_size_t,
                                         C.c_int, C.c_void_p]),
    "bd_imagine_workspace_bytes": (C.c_size_t, [C.POINTER(Rssm), C.POINTER(Mlp), C.c_int,
                                                C.c_int64, C.c_int]),
    "bd_imagine_saved_bytes": (C.c_size_t, [C.POINTER(Rssm), C.c_int, C.c_int64, C.c_int]),
    "bd_imagine_forward": (C.c_int, [C.POINTER(ImagineArgs), C.c_void_p, C.c_size_t, C.c_int,
                                     C.c_void_p]),
    "bd_imagine_backward": (C.c_int, [C.POINTER(ImagineBwdArgs), C.c_void_p, C.c_size_t, C.c_int,
                                      C.c_void_p]),
    "bd_imagine_returns_supported": (C.c_int, [C.POINTER(Rssm), C.POINTER(Mlp), C.POINTER(Mlp),
                                               C.POINTER(Mlp), C.c_int]),
    "bd_imagine_returns_workspace_bytes": (C.c_size_t, [C.POINTER(ImagineReturnsArgs), C.c_int]),
    "bd_imagine_returns_saved_bytes": (C.c_size_t, [C.POINTER(ImagineReturnsArgs)]),
    "bd_imagine_returns_forward": (C.c_int, [C.POINTER(ImagineReturnsArgs), C.c_void_p, C.c_size_t,
                                             C.c_int, C.c_void_p]),
    "bd_imagine_returns_backward": (C.c_int, [C.POINTER(ImagineReturnsBwdArgs), C.c_void_p, C.c_size_t,
                                              C.c_int, C.c_void_p]),
    "bd_cem_workspace_bytes": (C.c_size_t, [C.POINTER(Rssm), C.POINTER(Mlp), C.c_int, C.c_int,
                                            C.c_int]),
    "bd_cem_evaluate": (C.c_int, [C.POINTER(CemEvalArgs), C.c_void_p, C.c_size_t, C.c_int,
                                  C.c_void_p]),
    "bd_cem_refit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bd_cem_plan_workspace_bytes": (C.c_size_t, [C.POINTER(Rssm), C.POINTER(Mlp), C.c_int, C.c_int,
                                                 C.c_int, C.c_int]),
    "bd_cem_plan": (C.c_int, [C.POINTER(CemPlanArgs), C.c_void_p, C.c_size_t, C.c_int,
                              C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


# libbd_b200_test.so (include/bd_b200_test.h): self-test + micro-benchmarks, tests/ and scripts/ only
TEST_SIGNATURES = {
    "bd_tc_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    # debug micro-benchmarks (scripts/mmabench*.py, scripts/dsmembench.py)
    "bd_tc_mmabench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_tc_mmabench2": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "bd_tc_dsmembench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bd_test_last_error": (C.c_char_p, []),
}


class BdError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library and bind every symbol; fail loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise BdError(f"{LIB_PATH} not found: build it with `make -C big_dreamer_b200/csrc` "
                      "(__graft_entry__.build()). There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


_test_lib = None


def load_test() -> C.CDLL:
    """The separate debug library (never loaded by the product path)."""
    global _test_lib
    if _test_lib is None:
        if not os.path.isfile(TEST_LIB_PATH):
            raise BdError(f"{TEST_LIB_PATH} not found: build it with `make -C big_dreamer_b200/csrc`")
        lib = C.CDLL(TEST_LIB_PATH)
        for name, (res, args) in TEST_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _test_lib = lib
    return _test_lib


def check_test(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_test().bd_test_last_error()
        raise BdError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().bd_last_error()
        raise BdError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


# ------------------------------------------------------------------ helpers
def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise BdError("big_dreamer_b200 runs on CUDA tensors only (no CPU fallback)")
    if t.dtype not in (torch.float32, torch.int64):
        raise BdError(f"expected float32 tensor, got {t.dtype}")
    if not t.is_contiguous():
        raise BdError("expected a contiguous tensor")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def activation_id(act) -> int:
    name = act if isinstance(act, str) else getattr(act, "__name__", type(act).__name__)
    if name not in ACTIVATIONS:
        raise BdError(f"activation {name!r} is not implemented by the B200 path "
                      f"(supported: {sorted(ACTIVATIONS)})")
    return ACTIVATIONS[name]


def make_linear(w: torch.Tensor, b: torch.Tensor) -> Linear:
    return Linear(ptr(w), ptr(b), w.shape[1], w.shape[0])


def make_mlp(weights, biases, act_id: int) -> Mlp:
    if len(weights) > BD_MAX_LAYERS:
        raise BdError(f"MLP with {len(weights)} layers exceeds BD_MAX_LAYERS={BD_MAX_LAYERS}")
    m = Mlp()
    m.n_layers, m.activation = len(weights), act_id
    for i, (w, b) in enumerate(zip(weights, biases)):
        m.layer[i] = make_linear(w, b)
    return m


_ws_cache = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Per-(device, stream) scratch tensor (grown on demand; owned by PyTorch's allocator).
    During a CUDA-graph capture the scratch comes from the graph's private pool and is NOT cached:
    a cached entry would outlive its graph and leak into later captures on the same capture stream."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
    key = (torch.device(device).index, torch.cuda.current_stream().cuda_stream)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t
