"""torch.autograd.Function wrappers around the C ABI (include/bd_b200.h).

Every function returns ordinary contiguous fp32 CUDA tensors, so the reference's
own loss code, ``.detach()`` calls and logging keep working on the outputs
(SURVEY.md hard part 9).  Parameter gradients honour ``requires_grad`` exactly as
``FreezeParameters`` (src/utils.py:227-250) sets it.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import BdError

_precision = "fp32"


def set_precision(name: str) -> None:
    """'fp32' (check mode, FFMA kernels; the module default), 'fp16' (the fast mode: tcgen05 tensor
    cores, fp16 operands, fp32 accumulation and state) or 'bf16' (same kernels, bf16 operands)."""
    global _precision
    if name not in _lib.PRECISIONS:
        raise ValueError(f"unknown precision {name!r}")
    if not _lib.load().bd_precision_supported(_lib.PRECISIONS[name]):
        raise BdError(f"precision {name!r} is not implemented by this build of libbd_b200.so")
    _precision = name


def get_precision() -> str:
    return _precision


def _prec() -> int:
    return _lib.PRECISIONS[_precision]


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise BdError("big_dreamer_b200: CPU tensors are not supported (no CPU fallback)")
    if t.dtype != torch.float32:
        raise BdError(f"big_dreamer_b200: expected float32, got {t.dtype}")
    return t.contiguous()


def _zeros_like_if(flag: bool, p: torch.Tensor) -> Optional[torch.Tensor]:
    return torch.zeros_like(p, memory_format=torch.contiguous_format) if flag else None


def _zero_grads(flags, params):
    """Zero-initialised gradient buffers for the flagged parameters, carved out of ONE flat
    allocation (one fill kernel instead of one per parameter; the library accumulates into them)."""
    sizes = [p.numel() if f else 0 for f, p in zip(flags, params)]
    padded = [(n + 63) // 64 * 64 for n in sizes]          # keep every buffer 256-byte aligned
    total = sum(padded)
    if total == 0:
        return [None] * len(params)
    flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
    out, off = [], 0
    for n, m, p in zip(sizes, padded, params):
        out.append(flat[off:off + n].view(p.shape) if n else None)
        off += m
    return out


# =============================================================================
# MLP  (DenseModel.forward, src/models.py:393-408)
# =============================================================================
class MlpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, act_id: int, x1: torch.Tensor, x2: Optional[torch.Tensor], *params):
        lib = _lib.load()
        ws_, bs_ = [_f32c(p) for p in params[0::2]], [_f32c(p) for p in params[1::2]]
        lead = x1.shape[:-1]
        a1 = _f32c(x1).reshape(-1, x1.shape[-1])
        a2 = _f32c(x2).reshape(-1, x2.shape[-1]) if x2 is not None else None
        if a2 is not None and a2.shape[0] != a1.shape[0]:
            raise BdError("DenseModel: belief/state leading dims differ")
        k_in = a1.shape[1] + (a2.shape[1] if a2 is not None else 0)
        if k_in != ws_[0].shape[1]:
            raise BdError(f"DenseModel: input has {k_in} features, the first layer expects "
                          f"{ws_[0].shape[1]}")
        rows, out = a1.shape[0], ws_[-1].shape[0]
        mlp = _lib.make_mlp(ws_, bs_, act_id)
        y = torch.empty(rows, out, device=a1.device, dtype=torch.float32)
        nbytes = lib.bd_mlp_workspace_bytes(C.byref(mlp), rows, 0)
        ws = _lib.workspace(nbytes, a1.device)
        k2 = a2.shape[1] if a2 is not None else 0
        # tensor-core modes: keep the hidden-activation images so the backward skips recomputing them
        nsaved = lib.bd_mlp_saved_bytes(C.byref(mlp), a1.shape[1], k2, rows, _prec()) \
            if (rows and any(ctx.needs_input_grad)) else 0
        saved = torch.empty(nsaved, dtype=torch.uint8, device=a1.device) if nsaved else None
        _lib.check(lib.bd_mlp_forward_save(C.byref(mlp), _lib.ptr(a1), a1.shape[1], _lib.ptr(a2), k2,
                                           rows, _lib.ptr(y), saved.data_ptr() if nsaved else None,
                                           ws.data_ptr(), ws.numel(), _prec(), _lib.stream_ptr()),
                   "bd_mlp_forward_save")
        ctx.mlp_saved, ctx.prec = saved, _prec()
        ctx.act_id, ctx.has_x2 = act_id, x2 is not None
        ctx.x1_shape = x1.shape
        ctx.x2_shape = x2.shape if x2 is not None else None
        ctx.save_for_backward(a1, *( [a2] if a2 is not None else [] ), *ws_, *bs_)
        return y.reshape(*lead, out)

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        a1 = saved.pop(0)
        a2 = saved.pop(0) if ctx.has_x2 else None
        n = len(saved) // 2
        ws_, bs_ = saved[:n], saved[n:]
        mlp = _lib.make_mlp(ws_, bs_, ctx.act_id)
        rows, out = a1.shape[0], ws_[-1].shape[0]
        dy2 = _f32c(dy).reshape(rows, out)
        need = ctx.needs_input_grad  # (act_id, x1, x2, w0, b0, w1, b1, ...)
        dx1 = torch.empty_like(a1) if need[1] else None
        dx2 = torch.empty_like(a2) if (a2 is not None and need[2]) else None
        g = _zero_grads([need[3 + 2 * i] for i in range(n)] + [need[4 + 2 * i] for i in range(n)],
                        list(ws_) + list(bs_))
        dws, dbs = g[:n], g[n:]
        args = _lib.MlpBwdArgs()
        args.x1, args.k1 = _lib.ptr(a1), a1.shape[1]
        args.x2, args.k2 = _lib.ptr(a2), (a2.shape[1] if a2 is not None else 0)
        args.rows, args.dy = rows, _lib.ptr(dy2)
        args.dx1, args.dx2 = _lib.ptr(dx1), _lib.ptr(dx2)
        for i in range(n):
            args.dw[i], args.db[i] = _lib.ptr(dws[i]), _lib.ptr(dbs[i])
        if ctx.mlp_saved is not None:
            args.saved = ctx.mlp_saved.data_ptr()
        nbytes = lib.bd_mlp_workspace_bytes(C.byref(mlp), rows, 1)
        ws = _lib.workspace(nbytes, a1.device)
        _lib.check(lib.bd_mlp_backward(C.byref(mlp), C.byref(args), ws.data_ptr(), ws.numel(),
                                       ctx.prec, _lib.stream_ptr()), "bd_mlp_backward")
        grads: List[Optional[torch.Tensor]] = [None,
                                               dx1.reshape(ctx.x1_shape) if dx1 is not None else None,
                                               dx2.reshape(ctx.x2_shape) if dx2 is not None else None]
        for i in range(n):
            grads += [dws[i], dbs[i]]
        return tuple(grads)


def _mlp_backward_call(act_id, prec, a1, a2, ws_, bs_, dy2, mlp_saved, need_x1, need_x2, need_w, need_b):
    """One bd_mlp_backward call -> (dx1, dx2, dws, dbs); shared by MlpFunction and HeadsPairFunction."""
    lib = _lib.load()
    n = len(ws_)
    mlp = _lib.make_mlp(ws_, bs_, act_id)
    rows = a1.shape[0]
    dx1 = torch.empty_like(a1) if need_x1 else None
    dx2 = torch.empty_like(a2) if (a2 is not None and need_x2) else None
    g = _zero_grads(list(need_w) + list(need_b), list(ws_) + list(bs_))
    dws, dbs = g[:n], g[n:]
    args = _lib.MlpBwdArgs()
    args.x1, args.k1 = _lib.ptr(a1), a1.shape[1]
    args.x2, args.k2 = _lib.ptr(a2), (a2.shape[1] if a2 is not None else 0)
    args.rows, args.dy = rows, _lib.ptr(dy2)
    args.dx1, args.dx2 = _lib.ptr(dx1), _lib.ptr(dx2)
    for i in range(n):
        args.dw[i], args.db[i] = _lib.ptr(dws[i]), _lib.ptr(dbs[i])
    if mlp_saved is not None:
        args.saved = mlp_saved.data_ptr()
    nbytes = lib.bd_mlp_workspace_bytes(C.byref(mlp), rows, 1)
    ws = _lib.workspace(nbytes, a1.device)
    _lib.check(lib.bd_mlp_backward(C.byref(mlp), C.byref(args), ws.data_ptr(), ws.numel(), prec,
                                   _lib.stream_ptr()), "bd_mlp_backward")
    return dx1, dx2, dws, dbs


class HeadsPairFunction(torch.autograd.Function):
    """reward_model(beliefs, states) and value_model(beliefs, states) (src/dreamer.py:321-322) as ONE forward
    launch (bd_heads_forward: one tile prologue, the two chains interleaved); the backward is one
    bd_mlp_backward per head on the hidden images the forward left.  Tensor-core modes only."""

    @staticmethod
    def forward(ctx, act_id: int, n_layers: int, x1: torch.Tensor, x2: torch.Tensor, *params):
        lib = _lib.load()
        pr, pv = params[:2 * n_layers], params[2 * n_layers:]
        wr, br = [_f32c(p) for p in pr[0::2]], [_f32c(p) for p in pr[1::2]]
        wv, bv = [_f32c(p) for p in pv[0::2]], [_f32c(p) for p in pv[1::2]]
        lead = x1.shape[:-1]
        a1 = _f32c(x1).reshape(-1, x1.shape[-1])
        a2 = _f32c(x2).reshape(-1, x2.shape[-1])
        rows, k1, k2 = a1.shape[0], a1.shape[1], a2.shape[1]
        mr, mv = _lib.make_mlp(wr, br, act_id), _lib.make_mlp(wv, bv, act_id)
        yr = torch.empty(rows, 1, device=a1.device, dtype=torch.float32)
        yv = torch.empty(rows, 1, device=a1.device, dtype=torch.float32)
        ws = _lib.workspace(lib.bd_heads_forward_workspace_bytes(C.byref(mr), C.byref(mv)), a1.device)
        prec = _prec()
        want = rows and any(ctx.needs_input_grad)
        sr = torch.empty(lib.bd_mlp_saved_bytes(C.byref(mr), k1, k2, rows, prec), dtype=torch.uint8,
                         device=a1.device) if want else None
        sv = torch.empty(lib.bd_mlp_saved_bytes(C.byref(mv), k1, k2, rows, prec), dtype=torch.uint8,
                         device=a1.device) if want else None
        _lib.check(lib.bd_heads_forward(C.byref(mr), C.byref(mv), _lib.ptr(a1), k1, _lib.ptr(a2), k2, rows,
                                        _lib.ptr(yr), _lib.ptr(yv), sr.data_ptr() if want else None,
                                        sv.data_ptr() if want else None, ws.data_ptr(), ws.numel(), prec,
                                        _lib.stream_ptr()), "bd_heads_forward")
        ctx.saved_r, ctx.saved_v, ctx.prec, ctx.act_id, ctx.n_layers = sr, sv, prec, act_id, n_layers
        ctx.x1_shape, ctx.x2_shape = x1.shape, x2.shape
        ctx.save_for_backward(a1, a2, *wr, *br, *wv, *bv)
        return yr.reshape(*lead, 1), yv.reshape(*lead, 1)

    @staticmethod
    def backward(ctx, d_reward, d_value):
        saved = list(ctx.saved_tensors)
        a1, a2 = saved[0], saved[1]
        n = ctx.n_layers
        wr, br, wv, bv = saved[2:2 + n], saved[2 + n:2 + 2 * n], saved[2 + 2 * n:2 + 3 * n], saved[2 + 3 * n:]
        need = ctx.needs_input_grad     # (act_id, n_layers, x1, x2, reward w0, b0, ..., value w0, b0, ...)
        rows = a1.shape[0]
        if (not any(need[4:]) and (need[2] or need[3]) and d_reward is not None and d_value is not None
                and ctx.saved_r is not None and os.environ.get("BD_HEADS_PAIR_BWD", "1") != "0"):
            # frozen heads (the behaviour step, src/dreamer.py:320): both chains and the sum of their input
            # gradients in ONE launch
            lib = _lib.load()
            mr, mv = _lib.make_mlp(wr, br, ctx.act_id), _lib.make_mlp(wv, bv, ctx.act_id)
            k1, k2 = a1.shape[1], a2.shape[1]
            dx1 = torch.empty_like(a1) if need[2] else None
            dx2 = torch.empty_like(a2) if need[3] else None
            ws = _lib.workspace(lib.bd_heads_backward_workspace_bytes(C.byref(mr), C.byref(mv), k1, k2), a1.device)
            dr, dv = _f32c(d_reward).reshape(rows), _f32c(d_value).reshape(rows)
            _lib.check(lib.bd_heads_backward(C.byref(mr), C.byref(mv), k1, k2, rows, _lib.ptr(dr), _lib.ptr(dv),
                                             ctx.saved_r.data_ptr(), ctx.saved_v.data_ptr(), _lib.ptr(dx1),
                                             _lib.ptr(dx2), ws.data_ptr(), ws.numel(), ctx.prec,
                                             _lib.stream_ptr()), "bd_heads_backward")
            return (None, None, dx1.reshape(ctx.x1_shape) if dx1 is not None else None,
                    dx2.reshape(ctx.x2_shape) if dx2 is not None else None) + (None,) * (4 * n)
        outs = []
        for k, (ws_, bs_, dy, sv) in enumerate(((wr, br, d_reward, ctx.saved_r), (wv, bv, d_value, ctx.saved_v))):
            base = 4 + 2 * n * k
            if dy is None:
                dy = torch.zeros(rows, 1, device=a1.device, dtype=torch.float32)
            outs.append(_mlp_backward_call(ctx.act_id, ctx.prec, a1, a2, ws_, bs_, _f32c(dy).reshape(rows, 1), sv,
                                           need[2], need[3], [need[base + 2 * i] for i in range(n)],
                                           [need[base + 2 * i + 1] for i in range(n)]))
        (dx1r, dx2r, dwr, dbr), (dx1v, dx2v, dwv, dbv) = outs
        dx1 = (dx1r + dx1v).reshape(ctx.x1_shape) if need[2] else None
        dx2 = (dx2r + dx2v).reshape(ctx.x2_shape) if need[3] else None
        grads: List[Optional[torch.Tensor]] = [None, None, dx1, dx2]
        for dws, dbs in ((dwr, dbr), (dwv, dbv)):
            for i in range(n):
                grads += [dws[i], dbs[i]]
        return tuple(grads)


def heads_pair_supported(act_id: int, x1, x2, wr, br, wv, bv) -> bool:
    if _precision == "fp32" or not (x1.is_cuda and x2 is not None) or len(wr) != len(wv):
        return False
    lib = _lib.load()
    det = lambda ps: [_f32c(p.detach()) for p in ps]
    mr, mv = _lib.make_mlp(det(wr), det(br), act_id), _lib.make_mlp(det(wv), det(bv), act_id)
    return bool(lib.bd_heads_forward_supported(C.byref(mr), C.byref(mv), x1.shape[-1], x2.shape[-1], _prec()))


def mlp_apply(act_id: int, x1, x2, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]):
    params = []
    for w, b in zip(weights, biases):
        params += [w, b]
    return MlpFunction.apply(act_id, x1, x2, *params)


# =============================================================================
# lambda_return (src/dreamer.py:447-471)
# =============================================================================
class LambdaReturnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, reward, value, bootstrap, discount: float, lambda_: float):
        lib = _lib.load()
        shape = reward.shape
        T = shape[0]
        r, v, b = _f32c(reward).reshape(T, -1), _f32c(value).reshape(T, -1), _f32c(bootstrap).reshape(-1)
        N = r.shape[1]
        if v.shape != r.shape or b.shape[0] != N:
            raise BdError("lambda_return: reward/value/bootstrap shapes disagree")
        out = torch.empty_like(r)
        _lib.check(lib.bd_lambda_return_forward(_lib.ptr(r), _lib.ptr(v), _lib.ptr(b), T, N,
                                                float(discount), float(lambda_), _lib.ptr(out),
                                                _lib.stream_ptr()), "bd_lambda_return_forward")
        ctx.dims = (T, N, float(discount), float(lambda_), shape, bootstrap.shape)
        return out.reshape(shape)

    @staticmethod
    def backward(ctx, d_ret):
        lib = _lib.load()
        T, N, disc, lam, shape, bshape = ctx.dims
        g = _f32c(d_ret).reshape(T, N)
        need = ctx.needs_input_grad
        dr = torch.empty_like(g) if need[0] else None
        dv = torch.empty_like(g) if need[1] else None
        db = torch.empty(N, device=g.device, dtype=torch.float32) if need[2] else None
        _lib.check(lib.bd_lambda_return_backward(_lib.ptr(g), T, N, disc, lam, _lib.ptr(dr),
                                                 _lib.ptr(dv), _lib.ptr(db), _lib.stream_ptr()),
                   "bd_lambda_return_backward")
        return (dr.reshape(shape) if dr is not None else None,
                dv.reshape(shape) if dv is not None else None,
                db.reshape(bshape) if db is not None else None, None, None)


# =============================================================================
# TransitionModel.forward (src/models.py:190-299)
# =============================================================================
RSSM_PARAM_NAMES = ("embed_w", "embed_b", "w_ih", "w_hh", "b_ih", "b_hh", "prior1_w", "prior1_b",
                    "prior2_w", "prior2_b", "post1_w", "post1_b", "post2_w", "post2_b")


def make_rssm(p: Sequence[torch.Tensor], dims: dict) -> _lib.Rssm:
    """p: the 14 tensors in RSSM_PARAM_NAMES order (posterior entries may be None)."""
    r = _lib.Rssm()
    r.belief_size, r.state_size, r.action_size = dims["Be"], dims["S"], dims["A"]
    r.hidden_size, r.embedding_size = dims["Hi"], dims["E"]
    r.activation, r.min_std_dev = dims["act_id"], dims["min_std"]
    r.embed = _lib.make_linear(p[0], p[1])
    r.w_ih, r.w_hh, r.b_ih, r.b_hh = (_lib.ptr(t) for t in p[2:6])
    r.prior1, r.prior2 = _lib.make_linear(p[6], p[7]), _lib.make_linear(p[8], p[9])
    if p[10] is not None:
        r.post1, r.post2 = _lib.make_linear(p[10], p[11]), _lib.make_linear(p[12], p[13])
    return r


class TransitionFunction(torch.autograd.Function):
    """inputs: dims, init_state, actions, init_belief, embeddings|None, nonterminals|None,
    eps_prior, eps_post|None, *14 rssm params."""

    @staticmethod
    def forward(ctx, dims, init_state, actions, init_belief, embeddings, nonterminals, eps_prior,
                eps_post, *params):
        ctx.set_materialize_grads(False)     # outputs the loss does not use arrive as None, not as zero tensors
        lib = _lib.load()
        observe = embeddings is not None
        P = [(_f32c(p) if p is not None else None) for p in params]
        s0, b0, act = _f32c(init_state), _f32c(init_belief), _f32c(actions)
        L, B = act.shape[0], act.shape[1]
        Be, S = dims["Be"], dims["S"]
        if s0.shape != (B, S) or b0.shape != (B, Be) or act.shape[2] != dims["A"]:
            raise BdError(f"TransitionModel: bad input shapes {tuple(s0.shape)}, {tuple(act.shape)}, "
                          f"{tuple(b0.shape)}")
        emb = _f32c(embeddings) if observe else None
        nt = _f32c(nonterminals) if nonterminals is not None else None
        ep = _f32c(eps_prior)
        eq = _f32c(eps_post) if observe else None
        dev = s0.device
        new = lambda d: torch.empty(L, B, d, device=dev, dtype=torch.float32)
        outs = [new(Be), new(S), new(S), new(S)] + ([new(S), new(S), new(S)] if observe else [])
        a = _lib.TransitionArgs()
        a.rssm = make_rssm(P, dims)
        a.L, a.B = L, B
        a.init_state, a.init_belief, a.actions = _lib.ptr(s0), _lib.ptr(b0), _lib.ptr(act)
        a.embeddings, a.nonterminals = _lib.ptr(emb), _lib.ptr(nt)
        a.eps_prior, a.eps_post = _lib.ptr(ep), _lib.ptr(eq)
        a.beliefs, a.prior_states, a.prior_means, a.prior_stds = (_lib.ptr(t) for t in outs[:4])
        if observe:
            a.post_states, a.post_means, a.post_stds = (_lib.ptr(t) for t in outs[4:])
        nbytes = lib.bd_transition_workspace_bytes(C.byref(a.rssm), L, B, int(observe), 0)
        ws = _lib.workspace(nbytes, dev)
        _lib.check(lib.bd_transition_forward(C.byref(a), ws.data_ptr(), ws.numel(), _prec(),
                                             _lib.stream_ptr()), "bd_transition_forward")
        ctx.dims, ctx.observe, ctx.has_nt = dims, observe, nt is not None
        ctx.prec = _prec()
        ctx.n_params = len(P)
        keep = [s0, b0, act, ep] + ([emb, eq] if observe else []) + ([nt] if nt is not None else [])
        ctx.save_for_backward(*keep, *outs, *[p for p in P if p is not None])
        ctx.param_present = [p is not None for p in P]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        s0, b0, act, ep = saved[:4]
        del saved[:4]
        emb = eq = nt = None
        if ctx.observe:
            emb, eq = saved[:2]
            del saved[:2]
        if ctx.has_nt:
            nt = saved.pop(0)
        n_out = 7 if ctx.observe else 4
        outs = saved[:n_out]
        del saved[:n_out]
        P, it = [], iter(saved)
        for present in ctx.param_present:
            P.append(next(it) if present else None)
        dims = ctx.dims
        L, B = act.shape[0], act.shape[1]
        g = [(_f32c(t) if t is not None else None) for t in gouts]
        need = ctx.needs_input_grad   # dims, s0, actions, b0, emb, nt, eps_p, eps_q, *params
        a = _lib.TransitionBwdArgs()
        f = a.fwd
        f.rssm = make_rssm(P, dims)
        f.L, f.B = L, B
        f.init_state, f.init_belief, f.actions = _lib.ptr(s0), _lib.ptr(b0), _lib.ptr(act)
        f.embeddings, f.nonterminals = _lib.ptr(emb), _lib.ptr(nt)
        f.eps_prior, f.eps_post = _lib.ptr(ep), _lib.ptr(eq)
        f.beliefs, f.prior_states, f.prior_means, f.prior_stds = (_lib.ptr(t) for t in outs[:4])
        if ctx.observe:
            f.post_states, f.post_means, f.post_stds = (_lib.ptr(t) for t in outs[4:])
        a.g_beliefs, a.g_prior_states, a.g_prior_means, a.g_prior_stds = (_lib.ptr(t) for t in g[:4])
        if ctx.observe:
            a.g_post_states, a.g_post_means, a.g_post_stds = (_lib.ptr(t) for t in g[4:7])
        d_s0 = torch.empty_like(s0) if need[1] else None
        d_act = torch.empty_like(act) if need[2] else None
        d_b0 = torch.empty_like(b0) if need[3] else None
        d_emb = torch.empty_like(emb) if (ctx.observe and need[4]) else None
        a.d_init_state, a.d_init_belief = _lib.ptr(d_s0), _lib.ptr(d_b0)
        a.d_actions, a.d_embeddings = _lib.ptr(d_act), _lib.ptr(d_emb)
        dP = []
        for i, name in enumerate(RSSM_PARAM_NAMES):
            gp = _zeros_like_if(P[i] is not None and need[8 + i], P[i]) if P[i] is not None else None
            dP.append(gp)
            setattr(a.grads, name, _lib.ptr(gp))
        nbytes = lib.bd_transition_workspace_bytes(C.byref(f.rssm), L, B, int(ctx.observe), 1)
        ws = _lib.workspace(nbytes, s0.device)
        _lib.check(lib.bd_transition_backward(C.byref(a), ws.data_ptr(), ws.numel(), ctx.prec,
                                              _lib.stream_ptr()), "bd_transition_backward")
        return (None, d_s0, d_act, d_b0, d_emb, None, None, None, *dP)


# =============================================================================
# Dreamer.imagine_ahead (src/dreamer.py:178-237)
# =============================================================================
class ImagineFunction(torch.autograd.Function):
    """inputs: dims, actor_cfg(dict), T, prev_state (N,S), prev_belief (N,Be), eps_a, eps_e, eps_s,
    n_actor_layers, *actor params (w0,b0,...), *10 prior-path rssm params."""

    @staticmethod
    def forward(ctx, dims, actor_cfg, T, prev_state, prev_belief, eps_a, eps_e, eps_s, n_actor,
                *params):
        ctx.set_materialize_grads(False)     # outputs the loss does not use arrive as None, not as zero tensors
        lib = _lib.load()
        AP = [_f32c(p) for p in params[:2 * n_actor]]
        RP = [_f32c(p) for p in params[2 * n_actor:]] + [None] * 4
        s0, b0 = _f32c(prev_state), _f32c(prev_belief)
        N = s0.shape[0]
        Be, S, A, J = dims["Be"], dims["S"], dims["A"], actor_cfg["entropy_samples"]
        ea, ee, es = _f32c(eps_a), _f32c(eps_e), _f32c(eps_s)
        if ea.shape != (T, N, A) or ee.shape != (T, J, N, A) or es.shape != (T, N, S):
            raise BdError(f"imagine_ahead: noise shapes {tuple(ea.shape)}, {tuple(ee.shape)}, "
                          f"{tuple(es.shape)} do not match T={T}, N={N}, A={A}, S={S}, J={J}")
        dev = s0.device
        new = lambda *d: torch.empty(*d, device=dev, dtype=torch.float32)
        beliefs, states, means, stds = new(T, N, Be), new(T, N, S), new(T, N, S), new(T, N, S)
        entropy, actions = new(T, N), new(T, N, A)
        actor_raw, dent = new(T, N, 2 * A), new(T, N, 2 * A)
        a = _lib.ImagineArgs()
        a.rssm = make_rssm(RP, dims)
        a.actor = _lib.make_mlp(AP[0::2], AP[1::2], actor_cfg.get("act_id", dims["act_id"]))
        a.actor_cfg = _lib.ActorCfg(actor_cfg["mean_scale"], actor_cfg["raw_init_std"],
                                    actor_cfg["min_std"], J)
        a.T, a.N = T, N
        a.prev_state, a.prev_belief = _lib.ptr(s0), _lib.ptr(b0)
        a.eps_a, a.eps_e, a.eps_s = _lib.ptr(ea), _lib.ptr(ee), _lib.ptr(es)
        a.beliefs, a.states, a.means, a.stds = (_lib.ptr(t) for t in (beliefs, states, means, stds))
        a.entropy, a.actions = _lib.ptr(entropy), _lib.ptr(actions)
        a.actor_raw, a.dent = _lib.ptr(actor_raw), _lib.ptr(dent)
        saved = None
        need_bwd = any(ctx.needs_input_grad)   # (grad mode is always off inside Function.forward)
        nsaved = lib.bd_imagine_saved_bytes(C.byref(a.rssm), T, N, _prec()) if need_bwd else 0
        if nsaved:
            saved = torch.empty(nsaved, dtype=torch.uint8, device=dev)
            a.tc_saved = saved.data_ptr()
        ctx.tc_saved = saved
        ctx.prec = _prec()
        nbytes = lib.bd_imagine_workspace_bytes(C.byref(a.rssm), C.byref(a.actor), T, N, 0)
        ws = _lib.workspace(nbytes, dev)
        _lib.check(lib.bd_imagine_forward(C.byref(a), ws.data_ptr(), ws.numel(), _prec(),
                                          _lib.stream_ptr()), "bd_imagine_forward")
        ctx.dims, ctx.actor_cfg, ctx.T, ctx.n_actor = dims, actor_cfg, T, n_actor
        ctx.save_for_backward(s0, b0, ea, ee, es, beliefs, states, means, stds, entropy, actions,
                              actor_raw, dent, *AP, *RP[:10])
        ctx.mark_non_differentiable(actions)
        return beliefs, states, means, stds, entropy, actions

    @staticmethod
    def backward(ctx, g_b, g_s, g_m, g_sd, g_ent, _g_actions):
        lib = _lib.load()
        sv = list(ctx.saved_tensors)
        (s0, b0, ea, ee, es, beliefs, states, means, stds, entropy, actions, actor_raw,
         dent) = sv[:13]
        n_actor = ctx.n_actor
        AP = sv[13:13 + 2 * n_actor]
        RP = sv[13 + 2 * n_actor:] + [None] * 4
        dims, cfg, T = ctx.dims, ctx.actor_cfg, ctx.T
        N = s0.shape[0]
        a = _lib.ImagineBwdArgs()
        f = a.fwd
        f.rssm = make_rssm(RP, dims)
        f.actor = _lib.make_mlp(AP[0::2], AP[1::2], cfg.get("act_id", dims["act_id"]))
        f.actor_cfg = _lib.ActorCfg(cfg["mean_scale"], cfg["raw_init_std"], cfg["min_std"],
                                    cfg["entropy_samples"])
        f.T, f.N = T, N
        f.prev_state, f.prev_belief = _lib.ptr(s0), _lib.ptr(b0)
        f.eps_a, f.eps_e, f.eps_s = _lib.ptr(ea), _lib.ptr(ee), _lib.ptr(es)
        f.beliefs, f.states, f.means, f.stds = (_lib.ptr(t) for t in (beliefs, states, means, stds))
        f.entropy, f.actions = _lib.ptr(entropy), _lib.ptr(actions)
        f.actor_raw, f.dent = _lib.ptr(actor_raw), _lib.ptr(dent)
        if ctx.tc_saved is not None:
            f.tc_saved = ctx.tc_saved.data_ptr()
        gc = lambda t: _f32c(t) if t is not None else None
        g_b, g_s, g_m, g_sd, g_ent = gc(g_b), gc(g_s), gc(g_m), gc(g_sd), gc(g_ent)
        a.g_beliefs, a.g_states, a.g_means = _lib.ptr(g_b), _lib.ptr(g_s), _lib.ptr(g_m)
        a.g_stds, a.g_entropy = _lib.ptr(g_sd), _lib.ptr(g_ent)
        need = ctx.needs_input_grad  # dims,cfg,T,s0,b0,ea,ee,es,n_actor,*params
        d_s0 = torch.empty_like(s0) if need[3] else None
        d_b0 = torch.empty_like(b0) if need[4] else None
        a.d_prev_state, a.d_prev_belief = _lib.ptr(d_s0), _lib.ptr(d_b0)
        dA = _zero_grads([need[9 + j] for j in range(2 * n_actor)], list(AP[:2 * n_actor]))
        for i in range(n_actor):
            a.actor_dw[i], a.actor_db[i] = _lib.ptr(dA[2 * i]), _lib.ptr(dA[2 * i + 1])
        nbytes = lib.bd_imagine_workspace_bytes(C.byref(f.rssm), C.byref(f.actor), T, N, 1)
        ws = _lib.workspace(nbytes, s0.device)
        _lib.check(lib.bd_imagine_backward(C.byref(a), ws.data_ptr(), ws.numel(), ctx.prec,
                                           _lib.stream_ptr()), "bd_imagine_backward")
        return (None, None, None, d_s0, d_b0, None, None, None, None, *dA, *([None] * 10))


# =============================================================================
# imagine_ahead + reward / value heads + lambda_return, fused (src/dreamer.py:313-335)
# =============================================================================
def _fill_imagine_args(a, dims, actor_cfg, T, N, s0, b0, ea, ee, es, AP, RP, outs):
    a.rssm = make_rssm(RP, dims)
    a.actor = _lib.make_mlp(AP[0::2], AP[1::2], actor_cfg.get("act_id", dims["act_id"]))
    a.actor_cfg = _lib.ActorCfg(actor_cfg["mean_scale"], actor_cfg["raw_init_std"],
                                actor_cfg["min_std"], actor_cfg["entropy_samples"])
    a.T, a.N = T, N
    a.prev_state, a.prev_belief = _lib.ptr(s0), _lib.ptr(b0)
    a.eps_a, a.eps_e, a.eps_s = _lib.ptr(ea), _lib.ptr(ee), _lib.ptr(es)
    (a.beliefs, a.states, a.means, a.stds, a.entropy, a.actions, a.actor_raw,
     a.dent) = (_lib.ptr(t) for t in outs)


class ImagineReturnsFunction(torch.autograd.Function):
    """inputs: dims, actor_cfg, head_act_id, T, discount, lambda_, prev_state (N,S), prev_belief (N,Be),
    eps_a, eps_e, eps_s, n_actor, n_head, *actor params, *10 prior-path rssm params, *reward params,
    *value params.  One forward call (bd_imagine_returns_forward) and one backward call
    (bd_imagine_returns_backward).  Only the actor receives parameter gradients: transition and head
    weights are constants, as under the reference's FreezeParameters blocks (src/dreamer.py:313,320)."""

    @staticmethod
    def forward(ctx, dims, actor_cfg, head_act, T, discount, lambda_, prev_state, prev_belief, eps_a,
                eps_e, eps_s, n_actor, n_head, *params):
        ctx.set_materialize_grads(False)     # outputs the loss does not use arrive as None, not as zero tensors
        lib = _lib.load()
        AP = [_f32c(p) for p in params[:2 * n_actor]]
        RP = [_f32c(p) for p in params[2 * n_actor:2 * n_actor + 10]] + [None] * 4
        HP = [_f32c(p) for p in params[2 * n_actor + 10:]]
        RW, VW = HP[:2 * n_head], HP[2 * n_head:]
        s0, b0 = _f32c(prev_state), _f32c(prev_belief)
        N = s0.shape[0]
        Be, S, A, J = dims["Be"], dims["S"], dims["A"], actor_cfg["entropy_samples"]
        ea, ee, es = _f32c(eps_a), _f32c(eps_e), _f32c(eps_s)
        if ea.shape != (T, N, A) or ee.shape != (T, J, N, A) or es.shape != (T, N, S):
            raise BdError("imagine_and_returns: noise shapes do not match T, N, A, S, J")
        dev = s0.device
        new = lambda *d: torch.empty(*d, device=dev, dtype=torch.float32)
        beliefs, states, means, stds = new(T, N, Be), new(T, N, S), new(T, N, S), new(T, N, S)
        entropy, actions = new(T, N), new(T, N, A)
        actor_raw, dent = new(T, N, 2 * A), new(T, N, 2 * A)
        reward, value, returns = new(T, N, 1), new(T, N, 1), new(T, N, 1)
        a = _lib.ImagineReturnsArgs()
        outs = (beliefs, states, means, stds, entropy, actions, actor_raw, dent)
        _fill_imagine_args(a.img, dims, actor_cfg, T, N, s0, b0, ea, ee, es, AP, RP, outs)
        a.reward = _lib.make_mlp(RW[0::2], RW[1::2], head_act)
        a.value = _lib.make_mlp(VW[0::2], VW[1::2], head_act)
        a.discount, a.lambda_ = float(discount), float(lambda_)
        a.reward_out, a.value_out, a.returns = _lib.ptr(reward), _lib.ptr(value), _lib.ptr(returns)
        need_bwd = any(ctx.needs_input_grad)
        prec = _prec()
        saved = hsaved = None
        if need_bwd:
            saved = torch.empty(lib.bd_imagine_saved_bytes(C.byref(a.img.rssm), T, N, prec), dtype=torch.uint8,
                                device=dev)
            a.img.tc_saved = saved.data_ptr()
            hsaved = torch.empty(lib.bd_imagine_returns_saved_bytes(C.byref(a)), dtype=torch.uint8, device=dev)
            a.heads_saved = hsaved.data_ptr()
        ws = _lib.workspace(lib.bd_imagine_returns_workspace_bytes(C.byref(a), 0), dev)
        _lib.check(lib.bd_imagine_returns_forward(C.byref(a), ws.data_ptr(), ws.numel(), prec,
                                                  _lib.stream_ptr()), "bd_imagine_returns_forward")
        ctx.cfg = (dims, actor_cfg, head_act, T, float(discount), float(lambda_), n_actor, n_head, prec)
        ctx.bufs = (saved, hsaved)
        ctx.save_for_backward(s0, b0, ea, ee, es, beliefs, states, means, stds, entropy, actions, actor_raw,
                              dent, reward, value, returns, *AP, *RP[:10], *RW, *VW)
        ctx.mark_non_differentiable(actions)
        return beliefs, states, means, stds, entropy, actions, reward, value, returns

    @staticmethod
    def backward(ctx, g_b, g_s, g_m, g_sd, g_ent, _g_act, g_rew, g_val, g_ret):
        lib = _lib.load()
        dims, actor_cfg, head_act, T, discount, lambda_, n_actor, n_head, prec = ctx.cfg
        sv = list(ctx.saved_tensors)
        (s0, b0, ea, ee, es, beliefs, states, means, stds, entropy, actions, actor_raw, dent, reward, value,
         returns) = sv[:16]
        AP = sv[16:16 + 2 * n_actor]
        RP = sv[16 + 2 * n_actor:26 + 2 * n_actor] + [None] * 4
        HP = sv[26 + 2 * n_actor:]
        RW, VW = HP[:2 * n_head], HP[2 * n_head:]
        N = s0.shape[0]
        a = _lib.ImagineReturnsBwdArgs()
        f = a.fwd
        outs = (beliefs, states, means, stds, entropy, actions, actor_raw, dent)
        _fill_imagine_args(f.img, dims, actor_cfg, T, N, s0, b0, ea, ee, es, AP, RP, outs)
        f.reward = _lib.make_mlp(RW[0::2], RW[1::2], head_act)
        f.value = _lib.make_mlp(VW[0::2], VW[1::2], head_act)
        f.discount, f.lambda_ = discount, lambda_
        f.reward_out, f.value_out, f.returns = _lib.ptr(reward), _lib.ptr(value), _lib.ptr(returns)
        saved, hsaved = ctx.bufs
        f.img.tc_saved, f.heads_saved = saved.data_ptr(), hsaved.data_ptr()
        gc = lambda t: _f32c(t) if t is not None else None
        (a.g_beliefs, a.g_states, a.g_means, a.g_stds, a.g_entropy, a.g_reward, a.g_value,
         a.g_returns) = (_lib.ptr(gc(t)) for t in (g_b, g_s, g_m, g_sd, g_ent, g_rew, g_val, g_ret))
        need = ctx.needs_input_grad   # dims, cfg, head_act, T, disc, lam, s0, b0, ea, ee, es, n_actor, n_head, *params
        d_s0 = torch.empty_like(s0) if need[6] else None
        d_b0 = torch.empty_like(b0) if need[7] else None
        a.d_prev_state, a.d_prev_belief = _lib.ptr(d_s0), _lib.ptr(d_b0)
        dA = _zero_grads([need[13 + j] for j in range(2 * n_actor)], list(AP))
        for i in range(n_actor):
            a.actor_dw[i], a.actor_db[i] = _lib.ptr(dA[2 * i]), _lib.ptr(dA[2 * i + 1])
        ws = _lib.workspace(lib.bd_imagine_returns_workspace_bytes(C.byref(f), 1), s0.device)
        _lib.check(lib.bd_imagine_returns_backward(C.byref(a), ws.data_ptr(), ws.numel(), prec,
                                                   _lib.stream_ptr()), "bd_imagine_returns_backward")
        n_rest = 10 + 4 * n_head
        return (None,) * 6 + (d_s0, d_b0, None, None, None, None, None, *dA, *([None] * n_rest))


# =============================================================================
# KL loss (Planet._kl_loss src/planet.py:288-308, Dreamer._kl_loss src/dreamer.py:111-146)
# =============================================================================
class KlLossFunction(torch.autograd.Function):
    """(post_mean, post_std, prior_mean, prior_std) each (L,B,S), free_nats (1,) device tensor,
    balance (-1 = off) -> scalar loss.  Two kernels forward, one backward."""

    @staticmethod
    def forward(ctx, post_mean, post_std, prior_mean, prior_std, free_nats, balance: float):
        lib = _lib.load()
        mq, sq, mp, sp = (_f32c(t) for t in (post_mean, post_std, prior_mean, prior_std))
        if not (mq.shape == sq.shape == mp.shape == sp.shape) or mq.dim() < 2:
            raise BdError("kl_loss: parameter tensors must share one (..., S) shape")
        fn = _f32c(free_nats).reshape(-1)[:1]
        S = mq.shape[-1]
        rows = mq.numel() // S
        div = torch.empty(rows, device=mq.device, dtype=torch.float32)
        loss = torch.empty(2, device=mq.device, dtype=torch.float32)
        _lib.check(lib.bd_kl_loss_forward(_lib.ptr(mq), _lib.ptr(sq), _lib.ptr(mp), _lib.ptr(sp), rows, S,
                                          _lib.ptr(fn), float(balance), _lib.ptr(div), _lib.ptr(loss),
                                          _lib.stream_ptr()), "bd_kl_loss_forward")
        ctx.save_for_backward(mq, sq, mp, sp, fn, div, loss)
        ctx.balance, ctx.shape = float(balance), post_mean.shape
        # the reference returns a 0-dim tensor without balancing and shape (1,) with it
        return loss[0].clone() if balance < 0 else loss[:1].clone()

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        mq, sq, mp, sp, fn, div, loss = ctx.saved_tensors
        S = mq.shape[-1]
        rows = mq.numel() // S
        need = ctx.needs_input_grad
        outs = [torch.empty_like(t) if n else None for t, n in zip((mq, sq, mp, sp), need[:4])]
        gl = _f32c(g).reshape(-1)[:1]
        _lib.check(lib.bd_kl_loss_backward(_lib.ptr(mq), _lib.ptr(sq), _lib.ptr(mp), _lib.ptr(sp), rows, S,
                                           _lib.ptr(fn), ctx.balance, _lib.ptr(div), _lib.ptr(loss),
                                           _lib.ptr(gl), *(_lib.ptr(o) for o in outs), _lib.stream_ptr()),
                   "bd_kl_loss_backward")
        return (*(o.reshape(ctx.shape) if o is not None else None for o in outs), None, None)
