"""Host-side mirror of the reference's operator interface for the RSSM hot path.

Same constructor signatures, ``forward`` signatures, tensor layouts and
``state_dict`` keys as the reference classes they replace, so ``planet.py`` /
``dreamer.py`` / ``planner.py`` can use them as drop-ins:

    TransitionModel   <- src/models.py:120-299
    DenseModel        <- src/models.py:365-408
    MPCPlanner        <- src/planner.py:5-90
    imagine_ahead     <- Dreamer.imagine_ahead, src/dreamer.py:178-237
    lambda_return     <- src/dreamer.py:447-471

All arithmetic runs in libbd_b200.so on the current CUDA device.  CPU tensors,
Categorical latents and unknown activations raise -- there is no fallback.
Every function takes an optional ``noise=`` argument carrying the Gaussian draws
the reference makes internally, for bit-for-bit comparable parity runs.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional, Tuple

import torch
from torch import nn, Tensor

from . import _lib
from . import functions as F_
from ._lib import BdError

ENTROPY_SAMPLES = 100  # SampleDist(samples=100), src/models.py:684

# patch(fused=True): the last fused imagine_and_returns call made on behalf of an UNMODIFIED
# Dreamer.train_step (see imagine_ahead_fused): the heads / lambda_return calls that follow it in
# src/dreamer.py:321-335 recognise their inputs by identity and hand back the fused results.
_fused_record = None


def build_mlp(input_size: int, hidden_size: int, output_size: int, n_layers: int,
              activation="ELU", output_activation="Identity") -> nn.Sequential:
    """Same layer structure (and therefore the same parameter names
    ``{0,2,4,...}.{weight,bias}``) as the reference's build_mlp, src/utils.py:368-404."""
    act = getattr(nn, activation) if isinstance(activation, str) else activation
    out_act = getattr(nn, output_activation) if isinstance(output_activation, str) \
        else output_activation
    layers, in_size = [], input_size
    for _ in range(n_layers):
        layers += [nn.Linear(in_size, hidden_size), act()]
        in_size = hidden_size
    layers += [nn.Linear(in_size, output_size), out_act()]
    return nn.Sequential(*layers)


def _act_name(activation) -> str:
    if isinstance(activation, str):
        return activation
    return getattr(activation, "__name__", type(activation).__name__)


def _linears(seq: nn.Sequential):
    return [m for m in seq if isinstance(m, nn.Linear)]


# =============================================================================
# DenseModel
# =============================================================================
class DenseModel(nn.Module):
    """General fused MLP: ``forward(belief, state)`` (concatenated on the last dim) or
    ``forward(x)``; any leading dims.  Reference: src/models.py:365-408."""

    def __init__(self, input_size: int, hidden_size: int, output_size: int = 1,
                 activation: str = "ELU", n_layers: int = 4, distribution: str = "normal") -> None:
        super().__init__()
        self.model = build_mlp(input_size, hidden_size, output_size, n_layers, activation)
        self.distribution = distribution
        self._act_id = _lib.activation_id(_act_name(activation))

    def forward(self, *args: Tensor) -> Tensor:
        if len(args) == 2:
            x1, x2 = args
            rec = _fused_record
            if rec is not None and x1 is rec["beliefs"] and x2 is rec["states"]:
                # patch(fused=True): this head already ran inside the fused rollout on exactly these tensors
                if self is rec["reward_model"]:
                    return rec["reward"]
                if self is rec["value_model"]:
                    return rec["value"]
        else:
            (x1,), x2 = args, None
        lin = _linears(self.model)
        return F_.mlp_apply(self._act_id, x1, x2, [l.weight for l in lin], [l.bias for l in lin])


# =============================================================================
# TransitionModel
# =============================================================================
class GaussianBeliefModel(nn.Module):
    """Parameter container with the reference's layout (src/models.py:44-73).  ``forward`` keeps
    the reference behaviour for code that drives the RSSM step by step from Python."""

    def __init__(self, input_size, hidden_size, state_size, activation, min_std_dev) -> None:
        super().__init__()
        self.min_std_dev = min_std_dev
        self.model = build_mlp(input_size, hidden_size, 2 * state_size, 1, activation)

    def forward(self, belief: Tensor):
        mean, raw = torch.chunk(self.model(belief), 2, dim=1)
        std = torch.nn.functional.softplus(raw) + self.min_std_dev
        return mean + std * torch.randn_like(mean), (mean, std)


def rssm_params(tm, with_posterior: bool):
    """The 14 tensors of functions.RSSM_PARAM_NAMES read off a TransitionModel-shaped module
    (ours or the reference's own: same attribute names, src/models.py:149-167)."""
    emb = _linears(tm.fc_embed_state_action)[0]
    p1, p2 = _linears(tm.belief_prior.model)
    out = [emb.weight, emb.bias, tm.rnn.weight_ih, tm.rnn.weight_hh, tm.rnn.bias_ih,
           tm.rnn.bias_hh, p1.weight, p1.bias, p2.weight, p2.bias]
    if with_posterior:
        q1, q2 = _linears(tm.belief_posterior.model)
        out += [q1.weight, q1.bias, q2.weight, q2.bias]
    else:
        out += [None] * 4
    return out


def rssm_dims(tm) -> Dict:
    emb = _linears(tm.fc_embed_state_action)[0]
    p1, p2 = _linears(tm.belief_prior.model)
    q1 = _linears(tm.belief_posterior.model)[0]
    Be, S = emb.out_features, p2.out_features // 2
    act = None
    for m in tm.fc_embed_state_action:
        if not isinstance(m, nn.Linear):
            act = type(m).__name__
            break
    return dict(Be=Be, S=S, A=emb.in_features - S, Hi=p1.out_features, E=q1.in_features - Be,
                act_id=_lib.activation_id(act or "Identity"), min_std=float(tm.min_std_dev))


class TransitionModel(nn.Module):
    """RSSM transition model; reference: src/models.py:120-299."""

    def __init__(self, belief_size: int, state_size: int, action_size: int, hidden_size: int,
                 embedding_size: int, activation: Optional[str] = "ELU", min_std_dev: float = 0.1,
                 latent_distribution: Optional[str] = "Gaussian",
                 discrete_latent_dimensions: Optional[int] = 32,
                 discrete_latent_classes: Optional[int] = 32) -> None:
        super().__init__()
        assert latent_distribution in ["Gaussian", "Categorical"], f"{latent_distribution}"
        if latent_distribution != "Gaussian":
            raise NotImplementedError(
                "big_dreamer_b200.TransitionModel implements the Gaussian latent path only "
                "(latent_distribution='Categorical' stays on the reference module)")
        act = _act_name(activation)
        self._act_id = _lib.activation_id(act)
        self.min_std_dev = min_std_dev
        self.latent_distribution = latent_distribution
        self.rnn = nn.GRUCell(belief_size, belief_size)
        self.fc_embed_state_action = build_mlp(state_size + action_size, -1, belief_size, 0,
                                               output_activation=act)
        self.belief_prior = GaussianBeliefModel(belief_size, hidden_size, state_size, act,
                                                min_std_dev)
        self.belief_posterior = GaussianBeliefModel(belief_size + embedding_size, hidden_size,
                                                    state_size, act, min_std_dev)
        # the reference overwrites nn.Module.modules with this list (src/models.py:184-188);
        # FreezeParameters(self.transition_model.modules + ...) depends on it.
        self.modules = [self.fc_embed_state_action, self.belief_prior, self.belief_posterior]

    def forward(self, init_state: Tensor, actions: Tensor, init_belief: Tensor,
                embeddings: Optional[Tensor] = None, nonterminals: Optional[Tensor] = None,
                noise: Optional[Dict[str, Tensor]] = None):
        """init_state (B,S), actions (L,B,A), init_belief (B,Be), embeddings (L,B,E)|None,
        nonterminals (L,B,1)|None -> beliefs (L,B,Be), prior_states (L,B,S),
        (prior_means, prior_std_devs), posterior_states|None, (post_means, post_std_devs)|None.
        noise: {'eps_prior': (L,B,S), 'eps_post': (L,B,S)} (drawn with torch.randn if absent;
        the reference draws prior then posterior per step, src/models.py:256,267)."""
        observe = embeddings is not None
        L, B = actions.shape[0], actions.shape[1]
        dims = rssm_dims(self)
        S = dims["S"]
        if noise is None:
            eps = torch.randn(2 if observe else 1, L, B, S, device=actions.device,
                              dtype=torch.float32)
            eps_prior, eps_post = eps[0], (eps[1] if observe else None)
        else:
            eps_prior, eps_post = noise["eps_prior"], noise.get("eps_post")
            if observe and eps_post is None:
                raise BdError("TransitionModel: observe mode needs noise['eps_post']")
        params = rssm_params(self, with_posterior=observe)
        outs = F_.TransitionFunction.apply(dims, init_state, actions, init_belief, embeddings,
                                           nonterminals, eps_prior, eps_post, *params)
        if observe:
            return outs[0], outs[1], (outs[2], outs[3]), outs[4], (outs[5], outs[6])
        return outs[0], outs[1], (outs[2], outs[3]), None, None


# =============================================================================
# Dreamer.imagine_ahead / lambda_return
# =============================================================================
def _mlp_act_id(seq: nn.Sequential) -> int:
    """Activation id of a build_mlp chain: its first non-Linear module (Identity if none)."""
    for m in seq:
        if not isinstance(m, nn.Linear):
            return _lib.activation_id(type(m).__name__)
    return _lib.activation_id("Identity")


def actor_config(actor) -> Dict:
    """Squashing constants read from the reference's ActorModel (src/models.py:499-503) and the
    actor's OWN hidden activation (an independent constructor argument, src/models.py:469-481)."""
    raw = actor.raw_init_std
    raw = float(raw.item()) if isinstance(raw, torch.Tensor) else float(raw)
    return dict(mean_scale=float(actor._mean_scale), raw_init_std=raw,
                min_std=float(actor._min_std), entropy_samples=ENTROPY_SAMPLES,
                act_id=_mlp_act_id(actor.model))


def draw_imagine_noise(T: int, N: int, S: int, A: int, device, generator=None):
    """ε_a (T,N,A), ε_e (T,100,N,A), ε_s (T,N,S).  (The reference draws them per step in
    the order action, entropy, prior state: src/dreamer.py:443-444, src/models.py:72.)"""
    kw = dict(device=device, dtype=torch.float32, generator=generator)
    # ONE generator launch for the three tensors (views of one buffer; the entropy noise first: it is read with
    # 16-byte loads and stays 16-byte aligned there)
    ne, ns, na = T * ENTROPY_SAMPLES * N * A, T * N * S, T * N * A
    buf = torch.randn(ne + ns + na, **kw)
    return dict(eps_a=buf[ne + ns:].view(T, N, A), eps_e=buf[:ne].view(T, ENTROPY_SAMPLES, N, A),
                eps_s=buf[ne:ne + ns].view(T, N, S))


def imagine_ahead(self, prev_state: Tensor, prev_belief: Tensor,
                  noise: Optional[Dict[str, Tensor]] = None, return_actions: bool = False):
    """Drop-in for ``Dreamer.imagine_ahead`` (bind as a method; ``self`` needs
    ``transition_model``, ``actor``, ``planning_horizon``, ``latent_distribution``).

    prev_state (L,B,S), prev_belief (L,B,Be) -> beliefs (T,N,Be), prior_states (T,N,S),
    (prior_means, prior_std_devs), action_entropy (T,N); T = planning_horizon-1, N = L*B."""
    if getattr(self, "latent_distribution", "Gaussian") != "Gaussian":
        raise NotImplementedError("imagine_ahead: only Gaussian latents run on the B200 path")
    tm, actor = self.transition_model, self.actor
    if getattr(actor, "action_distribution", "Gaussian") != "Gaussian":
        raise NotImplementedError("imagine_ahead: only the Gaussian actor runs on the B200 path")
    T = self.planning_horizon - 1
    b0 = prev_belief.reshape(-1, prev_belief.shape[-1])
    s0 = prev_state.reshape(-1, prev_state.shape[-1])
    dims = rssm_dims(tm)
    N = s0.shape[0]
    rp = rssm_params(tm, with_posterior=False)[:10]
    if torch.is_grad_enabled() and any(p.requires_grad for p in rp):
        raise NotImplementedError(
            "imagine_ahead: transition-model parameters require grad; the B200 path computes "
            "actor gradients only -- call it under FreezeParameters(model_modules) as "
            "Dreamer.train_step does (src/dreamer.py:313)")
    if noise is None:
        noise = draw_imagine_noise(T, N, dims["S"], dims["A"], s0.device)
    lin = _linears(actor.model)
    ap = []
    for l in lin:
        ap += [l.weight, l.bias]
    beliefs, states, means, stds, entropy, actions = F_.ImagineFunction.apply(
        dims, actor_config(actor), T, s0, b0, noise["eps_a"], noise["eps_e"], noise["eps_s"],
        len(lin), *ap, *rp)
    if return_actions:
        return beliefs, states, (means, stds), entropy, actions
    return beliefs, states, (means, stds), entropy


def lambda_return(imged_reward: Tensor, value_pred: Tensor, bootstrap: Tensor,
                  discount: float = 0.99, lambda_: float = 0.95) -> Tensor:
    """Drop-in for src/dreamer.py:447-471.  (T,N,1),(T,N,1),(N,1) -> (T,N,1)."""
    rec = _fused_record
    if (rec is not None and imged_reward is rec["reward"] and value_pred is rec["value"]
            and float(discount) == rec["discount"] and float(lambda_) == rec["lambda_"]
            and bootstrap.shape == value_pred.shape[1:] and bootstrap.data_ptr() == value_pred[-1].data_ptr()):
        return rec["returns"]        # computed by the fused rollout's tail (bootstrap = value[-1])
    return F_.LambdaReturnFunction.apply(imged_reward, value_pred, bootstrap, discount, lambda_)


def kl_loss(posterior_params, prior_params, free_nats, kl_balance: float = -1) -> Tensor:
    """Fused KL(posterior || prior) loss with free nats and optional KL balancing
    (Planet._kl_loss src/planet.py:288-308; Dreamer._kl_loss src/dreamer.py:111-146).
    posterior_params / prior_params: (means, std_devs) each (L,B,S); free_nats: float or the
    reference's (1,) tensor.  Returns a 0-dim tensor (no balancing) or shape (1,) (balancing), as
    the reference does."""
    qm, qs = posterior_params
    pm, ps = prior_params
    if not isinstance(free_nats, torch.Tensor):
        free_nats = torch.full((1,), float(free_nats), device=qm.device, dtype=torch.float32)
    return F_.KlLossFunction.apply(qm, qs, pm, ps, free_nats, float(kl_balance))


def _kl_loss_method(self, posterior_params, prior_params):
    """Bound by patch() as Planet._kl_loss / Dreamer._kl_loss (Gaussian latents)."""
    return kl_loss(posterior_params, prior_params, self.free_nats, getattr(self, "kl_balance", -1))


def imagine_ahead_fused(self, prev_state: Tensor, prev_belief: Tensor,
                        noise: Optional[Dict[str, Tensor]] = None):
    """``Dreamer.imagine_ahead`` for ``patch(fused=True)``: runs the fused imagine + heads + lambda_return
    kernels with the agent's own ``reward_model`` / ``critic_target`` / ``discount`` / ``disclam`` -- what the
    unmodified ``train_step`` computes next (src/dreamer.py:320-335, heads under FreezeParameters) -- and
    remembers the results; the ``reward_model(b, s)``, ``critic_target(b, s)`` and ``lambda_return(...)`` calls
    that follow return them when handed exactly these tensors.  Agents without those attributes (or with
    heads the fused kernels do not cover) get the plain imagine_ahead."""
    global _fused_record
    _fused_record = None
    rm, vm = getattr(self, "reward_model", None), getattr(self, "critic_target", None)
    if rm is None or vm is None or not hasattr(self, "discount") or not hasattr(self, "disclam"):
        return imagine_ahead(self, prev_state, prev_belief, noise)
    beliefs, states, (means, stds), entropy, reward, value, returns = imagine_and_returns(
        self, prev_state, prev_belief, rm, vm, self.discount, self.disclam, noise, assume_frozen_heads=True,
        fused=True)
    _fused_record = dict(beliefs=beliefs, states=states, reward_model=rm, value_model=vm, reward=reward,
                         value=value, returns=returns, discount=float(self.discount), lambda_=float(self.disclam))
    return beliefs, states, (means, stds), entropy


def value_update(critic, beliefs: Tensor, states: Optional[Tensor], target: Tensor,
                 weight: Optional[Tensor] = None) -> Tensor:
    """The critic regression update of ``Dreamer.train_step`` (src/dreamer.py:369-391) behind one call:

        value_loss = -(weight * Normal(critic(beliefs, states), 1).log_prob(target)).mean()
        value_loss.backward()                 # gradients land in critic.parameters()[i].grad

    beliefs (..., Be), states (..., S) or None, target / weight (..., 1); inputs are treated as
    detached (the reference detaches them, :370-375).  Returns the loss (0-dim tensor); the caller runs
    ``clip_grad_norm_`` and the optimizer exactly as the reference does.  Parameter gradients of one
    call live in ONE flat buffer (``dist.allreduce_grads`` reduces it with a single collective)."""
    lib = _lib.load()
    lin = _linears(critic.model)
    ws_ = [F_._f32c(l.weight.detach()) for l in lin]
    bs_ = [F_._f32c(l.bias.detach()) for l in lin]
    if ws_[-1].shape[0] != 1:
        raise BdError("value_update: the critic must have one output")
    a1 = F_._f32c(beliefs.detach()).reshape(-1, beliefs.shape[-1])
    a2 = F_._f32c(states.detach()).reshape(-1, states.shape[-1]) if states is not None else None
    rows = a1.shape[0]
    k2 = a2.shape[1] if a2 is not None else 0
    if a1.shape[1] + k2 != ws_[0].shape[1] or (a2 is not None and a2.shape[0] != rows):
        raise BdError("value_update: input widths do not match the critic's first layer")
    tgt = F_._f32c(target.detach()).reshape(-1)
    wgt = F_._f32c(weight.detach()).reshape(-1) if weight is not None else None
    if tgt.numel() != rows or (wgt is not None and wgt.numel() != rows):
        raise BdError("value_update: target / weight must hold one value per row")
    prec = F_._prec()
    mlp = _lib.make_mlp(ws_, bs_, _mlp_act_id(critic.model))
    dev = a1.device
    v = torch.empty(rows, 1, device=dev, dtype=torch.float32)
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    if rows == 0:
        return loss.zero_()[0]
    ws = _lib.workspace(max(lib.bd_mlp_workspace_bytes(C.byref(mlp), rows, 1), 1 << 16), dev)
    nsaved = lib.bd_mlp_saved_bytes(C.byref(mlp), a1.shape[1], k2, rows, prec)
    saved = torch.empty(nsaved, dtype=torch.uint8, device=dev) if nsaved else None
    _lib.check(lib.bd_mlp_forward_save(C.byref(mlp), _lib.ptr(a1), a1.shape[1], _lib.ptr(a2), k2, rows,
                                       _lib.ptr(v), saved.data_ptr() if nsaved else None, ws.data_ptr(),
                                       ws.numel(), prec, _lib.stream_ptr()), "bd_mlp_forward_save")
    dv = torch.empty(rows, 1, device=dev, dtype=torch.float32)
    _lib.check(lib.bd_value_loss(_lib.ptr(v), _lib.ptr(tgt), _lib.ptr(wgt), rows, _lib.ptr(loss),
                                 _lib.ptr(dv), ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "bd_value_loss")
    params = [l.weight for l in lin] + [l.bias for l in lin]
    need = [p.requires_grad for p in params]
    g = F_._zero_grads(need, params)
    n = len(lin)
    args = _lib.MlpBwdArgs()
    args.x1, args.k1, args.x2, args.k2 = _lib.ptr(a1), a1.shape[1], _lib.ptr(a2), k2
    args.rows, args.dy = rows, _lib.ptr(dv)
    for i in range(n):
        args.dw[i], args.db[i] = _lib.ptr(g[i]), _lib.ptr(g[n + i])
    if saved is not None:
        args.saved = saved.data_ptr()
    _lib.check(lib.bd_mlp_backward(C.byref(mlp), C.byref(args), ws.data_ptr(), ws.numel(), prec,
                                   _lib.stream_ptr()), "bd_mlp_backward")
    for p, gp in zip(params, g):
        if gp is None:
            continue
        if p.grad is None:
            p.grad = gp
        else:
            p.grad.add_(gp)
    return loss[0]


def _fused_min_rows() -> int:
    return int(os.environ.get("BD_FUSED_MIN_ROWS", "8192"))


def _fused_heads_ok(tm, actor, reward_model, value_model) -> bool:
    """Can bd_imagine_returns_* run this configuration in the current precision mode?"""
    if F_.get_precision() == "fp32":
        return False                      # check mode stays piecewise (the fp32 kernels)
    for m in (reward_model, value_model):
        if not isinstance(m, DenseModel):
            return False
    lib = _lib.load()
    dims = rssm_dims(tm)
    rp = rssm_params(tm, with_posterior=False)
    det = lambda ps: [F_._f32c(p.detach()) for p in ps]
    r = F_.make_rssm(det(rp[:10]) + [None] * 4, dims)
    la, lr, lv = _linears(actor.model), _linears(reward_model.model), _linears(value_model.model)
    mk = lambda lin, act: _lib.make_mlp(det([l.weight for l in lin]), det([l.bias for l in lin]), act)
    if _mlp_act_id(reward_model.model) != _mlp_act_id(value_model.model):
        return False
    ma, mr, mv = mk(la, _mlp_act_id(actor.model)), mk(lr, _mlp_act_id(reward_model.model)), mk(lv, _mlp_act_id(value_model.model))
    return bool(lib.bd_imagine_returns_supported(C.byref(r), C.byref(ma), C.byref(mr), C.byref(mv), F_._prec()))


def imagine_and_returns(self, prev_state: Tensor, prev_belief: Tensor, reward_model, value_model,
                        discount: float, lambda_: float,
                        noise: Optional[Dict[str, Tensor]] = None, assume_frozen_heads: bool = False,
                        fused: Optional[bool] = None):
    """Fused entry (level L2): ``imagine_ahead`` + ``reward_model(b, s)`` +
    ``value_model(b, s)`` + ``lambda_return(reward, value, value[-1], discount, lambda_)`` of
    ``Dreamer.train_step`` (src/dreamer.py:313-335) as ONE forward call and ONE backward call.  Returns
    beliefs, states, (means, stds), entropy, reward, value, returns with the reference's shapes.

    In the tensor-core modes the heads ride inside the persistent rollout kernel and lambda_return is its
    tail (bd_imagine_returns_forward / _backward).  Head and transition weights are treated as constants,
    which is what the reference's FreezeParameters blocks make them (:313, :320).  Configurations the
    fused kernels do not cover (fp32 check mode, non-DenseModel heads, sizes beyond the tile limits,
    head parameters that require grad) run the same arithmetic through the piecewise entry points.

    ``fused``: True = the fused kernels whenever they cover the configuration, False = piecewise, None
    (default) = by row count.  The fused rollout puts the heads' eight layers on the per-step serial chain
    of every row tile, which pays once the row tiles fill the GPU several times over (measured on B200:
    2^17 rows 27.3 ms fused vs 30.2 ms piecewise, 2^14 rows 3.95 vs 4.36 ms) and costs when they do not
    (2 500 rows: 2.65 vs 2.00 ms, where the piecewise heads run as two batched MLP passes over all T*N rows
    on every SM); the switch point is ``BD_FUSED_MIN_ROWS`` (default 8 192 = 64 row tiles)."""
    tm, actor = self.transition_model, self.actor
    if fused is None:
        fused = prev_state.numel() // max(1, prev_state.shape[-1]) >= _fused_min_rows()
    heads_frozen = assume_frozen_heads or not (torch.is_grad_enabled() and any(
        p.requires_grad for m in (reward_model, value_model) for p in m.parameters()))
    if (getattr(self, "latent_distribution", "Gaussian") == "Gaussian"
            and getattr(actor, "action_distribution", "Gaussian") == "Gaussian"
            and fused and heads_frozen and prev_state.is_cuda
            and _fused_heads_ok(tm, actor, reward_model, value_model)):
        T = self.planning_horizon - 1
        b0 = prev_belief.reshape(-1, prev_belief.shape[-1])
        s0 = prev_state.reshape(-1, prev_state.shape[-1])
        dims = rssm_dims(tm)
        rp = rssm_params(tm, with_posterior=False)[:10]
        if torch.is_grad_enabled() and any(p.requires_grad for p in rp):
            raise NotImplementedError(
                "imagine_and_returns: transition-model parameters require grad; call it under "
                "FreezeParameters(model_modules) as Dreamer.train_step does (src/dreamer.py:313)")
        if noise is None:
            noise = draw_imagine_noise(T, s0.shape[0], dims["S"], dims["A"], s0.device)
        la, lr, lv = _linears(actor.model), _linears(reward_model.model), _linears(value_model.model)
        flat = lambda lin: [t for l in lin for t in (l.weight, l.bias)]
        (beliefs, states, means, stds, entropy, _actions, reward, value,
         returns) = F_.ImagineReturnsFunction.apply(
            dims, actor_config(actor), _mlp_act_id(reward_model.model), T, float(discount), float(lambda_),
            s0, b0, noise["eps_a"], noise["eps_e"], noise["eps_s"], len(la), len(lr),
            *flat(la), *rp, *[t.detach() for t in flat(lr)], *[t.detach() for t in flat(lv)])
        return beliefs, states, (means, stds), entropy, reward, value, returns
    beliefs, states, (means, stds), entropy = imagine_ahead(self, prev_state, prev_belief, noise)
    reward, value = heads_pair(reward_model, value_model, beliefs, states)
    returns = lambda_return(reward, value, value[-1], discount, lambda_)
    return beliefs, states, (means, stds), entropy, reward, value, returns


def heads_pair(reward_model, value_model, beliefs: Tensor, states: Tensor):
    """``reward_model(beliefs, states), value_model(beliefs, states)`` (src/dreamer.py:321-322).  Two DenseModels of
    the same shape run as ONE launch in the tensor-core modes (bd_heads_forward: one tile prologue, the two layer
    chains interleaved so the MMAs of one head run under the epilogue of the other); anything else is the two calls."""
    if (isinstance(reward_model, DenseModel) and isinstance(value_model, DenseModel) and beliefs.is_cuda
            and os.environ.get("BD_HEADS_PAIR", "1") != "0"
            and _mlp_act_id(reward_model.model) == _mlp_act_id(value_model.model)):
        lr, lv = _linears(reward_model.model), _linears(value_model.model)
        wr, br = [l.weight for l in lr], [l.bias for l in lr]
        wv, bv = [l.weight for l in lv], [l.bias for l in lv]
        act = _mlp_act_id(reward_model.model)
        if len(lr) == len(lv) and F_.heads_pair_supported(act, beliefs, states, wr, br, wv, bv):
            flat = lambda ws, bs: [t for w, b in zip(ws, bs) for t in (w, b)]
            return F_.HeadsPairFunction.apply(act, len(lr), beliefs, states, *flat(wr, br), *flat(wv, bv))
    return reward_model(beliefs, states), value_model(beliefs, states)


# =============================================================================
# MPCPlanner (CEM)
# =============================================================================
class MPCPlanner(nn.Module):
    """Cross-entropy-method planner; reference: src/planner.py:5-90."""

    def __init__(self, action_size, planning_horizon, optimisation_iters, candidates,
                 top_candidates, transition_model, reward_model):
        super().__init__()
        self.transition_model = transition_model
        self.reward_model = reward_model
        self.action_size = action_size
        self.planning_horizon = planning_horizon
        self.optimisation_iters = optimisation_iters
        self.candidates, self.top_candidates = candidates, top_candidates
        self.last_trace = None
        self.shard_candidates = True   # split candidates over ranks when torch.distributed is up

    def _models(self):
        tm, rm = self.transition_model, self.reward_model
        dims = rssm_dims(tm)
        rp = [F_._f32c(p.detach()) for p in rssm_params(tm, with_posterior=False)[:10]] + [None] * 4
        lin = _linears(rm.model)
        act = None
        for m in rm.model:
            if not isinstance(m, nn.Linear):
                act = type(m).__name__
                break
        ws_ = [F_._f32c(l.weight.detach()) for l in lin]
        bs_ = [F_._f32c(l.bias.detach()) for l in lin]
        return dims, rp, ws_, bs_, _lib.activation_id(act or "Identity")

    def draw_noise(self, B: int, device, generator=None):
        """ε_act (iters,H,B,C,A) and ε_s (iters,H,B*C,S); the reference draws, per iteration,
        randn(H,B,C,A) then H x randn_like((B*C,S)) (src/planner.py:53, src/models.py:72)."""
        dims = rssm_dims(self.transition_model)
        kw = dict(device=device, dtype=torch.float32, generator=generator)
        I, H, Cn = self.optimisation_iters, self.planning_horizon, self.candidates
        return dict(eps_act=torch.randn(I, H, B, Cn, self.action_size, **kw),
                    eps_s=torch.randn(I, H, B * Cn, dims["S"], **kw))

    @torch.no_grad()
    def forward(self, belief: Tensor, state: Tensor, noise: Optional[Dict[str, Tensor]] = None,
                trace: bool = False) -> Tensor:
        """belief (B,Be), state (B,S) -> first action mean (B,A)."""
        lib = _lib.load()
        belief, state = F_._f32c(belief), F_._f32c(state)
        B = belief.shape[0]
        dims, rp, ws_, bs_, ract = self._models()
        if dims["A"] != self.action_size:
            raise BdError("MPCPlanner: action_size differs from the transition model's")
        if noise is None:
            noise = self.draw_noise(B, belief.device)
        ea, es = F_._f32c(noise["eps_act"]), F_._f32c(noise["eps_s"])
        I, H, Cn, K, A = (self.optimisation_iters, self.planning_horizon, self.candidates,
                          self.top_candidates, self.action_size)
        if ea.shape != (I, H, B, Cn, A) or es.shape != (I, H, B * Cn, dims["S"]):
            raise BdError(f"MPCPlanner: noise shapes {tuple(ea.shape)}, {tuple(es.shape)} mismatch")
        from . import dist as D_
        if self.shard_candidates and D_.world_size() > 1:
            return self._forward_sharded(belief, state, ea, es, dims, rp, ws_, bs_, ract, trace)
        a = _lib.CemPlanArgs()
        a.rssm = F_.make_rssm(rp, dims)
        a.reward = _lib.make_mlp(ws_, bs_, ract)
        a.B, a.C, a.K, a.H, a.iters = B, Cn, K, H, I
        a.belief, a.state, a.eps_act, a.eps_s = (_lib.ptr(t) for t in (belief, state, ea, es))
        out = torch.empty(B, A, device=belief.device, dtype=torch.float32)
        a.action_out = _lib.ptr(out)
        if trace:
            rt = torch.empty(I, B, Cn, device=belief.device, dtype=torch.float32)
            tk = torch.empty(I, B, K, device=belief.device, dtype=torch.int64)
            a.returns_trace, a.topk_trace = _lib.ptr(rt), _lib.ptr(tk)
            self.last_trace = dict(returns=rt, topk=tk)
        nbytes = lib.bd_cem_plan_workspace_bytes(C.byref(a.rssm), C.byref(a.reward), B, Cn, K, H)
        ws = _lib.workspace(nbytes, belief.device)
        _lib.check(lib.bd_cem_plan(C.byref(a), ws.data_ptr(), ws.numel(), F_._prec(),
                                   _lib.stream_ptr()), "bd_cem_plan")
        return out

    def _forward_sharded(self, belief, state, ea, es, dims, rp, ws_, bs_, ract, trace):
        """Candidates split over ranks (weights replicated).  Per iteration: local rollout of
        C/G candidates -> all-gather of (returns, actions) -> identical global top-K + refit on
        every rank.  The noise tensors are the GLOBAL ones, sliced along the candidate axis, so
        the result does not depend on the number of ranks."""
        import torch.distributed as tdist
        from . import dist as D_
        lib = _lib.load()
        world, rank = D_.world_size(), tdist.get_rank()
        I, H, Cn, K, A = (self.optimisation_iters, self.planning_horizon, self.candidates,
                          self.top_candidates, self.action_size)
        B, dev = belief.shape[0], belief.device
        ranges = [D_.shard_range(Cn, r, world) for r in range(world)]
        sizes = [e - b for b, e in ranges]
        c0, c1 = ranges[rank]
        Cl = c1 - c0
        e = _lib.CemEvalArgs()
        e.rssm = F_.make_rssm(rp, dims)
        e.reward = _lib.make_mlp(ws_, bs_, ract)
        e.B, e.C, e.H, e.c_begin, e.c_end = B, Cn, H, c0, c1
        e.belief, e.state = _lib.ptr(belief), _lib.ptr(state)
        mean = torch.zeros(H, B, A, device=dev)
        std = torch.ones(H, B, A, device=dev)
        actions = torch.empty(H, B, Cl, A, device=dev)
        returns = torch.empty(B, Cl, device=dev)
        e.action_mean, e.action_std = _lib.ptr(mean), _lib.ptr(std)
        e.actions, e.returns = _lib.ptr(actions), _lib.ptr(returns)
        nbytes = lib.bd_cem_workspace_bytes(C.byref(e.rssm), C.byref(e.reward), B, Cl, H)
        ws = _lib.workspace(nbytes, dev)
        idx = torch.empty(B, K, device=dev, dtype=torch.int64)
        tr_r, tr_k = [], []
        for it in range(I):
            e.eps_act, e.eps_s = _lib.ptr(ea[it]), _lib.ptr(es[it])
            _lib.check(lib.bd_cem_evaluate(C.byref(e), ws.data_ptr(), ws.numel(), F_._prec(),
                                           _lib.stream_ptr()), "bd_cem_evaluate")
            g_ret, g_act = D_.gather_candidates(returns, actions, sizes)
            _lib.check(lib.bd_cem_refit(_lib.ptr(g_ret), _lib.ptr(g_act), B, Cn, K, H, A,
                                        _lib.ptr(idx), _lib.ptr(mean), _lib.ptr(std),
                                        _lib.stream_ptr()), "bd_cem_refit")
            if trace:
                tr_r.append(g_ret.clone())
                tr_k.append(idx.clone())
        if trace:
            self.last_trace = dict(returns=torch.stack(tr_r), topk=torch.stack(tr_k))
        return mean[0].clone()
