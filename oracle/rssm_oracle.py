"""CPU oracle for the RSSM latent-dynamics hot path of jgsimard/big-dreamer.

TEST INFRASTRUCTURE ONLY.  Nothing under ``big_dreamer_b200/`` imports this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may call it, and there only as the
checker / CPU baseline, never as the product path.

It is a plain-PyTorch (CPU, fp32 or fp64 -- the dtype follows the inputs)
restatement of the reference arithmetic with every Gaussian draw replaced by an
explicit noise argument.  Each function cites the reference lines it follows
(paths relative to /root/reference).

Parity pinning: the reference ships no golden vectors for this path (its test
suite is one FreezeParameters test, test/test_utils.py:5-20).  The restatement
is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/ref_harness.py
imports the unmodified reference here, replays a recorded noise tape through it
and (a) tests/test_oracle_vs_reference.py compares every function below with the
real reference in this container, (b) oracle/make_golden.py freezes reference
outputs as fixtures under tests/golden/ that travel to the GPU box.

Weights are passed as ``state_dict()``-style dicts with the reference's own keys
(src/models.py:149-167, src/utils.py:368-404):
  transition: rnn.{weight_ih,weight_hh,bias_ih,bias_hh}, fc_embed_state_action.0.*,
              belief_prior.model.{0,2}.*, belief_posterior.model.{0,2}.*
  dense/actor: model.{0,2,4,...}.{weight,bias}
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

ENTROPY_SAMPLES = 100  # SampleDist(samples=100), src/models.py:684


# ----------------------------------------------------------------------------
# activations (build_mlp resolves getattr(nn, name), src/utils.py:391-394)
# ----------------------------------------------------------------------------
def activation(name: str, x: Tensor) -> Tensor:
    if name == "ELU":
        return F.elu(x)
    if name == "ReLU":
        return F.relu(x)
    if name == "Tanh":
        return torch.tanh(x)
    if name == "Sigmoid":
        return torch.sigmoid(x)
    if name == "Identity":
        return x
    raise ValueError(f"oracle: unsupported activation {name}")


def mlp_num_layers(sd: SD, prefix: str = "model.") -> int:
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)})
    return len(idx)


def mlp(sd: SD, x: Tensor, act: str, prefix: str = "model.", out_act: str = "Identity") -> Tensor:
    """build_mlp: [Linear, act] * n + Linear + out_act (src/utils.py:396-404).
    Linear i lives at Sequential index 2*i."""
    n = mlp_num_layers(sd, prefix)
    for i in range(n):
        w, b = sd[f"{prefix}{2 * i}.weight"], sd[f"{prefix}{2 * i}.bias"]
        x = F.linear(x, w, b)
        x = activation(act if i < n - 1 else out_act, x)
    return x


def dense(sd: SD, act: str, *args: Tensor) -> Tensor:
    """DenseModel.forward (src/models.py:393-408): cat(belief, state, -1) -> MLP."""
    x = torch.cat(list(args), dim=-1) if len(args) == 2 else args[0]
    return mlp(sd, x, act)


# ----------------------------------------------------------------------------
# one transition step
# ----------------------------------------------------------------------------
def gru_cell(sd: SD, x: Tensor, h: Tensor) -> Tensor:
    """nn.GRUCell semantics (src/models.py:149,252); gate order r,z,n and b_hn
    inside the r-product."""
    gi = F.linear(x, sd["rnn.weight_ih"], sd["rnn.bias_ih"])
    gh = F.linear(h, sd["rnn.weight_hh"], sd["rnn.bias_hh"])
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h


def gaussian_belief(sd: SD, prefix: str, act: str, x: Tensor, eps: Tensor, min_std: float):
    """GaussianBeliefModel.forward (src/models.py:60-73)."""
    out = mlp(sd, x, act, prefix=prefix)
    mean, raw = torch.chunk(out, 2, dim=1)
    std = F.softplus(raw) + min_std
    return mean + std * eps, mean, std


def transition_step(sd: SD, act: str, min_std: float, state: Tensor, action: Tensor,
                    belief: Tensor, eps_prior: Tensor):
    """src/models.py:251-256 == src/dreamer.py:219-223."""
    hidden = activation(act, F.linear(torch.cat([state, action], dim=1),
                                      sd["fc_embed_state_action.0.weight"],
                                      sd["fc_embed_state_action.0.bias"]))
    belief = gru_cell(sd, hidden, belief)
    s, mean, std = gaussian_belief(sd, "belief_prior.model.", act, belief, eps_prior, min_std)
    return belief, s, mean, std


# ----------------------------------------------------------------------------
# TransitionModel.forward (prior-only and observe), src/models.py:190-299
# ----------------------------------------------------------------------------
def transition_forward(sd: SD, act: str, min_std: float, init_state: Tensor, actions: Tensor,
                       init_belief: Tensor, eps_prior: Tensor,
                       embeddings: Optional[Tensor] = None,
                       nonterminals: Optional[Tensor] = None,
                       eps_post: Optional[Tensor] = None):
    """eps_prior, eps_post: (L, B, S) -- draw order per step is prior then
    posterior (src/models.py:256, 267)."""
    L = actions.size(0)
    belief, state = init_belief, init_state
    beliefs, pri, pm, ps, pos, qm, qs = [], [], [], [], [], [], []
    for t in range(L):
        _state = state if nonterminals is None else state * nonterminals[t]      # :241-247
        belief, s_prior, m, sd_ = transition_step(sd, act, min_std, _state, actions[t], belief,
                                                  eps_prior[t])
        beliefs.append(belief); pri.append(s_prior); pm.append(m); ps.append(sd_)
        if embeddings is not None:
            # t_ = t - 1; embeddings[t_ + 1] == embeddings[t]  (src/models.py:265-266)
            s_post, mq, sq = gaussian_belief(sd, "belief_posterior.model.", act,
                                             torch.cat([belief, embeddings[t]], dim=1),
                                             eps_post[t], min_std)
            pos.append(s_post); qm.append(mq); qs.append(sq)
            state = s_post
        else:
            state = s_prior
    st = lambda xs: torch.stack(xs, dim=0)
    if embeddings is None:
        return st(beliefs), st(pri), (st(pm), st(ps)), None, None
    return st(beliefs), st(pri), (st(pm), st(ps)), st(pos), (st(qm), st(qs))


# ----------------------------------------------------------------------------
# KL loss of the dynamics update
# ----------------------------------------------------------------------------
def kl_normal(mq: Tensor, sq: Tensor, mp: Tensor, sp: Tensor) -> Tensor:
    """torch.distributions.kl._kl_normal_normal(Normal(mq, sq), Normal(mp, sp)), elementwise."""
    var_ratio = (sq / sp) ** 2
    t1 = ((mq - mp) / sp) ** 2
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def kl_loss(posterior_params, prior_params, free_nats: Tensor, kl_balance: float = -1) -> Tensor:
    """Planet._kl_loss (src/planet.py:288-308) when kl_balance == -1, else the balanced form of
    Dreamer._kl_loss (src/dreamer.py:129-144).  free_nats: (1,) tensor as in src/planet.py:98."""
    qm, qs = posterior_params
    pm, ps = prior_params
    if kl_balance == -1:
        div = kl_normal(qm, qs, pm, ps).sum(dim=2)
        return torch.max(div, free_nats).mean(dim=(0, 1))
    lhs = kl_normal(qm.detach(), qs.detach(), pm, ps).mean()
    rhs = kl_normal(qm, qs, pm.detach(), ps.detach()).mean()
    return kl_balance * torch.max(lhs, free_nats) + (1 - kl_balance) * torch.max(rhs, free_nats)


# ----------------------------------------------------------------------------
# actor: ActorModel.forward + Dreamer.get_action + SampleDist.entropy
# ----------------------------------------------------------------------------
TANH_CLAMP = 0.99999997  # src/models.py:663 (rounds to 0.99999994 in fp32)


def actor_mean_std(actor_sd: SD, act: str, belief: Tensor, state: Tensor,
                   mean_scale: float = 5.0, init_std: float = 5.0, min_std: float = 1e-4):
    """ActorModel.forward, Gaussian branch (src/models.py:506-517)."""
    out = mlp(actor_sd, torch.cat([belief, state], dim=1), act)
    m_raw, s_raw = torch.chunk(out, 2, dim=1)
    raw_init_std = torch.log(torch.exp(torch.tensor(init_std)) - 1).to(out.dtype)   # :503
    mean = mean_scale * torch.tanh(m_raw / mean_scale)
    std = F.softplus(s_raw + raw_init_std) + min_std
    return mean, std


def tanh_normal_logprob(y: Tensor, mean: Tensor, std: Tensor) -> Tensor:
    """log_prob of Independent(TransformedDistribution(Normal, TanhBijector), 1)
    evaluated at y (src/dreamer.py:435-437, src/models.py:656-673)."""
    yc = torch.where(torch.abs(y) <= 1.0, torch.clamp(y, -TANH_CLAMP, TANH_CLAMP), y)
    x = 0.5 * torch.log((1 + yc) / (1 - yc))
    var = std ** 2
    base = -((x - mean) ** 2) / (2 * var) - torch.log(std) - math.log(math.sqrt(2 * math.pi))
    ladj = 2.0 * (math.log(2.0) - x - F.softplus(-2.0 * x))
    return (base - ladj).sum(-1)


def get_action(actor_sd: SD, act: str, belief: Tensor, state: Tensor, eps_a: Tensor,
               eps_e: Tensor, **actor_kw):
    """Dreamer.get_action, deterministic=False (src/dreamer.py:429-444).
    eps_a: (N,A) rsample noise; eps_e: (100,N,A) entropy noise."""
    mean, std = actor_mean_std(actor_sd, act, belief, state, **actor_kw)
    action = torch.tanh(mean + eps_a * std)                 # Normal.rsample: loc + eps*scale
    y = torch.tanh(mean.unsqueeze(0) + eps_e * std.unsqueeze(0))
    logprob = tanh_normal_logprob(y, mean.unsqueeze(0), std.unsqueeze(0))
    entropy = -torch.mean(logprob, 0)                       # src/models.py:733
    return action, entropy


def get_action_mode(actor_sd: SD, act: str, belief: Tensor, state: Tensor, eps_m: Tensor, **actor_kw):
    """Dreamer.get_action, deterministic=True (src/dreamer.py:440-441): SampleDist.mode
    (src/models.py:707-723) -- of 100 samples the one with the largest log-probability.
    eps_m: (100,N,A), the draw of ``dist.rsample()`` on the expanded distribution."""
    mean, std = actor_mean_std(actor_sd, act, belief, state, **actor_kw)
    y = torch.tanh(mean.unsqueeze(0) + eps_m * std.unsqueeze(0))
    logprob = tanh_normal_logprob(y, mean.unsqueeze(0), std.unsqueeze(0))        # (100, N)
    idx = torch.argmax(logprob, dim=0).reshape(1, -1, 1).expand(1, y.size(1), y.size(2))
    return torch.gather(y, 0, idx).squeeze(0)


def act_step(trans_sd: SD, actor_sd: SD, act: str, min_std: float, belief: Tensor, state: Tensor,
             action: Tensor, embedding: Tensor, eps_prior: Tensor, eps_post: Tensor, eps_act: Tensor,
             deterministic: bool = False, **actor_kw):
    """Planet.update_belief_and_act (src/planet.py:370-403) for an actor policy, without the encoder,
    the exploration noise and env.step: one posterior step of the transition model (no nonterminals),
    then Dreamer.get_action on the new (belief, posterior state).
    belief (B,Be), state (B,S), action (B,A), embedding (B,E), eps_prior / eps_post (B,S),
    eps_act (B,A) or, deterministic, (100,B,A).  -> belief, posterior_state, action."""
    b, _, _, post, _ = transition_forward(trans_sd, act, min_std, state, action[None], belief, eps_prior[None],
                                          embedding[None], None, eps_post[None])
    b, post = b[0], post[0]
    if deterministic:
        a = get_action_mode(actor_sd, act, b, post, eps_act, **actor_kw)
    else:
        mean, std = actor_mean_std(actor_sd, act, b, post, **actor_kw)
        a = torch.tanh(mean + eps_act * std)
    return b, post, a


# ----------------------------------------------------------------------------
# Dreamer.imagine_ahead (src/dreamer.py:178-237)
# ----------------------------------------------------------------------------
def imagine_ahead(trans_sd: SD, actor_sd: SD, act: str, min_std: float, planning_horizon: int,
                  prev_state: Tensor, prev_belief: Tensor, eps_a: Tensor, eps_e: Tensor,
                  eps_s: Tensor, **actor_kw):
    """prev_state (L,B,S) / prev_belief (L,B,Be) are flattened to N=L*B rows.
    eps_a (T,N,A), eps_e (T,100,N,A), eps_s (T,N,S), T = planning_horizon-1.
    Returns beliefs (T,N,Be), states (T,N,S), (means, stds), entropy (T,N),
    plus the actions (T,N,A) (not a reference output; useful for checks)."""
    belief = prev_belief.reshape(-1, prev_belief.size(-1))
    state = prev_state.reshape(-1, prev_state.size(-1))
    T = planning_horizon - 1
    B_, S_, M_, D_, E_, A_ = [], [], [], [], [], []
    for t in range(T):
        action, ent = get_action(actor_sd, act, belief.detach(), state.detach(), eps_a[t],
                                 eps_e[t], **actor_kw)                     # :215
        belief, state, m, sd_ = transition_step(trans_sd, act, min_std, state, action, belief,
                                                eps_s[t])                  # :219-223
        B_.append(belief); S_.append(state); M_.append(m); D_.append(sd_); E_.append(ent)
        A_.append(action)
    st = lambda xs: torch.stack(xs, dim=0)
    return st(B_), st(S_), (st(M_), st(D_)), st(E_), st(A_)


# ----------------------------------------------------------------------------
# lambda_return (src/dreamer.py:447-471) and the actor loss (src/dreamer.py:329-353)
# ----------------------------------------------------------------------------
def lambda_return(imged_reward: Tensor, value_pred: Tensor, bootstrap: Tensor,
                  discount: float = 0.99, lambda_: float = 0.95) -> Tensor:
    next_values = torch.cat([value_pred[1:], bootstrap[None]], 0)
    inputs = imged_reward + discount * next_values * (1 - lambda_)
    last = bootstrap
    outs = []
    for t in reversed(range(inputs.size(0))):
        last = inputs[t] + discount * lambda_ * last
        outs.append(last)
    return torch.stack(list(reversed(outs)), 0)


def actor_loss(trans_sd: SD, actor_sd: SD, reward_sd: SD, value_sd: SD, act: str, min_std: float,
               planning_horizon: int, prev_state: Tensor, prev_belief: Tensor, eps_a: Tensor,
               eps_e: Tensor, eps_s: Tensor, discount: float = 0.995, lambda_: float = 0.95,
               entropy_weight: float = 1e-5, **actor_kw):
    """Behaviour-learning block of Dreamer.train_step (src/dreamer.py:313-353).
    Returns (loss, dict of intermediates)."""
    beliefs, states, (means, stds), entropy, actions = imagine_ahead(
        trans_sd, actor_sd, act, min_std, planning_horizon, prev_state, prev_belief,
        eps_a, eps_e, eps_s, **actor_kw)
    reward = dense(reward_sd, act, beliefs, states)
    value = dense(value_sd, act, beliefs, states)
    returns = lambda_return(reward, value, value[-1], discount, lambda_)
    objective = returns
    if entropy_weight != -1:
        objective = objective + entropy_weight * entropy.unsqueeze(-1)
    loss = -objective.mean()
    return loss, dict(beliefs=beliefs, states=states, means=means, stds=stds, entropy=entropy,
                      actions=actions, reward=reward, value=value, returns=returns)


# ----------------------------------------------------------------------------
# MPCPlanner.forward (src/planner.py:28-90)
# ----------------------------------------------------------------------------
def cem_plan(trans_sd: SD, reward_sd: SD, act: str, min_std: float, action_size: int,
             planning_horizon: int, optimisation_iters: int, candidates: int, top_candidates: int,
             belief: Tensor, state: Tensor, eps_act: Tensor, eps_s: Tensor,
             return_trace: bool = False):
    """eps_act (iters,H,B,C,A); eps_s (iters,H,B*C,S).  Row b*C+c is candidate c
    of batch row b (src/planner.py:37-39)."""
    B, Be, Z = belief.size(0), belief.size(1), state.size(1)
    H, C, K, A = planning_horizon, candidates, top_candidates, action_size
    belief = belief.unsqueeze(1).expand(B, C, Be).reshape(-1, Be)
    state = state.unsqueeze(1).expand(B, C, Z).reshape(-1, Z)
    mean = torch.zeros(H, B, 1, A, dtype=belief.dtype)
    std = torch.ones(H, B, 1, A, dtype=belief.dtype)
    trace = []
    for it in range(optimisation_iters):
        actions = (mean + std * eps_act[it]).view(H, B * C, A)                     # :53-62
        beliefs, states, _, _, _ = transition_forward(trans_sd, act, min_std, state, actions,
                                                      belief, eps_s[it])           # :65
        returns = dense(reward_sd, act, beliefs.view(-1, Be), states.view(-1, Z)) \
            .view(H, -1).sum(dim=0)                                                # :68-72
        _, topk = returns.reshape(B, C).topk(K, dim=1, largest=True, sorted=False)  # :74-76
        topk = topk + C * torch.arange(0, B, dtype=torch.int64).unsqueeze(1)       # :78-80
        best = actions[:, topk.view(-1)].reshape(H, B, K, A)                       # :81-83
        mean = best.mean(dim=2, keepdim=True)                                      # :86
        std = best.std(dim=2, unbiased=False, keepdim=True)                        # :87
        if return_trace:
            trace.append(dict(returns=returns.reshape(B, C).clone(),
                              topk=torch.sort(topk - C * torch.arange(B).unsqueeze(1), dim=1)[0],
                              mean=mean.clone(), std=std.clone()))
    out = mean[0].squeeze(dim=1)                                                   # :90
    return (out, trace) if return_trace else out


# ----------------------------------------------------------------------------
# helpers shared by tests and bench: synthetic weights / latents / noise
# ----------------------------------------------------------------------------
def _linear_init(gen: torch.Generator, out_f: int, in_f: int, dtype=torch.float32):
    """PyTorch default Linear/GRU init scale: U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    k = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, dtype=dtype) * 2 - 1) * k
    b = (torch.rand(out_f, generator=gen, dtype=dtype) * 2 - 1) * k
    return w, b


def make_mlp_sd(gen, sizes, prefix="model.", dtype=torch.float32) -> SD:
    sd = {}
    for i in range(len(sizes) - 1):
        w, b = _linear_init(gen, sizes[i + 1], sizes[i], dtype)
        sd[f"{prefix}{2 * i}.weight"], sd[f"{prefix}{2 * i}.bias"] = w, b
    return sd


def make_transition_sd(gen, Be, S, A, Hi, E, dtype=torch.float32) -> SD:
    sd = {}
    k = 1.0 / math.sqrt(Be)
    for name, shape in (("rnn.weight_ih", (3 * Be, Be)), ("rnn.weight_hh", (3 * Be, Be)),
                        ("rnn.bias_ih", (3 * Be,)), ("rnn.bias_hh", (3 * Be,))):
        sd[name] = (torch.rand(*shape, generator=gen, dtype=dtype) * 2 - 1) * k
    sd.update(make_mlp_sd(gen, [S + A, Be], "fc_embed_state_action.", dtype))
    sd.update(make_mlp_sd(gen, [Be, Hi, 2 * S], "belief_prior.model.", dtype))
    sd.update(make_mlp_sd(gen, [Be + E, Hi, 2 * S], "belief_posterior.model.", dtype))
    return sd


def make_models(seed: int, Be: int, S: int, A: int, Hi: int, E: int, n_layers: int = 4,
                dtype=torch.float32):
    """Random-init weights of the reference architecture (default-init scale)."""
    gen = torch.Generator().manual_seed(seed)
    trans = make_transition_sd(gen, Be, S, A, Hi, E, dtype)
    head = [Be + S] + [Hi] * n_layers
    reward = make_mlp_sd(gen, head + [1], dtype=dtype)
    value = make_mlp_sd(gen, head + [1], dtype=dtype)
    actor = make_mlp_sd(gen, head + [2 * A], dtype=dtype)
    return trans, actor, reward, value


def make_latents(seed: int, N: int, Be: int, S: int, dtype=torch.float32):
    """SURVEY 8d recipe: beliefs tanh(N(0,1)), states 0.5*N(0,1)."""
    gen = torch.Generator().manual_seed(seed + 1000)
    b = torch.tanh(torch.randn(N, Be, generator=gen, dtype=dtype))
    s = 0.5 * torch.randn(N, S, generator=gen, dtype=dtype)
    return s, b


def make_imagine_noise(seed: int, T: int, N: int, S: int, A: int, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed + 2000)
    eps_a = torch.randn(T, N, A, generator=gen, dtype=dtype)
    eps_e = torch.randn(T, ENTROPY_SAMPLES, N, A, generator=gen, dtype=dtype)
    eps_s = torch.randn(T, N, S, generator=gen, dtype=dtype)
    return eps_a, eps_e, eps_s
