"""Runs the UNMODIFIED reference (jgsimard/big-dreamer at /root/reference) with
injected noise.  TEST INFRASTRUCTURE ONLY -- used in this container to (a) pin
oracle/rssm_oracle.py against the real code and (b) generate tests/golden/*.
/root/reference does not exist on the GPU box, so nothing that runs there may
import this module (``available()`` is False there and the tests skip).

The reference draws its Gaussians internally: ``torch.randn_like``
(src/models.py:72), ``torch.randn`` (src/planner.py:53) and
``Normal.rsample`` -> ``torch.distributions.normal._standard_normal``
(src/dreamer.py:443, src/models.py:731).  ``NoiseTape`` replaces those three
symbols with a FIFO of pre-generated tensors (draw order: SURVEY.md A.3).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from typing import List

import torch

REFERENCE_ROOT = os.environ.get("BD_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models.py"))


def load():
    """Import the reference's models / planner / dreamer modules (with the four
    import shims ahead of site-packages)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("typeguard", "torchtyping", "plotly", "gym")}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "src"))
    sys.path.insert(0, _SHIMS)
    try:
        import models as ref_models          # noqa: E402
        import planner as ref_planner        # noqa: E402
        import dreamer as ref_dreamer        # noqa: E402
        import utils as ref_utils            # noqa: E402
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():      # leave real typeguard etc. for everyone else
            if v is not None:
                sys.modules[k] = v
    _mods = types.SimpleNamespace(models=ref_models, planner=ref_planner, dreamer=ref_dreamer,
                                  utils=ref_utils)
    return _mods


class NoiseTape:
    """FIFO replay of Gaussian draws."""

    def __init__(self, draws: List[torch.Tensor]):
        self.draws = list(draws)
        self.pos = 0

    def pop(self, shape, dtype=None):
        if self.pos >= len(self.draws):
            raise RuntimeError("noise tape exhausted")
        t = self.draws[self.pos]
        self.pos += 1
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"noise tape draw {self.pos - 1}: shape {tuple(t.shape)} "
                               f"!= requested {tuple(shape)}")
        return t if dtype is None else t.to(dtype)

    @contextlib.contextmanager
    def playing(self):
        import torch.distributions.normal as tdn
        o_randn, o_like, o_std = torch.randn, torch.randn_like, tdn._standard_normal

        def randn(*size, **kw):
            if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
                size = tuple(size[0])
            return self.pop(size, kw.get("dtype"))

        def randn_like(x, **kw):
            return self.pop(x.shape, x.dtype)

        def std_normal(shape, dtype, device):
            return self.pop(tuple(shape), dtype)

        torch.randn, torch.randn_like, tdn._standard_normal = randn, randn_like, std_normal
        try:
            yield self
        finally:
            torch.randn, torch.randn_like, tdn._standard_normal = o_randn, o_like, o_std
        if self.pos != len(self.draws):
            raise RuntimeError(f"noise tape: {len(self.draws) - self.pos} draws left unused")


# ----------------------------------------------------------------------------
# module construction
# ----------------------------------------------------------------------------
def build_modules(seed, Be, S, A, Hi, E, act="ELU", dtype=torch.float32):
    """Construct the reference's own modules with PyTorch default init
    (same positional args as src/planet.py:165-175, src/dreamer.py:30-50)."""
    m = load().models
    torch.manual_seed(seed)
    trans = m.TransitionModel(Be, S, A, Hi, E, act)
    reward = m.DenseModel(Be + S, Hi, activation=act)
    critic = m.DenseModel(Be + S, Hi, activation=act)
    actor = m.ActorModel(Be, S, Hi, A, act)
    mods = types.SimpleNamespace(transition=trans, reward=reward, critic=critic, actor=actor)
    if dtype != torch.float32:
        for mod in (trans, reward, critic, actor):
            mod.to(dtype)
        actor.raw_init_std = actor.raw_init_std.to(dtype)
    return mods


def fake_agent(mods, planning_horizon):
    """imagine_ahead / get_action only touch these attributes."""
    d = load().dreamer
    ns = types.SimpleNamespace(planning_horizon=planning_horizon, latent_distribution="Gaussian",
                               transition_model=mods.transition, actor=mods.actor)
    ns.get_action = lambda b, s, deterministic=False: d.Dreamer.get_action(ns, b, s, deterministic)
    return ns


# ----------------------------------------------------------------------------
# reference entry points with a noise tape
# ----------------------------------------------------------------------------
def imagine_tape(eps_a, eps_e, eps_s):
    draws = []
    for t in range(eps_a.size(0)):           # action, entropy, prior state
        draws += [eps_a[t], eps_e[t], eps_s[t]]
    return NoiseTape(draws)


def ref_imagine(mods, planning_horizon, prev_state, prev_belief, eps_a, eps_e, eps_s):
    d = load().dreamer
    agent = fake_agent(mods, planning_horizon)
    with imagine_tape(eps_a, eps_e, eps_s).playing():
        return d.Dreamer.imagine_ahead(agent, prev_state, prev_belief)


def ref_actor_loss(mods, planning_horizon, prev_state, prev_belief, eps_a, eps_e, eps_s,
                   discount=0.995, lambda_=0.95, entropy_weight=1e-5):
    """The behaviour-learning block exactly as src/dreamer.py:313-353 runs it
    (FreezeParameters included); returns loss, intermediates, actor grads."""
    r = load()
    d, u = r.dreamer, r.utils
    agent = fake_agent(mods, planning_horizon)
    model_modules = mods.transition.modules + [mods.reward]
    for p in mods.actor.parameters():
        p.grad = None
    with imagine_tape(eps_a, eps_e, eps_s).playing():
        with u.FreezeParameters(model_modules):
            beliefs, states, (means, stds), entropy = d.Dreamer.imagine_ahead(
                agent, prev_state.detach(), prev_belief.detach())
    with u.FreezeParameters(model_modules + [mods.critic]):
        reward = mods.reward(beliefs, states)
        value = mods.critic(beliefs, states)
    returns = d.lambda_return(reward, value, bootstrap=value[-1], discount=discount,
                              lambda_=lambda_)
    objective = returns
    if entropy_weight != -1:
        objective = objective + entropy_weight * entropy.unsqueeze(-1)
    loss = -objective.mean()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in mods.actor.named_parameters()}
    inter = dict(beliefs=beliefs, states=states, means=means, stds=stds, entropy=entropy,
                 reward=reward, value=value, returns=returns)
    return loss.detach(), {k: v.detach() for k, v in inter.items()}, grads


def ref_transition(mods, init_state, actions, init_belief, eps_prior, embeddings=None,
                   nonterminals=None, eps_post=None):
    draws = []
    for t in range(actions.size(0)):
        draws.append(eps_prior[t])
        if embeddings is not None:
            draws.append(eps_post[t])
    with NoiseTape(draws).playing():
        return mods.transition(init_state, actions, init_belief, embeddings, nonterminals)


def ref_cem(mods, action_size, planning_horizon, iters, candidates, top, belief, state,
            eps_act, eps_s):
    p = load().planner
    planner = p.MPCPlanner(action_size, planning_horizon, iters, candidates, top,
                           mods.transition, mods.reward)
    draws = []
    for it in range(iters):
        draws.append(eps_act[it])
        draws += [eps_s[it, h] for h in range(planning_horizon)]
    with NoiseTape(draws).playing(), torch.no_grad():
        return planner(belief, state)


def ref_lambda_return(*a, **k):
    return load().dreamer.lambda_return(*a, **k)


def ref_kl_loss(agent: str, posterior_params, prior_params, free_nats: float, kl_balance: float = -1):
    """The reference's own ``_kl_loss`` (agent = "planet": src/planet.py:288-308; "dreamer":
    src/dreamer.py:111-146) called on a stand-in ``self`` carrying the attributes it reads."""
    m = load()
    ns = types.SimpleNamespace(latent_distribution="Gaussian", kl_balance=kl_balance,
                               free_nats=torch.full((1,), free_nats, dtype=posterior_params[0].dtype))
    if agent == "planet":
        return sys.modules["planet"].Planet._kl_loss(ns, posterior_params, prior_params)
    ns._get_dist = types.MethodType(m.dreamer.Dreamer._get_dist, ns)
    return m.dreamer.Dreamer._kl_loss(ns, posterior_params, prior_params)
