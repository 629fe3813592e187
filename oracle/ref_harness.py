"""Runs the UNMODIFIED reference (jgsimard/big-dreamer at /root/reference) with
injected noise.  TEST INFRASTRUCTURE ONLY -- used in this container to (a) pin
oracle/rssm_oracle.py against the real code and (b) generate tests/golden/*.
/root/reference does not exist on the GPU box; there the harness finds the unmodified
copy installed under baseline/_ref (see ``_find_reference_root``) or reports
``available() == False`` and the tests that need it skip.

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

def _find_reference_root() -> str:
    """BD_REFERENCE_ROOT, else the read-only tree of the build container, else the unmodified copy
    `pip install --target baseline/_ref` made of it (git-ignored; it travels to the GPU box)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [os.environ.get("BD_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "models.py")):
            return c
    return cands[0] or "/root/reference"


REFERENCE_ROOT = _find_reference_root()
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


def ref_get_action(mods, belief, state, eps_first, eps_entropy, deterministic=False):
    """Dreamer.get_action (src/dreamer.py:429-444) of the unmodified reference.  Draw order: the sample
    (``rsample``: (B,A); ``mode``: (100,B,A)) and then the entropy's (100,B,A)."""
    d = load().dreamer
    agent = fake_agent(mods, 2)
    with NoiseTape([eps_first, eps_entropy]).playing():
        return d.Dreamer.get_action(agent, belief, state, deterministic)


def ref_act_step(mods, belief, state, action, embedding, eps_prior, eps_post, eps_first, eps_entropy,
                 deterministic=False):
    """The transition / get_action part of Planet.update_belief_and_act (src/planet.py:379-390) run on the
    reference's own modules: posterior step with a time dimension of 1, then get_action."""
    d = load().dreamer
    agent = fake_agent(mods, 2)
    with NoiseTape([eps_prior, eps_post, eps_first, eps_entropy]).playing():
        b, _, _, post, _ = mods.transition(state, action.unsqueeze(dim=0), belief, embedding.unsqueeze(dim=0))
        b, post = b.squeeze(dim=0), post.squeeze(dim=0)
        a, _ = d.Dreamer.get_action(agent, b, post, deterministic)
    return b, post, a


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


def ref_actor_step(mods, planning_horizon, prev_state, prev_belief, discount=0.995, lambda_=0.95,
                   entropy_weight=1e-5):
    """The same block as ref_actor_loss through the reference's STOCK code path (it draws its own
    noise): what `bench.py --impl reference` times.  Returns the loss as a float."""
    r = load()
    d, u = r.dreamer, r.utils
    agent = fake_agent(mods, planning_horizon)
    model_modules = mods.transition.modules + [mods.reward]
    for p in mods.actor.parameters():
        p.grad = None
    with u.FreezeParameters(model_modules):
        beliefs, states, _, entropy = d.Dreamer.imagine_ahead(agent, prev_state.detach(), prev_belief.detach())
    with u.FreezeParameters(model_modules + [mods.critic]):
        reward = mods.reward(beliefs, states)
        value = mods.critic(beliefs, states)
    returns = d.lambda_return(reward, value, bootstrap=value[-1], discount=discount, lambda_=lambda_)
    loss = -(returns + entropy_weight * entropy.unsqueeze(-1)).mean()
    loss.backward()
    return float(loss.detach())


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


def ref_value_update(mods, beliefs, states, target, discount=None):
    """The critic regression block exactly as src/dreamer.py:369-391 runs it (value_dist =
    Normal(critic(b, s), 1); -log_prob(target).mean(), optionally weighted by the cumulated discount);
    returns the loss and the critic's parameter gradients."""
    from torch.distributions import Normal
    for p in mods.critic.parameters():
        p.grad = None
    value_dist = Normal(mods.critic(beliefs.detach(), states.detach()), 1)
    if discount is not None:
        value_loss = -(discount.detach() * value_dist.log_prob(target.detach())).mean()
    else:
        value_loss = -value_dist.log_prob(target.detach()).mean()
    value_loss.backward()
    return value_loss.detach(), {k: p.grad.detach().clone() for k, p in mods.critic.named_parameters()}


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


# ----------------------------------------------------------------------------
# whole-agent harness (SURVEY 8c "whole-agent oracle"): the reference's own Dreamer(params, env)
# ----------------------------------------------------------------------------
def default_params() -> dict:
    """The reference's hydra config as the plain dict `my_app` hands to the agents
    (src/main.py:27-29).  Read from src/conf/config.yaml when the tree has it (PyYAML reads
    `2e-4` as a string where omegaconf gave a float: coerced here); the pip-installed copy under
    baseline/_ref carries no yaml, so the same defaults are restated below (src/conf/config.yaml:1-81)."""
    path = os.path.join(REFERENCE_ROOT, "src", "conf", "config.yaml")
    if os.path.isfile(path):
        import yaml

        def coerce(d):
            for k, v in list(d.items()):
                if isinstance(v, dict):
                    coerce(v)
                elif isinstance(v, str):
                    try:
                        d[k] = float(v)
                    except ValueError:
                        pass
            return d
        return coerce(yaml.safe_load(open(path)))
    return dict(
        algorithm="dreamer", exp_name="default", seed=0, disable_cuda=False, env="Pendulum-v0",
        max_episode_length=1000, experience_size=1000000, cnn_activation_function="ELU",
        dense_activation_function="ELU", embedding_size=1024, hidden_size=200, n_layers=4,
        belief_size=200, state_size=30, action_repeat=2, action_noise=0.3, episodes=1000,
        seed_episodes=1, seed_steps=5000, train_steps=1000000, collect_interval=5, batch_size=50,
        seq_len=50, free_nats=3.0, bit_depth=5, model_learning_rate=2e-4, adam_epsilon=1e-5,
        weight_decay=1e-6, grad_clip_norm=100.0, planning_horizon=15, discount=0.995, disclam=0.95,
        test=False, test_interval=25, test_episodes=10, checkpoint_interval=5,
        checkpoint_experience=False, models="", experience_replay="", render=False, log_freq=100,
        log_video_freq=-1, wandb_project=None, fps=10, kl_balance=0.8, kl_loss_weight=0.1,
        latent_distribution="Gaussian", discrete_latent_dimensions=32, discrete_latent_classes=32,
        action_distribution="Gaussian", jit=False,
        MPC=dict(optimisation_iters=10, candidates=1000, top_candidates=100),
        ActorCritic=dict(actor_learning_rate=4e-5, value_learning_rate=1e-4, entropy_weight=1e-5,
                         slow_critic_update_interval=100, polyak_avg=1.0, gradient_mixing=-1),
        use_discount=False, discount_weight=5.0, pixel_observation=True,
        environment_steps_per_update=10)


class FakeEnv:
    """What Planet.__init__ / the replay buffer read off an env (src/planet.py:33-77)."""

    def __init__(self, action_size=2, observation_size=(3, 64, 64)):
        self.action_size, self.observation_size = action_size, observation_size

    def reset(self):
        return torch.zeros(1, *self.observation_size)

    def step(self, action):
        return torch.zeros(1, *self.observation_size), 0.0, False

    def sample_random_action(self):
        return torch.zeros(self.action_size)

    def close(self):
        pass


@contextlib.contextmanager
def agent_modules():
    """sys.path / sys.modules set up so `import planet, dreamer, models, planner` gives the
    reference's modules (with the import shims); restored on exit."""
    load()
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k, None) for k in ("typeguard", "torchtyping", "plotly", "gym")}
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "src"))
    sys.path.insert(0, _SHIMS)
    try:
        import dreamer
        import models
        import planet
        import planner
        yield types.SimpleNamespace(planet=planet, dreamer=dreamer, models=models, planner=planner)
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v


def fill_buffer(agent, steps: int, seed: int = 0):
    """Synthetic experience through the reference's own buffer.append (src/memory.py:33-49)."""
    g = torch.Generator().manual_seed(seed)
    A = agent.env.action_size if hasattr(agent, "env") else agent.buffer.actions.shape[1]
    obs_shape = agent.buffer.observations.shape[1:]
    for i in range(steps):
        obs = torch.rand(1, *obs_shape, generator=g) - 0.5
        act = torch.rand(A, generator=g) * 2 - 1
        agent.buffer.append(obs, act, float(torch.randn((), generator=g)), (i + 1) % 37 == 0)
