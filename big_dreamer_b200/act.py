"""Acting path: what the agent runs once per environment step.

    get_action   <- Dreamer.get_action, src/dreamer.py:429-444 (the action; see below for the entropy)
    ActPath      <- Planet.update_belief_and_act, src/planet.py:370-403, between the encoder and env.step:
                    one posterior step of the transition model + the policy (CEM planner for PlaNet,
                    actor for Dreamer) + the exploration noise, replayed as ONE CUDA graph per step

The encoder (convolutions), the environment and the replay buffer stay the reference's own code: the
caller hands in the observation embedding and gets (belief, posterior_state, action) back.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
from torch import Tensor

from . import _lib
from . import functions as F_
from . import modules as M
from ._lib import BdError
from .graph import CapturedStep


def get_action(self, belief: Tensor, state: Tensor, deterministic: bool = False,
               noise: Optional[Tensor] = None):
    """Drop-in for ``Dreamer.get_action`` as the acting loop uses it (bind as a method; ``self`` needs
    ``actor``).  belief (B,Be), state (B,S) -> (action (B,A), None).

    deterministic=False: ``dist.rsample()`` = tanh(mean + eps * std); deterministic=True:
    ``SampleDist.mode()`` (of 100 samples the most probable, src/models.py:707-723).  ``noise`` is that
    draw -- (B,A) or (100,B,A) -- and is drawn with torch.randn when absent.  The reference also returns a
    100-sample entropy estimate which its only caller on this path discards (src/planet.py:390); it is not
    computed here (None) -- imagine_ahead, the caller that needs it, has its own fused implementation."""
    actor = self.actor if hasattr(self, "actor") else self
    if getattr(actor, "action_distribution", "Gaussian") != "Gaussian":
        raise NotImplementedError("get_action: only the Gaussian actor runs on the B200 path")
    lib = _lib.load()
    cfg = M.actor_config(actor)
    lin = M._linears(actor.model)
    with torch.no_grad():
        raw = F_.mlp_apply(cfg["act_id"], belief, state, [l.weight for l in lin], [l.bias for l in lin])
    raw = F_._f32c(raw)
    rows, A = raw.shape[0], raw.shape[1] // 2
    J = cfg["entropy_samples"]
    shape = (J, rows, A) if deterministic else (rows, A)
    if noise is None:
        noise = torch.randn(*shape, device=raw.device, dtype=torch.float32)
    noise = F_._f32c(noise)
    if tuple(noise.shape) != shape:
        raise BdError(f"get_action: noise must have shape {shape}, got {tuple(noise.shape)}")
    action = torch.empty(rows, A, device=raw.device, dtype=torch.float32)
    ccfg = _lib.ActorCfg(cfg["mean_scale"], cfg["raw_init_std"], cfg["min_std"], J)
    _lib.check(lib.bd_actor_act(_lib.ptr(raw), _lib.ptr(noise), C.byref(ccfg), rows, A, int(bool(deterministic)),
                                _lib.ptr(action), _lib.stream_ptr()), "bd_actor_act")
    return action, None


class ActPath:
    """One environment step of the acting loop behind one CUDA-graph replay.

        act = bd.ActPath(transition_model, policy, batch=1, action_noise=0.3)
        belief, posterior_state, action = act(belief, posterior_state, action, embedding, explore=True)

    is ``Planet.update_belief_and_act`` (src/planet.py:370-403) from ``self.transition_model(...)`` to the
    exploration noise: ``policy`` is an ``MPCPlanner`` (PlaNet: ``self.planner(belief, state)``) or an actor
    module (Dreamer: ``get_action``).  All Gaussian draws happen inside the graph (torch's graph-safe
    generator), so every replay sees fresh noise.  The modules' parameters are read in place at replay
    time: optimizer steps between environment steps are seen without re-capturing.  ``noise=`` runs the
    same arithmetic eagerly with explicit draws (parity tests)."""

    def __init__(self, transition_model, policy, batch: int = 1, action_noise: float = 0.3,
                 deterministic: bool = False, device=None):
        self.tm, self.policy = transition_model, policy
        self.is_planner = isinstance(policy, M.MPCPlanner) or hasattr(policy, "optimisation_iters")
        self.batch, self.action_noise, self.deterministic = batch, float(action_noise), deterministic
        d = M.rssm_dims(transition_model)
        self.dims = d
        dev = device if device is not None else next(transition_model.parameters()).device
        if torch.device(dev).type != "cuda":
            raise BdError("ActPath runs on a CUDA device only (no CPU fallback)")
        z = lambda n: torch.zeros(batch, n, device=dev, dtype=torch.float32)
        self._in = [z(d["Be"]), z(d["S"]), z(d["A"]), z(d["E"])]
        self._graphs: Dict[bool, CapturedStep] = {}

    def _step(self, belief, state, action, embedding, explore: bool, noise: Optional[Dict[str, Tensor]] = None):
        nz = noise or {}
        tn = None
        if noise is not None:
            tn = dict(eps_prior=nz["eps_prior"][None], eps_post=nz["eps_post"][None])
        with torch.no_grad():
            beliefs, _, _, post, _ = self.tm(state, action[None], belief, embedding[None], None, noise=tn)
            b, s = beliefs[0], post[0]
            if self.is_planner:
                a = self.policy(b, s, noise=nz["planner"]) if "planner" in nz else self.policy(b, s)
            else:
                a, _ = get_action(self.policy, b, s, self.deterministic, nz.get("eps_act"))
            if explore:     # torch.clamp(Normal(action, action_noise).rsample(), -1, 1), src/planet.py:392-395
                e = nz["eps_explore"] if "eps_explore" in nz else torch.randn_like(a)
                a = torch.clamp(a + self.action_noise * e, -1.0, 1.0)
        return b, s, a

    def __call__(self, belief: Tensor, posterior_state: Tensor, action: Tensor, embedding: Tensor,
                 explore: bool = False, noise: Optional[Dict[str, Tensor]] = None):
        if noise is not None:
            return self._step(F_._f32c(belief), F_._f32c(posterior_state), F_._f32c(action), F_._f32c(embedding),
                              explore, noise)
        g = self._graphs.get(bool(explore))
        if g is None:
            g = CapturedStep(lambda b, s, a, e: self._step(b, s, a, e, bool(explore)), self._in)
            self._graphs[bool(explore)] = g
        return g(belief, posterior_state, action, embedding)
