"""Swap the B200 path into the UNMODIFIED reference without editing it.

The reference's agents bind names at import time (``from models import ...`` in
src/planet.py:15-17 and src/dreamer.py:13) and construct the models in
``Planet.initialize_models`` (src/planet.py:161-223), so ``patch()`` must run
after the reference modules are importable and BEFORE agents are constructed:

    sys.path.insert(0, "<reference>/src")
    import big_dreamer_b200 as bd
    bd.patch()
    agent = dreamer.Dreamer(params, env)      # now builds B200-backed models
"""
from __future__ import annotations

import importlib
import sys

from . import modules as M

_saved = {}


def _get(name):
    return sys.modules.get(name) or importlib.import_module(name)


def patch(fused: bool = False) -> None:
    """Rebind models.TransitionModel / DenseModel, planner.MPCPlanner (and the copies
    ``planet`` / ``dreamer`` took with ``from ... import``), Dreamer.imagine_ahead,
    dreamer.lambda_return and Planet/Dreamer._kl_loss.

    fused=True (SURVEY 8b level L2): ``Dreamer.imagine_ahead`` runs the fused imagine + reward/value
    heads + lambda_return kernels (one forward launch chain, one backward) and the unmodified
    ``train_step``'s following ``reward_model(b, s)`` / ``critic_target(b, s)`` / ``lambda_return(...)``
    calls pick up those results (modules.imagine_ahead_fused); ``Dreamer.imagine_and_returns`` is
    also added as a method."""
    if _saved:
        return
    models, planner = _get("models"), _get("planner")
    planet, dreamer = _get("planet"), _get("dreamer")

    def swap(mod, attr, new):
        if hasattr(mod, attr):
            _saved[(mod.__name__, attr)] = (mod, getattr(mod, attr))
            setattr(mod, attr, new)

    # Categorical (DreamerV2) latents are not on the B200 path: the swapped-in name is a factory that
    # hands those configurations to the reference's own class (config.yaml:56, src/models.py:166-181)
    ref_tm = models.TransitionModel

    class TransitionModel(M.TransitionModel):
        def __new__(cls, *args, **kwargs):
            names = ("belief_size", "state_size", "action_size", "hidden_size", "embedding_size",
                     "activation", "min_std_dev", "latent_distribution")
            bound = dict(zip(names, args))
            bound.update(kwargs)
            if bound.get("latent_distribution", "Gaussian") != "Gaussian":
                return ref_tm(*args, **kwargs)       # not an instance of cls: __init__ is not re-run
            return super().__new__(cls)
    TransitionModel.__name__ = TransitionModel.__qualname__ = "TransitionModel"

    for mod in (models, planet, dreamer):
        swap(mod, "TransitionModel", TransitionModel)
        swap(mod, "DenseModel", M.DenseModel)
    for mod in (planner, planet, dreamer):
        swap(mod, "MPCPlanner", M.MPCPlanner)
    swap(dreamer, "lambda_return", M.lambda_return)
    cls = dreamer.Dreamer
    ref_imagine = cls.imagine_ahead
    _saved[("dreamer.Dreamer", "imagine_ahead")] = (cls, ref_imagine)

    def imagine_ahead(self, prev_state, prev_belief, *args, **kwargs):
        # Categorical latents / a Categorical actor / a reference (non-B200) transition model keep
        # the reference's own Python loop
        if (getattr(self, "latent_distribution", "Gaussian") != "Gaussian"
                or getattr(self.actor, "action_distribution", "Gaussian") != "Gaussian"
                or not isinstance(self.transition_model, M.TransitionModel)):
            return ref_imagine(self, prev_state, prev_belief)
        if fused:
            return M.imagine_ahead_fused(self, prev_state, prev_belief, *args, **kwargs)
        return M.imagine_ahead(self, prev_state, prev_belief, *args, **kwargs)
    cls.imagine_ahead = imagine_ahead
    # dynamics-update KL (Gaussian latents; the Categorical path keeps the reference's method)
    for owner in (planet.Planet, cls):
        if "_kl_loss" in vars(owner):
            orig = owner._kl_loss
            _saved[(owner.__module__ + "." + owner.__name__, "_kl_loss")] = (owner, orig)

            def _kl(self, posterior_params, prior_params, _orig=orig):
                if getattr(self, "latent_distribution", "Gaussian") != "Gaussian":
                    return _orig(self, posterior_params, prior_params)
                return M._kl_loss_method(self, posterior_params, prior_params)
            owner._kl_loss = _kl
    if fused:
        _saved[("dreamer.Dreamer", "imagine_and_returns")] = (cls, getattr(cls, "imagine_and_returns", None))
        cls.imagine_and_returns = M.imagine_and_returns


def unpatch() -> None:
    for (_, attr), (owner, old) in list(_saved.items()):
        if old is None:
            if hasattr(owner, attr):
                delattr(owner, attr)
        else:
            setattr(owner, attr, old)
    _saved.clear()
    M._fused_record = None
