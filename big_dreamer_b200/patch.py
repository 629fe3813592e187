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
    dreamer.lambda_return and Planet/Dreamer._kl_loss."""
    if _saved:
        return
    models, planner = _get("models"), _get("planner")
    planet, dreamer = _get("planet"), _get("dreamer")

    def swap(mod, attr, new):
        if hasattr(mod, attr):
            _saved[(mod.__name__, attr)] = (mod, getattr(mod, attr))
            setattr(mod, attr, new)

    for mod in (models, planet, dreamer):
        swap(mod, "TransitionModel", M.TransitionModel)
        swap(mod, "DenseModel", M.DenseModel)
    for mod in (planner, planet, dreamer):
        swap(mod, "MPCPlanner", M.MPCPlanner)
    swap(dreamer, "lambda_return", M.lambda_return)
    cls = dreamer.Dreamer
    _saved[("dreamer.Dreamer", "imagine_ahead")] = (cls, cls.imagine_ahead)
    cls.imagine_ahead = M.imagine_ahead
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
        cls.imagine_and_returns = M.imagine_and_returns


def unpatch() -> None:
    for (_, attr), (owner, old) in list(_saved.items()):
        setattr(owner, attr, old)
    _saved.clear()
