"""big_dreamer_b200 -- B200 (sm_100a) implementation of big-dreamer's RSSM
latent-dynamics hot path behind the reference's own Python API.

    import big_dreamer_b200 as bd
    bd.patch()            # rebind the reference's classes/functions (see patch.py)

Compute lives in libbd_b200.so (hand-written CUDA behind the C ABI in
include/bd_b200.h); this package is the thin host side.
"""
from ._lib import BdError, LIB_PATH, load as load_library
from .functions import get_precision, set_precision
from .modules import (DenseModel, MPCPlanner, TransitionModel, build_mlp, draw_imagine_noise, heads_pair,
                      imagine_ahead, imagine_and_returns, kl_loss, lambda_return, value_update)
from .patch import patch, unpatch
from .graph import CapturedStep
from .act import ActPath, get_action

__all__ = ["BdError", "LIB_PATH", "load_library", "get_precision", "set_precision", "DenseModel",
           "MPCPlanner", "TransitionModel", "build_mlp", "draw_imagine_noise", "heads_pair", "imagine_ahead",
           "imagine_and_returns", "kl_loss", "lambda_return", "value_update", "patch", "unpatch", "CapturedStep",
           "ActPath", "get_action"]
__version__ = "0.1.0"
