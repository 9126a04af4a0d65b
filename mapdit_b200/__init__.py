"""mapdit_b200 — B200-native (sm_100a) implementation of the MaP-DiT hot path.

Drop-in surface of ericbill21/map-dit for the path named in BASELINE.json:
    DIT_MODELS / DiT_models, DiT.forward(x, t, y), DiT.forward_with_cfg, create_diffusion(...)
    -> training_losses / p_sample_loop / p_sample / q_sample with respacing.
Everything numerical runs in hand-written CUDA kernels (libmapdit.so, C ABI in include/mapdit.h).
"""
from .models import DIT_MODELS, DiT_models, get_model  # noqa: F401
from .dit import DiT  # noqa: F401
from .diffusion import create_diffusion  # noqa: F401

__all__ = ["DIT_MODELS", "DiT_models", "get_model", "DiT", "create_diffusion"]
