"""hvae — B200-native (sm_100a) Poincare-ball VAE hot path behind the module API of
grisaitis/hyperbolic-vae (`hyperbolic_vae.{manifolds,layers,distributions}`).

Every op runs a hand-written CUDA kernel from libhvae_b200.so (C ABI, include/hvae_b200.h) through a
torch custom op with an analytic backward.  There is NO CPU path and no eager fallback: calling an op
on a non-CUDA tensor, or without the built library, raises.
"""
from . import _cabi  # noqa: F401
from . import ops  # noqa: F401
from . import manifolds, layers, distributions  # noqa: F401
from .manifolds import PoincareBall, PoincareBallWithExtras, ManifoldParameter  # noqa: F401

__version__ = "0.1.0"
