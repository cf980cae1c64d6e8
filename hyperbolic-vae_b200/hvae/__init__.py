"""hvae — B200-native (sm_100a) Poincare-ball VAE hot path behind the module API of
grisaitis/hyperbolic-vae (`hyperbolic_vae.{manifolds,layers,distributions}`).

Every op runs a hand-written CUDA kernel from libhvae_b200.so (C ABI, include/hvae_b200.h) through a
torch custom op with an analytic backward.  There is NO CPU path and no eager fallback: calling an op
on a non-CUDA tensor, or without the built library, raises.
"""
import torch as _torch

# Precision contract of the fp32 mode (BASELINE.json: 1e-5 relative against the reference's CPU fp32 path): the parts of
# the reference's graphs this package leaves to cuDNN / cuBLAS (the conv trunk of models/vae_hyperbolic.py, small Linear
# layers) must not run in TF32 (10-bit mantissa, ~1e-3) - torch enables TF32 for cuDNN convolutions by default.
_torch.backends.cudnn.allow_tf32 = False
_torch.backends.cuda.matmul.allow_tf32 = False

from . import _cabi  # noqa: F401,E402
from . import ops  # noqa: F401,E402
from . import manifolds, layers, distributions, optim  # noqa: F401,E402
from .manifolds import PoincareBall, PoincareBallWithExtras, ManifoldParameter  # noqa: F401,E402

__version__ = "0.1.0"
