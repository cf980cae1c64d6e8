"""ORACLE: `geoopt` stand-in (see oracle/__init__.py).  Exposes exactly the names the reference
imports: geoopt.{PoincareBall,Stereographic,Manifold,ManifoldTensor,ManifoldParameter,manifolds,
layers,optim,utils}."""
from . import manifolds, layers, optim, utils  # noqa: F401
from .tensor import ManifoldTensor, ManifoldParameter  # noqa: F401
from .manifolds import PoincareBall, Stereographic, Manifold  # noqa: F401
