from . import math  # noqa: F401
from .manifold import PoincareBall, Stereographic, Manifold  # noqa: F401
