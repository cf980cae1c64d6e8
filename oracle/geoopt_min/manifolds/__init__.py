from .stereographic import PoincareBall, Stereographic, Manifold  # noqa: F401
from . import stereographic  # noqa: F401
