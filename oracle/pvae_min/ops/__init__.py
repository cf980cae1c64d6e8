from . import manifold_layers  # noqa: F401
