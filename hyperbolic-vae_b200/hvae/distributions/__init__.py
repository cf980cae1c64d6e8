from .wrapped_normal import WrappedNormal  # noqa: F401
