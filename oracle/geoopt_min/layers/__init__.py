from . import stereographic  # noqa: F401
