from .wrapped_normal import WrappedNormal  # noqa: F401
from .riemannian_normal import HyperbolicRadius, HypersphericalUniform, RiemannianNormal  # noqa: F401
