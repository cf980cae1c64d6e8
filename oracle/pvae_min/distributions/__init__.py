from .hyperspherical_uniform import HypersphericalUniform  # noqa: F401
from .hyperbolic_radius import HyperbolicRadius  # noqa: F401
from .riemannian_normal import RiemannianNormal  # noqa: F401
from .wrapped_normal import WrappedNormal  # noqa: F401
