"""ORACLE: `pvae` stand-in (emilemathieu/pvae @ c04ec2149fc4d37fd83946a366780816c0cbe3c0, the commit
cited at /root/reference/hyperbolic_vae/layers.py:134).  pvae is not a declared dependency of the
reference (pyproject.toml:28 is commented out) and is absent here; this restates the published
algorithm per SURVEY.md Appendix A.2.  PARITY STATUS: unpinned third-party arithmetic."""
from . import utils, manifolds, distributions, ops  # noqa: F401
