"""ORACLE: import the reference's OWN modules, unmodified, from /root/reference over the shims.

Works only where /root/reference exists (the authoring container).  Used by
tests/golden/make_golden.py to mint fixtures and by tests that pin oracle/ref_port.py against the
reference's real files.  Nothing that runs on the GPU box may call this.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("HVAE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hyperbolic_vae", "layers.py"))


def _alias_package(real_name: str, alias: str):
    """Register every loaded submodule of `real_name` under `alias` in sys.modules."""
    importlib.import_module(real_name)
    for name, mod in list(sys.modules.items()):
        if name == real_name or name.startswith(real_name + "."):
            sys.modules[alias + name[len(real_name):]] = mod


def install_shims():
    from . import stubs

    stubs.install()
    if "geoopt" not in sys.modules:
        # make sure all submodules are loaded before aliasing
        for sub in ("", ".manifolds", ".manifolds.stereographic", ".manifolds.stereographic.math",
                    ".manifolds.stereographic.manifold", ".layers", ".layers.stereographic", ".optim",
                    ".utils", ".tensor"):
            importlib.import_module("oracle.geoopt_min" + sub)
        _alias_package("oracle.geoopt_min", "geoopt")
    if "pvae" not in sys.modules:
        for sub in ("", ".utils", ".manifolds", ".distributions", ".distributions.hyperbolic_radius",
                    ".distributions.hyperspherical_uniform", ".distributions.ars",
                    ".distributions.riemannian_normal", ".distributions.wrapped_normal", ".ops",
                    ".ops.manifold_layers"):
            importlib.import_module("oracle.pvae_min" + sub)
        _alias_package("oracle.pvae_min", "pvae")


def load():
    """Return the reference package `hyperbolic_vae` (its own files, executed verbatim)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import logging

    logging.getLogger("hyperbolic_vae").setLevel(logging.ERROR)
    return importlib.import_module("hyperbolic_vae")


def load_module(name: str):
    load()
    return importlib.import_module("hyperbolic_vae." + name)
