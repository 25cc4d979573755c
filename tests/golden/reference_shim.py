"""Import the reference's own source files from /root/reference on top of the NumPy-backed jax shim.

TEST INFRASTRUCTURE ONLY (see jax_shim/jax/__init__.py).  `/root/reference` exists only in the build
container, so callers must skip when `available()` is False; the vectors this produces are committed
as tests/golden/reference_v1.npz by make_reference_vectors.py.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("JXP_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "jaxoplanet"))


def load():
    """Returns a namespace with the reference modules on the hot path (imported once)."""
    if not available():
        raise RuntimeError(f"reference checkout not found under {REF_SRC}")
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__file__", "").startswith(HERE):
        raise RuntimeError("a real jax is already imported; use it directly instead of the shim")
    shim = os.path.join(HERE, "jax_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    if "jaxoplanet" not in sys.modules:
        # a bare package object: the reference's __init__ pulls in starry and a VCS-generated version file
        pkg = types.ModuleType("jaxoplanet")
        pkg.__path__ = [os.path.join(REF_SRC, "jaxoplanet")]
        sys.modules["jaxoplanet"] = pkg
        for sub in ("light_curves", "orbits", "core"):
            m = types.ModuleType(f"jaxoplanet.{sub}")
            m.__path__ = [os.path.join(REF_SRC, "jaxoplanet", sub)]
            sys.modules[f"jaxoplanet.{sub}"] = m
            setattr(pkg, sub, m)
    ns = types.SimpleNamespace()
    if not hasattr(sys.modules["jaxoplanet.orbits"], "TransitOrbit"):
        # ttv.py does `from jaxoplanet.orbits import TransitOrbit` (normally set by the package __init__)
        sys.modules["jaxoplanet.orbits"].TransitOrbit = importlib.import_module("jaxoplanet.orbits.transit").TransitOrbit
    for name, mod in (
        ("kepler", "jaxoplanet.core.kepler"),
        ("limb_dark", "jaxoplanet.core.limb_dark"),
        ("utils", "jaxoplanet.utils"),
        ("keplerian", "jaxoplanet.orbits.keplerian"),
        ("transit", "jaxoplanet.orbits.transit"),
        ("ttv", "jaxoplanet.orbits.ttv"),
        ("lc_limb_dark", "jaxoplanet.light_curves.limb_dark"),
        ("transforms", "jaxoplanet.light_curves.transforms"),
    ):
        setattr(ns, name, importlib.import_module(mod))
    return ns
