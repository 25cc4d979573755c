"""jaxoplanet_b200: B200-native (sm_100a) drop-in for jaxoplanet's hot path.

Mirrors the reference's public surface for that path only:
``core.kepler``, ``core.light_curve`` (``core.limb_dark.light_curve``),
``light_curves.limb_dark_light_curve``, ``light_curves.transforms.integrate`` and
``orbits.keplerian.{Central, Body, OrbitalBody, System}``.  All numerics run in the CUDA
library ``lib/libjxp_b200.so`` behind the C ABI of ``include/jaxoplanet_b200.h``.
"""

from jaxoplanet_b200 import core as core
from jaxoplanet_b200 import light_curves as light_curves
from jaxoplanet_b200 import orbits as orbits
from jaxoplanet_b200._array import batched as batched

__version__ = "0.1.0"
__all__ = ["core", "light_curves", "orbits", "batched"]
