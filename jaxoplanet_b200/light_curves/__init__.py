# mirrors src/jaxoplanet/light_curves/__init__.py:1-4
__all__ = ["transforms", "limb_dark_light_curve"]

from jaxoplanet_b200.light_curves import transforms as transforms
from jaxoplanet_b200.light_curves.limb_dark import light_curve as limb_dark_light_curve
