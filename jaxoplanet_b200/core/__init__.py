# mirrors src/jaxoplanet/core/__init__.py:1-2
from jaxoplanet_b200.core.kepler import kepler as kepler
from jaxoplanet_b200.core.limb_dark import light_curve as light_curve
