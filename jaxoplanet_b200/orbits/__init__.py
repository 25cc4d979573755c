# mirrors src/jaxoplanet/orbits/__init__.py:1 (TransitOrbit / TTVOrbit are out of scope)
from jaxoplanet_b200.orbits import keplerian as keplerian
