# mirrors src/jaxoplanet/orbits/__init__.py:1-2 (TTVOrbit is out of scope)
from jaxoplanet_b200.orbits import keplerian as keplerian
from jaxoplanet_b200.orbits.transit import TransitOrbit as TransitOrbit
