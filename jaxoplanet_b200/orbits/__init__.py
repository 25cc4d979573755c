# mirrors src/jaxoplanet/orbits/__init__.py:1-3
from jaxoplanet_b200.orbits import keplerian as keplerian
from jaxoplanet_b200.orbits.transit import TransitOrbit as TransitOrbit
from jaxoplanet_b200.orbits.ttv import TTVOrbit as TTVOrbit
