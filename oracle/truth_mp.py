"""mpmath (50-digit) truth models.  TEST INFRASTRUCTURE ONLY.

``quad_ld_discretised``  the reference's *discretised* flux formula
    (src/jaxoplanet/core/limb_dark.py:50-184, order-node Gauss-Legendre with the
    float64 nodes/weights SciPy returns) evaluated without rounding error: the
    value both the reference and our kernels approximate.  |ours - this| vs
    |oracle - this| separates our error from the reference's own rounding noise.
``quad_ld_exact``        the same with the 1-D integral converged (mp.quad): the
    physical answer (what exoplanet_core / starry return), used with the
    reference's own 5e-7 tolerance.
``kepler_exact``         Newton solution of Kepler's equation.
"""

from __future__ import annotations

import mpmath as mp
import numpy as np
from scipy.special import roots_legendre

mp.mp.dps = 50


def kepler_exact(M, e):
    """Returns (E, sin f, cos f) as mpf for scalar M, e (M any real)."""
    M = mp.mpf(float(M))
    e = mp.mpf(float(e))
    two_pi = 2 * mp.pi
    Mr = M - two_pi * mp.floor(M / two_pi)
    E = Mr if e < 0.8 else mp.pi
    for _ in range(200):
        dE = (E - e * mp.sin(E) - Mr) / (1 - e * mp.cos(E))
        E -= dE
        if abs(dE) < mp.mpf(10) ** (-45):
            break
    sf = mp.sqrt(1 - e * e) * mp.sin(E) / (1 - e * mp.cos(E))
    cf = (mp.cos(E) - e) / (1 - e * mp.cos(E))
    return E, sf, cf


def _g_quad(u):
    u1, u2 = (mp.mpf(float(x)) for x in u)
    g0 = 1 - u1 - mp.mpf(3) / 2 * u2
    g1 = u1 + 2 * u2
    g2 = -u2 / 4
    norm = mp.pi * (g0 + g1 / mp.mpf("1.5"))
    return g0 / norm, g1 / norm, g2 / norm


def _kite(a, b, c):
    a, b, c = sorted([a, b, c])
    sq = (a + (b + c)) * (c - (a - b)) * (c + (a - b)) * (a + (b - c))
    return mp.sqrt(max(mp.mpf(0), sq))


def _solution(b, r, integrate_p):
    b = abs(mp.mpf(float(b)))
    r = abs(mp.mpf(float(r)))
    pi = mp.pi
    # callers handle b >= 1 + r and 1 + b <= r (no / full occultation)
    cond = (b > abs(1 - r)) and (b < 1 + r)
    area = _kite(r, b, mp.mpf(1)) if cond else mp.mpf(0)
    b2, r2 = b * b, r * r
    factor = (r - 1) * (r + 1)
    k0 = mp.atan2(area, b2 + factor)
    k1 = mp.atan2(area, b2 - factor)
    eta2 = r2 * (r2 + 2 * b2) / 2
    if b + r < 1:
        s0 = pi * (1 - r2)
        s2 = 2 * s0 + 4 * pi * (eta2 - mp.mpf(1) / 2)
    else:
        Alens = k1 + r2 * k0 - area / 2
        s0 = pi - Alens
        s2 = 2 * s0 + 2 * (-(pi - k1) + 2 * eta2 * k0 - area * (1 + 5 * r2 + b2) / 4)

    def integrand(phi):
        c = mp.cos(phi)
        omz2 = max(mp.mpf(0), r2 + b2 - 2 * b * r * c)
        z2 = 1 - omz2
        if z2 <= 0:
            z3 = mp.mpf(0)
        else:
            z3 = z2 * mp.sqrt(z2)
        if omz2 == 0:
            return mp.mpf(0)
        return 2 * r * (r - b * c) * (1 - z3) / (3 * omz2)

    P0 = integrate_p(integrand, k0)
    s1 = -P0 - 2 * (k1 - pi) / 3
    return s0, s1, s2


def _flux(u, b, r, integrate_p):
    if abs(float(b)) >= 1 + abs(float(r)):
        return mp.mpf(0)
    g0, g1, g2 = _g_quad(u)
    if 1 + abs(float(b)) <= abs(float(r)):
        # full occultation: s = [0, -2(kappa1 - pi)/3 with kappa1 = atan2(0, .), 0]
        b_ = abs(mp.mpf(float(b)))
        r_ = abs(mp.mpf(float(r)))
        k1 = mp.atan2(mp.mpf(0), b_ * b_ - (r_ - 1) * (r_ + 1))
        return -2 * (k1 - mp.pi) / 3 * g1 - 1
    s0, s1, s2 = _solution(b, r, integrate_p)
    return s0 * g0 + s1 * g1 + s2 * g2 - 1


def quad_ld_discretised(u, b, r, order=10):
    roots, weights = roots_legendre(order)
    roots = [mp.mpf(float(x)) for x in roots]
    weights = [mp.mpf(float(w)) for w in weights]

    def gl(f, k0):
        rng = k0 / 2
        return rng * mp.fsum(w * f(rng * x + k0 / 2) for x, w in zip(roots, weights))

    return float(_flux(u, b, r, gl))


def quad_ld_discretised_mpf(u, b, r, order=10):
    roots, weights = roots_legendre(order)
    roots = [mp.mpf(float(x)) for x in roots]
    weights = [mp.mpf(float(w)) for w in weights]

    def gl(f, k0):
        rng = k0 / 2
        return rng * mp.fsum(w * f(rng * x + k0 / 2) for x, w in zip(roots, weights))

    return _flux(u, b, r, gl)


def quad_ld_exact(u, b, r):
    def conv(f, k0):
        if k0 == 0:
            return mp.mpf(0)
        return mp.quad(f, np.linspace(0, 1, 9) * k0)

    return float(_flux(u, b, r, conv))
