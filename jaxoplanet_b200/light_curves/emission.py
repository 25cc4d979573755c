"""``light_curve(system, order=10)(t)``: drop-in for src/jaxoplanet/light_curves/emission.py:11-52 --
the flux of a star AND of its (uniformly bright, limb-darkened) companions: transits of the bodies in
front of the star plus the occultation of every body behind it.

The reference takes a ``starry`` ``SurfaceSystem`` (spherical-harmonic maps, out of this path's scope,
SURVEY.md 2); what this light curve reads of it is only ``central_surface.u``, ``body_surfaces[i].u`` and
``body_surfaces[i].amplitude``.  ``Surface`` / ``SurfaceSystem`` below carry exactly that, with the
reference's constructor and ``add_body`` signatures (src/jaxoplanet/starry/orbit.py:13-69).

Kernels: the transits are the fused system kernel (K3); the occultations take the sky positions from the
orbit-state kernel (K8) and the flux from K2 with the roles of the two discs swapped
(``b = sqrt(x^2 + y^2) / R_body``, ``r = R_star / R_body``, emission.py:41-43).
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from jaxoplanet_b200 import _array as A
from jaxoplanet_b200.core.limb_dark import light_curve as _limb_dark_light_curve
from jaxoplanet_b200.light_curves.limb_dark import light_curve as _transit_light_curve
from jaxoplanet_b200.orbits.keplerian import Body, Central, OrbitalBody, System

__all__ = ["Surface", "SurfaceSystem", "light_curve"]


class Surface:
    """The two attributes of ``jaxoplanet.starry.surface.Surface`` this light curve reads:
    limb-darkening coefficients ``u`` and the ``amplitude`` of the (uniform) map."""

    def __init__(self, *, u=(), amplitude=1.0):
        self.u = tuple(u)
        self.amplitude = amplitude


class SurfaceSystem(System):
    """A Keplerian ``System`` whose central body and bodies carry a ``Surface``
    (src/jaxoplanet/starry/orbit.py:13-69)."""

    def __init__(self, central: Central | None = None, central_surface: Surface | None = None, *,
                 bodies=()):
        central = Central() if central is None else central
        self.central_surface = Surface() if central_surface is None else central_surface
        orbital, surfaces = [], []
        for body, surface in bodies:
            orbital.append(body if isinstance(body, OrbitalBody) else OrbitalBody(central, body))
            surfaces.append(getattr(body, "surface", None) if surface is None else surface)
        super().__init__(central, bodies=orbital)
        self._body_surfaces = tuple(surfaces)

    @property
    def body_surfaces(self):
        return self._body_surfaces

    def add_body(self, body: Body | None = None, surface: Surface | None = None, **kwargs: Any):
        if body is None:
            body = Body(**kwargs)
        if surface is None:
            surface = getattr(body, "surface", None)
        return SurfaceSystem(central=self.central, central_surface=self.central_surface,
                             bodies=list(zip(self.bodies, self.body_surfaces)) + [(body, surface)])


def light_curve(system: SurfaceSystem, order: int = 10):
    """-> function of time returning ``t.shape + (1 + n_bodies,)``: the star's flux
    ``1 + sum of transits`` followed by ``amplitude_i (1 + occultation_i)`` per body."""
    assert isinstance(system, SurfaceSystem), f"Expected an instance of 'SurfaceSystem', but got {type(system)}"
    if A.batch_ndim() and any(b.period.ndim == 1 for b in system.bodies):
        raise NotImplementedError("emission light curves of batched systems")
    transit = _transit_light_curve(system, *system.central_surface.u, order=order)

    def light_curve_impl(time):
        dt = A.result_dtype(time)
        t_dev = A.to_dev(time, dt)
        # emission.py:30-38: transits in front of the star, masked where z <= 0 by the fused kernel
        star = A.to_dev(transit(t_dev), dt).sum(-1) + 1
        r_star = system.central.radius
        cols = [star]
        for body, surface in zip(system.bodies, system.body_surfaces):
            if surface is None:
                raise ValueError("every body of an emission light curve needs a Surface")
            # emission.py:40-49: the star in front of the body (z < 0), discs swapped
            x, y, z = (A.to_dev(v, dt) for v in body.relative_position(t_dev))
            r_body = body.radius.to(device=x.device, dtype=dt)
            b = torch.sqrt(x**2 + y**2) / r_body
            r = (r_star.to(device=x.device, dtype=dt) / r_body).expand(b.shape).contiguous()
            u = np.asarray([float(v) for v in surface.u], dtype=np.float64)
            lc = A.to_dev(_limb_dark_light_curve(A.to_dev(u, dt), b.contiguous(), r, order=order), dt)
            cols.append(float(surface.amplitude) * (1 + torch.where(z < 0, lc, torch.zeros_like(lc))))
        return A.like_input(torch.stack(cols, dim=-1), time)

    return light_curve_impl
