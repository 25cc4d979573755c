"""``TransitOrbit``: an orbit parameterised by the transit signal (period, duration or speed,
t0, b, r) -- host-side mirror of src/jaxoplanet/orbits/transit.py:7-107.  There is no Kepler solve
on this path (linear motion across the disk); ``limb_dark_light_curve(TransitOrbit(...), u)(t)``
evaluates the positions here with torch element-wise ops and the flux in the K2 CUDA kernel
(through the ``LightCurveOrbit`` protocol, src/jaxoplanet/proto.py:8-20)."""

from __future__ import annotations

import numpy as np
import torch

from jaxoplanet_b200 import _array as A

__all__ = ["TransitOrbit"]


def _t(x):
    if isinstance(x, torch.Tensor):
        return x if x.dtype.is_floating_point else x.to(torch.float64)
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


class TransitOrbit:
    """Args: period, duration | speed (one required), time_transit (0), impact_param (0),
    radius_ratio (0).  ``shape`` is the shape of ``period`` (number of planets)."""

    def __init__(self, *, period, duration=None, speed=None, time_transit=None, impact_param=None,
                 radius_ratio=None):
        if duration is None and speed is None:
            raise ValueError("Either 'speed' or 'duration' must be provided")
        self.period = _t(period)
        self.time_transit = _t(0.0 if time_transit is None else time_transit)
        self.impact_param = _t(0.0 if impact_param is None else impact_param)
        self.radius_ratio = _t(0.0 if radius_ratio is None else radius_ratio)
        x2 = torch.square(1 + self.radius_ratio) - torch.square(self.impact_param)
        chord = 2 * torch.sqrt(torch.clamp(x2, min=0))
        if duration is None:
            self.speed = _t(speed)
            self.duration = chord / self.speed
        else:
            self.duration = _t(duration)
            self.speed = chord / self.duration

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(self.period.shape)

    @property
    def radius(self):
        return self.radius_ratio

    @property
    def central_radius(self):
        return torch.ones_like(self.period)

    def relative_position(self, t, parallax=None):
        """transit.py:93-107: returns (x, y, z) stacked on axis 0."""
        del parallax
        t_in = t
        t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t, dtype=np.float64))
        period = self.period.to(t.device)
        half_period = 0.5 * period
        ref_time = self.time_transit.to(t.device) - half_period
        dt = torch.remainder(t - ref_time, period) - half_period
        x = self.speed.to(t.device) * dt
        y = torch.zeros_like(dt) + self.impact_param.to(t.device)
        m = torch.abs(dt) < 0.5 * self.duration.to(t.device)
        z = torch.where(m, torch.ones_like(dt), -torch.ones_like(dt))
        out = torch.stack([x, y, z])
        return A.like_input(out, t_in) if not isinstance(t_in, torch.Tensor) else out
