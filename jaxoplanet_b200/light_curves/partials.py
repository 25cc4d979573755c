"""Flux and forward-mode partials of a single transiting planet w.r.t. the 8 sampled
parameters theta = (period, time_transit, eccentricity, omega_peri, impact_param, radius,
u1, u2) -- kernel K5 (jxp_transit_partials_*).

This is what ``jax.jvp`` / ``jax.grad`` evaluate when a reference user differentiates

    limb_dark_light_curve(System(Central(mass, radius)).add_body(period=..., time_transit=...,
        eccentricity=..., omega_peri=..., impact_param=..., radius=...), [u1, u2])(t)[..., 0]

(SURVEY.md 3.5): the kernel returns the partials, the tangent/cotangent contraction stays in
the host framework (``TransitLightCurve`` below does it for ``torch.autograd``).
"""

from __future__ import annotations

import numpy as np
import torch

from jaxoplanet_b200 import _array as A
from jaxoplanet_b200 import _lib

__all__ = ["transit_partials", "TransitLightCurve", "THETA_NAMES"]

THETA_NAMES = ("period", "time_transit", "eccentricity", "omega_peri", "impact_param", "radius",
               "u1", "u2")


def transit_partials(theta, t, *, central_mass=1.0, central_radius=1.0, order: int = 10,
                     exposure_time=None, exp_order: int = 0, num_samples: int = 7, y=None,
                     sigma=None, want_flux=True, want_partials=True, want_dloglike=True,
                     dtype=torch.float64, flags: int = 0):
    """theta: (n_sets, 8) or (8,) array in THETA_NAMES order; t: (n_times,).

    Returns a dict with ``flux`` (n_sets, n_times), ``partials`` (n_sets, 8, n_times) and, when
    ``y``/``sigma`` are given, ``loglike`` (n_sets,) and ``dloglike`` (n_sets, 8).
    Arrays are torch CUDA tensors if any input was a torch tensor, else numpy.

    A theta with 6 + n_u columns, n_u != 2 (the six orbit parameters followed by the coefficients u_1 .. u_n of a
    polynomial limb-darkening law of any length up to 8, as ``core.limb_dark.light_curve`` takes them), gives the
    partials w.r.t. those 6 + n_u parameters (double precision, no exposure integration).
    """
    dev = A.device()
    th = A.to_dev(theta, torch.float64)
    single = th.ndim == 1
    n_theta = int(th.shape[-1])
    if n_theta != _lib.JXP_NTHETA:
        return _transit_partials_poly(th, single, t, central_mass, central_radius, order, exposure_time, y, sigma,
                                      want_flux, want_partials, want_dloglike, dtype, flags, theta)
    th = th.reshape(-1, _lib.JXP_NTHETA).contiguous()
    n_sets = th.shape[0]
    star = torch.stack(torch.broadcast_tensors(
        torch.as_tensor(central_mass, dtype=torch.float64), torch.as_tensor(central_radius, dtype=torch.float64)), -1)
    if star.ndim == 1:
        star_stride = 0
    else:
        if star.shape[0] != n_sets:
            raise ValueError("central_mass / central_radius must be scalars or (n_sets,)")
        star_stride = 2
    star = star.to(dev).contiguous()
    td = A.to_dev(t, torch.float64).reshape(-1)
    n_times = td.numel()
    flux = torch.empty((n_sets, n_times), dtype=dtype, device=dev) if want_flux else None
    part = torch.empty((n_sets, _lib.JXP_NTHETA, n_times), dtype=dtype, device=dev) if want_partials else None
    ll = dll = yd = sd = None
    s_stride = 0
    if y is not None:
        yd = A.to_dev(y, dtype).reshape(-1)
        sd = A.to_dev(sigma, dtype).reshape(-1)
        if yd.numel() != n_times:
            raise ValueError("y must have the shape of the time array")
        if sd.numel() not in (1, n_times):
            raise ValueError("sigma must be a scalar or have the shape of the time array")
        s_stride = 0 if sd.numel() == 1 else 1
        ll = torch.empty((n_sets,), dtype=dtype, device=dev)
        dll = torch.empty((n_sets, _lib.JXP_NTHETA), dtype=dtype, device=dev) if want_dloglike else None
    wsb = int(_lib.load().jxp_transit_workspace_bytes(n_sets, n_times))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    name = "jxp_transit_partials_f64" if dtype == torch.float64 else "jxp_transit_partials_f32"
    expo = 0.0 if exposure_time is None else float(exposure_time)
    _lib.call(name, A.ptr(th), n_sets, A.ptr(star), star_stride, A.ptr(td), n_times, expo,
              int(exp_order), int(num_samples), int(order), int(flags), A.ptr(flux), A.ptr(part),
              A.ptr(yd), A.ptr(sd), s_stride, A.ptr(ll), A.ptr(dll), A.ptr(ws), ws.numel(),
              A.stream_ptr())
    out = {}
    for key, val in (("flux", flux), ("partials", part), ("loglike", ll), ("dloglike", dll)):
        if val is not None:
            if single:
                val = val[0]
            out[key] = A.like_input(val, theta, t)
    return out


def _transit_partials_poly(th, single, t, central_mass, central_radius, order, exposure_time, y, sigma, want_flux,
                           want_partials, want_dloglike, dtype, flags, theta_in):
    """General polynomial limb darkening: jxp_transit_partials_poly_f64 (theta rows of 6 + n_u)."""
    dev = A.device()
    n_theta = int(th.shape[-1])
    n_u = n_theta - 6
    if n_u < 0 or n_u > 8:
        raise ValueError("theta must have 6 + n_u columns with 0 <= n_u <= 8")
    if dtype != torch.float64 or exposure_time is not None or order != 10:
        raise NotImplementedError("partials for laws other than quadratic: float64, order 10, no exposure integration")
    th = th.reshape(-1, n_theta).contiguous()
    n_sets = th.shape[0]
    star = torch.stack(torch.broadcast_tensors(
        torch.as_tensor(central_mass, dtype=torch.float64), torch.as_tensor(central_radius, dtype=torch.float64)), -1)
    star_stride = 0 if star.ndim == 1 else 2
    if star.ndim != 1 and star.shape[0] != n_sets:
        raise ValueError("central_mass / central_radius must be scalars or (n_sets,)")
    star = star.to(dev).contiguous()
    td = A.to_dev(t, torch.float64).reshape(-1)
    n_times = td.numel()
    flux = torch.empty((n_sets, n_times), dtype=dtype, device=dev) if want_flux else None
    part = torch.empty((n_sets, n_theta, n_times), dtype=dtype, device=dev) if want_partials else None
    ll = dll = yd = sd = None
    s_stride = 0
    if y is not None:
        yd = A.to_dev(y, dtype).reshape(-1)
        sd = A.to_dev(sigma, dtype).reshape(-1)
        if yd.numel() != n_times:
            raise ValueError("y must have the shape of the time array")
        if sd.numel() not in (1, n_times):
            raise ValueError("sigma must be a scalar or have the shape of the time array")
        s_stride = 0 if sd.numel() == 1 else 1
        ll = torch.empty((n_sets,), dtype=dtype, device=dev)
        dll = torch.empty((n_sets, n_theta), dtype=dtype, device=dev) if want_dloglike else None
    wsb = int(_lib.load().jxp_transit_poly_workspace_bytes(n_sets, n_times))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    _lib.call("jxp_transit_partials_poly_f64", A.ptr(th), n_sets, n_u, A.ptr(star), star_stride, A.ptr(td), n_times,
              int(order), int(flags), A.ptr(flux), A.ptr(part), A.ptr(yd), A.ptr(sd), s_stride, A.ptr(ll), A.ptr(dll),
              A.ptr(ws), ws.numel(), A.stream_ptr())
    out = {}
    for key, val in (("flux", flux), ("partials", part), ("loglike", ll), ("dloglike", dll)):
        if val is not None:
            if single:
                val = val[0]
            out[key] = A.like_input(val, theta_in, t)
    return out


class TransitLightCurve(torch.autograd.Function):
    """flux(theta, t) with ``torch.autograd`` support: backward contracts the kernel's
    partials with the incoming cotangent (the custom-JVP pattern of SURVEY.md 3.5)."""

    @staticmethod
    def forward(ctx, theta, t, central_mass, central_radius):
        res = transit_partials(theta.detach(), t, central_mass=central_mass,
                               central_radius=central_radius)
        ctx.save_for_backward(res["partials"])
        ctx.single = theta.ndim == 1
        return res["flux"]

    @staticmethod
    def backward(ctx, g):
        (part,) = ctx.saved_tensors
        if ctx.single:
            gtheta = (part * g.unsqueeze(0)).sum(-1)
        else:
            gtheta = (part * g.unsqueeze(1)).sum(-1)
        return gtheta, None, None, None
