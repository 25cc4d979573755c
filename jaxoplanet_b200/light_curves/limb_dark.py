"""``light_curve(orbit, *u, order=10)(t)``: drop-in for
src/jaxoplanet/light_curves/limb_dark.py:15-62.

For a Keplerian ``System`` the whole chain time -> mean anomaly -> Kepler -> sky position
-> b -> limb-darkened flux (and optionally exposure integration and a Gaussian
log-likelihood) runs in ONE fused CUDA kernel (jxp_system_light_curve_*).  Any other
object satisfying the ``LightCurveOrbit`` protocol (src/jaxoplanet/proto.py:8-20) goes
through its own ``relative_position`` and the K2 kernel; ``TransitOrbit`` / ``TTVOrbit`` have their own
fused kernel (jxp_transit_orbit_light_curve_*).
"""

from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np
import torch

from jaxoplanet_b200 import _array as A
from jaxoplanet_b200 import _lib, constants
from jaxoplanet_b200.core.limb_dark import light_curve as _limb_dark_light_curve
from jaxoplanet_b200.orbits.keplerian import System
from jaxoplanet_b200.orbits.transit import TransitOrbit

__all__ = ["light_curve", "log_likelihood"]


def _field(x, n_sets, default=None):
    if x is None:
        x = default
    x = torch.as_tensor(x, dtype=torch.float64) if not isinstance(x, torch.Tensor) else x.detach().to(torch.float64)
    if x.ndim == 0:
        return x.expand(n_sets)
    return x


def pack_bodies(system: System):
    """-> (bodies [n_sets, n_bodies, 12] float64 tensor, n_sets, batched flag); the record
    layout of include/jaxoplanet_b200.h (JXP_BODY_*)."""
    # bodies whose constants were derived on the device (OrbitalBody._init_on_device): the records exist
    recs_dev = [getattr(body, "_record", None) for body in system.bodies]
    if recs_dev and all(r is not None for r in recs_dev) and len({r.shape[0] for r in recs_dev}) == 1:
        n = recs_dev[0].shape[0]
        return (recs_dev[0].unsqueeze(1) if len(recs_dev) == 1 else torch.stack(recs_dev, dim=1)), n, True
    # the parameter-set axis: the longest leading axis over EVERY packed field of every body (a fixed
    # ephemeris with sampled impact parameter / radius is a common vmap pattern)
    n_sets = 1
    batched = False
    for body in system.bodies:
        for x in (body.period, body.time_transit, body.time_ref, body.eccentricity, body.cos_omega_peri,
                  body.sin_omega_peri, body.cos_inclination, body.sin_inclination, body.semimajor,
                  body.central_radius, body.radius, body.parallax):
            if isinstance(x, torch.Tensor) and x.ndim == 1:
                batched = True
                if n_sets not in (1, x.shape[0]):
                    raise ValueError(
                        f"batched orbit parameters must share one leading axis; got lengths {n_sets} and {x.shape[0]}")
                n_sets = x.shape[0]
    recs = []
    for body in system.bodies:
        dev = body.period.device
        a = body.semimajor
        if body.parallax is not None:  # keplerian.py:620-625 (positions in arcsec)
            a = a * body.parallax / constants.au
        circular = body.eccentricity is None
        f = [None] * _lib.JXP_BODY_NFIELDS
        f[_lib.BODY_PERIOD] = body.period
        f[_lib.BODY_TIME_TRANSIT] = body.time_transit
        f[_lib.BODY_TIME_REF] = body.time_ref
        f[_lib.BODY_ECC] = -1.0 if circular else body.eccentricity
        f[_lib.BODY_COS_OMEGA] = 1.0 if circular else body.cos_omega_peri
        f[_lib.BODY_SIN_OMEGA] = 0.0 if circular else body.sin_omega_peri
        f[_lib.BODY_COS_INCL] = body.cos_inclination
        f[_lib.BODY_SIN_INCL] = body.sin_inclination
        f[_lib.BODY_SEMIMAJOR] = a
        f[_lib.BODY_CENTRAL_RADIUS] = body.central_radius
        f[_lib.BODY_RADIUS] = 0.0 if body.radius is None else body.radius
        f[_lib.BODY_RESERVED] = 0.0
        # fields that already live on one device are stacked there: ONE host-to-device copy of the
        # record array later instead of one per field
        fs = [_field(x, n_sets) for x in f]
        if len({x.device for x in fs}) > 1:
            fs = [x.to(dev) for x in fs]
        recs.append(torch.stack(fs, dim=-1))
    if len({r.device for r in recs}) > 1:
        recs = [r.to(A.device()) for r in recs]
    bodies = torch.stack(recs, dim=1)  # [n_sets, n_bodies, 12]
    return bodies, n_sets, batched


def pack_elements(n, *, central_mass=1.0, central_radius=1.0, **body):
    """Sampled ``Body`` arguments -> the [n, JXP_EL_NFIELDS] element records of kernel K0
    (include/jaxoplanet_b200.h); arguments that are not given become NaN (None in the reference)."""
    names = dict(period=_lib.EL_PERIOD, semimajor=_lib.EL_SEMIMAJOR, time_transit=_lib.EL_TIME_TRANSIT,
                 time_peri=_lib.EL_TIME_PERI, eccentricity=_lib.EL_ECC, omega_peri=_lib.EL_OMEGA,
                 sin_omega_peri=_lib.EL_SIN_OMEGA, cos_omega_peri=_lib.EL_COS_OMEGA,
                 impact_param=_lib.EL_IMPACT_PARAM, inclination=_lib.EL_INCLINATION, mass=_lib.EL_MASS,
                 radius=_lib.EL_RADIUS, parallax=_lib.EL_PARALLAX)
    el = np.full((n, _lib.JXP_EL_NFIELDS), np.nan)
    el[:, _lib.EL_CENTRAL_MASS] = central_mass
    el[:, _lib.EL_CENTRAL_RADIUS] = central_radius
    el[:, _lib.EL_RESERVED] = 0.0
    for k, v in body.items():
        if v is None:
            continue
        if k not in names:
            raise TypeError(f"unknown Body argument {k!r}")
        el[:, names[k]] = v
    return el


def derive_bodies(elements):
    """``OrbitalBody.__init__`` (src/jaxoplanet/orbits/keplerian.py:262-340) on the device: element
    records [..., JXP_EL_NFIELDS] -> body records [..., JXP_BODY_NFIELDS] (CUDA tensor), kernel K0."""
    el = A.to_dev(elements, torch.float64)
    out = torch.empty(tuple(el.shape[:-1]) + (_lib.JXP_BODY_NFIELDS,), dtype=torch.float64, device=el.device)
    _lib.call("jxp_orbital_elements_f64", el.data_ptr(), el.numel() // _lib.JXP_EL_NFIELDS, out.data_ptr(),
              A.stream_ptr())
    return out


@dataclass(frozen=True)
class FusedSpec:
    """Everything the fused kernel needs besides the time stamps."""

    orbit: System
    ld_u: object
    order: int = 10
    exposure_time: float | None = None
    exp_order: int = 0
    num_samples: int = 7
    flags: int = 0


def _run_fused(spec: FusedSpec, time, *, want="flux", y=None, sigma=None, dtype=None, host_sync=False):
    orbit = spec.orbit
    dt = dtype or A.result_dtype(time, spec.ld_u, *[b.period for b in orbit.bodies])
    dev = A.device()
    bodies, n_sets, batched = pack_bodies(orbit)
    n_bodies = bodies.shape[1]
    if bodies.device != dev or not bodies.is_contiguous():
        bodies = bodies.to(dev).contiguous()
    if want not in ("flux", "flux_sum", "loglike"):
        raise ValueError(want)
    # host inputs (NumPy arrays, lists, scalars) of the call go to the device together: one pinned arena, one DMA
    u = spec.ld_u
    host, slots = [], {}
    # host_sync: the caller reads the result back (NumPy out) before it returns to ITS caller, i.e. it
    # synchronises: large host arrays may then be read by DMA straight out of the user's (page-locked) memory
    sync_follows = bool(host_sync)

    def place(name, x, dtype_):
        if isinstance(x, torch.Tensor):
            slots[name] = x.detach().to(device=dev, dtype=dtype_).contiguous()
        elif dtype_ == torch.float64:
            arr = np.asarray(x, dtype=np.float64)
            direct = A.registered_to_dev(arr) if sync_follows else None
            if direct is not None:
                slots[name] = direct
            else:
                slots[name] = len(host)
                host.append(arr)
        else:
            slots[name] = A.to_dev(x, dtype_)

    place("u", u, torch.float64)
    place("t", time, torch.float64)
    if want == "loglike":
        place("y", y, dt)
        place("sigma", sigma, dt)
    if host:
        staged = A.stage_f64(host)
        slots = {k: (staged[v] if isinstance(v, int) else v) for k, v in slots.items()}
    u = slots["u"]
    if u.ndim == 2:
        if u.shape[0] != n_sets:
            raise ValueError("batched limb-darkening coefficients must have shape (n_sets, n_u)")
        u_stride = u.shape[1]
        n_u = u.shape[1]
    else:
        u_stride = 0
        n_u = u.shape[0]
    t = slots["t"]
    t_shape = tuple(t.shape)
    t = t.reshape(-1)
    n_times = t.numel()
    ws_bytes = _lib.load().jxp_system_workspace_bytes(n_sets, n_bodies, n_times)
    ws = torch.empty(max(int(ws_bytes), 256), dtype=torch.uint8, device=dev)
    flux = flux_sum = loglike = None
    yd = sd = None
    s_stride = 0
    if want == "flux":
        flux = torch.empty((n_sets, n_times, n_bodies), dtype=dt, device=dev)
    elif want == "flux_sum":
        flux_sum = torch.empty((n_sets, n_times), dtype=dt, device=dev)
    else:
        loglike = torch.empty((n_sets,), dtype=dt, device=dev)
        yd = slots["y"].reshape(-1)
        sd = slots["sigma"].reshape(-1)
        if yd.numel() != n_times:
            raise ValueError("y must have the shape of the time array")
        if sd.numel() == 1:
            s_stride = 0
        elif sd.numel() == n_times:
            s_stride = 1
        else:
            raise ValueError("sigma must be a scalar or have the shape of the time array")
    name = "jxp_system_light_curve_f64" if dt == torch.float64 else "jxp_system_light_curve_f32"
    expo = 0.0 if spec.exposure_time is None else float(spec.exposure_time)
    _lib.call(name, A.ptr(bodies), n_sets, n_bodies, A.ptr(u), n_u, u_stride, A.ptr(t), n_times,
              expo, int(spec.exp_order), int(spec.num_samples), int(spec.order), int(spec.flags),
              A.ptr(flux), A.ptr(flux_sum), A.ptr(yd), A.ptr(sd), s_stride, A.ptr(loglike),
              A.ptr(ws), ws.numel(), A.stream_ptr())
    if want == "flux":
        out = flux.reshape((n_sets,) + t_shape + (n_bodies,))
    elif want == "flux_sum":
        out = flux_sum.reshape((n_sets,) + t_shape)
    else:
        out = loglike
    if not batched:
        out = out[0]
    return out


def _all_host(*xs) -> bool:
    """No torch tensor among the inputs: like_input() will hand back NumPy, after a device-to-host copy."""
    return not any(isinstance(x, torch.Tensor) for x in xs)


def _concat_u(u):
    if not u:
        return np.array([], dtype=np.float64)
    if any(isinstance(u_, torch.Tensor) for u_ in u):
        return torch.cat([torch.atleast_1d(torch.as_tensor(u_)) for u_ in u], dim=-1)
    return np.concatenate([np.atleast_1d(np.asarray(u_, dtype=np.float64)) for u_ in u], axis=-1)


def light_curve(orbit, *u, order: int = 10):
    """Compute the light curve for (quadratic) polynomial limb darkening.

    Args:
        orbit: a Keplerian ``System`` (fused kernel) or any object with ``shape``,
            ``radius``, ``central_radius`` and ``relative_position(t)``.
        u: The coefficients of the polynomial limb darkening
        order: The order of the numerical integration used by the backend

    Returns:
        A function which takes the time in days and returns the light curve flux
        (``flux - 1`` per body, shape ``t.shape + orbit.shape``).
    """
    ld_u = _concat_u(u)

    if isinstance(orbit, System):
        spec = FusedSpec(orbit=orbit, ld_u=ld_u, order=order)

        def light_curve_impl(time):
            return A.like_input(_run_fused(spec, time, host_sync=_all_host(time)), time)

        light_curve_impl._jxp_spec = spec
        return light_curve_impl

    if isinstance(orbit, TransitOrbit):
        # TransitOrbit / TTVOrbit: time warp + linear motion + flux fused in k_transit_orbit (K6)
        def light_curve_transit(time):
            dt = A.result_dtype(time)
            td = A.to_dev(time, torch.float64)
            n_p = int(np.prod(orbit.shape)) if len(orbit.shape) else 1
            fields = [orbit.period, orbit.time_transit, orbit.speed, orbit.duration, orbit.impact_param,
                      orbit.radius_ratio]
            planets = torch.stack([torch.broadcast_to(torch.as_tensor(f, dtype=torch.float64).reshape(-1), (n_p,))
                                   for f in fields], dim=-1)
            planets = A.to_dev(planets, torch.float64)
            u_host = np.ascontiguousarray(ld_u.detach().cpu().numpy() if isinstance(ld_u, torch.Tensor) else ld_u,
                                          dtype=np.float64)
            if u_host.ndim != 1:
                raise ValueError("TransitOrbit light curves take one limb-darkening law (u.ndim == 1)")
            tabs = [None, None, None, None]
            if hasattr(orbit, "_ttv_tables"):
                e, v, off, t0 = orbit._ttv_tables()
                tabs = [A.to_dev(e, torch.float64), A.to_dev(v, torch.float64),
                        torch.as_tensor(off, dtype=torch.int32, device=A.device()), A.to_dev(t0, torch.float64)]
            flux = torch.empty(tuple(td.shape) + (n_p,), dtype=dt, device=td.device)
            name = "jxp_transit_orbit_light_curve_f64" if dt == torch.float64 else "jxp_transit_orbit_light_curve_f32"
            _lib.call(name, planets.data_ptr(), n_p, A.ptr(tabs[0]), A.ptr(tabs[1]), A.ptr(tabs[2]), A.ptr(tabs[3]),
                      u_host.ctypes.data, int(u_host.shape[0]), td.data_ptr(), td.numel(), int(order),
                      flux.data_ptr(), A.stream_ptr())
            if len(orbit.shape) == 0:
                flux = flux.reshape(td.shape)
            return A.like_input(flux, time)

        return light_curve_transit

    def light_curve_generic(time):
        # light_curves/limb_dark.py:49-60 for a foreign orbit object (LightCurveOrbit protocol)
        dt = A.result_dtype(time)
        pos = orbit.relative_position(time)
        x, y, z = (A.to_dev(v, dt) for v in pos)
        t_ndim = np.ndim(time) if not isinstance(time, torch.Tensor) else time.ndim
        r_star = A.to_dev(orbit.central_radius, dt)
        radius = A.to_dev(orbit.radius, dt)
        if x.ndim > t_ndim:  # (n_bodies, *t.shape) -> t.shape + (n_bodies,)
            x, y, z = (torch.movedim(v, 0, -1) for v in (x, y, z))
        b = torch.sqrt(x**2 + y**2) / r_star
        r = (radius / r_star).expand(b.shape).contiguous()
        lc = _limb_dark_light_curve(A.to_dev(ld_u, dt), b.contiguous(), r, order=order)
        lc = torch.where(z > 0, lc, torch.zeros_like(lc))
        return A.like_input(lc, time)

    return light_curve_generic


def log_likelihood(orbit: System, *u, order: int = 10, exposure_time=None, exp_order: int = 0,
                   num_samples: int = 7):
    """Fused ``sum_t Normal(1 + sum_bodies light_curve, sigma).log_prob(y)`` per parameter
    set (the NumPyro pattern of docs/tutorials/transit.ipynb cell 7) -- kernel K4.

    Returns a function ``(t, y, sigma) -> loglike`` (scalar, or ``(n_sets,)`` when batched).
    """
    spec = FusedSpec(orbit=orbit, ld_u=_concat_u(u), order=order, exposure_time=exposure_time,
                     exp_order=exp_order, num_samples=num_samples)

    def impl(time, y, sigma):
        return A.like_input(_run_fused(spec, time, want="loglike", y=y, sigma=sigma, host_sync=_all_host(time, y)),
                            time, y)

    impl._jxp_spec = spec
    impl._jxp_want = "loglike"
    return impl


def with_spec(func, **changes):
    """The fused function of `func` with some of its spec replaced; a log-likelihood stays a log-likelihood
    (signature (t, y, sigma)), a light curve a light curve."""
    spec = replace(func._jxp_spec, **changes)
    if getattr(func, "_jxp_want", "flux") == "loglike":
        def loglike_impl(time, y, sigma):
            return A.like_input(_run_fused(spec, time, want="loglike", y=y, sigma=sigma, host_sync=_all_host(time, y)),
                            time, y)

        loglike_impl._jxp_spec = spec
        loglike_impl._jxp_want = "loglike"
        return loglike_impl

    def light_curve_impl(time):
        return A.like_input(_run_fused(spec, time, host_sync=_all_host(time)), time)

    light_curve_impl._jxp_spec = spec
    return light_curve_impl
