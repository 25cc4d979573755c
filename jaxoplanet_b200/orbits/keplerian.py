"""Keplerian systems of bodies: the host-side mirror of
src/jaxoplanet/orbits/keplerian.py (same classes, fields, argument rules and errors).

The per-body derived constants (``OrbitalBody.__init__``, keplerian.py:262-340) are a few
dozen scalar operations per parameter set and are evaluated here with torch scalar ops;
everything that runs per time sample goes to the CUDA library: positions and velocities
through the Kepler kernel (K1), light curves through the fused kernel (K3/K4), see
``jaxoplanet_b200.light_curves.limb_dark``.

Inside ``jaxoplanet_b200.batched()`` every parameter may carry one leading axis of
parameter sets (the reference user writes ``jax.vmap`` for that).
"""

from __future__ import annotations

import math
from collections.abc import Callable, Iterable, Sequence
from typing import Any

import numpy as np
import torch

from jaxoplanet_b200 import _array as A
from jaxoplanet_b200 import constants

Scalar = Any


def _t(x):
    """Scalar-like -> float tensor (keeps device / float dtype of tensors)."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x if x.dtype.is_floating_point else x.to(torch.float64)
    arr = np.asarray(x)
    if arr.dtype == np.float32:
        return torch.as_tensor(arr)
    return torch.as_tensor(arr, dtype=torch.float64)


def _ndim(x) -> int:
    return x.ndim if isinstance(x, torch.Tensor) else np.ndim(x)


def _bad_rank(args) -> bool:
    return any(_ndim(a) > A.batch_ndim() for a in args if a is not None)


class Central:
    """A central body in an orbital system (keplerian.py:15-114).

    Args:
        mass: Mass of central body [mass unit].
        radius: Radius of central body [length unit].
        density: Density of central body [mass/length**3 unit].
    """

    def __init__(self, *, mass=None, radius=None, density=None):
        if radius is None and mass is None:
            radius = 1.0
            if density is None:
                mass = 1.0

        if _bad_rank((mass, radius, density)):
            raise ValueError("All parameters of a KeplerianCentral must be scalars")

        error_msg = "Values must be provided for exactly two of mass, radius, and density"
        if all(x is None or isinstance(x, (int, float)) for x in (mass, radius, density)):
            # plain Python scalars (the common case, and the default Central()): the two-of-three rule in
            # float arithmetic, three tensors at the end instead of a dozen 0-d tensor operations
            if sum(x is None for x in (mass, radius, density)) != 1:
                raise ValueError(error_msg)
            if density is None:
                density = 3 * mass / (4 * math.pi * radius**3)
            elif radius is None:
                radius = (3 * mass / (4 * math.pi * density)) ** (1 / 3)
            else:
                mass = 4 * math.pi * radius**3 * density / 3.0
            self.mass, self.radius, self.density = (torch.tensor(float(x), dtype=torch.float64)
                                                    for x in (mass, radius, density))
            return
        mass, radius, density = _t(mass), _t(radius), _t(density)
        if density is None:
            if mass is None or radius is None:
                raise ValueError(error_msg)
            self.mass = mass
            self.radius = radius
            self.density = 3 * mass / (4 * math.pi * radius**3)
        elif radius is None:
            if mass is None or density is None:
                raise ValueError(error_msg)
            self.mass = mass
            self.radius = (3 * mass / (4 * math.pi * density)) ** (1 / 3)
            self.density = density
        elif mass is None:
            if radius is None or density is None:
                raise ValueError(error_msg)
            self.mass = 4 * math.pi * radius**3 * density / 3.0
            self.radius = radius
            self.density = density
        else:
            raise ValueError(error_msg)

    @classmethod
    def from_orbital_properties(cls, *, period, semimajor, radius=None, body_mass=None):
        if _bad_rank((semimajor, period, body_mass)):
            raise ValueError(
                "All parameters of 'KeplerianCentral.from_orbital_properties' must be "
                "scalars; for multi-planet systems, use 'jax.vmap'"
            )
        radius = 1.0 if radius is None else radius
        mass = 4 * math.pi**2 * _t(semimajor) ** 3 / (constants.G * _t(period) ** 2)
        if body_mass is not None:
            mass = mass - _t(body_mass)
        return cls(mass=mass, radius=radius)

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(self.mass.shape)


_BODY_FIELDS = (
    "time_transit", "time_peri", "period", "semimajor", "inclination", "impact_param",
    "eccentricity", "omega_peri", "sin_omega_peri", "cos_omega_peri", "asc_node",
    "sin_asc_node", "cos_asc_node", "mass", "radius", "radial_velocity_semiamplitude",
    "parallax",
)


class Body:
    """An orbiting body defined by user-facing orbital parameters (keplerian.py:117-238)."""

    def __init__(self, **kwargs):
        for name in _BODY_FIELDS:
            setattr(self, name, kwargs.pop(name, None))
        if kwargs:
            raise TypeError(f"Body() got unexpected keyword arguments {sorted(kwargs)}")
        self.__check_init__()

    def __check_init__(self) -> None:
        if not ((self.period is None) ^ (self.semimajor is None)):
            raise ValueError("Exactly one of period or semimajor must be specified")

        provided = [getattr(self, n) for n in _BODY_FIELDS if getattr(self, n) is not None]
        if _bad_rank(provided):
            raise ValueError(
                "All input arguments to 'Body' must be scalars; "
                "for multi-planet systems, use a 'System'"
            )

        if self.omega_peri is not None and (
            self.sin_omega_peri is not None or self.cos_omega_peri is not None
        ):
            raise ValueError("Cannot specify both omega_peri and sin_omega_peri or cos_omega_peri")
        if (self.sin_omega_peri is not None) ^ (self.cos_omega_peri is not None):
            raise ValueError("Both sin_omega_peri and cos_omega_peri must be specified")

        if self.asc_node is not None and (
            self.sin_asc_node is not None or self.cos_asc_node is not None
        ):
            raise ValueError("Cannot specify both asc_node and sin_asc_node or cos_asc_node")
        if (self.sin_asc_node is not None) ^ (self.cos_asc_node is not None):
            raise ValueError("Both sin_asc_node and cos_asc_node must be specified")

        has_omega_peri = (
            self.omega_peri is not None
            or self.sin_omega_peri is not None
            or self.cos_omega_peri is not None
        )
        if (self.eccentricity is not None) ^ has_omega_peri:
            raise ValueError("Both or neither of eccentricity and omega_peri must be specified")

        if self.impact_param is not None and self.inclination is not None:
            raise ValueError("Only one of impact_param and inclination can be specified")

        if self.time_transit is not None and self.time_peri is not None:
            raise ValueError("Only one of time_transit or time_peri can be specified")


_EL_NAMES = ("period", "semimajor", "time_transit", "time_peri", "eccentricity", "omega_peri", "sin_omega_peri",
             "cos_omega_peri", "impact_param", "inclination", "mass", "radius", "parallax")
_EL_INDEX = dict(period=0, semimajor=1, time_transit=2, time_peri=3, eccentricity=4, omega_peri=5, sin_omega_peri=6,
                 cos_omega_peri=7, impact_param=8, inclination=9, mass=10, radius=11, parallax=14)  # JXP_EL_*


class OrbitalBody:
    """A body with its derived orbital constants (keplerian.py:241-637)."""

    def __init__(self, central: Central, body: Body):
        self.central = central
        self._record = None
        if self._init_on_device(central, body):
            return
        self.radius = _t(body.radius)
        self.mass = _t(body.mass)
        self.radial_velocity_semiamplitude = _t(body.radial_velocity_semiamplitude)
        self.parallax = _t(body.parallax)

        # period <-> semimajor (keplerian.py:272-282)
        mass_factor = constants.G * self.total_mass
        if body.semimajor is None:
            assert body.period is not None
            period = _t(body.period)
            x = mass_factor * period**2 / (4 * math.pi**2)
            self.semimajor = _cbrt(x)
            self.period = period + 0 * self.semimajor
        else:
            semimajor = _t(body.semimajor)
            self.semimajor = semimajor
            self.period = 2 * math.pi * semimajor * torch.sqrt(semimajor / mass_factor)

        if body.omega_peri is not None:
            w = _t(body.omega_peri)
            self.sin_omega_peri = torch.sin(w)
            self.cos_omega_peri = torch.cos(w)
        else:
            self.sin_omega_peri = _t(body.sin_omega_peri)
            self.cos_omega_peri = _t(body.cos_omega_peri)

        if body.asc_node is not None:
            o = _t(body.asc_node)
            self.sin_asc_node = torch.sin(o)
            self.cos_asc_node = torch.cos(o)
        else:
            self.sin_asc_node = _t(body.sin_asc_node)
            self.cos_asc_node = _t(body.cos_asc_node)

        # eccentric vs circular (keplerian.py:298-315)
        self.eccentricity = _t(body.eccentricity)
        if self.eccentricity is None:
            M0 = torch.full_like(self.period, 0.5 * math.pi)
            incl_factor = 1
        else:
            e = self.eccentricity
            opsw = 1 + self.sin_omega_peri
            E0 = 2 * torch.atan2(
                torch.sqrt(1 - e) * self.cos_omega_peri,
                torch.sqrt(1 + e) * opsw,
            )
            M0 = E0 - e * torch.sin(E0)
            ome2 = 1 - e**2
            incl_factor = (1 + e * self.sin_omega_peri) / ome2

        # inclination (keplerian.py:317-331)
        dcosidb = incl_factor * central.radius / self.semimajor
        if body.impact_param is not None:
            self.impact_param = _t(body.impact_param) + 0 * self.period
            self.cos_inclination = dcosidb * self.impact_param
            self.sin_inclination = torch.sqrt(1 - self.cos_inclination**2)
        elif body.inclination is not None:
            inc = _t(body.inclination)
            self.cos_inclination = torch.cos(inc) + 0 * self.period
            self.sin_inclination = torch.sin(inc) + 0 * self.period
            self.impact_param = self.cos_inclination / dcosidb
        else:
            z = torch.zeros_like(self.period)
            self.impact_param = z
            self.cos_inclination = z
            self.sin_inclination = torch.ones_like(self.period)

        # reference times (keplerian.py:333-340)
        self.time_ref = -M0 * self.period / (2 * math.pi)
        if body.time_transit is not None:
            self.time_transit = _t(body.time_transit) + 0 * self.time_ref
        elif body.time_peri is not None:
            self.time_transit = _t(body.time_peri) - self.time_ref
        else:
            self.time_transit = torch.zeros_like(self.time_ref)

    # -- large batches: the same constants, derived on the device by kernel K0 -----------
    _DEVICE_MIN_SETS = 1  # every batched call: results must not depend on the batch size

    def _init_on_device(self, central: Central, body: Body) -> bool:
        """``jaxoplanet_b200.batched()`` with the parameter sets given as host arrays: the sampled
        arguments go to the device as ONE record array and kernel K0 (jxp_orbital_elements_f64) derives
        what keplerian.py:262-340 derives; the attributes are views of its output.  Returns False when the
        call is not of that kind (torch inputs, ascending node, RV semi-amplitude, no CUDA device): the
        host path below then does the same arithmetic with torch."""
        if not A.batch_ndim():
            return False
        if (body.asc_node is not None or body.sin_asc_node is not None or body.cos_asc_node is not None
                or body.radial_velocity_semiamplitude is not None):
            return False
        given, n = {}, 0
        for k in _EL_NAMES:
            v = getattr(body, k)
            if v is not None:
                given[k] = v
        for v in (*given.values(), central.mass, central.radius):
            if isinstance(v, torch.Tensor):
                if v.is_cuda or v.requires_grad or v.dtype == torch.float32:
                    return False
                nd, n0 = v.ndim, (v.shape[0] if v.ndim == 1 else 0)
            else:
                if isinstance(v, np.ndarray) and v.dtype == np.float32:
                    return False  # (the fp32 mode keeps its inputs' dtype through the host path)
                nd = np.ndim(v)
                n0 = np.shape(v)[0] if nd == 1 else 0
            if nd == 1:
                if n and n0 != n:
                    return False
                n = n0
        if n < self._DEVICE_MIN_SETS:
            return False
        try:
            dev = A.device()
        except RuntimeError:
            return False  # no CUDA device: the host arithmetic below (records can still be packed and checked)
        from jaxoplanet_b200 import _lib

        # the element records are packed straight into a page-locked arena, FIELD-major (one contiguous row
        # per Body argument; include/jaxoplanet_b200.h, JXP_EL_*; NaN = not given), and go over in one DMA
        nf = _lib.JXP_EL_NFIELDS
        view, commit = A.stage_view(n * nf)
        el = view[: n * nf].reshape(nf, n)
        el[_lib.EL_CENTRAL_MASS] = float(central.mass) if np.ndim(central.mass) == 0 else central.mass
        el[_lib.EL_CENTRAL_RADIUS] = float(central.radius) if np.ndim(central.radius) == 0 else central.radius
        el[_lib.EL_RESERVED] = 0.0
        for k in _EL_NAMES:
            v = given.get(k)
            el[_EL_INDEX[k]] = np.nan if v is None else v
        eld = commit()[: n * nf].view(nf, n)
        rec = torch.empty((n, _lib.JXP_BODY_NFIELDS), dtype=torch.float64, device=dev)
        _lib.call("jxp_orbital_elements_soa_f64", eld.data_ptr(), n, rec.data_ptr(), A.stream_ptr())
        self._record = rec
        circular = body.eccentricity is None
        cols = rec.unbind(1)
        self.period = cols[_lib.BODY_PERIOD]
        self.time_transit = cols[_lib.BODY_TIME_TRANSIT]
        self.time_ref = cols[_lib.BODY_TIME_REF]
        self.eccentricity = None if circular else cols[_lib.BODY_ECC]
        self.cos_omega_peri = None if circular else cols[_lib.BODY_COS_OMEGA]
        self.sin_omega_peri = None if circular else cols[_lib.BODY_SIN_OMEGA]
        self.cos_inclination = cols[_lib.BODY_COS_INCL]
        self.sin_inclination = cols[_lib.BODY_SIN_INCL]
        self.semimajor = cols[_lib.BODY_SEMIMAJOR] if body.parallax is None else \
            cols[_lib.BODY_SEMIMAJOR] * constants.au / eld[_lib.EL_PARALLAX]
        self.radius = None if body.radius is None else cols[_lib.BODY_RADIUS]
        self.mass = None if body.mass is None else eld[_lib.EL_MASS]
        self.parallax = None if body.parallax is None else eld[_lib.EL_PARALLAX]
        self.radial_velocity_semiamplitude = None
        self.sin_asc_node = self.cos_asc_node = None
        self._impact_param_given = None if body.impact_param is None else eld[_lib.EL_IMPACT_PARAM]
        return True

    @property
    def impact_param(self):
        ip = self.__dict__.get("_impact_param")
        if ip is not None:
            return ip
        if self._record is None:
            raise AttributeError("impact_param")
        if self._impact_param_given is not None:
            return self._impact_param_given
        # keplerian.py:317-331: cos i / (d cos i / d b)
        if self.eccentricity is None:
            incl_factor = 1.0
        else:
            incl_factor = (1 + self.eccentricity * self.sin_omega_peri) / (1 - self.eccentricity**2)
        return self.cos_inclination * self.semimajor / (incl_factor * _on(self.central.radius, self.semimajor))

    @impact_param.setter
    def impact_param(self, value):
        self.__dict__["_impact_param"] = value

    # -- properties (keplerian.py:342-364) ------------------------------------------
    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(self.period.shape)

    @property
    def central_radius(self):
        return self.central.radius

    @property
    def time_peri(self):
        return self.time_transit + self.time_ref

    @property
    def inclination(self):
        return torch.atan2(self.sin_inclination, self.cos_inclination)

    @property
    def omega_peri(self):
        if self.eccentricity is None:
            return None
        return torch.atan2(self.sin_omega_peri, self.cos_omega_peri)

    @property
    def total_mass(self):
        return self.central.mass if self.mass is None else self.mass + self.central.mass

    # -- positions and velocities (keplerian.py:366-532) ----------------------------
    def position(self, t, parallax=None):
        semimajor = -self.semimajor * self.central.mass / self.total_mass
        return self._get_position_and_velocity(t, semimajor=semimajor, parallax=parallax)[0]

    def central_position(self, t, parallax=None):
        semimajor = self.semimajor * self.mass / self.total_mass
        return self._get_position_and_velocity(t, semimajor=semimajor, parallax=parallax)[0]

    def relative_position(self, t, parallax=None):
        return self._get_position_and_velocity(t, semimajor=-self.semimajor, parallax=parallax)[0]

    def relative_angles(self, t, parallax=None):
        X, Y, _ = self.relative_position(t, parallax=parallax)
        if isinstance(X, torch.Tensor):
            return torch.sqrt(X**2 + Y**2), torch.atan2(Y, X)
        return np.sqrt(X**2 + Y**2), np.arctan2(Y, X)

    def velocity(self, t, semiamplitude=None):
        if semiamplitude is None:
            return self._get_position_and_velocity(t, mass=-self.central.mass)[1]
        k = -semiamplitude * self.central.mass / self.mass
        return self._get_position_and_velocity(t, semiamplitude=k)[1]

    def central_velocity(self, t, semiamplitude=None):
        if semiamplitude is None:
            return self._get_position_and_velocity(t, mass=self.mass)[1]
        return self._get_position_and_velocity(t, semiamplitude=semiamplitude)[1]

    def relative_velocity(self, t, semiamplitude=None):
        if semiamplitude is None:
            return self._get_position_and_velocity(t, mass=-self.total_mass)[1]
        k = -semiamplitude * self.total_mass / self.mass
        return self._get_position_and_velocity(t, semiamplitude=k)[1]

    def radial_velocity(self, t, semiamplitude=None):
        rv = -self.central_velocity(t, semiamplitude=semiamplitude)[2]
        if (
            self.parallax is not None
            and semiamplitude is None
            and self.radial_velocity_semiamplitude is None
        ):
            rv = rv / _on(self.parallax, rv) * constants.au
        return rv

    # -- internals --------------------------------------------------------------------
    def _warp_times(self, t):
        return t - self.time_transit

    def _state_record(self, dev) -> torch.Tensor:
        """[n_sets, JXP_BODY_NFIELDS] float64 record of this body on the device (the layout of
        include/jaxoplanet_b200.h; what the orbit-state kernel K8 reads)."""
        from jaxoplanet_b200 import _lib

        if self._record is not None:  # derived on the device by K0: the record exists
            return self._record if self._record.device == dev else self._record.to(dev)
        n = self.period.shape[0] if self.period.ndim == 1 else 1
        circular = self.eccentricity is None
        f = [None] * _lib.JXP_BODY_NFIELDS
        f[_lib.BODY_PERIOD] = self.period
        f[_lib.BODY_TIME_TRANSIT] = self.time_transit
        f[_lib.BODY_TIME_REF] = self.time_ref
        f[_lib.BODY_ECC] = -1.0 if circular else self.eccentricity
        f[_lib.BODY_COS_OMEGA] = 1.0 if circular else self.cos_omega_peri
        f[_lib.BODY_SIN_OMEGA] = 0.0 if circular else self.sin_omega_peri
        f[_lib.BODY_COS_INCL] = self.cos_inclination
        f[_lib.BODY_SIN_INCL] = self.sin_inclination
        f[_lib.BODY_SEMIMAJOR] = self.semimajor
        f[_lib.BODY_CENTRAL_RADIUS] = self.central.radius
        f[_lib.BODY_RADIUS] = 0.0 if self.radius is None else self.radius
        f[_lib.BODY_RESERVED] = 0.0
        cols = [_per_set(x, n).to(dev) for x in f]
        return torch.stack(cols, dim=-1).contiguous()

    def _get_position_and_velocity(self, t, semimajor=None, mass=None, semiamplitude=None,
                                   parallax=None):
        """keplerian.py:590-637.  The per-set scalars (the factor of the orbital radius, the velocity
        amplitude k0) are derived here; everything per time sample -- mean anomaly, Kepler solve, radius,
        the three rotations -- runs in ONE kernel (K8, jxp_orbit_state_*)."""
        from jaxoplanet_b200 import _lib

        t_in = t
        dt = A.result_dtype(t)
        dev = A.device()
        td = A.to_dev(t, torch.float64)
        t_shape = tuple(td.shape)
        batched = bool(A.batch_ndim()) and self.period.ndim == 1
        n_sets = self.period.shape[0] if self.period.ndim == 1 else 1

        if semiamplitude is None:
            semiamplitude = self.radial_velocity_semiamplitude
        if parallax is None:
            parallax = self.parallax

        if semiamplitude is None:
            if self.radial_velocity_semiamplitude is None:
                m = 1.0 if mass is None else mass
                k0 = 2 * math.pi * self.semimajor * m / (self.total_mass * self.period)
                if self.eccentricity is not None:
                    k0 = k0 / torch.sqrt(1 - self.eccentricity**2)
            else:
                k0 = self.radial_velocity_semiamplitude
            if parallax is not None:
                k0 = k0 * parallax / constants.au
        else:
            k0 = semiamplitude

        r0 = 1.0
        if semimajor is not None:
            r0 = semimajor if parallax is None else semimajor * parallax / constants.au

        rec = self._state_record(dev)
        r0d = _per_set(r0, n_sets).to(device=dev, dtype=torch.float64).contiguous()
        k0d = _per_set(k0, n_sets).to(device=dev, dtype=torch.float64).contiguous()
        asc = None
        if self.cos_asc_node is not None:
            asc = torch.stack([_per_set(self.cos_asc_node, n_sets), _per_set(self.sin_asc_node, n_sets)], dim=-1)
            asc = asc.to(device=dev, dtype=torch.float64).contiguous()
        pos = torch.empty((3, n_sets) + t_shape, dtype=dt, device=dev)
        vel = torch.empty((3, n_sets) + t_shape, dtype=dt, device=dev)
        # (the reference takes include_inclination = (semiamplitude is None) AFTER the default
        # radial_velocity_semiamplitude has been filled in, keplerian.py:600-603, 634)
        flags = 0 if semiamplitude is None else _lib.JXP_STATE_NO_INCLINATION
        name = "jxp_orbit_state_f64" if dt == torch.float64 else "jxp_orbit_state_f32"
        _lib.call(name, rec.data_ptr(), n_sets, r0d.data_ptr(), 1, k0d.data_ptr(), 1, A.ptr(asc),
                  td.data_ptr(), td.numel(), flags, pos.data_ptr(), vel.data_ptr(), A.stream_ptr())
        if not batched:
            pos, vel = pos[:, 0], vel[:, 0]
        return (
            tuple(A.like_input(pos[i], t_in) for i in range(3)),
            tuple(A.like_input(vel[i], t_in) for i in range(3)),
        )


def _per_set(x, n: int) -> torch.Tensor:
    """Per-set constant (python scalar, 0-d or (n,) tensor) -> (n,) float64 tensor."""
    x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)
    x = x.detach().to(torch.float64)
    if x.ndim == 0:
        return x.expand(n)
    return x


def _cbrt(x: torch.Tensor) -> torch.Tensor:
    # torch has no cbrt; float64 numpy cbrt is correctly rounded-ish and this is per-set glue
    if x.device.type == "cpu":
        return torch.as_tensor(np.cbrt(x.detach().numpy()), dtype=x.dtype) if not x.requires_grad \
            else torch.sign(x) * torch.abs(x) ** (1.0 / 3.0)
    return torch.sign(x) * torch.abs(x) ** (1.0 / 3.0)


def _on(c, ref: torch.Tensor):
    """Per-set constant -> tensor on ref's device/dtype, shaped to broadcast against
    (n_sets, *t.shape) when batched."""
    if not isinstance(c, torch.Tensor):
        return c
    c = c.to(device=ref.device, dtype=ref.dtype)
    if c.ndim == 1 and A.batch_ndim():
        c = c.reshape((-1,) + (1,) * (ref.ndim - 1))
    return c


class System:
    """A Keplerian orbital system (keplerian.py:640-762)."""

    def __init__(self, central: Central | None = None, *, bodies: Iterable[Body | OrbitalBody] = ()):
        self.central = Central() if central is None else central
        self._bodies = tuple(
            b if isinstance(b, OrbitalBody) else OrbitalBody(self.central, b) for b in bodies
        )

    def __repr__(self) -> str:
        return f"System(central={self.central!r}, n_bodies={len(self._bodies)})"

    @property
    def shape(self) -> tuple[int, ...]:
        return (len(self._bodies),)

    @property
    def bodies(self) -> tuple[OrbitalBody, ...]:
        return self._bodies

    @property
    def radius(self):
        return self.body_vmap(lambda body: body.radius)()

    @property
    def central_radius(self):
        return self.body_vmap(lambda body: body.central_radius)()

    def add_body(self, body: Body | None = None, central: Central | None = None, **kwargs: Any):
        body_: Body | OrbitalBody | None = body
        if body_ is None:
            body_ = Body(**kwargs)
        if central is not None:
            body_ = OrbitalBody(central, body_)
        return System(central=self.central, bodies=self.bodies + (body_,))

    def body_vmap(self, func: Callable, in_axes: int | None | Sequence[Any] = 0, out_axes: Any = 0):
        """Map ``func(body, *args)`` over the bodies and stack the results on ``out_axes``
        (the reference's loop-and-stack path, object_stack.py:102-127)."""

        def mapped(*args):
            n_args = len(args)
            axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * n_args
            outs = []
            for i, body in enumerate(self._bodies):
                sel = [a if ax is None else _take(a, i, ax) for a, ax in zip(args, axes)]
                outs.append(func(body, *sel))
            return _stack_tree(outs, out_axes)

        return mapped

    def position(self, t):
        return self.body_vmap(OrbitalBody.position, in_axes=None)(t)

    def central_position(self, t):
        return self.body_vmap(OrbitalBody.central_position, in_axes=None)(t)

    def relative_position(self, t):
        return self.body_vmap(OrbitalBody.relative_position, in_axes=None)(t)

    def velocity(self, t):
        return self.body_vmap(OrbitalBody.velocity, in_axes=None)(t)

    def central_velocity(self, t):
        return self.body_vmap(OrbitalBody.central_velocity, in_axes=None)(t)

    def relative_velocity(self, t):
        return self.body_vmap(OrbitalBody.relative_velocity, in_axes=None)(t)

    def radial_velocity(self, t):
        return self.body_vmap(OrbitalBody.radial_velocity, in_axes=None)(t)


def _take(a, i, axis):
    if isinstance(a, torch.Tensor):
        return a.select(axis, i)
    return np.take(a, i, axis=axis)


def _stack_tree(outs, axis):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_stack_tree([o[k] for o in outs], axis) for k in range(len(first)))
    if isinstance(first, torch.Tensor):
        return torch.stack([o if isinstance(o, torch.Tensor) else torch.as_tensor(o) for o in outs], dim=axis)
    return np.stack([np.asarray(o) for o in outs], axis=axis)
