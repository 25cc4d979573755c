"""Generates tests/golden/reference_v1.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES
(/root/reference/src/jaxoplanet/...) on the NumPy-backed jax shim (tests/golden/jax_shim).

These are reference outputs, not oracle outputs: nothing under oracle/ or jaxoplanet_b200/ is
imported here.  Limits of the pin are stated in jax_shim/jax/__init__.py (NumPy's libm instead of
XLA's, no derivatives).  Re-run in the build container:  python tests/golden/make_reference_vectors.py
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import reference_shim  # noqa: E402

# the named single-planet case and companions used by several vectors
KW = [
    dict(period=3.456, time_transit=0.0, impact_param=0.3, radius=0.1, eccentricity=0.1, omega_peri=0.5),
    dict(period=7.77, time_transit=1.2, impact_param=0.6, radius=0.05, eccentricity=0.35, omega_peri=-2.1, asc_node=0.7),
    dict(period=1.2345, time_transit=0.31, impact_param=0.05, radius=0.02, eccentricity=0.0, omega_peri=0.0),
]
KW_CIRC = [
    dict(period=2.2, time_transit=0.4, impact_param=0.2, radius=0.08, mass=0.0004),
    dict(period=5.1, time_transit=-0.3, inclination=1.54, radius=0.12, mass=0.001),
]


def main_ttv():
    """TTVOrbit (orbits/ttv.py:46-225) -> tests/golden/reference_ttv_v1.npz (kept separate so that
    reference_v1.npz does not have to be regenerated)."""
    R = reference_shim.load()
    rng = np.random.default_rng(20261021)
    out = {}
    us = np.array([0.3, 0.2])
    with np.errstate(all="ignore"):
        # two planets given by observed transit times (linear fit + residuals), ttv.py CASE 1
        tt0 = 1.3 + 3.1 * np.arange(9) + 0.02 * np.sin(np.arange(9) * 0.9)
        tt1 = 0.4 + 7.45 * np.arange(4) + np.array([0.01, -0.015, 0.02, -0.005])
        orb = R.ttv.TTVOrbit(duration=np.array([0.15, 0.25]), impact_param=np.array([0.2, 0.55]),
                             radius_ratio=np.array([0.08, 0.12]), transit_times=(tt0, tt1))
        t = np.sort(np.concatenate([np.linspace(-1.0, 29.0, 700),
                                    tt0[rng.integers(0, 9, 500)] + rng.uniform(-0.12, 0.12, 500),
                                    tt1[rng.integers(0, 4, 400)] + rng.uniform(-0.2, 0.2, 400)]))
        out.update(ttv_tt0=tt0, ttv_tt1=tt1, ttv_t=t, ttv_u=us,
                   ttv_flux=np.asarray(R.lc_limb_dark.light_curve(orb, us)(t)),
                   ttv_t0=np.asarray(orb.t0), ttv_period=np.asarray(orb.ttv_period),
                   ttv_res0=np.asarray(orb.ttvs[0]), ttv_res1=np.asarray(orb.ttvs[1]))
        # one planet given by ttvs + period + time_transit, CASE 2, with a speed instead of a duration
        ttvs = 0.03 * np.cos(np.arange(7) * 1.1)
        orb2 = R.ttv.TTVOrbit(period=4.2, time_transit=0.7, speed=9.0, impact_param=0.4, radius_ratio=0.1, ttvs=(ttvs,))
        t2 = np.sort(np.concatenate([np.linspace(-2.0, 28.0, 400),
                                     np.asarray(orb2.transit_times[0])[rng.integers(0, 7, 500)] + rng.uniform(-0.15, 0.15, 500)]))
        out.update(ttv2_ttvs=ttvs, ttv2_t=t2, ttv2_flux=np.asarray(R.lc_limb_dark.light_curve(orb2, us)(t2)),
                   ttv2_transit_times=np.asarray(orb2.transit_times[0]))
        exp = R.ttv.compute_expected_transit_times(0.0, 30.0, np.array([3.1, 7.45]), np.array([1.3, 0.4]))
        out.update(ttv_expected0=np.asarray(exp[0]), ttv_expected1=np.asarray(exp[1]))
    path = os.path.join(HERE, "reference_ttv_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays; in transit:", np.mean(out["ttv_flux"] != 0))


def main_emission():
    """light_curves/emission.py:11-52 -> tests/golden/reference_emission_v1.npz.  emission.py itself is
    executed unmodified; the `SurfaceSystem` it asserts on lives in starry (spherical-harmonic maps, outside
    this path), so a stand-in built on the reference's own `System` carries the three attributes the
    function reads (central_surface.u, body_surfaces[i].u / .amplitude)."""
    import importlib
    import types

    R = reference_shim.load()
    K = R.keplerian

    class Surface:
        def __init__(self, u=(), amplitude=1.0):
            self.u, self.amplitude = tuple(u), amplitude

    class SurfaceSystem(K.System):
        def __init__(self, central, central_surface, bodies_and_surfaces):
            K.System.__init__(self, central, bodies=[b for b, _ in bodies_and_surfaces])
            object.__setattr__(self, "central_surface", central_surface)
            object.__setattr__(self, "_surfaces", tuple(s for _, s in bodies_and_surfaces))

        @property
        def body_surfaces(self):
            return self._surfaces

    starry = types.ModuleType("jaxoplanet.starry")
    starry.__path__ = []
    orbit_mod = types.ModuleType("jaxoplanet.starry.orbit")
    orbit_mod.SurfaceSystem = SurfaceSystem
    sys.modules["jaxoplanet.starry"] = starry
    sys.modules["jaxoplanet.starry.orbit"] = orbit_mod
    em = importlib.import_module("jaxoplanet.light_curves.emission")
    kws = [dict(period=3.456, time_transit=0.1, impact_param=0.3, radius=0.1, eccentricity=0.1, omega_peri=0.5),
           dict(period=1.7, time_transit=0.6, impact_param=0.1, radius=0.25, mass=0.002)]
    surfaces = [Surface(u=(0.4, 0.1), amplitude=0.002), Surface(u=(), amplitude=0.0005)]
    central = K.Central(mass=1.1, radius=0.9)
    try:
        sysm = SurfaceSystem(central, Surface(u=(0.3, 0.2)), [(K.Body(**kw), s) for kw, s in zip(kws, surfaces)])
    except Exception:  # an equinox-style frozen module in the shim: build through the plain System and graft
        raise
    t = np.linspace(-1.0, 8.0, 4001)
    with np.errstate(all="ignore"):
        flux = np.asarray(em.light_curve(sysm)(t))
    out = dict(em_t=t, em_flux=flux, em_central=np.array([1.1, 0.9]), em_star_u=np.array([0.3, 0.2]),
               em_amp=np.array([s.amplitude for s in surfaces]))
    path = os.path.join(HERE, "reference_emission_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; shape", flux.shape, "occulted:", np.mean(flux[:, 1] < 0.002),
          np.mean(flux[:, 2] < 0.0005), "transit:", np.mean(flux[:, 0] < 1))


def main():
    R = reference_shim.load()
    K = R.keplerian
    rng = np.random.default_rng(20261020)
    out = {}
    with np.errstate(all="ignore"):
        # ---- core.kepler.kepler (core/kepler.py:14-56), random + the edge cases of tests/core/kepler_test.py
        M = np.concatenate([rng.uniform(-300, 300, 3000), rng.uniform(0, 2 * np.pi, 1000),
                            [0.0, 2 * np.pi, np.pi, 1e-12, -226.2, np.pi - 1e-9, np.pi + 1e-9]])
        e = np.concatenate([rng.uniform(0, 0.99, 4000), [1 - 1e-6, 1 - 1e-6, 0.5, 0.3, 0.9939879759519037, 0.7, 0.7]])
        s, c = R.kepler.kepler(M, e)
        out.update(kep_M=M, kep_e=e, kep_sinf=np.asarray(s), kep_cosf=np.asarray(c))
        # the reference's own JVP rule (core/kepler.py:59-79) evaluated on unit tangents
        (_, _), (ds_dM, dc_dM) = R.kepler._kepler.jvp_rule((M, e), (np.ones_like(M), np.zeros_like(M)))
        (_, _), (ds_de, dc_de) = R.kepler._kepler.jvp_rule((M, e), (np.zeros_like(M), np.ones_like(M)))
        out.update(kep_ds_dM=ds_dM, kep_dc_dM=dc_dM, kep_ds_de=ds_de, kep_dc_de=dc_de)
        # zero_safe_sqrt rule (utils.py:16-24)
        x = np.concatenate([[0.0, 1e-16, 2.2e-15, 2.3e-15, 1e-8], rng.uniform(0, 2, 20)])
        p, tgt = R.utils.zero_safe_sqrt.jvp_rule((x,), (np.ones_like(x),))
        out.update(zss_x=x, zss_primal=p, zss_tangent=tgt)

        # ---- core.limb_dark.light_curve (core/limb_dark.py:19-196): grid of tests/core/limb_dark_test.py + edges
        bs, rs = [], []
        for r in (0.01, 0.1, 0.5, 1.1, 2.0):
            b = np.concatenate([np.linspace(-1 - 2 * r, 1 + 2 * r, 301), [0.0, 0.5, 1.0, r, abs(1 - r), 1 + r, 1 - r, r - 1]])
            bs.append(b)
            rs.append(np.full_like(b, r))
        rr = rng.uniform(0.01, 0.3, 600)
        bs.append(rng.uniform(0, 1, 600) * (1 + rr))
        rs.append(rr)
        b = np.concatenate(bs)
        r = np.concatenate(rs)
        u_all = np.array([0.3, 0.2, 0.1, -0.05, 0.02, 0.04])
        out.update(ld_b=b, ld_r=r, ld_u=u_all)
        for n_u in (0, 1, 2, 3, 4, 6):
            out[f"ld_flux_u{n_u}"] = np.asarray(R.limb_dark.light_curve(u_all[:n_u], b, r))
        out["ld_flux_order20"] = np.asarray(R.limb_dark.light_curve(u_all[:2], b, r, order=20))
        # solution vectors (core/limb_dark.py:50-84)
        out["ld_solution_l2"] = np.asarray(R.limb_dark.solution_vector(2, order=10)(b, r))
        out["ld_solution_l4"] = np.asarray(R.limb_dark.solution_vector(4, order=10)(b, r))

        # ---- OrbitalBody derived constants (orbits/keplerian.py:262-340) and per-time outputs
        system = K.System(K.Central())
        for kw in KW:
            system = system.add_body(**kw)
        for i, bd in enumerate(system.bodies):
            out[f"body{i}_derived"] = np.array([float(np.asarray(getattr(bd, k))) for k in (
                "semimajor", "time_ref", "sin_inclination", "cos_inclination", "impact_param", "sin_omega_peri",
                "cos_omega_peri")])
        star = K.Central(mass=0.8, radius=0.75)
        system2 = K.System(star)
        for kw in KW_CIRC:
            system2 = system2.add_body(**kw)
        for i, bd in enumerate(system2.bodies):
            out[f"circ{i}_derived"] = np.array([float(np.asarray(getattr(bd, k))) for k in (
                "semimajor", "time_ref", "sin_inclination", "cos_inclination", "impact_param")])
        tpos = np.linspace(-4.0, 9.0, 257)
        for name, sysm in (("sys", system), ("circ", system2)):
            for meth in ("position", "central_position", "relative_position", "velocity", "central_velocity",
                         "relative_velocity"):
                if name == "sys" and meth.startswith("central"):
                    continue  # these bodies carry no mass: the reference raises TypeError there
                vals = [getattr(sysm, meth)(tt) for tt in tpos]
                out[f"{name}_{meth}"] = np.stack([np.stack([np.asarray(v[k]) for v in vals]) for k in range(3)])  # (3, n_t, n_bodies)
        out["circ_radial_velocity"] = np.stack([np.asarray(system2.radial_velocity(tt)) for tt in tpos])
        out["pos_t"] = tpos

        # ---- light_curves.limb_dark.light_curve(System, u)(t) (light_curves/limb_dark.py:15-62)
        us = np.array([0.3, 0.2])
        # one TESS-like stretch, sampled densely around transits so that most samples are in transit
        t = np.sort(np.concatenate([np.linspace(-0.5, 27.0, 1500),
                                    (rng.integers(0, 7, 900) * 3.456 + rng.uniform(-0.12, 0.12, 900)),
                                    (1.2 + rng.integers(0, 3, 500) * 7.77 + rng.uniform(-0.2, 0.2, 500)),
                                    (0.31 + rng.integers(0, 20, 300) * 1.2345 + rng.uniform(-0.06, 0.06, 300))]))
        lc = R.lc_limb_dark.light_curve(system, us)
        out.update(sys_t=t, sys_u=us, sys_flux=np.asarray(lc(t)))
        tc = t[:1200] * 0.4
        lc2 = R.lc_limb_dark.light_curve(system2, us)
        out.update(circ_t=tc, circ_flux=np.asarray(lc2(tc)))
        # eccentric + circular bodies: different pytree structure => ObjectStack's python-loop branch
        mixed = K.System(K.Central()).add_body(**KW[0]).add_body(period=7.77, time_transit=1.2, impact_param=0.6, radius=0.05)
        out.update(mixed_flux=np.asarray(R.lc_limb_dark.light_curve(mixed, us)(t[::2])))
        lc6 = R.lc_limb_dark.light_curve(system, u_all)
        t6 = t[::4]
        out.update(sys6_t=t6, sys6_flux=np.asarray(lc6(t6)))

        # ---- transforms.integrate (light_curves/transforms.py:21-116), all three orders
        ti = t[::3]
        out["int_t"] = ti
        for order, ns in ((0, 7), (1, 5), (2, 9), (0, 4)):
            f = R.transforms.integrate(lc, exposure_time=0.02, order=order, num_samples=ns)
            out[f"int_flux_o{order}_n{ns}"] = np.asarray(f(ti))

        # ---- TransitOrbit (orbits/transit.py:7-107)
        orb = R.transit.TransitOrbit(period=3.3, duration=0.21, time_transit=0.4, impact_param=0.35, radius_ratio=0.09)
        tt = np.sort(np.concatenate([np.linspace(-5, 12, 400), 0.4 + rng.integers(-1, 4, 600) * 3.3 + rng.uniform(-0.15, 0.15, 600)]))
        out.update(tr_t=tt, tr_flux=np.asarray(R.lc_limb_dark.light_curve(orb, us)(tt)),
                   tr_speed=np.float64(orb.speed))
        orb2 = R.transit.TransitOrbit(period=2.1, speed=11.0, impact_param=0.8, radius_ratio=0.15)
        out.update(tr2_flux=np.asarray(R.lc_limb_dark.light_curve(orb2, us)(tt)), tr2_duration=np.float64(orb2.duration))

        # ---- transforms.interpolate (light_curves/transforms.py:119-185) on the single-planet case
        # (jnp.interp needs a 1-D flux grid, so the orbit is a bare OrbitalBody, not a System)
        single = K.OrbitalBody(K.Central(), K.Body(**KW[0]))
        f1 = R.lc_limb_dark.light_curve(single, us)
        fi = R.transforms.interpolate(f1, period=3.456, time_transit=0.0, num_samples=1501, duration=0.6)
        tq = t[::5]
        out.update(itp_t=tq, itp_flux=np.asarray(fi(tq)))

    path = os.path.join(HERE, "reference_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ttv":
        main_ttv()
    elif len(sys.argv) > 1 and sys.argv[1] == "emission":
        main_emission()
    else:
        main()
