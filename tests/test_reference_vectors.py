"""Reference-produced vectors (tests/golden/reference_v1.npz).

They were made by executing the reference's own source files on a NumPy-backed jax shim
(tests/golden/make_reference_vectors.py, tests/golden/jax_shim/): reference outputs, not oracle
outputs.  CPU: the oracle reproduces them (this is what pins the oracle).  GPU: the CUDA path
reproduces them at the north-star tolerances, without the oracle in the loop.
"""

import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "reference_v1.npz"))

# the bodies make_reference_vectors.py builds (kept literal here: the GPU box has no /root/reference)
KW = [
    dict(period=3.456, time_transit=0.0, impact_param=0.3, radius=0.1, eccentricity=0.1, omega_peri=0.5),
    dict(period=7.77, time_transit=1.2, impact_param=0.6, radius=0.05, eccentricity=0.35, omega_peri=-2.1, asc_node=0.7),
    dict(period=1.2345, time_transit=0.31, impact_param=0.05, radius=0.02, eccentricity=0.0, omega_peri=0.0),
]
KW_CIRC = [
    dict(period=2.2, time_transit=0.4, impact_param=0.2, radius=0.08, mass=0.0004),
    dict(period=5.1, time_transit=-0.3, inclination=1.54, radius=0.12, mass=0.001),
]
KW_MIXED = [KW[0], dict(period=7.77, time_transit=1.2, impact_param=0.6, radius=0.05)]
DERIVED = ("semimajor", "time_ref", "sin_inclination", "cos_inclination", "impact_param", "sin_omega_peri", "cos_omega_peri")
EPS = np.finfo(np.float64).eps


def _flux_err(got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    return np.nanmax(np.abs(got - want) / np.maximum(1.0, np.abs(1 + want)))


# ----------------------------------------------------------------------------- CPU: oracle == reference
def test_oracle_kepler_bit_exact_with_reference():
    from oracle import ref_numpy as ref

    with np.errstate(all="ignore"):
        s, c = ref.kepler(G["kep_M"], G["kep_e"])
    np.testing.assert_array_equal(s, G["kep_sinf"])
    np.testing.assert_array_equal(c, G["kep_cosf"])


def test_oracle_kepler_jvp_rule_matches_reference_rule():
    from oracle import ref_numpy as ref

    M, e = G["kep_M"], G["kep_e"]
    ok = e < 0.999
    with np.errstate(all="ignore"):
        (_, _), (ds, dc) = ref.kepler_jvp(M, e, np.ones_like(M), np.zeros_like(M))
        np.testing.assert_allclose(ds[ok], G["kep_ds_dM"][ok], rtol=1e-14, atol=1e-14)
        np.testing.assert_allclose(dc[ok], G["kep_dc_dM"][ok], rtol=1e-14, atol=1e-14)
        (_, _), (ds, dc) = ref.kepler_jvp(M, e, np.zeros_like(M), np.ones_like(M))
        np.testing.assert_allclose(ds[ok], G["kep_ds_de"][ok], rtol=1e-14, atol=1e-14)
        np.testing.assert_allclose(dc[ok], G["kep_dc_de"][ok], rtol=1e-14, atol=1e-14)


def test_oracle_zero_safe_sqrt_rule_matches_reference_rule():
    import torch

    from oracle import ref_numpy as ref

    x = torch.tensor(G["zss_x"], requires_grad=True)
    y = ref.torch_backend().zero_safe_sqrt(x)
    (g,) = torch.autograd.grad(y.sum(), x)
    np.testing.assert_array_equal(y.detach().numpy(), G["zss_primal"])
    np.testing.assert_allclose(g.numpy(), G["zss_tangent"], rtol=1e-15, atol=0)


@pytest.mark.parametrize("n_u", [0, 1, 2, 3, 4, 6])
def test_oracle_limb_dark_matches_reference(n_u):
    from oracle import ref_numpy as ref

    got = ref.ld_light_curve(G["ld_u"][:n_u], G["ld_b"], G["ld_r"])
    # the only difference is the summation order of the final dot product s @ g
    assert _flux_err(got, G[f"ld_flux_u{n_u}"]) <= 4 * EPS


def test_oracle_limb_dark_order20_and_solution_vector():
    from oracle import ref_numpy as ref

    assert _flux_err(ref.ld_light_curve(G["ld_u"][:2], G["ld_b"], G["ld_r"], order=20), G["ld_flux_order20"]) <= 4 * EPS
    with np.errstate(all="ignore"):
        for l_max in (2, 4):
            b, r = np.broadcast_arrays(G["ld_b"], G["ld_r"])
            s = np.stack(ref.solution_vector(l_max, b.copy(), r.copy()), axis=-1)
            # identical up to the summation order of the quadrature (values are O(pi))
            np.testing.assert_allclose(s, G[f"ld_solution_l{l_max}"], rtol=0, atol=8 * EPS)


def test_oracle_orbit_constants_positions_velocities_match_reference():
    from oracle import ref_numpy as ref

    bodies = [ref.orbital_body(**kw) for kw in KW]
    for i, bd in enumerate(bodies):
        np.testing.assert_array_equal(np.array([bd[k] for k in DERIVED], dtype=float), G[f"body{i}_derived"])
    circ = [ref.orbital_body(central_mass=0.8, central_radius=0.75, **kw) for kw in KW_CIRC]
    for i, bd in enumerate(circ):
        np.testing.assert_array_equal(np.array([bd[k] for k in DERIVED[:5]], dtype=float), G[f"circ{i}_derived"])
    t = G["pos_t"]
    with np.errstate(all="ignore"):
        np.testing.assert_array_equal(np.stack([np.stack(ref.relative_position(b, t)) for b in bodies], -1), G["sys_relative_position"])
        np.testing.assert_array_equal(np.stack([np.stack(ref.relative_position(b, t)) for b in circ], -1), G["circ_relative_position"])
        for kind, key in (("relative", "relative_velocity"), ("", "velocity"), ("central", "central_velocity")):
            got = np.stack([np.stack(ref.velocity(b, t, kind)) for b in circ], -1)
            np.testing.assert_allclose(got, G[f"circ_{key}"], rtol=4 * EPS, atol=0)
        got = np.stack([np.stack(ref.velocity(b, t, "relative")) for b in bodies], -1)
        np.testing.assert_allclose(got, G["sys_relative_velocity"], rtol=4 * EPS, atol=1e-300)


def test_oracle_system_light_curves_match_reference():
    from oracle import ref_numpy as ref

    bodies = [ref.orbital_body(**kw) for kw in KW]
    u = G["sys_u"]
    assert _flux_err(ref.system_light_curve(bodies, u, G["sys_t"]), G["sys_flux"]) <= 4 * EPS
    assert np.mean(G["sys_flux"] != 0) > 0.1  # the vectors really are in transit
    assert _flux_err(ref.system_light_curve(bodies, G["ld_u"], G["sys6_t"]), G["sys6_flux"]) <= 4 * EPS
    circ = [ref.orbital_body(central_mass=0.8, central_radius=0.75, **kw) for kw in KW_CIRC]
    assert _flux_err(ref.system_light_curve(circ, u, G["circ_t"]), G["circ_flux"]) <= 4 * EPS
    mixed = [ref.orbital_body(**kw) for kw in KW_MIXED]
    assert _flux_err(ref.system_light_curve(mixed, u, G["sys_t"][::2]), G["mixed_flux"]) <= 4 * EPS
    for order, ns in ((0, 7), (1, 5), (2, 9), (0, 4)):
        f = ref.integrate(lambda tt: ref.system_light_curve(bodies, u, tt), 0.02, order, ns)
        assert _flux_err(f(G["int_t"]), G[f"int_flux_o{order}_n{ns}"]) <= 4 * EPS


# ----------------------------------------------------------------------------- GPU: CUDA path == reference
def _system(kws, central=None):
    from jaxoplanet_b200.orbits.keplerian import System

    s = System(central)
    for kw in kws:
        s = s.add_body(**kw)
    return s


@pytest.mark.gpu
def test_cuda_kepler_matches_reference_vectors():
    from jaxoplanet_b200.core import kepler

    M, e = G["kep_M"], G["kep_e"]
    s, c = kepler(M, e)
    lo = e <= 0.99
    # |d sin f|, |d cos f| <= 1e-12 (north star) where the reference algorithm itself is that accurate
    assert np.max(np.abs(s - G["kep_sinf"])[lo]) <= 1e-12 and np.max(np.abs(c - G["kep_cosf"])[lo]) <= 1e-12
    # true-anomaly angle difference scaled by the conditioning 1/(1-e) of the half-angle formula above e = 0.99
    ang = np.abs(np.arctan2(s, c) - np.arctan2(G["kep_sinf"], G["kep_cosf"]))
    ang = np.minimum(ang, 2 * np.pi - ang)
    assert np.max(ang * np.sqrt(1 - e)) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("n_u", [0, 1, 2, 3, 4, 6])
def test_cuda_limb_dark_matches_reference_vectors(n_u):
    from jaxoplanet_b200.core.limb_dark import light_curve

    got = light_curve(G["ld_u"][:n_u], G["ld_b"], G["ld_r"])
    assert _flux_err(got, G[f"ld_flux_u{n_u}"]) <= 1e-12
    if n_u == 2:
        assert _flux_err(light_curve(G["ld_u"][:2], G["ld_b"], G["ld_r"], order=20), G["ld_flux_order20"]) <= 1e-12


@pytest.mark.gpu
def test_cuda_system_light_curves_match_reference_vectors():
    from jaxoplanet_b200.light_curves import limb_dark_light_curve, transforms
    from jaxoplanet_b200.orbits.keplerian import Central

    system = _system(KW)
    u = G["sys_u"]
    lc = limb_dark_light_curve(system, u)
    assert _flux_err(lc(G["sys_t"]), G["sys_flux"]) <= 1e-12
    assert _flux_err(limb_dark_light_curve(system, G["ld_u"])(G["sys6_t"]), G["sys6_flux"]) <= 1e-12
    circ = _system(KW_CIRC, Central(mass=0.8, radius=0.75))
    assert _flux_err(limb_dark_light_curve(circ, u)(G["circ_t"]), G["circ_flux"]) <= 1e-12
    assert _flux_err(limb_dark_light_curve(_system(KW_MIXED), u)(G["sys_t"][::2]), G["mixed_flux"]) <= 1e-12
    for order, ns in ((0, 7), (1, 5), (2, 9), (0, 4)):
        f = transforms.integrate(lc, exposure_time=0.02, order=order, num_samples=ns)
        assert _flux_err(f(G["int_t"]), G[f"int_flux_o{order}_n{ns}"]) <= 1e-12


@pytest.mark.gpu
def test_cuda_orbit_positions_velocities_match_reference_vectors():
    from jaxoplanet_b200.orbits.keplerian import Central

    t = G["pos_t"]
    system = _system(KW)
    circ = _system(KW_CIRC, Central(mass=0.8, radius=0.75))
    for i, bd in enumerate(system.bodies):
        got = np.array([float(np.asarray(getattr(bd, k))) for k in DERIVED])
        np.testing.assert_allclose(got, G[f"body{i}_derived"], rtol=1e-14, atol=1e-15)

    def stacked(out):
        # body_vmap puts the body axis first, (n_bodies, n_t) for array t; the vectors were made one
        # scalar time at a time, i.e. (n_t, n_bodies)
        return np.stack([np.asarray(o).T for o in out])

    for name, sysm, meths in (
        ("sys", system, ("position", "relative_position", "velocity", "relative_velocity")),
        ("circ", circ, ("position", "central_position", "relative_position", "velocity", "central_velocity", "relative_velocity")),
    ):
        for meth in meths:
            want = G[f"{name}_{meth}"]
            got = stacked(getattr(sysm, meth)(t))
            assert got.shape == want.shape, (meth, got.shape, want.shape)
            scale = np.max(np.abs(want), axis=(0, 1), keepdims=True)
            assert np.max(np.abs(got - want) / scale) <= 1e-12, (name, meth)
    rv = np.asarray(circ.radial_velocity(t)).T
    assert np.max(np.abs(rv - G["circ_radial_velocity"])) <= 1e-12 * np.max(np.abs(G["circ_radial_velocity"]))


@pytest.mark.gpu
def test_cuda_transit_orbit_and_interpolate_match_reference_vectors():
    from jaxoplanet_b200.light_curves import limb_dark_light_curve, transforms
    from jaxoplanet_b200.orbits import TransitOrbit
    from jaxoplanet_b200.orbits.keplerian import Body, Central, OrbitalBody

    u = G["sys_u"]
    orb = TransitOrbit(period=3.3, duration=0.21, time_transit=0.4, impact_param=0.35, radius_ratio=0.09)
    assert abs(float(orb.speed) - float(G["tr_speed"])) <= 1e-14 * float(G["tr_speed"])
    assert _flux_err(limb_dark_light_curve(orb, u)(G["tr_t"]), G["tr_flux"]) <= 1e-12
    orb2 = TransitOrbit(period=2.1, speed=11.0, impact_param=0.8, radius_ratio=0.15)
    assert abs(float(orb2.duration) - float(G["tr2_duration"])) <= 1e-14
    assert _flux_err(limb_dark_light_curve(orb2, u)(G["tr_t"]), G["tr2_flux"]) <= 1e-12
    assert np.sum(G["tr_flux"] < 0) > 300
    single = OrbitalBody(Central(), Body(**KW[0]))
    fi = transforms.interpolate(limb_dark_light_curve(single, u), period=3.456, time_transit=0.0, num_samples=1501, duration=0.6)
    assert _flux_err(fi(G["itp_t"]), G["itp_flux"]) <= 1e-12


# ----------------------------------------------------------------------------- TTVOrbit (orbits/ttv.py)
GT = np.load(os.path.join(HERE, "golden", "reference_ttv_v1.npz"))


def test_ttv_orbit_host_setup_matches_reference():
    """Linear ephemeris fit, residuals, bins, expected transit times: host-side NumPy, no GPU needed."""
    from jaxoplanet_b200.orbits.ttv import TTVOrbit, compute_expected_transit_times

    orb = TTVOrbit(duration=np.array([0.15, 0.25]), impact_param=np.array([0.2, 0.55]),
                   radius_ratio=np.array([0.08, 0.12]), transit_times=(GT["ttv_tt0"], GT["ttv_tt1"]))
    np.testing.assert_allclose(orb.t0, GT["ttv_t0"], rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(orb.ttv_period, GT["ttv_period"], rtol=1e-14)
    np.testing.assert_allclose(orb.ttvs[0], GT["ttv_res0"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(orb.ttvs[1], GT["ttv_res1"], rtol=0, atol=1e-13)
    orb2 = TTVOrbit(period=4.2, time_transit=0.7, speed=9.0, impact_param=0.4, radius_ratio=0.1, ttvs=(GT["ttv2_ttvs"],))
    np.testing.assert_allclose(orb2.transit_times[0], GT["ttv2_transit_times"], rtol=1e-15, atol=1e-15)
    exp = compute_expected_transit_times(0.0, 30.0, np.array([3.1, 7.45]), np.array([1.3, 0.4]))
    np.testing.assert_array_equal(exp[0], GT["ttv_expected0"])
    np.testing.assert_array_equal(exp[1], GT["ttv_expected1"])
    with pytest.raises(ValueError, match="Supply either ttvs or transit_times, not both."):
        TTVOrbit(period=1.0, duration=0.1, ttvs=(np.zeros(3),), transit_times=(np.arange(3.0),))
    with pytest.raises(ValueError, match="You must supply either transit_times or ttvs."):
        TTVOrbit(period=1.0, duration=0.1)
    with pytest.raises(ValueError, match="When supplying ttvs, period must be provided."):
        TTVOrbit(duration=0.1, ttvs=(np.zeros(3),))


@pytest.mark.gpu
def test_cuda_ttv_orbit_matches_reference_vectors():
    from jaxoplanet_b200.light_curves import limb_dark_light_curve
    from jaxoplanet_b200.orbits import TTVOrbit

    u = GT["ttv_u"]
    orb = TTVOrbit(duration=np.array([0.15, 0.25]), impact_param=np.array([0.2, 0.55]),
                   radius_ratio=np.array([0.08, 0.12]), transit_times=(GT["ttv_tt0"], GT["ttv_tt1"]))
    got = limb_dark_light_curve(orb, u)(GT["ttv_t"])
    assert np.mean(GT["ttv_flux"] != 0) > 0.15
    assert _flux_err(got, GT["ttv_flux"]) <= 1e-12
    orb2 = TTVOrbit(period=4.2, time_transit=0.7, speed=9.0, impact_param=0.4, radius_ratio=0.1, ttvs=(GT["ttv2_ttvs"],))
    got2 = limb_dark_light_curve(orb2, u)(GT["ttv2_t"])
    assert _flux_err(got2, GT["ttv2_flux"]) <= 1e-12
    # the orbit protocol path (relative_position + K2) agrees with the fused kernel
    x, y, z = orb.relative_position(GT["ttv_t"])
    b = np.sqrt(x**2 + y**2)
    from jaxoplanet_b200.core.limb_dark import light_curve as core_lc

    want = np.where(z > 0, core_lc(u, b, np.broadcast_to(np.array([0.08, 0.12]), b.shape).copy()), 0.0)
    assert _flux_err(got, want) <= 1e-12


# ----------------------------------------------------------------------------- build container only
@pytest.mark.skipif(not os.path.isdir("/root/reference/src/jaxoplanet"), reason="reference checkout only exists in the build container")
def test_live_reference_source_on_shim_agrees_with_oracle_on_fresh_inputs():
    """Runs the reference's source on fresh random inputs (not the committed vectors)."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import reference_shim

    from oracle import ref_numpy as ref

    R = reference_shim.load()
    rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little"))
    with np.errstate(all="ignore"):
        M = rng.uniform(-50, 50, 500)
        e = rng.uniform(0, 0.999, 500)
        s, c = R.kepler.kepler(M, e)
        s0, c0 = ref.kepler(M, e)
        np.testing.assert_array_equal(np.asarray(s), s0)
        np.testing.assert_array_equal(np.asarray(c), c0)
        r = rng.uniform(0.01, 0.5, 300)
        b = rng.uniform(0, 1.1, 300) * (1 + r)
        u = rng.uniform(0, 0.4, 3)
        assert _flux_err(ref.ld_light_curve(u[:2], b, r), np.asarray(R.limb_dark.light_curve(u[:2], b, r))) <= 4 * EPS
        assert _flux_err(ref.ld_light_curve(u, b, r), np.asarray(R.limb_dark.light_curve(u, b, r))) <= 4 * EPS
        kw = dict(period=rng.uniform(2, 5), time_transit=rng.uniform(0, 2), impact_param=rng.uniform(0, 0.9),
                  radius=rng.uniform(0.03, 0.15), eccentricity=rng.uniform(0, 0.5), omega_peri=rng.uniform(-3, 3))
        K = R.keplerian
        system = K.System(K.Central()).add_body(**kw)
        t = kw["time_transit"] + rng.integers(-3, 4, 200) * kw["period"] + rng.uniform(-0.2, 0.2, 200)
        want = np.asarray(R.lc_limb_dark.light_curve(system, u[:2])(t))
        got = ref.system_light_curve([ref.orbital_body(**kw)], u[:2], t)
        assert _flux_err(got, want) <= 4 * EPS


# ----------------------------------------------------------------------------- emission light curve
GE = np.load(os.path.join(HERE, "golden", "reference_emission_v1.npz"))
EM_KW = [dict(period=3.456, time_transit=0.1, impact_param=0.3, radius=0.1, eccentricity=0.1, omega_peri=0.5),
         dict(period=1.7, time_transit=0.6, impact_param=0.1, radius=0.25, mass=0.002)]
EM_SURF = [((0.4, 0.1), 0.002), ((), 0.0005)]


def test_oracle_emission_light_curve_matches_reference():
    """emission.py executed unmodified on the shim (make_reference_vectors.py emission) vs the restatement."""
    from oracle import ref_numpy as ref

    bodies = [ref.orbital_body(central_mass=1.1, central_radius=0.9, **kw) for kw in EM_KW]
    with np.errstate(all="ignore"):
        got = ref.emission_light_curve(bodies, GE["em_star_u"], EM_SURF, GE["em_t"])
    assert got.shape == GE["em_flux"].shape
    assert np.max(np.abs(got - GE["em_flux"])) <= 4 * EPS


@pytest.mark.gpu
def test_cuda_emission_light_curve_matches_reference_vectors():
    from jaxoplanet_b200.light_curves.emission import Surface, SurfaceSystem, light_curve
    from jaxoplanet_b200.orbits.keplerian import Central

    sysm = SurfaceSystem(Central(mass=1.1, radius=0.9), Surface(u=(0.3, 0.2)))
    for kw, (u_b, amp) in zip(EM_KW, EM_SURF):
        sysm = sysm.add_body(surface=Surface(u=u_b, amplitude=amp), **kw)
    got = np.asarray(light_curve(sysm)(GE["em_t"]))
    want = GE["em_flux"]
    assert got.shape == want.shape
    # star flux to 1e-12; the companions' fluxes to 1e-12 of their amplitude
    scale = np.array([1.0, 0.002, 0.0005])
    assert np.max(np.abs(got - want) / scale) <= 1e-12
    assert np.sum(want[:, 0] < 1) > 100 and np.sum(want[:, 1] < 0.002 * (1 - 1e-6)) > 20


@pytest.mark.gpu
def test_cuda_batched_positions_and_velocities_one_kernel_matches_oracle():
    """OrbitalBody.position / velocity families of a BATCH of parameter sets (kernel K8, one launch per call)
    against the oracle, set by set; large batches take the device-side constants (kernel K0)."""
    import jaxoplanet_b200 as jxp
    from jaxoplanet_b200.orbits.keplerian import Central, System
    from oracle import ref_numpy as ref

    rng = np.random.default_rng(5)
    n = 300  # >= 256: OrbitalBody._init_on_device
    P = rng.uniform(1, 20, n)
    kw = dict(period=P, time_transit=rng.uniform(0, 1, n) * P, eccentricity=rng.uniform(0, 0.8, n),
              omega_peri=rng.uniform(-np.pi, np.pi, n), impact_param=rng.uniform(0, 1.2, n),
              radius=rng.uniform(0.01, 0.2, n), mass=rng.uniform(1e-5, 1e-2, n))
    t = np.sort(rng.uniform(-30, 60, 257))
    with jxp.batched():
        sysm = System(Central(mass=1.2, radius=0.8)).add_body(**kw)
        body = sysm.bodies[0]
        assert body._record is not None
        outs = {m: [np.asarray(c) for c in getattr(body, m)(t)]
                for m in ("position", "central_position", "relative_position", "velocity", "central_velocity",
                          "relative_velocity")}
        rv = np.asarray(body.radial_velocity(t))
        ip = np.asarray(body.impact_param.cpu()) if hasattr(body.impact_param, "cpu") else np.asarray(body.impact_param)
    np.testing.assert_allclose(ip, kw["impact_param"], rtol=0, atol=1e-14)
    for s in range(0, n, 37):
        bd = ref.orbital_body(central_mass=1.2, central_radius=0.8, **{k: v[s] for k, v in kw.items()})
        a = bd["semimajor"]
        mt = bd["mass"] + 1.2
        want = {
            "position": ref.position(bd, t, -a * 1.2 / mt),
            "central_position": ref.position(bd, t, a * bd["mass"] / mt),
            "relative_position": ref.relative_position(bd, t),
        }
        for m, w in want.items():
            for got, ww in zip(outs[m], w):
                assert got.shape == (n, t.size)
                np.testing.assert_allclose(got[s], ww, rtol=0, atol=1e-11 * a, err_msg=m)
        for m, kind in (("velocity", ""), ("central_velocity", "central"), ("relative_velocity", "relative")):
            w = ref.velocity(bd, t, kind)
            vs = np.max(np.abs(np.asarray(ref.velocity(bd, t, "relative"))))
            for got, ww in zip(outs[m], w):
                np.testing.assert_allclose(got[s], ww, rtol=0, atol=1e-11 * vs, err_msg=m)
        np.testing.assert_allclose(rv[s], -np.asarray(ref.velocity(bd, t, "central")[2]), rtol=0,
                                   atol=1e-11 * np.max(np.abs(rv[s])) + 1e-300)
