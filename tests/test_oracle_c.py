"""The C/OpenMP restatement (CPU baseline of bench.py) against the NumPy oracle.  CPU only."""

import numpy as np

from oracle import ref_c
from oracle import ref_numpy as ref


def test_c_kepler_matches_numpy():
    rng = np.random.default_rng(0)
    M = rng.uniform(-300, 300, 100_000)
    e = rng.uniform(0, 0.99, 100_000)
    s, c = ref_c.kepler(M, e)
    s0, c0 = ref.kepler(M, e)
    assert np.max(np.abs(s - s0)) < 1e-13 and np.max(np.abs(c - c0)) < 1e-13


def test_c_limb_dark_matches_numpy():
    rng = np.random.default_rng(1)
    r = rng.uniform(0.01, 0.3, 50_000)
    b = rng.uniform(0, 1.1, 50_000) * (1 + r)
    u = np.array([0.3, 0.2])
    assert np.max(np.abs(ref_c.ld_quad(u, b, r) - ref.ld_light_curve(u, b, r))) < 2e-15
    for rr in (0.01, 0.1, 0.5, 1.1, 2.0):
        bb = np.array([0.0, 0.5, 1.0, rr, abs(1 - rr), 1 + rr])
        assert np.max(np.abs(ref_c.ld_quad(u, bb, rr) - ref.ld_light_curve(u, bb, np.full_like(bb, rr)))) < 2e-15


def test_c_system_matches_numpy():
    rng = np.random.default_rng(2)
    n_sets, n_t = 3, 4000
    t = np.arange(n_t) * (10.0 / 1440.0)
    u = np.array([0.3, 0.2])
    recs = []
    bodies_np = []
    for s in range(n_sets):
        row, bl = [], []
        for k in range(2):
            kw = dict(period=rng.uniform(2, 5), time_transit=rng.uniform(0, 2), impact_param=rng.uniform(0, 0.9),
                      radius=rng.uniform(0.03, 0.12))
            if k == 0:
                kw.update(eccentricity=rng.uniform(0, 0.3), omega_peri=rng.uniform(-3, 3))
            bd = ref.orbital_body(**kw)
            bl.append(bd)
            circ = bd["eccentricity"] is None
            row.append([bd["period"], bd["time_transit"], bd["time_ref"], -1.0 if circ else bd["eccentricity"],
                        1.0 if circ else bd["cos_omega_peri"], 0.0 if circ else bd["sin_omega_peri"],
                        bd["cos_inclination"], bd["sin_inclination"], bd["semimajor"], 1.0, bd["radius"], 0.0])
        recs.append(row)
        bodies_np.append(bl)
    dt, w = ref.integrate_stencil(0.02, 0, 5)
    y = 1 + 1e-3 * rng.normal(size=n_t)
    flux, ll = ref_c.system(np.array(recs), u, t, dt=dt, w=w, y=y, sigma=1e-3)
    for s in range(n_sets):
        want = ref.integrate(lambda tt: ref.system_light_curve(bodies_np[s], u, tt), 0.02, 0, 5)(t).sum(-1)
        assert np.max(np.abs(flux[s] - want)) < 1e-14
        np.testing.assert_allclose(ll[s], ref.gaussian_loglike(1 + want, y, np.full(n_t, 1e-3)), rtol=1e-13)
    assert ref_c.max_threads() >= 1
