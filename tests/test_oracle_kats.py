"""Pin the oracle against every self-contained known-answer test the reference's
own suite holds for the hot path.  CPU only.

Tolerance: the reference's own ``assert_allclose`` (src/jaxoplanet/test_utils.py:11-17)
= rtol 5e-7 (f64) / 5e-4 (f32) with jax's default atol (1e-5 f32 / 1e-15.. for f64 we
use atol 1e-12, stricter than needed).
"""

import numpy as np
import pytest

from oracle import ref_numpy as ref

RTOL64 = 5e-7


def _mean_true(e, E):
    # tests/core/kepler_test.py:12-18
    M = E - e * np.sin(E)
    f = 2 * np.arctan(np.sqrt((1 + e) / (1 - e)) * np.tan(0.5 * E))
    return M, f


def _check_kepler(e, E, rtol=RTOL64, atol=1e-12):
    # tests/core/kepler_test.py:21-27
    E = E % (2 * np.pi)
    M, f = _mean_true(e, E)
    s, c = ref.kepler(M, e)
    np.testing.assert_allclose(s, np.sin(f), rtol=rtol, atol=atol)
    np.testing.assert_allclose(c, np.cos(f), rtol=rtol, atol=atol)


def test_kepler_basic_low_e():
    # tests/core/kepler_test.py:30-35
    e = np.linspace(0, 0.85, 500)
    E = np.linspace(-300, 300, 1001)
    e = e[None, :] + np.zeros((E.shape[0], e.shape[0]))
    E = E[:, None] + np.zeros_like(e)
    _check_kepler(e, E)


def test_kepler_basic_low_e_f32():
    e = np.linspace(0, 0.85, 500, dtype=np.float32)
    E = np.linspace(-300, 300, 1001, dtype=np.float32)
    e = e[None, :] + np.zeros((E.shape[0], e.shape[0]), dtype=np.float32)
    E = E[:, None] + np.zeros_like(e)
    E = E % np.float32(2 * np.pi)
    M, f = _mean_true(e, E)
    s, c = ref.kepler(M, e)
    assert s.dtype == np.float32
    np.testing.assert_allclose(s, np.sin(f), rtol=5e-4, atol=5e-4)
    np.testing.assert_allclose(c, np.cos(f), rtol=5e-4, atol=5e-4)


def test_kepler_basic_high_e():
    # tests/core/kepler_test.py:42-47
    e = np.linspace(0.85, 1.0, 500)[:-1]
    E = np.linspace(-300, 300, 1001)
    e = e[None, :] + np.zeros((E.shape[0], e.shape[0]))
    E = E[:, None] + np.zeros_like(e)
    _check_kepler(e, E)


def test_kepler_edge_cases():
    # tests/core/kepler_test.py:54-58
    E = np.array([0.0, 2 * np.pi, -226.2, -170.4])
    e = (1 - 1e-6) * np.ones_like(E)
    e = np.append(e[:-1], 0.9939879759519037)
    _check_kepler(e, E)


def test_kepler_pi():
    # tests/core/kepler_test.py:61-64
    e = np.linspace(0, 1.0, 100)[:-1]
    E = np.pi + np.zeros_like(e)
    _check_kepler(e, E)


@pytest.mark.parametrize("e", [0.01, 0.1, 0.5, 0.6])
def test_kepler_grad(e):
    # tests/core/kepler_test.py:67-74 (check_grads order 1 == finite differences)
    M = np.linspace(-5, 5, 100)
    ee = np.full_like(M, e)
    (_, _), (ds, dc) = ref.kepler_jvp(M, ee, np.ones_like(M), np.zeros_like(M))
    h = 1e-6
    sp, cp = ref.kepler(M + h, ee)
    sm, cm = ref.kepler(M - h, ee)
    np.testing.assert_allclose(ds, (sp - sm) / (2 * h), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dc, (cp - cm) / (2 * h), rtol=1e-6, atol=1e-7)
    (_, _), (ds, dc) = ref.kepler_jvp(M, ee, np.zeros_like(M), np.ones_like(M))
    sp, cp = ref.kepler(M, ee + h)
    sm, cm = ref.kepler(M, ee - h)
    np.testing.assert_allclose(ds, (sp - sm) / (2 * h), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dc, (cp - cm) / (2 * h), rtol=1e-6, atol=1e-7)


def test_kepler_E_true_solution():
    # the internal E solves Kepler's equation (docs/tutorials/core-from-scratch.ipynb:188)
    rng = np.random.default_rng(0)
    M = rng.uniform(0, 2 * np.pi, 20000)
    e = rng.uniform(0, 0.99, 20000)
    _, _, E = ref.kepler(M, e, return_E=True)
    resid = E - e * np.sin(E) - M
    assert np.max(np.abs(resid / (1 - e * np.cos(E)))) < 2e-13


# ---- limb darkening ------------------------------------------------------
@pytest.mark.parametrize("u", [[0.2], [0.2, 0.3], [0.2, 0.3, 0.1, 0.5, 0.02]])
@pytest.mark.parametrize("r", [0.01, 0.1, 0.5, 1.1, 2.0])
def test_ld_edge_cases_finite(u, r):
    # tests/core/limb_dark_test.py:26-33 (finiteness part; exoplanet_core absent)
    u = np.array(u)
    for b in [0.0, 0.5, 1.0, r, np.abs(1 - r), 1 + r]:
        calc = ref.ld_light_curve(u, b, r)
        assert np.isfinite(calc)


@pytest.mark.parametrize("r", [0.01, 0.1, 0.5])
def test_ld_uniform_depth(r):
    # tests/light_curves/transit_test.py:9-37: u=(0,), min flux = -(Rp/R*)^2
    b = np.linspace(0, 1 - r, 50)
    lc = ref.ld_light_curve(np.array([0.0]), b, r)
    np.testing.assert_allclose(lc, -(r**2), rtol=RTOL64, atol=1e-14)
    lc0 = ref.ld_light_curve(np.array([]), b, r)
    np.testing.assert_allclose(lc0, -(r**2), rtol=RTOL64, atol=1e-14)


def test_ld_kappas_identity():
    # tests/starry/core/solution_test.py:13-38 (same formula as the hot-path kappas)
    r = 0.3
    b = np.linspace(abs(1 - r) + 1e-3, 1 + r - 1e-3, 200)
    area, k0, k1 = ref.kappas(b, np.full_like(b, r))
    # kappa0 = angle at the occultor centre, kappa1 at the star centre
    cosk0 = (b**2 + r**2 - 1) / (2 * b * r)
    cosk1 = (b**2 + 1 - r**2) / (2 * b)
    np.testing.assert_allclose(np.cos(k0), cosk0, atol=1e-12)
    np.testing.assert_allclose(np.cos(k1), cosk1, atol=1e-12)


def test_ld_matches_mpmath_truth():
    # independent truth: closed-form-free 2-D integral is too slow; use the
    # converged 1-D quadrature in mpmath (oracle/truth_mp.py), tolerance = the
    # reference's own 5e-7 vs independent codes (test_utils.py:11-17).
    from oracle import truth_mp

    u = np.array([0.2, 0.3])
    for r in [0.01, 0.1]:
        for b in [0.0, 0.3, 1 - r, 1.0, 1 + 0.5 * r]:
            got = float(ref.ld_light_curve(u, b, r))
            want = truth_mp.quad_ld_exact(u, b, r)
            assert abs(got - want) <= 5e-7 * max(abs(want), 1e-3), (r, b, got, want)


# ---- orbits --------------------------------------------------------------
SYSTEMS = [
    dict(central_mass=1.3, central_radius=1.1, mass=0.1, time_transit=0.1, period=12.5,
         inclination=0.3),
    dict(central_mass=1.3, central_radius=1.1, mass=0.1, time_transit=0.1, period=12.5,
         inclination=0.3, eccentricity=0.3, omega_peri=-1.5, asc_node=0.3),
]


def test_central_density():
    # tests/orbits/keplerian_test.py:80-86
    m, r, rho = ref.central()
    assert abs(rho * 5.905271918964842 - 1.4) < 0.01


def test_keplers_law():
    # tests/orbits/keplerian_test.py:101-108
    au = 215.03215567054764
    bd = ref.orbital_body(semimajor=au)
    assert abs(bd["period"] - 365.25) < 0.01
    bd = ref.orbital_body(period=365.25)
    assert abs(bd["semimajor"] - au) < 0.01


@pytest.mark.parametrize("kw", SYSTEMS)
def test_impact_parameter(kw):
    # tests/orbits/keplerian_test.py:134-142
    bd = ref.orbital_body(**kw)
    x, y, z = ref.relative_position(bd, np.asarray(bd["time_transit"]))
    np.testing.assert_allclose(
        np.sqrt(x**2 + y**2) / kw["central_radius"], bd["impact_param"], rtol=RTOL64
    )
    assert z > 0


@pytest.mark.parametrize("kw", SYSTEMS)
def test_velocity_is_position_derivative(kw):
    # tests/orbits/keplerian_test.py:111-125 (relative_ prefix), central differences
    bd = ref.orbital_body(**kw)
    t = np.linspace(-50.0, 50.0, 500)
    h = 1e-5
    v = ref.velocity(bd, t, "relative")
    pp = ref.relative_position(bd, t + h)
    pm = ref.relative_position(bd, t - h)
    for i in range(3):
        np.testing.assert_allclose(v[i], (pp[i] - pm[i]) / (2 * h), rtol=2e-6, atol=1e-6)


def test_circular_vs_e0():
    # tests/orbits/keplerian_test.py:326-349 atol 1e-12 (positions scaled by a)
    t = np.linspace(-50.0, 50.0, 500)
    b1 = ref.orbital_body(period=12.5, impact_param=0.3, radius=0.1)
    b2 = ref.orbital_body(period=12.5, impact_param=0.3, radius=0.1, eccentricity=0.0,
                          omega_peri=0.0)
    p1 = ref.relative_position(b1, t)
    p2 = ref.relative_position(b2, t)
    for a, b in zip(p1, p2):
        np.testing.assert_allclose(a, b, atol=1e-11 * b1["semimajor"])


def test_light_curve_shape_and_depth():
    # tests/light_curves/limb_dark_test.py:7-25 and transit_test.py:24-37
    bd = ref.orbital_body(period=300.456, radius=0.1, impact_param=0.8)
    t = np.linspace(-0.3, 0.3, 1000)
    lc = ref.system_light_curve([bd], np.array([0.3, 0.2]), t)
    assert lc.shape == t.shape + (1,)
    m, rs, _ = ref.central(mass=0.5, radius=0.5)
    bd = ref.orbital_body(central_mass=m, central_radius=rs, period=1.0, radius=0.1)
    t = np.linspace(-0.1, 0.1, 10)
    lc = ref.system_light_curve([bd], np.array([0.0]), t)
    np.testing.assert_allclose(lc.min(), -((0.1 / 0.5) ** 2), rtol=RTOL64)


@pytest.mark.parametrize("order", [0, 1, 2])
def test_integrate_invariant(order):
    # tests/light_curves/transforms_test.py:16-21
    def f1(t):
        return np.ones_like(t)

    def f2(t):
        return np.stack([0.5 * t + 0.1, -1.5 * t + 3.6], axis=-1)

    t = np.linspace(0, 10, 50)
    for f in (f1, f2):
        g = ref.integrate(f, exposure_time=0.1, order=order)
        np.testing.assert_allclose(g(t), f(t), rtol=RTOL64)


def test_torch_backend_matches_numpy_and_grads():
    import torch

    TH = ref.torch_backend()
    rng = np.random.default_rng(5)
    n = 200
    r = rng.uniform(0.02, 0.2, n)
    b = rng.uniform(0, 1.0, n) * (1 + r)
    u = np.array([0.3, 0.2])
    want = ref.ld_light_curve(u, b, r)
    bt = torch.tensor(b, requires_grad=True)
    rt = torch.tensor(r, requires_grad=True)
    ut = torch.tensor(u, requires_grad=True)
    got = ref.ld_light_curve(ut, bt, rt, xp=TH)
    np.testing.assert_allclose(got.detach().numpy(), want, rtol=0, atol=1e-14)
    (gb, gr) = torch.autograd.grad(got.sum(), (bt, rt))
    h = 1e-7
    fd_b = (ref.ld_light_curve(u, b + h, r) - ref.ld_light_curve(u, b - h, r)) / (2 * h)
    fd_r = (ref.ld_light_curve(u, b, r + h) - ref.ld_light_curve(u, b, r - h)) / (2 * h)
    ok = (np.abs(b - r) > 1e-3) & (np.abs(np.abs(b - r) - 1) > 1e-3) & (np.abs(b - 1 - r) > 1e-3) & (b > 1e-3)
    np.testing.assert_allclose(gb.numpy()[ok], fd_b[ok], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(gr.numpy()[ok], fd_r[ok], rtol=2e-5, atol=2e-7)
