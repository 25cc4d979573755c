"""Randomised comparison of the transit-list kernels with the window-test kernels (both on the GPU; the latter
are the ones the parity tests pin to the oracle): log-likelihood, flux and gradient step on small random
problems -- irregular sorted time stamps, circular and eccentric bodies, grazing / non-transiting / huge
planets, tiny and huge periods.  usage: python tools/tl_fuzz.py [n_trials] [seed]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jaxoplanet_b200 as jxp
from jaxoplanet_b200.orbits.keplerian import Central, System
from jaxoplanet_b200.light_curves.limb_dark import log_likelihood, light_curve
from jaxoplanet_b200.light_curves.partials import transit_partials

n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
from jaxoplanet_b200 import _lib
_lib.load().jxp_set_dev_options(_lib.JXP_DEV_WALK_ANY_SIZE)
worst = dict(ll=0.0, flux=0.0, grad=0.0)
def both(fn):
    a = fn()
    with _lib.dev_options(_lib.JXP_DEV_NO_WALK):
        b = fn()
    return a, b
for trial in range(n_trials):
    n_sets = int(rng.integers(1, 40)); n_t = int(rng.integers(1, 4000))
    span = 10 ** rng.uniform(-1, 2.5)
    t = np.sort(rng.uniform(-span, span, n_t))
    if rng.uniform() < 0.3: t = np.round(t, 2)  # duplicates
    P = 10 ** rng.uniform(-1.5, 3, n_sets)
    p = dict(period=P, time_transit=rng.uniform(-1, 1, n_sets) * P, impact_param=rng.uniform(0, 1.6, n_sets),
             radius=10 ** rng.uniform(-2.5, -0.1, n_sets))
    if rng.uniform() < 0.7:
        p["eccentricity"] = rng.uniform(0, 0.95, n_sets) ** 2; p["omega_peri"] = rng.uniform(-np.pi, np.pi, n_sets)
    u = rng.uniform(0, 0.5, (n_sets, 2)) if rng.uniform() < 0.5 else rng.uniform(0, 0.5, 2)
    sigma = 10 ** rng.uniform(-4, -2)
    y = 1 + sigma * rng.normal(size=n_t)
    with jxp.batched():
        sys_ = System(Central(mass=rng.uniform(0.5, 1.5), radius=rng.uniform(0.5, 1.5))).add_body(**p)
    a, b = both(lambda: log_likelihood(sys_, u)(t, y, sigma))
    scale = np.abs(b) + n_t  # log-likelihoods can cancel to ~0: compare against the size of their terms
    worst["ll"] = max(worst["ll"], float(np.max(np.abs(a - b) / scale)))
    assert np.all(np.isfinite(a) == np.isfinite(b)), (trial, "ll finite")
    a, b = both(lambda: light_curve(sys_, u)(t))
    assert np.array_equal(a != 0, b != 0), (trial, "flux support")
    worst["flux"] = max(worst["flux"], float(np.max(np.abs(a - b))) if a.size else 0.0)
    if "eccentricity" in p and np.ndim(u) == 2:
        theta = np.stack([p["period"], p["time_transit"], p["eccentricity"], p["omega_peri"], p["impact_param"],
                          p["radius"], u[:, 0], u[:, 1]], 1)
        f = lambda: transit_partials(theta, t, y=y, sigma=sigma, want_flux=False, want_partials=False)
        a, b = both(f)
        ok = np.isfinite(b["dloglike"])
        assert np.array_equal(np.isfinite(a["dloglike"]), ok), (trial, "grad finite")
        sc = np.abs(b["dloglike"])[ok].max() if ok.any() else 1.0
        if ok.any():
            worst["grad"] = max(worst["grad"], float(np.max(np.abs(a["dloglike"][ok] - b["dloglike"][ok])) / max(sc, 1e-300)))
print("trials", n_trials, "worst |d loglike| / (|loglike| + n_t)", worst["ll"], "worst |d flux|", worst["flux"],
      "worst |d grad| / max|grad|", worst["grad"])
assert worst["ll"] < 1e-11 and worst["flux"] < 1e-13 and worst["grad"] < 1e-9
print("fuzz ok")
