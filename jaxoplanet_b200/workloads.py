"""Seeded synthetic inputs for the BASELINE.json configs.  NumPy only; used by
bench.py, the tools and the tests so that every leg sees identical inputs."""

from __future__ import annotations

import numpy as np

TWO_MIN = 2.0 / 1440.0
THIRTY_MIN = 30.0 / 1440.0


def kipping_u(q):
    """Kipping (2013) q -> (u1, u2)."""
    return np.stack([2 * np.sqrt(q[..., 0]) * q[..., 1], np.sqrt(q[..., 0]) * (1 - 2 * q[..., 1])], -1)


def c1():
    """Config 1: single planet, e = 0.1, 1e5 samples (one TESS sector)."""
    body = dict(period=3.456, time_transit=0.0, impact_param=0.3, radius=0.1, eccentricity=0.1,
                omega_peri=0.5)
    return body, np.array([0.3, 0.2]), np.linspace(-0.5, 27.0, 100_000)


def c2(n=100_000_000, seed=0, wide=False):
    """Config 2: n (M, e) pairs, e ~ U[0, 0.99); wide: M ~ U[-300, 300] (range reduction)."""
    rng = np.random.default_rng(seed)
    M = rng.uniform(-300.0, 300.0, n) if wide else rng.uniform(0.0, 2 * np.pi, n)
    return M, rng.uniform(0.0, 0.99, n)


def transit_sets(n_sets, seed=1, stress=False):
    """Parameter sets around the docs/tutorials/transit.ipynb truth: dict of (n_sets,) arrays
    and u (n_sets, 2).  stress: a/R* small and b spread so that far more samples are in transit."""
    rng = np.random.default_rng(seed)
    P = rng.uniform(2.0, 5.0, n_sets)
    p = dict(
        period=P,
        time_transit=rng.uniform(0.0, 1.0, n_sets) * P,
        eccentricity=rng.uniform(0.0, 0.3, n_sets),
        omega_peri=rng.uniform(-np.pi, np.pi, n_sets),
        impact_param=rng.uniform(0.0, 1.0, n_sets),
        radius=rng.uniform(0.03, 0.15, n_sets),
    )
    u = kipping_u(rng.uniform(0.0, 1.0, (n_sets, 2)))
    return p, u


def theta_from(p, u):
    """(n_sets, 8) in the K5 order (period, t0, e, omega, b, r, u1, u2)."""
    return np.stack([p["period"], p["time_transit"], p["eccentricity"], p["omega_peri"],
                     p["impact_param"], p["radius"], u[:, 0], u[:, 1]], 1)


def c3(n_sets=4096, n_times=100_000, seed=1):
    """Config 3: n_sets x n_times 2-min cadence + Gaussian likelihood (sigma = 5e-4).
    y is 1 + noise here; bench/tests add the model of set 0 when they need a signal."""
    p, u = transit_sets(n_sets, seed)
    t = np.arange(n_times) * TWO_MIN
    noise = 5e-4 * np.random.default_rng(42).normal(size=n_times)
    return p, u, t, noise, 5e-4


def c4(n_sets=1024, n_times=200_000, seed=2):
    """Config 4: 3 planets, 30-min cadence, 15x order-0 supersampling."""
    rng = np.random.default_rng(seed)
    P = np.stack([rng.uniform(1, 3, n_sets), rng.uniform(4, 9, n_sets), rng.uniform(12, 30, n_sets)], 1)
    p = dict(
        period=P,
        time_transit=rng.uniform(0, 1, (n_sets, 3)) * P,
        eccentricity=rng.uniform(0, 0.2, (n_sets, 3)),
        omega_peri=rng.uniform(-np.pi, np.pi, (n_sets, 3)),
        impact_param=rng.uniform(0, 0.9, (n_sets, 3)),
        radius=rng.uniform(0.02, 0.1, (n_sets, 3)),
    )
    u = kipping_u(rng.uniform(0, 1, (n_sets, 2)))
    t = np.arange(n_times) * THIRTY_MIN
    return p, u, t, dict(exposure_time=THIRTY_MIN, order=0, num_samples=15)


def c5(n_sets=16384, n_times=50_000, seed=3):
    """Config 5: gradient batch, flux + 8 partials."""
    p, u = transit_sets(n_sets, seed)
    return theta_from(p, u), np.arange(n_times) * TWO_MIN


def system_from(p, n_bodies=None):
    """Build a (batched) facade System from a parameter dict; host-side only."""
    import jaxoplanet_b200 as jxp
    from jaxoplanet_b200.orbits.keplerian import Central, System

    with jxp.batched():
        sys_ = System(Central())
        if p["period"].ndim == 1:
            sys_ = sys_.add_body(**p)
        else:
            for k in range(p["period"].shape[1]):
                sys_ = sys_.add_body(**{name: v[:, k] for name, v in p.items()})
    return sys_
