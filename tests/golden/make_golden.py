"""Generates tests/golden/hotpath_v1.npz from the CPU oracle (oracle/ref_numpy.py).

The reference itself (JAX) cannot run in this image, so these vectors are oracle outputs, not
reference outputs; the oracle is pinned to the reference's known-answer tests in
tests/test_oracle_kats.py.  Re-run:  python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_numpy as ref  # noqa: E402


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    # Kepler: random + the reference's edge cases (tests/core/kepler_test.py:54-64)
    M = np.concatenate([rng.uniform(-300, 300, 4000), [0.0, 2 * np.pi, np.pi, 1e-12, -226.2]])
    e = np.concatenate([rng.uniform(0, 0.99, 4000), [1 - 1e-6, 1 - 1e-6, 0.5, 0.3, 0.9939879759519037]])
    s, c, E = ref.kepler(M, e, return_E=True)
    out.update(kep_M=M, kep_e=e, kep_sinf=s, kep_cosf=c, kep_E=E)
    # limb darkening: grid of tests/core/limb_dark_test.py:16-23 (thinned) + edge list :30
    u = np.array([0.2, 0.3])
    bs, rs = [], []
    for r in (0.01, 0.1, 0.5, 1.1, 2.0):
        b = np.concatenate([np.linspace(-1 - 2 * r, 1 + 2 * r, 401), [0.0, 0.5, 1.0, r, abs(1 - r), 1 + r]])
        bs.append(b)
        rs.append(np.full_like(b, r))
    b = np.concatenate(bs)
    r = np.concatenate(rs)
    out.update(ld_u=u, ld_b=b, ld_r=r, ld_flux=ref.ld_light_curve(u, b, r),
               ld_flux_lin=ref.ld_light_curve(u[:1], b, r), ld_flux_uni=ref.ld_light_curve(u[:0], b, r))
    # system: config-1 planet + a circular companion, exposure integration, log-likelihood
    kw1 = dict(period=3.456, time_transit=0.0, impact_param=0.3, radius=0.1, eccentricity=0.1, omega_peri=0.5)
    kw2 = dict(period=7.77, time_transit=1.2, impact_param=0.6, radius=0.05)
    bodies = [ref.orbital_body(**kw1), ref.orbital_body(**kw2)]
    t = np.linspace(-0.5, 27.0, 6000)
    us = np.array([0.3, 0.2])
    flux = ref.system_light_curve(bodies, us, t)
    flux_int = ref.integrate(lambda tt: ref.system_light_curve(bodies, us, tt), 0.02, 0, 7)(t)
    y = 1 + flux.sum(-1) + 5e-4 * rng.normal(size=t.size)
    ll = ref.gaussian_loglike(1 + flux_int.sum(-1), y, np.full_like(t, 5e-4))
    out.update(sys_t=t, sys_u=us, sys_flux=flux, sys_flux_int=flux_int, sys_y=y, sys_loglike=ll)
    # partials (torch-autograd oracle), one parameter set
    from test_gpu_parity import _partials_oracle

    theta = np.array([3.456, 0.4, 0.17, 0.8, 0.45, 0.09, 0.35, 0.2])
    tp = np.arange(2500) * (2.0 / 1440.0)
    f, p = _partials_oracle(theta, tp)
    out.update(par_theta=theta, par_t=tp, par_flux=f, par_partials=p)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
