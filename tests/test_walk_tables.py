"""GPU tests of the per-set function tables of the transit walk (csrc/jxp_tables.cuh: k_walk_tables, the b^2 series
of k_walk_plan, k_walk_items, k_walk_fix).

With many samples per parameter set the walk does not solve Kepler's equation and run the ten-node quadrature at
every sample: per set it fits b^2 as a series in the mean anomaly across the transit window and the flux F(b) on
pieces graded towards the contact points, from node values of the SAME device code the direct path uses, and a
sample costs two Horner chains.  The fits must reproduce the direct code to its rounding noise; here they are
compared with the direct walk (JXP_DEV_NO_TABLES), the window-test kernels (JXP_DEV_NO_WALK) and the CPU oracle at
the north-star tolerances (flux 1e-12, log-likelihood 1e-12 relative), on the benchmark's parameter distribution, on
adversarial orbits, and on the cases the tables must refuse (large r, exposure stencils, unsorted time stamps,
short series)."""

import ctypes

import numpy as np
import pytest

from oracle import ref_numpy as ref

pytestmark = pytest.mark.gpu


def _call(p, u, t, y=None, sigma=None, opts=0, per_body=False, exposure=None, n_bodies=1, system=None):
    """jxp_system_light_curve_f64 through the C ABI -> (loglike or flux rows, list stats, walk stats)."""
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    dev = A.device()
    bodies, n_sets, _ = pack_bodies(system if system is not None else W.system_from(p))
    bodies = bodies.to(dev).contiguous()
    nb = int(bodies.shape[1])
    ud = torch.as_tensor(np.asarray(u, dtype=np.float64), device=dev).contiguous()
    u_stride = ud.shape[1] if ud.ndim == 2 else 0
    td = torch.as_tensor(np.asarray(t, dtype=np.float64), device=dev).contiguous()
    n_t = td.numel()
    lib = _lib.load()
    nbytes = int(lib.jxp_system_workspace_bytes(n_sets, nb, n_t))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ex = exposure or (0.0, 0, 0)
    with _lib.dev_options(opts | _lib.JXP_DEV_WALK_ANY_SIZE):
        if y is not None:
            yd = torch.as_tensor(np.asarray(y, dtype=np.float64), device=dev).contiguous()
            sg = np.atleast_1d(np.asarray(sigma, dtype=np.float64))
            sd = torch.as_tensor(sg, device=dev).contiguous()
            out = torch.full((n_sets,), float("nan"), dtype=torch.float64, device=dev)
            _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, nb, ud.data_ptr(), ud.shape[-1], u_stride,
                      td.data_ptr(), n_t, ex[0], ex[1], ex[2], 10, 0, None, None, yd.data_ptr(), sd.data_ptr(),
                      1 if sg.size > 1 else 0, out.data_ptr(), ws.data_ptr(), nbytes, A.stream_ptr())
        else:
            shape = (n_sets, n_t, nb) if per_body else (n_sets, n_t)
            out = torch.full(shape, float("nan"), dtype=torch.float64, device=dev)
            _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, nb, ud.data_ptr(), ud.shape[-1], u_stride,
                      td.data_ptr(), n_t, ex[0], ex[1], ex[2], 10, 0, out.data_ptr() if per_body else None,
                      None if per_body else out.data_ptr(), None, None, 0, None, ws.data_ptr(), nbytes, A.stream_ptr())
    ls = (ctypes.c_int64 * 4)()
    _lib.call("jxp_system_list_stats", ws.data_ptr(), n_sets, nb, n_t, ctypes.addressof(ls), A.stream_ptr())
    wk = (ctypes.c_int64 * 4)()
    _lib.call("jxp_system_walk_stats", ws.data_ptr(), n_sets, nb, n_t, ctypes.addressof(wk), A.stream_ptr())
    return (out.cpu().numpy(), dict(items=int(ls[0]), flags=int(ls[1])),
            dict(tab_evals=int(wk[0]), n_fix=int(wk[1]), tab_sets=int(wk[2]), no_tables=int(wk[3])))


def _ll_scale(ll, n_t, sigma):
    """The natural scale of a Gaussian log-likelihood: |value| + |constant term| (the constant and chi^2 / 2 nearly
    cancel for some sets, so |value| alone is not a scale)."""
    return np.abs(ll) + n_t * abs(np.log(sigma) + 0.5 * np.log(2 * np.pi))


def test_tables_match_direct_walk_window_kernels_and_oracle_on_config3_inputs():
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W

    p, ub, t, noise, sigma = W.c3(320, 50_000)
    truth = ref.system_light_curve([ref.orbital_body(**{k: v[3] for k, v in p.items()})], ub[3], t)[:, 0]
    y = 1 + truth + noise
    ll, ls, wk = _call(p, ub, t, y, sigma)
    ll_d, _, wk_d = _call(p, ub, t, y, sigma, opts=_lib.JXP_DEV_NO_TABLES)
    ll_w, ls_w, _ = _call(p, ub, t, y, sigma, opts=_lib.JXP_DEV_NO_WALK)
    assert ls["flags"] == 0 and ls_w["items"] == 0 and wk_d["tab_evals"] == 0 and wk_d["tab_sets"] == 0
    # (nearly) every set and every evaluation go through the tables; the rest -- samples within dx < 2^-13 of the
    # contact point b = 1 - r -- through the fix list
    assert wk["tab_sets"] >= 0.98 * 320 and wk["tab_evals"] > 0.9 * ls["items"] and 0 < wk["n_fix"] < 2e-3 * ls["items"]
    scale = _ll_scale(ll_d, t.size, sigma)
    assert np.max(np.abs(ll - ll_d) / scale) <= 1e-12 and np.max(np.abs(ll - ll_w) / scale) <= 1e-12
    for s in range(0, 320, 9):
        bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
        m = 1 + ref.system_light_curve([bd], ub[s], t).sum(-1)
        w = ref.gaussian_loglike(m, y, sigma * np.ones(t.size))
        assert abs(ll[s] - w) <= 1e-12 * (abs(w) + scale[s] - abs(ll_d[s]))
    # flux rows: the same zeros, values to the rounding noise of the direct code
    f, _, wkf = _call(p, ub, t)
    f_d, _, _ = _call(p, ub, t, opts=_lib.JXP_DEV_NO_TABLES)
    assert wkf["tab_evals"] > 0.9 * ls["items"]
    assert np.array_equal(f != 0, f_d != 0) and np.max(np.abs(f - f_d)) <= 2e-15
    for s in range(0, 320, 13):
        bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
        want = ref.system_light_curve([bd], ub[s], t)[:, 0]
        assert np.max(np.abs(f[s] - want)) <= 1e-14


def test_tables_are_bit_reproducible_and_shard_invariant():
    """Item partial sums are added in item order and the fix list in fixed point: repeated calls and shards of the
    set axis (what the ranks of a multi-GPU run evaluate) give the same bits."""
    from jaxoplanet_b200 import workloads as W

    p, ub, t, noise, sigma = W.c3(384, 30_000)
    y = 1 + noise
    base, _, wk = _call(p, ub, t, y, sigma)
    assert wk["tab_sets"] > 370 and wk["n_fix"] > 0
    for _ in range(3):
        assert np.array_equal(_call(p, ub, t, y, sigma)[0], base)
    for lo, hi in ((0, 96), (96, 384), (200, 201)):
        part = _call({k: v[lo:hi] for k, v in p.items()}, ub[lo:hi], t, y, sigma)[0]
        assert np.array_equal(part, base[lo:hi])
    f0 = _call(p, ub, t)[0]
    assert np.array_equal(_call(p, ub, t)[0], f0)


def test_tables_on_adversarial_orbits_irregular_times_and_per_sample_sigma():
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W

    rng = np.random.default_rng(77)
    n_sets = 300
    P = 10 ** rng.uniform(-0.5, 1.3, n_sets)
    p = dict(period=P, time_transit=rng.uniform(-1, 1, n_sets) * P, eccentricity=rng.uniform(0, 0.93, n_sets),
             omega_peri=rng.uniform(-np.pi, np.pi, n_sets), impact_param=rng.uniform(0, 1.4, n_sets),
             radius=10 ** rng.uniform(-2.3, -0.3, n_sets))
    u = W.kipping_u(rng.uniform(0, 1, (n_sets, 2)))
    t = np.sort(rng.uniform(-30, 30, 60_000))  # irregular cadence
    sigma = 5e-4 * rng.uniform(0.5, 2.0, t.size)
    y = 1 + sigma * rng.normal(size=t.size)
    f, ls, wk = _call(p, u, t)
    f_d, _, _ = _call(p, u, t, opts=_lib.JXP_DEV_NO_TABLES)
    assert ls["flags"] == 0 and wk["tab_sets"] > 100 and wk["tab_evals"] > 0
    assert np.array_equal(f != 0, f_d != 0) and np.max(np.abs(f - f_d)) <= 2e-14  # (depths up to 0.3 here)
    for s in range(0, n_sets, 17):
        bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
        want = ref.system_light_curve([bd], u[s], t)[:, 0]
        assert np.max(np.abs(f[s] - want)) <= 1e-12
    ll, _, wkl = _call(p, u, t, y, sigma)
    ll_d, _, _ = _call(p, u, t, y, sigma, opts=_lib.JXP_DEV_NO_TABLES)
    assert wkl["tab_sets"] == wk["tab_sets"]
    scale = np.abs(ll_d) + np.sum(np.abs(np.log(sigma) + 0.5 * np.log(2 * np.pi)))
    assert np.max(np.abs(ll - ll_d) / scale) <= 1e-12


def test_tables_are_refused_where_they_do_not_apply_and_serve_the_chunk_lists_too():
    """Large radius ratios and unsorted time stamps: the direct code does the work and the results are those of the
    direct walk / the window-test kernels bit for bit.  Exposure stencils and transits of a few samples: the chunk
    lists, with the tables."""
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W

    p, ub, t, noise, sigma = W.c3(96, 20_000)
    y = 1 + noise
    # r beyond the validated range
    big = dict(p)
    big["radius"] = np.linspace(0.51, 0.9, 96)
    ll, _, wk = _call(big, ub, t, y, sigma)
    ll_d, _, _ = _call(big, ub, t, y, sigma, opts=_lib.JXP_DEV_NO_TABLES)
    assert wk["tab_sets"] == 0 and wk["tab_evals"] == 0 and np.array_equal(ll, ll_d)
    # a mix: the sets in range use tables, the others do not, in one call
    mix = dict(p)
    mix["radius"] = np.where(np.arange(96) % 3 == 0, 0.6, p["radius"])
    ll, _, wk = _call(mix, ub, t, y, sigma)
    ll_d, _, _ = _call(mix, ub, t, y, sigma, opts=_lib.JXP_DEV_NO_TABLES)
    assert 55 <= wk["tab_sets"] <= 64 and np.array_equal(ll[::3], ll_d[::3])
    assert np.max(np.abs(ll - ll_d) / _ll_scale(ll_d, t.size, sigma)) <= 1e-12
    # exposure integration and transits of a few samples stay on the chunk lists (positions of many short transits
    # packed into chunks of 64) -- k_walk_eval reads the same tables there, per (sub-)sample
    f, ls, wk = _call(p, ub, t, exposure=(0.01, 0, 5))
    f_d, _, wk_d = _call(p, ub, t, exposure=(0.01, 0, 5), opts=_lib.JXP_DEV_NO_TABLES)
    assert wk["tab_evals"] > 4 * ls["items"] and wk_d["tab_evals"] == 0
    assert np.array_equal(f != 0, f_d != 0) and np.max(np.abs(f - f_d)) <= 2e-15
    want = ref.integrate(lambda tt: ref.system_light_curve([ref.orbital_body(**{k: v[5] for k, v in p.items()})], ub[5], tt),
                         0.01, 0, 5)(t)[:, 0]
    assert np.max(np.abs(f[5] - want)) <= 1e-14
    tc = np.arange(3000) * (30.0 / 1440.0)  # a cadence so coarse that a transit holds fewer than 12 samples
    f, _, wk = _call(p, ub, tc)
    f_d, _, _ = _call(p, ub, tc, opts=_lib.JXP_DEV_NO_TABLES)
    assert wk["tab_sets"] > 90 and wk["tab_evals"] > 0
    assert np.array_equal(f != 0, f_d != 0) and np.max(np.abs(f - f_d)) <= 2e-15
    # unsorted time stamps: the window-test kernels
    perm = np.random.default_rng(3).permutation(t.size)
    ll, ls, _ = _call(p, ub, t[perm], y[perm], sigma)
    ll_w, _, _ = _call(p, ub, t[perm], y[perm], sigma, opts=_lib.JXP_DEV_NO_WALK)
    assert ls["flags"] & 1 and np.array_equal(ll, ll_w)


def test_radius_ratios_near_the_end_of_the_table_range_and_two_bodies():
    """r up to 0.5: fits that do not converge make the whole set's pieces invalid and its samples take the fix list
    (or, when that overflows, the window-test kernels): values stay those of the direct code.  Two bodies: one plan
    + tables + evaluation per body, flux_sum in body order."""
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W

    p, ub, t, noise, sigma = W.c3(128, 30_000)
    y = 1 + noise
    hi = dict(p)
    hi["radius"] = np.linspace(0.2, 0.5, 128)
    hi["impact_param"] = p["impact_param"] * 0.8
    ll, ls, wk = _call(hi, ub, t, y, sigma)
    ll_d, _, _ = _call(hi, ub, t, y, sigma, opts=_lib.JXP_DEV_NO_TABLES)
    assert np.max(np.abs(ll - ll_d) / _ll_scale(ll_d, t.size, sigma)) <= 1e-12
    f, _, _ = _call(hi, ub, t)
    f_d, _, _ = _call(hi, ub, t, opts=_lib.JXP_DEV_NO_TABLES)
    assert np.array_equal(f != 0, f_d != 0) and np.max(np.abs(f - f_d)) <= 5e-14  # (depths up to 0.3)
    # two bodies per set
    import jaxoplanet_b200 as jxp
    from jaxoplanet_b200.orbits.keplerian import Central, System

    with jxp.batched():
        sys2 = System(Central()).add_body(**p)
        sys2 = sys2.add_body(period=p["period"] * 1.61, time_transit=p["time_transit"] + 0.3, radius=p["radius"] * 0.7,
                             impact_param=p["impact_param"] * 0.5)
    for per_body in (False, True):
        f, _, wk = _call(None, ub, t, per_body=per_body, system=sys2)
        f_d, _, _ = _call(None, ub, t, per_body=per_body, system=sys2, opts=_lib.JXP_DEV_NO_TABLES)
        assert wk["tab_evals"] > 0
        assert np.array_equal(f != 0, f_d != 0) and np.max(np.abs(f - f_d)) <= 3e-15


@pytest.mark.parametrize("n_u", [0, 1])
def test_tables_with_uniform_and_linear_limb_darkening(n_u):
    """The tables tabulate F(b) whatever the law: uniform disk (u = []) and linear limb darkening against the direct
    walk and the oracle."""
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W

    p, ub, t, noise, sigma = W.c3(64, 30_000)
    u = np.zeros((64, 0)) if n_u == 0 else ub[:, :1].copy()
    y = 1 + noise
    f, ls, wk = _call(p, u, t)
    f_d, _, _ = _call(p, u, t, opts=_lib.JXP_DEV_NO_TABLES)
    assert ls["flags"] == 0 and wk["tab_sets"] >= 60 and wk["tab_evals"] > 0.9 * ls["items"]
    assert np.array_equal(f != 0, f_d != 0) and np.max(np.abs(f - f_d)) <= 2e-15
    for s in (0, 17, 63):
        bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
        want = ref.system_light_curve([bd], u[s], t)[:, 0]
        assert np.max(np.abs(f[s] - want)) <= 1e-14
    ll, _, _ = _call(p, u, t, y, sigma)
    ll_d, _, _ = _call(p, u, t, y, sigma, opts=_lib.JXP_DEV_NO_TABLES)
    assert np.max(np.abs(ll - ll_d) / _ll_scale(ll_d, t.size, sigma)) <= 1e-12
