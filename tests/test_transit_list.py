"""GPU tests of the transit-list path of the fused log-likelihood (k_tl_build / k_tl_eval /
k_tl_reduce, DESIGN.md "transit list"): sorted time stamps, one body, no exposure integration.

The list kernels and the window-test kernels (k_system_pair) are two different programs for the
same function; both are compared with the CPU oracle at the north-star tolerance, with each other,
and the device-side choice between them (sorted? fits the workspace?) is checked through
jxp_system_list_stats.
"""

import ctypes
import os

import numpy as np
import pytest

from oracle import ref_numpy as ref

pytestmark = pytest.mark.gpu

LAST_WALK = {}  # jxp_system_walk_stats of the last _loglike_abi / _flux_abi call (per-set function tables)


def _walk_stats(ws, n_sets, n_t):
    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib

    st = (ctypes.c_int64 * 4)()
    _lib.call("jxp_system_walk_stats", ws.data_ptr(), n_sets, 1, n_t, ctypes.addressof(st), A.stream_ptr())
    LAST_WALK.update(tab_evals=int(st[0]), n_fix=int(st[1]), tab_sets=int(st[2]), no_tables=int(st[3]))


def _loglike_abi(system, u, t, y, sigma):
    """Log-likelihood through the C ABI with a workspace we keep -> (loglike, list items, list flags)."""
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    dev = A.device()
    bodies, n_sets, _ = pack_bodies(system)
    bodies = bodies.to(dev).contiguous()
    u = torch.as_tensor(np.asarray(u, dtype=np.float64), device=dev).contiguous()
    u_stride = u.shape[1] if u.ndim == 2 else 0
    n_u = u.shape[-1]
    td = torch.as_tensor(np.asarray(t, dtype=np.float64), device=dev).contiguous()
    yd = torch.as_tensor(np.asarray(y, dtype=np.float64), device=dev).contiguous()
    sg = np.atleast_1d(np.asarray(sigma, dtype=np.float64))
    sd = torch.as_tensor(sg, device=dev).contiguous()
    n_t = td.numel()
    nbytes = _lib.load().jxp_system_workspace_bytes(n_sets, 1, n_t)
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    out = torch.full((n_sets,), float("nan"), dtype=torch.float64, device=dev)
    _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, 1, u.data_ptr(), n_u, u_stride,
              td.data_ptr(), n_t, 0.0, 0, 0, 10, 0, None, None, yd.data_ptr(), sd.data_ptr(),
              1 if sg.size > 1 else 0, out.data_ptr(), ws.data_ptr(), ws.numel(), A.stream_ptr())
    stats = (ctypes.c_int64 * 4)()
    _lib.call("jxp_system_list_stats", ws.data_ptr(), n_sets, 1, n_t, ctypes.addressof(stats), A.stream_ptr())
    _walk_stats(ws, n_sets, n_t)
    return out.cpu().numpy(), int(stats[0]), int(stats[1]), int(stats[2]), int(stats[3])


def _without_list(fn):
    from jaxoplanet_b200 import _lib

    with _lib.dev_options(_lib.JXP_DEV_NO_WALK):
        return fn()


def _oracle_loglike(p, u, t, y, sigma, sets):
    want = {}
    for s in sets:
        kw = {k: (None if v is None else v[s]) for k, v in p.items()}
        bd = ref.orbital_body(**{k: v for k, v in kw.items() if v is not None})
        m = 1 + ref.system_light_curve([bd], u if np.ndim(u) == 1 else u[s], t).sum(-1)
        want[s] = ref.gaussian_loglike(m, y, sigma * np.ones(t.size))
    return want


def _system(p, central=None):
    import jaxoplanet_b200 as jxp
    from jaxoplanet_b200.orbits.keplerian import Central, System

    with jxp.batched():
        return System(central or Central()).add_body(**p)


def test_list_path_matches_oracle_and_window_kernels_c3_like():
    from jaxoplanet_b200 import workloads as W

    p, ub, t, noise, sigma = W.c3(300, 40_000)
    truth = ref.system_light_curve([ref.orbital_body(**{k: v[0] for k, v in p.items()})], ub[0], t)[:, 0]
    y = 1 + truth + noise
    sys_ = _system(p)
    from jaxoplanet_b200 import _lib

    want = _oracle_loglike(p, ub, t, y, sigma, range(0, 300, 7))
    ll2, items2, flags2 = _without_list(lambda: _loglike_abi(sys_, ub, t, y, sigma))[:3]
    assert items2 == 0  # the list kernels were not launched
    # the walk with the per-set function tables (the default: csrc/jxp_tables.cuh) and with the direct Kepler +
    # quadrature code on the chunk lists
    lls = {}
    for tables in (True, False):
        with _lib.dev_options(0 if tables else _lib.JXP_DEV_NO_TABLES):
            ll, items, flags, inside, fast = _loglike_abi(sys_, ub, t, y, sigma)
        lls[tables] = ll
        assert flags == 0 and 0.01 * 300 * 40_000 < items < 0.08 * 300 * 40_000
        # different programs, different summation orders: terms of 1e-16 * 1 / sigma^2
        assert np.max(np.abs(ll - ll2) / np.abs(ll2)) <= 2e-12
        for s, w in want.items():
            assert abs(ll[s] - w) <= 1e-12 * abs(w)
        if tables:  # nearly every evaluation is answered by a table; the rest (next to b = 1 - r) by k_walk_fix
            assert LAST_WALK["tab_sets"] >= 0.98 * 300 and LAST_WALK["no_tables"] <= 6
            assert LAST_WALK["tab_evals"] > 0.9 * items and 0 < LAST_WALK["n_fix"] < 2e-3 * items
        else:
            assert LAST_WALK["tab_sets"] == 0 and LAST_WALK["tab_evals"] == 0
    assert inside > 0.5 * items and fast > 0.9 * (inside // 64)  # most of the list takes the inside-the-disk code
    # flux-level check of both flux codes: with y = 1 and a tiny sigma the likelihood is dominated by
    # sum (lc / sigma)^2, so its relative error is twice the (weighted) relative error of lc itself --
    # a systematic 1e-12 error of 1 + lc (the north-star bound) would show as ~2e-10 here
    sharp, _, fs = _loglike_abi(sys_, ub, t, np.ones_like(t), 1e-6)[:3]
    assert fs == 0
    sharp_w = _without_list(lambda: _loglike_abi(sys_, ub, t, np.ones_like(t), 1e-6))[0]
    assert np.max(np.abs(sharp - sharp_w) / np.abs(sharp_w)) <= 2e-12
    want = _oracle_loglike(p, ub, t, np.ones_like(t), 1e-6, range(0, 300, 13))
    worst = max(abs(sharp[s] - w) / abs(w) for s, w in want.items())
    assert worst <= 2e-12, worst
    # bit-reproducible although the slices of the list are handed out by atomics
    for _ in range(3):
        again, _, _ = _loglike_abi(sys_, ub, t, y, sigma)[:3]
        assert np.array_equal(again, lls[True])
        with _lib.dev_options(_lib.JXP_DEV_NO_TABLES):
            again, _, _ = _loglike_abi(sys_, ub, t, y, sigma)[:3]
        assert np.array_equal(again, lls[False])
    # per-sample sigma
    sig_t = sigma * (1 + 0.2 * np.random.default_rng(5).uniform(size=t.size))
    ll3, _, f3 = _loglike_abi(sys_, ub, t, y, sig_t)[:3]
    assert f3 == 0
    want = _oracle_loglike(p, ub, t, y, sig_t, [0, 11, 299])
    for s, w in want.items():
        assert abs(ll3[s] - w) <= 1e-12 * abs(w)


def test_list_path_adversarial_orbits_and_irregular_times():
    # short and long periods, e up to 0.93, grazing and non-transiting bodies, a non-default star;
    # irregular sorted time stamps with duplicates, gaps and negative times.  Bodies without a
    # provable window put every sample on the list, so the sets go in small batches (a list that
    # does not fit falls back, which the last test covers).
    from jaxoplanet_b200.orbits.keplerian import Central

    rng = np.random.default_rng(123)
    u = np.array([0.3, 0.2])
    t = np.sort(np.concatenate([rng.uniform(-40, -5, 9_000), rng.uniform(3, 30, 9_000)]))
    t[100:110] = t[100]
    t[-5:] = t[-5]
    y = 1 + 1e-3 * rng.normal(size=t.size)
    sigma = 1e-3
    central = Central(mass=0.8, radius=1.3)
    n_list = 0
    for batch in range(24):
        n_sets = 16
        P = 10 ** rng.uniform(-0.7, 1.6, n_sets)
        p = dict(period=P, time_transit=rng.uniform(-1, 1, n_sets) * P,
                 eccentricity=rng.uniform(0, 0.93, n_sets), omega_peri=rng.uniform(-np.pi, np.pi, n_sets),
                 impact_param=rng.uniform(0, 1.4, n_sets), radius=10 ** rng.uniform(-2.3, -0.5, n_sets))
        sys_ = _system(p, central)
        ll, items, flags = _loglike_abi(sys_, u, t, y, sigma)[:3]
        assert flags in (0, 2)
        n_list += flags == 0
        ll2, _, _ = _without_list(lambda: _loglike_abi(sys_, u, t, y, sigma))[:3]
        assert np.max(np.abs(ll - ll2) / np.abs(ll2)) <= 2e-12
        for s in range(batch % 3, n_sets, 3):
            bd = ref.orbital_body(central_mass=0.8, central_radius=1.3, **{k: v[s] for k, v in p.items()})
            m = 1 + ref.system_light_curve([bd], u, t).sum(-1)
            w = ref.gaussian_loglike(m, y, sigma * np.ones(t.size))
            assert abs(ll[s] - w) <= 1e-12 * abs(w)
    assert n_list >= 12  # most batches went through the list kernels


def test_list_path_circular_bodies_and_single_sample():
    rng = np.random.default_rng(9)
    n_sets = 64
    P = rng.uniform(0.8, 12, n_sets)
    p = dict(period=P, time_transit=rng.uniform(0, 1, n_sets) * P, impact_param=rng.uniform(0, 1.1, n_sets),
             radius=rng.uniform(0.02, 0.2, n_sets))
    u = np.array([0.4, 0.1])
    t = np.arange(15_000) * (10.0 / 1440.0) - 20.0
    y = 1 + 2e-4 * rng.normal(size=t.size)
    sys_ = _system(p)
    ll, items, flags = _loglike_abi(sys_, u, t, y, 2e-4)[:3]
    assert flags == 0 and items > 0
    want = _oracle_loglike(p, u, t, y, 2e-4, range(0, n_sets, 5))
    for s, w in want.items():
        assert abs(ll[s] - w) <= 1e-12 * abs(w)
    # one time stamp: the list has at most n_sets items
    ll1, items1, flags1 = _loglike_abi(sys_, u, t[7000:7001], y[7000:7001], 2e-4)[:3]
    assert flags1 == 0 and items1 <= n_sets
    want = _oracle_loglike(p, u, t[7000:7001], y[7000:7001], 2e-4, range(0, n_sets, 5))
    for s, w in want.items():
        assert abs(ll1[s] - w) <= 1e-12 * abs(w)


def test_unsorted_or_nan_times_fall_back_on_the_device():
    from jaxoplanet_b200 import workloads as W

    p, ub, t, noise, sigma = W.c3(96, 12_000)
    y = 1 + noise
    sys_ = _system(p)
    rng = np.random.default_rng(3)
    perm = rng.permutation(t.size)
    tu, yu = t[perm], y[perm]
    ll, items, flags = _loglike_abi(sys_, ub, tu, yu, sigma)[:3]
    assert flags & 1
    ll2, _, _ = _without_list(lambda: _loglike_abi(sys_, ub, tu, yu, sigma))[:3]
    assert np.array_equal(ll, ll2)  # the same kernels produced both
    want = _oracle_loglike(p, ub, tu, yu, sigma, [0, 50, 95])
    for s, w in want.items():
        assert abs(ll[s] - w) <= 1e-12 * abs(w)
    # a single inversion at the very end is enough
    t1 = t.copy()
    t1[-1] = t1[-2] - 1e-9
    _, _, f1 = _loglike_abi(sys_, ub, t1, y, sigma)[:3]
    assert f1 & 1
    # a NaN time stamp breaks the binary searches: not sorted
    t2 = t.copy()
    t2[777] = np.nan
    _, _, f3 = _loglike_abi(sys_, ub, t2, y, sigma)[:3]
    assert f3 & 1


def test_list_overflow_falls_back_and_always_bodies_fit_when_few():
    # e >= 0.99 bodies have no provable window (BODY_ALWAYS): every sample goes on the list
    rng = np.random.default_rng(11)
    n_t = 30_000
    t = np.arange(n_t) * (2.0 / 1440.0)
    y = 1 + 5e-4 * rng.normal(size=n_t)
    u = np.array([0.3, 0.2])

    def sets(n_sets, n_always):
        P = rng.uniform(2, 5, n_sets)
        e = rng.uniform(0, 0.3, n_sets)
        e[:n_always] = 0.992
        return dict(period=P, time_transit=rng.uniform(0, 1, n_sets) * P, eccentricity=e,
                    omega_peri=rng.uniform(-np.pi, np.pi, n_sets), impact_param=rng.uniform(0, 1, n_sets),
                    radius=rng.uniform(0.03, 0.15, n_sets))

    p = sets(64, 64)  # 64 x 30000 / 64 chunks > room for 1/8 of the grid in chunks of 32 + 1152
    sys_ = _system(p)
    ll, items, flags = _loglike_abi(sys_, u, t, y, 5e-4)[:3]
    assert flags == 2
    ll2, _, _ = _without_list(lambda: _loglike_abi(sys_, u, t, y, 5e-4))[:3]
    assert np.array_equal(ll, ll2)
    p = sets(64, 2)  # two full-length slices fit
    sys_ = _system(p)
    ll, items, flags = _loglike_abi(sys_, u, t, y, 5e-4)[:3]
    assert flags == 0 and items >= 2 * n_t
    want = _oracle_loglike(p, u, t, y, 5e-4, [0, 1, 2, 33, 63])
    for s, w in want.items():
        assert abs(ll[s] - w) <= 1e-12 * abs(w)


def test_list_path_no_out_of_bounds_writes():
    # sentinels around the workspace and the output; ragged sizes
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    dev = A.device()
    for n_sets, n_t in ((1, 1), (3, 77), (33, 4099), (130, 10_001)):
        p, ub, t, noise, sigma = W.c3(n_sets, n_t)
        bodies, _, _ = pack_bodies(_system(p))
        bodies = bodies.to(dev).contiguous()
        u = torch.as_tensor(ub, device=dev).contiguous()
        td = torch.as_tensor(t, device=dev)
        yd = torch.as_tensor(1 + noise, device=dev)
        sd = torch.full((1,), sigma, dtype=torch.float64, device=dev)
        nbytes = int(_lib.load().jxp_system_workspace_bytes(n_sets, 1, n_t))
        guard = 4096
        buf = torch.full((nbytes + 2 * guard,), 0xA5, dtype=torch.uint8, device=dev)
        ws = buf[guard:guard + nbytes]
        assert ws.data_ptr() % 256 == 0
        out = torch.full((n_sets + 16,), -7.0, dtype=torch.float64, device=dev)
        _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, 1, u.data_ptr(), 2, 2,
                  td.data_ptr(), n_t, 0.0, 0, 0, 10, 0, None, None, yd.data_ptr(), sd.data_ptr(), 0,
                  out[8:].data_ptr(), ws.data_ptr(), nbytes, A.stream_ptr())
        torch.cuda.synchronize()
        assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + nbytes:] == 0xA5).all())
        assert bool((out[:8] == -7.0).all()) and bool((out[8 + n_sets:] == -7.0).all())
        assert bool(torch.isfinite(out[8:8 + n_sets]).all())


# ----------------------------------------------------------------------------- gradient step (K5)
def _grad_step(theta, t, y, sigma):
    from jaxoplanet_b200.light_curves.partials import transit_partials

    r = transit_partials(theta, t, y=y, sigma=sigma, want_flux=False, want_partials=False)
    return r["loglike"], r["dloglike"]


def test_gradient_step_list_path_matches_rows_reduction_window_kernels_and_autograd_oracle():
    """Log-likelihood + 8-gradient without stored rows goes through k_tl_build / k_tlg_eval /
    k_tlg_reduce (sorted time stamps).  Checked against (1) the reductions of the per-point flux and
    partials the row-storing kernel writes -- which tests/test_gpu_parity.py checks point by point against
    the torch-autograd oracle -- at the north-star tolerance for gradients (1e-9 relative), and (2) the
    window-test kernels (JXP_NO_TL)."""
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.partials import transit_partials

    n_sets, n_t = 200, 30_000
    theta, t = W.c5(n_sets, n_t)
    rows = transit_partials(theta, t)  # flux (n_sets, n_t), partials (n_sets, 8, n_t)
    rng = np.random.default_rng(8)
    y = 1 + rows["flux"][3] + 5e-4 * rng.normal(size=n_t)
    sigma = 5e-4 * (1 + 0.2 * rng.uniform(size=n_t))
    ll, g = _grad_step(theta, t, y, sigma)
    assert ll.shape == (n_sets,) and g.shape == (n_sets, 8)
    # (1) reductions of the stored rows (different kernel, general flux code everywhere)
    m = 1 + rows["flux"]
    want_ll = ref.gaussian_loglike(m, y[None, :], sigma[None, :])
    want_g = np.einsum("st,skt->sk", (y[None, :] - m) / sigma[None, :] ** 2, rows["partials"])
    scale = np.abs(want_g).max(0, keepdims=True)
    assert np.max(np.abs(ll - want_ll) / np.abs(want_ll)) <= 1e-11
    assert np.max(np.abs(g - want_g) / np.maximum(np.abs(want_g), scale)) <= 1e-9
    # (2) the window-test kernels
    ll_w, g_w = _without_list(lambda: _grad_step(theta, t, y, sigma))
    assert np.max(np.abs(ll - ll_w) / np.abs(ll_w)) <= 1e-11
    assert np.max(np.abs(g - g_w) / np.maximum(np.abs(g_w), scale)) <= 1e-11
    # bit-reproducible
    for _ in range(2):
        ll2, g2 = _grad_step(theta, t, y, sigma)
        assert np.array_equal(ll2, ll) and np.array_equal(g2, g)


def _grad_step_abi(theta, t, y, sigma):
    """The gradient step through the C ABI with a workspace we keep -> (loglike, dloglike, list stats)."""
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib

    dev = A.device()
    th = torch.as_tensor(theta, device=dev).contiguous()
    td = torch.as_tensor(t, device=dev)
    yd = torch.as_tensor(y, device=dev)
    sd = torch.full((1,), sigma, dtype=torch.float64, device=dev)
    star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
    ns, nt = th.shape[0], td.numel()
    wsb = int(_lib.load().jxp_transit_workspace_bytes(ns, nt))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ll = torch.empty(ns, dtype=torch.float64, device=dev)
    dll = torch.empty((ns, 8), dtype=torch.float64, device=dev)
    _lib.call("jxp_transit_partials_f64", th.data_ptr(), ns, star.data_ptr(), 0, td.data_ptr(), nt, 0.0, 0, 0,
              10, 0, None, None, yd.data_ptr(), sd.data_ptr(), 0, ll.data_ptr(), dll.data_ptr(), ws.data_ptr(), wsb,
              A.stream_ptr())
    st = (ctypes.c_int64 * 4)()
    _lib.call("jxp_transit_list_stats", ws.data_ptr(), ns, nt, ctypes.addressof(st), A.stream_ptr())
    return ll.cpu().numpy(), dll.cpu().numpy(), [int(v) for v in st]


def test_gradient_step_uses_the_list_and_reports_it():
    from jaxoplanet_b200 import workloads as W

    theta, t = W.c5(128, 20_000)
    y = 1 + 5e-4 * np.random.default_rng(1).normal(size=t.size)
    ll, g, st = _grad_step_abi(theta, t, y, 5e-4)
    assert st[1] == 0 and 0.01 * 128 * 20_000 < st[0] < 0.06 * 128 * 20_000 and st[2] > 0.5 * st[0]
    ll2, g2 = _grad_step(theta, t, y, 5e-4)
    assert np.array_equal(ll, ll2) and np.array_equal(g, g2)
    perm = np.random.default_rng(2).permutation(t.size)
    _, _, st = _grad_step_abi(theta, t[perm], y[perm], 5e-4)
    assert st[1] & 1  # unsorted: the window-test kernels did the work
    _, _, st = _without_list(lambda: _grad_step_abi(theta, t, y, 5e-4))
    assert st[0] == 0


def test_gradient_step_falls_back_on_unsorted_times():
    from jaxoplanet_b200 import workloads as W

    theta, t = W.c5(64, 9_000)
    rng = np.random.default_rng(2)
    perm = rng.permutation(t.size)
    y = 1 + 5e-4 * rng.normal(size=t.size)
    ll, g = _grad_step(theta, t[perm], y[perm], 5e-4)
    ll_w, g_w = _without_list(lambda: _grad_step(theta, t[perm], y[perm], 5e-4))
    assert np.array_equal(ll, ll_w) and np.array_equal(g, g_w)  # the same kernels produced both
    ll_s, g_s = _grad_step(theta, t, y, 5e-4)
    assert np.max(np.abs(ll - ll_s) / np.abs(ll_s)) <= 1e-11
    scale = np.abs(g_s).max(0, keepdims=True)
    assert np.max(np.abs(g - g_s) / np.maximum(np.abs(g_s), scale)) <= 1e-11


# ----------------------------------------------------------------------------- flux outputs
def _flux_abi(system, u, t, per_body=False):
    """light_curve(orbit, u)(t) of a batch of single-planet systems through the C ABI -> (flux, list stats)."""
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    dev = A.device()
    bodies, n_sets, _ = pack_bodies(system)
    bodies = bodies.to(dev).contiguous()
    ud = torch.as_tensor(np.asarray(u, dtype=np.float64), device=dev).contiguous()
    u_stride = ud.shape[1] if ud.ndim == 2 else 0
    td = torch.as_tensor(np.asarray(t, dtype=np.float64), device=dev).contiguous()
    n_t = td.numel()
    nbytes = int(_lib.load().jxp_system_workspace_bytes(n_sets, 1, n_t))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.full((n_sets, n_t), float("nan"), dtype=torch.float64, device=dev)
    _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, 1, ud.data_ptr(), ud.shape[-1], u_stride,
              td.data_ptr(), n_t, 0.0, 0, 0, 10, 0, out.data_ptr() if per_body else None,
              None if per_body else out.data_ptr(), None, None, 0, None, ws.data_ptr(), nbytes, A.stream_ptr())
    st = (ctypes.c_int64 * 4)()
    _lib.call("jxp_system_list_stats", ws.data_ptr(), n_sets, 1, n_t, ctypes.addressof(st), A.stream_ptr())
    _walk_stats(ws, n_sets, n_t)
    return out.cpu().numpy(), [int(v) for v in st]


def test_flux_outputs_take_the_list_and_match_oracle_and_window_kernels():
    from jaxoplanet_b200 import workloads as W

    p, ub, t, _, _ = W.c3(150, 25_000)
    sys_ = _system(p)
    from jaxoplanet_b200 import _lib

    for per_body, tables in ((False, True), (True, True), (False, False), (True, False)):
        with _lib.dev_options(0 if tables else _lib.JXP_DEV_NO_TABLES):
            f, st = _flux_abi(sys_, ub, t, per_body)
        assert st[1] == 0 and st[0] > 0
        if tables:  # per-set function tables: the samples are items of whole transits, not list positions
            assert LAST_WALK["tab_sets"] >= 0.98 * 150 and LAST_WALK["tab_evals"] > 0.9 * st[0]
        else:
            assert st[2] > 0.5 * st[0] and LAST_WALK["tab_evals"] == 0
        fw, stw = _without_list(lambda: _flux_abi(sys_, ub, t, per_body))
        assert stw[0] == 0
        # (the tables reproduce the direct code to its own rounding noise, a few 1e-16)
        assert np.array_equal(f != 0, fw != 0) and np.max(np.abs(f - fw)) <= (2e-15 if tables else 1e-15)
        for s in range(0, 150, 11):
            bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
            want = ref.system_light_curve([bd], ub[s], t)[:, 0]
            assert np.max(np.abs(f[s] - want)) <= 1e-12
    # unsorted time stamps: the window-test kernel writes every row
    perm = np.random.default_rng(0).permutation(t.size)
    f, st = _flux_abi(sys_, ub, t[perm])
    fw, _ = _without_list(lambda: _flux_abi(sys_, ub, t[perm]))
    assert st[1] & 1 and np.array_equal(f, fw)
    # the facade call a reference user writes
    from jaxoplanet_b200.light_curves import limb_dark_light_curve

    lc = limb_dark_light_curve(sys_, ub)(t)
    assert lc.shape == (150, 25_000, 1)
    f, _ = _flux_abi(sys_, ub, t)
    assert np.array_equal(lc[..., 0], f)


# ----------------------------------------------------------------------------- K2 sorted by class
def test_limb_dark_class_sorted_kernel_matches_plain_kernel_and_oracle():
    """core.limb_dark.light_curve on >= 8192 points (double, quadratic, order 10) sorts each tile's points
    into "planet fully inside the disk" and "everything else" and runs one flux code per class
    (k_limb_dark_cls); the plain kernel (JXP_NO_LD_CLASSES) and the oracle are the references.  Edge
    values of the reference's own test (tests/core/limb_dark_test.py:27-31) are mixed in."""
    from jaxoplanet_b200.core.limb_dark import light_curve

    rng = np.random.default_rng(17)
    n = 60_001
    r = rng.uniform(0.01, 1.3, n)
    b = rng.uniform(0, 1, n) * (1.2 + r)
    r[:5] = [0.01, 0.1, 0.5, 1.1, 2.0]
    for k, rk in enumerate([0.01, 0.1, 0.5, 1.1, 2.0]):  # contact points and the centre
        sl = slice(10 + 8 * k, 18 + 8 * k)
        r[sl] = rk
        b[sl] = [0.0, abs(1 - rk), 1 + rk, rk, abs(1 - rk) - 1e-14, abs(1 - rk) + 1e-14, 1 + rk - 1e-14, 1e-300]
    b[100] = np.nan
    r[101] = np.nan
    b[102] = -0.3  # the reference takes |b|
    for u in ([0.3, 0.2], [0.4], []):
        got = light_curve(np.array(u), b, r)
        from jaxoplanet_b200 import _lib as _l

        with _l.dev_options(_l.JXP_DEV_NO_LD_CLASSES):
            plain = light_curve(np.array(u), b, r)
        with np.errstate(all="ignore"):
            want = ref.ld_light_curve(np.array(u), b[:20_000], r[:20_000])
        ok = np.isfinite(want)
        # a NaN b or r is a NaN flux in the reference (it goes through atan2 and the kite area), for every
        # law length; both kernels must agree with that
        assert not ok[100] and not ok[101] and ok[102]
        assert np.array_equal(np.isfinite(got[:20_000]), ok) and np.array_equal(np.isfinite(plain[:20_000]), ok)
        fin = np.isfinite(plain)
        assert np.array_equal(np.isfinite(got), fin)
        assert np.max(np.abs(got[fin] - plain[fin])) <= 2e-15
        assert np.max(np.abs(got[:20_000][ok] - want[ok])) <= 1e-12


def test_randomised_list_kernels_against_window_kernels():
    """tools/tl_fuzz.py: 80 small random problems (irregular sorted time stamps with duplicates, circular and
    eccentric bodies, grazing / non-transiting / huge planets, periods from 0.03 to 1000 d): log-likelihood,
    flux and gradient step of the list kernels against the window-test kernels -- same support, same
    finiteness, differences at rounding level."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "tl_fuzz.py"), "80", "3"], capture_output=True,
                       text=True, cwd=root)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_gradient_and_flux_list_paths_no_out_of_bounds_writes():
    # sentinels around the workspace and every output; ragged sizes
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    dev = A.device()
    guard = 4096
    with _lib.dev_options(_lib.JXP_DEV_WALK_ANY_SIZE):
        for n_sets, n_t in ((1, 1), (5, 333), (33, 4099), (70, 9001)):
            theta, t = W.c5(n_sets, n_t)
            th = torch.as_tensor(theta, device=dev).contiguous()
            td = torch.as_tensor(t, device=dev)
            yd = torch.ones(n_t, dtype=torch.float64, device=dev)
            sd = torch.full((1,), 5e-4, dtype=torch.float64, device=dev)
            star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
            nbytes = int(_lib.load().jxp_transit_workspace_bytes(n_sets, n_t))
            buf = torch.full((nbytes + 2 * guard,), 0xA5, dtype=torch.uint8, device=dev)
            ws = buf[guard:guard + nbytes]
            ll = torch.full((n_sets + 16,), -7.0, dtype=torch.float64, device=dev)
            dll = torch.full((n_sets * 8 + 16,), -7.0, dtype=torch.float64, device=dev)
            _lib.call("jxp_transit_partials_f64", th.data_ptr(), n_sets, star.data_ptr(), 0, td.data_ptr(), n_t, 0.0, 0,
                      0, 10, 0, None, None, yd.data_ptr(), sd.data_ptr(), 0, ll[8:].data_ptr(), dll[8:].data_ptr(),
                      ws.data_ptr(), nbytes, A.stream_ptr())
            torch.cuda.synchronize()
            assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + nbytes:] == 0xA5).all())
            assert bool((ll[:8] == -7.0).all()) and bool((ll[8 + n_sets:] == -7.0).all())
            assert bool((dll[:8] == -7.0).all()) and bool((dll[8 + 8 * n_sets:] == -7.0).all())
            assert bool(torch.isfinite(ll[8:8 + n_sets]).all()) and bool(torch.isfinite(dll[8:8 + 8 * n_sets]).all())
            # flux list
            p, ub, t3, _, _ = W.c3(n_sets, n_t)
            bodies, _, _ = pack_bodies(_system(p))
            bodies = bodies.to(dev).contiguous()
            u = torch.as_tensor(ub, device=dev).contiguous()
            t3d = torch.as_tensor(t3, device=dev)
            nbytes = int(_lib.load().jxp_system_workspace_bytes(n_sets, 1, n_t))
            buf = torch.full((nbytes + 2 * guard,), 0xA5, dtype=torch.uint8, device=dev)
            ws = buf[guard:guard + nbytes]
            out = torch.full((n_sets * n_t + 16,), -7.0, dtype=torch.float64, device=dev)
            _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, 1, u.data_ptr(), 2, 2, t3d.data_ptr(), n_t,
                      0.0, 0, 0, 10, 0, None, out[8:].data_ptr(), None, None, 0, None, ws.data_ptr(), nbytes,
                      A.stream_ptr())
            st = (ctypes.c_int64 * 4)()
            _lib.call("jxp_system_list_stats", ws.data_ptr(), n_sets, 1, n_t, ctypes.addressof(st), A.stream_ptr())
            assert st[1] == 0
            assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + nbytes:] == 0xA5).all())
            assert bool((out[:8] == -7.0).all()) and bool((out[8 + n_sets * n_t:] == -7.0).all())
            assert bool(torch.isfinite(out[8:8 + n_sets * n_t]).all()) and bool((out[8:8 + n_sets * n_t] <= 0).all())


def _system_flux_abi(system, u, t, expo=None, per_body=False):
    """Flux of a (multi-body) batch through the C ABI with a workspace we keep -> (flux, list items, flags)."""
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    dev = A.device()
    bodies, n_sets, _ = pack_bodies(system)
    n_bodies = bodies.shape[1]
    bodies = bodies.to(dev).contiguous()
    ud = torch.as_tensor(np.asarray(u, dtype=np.float64), device=dev).contiguous()
    u_stride = ud.shape[1] if ud.ndim == 2 else 0
    td = torch.as_tensor(np.asarray(t, dtype=np.float64), device=dev).contiguous()
    n_t = td.numel()
    nbytes = _lib.load().jxp_system_workspace_bytes(n_sets, n_bodies, n_t)
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    shape = (n_sets, n_t, n_bodies) if per_body else (n_sets, n_t)
    out = torch.full(shape, float("nan"), dtype=torch.float64, device=dev)
    ex = expo or dict(exposure_time=0.0, order=0, num_samples=0)
    _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, n_bodies, ud.data_ptr(), ud.shape[-1], u_stride,
              td.data_ptr(), n_t, ex["exposure_time"], ex["order"], ex["num_samples"], 10, 0,
              out.data_ptr() if per_body else None, None if per_body else out.data_ptr(), None, None, 0, None,
              ws.data_ptr(), ws.numel(), A.stream_ptr())
    stats = (ctypes.c_int64 * 4)()
    _lib.call("jxp_system_list_stats", ws.data_ptr(), n_sets, n_bodies, n_t, ctypes.addressof(stats), A.stream_ptr())
    return out.cpu().numpy(), int(stats[0]), int(stats[1])


@pytest.mark.parametrize("order,num_samples", [(0, 15), (1, 7), (2, 5), (0, 0)])
def test_walk_with_several_bodies_and_exposure_integration_matches_oracle_and_window_kernels(order, num_samples):
    """Config-4-shaped calls on the transit walk (one plan + evaluation per body, a lane walks the sub-samples
    of its sample): summed and per-body flux against the oracle and against the window-test kernels, all three
    stencils of transforms.py:67-89, eccentric + circular bodies, one body that never transits."""
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W

    rng = np.random.default_rng(17 + order)
    n_sets, n_t = 24, 6000
    P = np.stack([rng.uniform(1, 3, n_sets), rng.uniform(4, 9, n_sets), rng.uniform(12, 30, n_sets)], 1)
    p = dict(period=P, time_transit=rng.uniform(0, 1, (n_sets, 3)) * P, eccentricity=rng.uniform(0, 0.4, (n_sets, 3)),
             omega_peri=rng.uniform(-np.pi, np.pi, (n_sets, 3)), impact_param=rng.uniform(0, 1.15, (n_sets, 3)),
             radius=rng.uniform(0.02, 0.2, (n_sets, 3)))
    p["impact_param"][:, 2] = np.where(np.arange(n_sets) % 5 == 0, 3.0, p["impact_param"][:, 2])  # never transits
    u = W.kipping_u(rng.uniform(0, 1, (n_sets, 2)))
    t = np.sort(np.arange(n_t) * W.THIRTY_MIN + rng.uniform(-2e-3, 2e-3, n_t))
    expo = dict(exposure_time=W.THIRTY_MIN, order=order, num_samples=num_samples) if num_samples else None
    sys_ = W.system_from(p)
    with _lib.dev_options(_lib.JXP_DEV_WALK_ANY_SIZE):
        summed, items, flags = _system_flux_abi(sys_, u, t, expo)
        per_body, items_b, flags_b = _system_flux_abi(sys_, u, t, expo, per_body=True)
    assert flags == 0 and items > 0 and flags_b == 0 and items_b == items  # the walk produced both
    summed_w, items_w, _ = _without_list(lambda: _system_flux_abi(sys_, u, t, expo))
    per_body_w, _, _ = _without_list(lambda: _system_flux_abi(sys_, u, t, expo, per_body=True))
    assert items_w == 0
    assert np.max(np.abs(summed - summed_w)) <= 2e-13 and np.max(np.abs(per_body - per_body_w)) <= 2e-13
    assert np.array_equal(summed == 0, summed_w == 0)  # exact zeros out of transit, on both paths
    assert np.max(np.abs(per_body.sum(-1) - summed)) <= 1e-15
    for s in range(0, n_sets, 5):
        bodies = [ref.orbital_body(**{k: v[s, b] for k, v in p.items()}) for b in range(3)]
        f = lambda tt: ref.system_light_curve(bodies, u[s], tt)  # noqa: E731
        want = ref.integrate(f, expo["exposure_time"], order, num_samples)(t) if expo else f(t)
        assert np.max(np.abs(per_body[s] - want) / np.maximum(1.0, np.abs(1.0 + want))) <= 1e-12
        assert np.all(per_body[s][:, 2] == 0) if s % 5 == 0 else True
    assert np.mean(summed != 0) > 0.03
