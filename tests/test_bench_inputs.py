"""Oracle parity ON THE BENCH INPUTS: the CUDA path at BASELINE.json's full sizes, on exactly the seeded
arrays bench.py times (``workloads.c3() / c4() / c5()``), compared with the CPU oracle on a random sample
of >= 16 parameter sets over ALL their time samples -- at the north-star tolerances (fp64 flux 1e-12,
fp32 mode 1 ppm, gradients 1e-9), through the walk path and through the window-test fallback.

(tests/test_full_size_properties.py checks the same sizes through kernel-vs-kernel identities; the oracle
comparisons of tests/test_gpu_parity.py run at reduced sizes.)
"""

import contextlib

import numpy as np
import pytest

from oracle import ref_numpy as ref

pytestmark = pytest.mark.gpu

N_SAMPLE = 16


@contextlib.contextmanager
def dev_options(mask):
    from jaxoplanet_b200 import _lib

    lib = _lib.load()
    keep = lib.jxp_get_dev_options()
    lib.jxp_set_dev_options(mask)
    try:
        yield
    finally:
        lib.jxp_set_dev_options(keep)


def _sample(n_sets, seed):
    # the generating set and the first / last sets always, the rest drawn
    rng = np.random.default_rng(seed)
    pick = {0, n_sets - 1} | set(int(i) for i in rng.choice(n_sets, N_SAMPLE, replace=False))
    return sorted(pick)


def _flux_close(got, want, tol=1e-12):
    err = np.abs(got - want) / np.maximum(1.0, np.abs(1.0 + want))
    assert np.max(err) <= tol, float(np.max(err))


def test_c3_bench_inputs_flux_and_loglike_match_oracle_walk_and_fallback_f64_f32():
    """Config 3 exactly as bench.py builds it (seed 1, y = 1 + model of set 0 + noise, sigma = 5e-4)."""
    import torch

    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import FusedSpec, _run_fused, log_likelihood

    p, u, t, noise, sigma = W.c3()
    assert p["period"].shape == (4096,) and t.shape == (100_000,)
    system = W.system_from(p)
    td = torch.as_tensor(t).cuda()
    spec = FusedSpec(system, u)
    flux = _run_fused(spec, td, want="flux_sum")  # transit walk, flux mode (4096 x 1e5)
    y = (1.0 + flux[0] + torch.as_tensor(noise).cuda()).contiguous()
    ll_walk = log_likelihood(system, u)(td, y, sigma)
    with dev_options(_lib.JXP_DEV_NO_WALK):
        ll_win = log_likelihood(system, u)(td, y, sigma)
        flux_win = _run_fused(spec, td, want="flux_sum")
    flux32 = _run_fused(spec, td, want="flux_sum", dtype=torch.float32)
    assert flux32.dtype == torch.float32
    y_np = y.cpu().numpy()
    pick = _sample(4096, 11)
    n_in = 0
    for s in pick:
        bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
        want = ref.system_light_curve([bd], u[s], t)[:, 0]
        n_in += int(np.sum(want != 0))
        _flux_close(flux[s].cpu().numpy(), want)
        _flux_close(flux_win[s].cpu().numpy(), want)
        assert np.max(np.abs(flux32[s].cpu().numpy().astype(np.float64) - want)) <= 1e-6  # fp32 mode: 1 ppm
        w = ref.gaussian_loglike(1.0 + want, y_np, np.full_like(t, sigma))
        assert abs(float(ll_walk[s]) - w) <= 1e-12 * abs(w), (s, float(ll_walk[s]), w)
        assert abs(float(ll_win[s]) - w) <= 1e-12 * abs(w), (s, float(ll_win[s]), w)
    assert n_in > 20_000  # the sample does contain transits
    # every set: the two device paths agree to the same bar (scale: the constant term n |log(sigma sqrt(2 pi))|
    # of the sum -- a log-likelihood that happens to pass through 0 has no meaningful relative error)
    scale = t.size * abs(np.log(sigma) + 0.5 * np.log(2 * np.pi))
    assert float(torch.max(torch.abs(ll_walk - ll_win) / torch.clamp(torch.abs(ll_win), min=scale))) <= 1e-12


def test_c4_bench_inputs_match_oracle_per_body_and_summed():
    """Config 4 as bench.py runs it (3 planets, 30-min cadence, 15x order-0 exposure integration,
    1024 sets x 2e5 times, summed flux): 16 sampled sets against the C restatement of the oracle (all 2e5
    times; oracle/ref_c.c is itself pinned to oracle/ref_numpy.py by tests/test_oracle_c.py), and two of them
    per body against the NumPy oracle."""
    import torch

    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import FusedSpec, _run_fused
    from oracle import ref_c

    p, u, t, expo = W.c4()
    system = W.system_from(p)
    td = torch.as_tensor(t).cuda()
    spec = FusedSpec(system, u, exposure_time=expo["exposure_time"], exp_order=expo["order"],
                     num_samples=expo["num_samples"])
    summed = _run_fused(spec, td, want="flux_sum")
    assert summed.shape == (1024, 200_000)
    pick = _sample(1024, 12)
    dt, w = ref.integrate_stencil(expo["exposure_time"], expo["order"], expo["num_samples"])
    ref_c.use_all_cores()
    rec = np.stack([[_record(ref.orbital_body(**{k: v[s, b] for k, v in p.items()})) for b in range(3)]
                    for s in pick])
    want, _ = ref_c.system(rec, u[pick], t, dt=dt, w=w)
    got = summed[pick].cpu().numpy()
    _flux_close(got, want)
    assert np.mean(want != 0) > 0.05
    del summed
    # per body, NumPy oracle, two sets
    sub = {k: v[pick[:2]] for k, v in p.items()}
    per_body = _run_fused(FusedSpec(W.system_from(sub), u[pick[:2]], exposure_time=expo["exposure_time"],
                                    exp_order=expo["order"], num_samples=expo["num_samples"]), td).cpu().numpy()
    for i, s in enumerate(pick[:2]):
        bodies = [ref.orbital_body(**{k: v[s, b] for k, v in p.items()}) for b in range(3)]
        want_b = ref.integrate(lambda tt: ref.system_light_curve(bodies, u[s], tt), expo["exposure_time"],
                               expo["order"], expo["num_samples"])(t)
        _flux_close(per_body[i], want_b)


def _record(bd):
    """Oracle body dict -> the 12-double record of include/jaxoplanet_b200.h (what oracle/ref_c.c reads)."""
    from jaxoplanet_b200 import _lib

    circular = bd["eccentricity"] is None
    r = np.zeros(_lib.JXP_BODY_NFIELDS)
    r[_lib.BODY_PERIOD] = bd["period"]
    r[_lib.BODY_TIME_TRANSIT] = bd["time_transit"]
    r[_lib.BODY_TIME_REF] = bd["time_ref"]
    r[_lib.BODY_ECC] = -1.0 if circular else bd["eccentricity"]
    r[_lib.BODY_COS_OMEGA] = 1.0 if circular else bd["cos_omega_peri"]
    r[_lib.BODY_SIN_OMEGA] = 0.0 if circular else bd["sin_omega_peri"]
    r[_lib.BODY_COS_INCL] = bd["cos_inclination"]
    r[_lib.BODY_SIN_INCL] = bd["sin_inclination"]
    r[_lib.BODY_SEMIMAJOR] = bd["semimajor"]
    r[_lib.BODY_CENTRAL_RADIUS] = bd["central_radius"]
    r[_lib.BODY_RADIUS] = 0.0 if bd["radius"] is None else bd["radius"]
    return r


def test_c5_bench_inputs_partials_and_gradient_step_match_autograd_oracle():
    """Config 5 as bench.py runs it (16384 sets x 5e4 times): the stored flux + 8 partials (f64, and the f32
    mode at 1 ppm of flux), and the gradient step on the transit list (k_tlg_eval: log-likelihood + its
    8-gradient, no rows stored) and on the window-test kernel, against torch-f64 autograd over the oracle
    restatement for 16 sampled sets over all 5e4 times."""
    import torch

    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.partials import transit_partials
    from test_gpu_parity import _partials_oracle  # (tests/ is on sys.path under pytest)

    theta, t = W.c5()
    assert theta.shape == (16384, 8) and t.shape == (50_000,)
    td = torch.as_tensor(t).cuda()
    thd = torch.as_tensor(theta).cuda()
    sigma = 5e-4
    stored = transit_partials(thd, td)
    rng = np.random.default_rng(42)
    y = 1.0 + stored["flux"][3] + torch.as_tensor(sigma * rng.normal(size=t.size)).cuda()
    pick = _sample(16384, 13)
    got_flux = stored["flux"][pick].cpu().numpy()
    got_part = stored["partials"][pick].cpu().numpy()
    del stored
    torch.cuda.empty_cache()
    f32 = transit_partials(thd, td, want_partials=False, dtype=torch.float32)["flux"][pick].cpu().numpy()
    torch.cuda.empty_cache()
    grad_list = transit_partials(thd, td, y=y, sigma=sigma, want_flux=False, want_partials=False)
    with dev_options(_lib.JXP_DEV_NO_WALK):
        grad_win = transit_partials(thd, td, y=y, sigma=sigma, want_flux=False, want_partials=False)
    y_np = y.cpu().numpy()
    scale = np.array([1e-1, 1e-1, 1e-2, 1e-2, 1e-2, 1e-2, 1e-3, 1e-3])[:, None]
    dll_scale = torch.abs(grad_list["dloglike"]).amax(dim=0).cpu().numpy()
    n_checked = 0
    for i, s in enumerate(pick):
        flux0, part0 = _partials_oracle(theta[s], t)
        _flux_close(got_flux[i], flux0)
        assert np.max(np.abs(f32[i].astype(np.float64) - flux0)) <= 1e-6
        bd = ref.orbital_body(**dict(zip(("period", "time_transit", "eccentricity", "omega_peri",
                                          "impact_param", "radius"), theta[s, :6])))
        x, yy, z = ref.relative_position(bd, t)
        b = np.sqrt(x**2 + yy**2)
        r = theta[s, 5]
        ok = (np.abs(b - r) > 1e-6) & (np.abs(np.abs(b - r) - 1) > 1e-6) & (np.abs(b - 1 - r) > 1e-6) & (b > 1e-6)
        err = np.abs(got_part[i] - part0) / np.maximum(np.abs(part0), scale)
        assert np.max(err[:, ok]) <= 1e-9, (s, np.max(err[:, ok], axis=1))
        assert np.all(got_part[i][:, flux0 == 0] == 0)
        n_checked += int(np.sum(flux0 != 0))
        # gradient step: d loglike / d theta_k = sum_t (y - 1 - m) / sigma^2 * d m / d theta_k
        res = (y_np - 1.0 - flux0) / sigma
        want_ll = ref.gaussian_loglike(1.0 + flux0, y_np, np.full_like(t, sigma))
        want_g = (part0 * (res / sigma)[None, :]).sum(1)
        for name, g in (("list", grad_list), ("window", grad_win)):
            assert abs(float(g["loglike"][s]) - want_ll) <= 1e-12 * abs(want_ll), (name, s)
            gerr = np.abs(g["dloglike"][s].cpu().numpy() - want_g) / np.maximum(np.abs(want_g), 1e-6 * dll_scale)
            assert np.max(gerr) <= 1e-9, (name, s, gerr)
    assert n_checked > 10_000
