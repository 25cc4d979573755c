"""The entry points are stream-capturable (no synchronisation, no host callback inside a call; the
side stream of the table launch forks from and joins the caller's stream by events, which a capture follows): a
call captured into a CUDA graph and replayed writes the same bits as the stream-ordered call."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _capture(fn):
    from jaxoplanet_b200.graphs import capture

    g = capture(fn)
    assert g is not None
    return g


def test_system_light_curve_and_partials_replay_from_a_cuda_graph():
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import pack_elements

    lib = _lib.load()
    n_sets, n_times = 300, 30_000
    p, u, t, noise, sigma = W.c3(n_sets, n_times, seed=4)
    el = torch.as_tensor(np.ascontiguousarray(pack_elements(n_sets, **p))).cuda()
    bodies = torch.empty((n_sets, 1, _lib.JXP_BODY_NFIELDS), dtype=torch.float64, device="cuda")
    _lib.call("jxp_orbital_elements_f64", el.data_ptr(), n_sets, bodies.data_ptr(), A.stream_ptr())
    ud, td = torch.as_tensor(np.ascontiguousarray(u)).cuda(), torch.as_tensor(t).cuda()
    yd = torch.as_tensor(1.0 + noise).cuda()
    sg = torch.full((1,), sigma, dtype=torch.float64, device="cuda")
    wsb = int(lib.jxp_system_workspace_bytes(n_sets, 1, n_times))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    ll = torch.empty(n_sets, dtype=torch.float64, device="cuda")
    flux = torch.empty((n_sets, n_times), dtype=torch.float64, device="cuda")

    def loglike():  # transit walk + function tables (side stream) + fallback kernels
        _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_times,
                  0.0, 0, 0, 10, 0, 0, 0, yd.data_ptr(), sg.data_ptr(), 0, ll.data_ptr(), ws.data_ptr(), wsb, A.stream_ptr())

    def fluxes():  # memset + walk with stored rows
        _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_times,
                  0.0, 0, 0, 10, 0, 0, flux.data_ptr(), 0, 0, 0, 0, ws.data_ptr(), wsb, A.stream_ptr())

    theta = torch.as_tensor(W.theta_from(p, u)).cuda()
    star = torch.tensor([1.0, 1.0], dtype=torch.float64, device="cuda")
    wsb5 = int(lib.jxp_transit_workspace_bytes(n_sets, n_times))
    ws5 = torch.empty(wsb5, dtype=torch.uint8, device="cuda")
    fl5 = torch.empty((n_sets, n_times), dtype=torch.float64, device="cuda")
    pt5 = torch.empty((n_sets, 8, n_times), dtype=torch.float64, device="cuda")

    def partials():  # K5 with the bulk-copy zero fill
        _lib.call("jxp_transit_partials_f64", theta.data_ptr(), n_sets, star.data_ptr(), 0, td.data_ptr(), n_times, 0.0, 0, 0,
                  10, 0, fl5.data_ptr(), pt5.data_ptr(), 0, 0, 0, 0, 0, ws5.data_ptr(), wsb5, A.stream_ptr())

    for fn, outs in ((loglike, (ll,)), (fluxes, (flux,)), (partials, (fl5, pt5))):
        fn()
        torch.cuda.synchronize()
        want = [o.clone() for o in outs]
        g = _capture(fn)
        for _ in range(3):
            for o in outs:
                o.fill_(7.0)
            g.replay()
            torch.cuda.synchronize()
            for o, w in zip(outs, want):
                assert torch.equal(o, w)
    assert float(flux.min()) < 0.0 and bool(torch.isfinite(ll).all())
