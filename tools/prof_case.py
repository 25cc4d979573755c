"""Small single-purpose driver for ncu captures: python tools/prof_case.py <case> [n_sets] [n_t] [flags]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import _lib, _array as A
import jaxoplanet_b200 as jxp
from jaxoplanet_b200.orbits.keplerian import Central, System
from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

case = sys.argv[1]
n_sets = int(sys.argv[2]) if len(sys.argv) > 2 else 512
n_t = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = A.device()

def timed(f):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

if case == "kepler":
    n = n_sets * n_t
    M = torch.rand(n, dtype=torch.float64, device=dev) * (2 * np.pi)
    e = torch.rand(n, dtype=torch.float64, device=dev) * 0.99
    s = torch.empty_like(M); c = torch.empty_like(M)
    ms = timed(lambda: _lib.call("jxp_kepler_f64", M.data_ptr(), 1, e.data_ptr(), 1, n, s.data_ptr(), c.data_ptr(), 0, 0, 0, A.stream_ptr()))
    print("kepler", n, "ms", ms, "solves/s", n / ms * 1e3)
elif case == "ld":
    n = n_sets * n_t
    rng = np.random.default_rng(0)
    r = torch.rand(n, dtype=torch.float64, device=dev) * 0.12 + 0.03
    b = torch.rand(n, dtype=torch.float64, device=dev) * (1 + r)
    u = torch.tensor([0.3, 0.2], dtype=torch.float64, device=dev)
    f = torch.empty_like(b)
    ms = timed(lambda: _lib.call("jxp_limb_dark_f64", u.data_ptr(), 2, b.data_ptr(), 1, r.data_ptr(), 1, n, 10, f.data_ptr(), 0, 0, 0, A.stream_ptr()))
    print("ld", n, "ms", ms, "points/s", n / ms * 1e3)
elif case == "c4":
    from jaxoplanet_b200 import workloads as W
    p4, u4, t4, expo = W.c4(n_sets, n_t)
    bodies, ns, _ = pack_bodies(W.system_from(p4))
    bodies = bodies.to(dev).contiguous(); ud = torch.as_tensor(u4).to(dev).contiguous(); td = torch.as_tensor(t4).to(dev)
    wsb = _lib.load().jxp_system_workspace_bytes(ns, 3, n_t); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    fs = torch.empty((ns, n_t), dtype=torch.float64, device=dev)
    ms = timed(lambda: _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), ns, 3, ud.data_ptr(), 2, 2, td.data_ptr(), n_t,
                                 expo["exposure_time"], expo["order"], expo["num_samples"], 10, flags, 0, fs.data_ptr(), 0, 0, 0, 0,
                                 ws.data_ptr(), wsb, A.stream_ptr()))
    print("c4", ns, n_t, "flags", flags, "ms", ms, "points/s", ns * n_t / ms * 1e3, "inner/s", ns * n_t * 45 / ms * 1e3)
elif case == "c5":
    from jaxoplanet_b200 import workloads as W
    theta, t5 = W.c5(n_sets, n_t)
    thd = torch.as_tensor(theta).to(dev).contiguous(); t5d = torch.as_tensor(t5).to(dev)
    star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
    wsb5 = _lib.load().jxp_transit_workspace_bytes(n_sets, n_t); ws5 = torch.empty(wsb5, dtype=torch.uint8, device=dev)
    fl = torch.empty((n_sets, n_t), dtype=torch.float64, device=dev); pt = torch.empty((n_sets, 8, n_t), dtype=torch.float64, device=dev)
    ms = timed(lambda: _lib.call("jxp_transit_partials_f64", thd.data_ptr(), n_sets, star.data_ptr(), 0, t5d.data_ptr(), n_t, 0.0, 0, 0, 10, flags,
                                 fl.data_ptr(), pt.data_ptr(), 0, 0, 0, 0, 0, ws5.data_ptr(), wsb5, A.stream_ptr()))
    print("c5", n_sets, n_t, "flags", flags, "ms", ms, "points/s", n_sets * n_t / ms * 1e3, "GB/s", n_sets * n_t * 72 / ms / 1e6)
else:
    rng = np.random.default_rng(1)
    P = rng.uniform(2, 5, n_sets)
    p = dict(period=P, time_transit=rng.uniform(0, 1, n_sets) * P, eccentricity=rng.uniform(0, .3, n_sets),
             omega_peri=rng.uniform(-np.pi, np.pi, n_sets), impact_param=rng.uniform(0, 1, n_sets), radius=rng.uniform(.03, .15, n_sets))
    q = rng.uniform(0, 1, (n_sets, 2))
    u = np.stack([2 * np.sqrt(q[:, 0]) * q[:, 1], np.sqrt(q[:, 0]) * (1 - 2 * q[:, 1])], 1)
    t = np.arange(n_t) * (2.0 / 1440.0)
    with jxp.batched():
        sys_ = System(Central()).add_body(**p)
    bodies, ns, _ = pack_bodies(sys_)
    bodies = bodies.to(dev).contiguous(); ud = torch.as_tensor(u).to(dev).contiguous(); td = torch.as_tensor(t).to(dev)
    yd = torch.as_tensor(1 + 5e-4 * np.random.default_rng(42).normal(size=t.shape)).to(dev)
    sd = torch.full((1,), 5e-4, dtype=torch.float64, device=dev)
    wsb = _lib.load().jxp_system_workspace_bytes(ns, 1, n_t); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ll = torch.empty(ns, dtype=torch.float64, device=dev)
    fs = torch.empty((ns, n_t), dtype=torch.float64, device=dev) if case == "c3flux" else None
    def f():
        _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), ns, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_t, 0.0, 0, 0, 10, flags,
                  0, A.ptr(fs), A.ptr(yd) if fs is None else 0, A.ptr(sd) if fs is None else 0, 0, A.ptr(ll) if fs is None else 0, ws.data_ptr(), wsb, A.stream_ptr())
    ms = timed(f)
    print(case, ns, n_t, "flags", flags, "ms", ms, "points/s", ns * n_t / ms * 1e3)
