"""Host-side cost of the pieces of one facade step (no device sync inside the timed pieces)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
import jaxoplanet_b200 as jxp
from jaxoplanet_b200 import _array as A, _lib, workloads as W
from jaxoplanet_b200.orbits.keplerian import Central, Body, OrbitalBody, System
from jaxoplanet_b200.light_curves.limb_dark import log_likelihood, pack_bodies, FusedSpec, _run_fused
p, u, t, noise, sigma = W.c3(); y = 1 + noise
T = {}
def tick(name, t0):
    T.setdefault(name, []).append(time.perf_counter() - t0)
for it in range(40):
    torch.cuda.synchronize()
    with jxp.batched():
        t0 = time.perf_counter(); c = Central(); tick("Central()", t0)
        t0 = time.perf_counter(); b = Body(**p); tick("Body()", t0)
        t0 = time.perf_counter(); ob = OrbitalBody(c, b); tick("OrbitalBody()", t0)
        t0 = time.perf_counter(); s = System(c, bodies=[ob]); tick("System()", t0)
    t0 = time.perf_counter(); f = log_likelihood(s, u); tick("log_likelihood()", t0)
    t0 = time.perf_counter(); bodies, ns, _ = pack_bodies(s); tick("pack_bodies", t0)
    t0 = time.perf_counter(); td = A.registered_to_dev(t); yd = A.registered_to_dev(y); tick("registered t,y", t0)
    t0 = time.perf_counter(); ud, sd = A.stage_f64([u, np.asarray(sigma)]); tick("stage u,sigma", t0)
    t0 = time.perf_counter()
    wsb = _lib.load().jxp_system_workspace_bytes(ns, 1, t.size); ws = torch.empty(wsb, dtype=torch.uint8, device=A.device())
    ll = torch.empty(ns, dtype=torch.float64, device=A.device()); tick("ws+out alloc", t0)
    t0 = time.perf_counter()
    _lib.call("jxp_system_light_curve_f64", A.ptr(bodies), ns, 1, A.ptr(ud), 2, 2, A.ptr(td), t.size, 0.0, 0, 0, 10, 0, 0, 0,
              A.ptr(yd), A.ptr(sd), 0, A.ptr(ll), A.ptr(ws), ws.numel(), A.stream_ptr()); tick("C ABI call (4 launches)", t0)
    t0 = time.perf_counter(); out = ll.cpu().numpy(); tick("D2H + wait for the GPU", t0)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); r = _run_fused(FusedSpec(s, u), t, want="loglike", y=y, sigma=sigma, host_sync=True); tick("_run_fused (whole, no sync)", t0)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); r = f(t, y, sigma); tick("facade call (incl. GPU wait)", t0)
tot = 0
for k, v in T.items():
    m = 1e3 * float(np.median(v[5:])); print(f"{k:32s} {m:8.4f} ms")
print("workspace bytes", wsb)
