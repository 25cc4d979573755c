"""Scratch timing of the config-3 log-likelihood: transit walk vs window-test kernels, at NS sets
(env NS, default 4096 and 512), inputs resident, L2 flushed between steps.  Prints one JSON line per NS."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxoplanet_b200 import _lib, _array as A, workloads as W
from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

dev = A.device()
lib = _lib.load()
n_t = int(os.environ.get("NT", 100_000))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n_sets in [int(x) for x in os.environ.get("NS", "4096,512").split(",")]:
    p, u, t, noise, sigma = W.c3(n_sets, n_t)
    bodies, ns, _ = pack_bodies(W.system_from(p))
    bodies = bodies.to(dev).contiguous(); ud = torch.as_tensor(u).to(dev).contiguous(); td = torch.as_tensor(t).to(dev)
    yd = torch.as_tensor(1 + noise).to(dev); sd = torch.full((1,), sigma, dtype=torch.float64, device=dev)
    wsb = lib.jxp_system_workspace_bytes(ns, 1, n_t); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ll = torch.empty(ns, dtype=torch.float64, device=dev)
    call = lambda: _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), ns, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_t,
                             0.0, 0, 0, 10, 0, None, None, yd.data_ptr(), sd.data_ptr(), 0, ll.data_ptr(), ws.data_ptr(), wsb,
                             A.stream_ptr())
    res = {"n_sets": n_sets}
    lls = {}
    for label, opt in (("walk", 0), ("walk_plan", _lib.JXP_DEV_PROFILE_PLAN), ("window", _lib.JXP_DEV_NO_WALK)):
        lib.jxp_set_dev_options(opt)
        for _ in range(3): call()
        torch.cuda.synchronize()
        tot = ktot = 0.0; n = 10
        for _ in range(n):
            flush.fill_(1)
            a, b, ka, kb = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            for e_ in (a, b, ka, kb): e_.record()
            lib.jxp_profile_events(ctypes.c_void_p(ka.cuda_event), ctypes.c_void_p(kb.cuda_event))
            a.record(); call(); b.record()
            lib.jxp_profile_events(None, None)
            torch.cuda.synchronize()
            tot += a.elapsed_time(b); ktot += ka.elapsed_time(kb)
        # back-to-back (no flush, no per-step sync): what a sampler loop sees
        torch.cuda.synchronize(); a.record()
        for _ in range(20): call()
        b.record(); torch.cuda.synchronize()
        st = (ctypes.c_int64 * 2)(); _lib.call("jxp_system_stats", ws.data_ptr(), ns, 1, n_t, ctypes.addressof(st), A.stream_ptr())
        ls = (ctypes.c_int64 * 4)(); _lib.call("jxp_system_list_stats", ws.data_ptr(), ns, 1, n_t, ctypes.addressof(ls), A.stream_ptr())
        res[label] = dict(step_ms=round(tot / n, 4), kernel_ms=round(ktot / n, 4), b2b_ms=round(a.elapsed_time(b) / 20, 4),
                          kepler=int(st[0]), ld=int(st[1]), items=int(ls[0]), flags=int(ls[1]), inside=int(ls[2]), fast=int(ls[3]))
        lls[label] = ll.clone()
    lib.jxp_set_dev_options(0)
    d = (lls["walk"] - lls["window"]).abs() / lls["window"].abs()
    res["max_rel_diff_vs_window"] = float(d.max())
    res["nan"] = int(torch.isnan(lls["walk"]).sum())
    print(json.dumps(res))
