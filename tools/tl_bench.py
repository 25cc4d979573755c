"""Scratch timing of the config-3 log-likelihood: transit-list kernels vs window-test kernels."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import _lib, _array as A, workloads as W
from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

dev = A.device()
lib = _lib.load()
n_sets, n_t = int(os.environ.get("NS", 4096)), int(os.environ.get("NT", 100_000))
p, u, t, noise, sigma = W.c3(n_sets, n_t)
bodies, ns, _ = pack_bodies(W.system_from(p))
bodies = bodies.to(dev).contiguous(); ud = torch.as_tensor(u).to(dev).contiguous(); td = torch.as_tensor(t).to(dev)
yd = torch.as_tensor(1 + noise).to(dev); sd = torch.full((1,), sigma, dtype=torch.float64, device=dev)
wsb = lib.jxp_system_workspace_bytes(ns, 1, n_t); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
ll = torch.empty(ns, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
call = lambda: _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), ns, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_t,
                         0.0, 0, 0, 10, 0, None, None, yd.data_ptr(), sd.data_ptr(), 0, ll.data_ptr(), ws.data_ptr(), wsb,
                         A.stream_ptr())
res = {}
for label, env in (("list", None), ("list_prep", "prep"), ("list_build", "build"), ("list_reduce", "reduce"), ("window", "1")):
    os.environ.pop("JXP_NO_TL", None); os.environ.pop("JXP_PROFILE_KERNEL", None)
    if env == "1": os.environ["JXP_NO_TL"] = env
    elif env: os.environ["JXP_PROFILE_KERNEL"] = env
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
    st = (ctypes.c_int64 * 2)(); _lib.call("jxp_system_stats", ws.data_ptr(), ns, 1, n_t, ctypes.addressof(st), A.stream_ptr())
    ls = (ctypes.c_int64 * 4)(); _lib.call("jxp_system_list_stats", ws.data_ptr(), ns, 1, n_t, ctypes.addressof(ls), A.stream_ptr())
    res[label] = dict(step_ms=tot / n, kernel_ms=ktot / n, kepler=int(st[0]), ld=int(st[1]), items=int(ls[0]), flags=int(ls[1]), inside_items=int(ls[2]), fast_chunks=int(ls[3]),
                      ll0=float(ll[0]), tflops=(st[0] * 302 + st[1] * 746) / (ktot / n * 1e-3) / 1e12)
    res[label + "_ll"] = ll.clone()
d = (res["list_ll"] - res["window_ll"]).abs() / res["window_ll"].abs()
for k in [k for k in res if k.endswith("_ll")]: res.pop(k)
for k in ("list_prep", "list_build", "list_reduce"): res[k] = {"kernel_ms": res[k]["kernel_ms"], "step_ms": res[k]["step_ms"]}
res["max_rel_diff"] = float(d.max())
print(json.dumps(res, indent=1))
