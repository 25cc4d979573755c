"""Config 4 (3 planets, 15x exposure integration, 1024 sets x 2e5 times, summed flux): transit walk vs
window-test kernels, best of 3, and the difference between the two results."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxoplanet_b200 import _lib, _array as A, workloads as W
from jaxoplanet_b200.light_curves.limb_dark import pack_bodies
dev = A.device(); lib = _lib.load()
p, u, t, expo = W.c4()
bodies, n_sets, _ = pack_bodies(W.system_from(p)); bodies = bodies.to(dev).contiguous()
ud = torch.as_tensor(u).to(dev).contiguous(); td = torch.as_tensor(t).to(dev); n_times = td.numel()
wsb = int(lib.jxp_system_workspace_bytes(n_sets, 3, n_times)); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for label, opt, per_body in (("walk", 0, False), ("walk direct", _lib.JXP_DEV_NO_TABLES, False), ("window", _lib.JXP_DEV_NO_WALK, False), ("walk per-body", 0, True)):
    lib.jxp_set_dev_options(opt)
    fs = torch.empty((n_sets, n_times, 3) if per_body else (n_sets, n_times), dtype=torch.float64, device=dev)
    best = 1e30
    for i in range(4):
        ev0.record()
        _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, 3, ud.data_ptr(), 2, 2, td.data_ptr(), n_times,
                  expo["exposure_time"], expo["order"], expo["num_samples"], 10, 0, fs.data_ptr() if per_body else 0,
                  0 if per_body else fs.data_ptr(), 0, 0, 0, 0, ws.data_ptr(), wsb, A.stream_ptr())
        ev1.record(); torch.cuda.synchronize()
        if i: best = min(best, ev0.elapsed_time(ev1))
    st = (ctypes.c_int64 * 2)(); _lib.call("jxp_system_stats", ws.data_ptr(), n_sets, 3, n_times, ctypes.addressof(st), A.stream_ptr())
    ls = (ctypes.c_int64 * 4)(); _lib.call("jxp_system_list_stats", ws.data_ptr(), n_sets, 3, n_times, ctypes.addressof(ls), A.stream_ptr())
    wk = (ctypes.c_int64 * 4)(); _lib.call("jxp_system_walk_stats", ws.data_ptr(), n_sets, 3, n_times, ctypes.addressof(wk), A.stream_ptr())
    print(label, "ms", round(best, 3), "kepler", st[0], "ld", st[1], "items", ls[0], "flags", ls[1], "inside", ls[2], "fast", ls[3], "tab_evals", wk[0], "tab_units", wk[2], "no_tables", wk[3])
    res[label] = fs.sum(-1) if per_body else fs
lib.jxp_set_dev_options(0)
print("max |walk - window|", float((res["walk"] - res["window"]).abs().max()), "max |walk - walk direct|", float((res["walk"] - res["walk direct"]).abs().max()), "max |per-body sum - summed|", float((res["walk per-body"] - res["walk"]).abs().max()),
      "in-transit fraction", float((res["walk"] != 0).double().mean()))
