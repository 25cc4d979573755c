"""Scratch timing of the config-5 gradient step (log-likelihood + 8-gradient, no rows): list path vs window-test kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import _lib, _array as A, workloads as W
dev = A.device(); lib = _lib.load()
ns, nt = int(os.environ.get("NS", 16384)), int(os.environ.get("NT", 50_000))
theta, t5 = W.c5(ns, nt)
thd = torch.as_tensor(theta).to(dev).contiguous(); td = torch.as_tensor(t5).to(dev)
star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
y = torch.as_tensor(1 + 5e-4 * np.random.default_rng(42).normal(size=nt)).to(dev); sd = torch.full((1,), 5e-4, dtype=torch.float64, device=dev)
wsb = lib.jxp_transit_workspace_bytes(ns, nt); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
ll = torch.empty(ns, dtype=torch.float64, device=dev); dll = torch.empty((ns, 8), dtype=torch.float64, device=dev)
call = lambda: _lib.call("jxp_transit_partials_f64", thd.data_ptr(), ns, star.data_ptr(), 0, td.data_ptr(), nt, 0.0, 0, 0, 10, 0,
                         None, None, y.data_ptr(), sd.data_ptr(), 0, ll.data_ptr(), dll.data_ptr(), ws.data_ptr(), wsb, A.stream_ptr())
res = {}
for label in ("list", "window"):
    lib.jxp_set_dev_options(_lib.JXP_DEV_NO_WALK if label == "window" else 0)
    for _ in range(2): call()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    import ctypes
    ls = (ctypes.c_int64 * 4)(); _lib.call("jxp_transit_list_stats", ws.data_ptr(), ns, nt, ctypes.addressof(ls), A.stream_ptr())
    print(label, "list items", int(ls[0]), "flags", int(ls[1]), "inside", int(ls[2]))
    res[label] = (best, ll.clone(), dll.clone())
    print(label, "ms", best, "points/s", ns * nt / best * 1e3, "workspace GB", wsb / 1e9)
l, w = res["list"], res["window"]
print("max rel diff loglike", float(((l[1] - w[1]).abs() / w[1].abs()).max()),
      "grad (scaled by column max)", float(((l[2] - w[2]).abs() / w[2].abs().amax(0, keepdim=True)).max()))
