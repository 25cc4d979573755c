"""Scratch: config-3 log-likelihood calls (for ncu captures of the walk kernels)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxoplanet_b200 import _lib, _array as A, workloads as W
from jaxoplanet_b200.light_curves.limb_dark import pack_bodies
dev = A.device(); lib = _lib.load()
n_t = 100_000; ns = int(os.environ.get("NS", 4096))
p, u, t, noise, sigma = W.c3(ns, n_t)
bodies, ns, _ = pack_bodies(W.system_from(p))
bodies = bodies.to(dev).contiguous(); ud = torch.as_tensor(u).to(dev).contiguous(); td = torch.as_tensor(t).to(dev)
yd = torch.as_tensor(1 + noise).to(dev); sd = torch.full((1,), sigma, dtype=torch.float64, device=dev)
wsb = lib.jxp_system_workspace_bytes(ns, 1, n_t); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
ll = torch.empty(ns, dtype=torch.float64, device=dev)
for _ in range(int(os.environ.get("REPS", 3))):
    _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), ns, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_t,
              0.0, 0, 0, 10, 0, None, None, yd.data_ptr(), sd.data_ptr(), 0, ll.data_ptr(), ws.data_ptr(), wsb, A.stream_ptr())
torch.cuda.synchronize()
print("ok", float(ll[0]))
