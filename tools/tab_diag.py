"""Scratch: per-set flux difference (tables vs direct) over all config-3 sets, worst sets with their parameters."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxoplanet_b200 import _lib, _array as A, workloads as W
from jaxoplanet_b200.light_curves.limb_dark import pack_bodies
from oracle import ref_numpy as ref
dev = A.device(); lib = _lib.load()
n_t = 100_000
p, u, t, noise, sigma = W.c3(4096, n_t)
bodies_all, ns_all, _ = pack_bodies(W.system_from(p))
td = torch.as_tensor(t).to(dev)
B = 256
worst = []
for s0 in range(0, 4096, B):
    bodies = bodies_all[s0:s0 + B].to(dev).contiguous(); ud = torch.as_tensor(u[s0:s0 + B]).to(dev).contiguous()
    wsb = lib.jxp_system_workspace_bytes(B, 1, n_t); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    fx = {}
    for name, opt in (("tab", 0), ("direct", _lib.JXP_DEV_NO_TABLES)):
        lib.jxp_set_dev_options(opt | _lib.JXP_DEV_WALK_ANY_SIZE)
        f = torch.empty(B, n_t, dtype=torch.float64, device=dev)
        _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), B, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_t,
                  0.0, 0, 0, 10, 0, None, f.data_ptr(), None, None, 0, None, ws.data_ptr(), wsb, A.stream_ptr())
        torch.cuda.synchronize(); fx[name] = f
    d = (fx["tab"] - fx["direct"]).abs()
    m, am = d.max(dim=1)
    for i in torch.argsort(m, descending=True)[:3].tolist():
        worst.append((float(m[i]), s0 + i, int(am[i]), float(fx["direct"][i, am[i]])))
lib.jxp_set_dev_options(0)
worst.sort(reverse=True)
for w in worst[:12]:
    i = w[1]
    kw = {k: float(v[i]) for k, v in p.items()}
    bd = ref.orbital_body(**{k: v[i] for k, v in p.items()})
    x, y, z = ref.relative_position(bd, t[w[2]:w[2] + 1]); b = float(np.sqrt(x**2 + y**2)[0])
    print("diff %.2e set %d idx %d F %.3e b %.6f 1-r %.6f 1+r %.6f" % (w[0], i, w[2], w[3], b, 1 - kw['radius'], 1 + kw['radius']), {k: round(v, 4) for k, v in kw.items()}, flush=True)
