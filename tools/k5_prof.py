"""K5 with rows stored on a config-5-shaped batch (n_sets x 50000): python tools/k5_prof.py [n_sets] [f64|f32] [mask]"""
import sys, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import _lib, _array as A, workloads as W
dev = A.device()
n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nm = sys.argv[2] if len(sys.argv) > 2 else "f64"
mask = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dt = torch.float64 if nm == "f64" else torch.float32
theta, t5 = W.c5(n_sets, 50000)
thd = torch.as_tensor(theta).to(dev).contiguous(); t5d = torch.as_tensor(t5).to(dev)
star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
wsb = _lib.load().jxp_transit_workspace_bytes(n_sets, 50000); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
fl = torch.empty((n_sets, 50000), dtype=dt, device=dev); pt = torch.empty((n_sets, 8, 50000), dtype=dt, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with _lib.dev_options(mask):
    for i in range(3):
        ev0.record()
        _lib.call("jxp_transit_partials_" + nm, thd.data_ptr(), n_sets, star.data_ptr(), 0, t5d.data_ptr(), 50000, 0.0, 0, 0, 10, 0,
                  fl.data_ptr(), pt.data_ptr(), 0, 0, 0, 0, 0, ws.data_ptr(), wsb, A.stream_ptr())
        ev1.record(); torch.cuda.synchronize()
        print(nm, n_sets, "sets, mask", mask, ":", round(ev0.elapsed_time(ev1), 3), "ms", flush=True)
