"""Does the float K5 time depend on what was allocated / run before it?  (bench.py 6.7 ms, c5_store_bench.py 5.85 ms)"""
import sys, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import _lib, _array as A, workloads as W
dev = A.device(); lib = _lib.load()
theta, t5 = W.c5()
thd = torch.as_tensor(theta).to(dev); t5d = torch.as_tensor(t5).to(dev)
ns5, nt5 = thd.shape[0], t5d.numel()
star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
wsb5 = int(lib.jxp_transit_workspace_bytes(ns5, nt5)); ws5 = torch.empty(wsb5, dtype=torch.uint8, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run(dt, nm, reps, label):
    fl = torch.empty((ns5, nt5), dtype=dt, device=dev); pt = torch.empty((ns5, 8, nt5), dtype=dt, device=dev)
    ts = []
    for i in range(reps):
        ev0.record()
        _lib.call("jxp_transit_partials_" + nm, thd.data_ptr(), ns5, star.data_ptr(), 0, t5d.data_ptr(), nt5, 0.0, 0, 0, 10, 0,
                  fl.data_ptr(), pt.data_ptr(), 0, 0, 0, 0, 0, ws5.data_ptr(), wsb5, A.stream_ptr())
        ev1.record(); torch.cuda.synchronize(); ts.append(round(ev0.elapsed_time(ev1), 3))
    print(label, nm, ts, "flux ptr % 2MB", fl.data_ptr() % (1 << 21), "partials ptr % 2MB", pt.data_ptr() % (1 << 21), flush=True)
    del fl, pt; torch.cuda.empty_cache()
run(torch.float32, "f32", 5, "first thing in the process")
run(torch.float64, "f64", 4, "then")
run(torch.float32, "f32", 5, "after 4 runs in double")
run(torch.float64, "f64", 20, "then")
run(torch.float32, "f32", 5, "after 20 runs in double")
