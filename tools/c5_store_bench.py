"""Config 5 with rows stored (flux + 8 partials), f64 and f32: best of 4, HBM store rate."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxoplanet_b200 import _lib, _array as A, workloads as W
dev = A.device(); lib = _lib.load()
theta, t5 = W.c5()
thd = torch.as_tensor(theta).to(dev); t5d = torch.as_tensor(t5).to(dev)
ns5, nt5 = thd.shape[0], t5d.numel()
star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
wsb5 = int(lib.jxp_transit_workspace_bytes(ns5, nt5)); ws5 = torch.empty(wsb5, dtype=torch.uint8, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for dt, nm in ((torch.float64, "f64"), (torch.float32, "f32")):
    fl = torch.empty((ns5, nt5), dtype=dt, device=dev); pt = torch.empty((ns5, 8, nt5), dtype=dt, device=dev)
    for mask, what in ((0, "bulk-copy zero fill"), (_lib.JXP_DEV_TR_TPW(2), "bulk-copy zero fill, two tiles per warp"),
                       (_lib.JXP_DEV_BULK_FILL_SWAP, "bulk-copy zero fill, other mode"),
                       (_lib.JXP_DEV_NO_BULK_FILL, "warp-store zero fill")):
        best = 1e30
        with _lib.dev_options(mask):
            for i in range(5):
                ev0.record()
                _lib.call("jxp_transit_partials_" + nm, thd.data_ptr(), ns5, star.data_ptr(), 0, t5d.data_ptr(), nt5, 0.0, 0, 0,
                          10, 0, fl.data_ptr(), pt.data_ptr(), 0, 0, 0, 0, 0, ws5.data_ptr(), wsb5, A.stream_ptr())
                ev1.record(); torch.cuda.synchronize()
                if i: best = min(best, ev0.elapsed_time(ev1))
        gbs = ns5 * nt5 * 9 * fl.element_size() / (best * 1e-3) / 1e9
        print(nm, what, "ms", round(best, 3), "GB/s", round(gbs, 1), "frac of 6547.8:", round(gbs / 6547.8, 3), flush=True)
    del fl, pt; torch.cuda.empty_cache()
