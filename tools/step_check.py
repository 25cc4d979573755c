"""Config 3 step with and without programmatic dependent launches (device-timed, L2 flushed), 4096 and 512 sets."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, ".")
import bench
from jaxoplanet_b200 import _lib, _array as A
dev = A.device(); lib = _lib.load()
p, u_h, t_h, noise, sigma = bench.c3_inputs(seed=1)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (4096, 2048, 512):
    prob = bench.Problem({k: v[:n] for k, v in p.items()}, u_h[:n], t_h, dev)
    y = torch.as_tensor(1.0 + noise).to(dev); sig = torch.full((1,), sigma, dtype=torch.float64, device=dev)
    ref = None
    for mask, what in ((0, "default"), (0, "again"), (0, "default"), (0, "again")):
        with _lib.dev_options(mask):
            ts = []
            for i in range(25):
                flush.fill_(1)
                ev0.record(); prob.call(y=y, sig=sig); ev1.record(); torch.cuda.synchronize()
                if i >= 5: ts.append(ev0.elapsed_time(ev1))
        ll = prob.ll.clone()
        if ref is None: ref = ll
        print(n, "sets", what, "ms/step mean", round(float(np.mean(ts)), 4), "min", round(min(ts), 4),
              "identical" if torch.equal(ll, ref) else "DIFFERENT", flush=True)
