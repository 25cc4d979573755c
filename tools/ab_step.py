"""A/B of library builds on config 3 (one GPU, 4096 sets) and config 4: run with JXP_LIB_PATH set per variant."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from jaxoplanet_b200 import _lib, _array as A
dev = A.device(); lib = _lib.load()
p, u_h, t_h, noise, sigma = bench.c3_inputs(seed=1)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
y = torch.as_tensor(1.0 + noise).to(dev); sig = torch.full((1,), sigma, dtype=torch.float64, device=dev)
for n in (4096, 512):
    prob = bench.Problem({k: v[:n] for k, v in p.items()}, u_h[:n], t_h, dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(120)]
    for i in range(60):
        flush.fill_(1)
        evs[2 * i].record(); prob.call(y=y, sig=sig); evs[2 * i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(10, 60)]
    print(n, "sets: ms per step mean", round(float(np.mean(ts)), 4), "min", round(min(ts), 4), flush=True)
