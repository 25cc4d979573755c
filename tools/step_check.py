"""Config 3 step (device-timed, L2 flushed, no exchange) on the shard every rank of an N-GPU run would own."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from jaxoplanet_b200 import _lib, _array as A
dev = A.device(); lib = _lib.load()
p, u_h, t_h, noise, sigma = bench.c3_inputs(seed=1)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
y = torch.as_tensor(1.0 + noise).to(dev); sig = torch.full((1,), sigma, dtype=torch.float64, device=dev)
for world in (1, 2, 4, 8):
    out = []
    for rank in range(world):
        lo, hi = rank * 4096 // world, (rank + 1) * 4096 // world
        prob = bench.Problem({k: v[lo:hi] for k, v in p.items()}, u_h[lo:hi], t_h, dev)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(60)]
        for i in range(30):
            flush.fill_(1)
            evs[2 * i].record(); prob.call(y=y, sig=sig); evs[2 * i + 1].record()
        torch.cuda.synchronize()
        ts = [evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(5, 30)]
        st = prob.stats()
        out.append((round(float(np.mean(ts)), 4), st["items"]))
        del prob
    print(f"{world} ranks: ms per step of each shard {[o[0] for o in out]} (max {max(o[0] for o in out)}; "
          f"one GPU / N = {round(0.2776 / world, 4)}); in-window samples {[o[1] for o in out]}", flush=True)
