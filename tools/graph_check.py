"""The config-3 step captured into a CUDA graph (stream capture through the C ABI, side-stream fork / join
included) and replayed: same bits as the eager call? faster?"""
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
    prob.call(y=y, sig=sig); torch.cuda.synchronize()
    want = prob.ll.clone()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        prob.call(y=y, sig=sig)  # (warm-up on the capture stream)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        prob.call(y=y, sig=sig)
    prob.ll.zero_(); g.replay(); torch.cuda.synchronize()
    same = bool(torch.equal(prob.ll, want))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(120)]
    out = {}
    for what in ("eager", "graph", "eager", "graph"):
        for i in range(60):
            flush.fill_(1)
            evs[2 * i].record()
            if what == "eager":
                prob.call(y=y, sig=sig)
            else:
                g.replay()
            evs[2 * i + 1].record()
        torch.cuda.synchronize()
        ts = [evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(10, 60)]
        out.setdefault(what, []).append(round(float(np.mean(ts)), 4))
    print(n, "sets: replay bit-identical", same, "ms per step", out, flush=True)
