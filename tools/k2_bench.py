"""Scratch timing of K2 (core.limb_dark.light_curve) on 1e8 random in-transit points: class-sorted kernel vs plain."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import _lib, _array as A
dev = A.device(); n = 100_000_000
torch.manual_seed(0)
rr = torch.rand(n, dtype=torch.float64, device=dev) * 0.12 + 0.03
bb = torch.rand(n, dtype=torch.float64, device=dev) * (1 + rr)
u = torch.tensor([0.3, 0.2], dtype=torch.float64, device=dev)
out = torch.empty(n, dtype=torch.float64, device=dev)
f = lambda: _lib.call("jxp_limb_dark_f64", u.data_ptr(), 2, bb.data_ptr(), 1, rr.data_ptr(), 1, n, 10, out.data_ptr(), None, None, None, A.stream_ptr())
res = {}
for label in ("classes", "plain"):
    _lib.load().jxp_set_dev_options(_lib.JXP_DEV_NO_LD_CLASSES if label == "plain" else 0)
    for _ in range(2): f()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record(); f(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    res[label] = out.clone()
    print(label, "ms", best, "evals/s", n / best * 1e3)
print("max abs diff", float((res["classes"] - res["plain"]).abs().max()), "inside fraction", float(((bb < 1 - rr)).double().mean()))
