"""bench.secondary_configs alone in a fresh process (is the float K5 figure of bench.py a property of the call or of
what ran before it?)."""
import sys, json, torch
sys.path.insert(0, ".")
import bench
from jaxoplanet_b200 import _lib, _array as A
dev = A.device(); lib = _lib.load()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = bench.secondary_configs(dev, lib, ev0, ev1, 0, 1, 36.6, lambda x: x, lambda: torch.cuda.synchronize())
for k in ("c5_flux_plus_8_partials_f64", "c5_flux_plus_8_partials_f32"):
    print(k, round(out[k]["ms"], 3), round(out[k]["hbm_frac"], 3))
