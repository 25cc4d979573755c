"""Where does the fp32 mode lose precision?  C3 bench inputs: flux f32 vs f64 (device), worst points with
their (b, r) from the oracle; and K2 alone (f32 vs f64) on the same (b, r)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxoplanet_b200 import workloads as W
from jaxoplanet_b200.light_curves.limb_dark import FusedSpec, _run_fused
from jaxoplanet_b200.core.limb_dark import light_curve as ld
from oracle import ref_numpy as ref

p, u, t, noise, sigma = W.c3()
system = W.system_from(p)
td = torch.as_tensor(t).cuda()
spec = FusedSpec(system, u)
f64 = _run_fused(spec, td, want="flux_sum")
f32 = _run_fused(spec, td, want="flux_sum", dtype=torch.float32)
err = (f32.double() - f64).abs()
print("max err", float(err.max()), "n > 1e-6:", int((err > 1e-6).sum()), "n > 3e-7:", int((err > 3e-7).sum()),
      "in transit", int((f64 != 0).sum()))
flat = torch.topk(err.reshape(-1), 12).indices.cpu().numpy()
for ix in flat:
    s, i = divmod(int(ix), t.size)
    bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
    x, y, z = ref.relative_position(bd, t[i:i + 1])
    b = float(np.sqrt(x**2 + y**2)[0]); r = float(p["radius"][s])
    k2_64 = float(ld(u[s], np.array([b]), np.array([r]))[0])
    k2_32 = float(ld(u[s].astype(np.float32), np.array([b], dtype=np.float32), np.array([r], dtype=np.float32))[0])
    print(f"set {s} i {i} b {b:.9f} r {r:.6f} 1-r {1-r:.6f} 1+r {1+r:.6f} b-(1-r) {b-(1-r):+.3e} b-(1+r) {b-(1+r):+.3e} "
          f"f64 {float(f64[s,i]):+.9e} f32 {float(f32[s,i]):+.9e} err {float(err[s,i]):.2e} | K2 f64 {k2_64:+.9e} f32 {k2_32:+.9e} u {u[s]}")
# histogram of err vs distance to contact
