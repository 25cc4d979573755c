import sys; sys.path.insert(0,'.')
import cProfile, pstats, io, time
import numpy as np, torch
from jaxoplanet_b200 import workloads as W
from jaxoplanet_b200.light_curves.limb_dark import log_likelihood
p, u, t, noise, sigma = W.c3()
y = 1 + noise
def step():
    return log_likelihood(W.system_from(p), u)(t, y, sigma)
for _ in range(5): step()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(50): step()
print("ms/step", (time.perf_counter()-t0)/50*1e3)
# host time before the result is awaited: the same step with the final D2H replaced by nothing
from jaxoplanet_b200.light_curves.limb_dark import FusedSpec, _run_fused
def host_only():
    s = W.system_from(p)
    return _run_fused(FusedSpec(s, u), t, want="loglike", y=y, sigma=sigma)
for _ in range(3): host_only()
torch.cuda.synchronize()
ts=[]
for _ in range(30):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=host_only(); ts.append(time.perf_counter()-t0)
print("host prelude (launches issued, no sync) ms", 1e3*np.median(ts))
ts=[]
for _ in range(30):
    torch.cuda.synchronize(); t0=time.perf_counter(); s=W.system_from(p); ts.append(time.perf_counter()-t0)
print("system_from ms", 1e3*np.median(ts))
pr=cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('tottime').print_stats(45); print(s.getvalue()[:7000])
