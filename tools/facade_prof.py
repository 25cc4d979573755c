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
for _ in range(20): step()
print("ms/step", (time.perf_counter()-t0)/20*1e3)
pr=cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:4500])
