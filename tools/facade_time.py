import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import jaxoplanet_b200 as jxp
from jaxoplanet_b200 import workloads as W
from jaxoplanet_b200.light_curves.limb_dark import log_likelihood
p, ub, tb, noise, sigma = W.c3(4096, 100000)
y = 1.0 + noise
td = torch.as_tensor(tb).cuda(); yd = torch.as_tensor(y).cuda()
def step():
    sys_ = W.system_from(p)
    return log_likelihood(sys_, ub)(td, yd, sigma)
for _ in range(20): step()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(50): ll = step()
    torch.cuda.synchronize()
    print("facade step (System build + log_likelihood, numpy params):", (time.perf_counter() - t0) / 50 * 1e3, "ms")
t0 = time.perf_counter()
for _ in range(10): sys_ = W.system_from(p)
print("  of which System construction:", (time.perf_counter() - t0) / 10 * 1e3, "ms")
f = log_likelihood(sys_, ub)
for _ in range(3): f(td, yd, sigma)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): ll = f(td, yd, sigma)
torch.cuda.synchronize()
print("  log_likelihood call only:", (time.perf_counter() - t0) / 10 * 1e3, "ms")
