import sys; sys.path.insert(0,'.')
import numpy as np
sys.path.insert(0,'tests')
import test_transit_list as T
from jaxoplanet_b200 import workloads as W, _lib
p, ub, t, noise, sigma = W.c3(96, 12_000)
y = 1 + noise
sys_ = T._system(p)
rng = np.random.default_rng(3)
perm = rng.permutation(t.size)
tu, yu = t[perm], y[perm]
for rep in range(4):
    ll, items, flags = T._loglike_abi(sys_, ub, tu, yu, sigma)[:3]
    ll2, _, f2 = T._without_list(lambda: T._loglike_abi(sys_, ub, tu, yu, sigma))[:3]
    ll3, _, f3 = T._without_list(lambda: T._loglike_abi(sys_, ub, tu, yu, sigma))[:3]
    d = np.abs(ll-ll2); d23=np.abs(ll2-ll3)
    print(rep, 'flags', flags, f2, 'ndiff', int((d>0).sum()), 'max rel', float((d/np.abs(ll2)).max()), 'where', np.nonzero(d>0)[0][:10], 'll2 vs ll3', float(d23.max()))
