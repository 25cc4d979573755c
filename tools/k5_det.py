import sys, numpy as np, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import workloads as W
from jaxoplanet_b200.light_curves.partials import transit_partials
theta, t = W.c5()
td = torch.as_tensor(t).cuda(); thd = torch.as_tensor(theta).cuda()
sigma = 5e-4
gen = transit_partials(thd[3:4], td, want_partials=False)["flux"][0]
y = 1.0 + gen
r1 = transit_partials(thd, td, y=y, sigma=sigma, want_flux=False, want_partials=False)
r2 = transit_partials(thd, td, y=y, sigma=sigma, want_flux=False, want_partials=False)
print("reduced vs reduced: ll equal", torch.equal(r1["loglike"], r2["loglike"]), "max diff", float((r1["loglike"]-r2["loglike"]).abs().max()))
s1 = transit_partials(thd, td, y=y, sigma=sigma, want_partials=False)
s2 = transit_partials(thd, td, y=y, sigma=sigma, want_partials=False)
print("stored vs stored: ll equal", torch.equal(s1["loglike"], s2["loglike"]), "max diff", float((s1["loglike"]-s2["loglike"]).abs().max()))
d = (s1["loglike"] - r1["loglike"]).abs()
print("stored vs reduced: n differing >1e-6:", int((d > 1e-6).sum()), "max", float(d.max()))
const = -td.numel() * (np.log(sigma) + 0.5 * np.log(2 * np.pi))
f = s1["flux"]
want = torch.stack([-0.5 * torch.sum(((y[None, :] - 1.0 - f[lo:lo+1024]) / sigma) ** 2, dim=1) + const for lo in range(0, 16384, 1024)]).reshape(-1)
print("stored ll vs torch reduction: n bad", int(((s1["loglike"] - want).abs() > 1e-5).sum()), " reduced ll vs torch: n bad", int(((r1["loglike"] - want).abs() > 1e-5).sum()))
bad = torch.nonzero((r1["loglike"] - want).abs() > 1e-5).flatten()[:10]; print("bad sets (reduced):", bad.tolist())
bad = torch.nonzero((s1["loglike"] - want).abs() > 1e-5).flatten()[:10]; print("bad sets (stored):", bad.tolist())
