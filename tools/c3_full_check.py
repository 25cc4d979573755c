import sys; sys.path.insert(0,'.')
import numpy as np, torch, os
from jaxoplanet_b200 import workloads as W
from jaxoplanet_b200.light_curves import limb_dark_light_curve
from jaxoplanet_b200.light_curves.limb_dark import log_likelihood
p, u, t, _, sigma = W.c3()
system = W.system_from(p)
td = torch.as_tensor(t).cuda()
flux = limb_dark_light_curve(system, u)(td)[..., 0]
y = 1.0 + flux[7]
n = td.numel()
const = -n * (np.log(sigma) + 0.5 * np.log(2 * np.pi))
resid = (y[None, :] - 1.0 - flux)
want = -0.5 * torch.sum((resid / sigma) ** 2, dim=1) + const
bound = torch.sum(torch.abs(resid), dim=1) / sigma**2   # |d ll| per unit flux error
for mode in ("list","window"):
    if mode=="window": os.environ["JXP_NO_TL"]="1"
    ll = log_likelihood(system, u)(td, y, sigma)
    d = torch.abs(ll - want)
    rel = d/torch.abs(want)
    i = int(torch.argmax(rel))
    print(mode, "max rel", float(rel.max()), "at", i, "ll", float(ll[i]), "want", float(want[i]), "absdiff", float(d[i]),
          "max d/bound (flux-equivalent error)", float((d/bound.clamp_min(1e-300))[bound>0].max()), "argmax ok", int(torch.argmax(ll)), float(ll[7])-const)
