import sys, os; sys.path.insert(0,'.')
import numpy as np
from jaxoplanet_b200.core.limb_dark import light_curve
from oracle import ref_numpy as ref
n=20000
rng=np.random.default_rng(1)
r=rng.uniform(0.01,0.3,n); b=rng.uniform(0,1,n)
b[100]=np.nan; r[101]=np.nan; b[102]=-0.3
for u in ([0.3,0.2],[0.4],[]):
    got=light_curve(np.array(u),b,r)
    os.environ["JXP_NO_LD_CLASSES"]="1"; plain=light_curve(np.array(u),b,r); os.environ.pop("JXP_NO_LD_CLASSES")
    with np.errstate(all='ignore'):
        want=ref.ld_light_curve(np.array(u),b,r)
    print(u, "cls", got[100:103], "plain", plain[100:103], "oracle", want[100:103])
