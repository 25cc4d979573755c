import sys; sys.path.insert(0,'.')
import numpy as np, torch
from jaxoplanet_b200 import _lib, _array as A
dev=A.device(); n=100_000_000
M=torch.rand(n,dtype=torch.float64,device=dev)*(2*np.pi); e=torch.rand(n,dtype=torch.float64,device=dev)*0.99
s=torch.empty_like(M); c=torch.empty_like(M)
f=lambda: _lib.call("jxp_kepler_f64", M.data_ptr(),1,e.data_ptr(),1,n,s.data_ptr(),c.data_ptr(),0,0,0,A.stream_ptr())
for _ in range(2): f()
torch.cuda.synchronize(); best=1e9
for _ in range(5):
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True); a.record(); f(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
print("kepler 1e8 ms", best, "solves/s", n/best*1e3)
