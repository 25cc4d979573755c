"""Scratch timing of the kernels (development aid; bench.py is the contract)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import _lib, _array as A
import jaxoplanet_b200 as jxp
from jaxoplanet_b200.orbits.keplerian import Central, System
from jaxoplanet_b200.light_curves.limb_dark import FusedSpec, _run_fused, pack_bodies

def ev_time(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

out={}
dev=A.device()
# FP64 / FP32 peak
for name, dt, key in (("jxp_peak_fma_f64", torch.float64, "fp64"), ("jxp_peak_fma_f32", torch.float32, "fp32")):
    for threads in (256, 512, 1024):
        blocks=148*(2048//threads)
        sink=torch.empty(blocks*threads, dtype=dt, device=dev)
        iters=20000
        f=lambda: _lib.call(name, sink.data_ptr(), blocks, threads, iters, A.stream_ptr())
        tmin,tmed=ev_time(f)
        flops=2.0*blocks*threads*iters*16
        out[f"peak_{key}_t{threads}"]={"ms":tmin,"tflops":flops/tmin/1e9}
print(json.dumps(out, indent=1)); sys.stdout.flush()

# K1: 1e8 pairs
rng=np.random.default_rng(0)
n=100_000_000
M=torch.rand(n, dtype=torch.float64, device=dev)*(2*np.pi)
e=torch.rand(n, dtype=torch.float64, device=dev)*0.99
s=torch.empty_like(M); c=torch.empty_like(M)
f=lambda: _lib.call("jxp_kepler_f64", M.data_ptr(),1,e.data_ptr(),1,n,s.data_ptr(),c.data_ptr(),0,0,0,A.stream_ptr())
tmin,tmed=ev_time(f)
out["kepler_1e8"]={"ms":tmin,"solves_per_s":n/tmin*1e3,"GBps":32*n/tmin/1e6}
print(json.dumps(out["kepler_1e8"])); sys.stdout.flush()
del M,e,s,c

# C3: 4096 sets x 1e5 times
def c3(n_sets, n_t, seed=1):
    rng=np.random.default_rng(seed)
    P=rng.uniform(2,5,n_sets)
    p=dict(period=P,time_transit=rng.uniform(0,1,n_sets)*P,eccentricity=rng.uniform(0,.3,n_sets),omega_peri=rng.uniform(-np.pi,np.pi,n_sets),impact_param=rng.uniform(0,1,n_sets),radius=rng.uniform(.03,.15,n_sets))
    q=rng.uniform(0,1,(n_sets,2))
    u=np.stack([2*np.sqrt(q[:,0])*q[:,1],np.sqrt(q[:,0])*(1-2*q[:,1])],1)
    t=np.arange(n_t)*(2.0/1440.0)
    with jxp.batched():
        sys_=System(Central()).add_body(**p)
    return sys_,u,t
sys_,u,t=c3(4096,100_000)
y=1+5e-4*np.random.default_rng(42).normal(size=t.shape); sigma=5e-4
bodies,ns,_=pack_bodies(sys_)
bodies=bodies.to(dev).contiguous(); ud=torch.as_tensor(u).to(dev).contiguous(); td=torch.as_tensor(t).to(dev)
yd=torch.as_tensor(y).to(dev); sd=torch.full((1,),sigma,dtype=torch.float64,device=dev)
wsb=_lib.load().jxp_system_workspace_bytes(ns,1,td.numel()); ws=torch.empty(wsb,dtype=torch.uint8,device=dev)
ll=torch.empty(ns,dtype=torch.float64,device=dev)
for flags,label in ((0,"window"),(1,"dense")):
    f=lambda: _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(),ns,1,ud.data_ptr(),2,2,td.data_ptr(),td.numel(),0.0,0,0,10,flags,0,0,yd.data_ptr(),sd.data_ptr(),0,ll.data_ptr(),ws.data_ptr(),wsb,A.stream_ptr())
    tmin,tmed=ev_time(f,n=3,warm=1)
    out[f"c3_loglike_{label}"]={"ms":tmin,"points_per_s":ns*td.numel()/tmin*1e3}
    print(label, json.dumps(out[f"c3_loglike_{label}"])); sys.stdout.flush()
fs=torch.empty((ns,td.numel()),dtype=torch.float64,device=dev)
for flags,label in ((0,"window"),(1,"dense")):
    f=lambda: _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(),ns,1,ud.data_ptr(),2,2,td.data_ptr(),td.numel(),0.0,0,0,10,flags,0,fs.data_ptr(),0,0,0,0,ws.data_ptr(),wsb,A.stream_ptr())
    tmin,tmed=ev_time(f,n=3,warm=1)
    out[f"c3_flux_{label}"]={"ms":tmin,"points_per_s":ns*td.numel()/tmin*1e3,"f_in":float((fs!=0).double().mean())}
    print(label, json.dumps(out[f"c3_flux_{label}"])); sys.stdout.flush()
del fs
torch.cuda.empty_cache()
from jaxoplanet_b200 import workloads as W
theta,t5=W.c5(16384,50_000)
thd=torch.as_tensor(theta).to(dev).contiguous(); t5d=torch.as_tensor(t5).to(dev)
star=torch.tensor([1.0,1.0],dtype=torch.float64,device=dev)
ns5=thd.shape[0]; nt5=t5d.numel()
wsb5=_lib.load().jxp_transit_workspace_bytes(ns5,nt5); ws5=torch.empty(wsb5,dtype=torch.uint8,device=dev)
for dt,nm in ((torch.float64,"f64"),(torch.float32,"f32")):
    fl=torch.empty((ns5,nt5),dtype=dt,device=dev); pt=torch.empty((ns5,8,nt5),dtype=dt,device=dev)
    for flags,label in ((0,"window"),(1,"dense")):
        f=lambda: _lib.call("jxp_transit_partials_"+nm, thd.data_ptr(),ns5,star.data_ptr(),0,t5d.data_ptr(),nt5,0.0,0,0,10,flags,fl.data_ptr(),pt.data_ptr(),0,0,0,0,0,ws5.data_ptr(),wsb5,A.stream_ptr())
        tmin,tmed=ev_time(f,n=3,warm=1)
        bytes_=ns5*nt5*9*fl.element_size()
        out[f"c5_{nm}_{label}"]={"ms":tmin,"points_per_s":ns5*nt5/tmin*1e3,"GBps":bytes_/tmin/1e6}
        print("c5",nm,label,json.dumps(out[f"c5_{nm}_{label}"])); sys.stdout.flush()
    del fl,pt
json.dump(out, open("gpurun_out/quick_bench.json","w"), indent=1)
