import sys, numpy as np, torch
sys.path.insert(0, ".")
from jaxoplanet_b200 import _lib, _array as A, workloads as W
dev = A.device()
n_sets, n_t = 16384, 50000
theta, t5 = W.c5(n_sets, n_t)
DT = torch.float64
NM = "f64"
def run(theta, want_out, want_ll, label):
    thd = torch.as_tensor(theta).to(dev).contiguous(); t5d = torch.as_tensor(t5).to(dev)
    star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
    wsb5 = _lib.load().jxp_transit_workspace_bytes(n_sets, n_t); ws5 = torch.empty(wsb5, dtype=torch.uint8, device=dev)
    fl = torch.empty((n_sets, n_t), dtype=DT, device=dev) if want_out else None
    pt = torch.empty((n_sets, 8, n_t), dtype=DT, device=dev) if want_out else None
    y = torch.ones(n_t, dtype=DT, device=dev); sg = torch.full((1,), 5e-4, dtype=DT, device=dev)
    ll = torch.empty(n_sets, dtype=DT, device=dev) if want_ll else None
    dll = torch.empty((n_sets, 8), dtype=DT, device=dev) if want_ll else None
    f = lambda: _lib.call("jxp_transit_partials_" + NM, thd.data_ptr(), n_sets, star.data_ptr(), 0, t5d.data_ptr(), n_t, 0.0, 0, 0, 10, 0,
                          A.ptr(fl), A.ptr(pt), y.data_ptr() if want_ll else 0, sg.data_ptr() if want_ll else 0, 0, A.ptr(ll), A.ptr(dll), ws5.data_ptr(), wsb5, A.stream_ptr())
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    print(label, best, "ms")
th2 = theta.copy(); th2[:, 4] = 5.0   # impact parameter 5: never transits
for DT, NM in ((torch.float64, "f64"), (torch.float32, "f32")):
    for mask, what in ((0, "bulk"), (_lib.JXP_DEV_NO_BULK_FILL, "warp stores")):
        with _lib.dev_options(mask):
            run(theta, True, False, f"{NM} {what}: flux+partials stored")
            run(th2, True, False, f"{NM} {what}: never-transiting sets (pure zero fill)")
    with _lib.dev_options(_lib.JXP_DEV_NO_WALK):
        run(theta, False, True, f"{NM}: loglike + gradient only (window-test kernel, no row stores)")
