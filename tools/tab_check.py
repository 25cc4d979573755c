"""Scratch check of the walk path's function tables (csrc/jxp_tables.cuh): config-3 log-likelihoods and fluxes with
and without the tables, against each other and (a few sets) against the oracle; table statistics; timing."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxoplanet_b200 import _lib, _array as A, workloads as W
from jaxoplanet_b200.light_curves.limb_dark import pack_bodies
from oracle import ref_numpy as ref

dev = A.device()
lib = _lib.load()
n_t = int(os.environ.get("NT", 100_000))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def stats(ws, ns, nb):
    st = (ctypes.c_int64 * 2)(); _lib.call("jxp_system_stats", ws.data_ptr(), ns, nb, n_t, ctypes.addressof(st), A.stream_ptr())
    ls = (ctypes.c_int64 * 4)(); _lib.call("jxp_system_list_stats", ws.data_ptr(), ns, nb, n_t, ctypes.addressof(ls), A.stream_ptr())
    ts = (ctypes.c_int64 * 4)(); _lib.call("jxp_system_walk_stats", ws.data_ptr(), ns, nb, n_t, ctypes.addressof(ts), A.stream_ptr())
    return dict(kepler=int(st[0]), ld=int(st[1]), items=int(ls[0]), flags=int(ls[1]), tab_evals=int(ts[0]), tab_sets=int(ts[2]), n_fix=int(ts[1]), no_tables=int(ts[3]))


def run(p, u, t, noise, sigma, label, n_oracle=4, timing=True):
    n_sets = len(p["period"])
    bodies, ns, _ = pack_bodies(W.system_from(p))
    bodies = bodies.to(dev).contiguous(); ud = torch.as_tensor(u).to(dev).contiguous(); td = torch.as_tensor(t).to(dev)
    y = 1 + noise
    yd = torch.as_tensor(y).to(dev); sd = torch.full((1,), sigma, dtype=torch.float64, device=dev)
    wsb = lib.jxp_system_workspace_bytes(ns, 1, n_t); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ll = torch.empty(ns, dtype=torch.float64, device=dev)
    fx = torch.empty(ns, n_t, dtype=torch.float64, device=dev) if ns <= 512 else None
    call_ll = lambda: _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), ns, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_t,
                                0.0, 0, 0, 10, 0, None, None, yd.data_ptr(), sd.data_ptr(), 0, ll.data_ptr(), ws.data_ptr(), wsb,
                                A.stream_ptr())
    call_fx = lambda: _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), ns, 1, ud.data_ptr(), 2, 2, td.data_ptr(), n_t,
                                0.0, 0, 0, 10, 0, None, fx.data_ptr(), None, None, 0, None, ws.data_ptr(), wsb, A.stream_ptr())
    res = {"case": label, "n_sets": n_sets}
    out = {}
    for name, opt in (("tab", 0), ("direct", _lib.JXP_DEV_NO_TABLES)):
        lib.jxp_set_dev_options(opt | _lib.JXP_DEV_WALK_ANY_SIZE)
        call_ll(); torch.cuda.synchronize()
        res[name] = stats(ws, ns, 1)
        out[name + "_ll"] = ll.clone()
        if fx is not None:
            call_fx(); torch.cuda.synchronize()
            out[name + "_fx"] = fx.clone()
            res[name + "_fx"] = stats(ws, ns, 1)
        if timing:
            for _ in range(3): call_ll()
            torch.cuda.synchronize()
            tot = 0.0; n = 10
            for _ in range(n):
                flush.fill_(1)
                a, b = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                a.record(); call_ll(); b.record(); torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            res[name]["step_ms"] = round(tot / n, 4)
            for which, o2 in (("eval_ms", 0), ("plan_ms", _lib.JXP_DEV_PROFILE_PLAN), ("tables_ms", _lib.JXP_DEV_PROFILE_TABLES), ("items_ms", _lib.JXP_DEV_PROFILE_ITEMS)):
                lib.jxp_set_dev_options(opt | o2 | _lib.JXP_DEV_WALK_ANY_SIZE)
                kt = 0.0
                for _ in range(n):
                    flush.fill_(1)
                    ka, kb = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                    ka.record(); kb.record()
                    lib.jxp_profile_events(ctypes.c_void_p(ka.cuda_event), ctypes.c_void_p(kb.cuda_event))
                    call_ll()
                    lib.jxp_profile_events(None, None)
                    torch.cuda.synchronize()
                    kt += ka.elapsed_time(kb)
                res[name][which] = round(kt / n, 4)
    lib.jxp_set_dev_options(0)
    d = (out["tab_ll"] - out["direct_ll"]).abs() / out["direct_ll"].abs()
    res["ll_rel_diff_tab_vs_direct"] = float(d.max())
    res["ll_abs_diff_tab_vs_direct"] = float((out["tab_ll"] - out["direct_ll"]).abs().max())
    res["ll_abs_max"] = float(out["direct_ll"].abs().max())
    res["ll_nan"] = int(torch.isnan(out["tab_ll"]).sum())
    if fx is not None:
        df = (out["tab_fx"] - out["direct_fx"]).abs()
        res["flux_abs_diff_tab_vs_direct"] = float(df.max())
        res["flux_zero_pattern_equal"] = bool(torch.equal(out["tab_fx"] == 0, out["direct_fx"] == 0))
        # oracle on a few sets
        worst_t = worst_d = 0.0
        for i in np.linspace(0, n_sets - 1, n_oracle).astype(int):
            bd = ref.orbital_body(**{k: v[i] for k, v in p.items()})
            want = ref.system_light_curve([bd], u[i], t).sum(-1)
            worst_t = max(worst_t, float(np.max(np.abs(out["tab_fx"][i].cpu().numpy() - want))))
            worst_d = max(worst_d, float(np.max(np.abs(out["direct_fx"][i].cpu().numpy() - want))))
        res["flux_abs_err_vs_oracle"] = dict(tab=worst_t, direct=worst_d)
    print(json.dumps(res), flush=True)


for n_sets in [int(x) for x in os.environ.get("NS", "256,4096").split(",")]:
    p, u, t, noise, sigma = W.c3(n_sets, n_t)
    run(p, u, t, noise, sigma, "c3")
if os.environ.get('NOADV'):
    sys.exit(0)
# adversarial orbits (tests/test_gpu_parity.py::test_window_is_conservative_on_adversarial_orbits)
rng = np.random.default_rng(77)
n_sets = 300
P = 10 ** rng.uniform(-0.5, 1.3, n_sets)
p = dict(period=P, time_transit=rng.uniform(-1, 1, n_sets) * P, eccentricity=rng.uniform(0, 0.93, n_sets),
         omega_peri=rng.uniform(-np.pi, np.pi, n_sets), impact_param=rng.uniform(0, 1.4, n_sets),
         radius=10 ** rng.uniform(-2.3, -0.3, n_sets))
u = np.tile(np.array([0.3, 0.2]), (n_sets, 1))
t = np.sort(rng.uniform(-30, 30, n_t))
run(p, u, t, 5e-4 * rng.normal(size=n_t), 5e-4, "adversarial", n_oracle=12, timing=False)
