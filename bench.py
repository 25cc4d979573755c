#!/usr/bin/env python
"""Benchmark of the jaxoplanet hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json config 3 -- 4096 parameter sets x 1e5 time samples
of a Keplerian transit light curve with quadratic limb darkening, fused with the Gaussian
log-likelihood -- PER GPU (weak scaling: rank g evaluates its own 4096 sets; the per-set
log-likelihoods are all-gathered over NCCL).  A "step" is one evaluation of that batch, the
unit a sampler performs per proposal.  metric = light-curve points/s (one point = one time
sample of one parameter set).

value    device-resident inputs, CUDA events around each step, L2 flushed between steps
e2e      same step through the C ABI with pinned HOST buffers: H2D of every input and D2H of
         the log-likelihoods inside the timed region
roofline dominant kernel (k_system) timed alone with CUDA events inside the timed steps;
         achieved = A-table flops of the work it EXECUTED (counters read back from the
         device) / its mean duration; peak = FP64 FMA peak measured in this run
cpu_baseline / --impl reference: the C/OpenMP restatement of the reference (oracle/ref_c.c)
         on the host cores, on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# A-table v1 (SURVEY.md 8d / DESIGN.md): algorithmic flops per unit of the reference formulas
W_KEPLER = 266.0
W_ORBIT = 36.0
W_LD = 746.0
N_SETS = 4096
N_TIMES = 100_000
METRIC = "light-curve points/s (Keplerian orbit + quadratic limb darkening + Gaussian log-likelihood, fp64)"


_STDOUT_FD = None


def capture_stdout():
    """Libraries (NCCL banners, torchrun notices) may print to stdout; the contract is ONE JSON
    line there.  Route fd 1 to stderr for the run and keep the real stdout for emit()."""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def workload_config(n_gpus, extra=None):
    cfg = {
        "workload": f"C3: {N_SETS} parameter sets x {N_TIMES} times (2-min cadence), 1 planet, quadratic LD, "
                    "fused Gaussian log-likelihood, per GPU",
        "sets_per_gpu": N_SETS,
        "n_times": N_TIMES,
        "n_bodies": 1,
        "parallelism": f"set-sharded x{n_gpus}",
        "cache": "L2 flushed between steps (256 MiB write); per-step CUDA events summed",
    }
    if extra:
        cfg.update(extra)
    return cfg


def build_inputs(seed):
    """Host-side inputs of one rank: C-ABI body records, u, t, noise, sigma."""
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    p, u, t, noise, sigma = W.c3(N_SETS, N_TIMES, seed=seed)
    bodies, n_sets, _ = pack_bodies(W.system_from(p))
    return p, bodies.cpu().numpy().copy(), u, t, noise, sigma


# ------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (C/OpenMP restatement) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_points_per_s(bodies, u, t, y, sigma, budget_s=12.0):
    from oracle import ref_c

    ref_c.build()
    ref_c.use_all_cores()
    n_try = 2
    t0 = time.perf_counter()
    ref_c.system(bodies[:n_try], u[:n_try], t, y=y, sigma=sigma, want_flux=False)
    dt = time.perf_counter() - t0
    threads = ref_c.max_threads()
    n_sample = int(min(len(bodies), max(threads, budget_s / max(dt / n_try, 1e-9))))
    n_sample = max(threads, (n_sample // threads) * threads)
    n_sample = min(n_sample, len(bodies))
    t0 = time.perf_counter()
    ref_c.system(bodies[:n_sample], u[:n_sample], t, y=y, sigma=sigma, want_flux=False)
    dt = time.perf_counter() - t0
    return n_sample * len(t) / dt, n_sample, threads, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    _, bodies, u, t, noise, sigma = build_inputs(seed=1)
    y = 1.0 + noise
    from oracle import ref_c

    ref_c.build()
    threads = ref_c.use_all_cores()
    # size one step to ~2 s of CPU work
    n_try = max(2, threads)
    t0 = time.perf_counter()
    ref_c.system(bodies[:n_try], u[:n_try], t, y=y, sigma=sigma, want_flux=False)
    per_set = (time.perf_counter() - t0) / n_try
    n_sample = int(min(N_SETS, max(threads, 2.0 / max(per_set, 1e-9))))
    n_sample = max(threads, (n_sample // threads) * threads)
    for _ in range(args.warmup):
        ref_c.system(bodies[:n_sample], u[:n_sample], t, y=y, sigma=sigma, want_flux=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_c.system(bodies[:n_sample], u[:n_sample], t, y=y, sigma=sigma, want_flux=False)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_sample * N_TIMES / dt
    sample = f"{n_sample} of {N_SETS} sets x {N_TIMES} times per step (linear in sets)"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, {"cache": "n/a (CPU)"}),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": sample,
                         "note": "C/OpenMP restatement of the reference formulas (oracle/ref_c.c); "
                                 "jax is not installable in this image, so this is not jax[cpu]"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t_lo, t_hi):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < t_lo - 0.05 or ts > t_hi + 0.15:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    stream = A.stream_ptr

    p, bodies_h, u_h, t_h, noise, sigma = build_inputs(seed=1 + rank)
    n_sets, n_times = bodies_h.shape[0], t_h.shape[0]
    bodies = torch.as_tensor(bodies_h).to(dev).contiguous()
    u = torch.as_tensor(u_h).to(dev).contiguous()
    t = torch.as_tensor(t_h).to(dev)
    sig = torch.full((1,), sigma, dtype=torch.float64, device=dev)
    wsb = int(lib.jxp_system_workspace_bytes(n_sets, 1, n_times))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ll = torch.empty(n_sets, dtype=torch.float64, device=dev)
    ll_all = torch.empty(world * n_sets, dtype=torch.float64, device=dev) if world > 1 else None

    def call(flags=0, y_=None, flux_sum=None, loglike=ll, bodies_=bodies, u_=u, t_=t, sig_=sig):
        _lib.call("jxp_system_light_curve_f64", bodies_.data_ptr(), n_sets, 1, u_.data_ptr(), 2, 2,
                  t_.data_ptr(), n_times, 0.0, 0, 0, 10, flags, 0, A.ptr(flux_sum), A.ptr(y_),
                  sig_.data_ptr() if y_ is not None else 0, 0, A.ptr(loglike) if y_ is not None else 0,
                  ws.data_ptr(), wsb, stream())

    # observed flux: 1 + model of this rank's set 0 + noise (synthetic), computed once, untimed
    f0 = torch.empty((n_sets, n_times), dtype=torch.float64, device=dev)
    call(flux_sum=f0, loglike=None)
    y = (1.0 + f0[0] + torch.as_tensor(noise).to(dev)).contiguous()
    f_in = float((f0 != 0).double().mean())
    del f0
    torch.cuda.synchronize()

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush_buf.fill_(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(flags=0):
        call(flags=flags, y_=y)
        if world > 1:
            w = dist.all_gather_into_tensor(ll_all, ll, async_op=True)
            w.wait()

    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    kev0 = torch.cuda.Event(enable_timing=True)
    kev1 = torch.cuda.Event(enable_timing=True)
    for e in (ev0, ev1, kev0, kev1):
        e.record()
    torch.cuda.synchronize()

    def timed_steps(n, flags=0, kernel_events=True):
        tot, ktot = 0.0, 0.0
        for _ in range(n):
            flush_l2()
            if kernel_events:
                lib.jxp_profile_events(ctypes.c_void_p(kev0.cuda_event), ctypes.c_void_p(kev1.cuda_event))
            ev0.record()
            step(flags)
            ev1.record()
            lib.jxp_profile_events(None, None)
            torch.cuda.synchronize()
            tot += ev0.elapsed_time(ev1)
            if kernel_events:
                ktot += kev0.elapsed_time(kev1)
        return tot, ktot

    # FP64 pipe peak, measured now (MEASURED_PEAKS.json has no FP64 figure)
    def fma_peak(name, dtype):
        blocks, threads, iters = 148 * 2, 1024, 20000
        sink = torch.empty(blocks * threads, dtype=dtype, device=dev)
        best = 1e30
        for i in range(5):
            ev0.record()
            _lib.call(name, sink.data_ptr(), blocks, threads, iters, stream())
            ev1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, ev0.elapsed_time(ev1))
        return 2.0 * blocks * threads * iters * 16 / best / 1e9  # TFLOP/s

    fp64_peak = fma_peak("jxp_peak_fma_f64", torch.float64)

    # ---- value: device-resident inputs --------------------------------------------------
    timed_steps(max(args.warmup, 3), kernel_events=False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    t_lo = time.perf_counter()
    tot_ms, k_ms = timed_steps(args.steps)
    barrier()
    t_hi = time.perf_counter()
    clocks = sampler.stop(t_lo, t_hi) if rank == 0 else None
    stats = (ctypes.c_int64 * 2)()
    _lib.call("jxp_system_stats", ws.data_ptr(), n_sets, 1, n_times, ctypes.addressof(stats), stream())
    n_kep, n_ld = int(stats[0]), int(stats[1])
    lstats = (ctypes.c_int64 * 4)()
    _lib.call("jxp_system_list_stats", ws.data_ptr(), n_sets, 1, n_times, ctypes.addressof(lstats), stream())
    list_items, list_used = int(lstats[0]), int(lstats[0]) > 0 and int(lstats[1]) == 0
    # sorted time stamps: the in-window samples go on a work list (k_tl_*); the window-test kernels are
    # still launched behind them and return at once (the choice is made on the device)
    launches = (["k_sys_prep_ll", "k_tl_build", "k_tl_eval", "k_tl_reduce", "k_system_pair (returns at once)",
                 "k_loglike_final (returns at once)"] if list_used
                else ["k_sys_prep_ll", "k_system_pair", "k_loglike_final"])

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    tot_ms = max_over_ranks(tot_ms)
    ms_per_step = tot_ms / args.steps
    points_per_step = world * n_sets * n_times
    value = points_per_step / (ms_per_step * 1e-3)

    # ---- e2e: pinned host buffers through the C ABI ------------------------------------------
    # the step starts from the SAMPLED orbital elements (what an MCMC step has on the host): K0 derives
    # the per-set constants of OrbitalBody.__init__ on the device, inside the timed region
    from jaxoplanet_b200.light_curves.limb_dark import pack_elements

    # one pinned arena on the host and its mirror on the device: [elements | u | t | y | sigma] go over in
    # ONE copy per step (five separate copies cost ~8 us of latency each on a 2 MB transfer)
    parts = [np.ascontiguousarray(pack_elements(n_sets, **p), dtype=np.float64).reshape(-1),
             np.ascontiguousarray(u_h, dtype=np.float64).reshape(-1), np.ascontiguousarray(t_h, dtype=np.float64),
             y.cpu().numpy().astype(np.float64), np.array([sigma], dtype=np.float64)]
    offs = np.concatenate([[0], np.cumsum([((x.size + 31) // 32) * 32 for x in parts])])  # 256-byte aligned
    harena = torch.zeros(int(offs[-1]), dtype=torch.float64).pin_memory()
    for x, o in zip(parts, offs[:-1]):
        harena[int(o):int(o) + x.size] = torch.as_tensor(x)
    darena = torch.empty_like(harena, device=dev)
    del_, du, dt_, dy, ds = (darena[int(o):int(o) + x.size] for x, o in zip(parts, offs[:-1]))
    hll = torch.empty(n_sets, dtype=torch.float64).pin_memory()
    db = torch.empty((n_sets, 1, _lib.JXP_BODY_NFIELDS), dtype=torch.float64, device=dev)
    h2d = sum(x.size * 8 for x in parts)
    d2h = hll.numel() * hll.element_size()

    def e2e_step():
        darena.copy_(harena, non_blocking=True)
        _lib.call("jxp_orbital_elements_f64", del_.data_ptr(), n_sets, db.data_ptr(), stream())
        _lib.call("jxp_system_light_curve_f64", db.data_ptr(), n_sets, 1, du.data_ptr(), 2, 2,
                  dt_.data_ptr(), n_times, 0.0, 0, 0, 10, 0, 0, 0, dy.data_ptr(), ds.data_ptr(), 0,
                  ll.data_ptr(), ws.data_ptr(), wsb, stream())
        if world > 1:
            w = dist.all_gather_into_tensor(ll_all, ll, async_op=True)
            w.wait()
        hll.copy_(ll, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step()
        e2e_ms += (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(e2e_ms) / args.steps
    e2e_value = points_per_step / (e2e_ms * 1e-3)

    # ---- the same step through the Python facade (what a user of the reference API writes): the
    # sampled orbital elements, t and y are HOST NumPy arrays, System(...) is rebuilt (per-set
    # constants of OrbitalBody.__init__ re-derived) every step, the result comes back as NumPy
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import log_likelihood as _facade_ll

    y_np = y.cpu().numpy()

    def facade_step():
        return _facade_ll(W.system_from(p), u_h)(t_h, y_np, sigma)

    for _ in range(5):
        facade_step()
    barrier()
    py_ms = 0.0
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ll_np = facade_step()
        py_ms += (time.perf_counter() - t0) * 1e3
    assert isinstance(ll_np, np.ndarray) and ll_np.shape == (n_sets,)
    py_ms = max_over_ranks(py_ms) / args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0) -------------------------------------------
    k_ms_avg = k_ms / args.steps
    flops = n_kep * (W_KEPLER + W_ORBIT) + n_ld * W_LD
    achieved = flops / (k_ms_avg * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[
            "k_tl_eval" if list_used else "k_system_pair"]
        traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "kernel": "k_tl_eval<1>" if list_used else "k_system_pair<1>", "achieved": achieved,
        "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": traffic,
        "traffic_note": "DRAM bytes per launch from the ncu capture in profiles/ (FP64-bound kernel; the list "
                        "path reads its 8-byte items and writes the chi^2 terms over them: 13 M items x 8 B each way, "
                        "part of the writes still in L2 when the kernel ends)",
        "list_items": list_items,
        "peak_source": "DFMA loop (jxp_peak_fma_f64) measured in this run; MEASURED_PEAKS.json has no FP64 figure",
        "kernel_ms": k_ms_avg, "kernel_share_of_step": k_ms_avg / ms_per_step,
        "executed": {"kepler_solves": n_kep, "ld_evaluations": n_ld, "points": n_sets * n_times,
                     "in_transit_fraction": f_in},
        "flops_convention": "A-table v1: 302 flop per executed Kepler solve + orbit, 746 per executed LD evaluation; "
                            "window pre-tests counted as 0. The A-table charges every flux evaluation the reference's "
                            "formula (10 node cosines, 2 atan2); for samples with the planet fully inside the disk the list "
                            "path evaluates the same function with constant kappas and node cosines, i.e. fewer executed "
                            "instructions than charged (hardware side: FP64 pipe 58 % active, profiles/r1_bench_tl_eval.md)",
        "hbm_peak_gbs": peaks.get("hbm_gbs"),
    }

    # ---- secondary numbers (rank 0, N = 1 only): dense mode, C2 Kepler ------------------------
    secondary = {}
    if world == 1 and not args.no_secondary:
        n_sec = 3
        timed_steps(2, flags=_lib.JXP_FLAG_NO_WINDOW, kernel_events=False)
        d_tot, d_k = timed_steps(n_sec, flags=_lib.JXP_FLAG_NO_WINDOW)
        _lib.call("jxp_system_stats", ws.data_ptr(), n_sets, 1, n_times, ctypes.addressof(stats), stream())
        dk, dl = int(stats[0]), int(stats[1])
        d_flops = dk * (W_KEPLER + W_ORBIT) + dl * W_LD
        secondary["c3_dense_no_window"] = {
            "points_per_s": n_sets * n_times / (d_tot / n_sec * 1e-3), "ms_per_step": d_tot / n_sec,
            "fp64_frac": d_flops / (d_k / n_sec * 1e-3) / 1e12 / fp64_peak,
            "executed": {"kepler_solves": dk, "ld_evaluations": dl}}
        n2 = 100_000_000
        M = torch.rand(n2, dtype=torch.float64, device=dev) * (2 * np.pi)
        e = torch.rand(n2, dtype=torch.float64, device=dev) * 0.99
        s_ = torch.empty_like(M)
        c_ = torch.empty_like(M)
        best = 1e30
        for i in range(4):
            ev0.record()
            _lib.call("jxp_kepler_f64", M.data_ptr(), 1, e.data_ptr(), 1, n2, s_.data_ptr(), c_.data_ptr(),
                      0, 0, 0, stream())
            ev1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, ev0.elapsed_time(ev1))
        secondary["c2_kepler_1e8"] = {
            "solves_per_s": n2 / (best * 1e-3), "ms": best,
            "fp64_frac": n2 * W_KEPLER / (best * 1e-3) / 1e12 / fp64_peak,
            "hbm_gbs": 32.0 * n2 / (best * 1e-3) / 1e9,
            "hbm_frac": (32.0 * n2 / (best * 1e-3) / 1e9) / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None}
        del M, e, s_, c_
        # K2: core.limb_dark.light_curve on 1e8 random in-transit (b, r) points
        rr = torch.rand(n2, dtype=torch.float64, device=dev) * 0.12 + 0.03
        bb = torch.rand(n2, dtype=torch.float64, device=dev) * (1 + rr)
        u2 = torch.tensor([0.3, 0.2], dtype=torch.float64, device=dev)
        f_ = torch.empty_like(bb)
        best = 1e30
        for i in range(4):
            ev0.record()
            _lib.call("jxp_limb_dark_f64", u2.data_ptr(), 2, bb.data_ptr(), 1, rr.data_ptr(), 1, n2, 10,
                      f_.data_ptr(), 0, 0, 0, stream())
            ev1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, ev0.elapsed_time(ev1))
        secondary["k2_limb_dark_1e8_in_transit"] = {
            "evaluations_per_s": n2 / (best * 1e-3), "ms": best,
            "fp64_frac": n2 * W_LD / (best * 1e-3) / 1e12 / fp64_peak}
        del rr, bb, f_
        secondary.update(secondary_c4_c5(dev, lib, ev0, ev1, peaks))

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        v, n_sample, threads, secs = cpu_points_per_s(bodies_h, u_h, t_h, y.cpu().numpy(), sigma)
        cpu = {"value": v, "unit": "points/s", "cores": threads, "kind": "port",
               "sample": f"{n_sample} of {n_sets} sets x {n_times} times, {secs:.1f} s",
               "note": "C/OpenMP restatement of the reference formulas (oracle/ref_c.c); not jax[cpu]"}

    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "path": "jxp_orbital_elements_f64 + jxp_system_light_curve_f64 (C ABI) from the sampled elements; inputs in one pinned host arena, one H2D copy per step",
                "python_api": {"value": points_per_step / (py_ms * 1e-3), "unit": "points/s", "ms_per_step": py_ms,
                               "path": "light_curves.limb_dark.log_likelihood(System(...).add_body(**sampled elements), u)"
                                       "(t, y, sigma): NumPy in, NumPy out, System rebuilt every step"}},
        "gpu_launches": len(launches) * args.steps,
        "gpu_launches_per_step": launches,
        "clocks": clocks, "secondary": secondary,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def secondary_c4_c5(dev, lib, ev0, ev1, peaks):
    """Configs 4 and 5 at full size (timed, best of 3, device-resident inputs)."""
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    out = {}

    def best_of(fn, n=3):
        best = 1e30
        for i in range(n + 1):
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, ev0.elapsed_time(ev1))
        return best

    # C1: the reference's own CPU-runnable case, 1 planet x 1e5 times, flux materialised
    from jaxoplanet_b200.orbits.keplerian import System

    body1, u1, t1 = W.c1()
    b1, _, _ = pack_bodies(System().add_body(**body1))
    b1 = b1.to(dev).contiguous()
    u1d = torch.as_tensor(np.asarray(u1, dtype=np.float64)).to(dev)
    t1d = torch.as_tensor(t1).to(dev)
    n1 = t1d.numel()
    wsb1 = int(lib.jxp_system_workspace_bytes(1, 1, n1))
    ws1 = torch.empty(wsb1, dtype=torch.uint8, device=dev)
    f1 = torch.empty((1, n1, 1), dtype=torch.float64, device=dev)
    ms = best_of(lambda: _lib.call("jxp_system_light_curve_f64", b1.data_ptr(), 1, 1, u1d.data_ptr(), 2, 0,
                                   t1d.data_ptr(), n1, 0.0, 0, 0, 10, 0, f1.data_ptr(), 0, 0, 0, 0, 0,
                                   ws1.data_ptr(), wsb1, A.stream_ptr()), n=5)
    out["c1_single_planet_1e5_times"] = {"points_per_s": n1 / (ms * 1e-3), "ms": ms,
                                         "in_transit_fraction": float((f1 != 0).double().mean())}

    # C4: 3 planets, 30-min cadence, 15x exposure supersampling, 1024 sets x 2e5 times
    p, u, t, expo = W.c4()
    bodies, n_sets, _ = pack_bodies(W.system_from(p))
    bodies = bodies.to(dev).contiguous()
    ud = torch.as_tensor(u).to(dev).contiguous()
    td = torch.as_tensor(t).to(dev)
    n_times = td.numel()
    wsb = int(lib.jxp_system_workspace_bytes(n_sets, 3, n_times))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    fs = torch.empty((n_sets, n_times), dtype=torch.float64, device=dev)
    ms = best_of(lambda: _lib.call("jxp_system_light_curve_f64", bodies.data_ptr(), n_sets, 3, ud.data_ptr(), 2, 2,
                                   td.data_ptr(), n_times, expo["exposure_time"], expo["order"], expo["num_samples"],
                                   10, 0, 0, fs.data_ptr(), 0, 0, 0, 0, ws.data_ptr(), wsb, A.stream_ptr()))
    stats = (ctypes.c_int64 * 2)()
    _lib.call("jxp_system_stats", ws.data_ptr(), n_sets, 3, n_times, ctypes.addressof(stats), A.stream_ptr())
    out["c4_3planets_15x_supersampled"] = {
        "points_per_s": n_sets * n_times / (ms * 1e-3), "ms": ms,
        "inner_evaluations_per_s": n_sets * n_times * 45 / (ms * 1e-3),
        "executed": {"kepler_solves": int(stats[0]), "ld_evaluations": int(stats[1])},
        "tflops_executed": (int(stats[0]) * (W_KEPLER + W_ORBIT) + int(stats[1]) * W_LD) / (ms * 1e-3) / 1e12,
        "in_transit_fraction": float((fs != 0).double().mean())}
    del fs, ws, bodies
    torch.cuda.empty_cache()

    # C5: flux + 8 partials, 16384 sets x 5e4 times, f64 and f32
    theta, t5 = W.c5()
    thd = torch.as_tensor(theta).to(dev).contiguous()
    t5d = torch.as_tensor(t5).to(dev)
    ns5, nt5 = thd.shape[0], t5d.numel()
    star = torch.tensor([1.0, 1.0], dtype=torch.float64, device=dev)
    wsb5 = int(lib.jxp_transit_workspace_bytes(ns5, nt5))
    ws5 = torch.empty(wsb5, dtype=torch.uint8, device=dev)
    for dt, nm in ((torch.float64, "f64"), (torch.float32, "f32")):
        fl = torch.empty((ns5, nt5), dtype=dt, device=dev)
        pt = torch.empty((ns5, 8, nt5), dtype=dt, device=dev)
        ms = best_of(lambda: _lib.call("jxp_transit_partials_" + nm, thd.data_ptr(), ns5, star.data_ptr(), 0,
                                       t5d.data_ptr(), nt5, 0.0, 0, 0, 10, 0, fl.data_ptr(), pt.data_ptr(), 0, 0, 0,
                                       0, 0, ws5.data_ptr(), wsb5, A.stream_ptr()))
        gbs = ns5 * nt5 * 9 * fl.element_size() / (ms * 1e-3) / 1e9
        out["c5_flux_plus_8_partials_" + nm] = {
            "points_per_s": ns5 * nt5 / (ms * 1e-3), "ms": ms, "store_gbs": gbs,
            "hbm_frac": gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None}
        del fl, pt
        torch.cuda.empty_cache()
    # C5 as a gradient step: log-likelihood and its 8-gradient reduced in the kernel, no rows stored
    y5 = torch.ones(nt5, dtype=torch.float64, device=dev)
    sg5 = torch.full((1,), 5e-4, dtype=torch.float64, device=dev)
    ll5 = torch.empty(ns5, dtype=torch.float64, device=dev)
    dll5 = torch.empty((ns5, 8), dtype=torch.float64, device=dev)
    ms = best_of(lambda: _lib.call("jxp_transit_partials_f64", thd.data_ptr(), ns5, star.data_ptr(), 0,
                                   t5d.data_ptr(), nt5, 0.0, 0, 0, 10, 0, 0, 0, y5.data_ptr(), sg5.data_ptr(), 0,
                                   ll5.data_ptr(), dll5.data_ptr(), ws5.data_ptr(), wsb5, A.stream_ptr()))
    ls5 = (ctypes.c_int64 * 4)()
    _lib.call("jxp_transit_list_stats", ws5.data_ptr(), ns5, nt5, ctypes.addressof(ls5), A.stream_ptr())
    out["c5_loglike_plus_8_gradient_f64"] = {
        "points_per_s": ns5 * nt5 / (ms * 1e-3), "ms": ms,
        "path": "transit list (k_tl_build, k_tlg_eval, k_tlg_reduce)" if (ls5[0] > 0 and ls5[1] == 0)
                else "window-test kernel (k_transit_partials)",
        "list_items": int(ls5[0])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (WORLD_SIZE={world}); running 1 GPU", file=sys.stderr)
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
