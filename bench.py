#!/usr/bin/env python
"""Benchmark of the jaxoplanet hot path on B200 (contract: the task statement / DESIGN.md 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json config 3 AS WRITTEN -- 4096 parameter sets x 1e5 time samples of
a Keplerian transit light curve with quadratic limb darkening, fused with the Gaussian log-likelihood,
the 4096 sets SHARDED over the N GPUs (strong scaling: rank g evaluates sets [g 4096/N, (g+1) 4096/N);
t, y, sigma are replicated; the per-set log-likelihoods are all-gathered over NCCL inside every step).
A "step" is one evaluation of the batch, the unit a sampler performs per proposal.
metric = light-curve points/s (one point = one time sample of one parameter set).

value      device-resident inputs, CUDA events around each step, L2 flushed between steps, max over ranks
weak       (N > 1) the same step with 4096 sets PER GPU (round 1's definition), as a second field
e2e        the same step through the repo's public Python API (the reference's call: System(...).add_body(
           **sampled elements) -> log_likelihood(orbit, u)(t, y, sigma)) with HOST NumPy arrays in and out:
           host-to-device copies of every input and the device-to-host read of the result inside the timed
           region; e2e.c_abi: the same through the bare C ABI from one pinned arena
roofline   dominant kernel timed alone with CUDA events inside the timed steps: k_walk_items, which answers every
           in-window sample from the per-set function tables (csrc/jxp_tables.cuh: two Horner chains per sample
           instead of a Kepler solve + the ten-node quadrature).  frac = FP64 flops the launch EXECUTED (thread
           instructions counted by ncu on the same seeded inputs, profiles/r2_walk_items.json: 2 per DFMA, 1 per
           DMUL / DADD) / its live duration / the FP64 FMA peak measured in this run; pipe_fp64_pct from the same
           capture.  roofline.algorithmic: the SURVEY 8d figure -- the flops of the REFERENCE formula (302 per Kepler
           solve + orbit, 746 per flux evaluation) for the samples the launch answered / its duration; above the
           peak because the tables replace most of those flops, not because the pipe runs faster.
cpu_baseline / --impl reference: the reference's algorithm on the host cores -- real jaxoplanet on jax[cpu]
           when `import jax` and a reference checkout (baseline/_ref) exist, else the C/OpenMP restatement
           (oracle/ref_c.c, -O3 -march=native) -- over ALL 4096 sets per step.
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
# a flux evaluation with the planet fully inside the disk (kappa0 = pi, kappa1 = 0, kite area 0: constants):
# the reference formula without its 10 node cosines (250), 2 atan2 (80) and the kite area (22)
W_LD_INSIDE = 394.0
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
                    f"fused Gaussian log-likelihood; the {N_SETS} sets sharded over {n_gpus} GPU(s)",
        "n_sets": N_SETS,
        "sets_per_gpu": N_SETS // max(n_gpus, 1),
        "n_times": N_TIMES,
        "n_bodies": 1,
        "parallelism": f"set-sharded x{n_gpus} (no data-path collective; all-gather of 8 B per set per step: see exchange)",
        "cache": "L2 flushed between steps (256 MiB write); per-step CUDA events summed, the steps enqueued back to back",
    }
    if extra:
        cfg.update(extra)
    return cfg


def c3_inputs(seed=1):
    from jaxoplanet_b200 import workloads as W

    return W.c3(N_SETS, N_TIMES, seed=seed)


# ------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm on the host cores
# ------------------------------------------------------------------------------------------
def oracle_records(p):
    """Oracle-side body records (host arithmetic of oracle/ref_numpy.py) for the C restatement."""
    from oracle import ref_numpy as ref

    bd = ref.orbital_body(**p)
    n = len(p["period"])
    rec = np.zeros((n, 1, 12))
    rec[:, 0, 0] = bd["period"]
    rec[:, 0, 1] = bd["time_transit"]
    rec[:, 0, 2] = bd["time_ref"]
    rec[:, 0, 3] = bd["eccentricity"]
    rec[:, 0, 4] = bd["cos_omega_peri"]
    rec[:, 0, 5] = bd["sin_omega_peri"]
    rec[:, 0, 6] = bd["cos_inclination"]
    rec[:, 0, 7] = bd["sin_inclination"]
    rec[:, 0, 8] = bd["semimajor"]
    rec[:, 0, 9] = 1.0
    rec[:, 0, 10] = bd["radius"]
    return rec


def reference_y(rec, u, t, noise):
    """y = 1 + model of set 0 + noise, from the CPU restatement (the GPU arm derives the same y from its own
    set-0 light curve; the two agree to 1e-15, irrelevant for timing)."""
    from oracle import ref_c

    f0, _ = ref_c.system(rec[:1], u[:1], t)
    return 1.0 + f0[0] + noise


def jax_reference():
    """Real jaxoplanet on jax[cpu], if this box has it: `import jax` plus the reference package importable
    (site-packages or a checkout under baseline/_ref).  Returns (callable (p, u, t, y, sigma) -> (n_sets,)
    log-likelihoods, description) or (None, reason)."""
    try:
        import jax  # noqa: F401
    except Exception as e:  # noqa: BLE001
        return None, f"jax not importable ({type(e).__name__})"
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.path.join(ROOT, "baseline", "_ref", "src")):
        if os.path.isdir(cand) and cand not in sys.path:
            sys.path.insert(0, cand)
    try:
        import jax
        import jax.numpy as jnp

        jax.config.update("jax_enable_x64", True)
        jax.config.update("jax_platform_name", "cpu")
        from jaxoplanet.light_curves import limb_dark_light_curve
        from jaxoplanet.orbits.keplerian import Central, System
    except Exception as e:  # noqa: BLE001
        return None, f"jaxoplanet not importable ({type(e).__name__}: {e})"

    def one(params, uu, t, y, sigma):
        orbit = System(Central()).add_body(**params)
        lc = limb_dark_light_curve(orbit, uu)(t)[..., 0]
        r = (y - (1.0 + lc)) / sigma
        return jnp.sum(-0.5 * r * r - jnp.log(sigma) - 0.5 * jnp.log(2 * jnp.pi))

    batched = jax.jit(jax.vmap(one, in_axes=(0, 0, None, None, None)))

    def run(p, u, t, y, sigma):
        return np.asarray(batched({k: jnp.asarray(v) for k, v in p.items()}, jnp.asarray(u), jnp.asarray(t),
                                  jnp.asarray(y), sigma).block_until_ready())

    return run, "jaxoplanet on jax[cpu], x64, jit + vmap over the parameter sets"


def run_reference(args, rank, world):
    if rank != 0:
        return
    p, u, t, noise, sigma = c3_inputs()
    from oracle import ref_c

    ref_c.build()
    threads = ref_c.use_all_cores()
    rec = oracle_records(p)
    y = reference_y(rec, u, t, noise)
    jax_run, how = jax_reference()
    kind = "reference" if jax_run is not None else "port"

    def step(n):
        if jax_run is not None:
            return jax_run({k: v[:n] for k, v in p.items()}, u[:n], t, y, sigma)
        return ref_c.system(rec[:n], u[:n], t, y=y, sigma=sigma, want_flux=False)[1]

    # every step evaluates ALL 4096 sets unless that would take the whole run past ~4 minutes
    n_try = max(2 * threads, 32)
    step(n_try)
    t0 = time.perf_counter()
    step(n_try)
    per_set = (time.perf_counter() - t0) / n_try
    total_steps = args.steps + args.warmup
    n_sample = N_SETS
    if per_set * N_SETS * total_steps > 240.0:
        n_sample = int(240.0 / (per_set * total_steps))
        n_sample = max(threads, min(N_SETS, (n_sample // threads) * threads))
    for _ in range(args.warmup):
        step(n_sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(n_sample)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_sample * N_TIMES / dt
    sample = (f"all {N_SETS} sets x {N_TIMES} times per step" if n_sample == N_SETS else
              f"{n_sample} of {N_SETS} sets x {N_TIMES} times per step (bounded to keep the run within minutes; "
              "the work is linear in sets)")
    note = (how if jax_run is not None else
            "C/OpenMP restatement of the reference formulas (oracle/ref_c.c, -O3 -march=native, no fast-math), every "
            f"sample evaluated in full as the reference does; not jax[cpu]: {how}")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 * (N_SETS / n_sample), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, {"cache": "n/a (CPU)"}),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": kind, "sample": sample,
                         "note": note},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline_leg(p, u, t, y, sigma):
    """N = 1 only: the CPU arm on a bounded sample (one pass over all sets, ~10-30 s of CPU work), plus the
    Kepler solver alone (config 2) next to the reference's published 14 ns/solve."""
    from oracle import ref_c

    ref_c.build()
    threads = ref_c.use_all_cores()
    rec = oracle_records(p)
    jax_run, how = jax_reference()
    n_try = max(2 * threads, 32)
    if jax_run is not None:
        run = lambda n: jax_run({k: v[:n] for k, v in p.items()}, u[:n], t, y, sigma)  # noqa: E731
    else:
        run = lambda n: ref_c.system(rec[:n], u[:n], t, y=y, sigma=sigma, want_flux=False)  # noqa: E731
    run(n_try)
    t0 = time.perf_counter()
    run(n_try)
    per_set = (time.perf_counter() - t0) / n_try
    n_sample = N_SETS if per_set * N_SETS <= 30.0 else max(threads, int(30.0 / per_set) // threads * threads)
    t0 = time.perf_counter()
    run(n_sample)
    secs = time.perf_counter() - t0
    out = {"value": n_sample * N_TIMES / secs, "unit": "points/s", "cores": threads,
           "kind": "reference" if jax_run is not None else "port",
           "sample": f"{n_sample} of {N_SETS} sets x {N_TIMES} times, one pass, {secs:.1f} s",
           "note": how if jax_run is not None else
           f"C/OpenMP restatement of the reference formulas (oracle/ref_c.c, -O3 -march=native); not jax[cpu]: {how}"}
    # C2 on the CPU: the Kepler solve alone
    rng = np.random.default_rng(0)
    n2 = 20_000_000
    M = rng.uniform(0.0, 2 * np.pi, n2)
    e = rng.uniform(0.0, 0.99, n2)
    ref_c.kepler(M[:100_000], e[:100_000])
    t0 = time.perf_counter()
    ref_c.kepler(M, e)
    k_secs = time.perf_counter() - t0
    out["c2_kepler"] = {"solves_per_s": n2 / k_secs, "ns_per_solve_per_core": k_secs / n2 * 1e9 * threads,
                        "sample": f"{n2} (M, e) pairs, e ~ U[0, 0.99), {threads} threads",
                        "published_reference": "14 ns/solve on one M1 core at fixed e = 0.5 "
                                               "(docs/tutorials/core-from-scratch.ipynb:215)"}
    return out


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


class NvmlSampler:
    """Clocks and throttle reasons of this rank's GPU straight from NVML (nvidia_ml_py), sampled by the MAIN thread
    while the GPU executes the timed steps and the host has nothing left to enqueue (timed_steps calls sample()
    between its last launch and its wait).  The nvidia-smi loop of ClockSampler -- a second process polling the
    driver every 50 ms -- showed up in the measurement at N > 1: its queries on rank 0 delay that rank's launches
    and the exchange makes every rank wait for them (0.142 against 0.133 ms per step at 4 GPUs)."""

    def __init__(self, torch_index):
        import pynvml
        import torch

        self.nv = pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:  # noqa: BLE001
            self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_index)
        self.samples = []

    def sample(self, n=4, gap_s=3e-4):
        nv = self.nv
        for _ in range(n):
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:  # noqa: BLE001
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.samples.append((sm, mx, int(rs)))
            time.sleep(gap_s)

    def result(self):
        nv = self.nv
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML sample"]}
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r & bit for _, _, r in self.samples))
        return {"sm_mhz": statistics.median([x[0] for x in self.samples]), "sm_max_mhz": max(x[1] for x in self.samples),
                "samples": len(self.samples), "reasons": reasons,
                "how": "NVML from the main thread while the GPU executes the timed steps (after the last launch, "
                       "before the wait)"}


# ------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------
class Problem:
    """Device-resident inputs + workspace of one rank's block of sets; call() = one step through the C ABI."""

    def __init__(self, p, u, t, dev):
        import torch

        from jaxoplanet_b200 import _array as A
        from jaxoplanet_b200 import _lib
        from jaxoplanet_b200.light_curves.limb_dark import pack_elements

        self.A, self._lib, self.lib = A, _lib, _lib.load()
        self.n_sets, self.n_times = len(p["period"]), len(t)
        self.elements_h = np.ascontiguousarray(pack_elements(self.n_sets, **p), dtype=np.float64)
        el = torch.as_tensor(self.elements_h).to(dev)
        self.bodies = torch.empty((self.n_sets, 1, _lib.JXP_BODY_NFIELDS), dtype=torch.float64, device=dev)
        _lib.call("jxp_orbital_elements_f64", el.data_ptr(), self.n_sets, self.bodies.data_ptr(), A.stream_ptr())
        self.u = torch.as_tensor(np.ascontiguousarray(u)).to(dev)
        self.t = torch.as_tensor(t).to(dev)
        self.wsb = int(self.lib.jxp_system_workspace_bytes(self.n_sets, 1, self.n_times))
        self.ws = torch.empty(self.wsb, dtype=torch.uint8, device=dev)
        self.ll = torch.empty(self.n_sets, dtype=torch.float64, device=dev)

    def call(self, y=None, sig=None, flags=0, flux_sum=None):
        A = self.A
        self._lib.call("jxp_system_light_curve_f64", self.bodies.data_ptr(), self.n_sets, 1, self.u.data_ptr(), 2, 2,
                       self.t.data_ptr(), self.n_times, 0.0, 0, 0, 10, flags, 0, A.ptr(flux_sum), A.ptr(y),
                       A.ptr(sig) if y is not None else 0, 0, self.ll.data_ptr() if y is not None else 0,
                       self.ws.data_ptr(), self.wsb, A.stream_ptr())

    def stats(self):
        st = (ctypes.c_int64 * 2)()
        self._lib.call("jxp_system_stats", self.ws.data_ptr(), self.n_sets, 1, self.n_times, ctypes.addressof(st),
                       self.A.stream_ptr())
        ls = (ctypes.c_int64 * 4)()
        self._lib.call("jxp_system_list_stats", self.ws.data_ptr(), self.n_sets, 1, self.n_times, ctypes.addressof(ls),
                       self.A.stream_ptr())
        wk = (ctypes.c_int64 * 4)()
        self._lib.call("jxp_system_walk_stats", self.ws.data_ptr(), self.n_sets, 1, self.n_times, ctypes.addressof(wk),
                       self.A.stream_ptr())
        return dict(kepler=int(st[0]), ld=int(st[1]), items=int(ls[0]), flags=int(ls[1]), inside=int(ls[2]),
                    fast=int(ls[3]), tab_evals=int(wk[0]), n_fix=int(wk[1]), tab_sets=int(wk[2]), no_tables=int(wk[3]))


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.distributed import log_likelihood_distributed, peer_gather_for, shard_range

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    stream = A.stream_ptr

    p, u_h, t_h, noise, sigma = c3_inputs(seed=1)
    lo, hi = shard_range(N_SETS, rank, world)
    p_loc = {k: v[lo:hi] for k, v in p.items()}
    prob = Problem(p_loc, u_h[lo:hi], t_h, dev)
    n_times = prob.n_times
    sig = torch.full((1,), sigma, dtype=torch.float64, device=dev)

    # observed flux: 1 + model of set 0 + noise (synthetic); every rank evaluates set 0 itself, untimed
    p0 = Problem({k: v[:1] for k, v in p.items()}, u_h[:1], t_h, dev)
    f0 = torch.empty((1, n_times), dtype=torch.float64, device=dev)
    p0.call(flux_sum=f0)
    y = (1.0 + f0[0] + torch.as_tensor(noise).to(dev)).contiguous()
    y_np = y.cpu().numpy()
    del p0, f0
    torch.cuda.synchronize()

    ll_all = torch.empty(N_SETS, dtype=torch.float64, device=dev) if world > 1 else None
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush_buf.fill_(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    ev0, ev1, kev0, kev1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    for e in (ev0, ev1, kev0, kev1):
        e.record()
    torch.cuda.synchronize()

    # the exchange of a step: every rank's block of log-likelihoods to every rank -- peer stores over NVLink
    # (jxp_peer_allgather_f64, one launch on the step's stream), NCCL's all-gather if the buffers cannot be shared
    def make_step(problem, gathered, pg=None, offset=0):
        def step(flags=0):
            problem.call(y=y, sig=sig, flags=flags)
            if world > 1:
                if pg is not None:
                    pg.gather(problem.ll, offset)
                else:
                    dist.all_gather_into_tensor(gathered, problem.ll)

        return step

    peer = peer_gather_for(N_SETS, dev) if world > 1 else None

    step_events, last_steps = [], []

    def timed_steps(step, n, flags=0, kernel_events=True, busy_hook=None):
        """Sum of the per-step device times (CUDA events around every step, the L2 flush between steps outside
        them).  Without the kernel events the n steps are enqueued back to back and the host waits once at the
        end: at N > 1 the ranks then stay in step through the exchange itself instead of drifting apart by the
        host's per-step jitter.  With them (one event pair inside the C-ABI call) every step is awaited."""
        tot, ktot = 0.0, 0.0
        if not kernel_events:
            while len(step_events) < 2 * n:
                step_events.append(torch.cuda.Event(enable_timing=True))
            if world > 1:
                # one more untimed step in front, enqueued with the timed ones: the ranks leave the host barrier up to
                # ~0.1 ms apart, and without it the first timed step of the early ranks is a wait for the late ones
                # (0.27 ms against 0.133 at 4 GPUs); the exchange of this step lines the GPUs up
                flush_l2()
                step(flags)
            for i in range(n):
                flush_l2()
                step_events[2 * i].record()
                step(flags)
                step_events[2 * i + 1].record()
            if busy_hook is not None:
                busy_hook()  # (the GPU is executing the steps: clocks under load)
            torch.cuda.synchronize()
            last_steps[:] = [step_events[2 * i].elapsed_time(step_events[2 * i + 1]) for i in range(n)]
            return sum(last_steps), 0.0
        for _ in range(n):
            flush_l2()
            lib.jxp_profile_events(ctypes.c_void_p(kev0.cuda_event), ctypes.c_void_p(kev1.cuda_event))
            ev0.record()
            step(flags)
            ev1.record()
            lib.jxp_profile_events(None, None)
            torch.cuda.synchronize()
            tot += ev0.elapsed_time(ev1)
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

    # ---- value: device-resident inputs, the 4096 sets sharded over the ranks ---------------------
    step = make_step(prob, ll_all, peer, lo)

    # The timed step is the C-ABI call (and, at N > 1, the peer-store exchange behind it) captured ONCE into a CUDA
    # graph and replayed: the launches of a step are fixed -- everything that depends on the data is decided on the
    # device --, so a step is a launch-bound chain of seven small kernels and two streams, which is what graphs are for
    # (an XLA program would run the custom call under a command buffer the same way).  The replay writes the same bits
    # as the stream-ordered call (tests/test_cuda_graph.py; checked below).  Stream-ordered launches are timed beside
    # it (`ms_per_step_stream_launches`).  Not with the NCCL fallback of the exchange, nor with --no-graph.
    def capture(step_fn, pg):
        from jaxoplanet_b200.graphs import capture as capture_call

        g = capture_call(step_fn, after_capture=pg.sync_epoch if pg is not None else None)
        if g is None:
            print("bench.py: CUDA graph capture failed; stream-ordered launches", file=sys.stderr)
        return g

    graph = None
    if not args.no_graph and (world == 1 or peer is not None):
        step()
        torch.cuda.synchronize()
        ll_eager = prob.ll.clone()
        graph = capture(step, peer)
        if graph is not None:
            prob.ll.zero_()
            graph.replay()
            torch.cuda.synchronize()
            if not torch.equal(prob.ll, ll_eager):
                raise AssertionError("graph replay and stream-ordered call disagree")
    flag = torch.tensor([1 if graph is not None else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # (all ranks or none: the exchange must be launched by everyone)
    use_graph = bool(int(flag.item()))
    step_eager = step
    if use_graph:
        def step(flags=0):  # noqa: F811
            graph.replay()
    warm = max(args.warmup, 3)
    timed_steps(step, warm, kernel_events=False)
    # which kernel dominates: with the per-set function tables (the default) k_walk_items, else k_walk_eval; the
    # profile events bracket that launch (a development option that moves the event records, nothing else)
    st_w = prob.stats()
    tables_used = st_w["tab_evals"] > 0.5 * max(st_w["ld"], 1)
    sampler, nvml = None, None
    if rank == 0:
        try:
            nvml = NvmlSampler(local_rank)
        except Exception:  # noqa: BLE001 -- no NVML bindings: the nvidia-smi loop
            sampler = ClockSampler(local_rank)
            sampler.start()
            time.sleep(0.25)
    barrier()
    t_lo = time.perf_counter()
    # (no event records inside the call during the timed steps -- the pair around the dominant launch costs ~14 us of
    # a 0.28 ms step --: the dominant kernel's own duration is measured in separate steps right after)
    tot_ms, _ = timed_steps(step, args.steps, kernel_events=False, busy_hook=nvml.sample if nvml else None)
    barrier()
    t_hi = time.perf_counter()
    per_step_ms = [round(x, 4) for x in last_steps]
    clocks = (nvml.result() if nvml else sampler.stop(t_lo, t_hi)) if rank == 0 else None
    st = prob.stats()
    walk_used = st["items"] > 0 and st["flags"] == 0
    eager_ms = None
    if use_graph:
        if peer is not None:
            peer.sync_epoch()
        step = step_eager
        timed_steps(step, 3, kernel_events=False)
        barrier()
        e_ms, _ = timed_steps(step, args.steps, kernel_events=False)
        barrier()
        eager_ms = max_over_ranks(e_ms) / args.steps
    lib.jxp_set_dev_options(_lib.JXP_DEV_PROFILE_ITEMS if tables_used else 0)
    n_k = min(args.steps, 10)
    _, k_ms = timed_steps(step, n_k)
    k_ms *= args.steps / n_k  # (k_ms_avg below divides by args.steps)
    # plan launch alone (same events, other bracket; the table launch runs beside it on a side stream), and the
    # table launch alone (serial in that measurement)
    lib.jxp_set_dev_options(_lib.JXP_DEV_PROFILE_PLAN)
    _, plan_ms = timed_steps(step, 5)
    lib.jxp_set_dev_options(_lib.JXP_DEV_PROFILE_TABLES)
    _, tables_ms = timed_steps(step, 5)
    lib.jxp_set_dev_options(0)
    plan_ms /= 5
    tables_ms /= 5
    # sorted time stamps: only the in-window samples are visited (k_walk_*); the window-test kernels are still
    # launched behind them and return at once (the choice is made on the device)
    if walk_used and tables_used:
        launches = ["k_walk_tables (side stream, beside the plan)", "k_walk_plan", "k_walk_eval (sets without tables)",
                    "k_walk_items", "k_walk_fix", "k_system_pair (returns at once)", "k_loglike_final (returns at once)"]
    elif walk_used:
        launches = ["k_walk_plan", "k_walk_eval", "k_system_pair (returns at once)", "k_loglike_final (returns at once)"]
    else:
        launches = ["k_sys_prep_ll", "k_system_pair", "k_loglike_final"]
    tot_ms = max_over_ranks(tot_ms)
    ms_per_step = tot_ms / args.steps
    points_per_step = N_SETS * n_times
    value = points_per_step / (ms_per_step * 1e-3)

    # ---- the exchange: what was used, checked against NCCL's all-gather, and the step with NCCL beside it -----
    exchange = None
    if world > 1:
        step_nccl = make_step(prob, ll_all)
        step_nccl()
        torch.cuda.synchronize()
        exchange = {"kind": "ncclAllGather (torch.distributed)"}
        if peer is not None:
            got = peer.gather(prob.ll, lo).clone()
            torch.cuda.synchronize()
            # the step with either exchange, measured alternately under the same conditions
            ab = {"peer": [], "nccl": []}
            for _ in range(2):
                for which, fn in (("peer", step), ("nccl", step_nccl)):
                    timed_steps(fn, 3, kernel_events=False)
                    barrier()
                    x_ms, _ = timed_steps(fn, args.steps, kernel_events=False)
                    barrier()
                    ab[which].append(max_over_ranks(x_ms) / args.steps)
            exchange = {"kind": "peer stores over NVLink into CUDA-IPC exchange buffers + epoch flags, one launch per step "
                                "(jxp_peer_allgather_f64)",
                        "identical_to_nccl_all_gather": bool(torch.equal(got, ll_all)), "status": peer.status(),
                        "ms_per_step_with_peer_stores": min(ab["peer"]),
                        "ms_per_step_with_nccl_all_gather": min(ab["nccl"]),
                        "how": "the timed step with either exchange, 2 x K steps each, alternating, best of the two; "
                               "measured right after the headline"}

    # ---- weak scaling as a second field (N > 1): 4096 sets per GPU, round 1's definition -----------
    weak = None
    if world > 1:
        pw, uw, _, _, _ = c3_inputs(seed=1 + rank)
        probw = Problem(pw, uw, t_h, dev)
        llw_all = torch.empty(world * N_SETS, dtype=torch.float64, device=dev)
        pgw = peer_gather_for(world * N_SETS, dev)
        stepw = make_step(probw, llw_all, pgw, rank * N_SETS)
        if use_graph and pgw is not None:  # (the same launch mode as the headline)
            stepw()
            gw = capture(stepw, pgw)
            fw = torch.tensor([1 if gw is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(fw, op=dist.ReduceOp.MIN)
            if int(fw.item()):
                def stepw(flags=0):  # noqa: F811
                    gw.replay()
        timed_steps(stepw, 3, kernel_events=False)
        barrier()
        w_ms, _ = timed_steps(stepw, args.steps, kernel_events=False)
        barrier()
        w_ms = max_over_ranks(w_ms) / args.steps
        weak = {"value": world * N_SETS * n_times / (w_ms * 1e-3), "unit": "points/s", "ms_per_step": w_ms,
                "sets_per_gpu": N_SETS, "scaling": "weak"}
        del probw, llw_all
        torch.cuda.empty_cache()

    # ---- optional flux all-gather (north_star): the materialised light curves, 3.28 GB in all, on every rank ----
    flux_gather = None
    if world > 1 and not args.no_secondary:
        n_loc_ = hi - lo
        f_loc = torch.empty((n_loc_, n_times), dtype=torch.float64, device=dev)
        f_all = torch.empty((N_SETS, n_times), dtype=torch.float64, device=dev)
        prob.call(flux_sum=f_loc)  # this rank's rows (transit walk, flux mode)
        best = 1e30
        for i in range(4):
            barrier()
            ev0.record()
            dist.all_gather_into_tensor(f_all, f_loc)
            ev1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, ev0.elapsed_time(ev1))
        best = max_over_ranks(best)
        recv = (N_SETS - n_loc_) * n_times * 8
        ok = bool(torch.equal(f_all[lo:hi], f_loc))
        flux_gather = {"ms": best, "bytes_total": N_SETS * n_times * 8, "bytes_received_per_rank": recv,
                       "gbs_received_per_rank": recv / (best * 1e-3) / 1e9, "own_rows_intact": ok,
                       "how": "ncclAllGather (torch.distributed.all_gather_into_tensor) of the per-rank flux rows over "
                              "NVLink, best of 3, max over ranks; not part of the timed step"}
        del f_loc, f_all
        torch.cuda.empty_cache()

    # ---- e2e through the public Python API (headline): NumPy in, NumPy out --------------------------
    from jaxoplanet_b200.light_curves.limb_dark import log_likelihood as facade_ll

    if world == 1:
        def api_step():
            return facade_ll(W.system_from(p), u_h)(t_h, y_np, sigma)
    else:
        def api_step():
            return log_likelihood_distributed(p, u_h, t_h, y_np, sigma).cpu().numpy()

    for _ in range(5):
        api_step()
    barrier()
    py_ms = 0.0
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ll_np = api_step()
        py_ms += (time.perf_counter() - t0) * 1e3
    assert isinstance(ll_np, np.ndarray) and ll_np.shape == (N_SETS,)
    barrier()
    py_ms = max_over_ranks(py_ms) / args.steps
    # the bytes the API moved per step and rank: element records, u, t, y, sigma in; the log-likelihoods out
    n_loc = hi - lo
    h2d = (n_loc * _lib.JXP_EL_NFIELDS + n_loc * 2 + 2 * n_times + 1) * 8
    d2h = N_SETS * 8

    # ---- e2e through the bare C ABI: one pinned arena [elements | u | t | y | sigma], one H2D copy -----
    parts = [prob.elements_h.reshape(-1), np.ascontiguousarray(u_h[lo:hi], dtype=np.float64).reshape(-1),
             np.ascontiguousarray(t_h, dtype=np.float64), y_np.astype(np.float64), np.array([sigma], dtype=np.float64)]
    offs = np.concatenate([[0], np.cumsum([((x.size + 31) // 32) * 32 for x in parts])])  # 256-byte aligned
    harena = torch.zeros(int(offs[-1]), dtype=torch.float64).pin_memory()
    for x, o in zip(parts, offs[:-1]):
        harena[int(o):int(o) + x.size] = torch.as_tensor(x)
    darena = torch.empty_like(harena, device=dev)
    del_, du, dt_, dy, ds = (darena[int(o):int(o) + x.size] for x, o in zip(parts, offs[:-1]))
    hll = torch.empty(N_SETS, dtype=torch.float64).pin_memory()
    db = torch.empty((n_loc, 1, _lib.JXP_BODY_NFIELDS), dtype=torch.float64, device=dev)

    def c_abi_step():
        darena.copy_(harena, non_blocking=True)
        _lib.call("jxp_orbital_elements_f64", del_.data_ptr(), n_loc, db.data_ptr(), stream())
        _lib.call("jxp_system_light_curve_f64", db.data_ptr(), n_loc, 1, du.data_ptr(), 2, 2,
                  dt_.data_ptr(), n_times, 0.0, 0, 0, 10, 0, 0, 0, dy.data_ptr(), ds.data_ptr(), 0,
                  prob.ll.data_ptr(), prob.ws.data_ptr(), prob.wsb, stream())
        if world > 1:
            if peer is not None:
                hll.copy_(peer.gather(prob.ll, lo), non_blocking=True)
            else:
                dist.all_gather_into_tensor(ll_all, prob.ll)
                hll.copy_(ll_all, non_blocking=True)
        else:
            hll.copy_(prob.ll, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        c_abi_step()
    barrier()
    c_ms = 0.0
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c_abi_step()
        c_ms += (time.perf_counter() - t0) * 1e3
    barrier()
    c_ms = max_over_ranks(c_ms) / args.steps
    assert np.array_equal(hll.numpy(), ll_np), "the Python API and the C ABI must return the same numbers"

    # ---- secondary configs (C2 / C4 / C5), sharded over the ranks like the headline -----------------
    secondary = {}
    if not args.no_secondary:
        secondary = secondary_configs(dev, lib, ev0, ev1, rank, world, fp64_peak, max_over_ranks, barrier)
        if world == 1:
            timed_steps(step, 2, flags=_lib.JXP_FLAG_NO_WINDOW, kernel_events=False)
            d_tot, d_k = timed_steps(step, 3, flags=_lib.JXP_FLAG_NO_WINDOW)
            sd = prob.stats()
            d_flops = sd["kepler"] * (W_KEPLER + W_ORBIT) + sd["ld"] * W_LD
            secondary["c3_dense_no_window"] = {
                "points_per_s": N_SETS * n_times / (d_tot / 3 * 1e-3), "ms_per_step": d_tot / 3,
                "fp64_frac": d_flops / (d_k / 3 * 1e-3) / 1e12 / fp64_peak,
                "executed": {"kepler_solves": sd["kepler"], "ld_evaluations": sd["ld"]}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0's shard) -------------------------------------------
    k_ms_avg = k_ms / args.steps
    n_gen = st["ld"] - st["fast"]
    flops_ref = st["kepler"] * (W_KEPLER + W_ORBIT) + st["ld"] * W_LD  # the reference formula on the samples visited
    flops_direct = st["kepler"] * (W_KEPLER + W_ORBIT) + st["fast"] * W_LD_INSIDE + n_gen * W_LD
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    kernel = ("k_walk_items" if tables_used else "k_walk_eval<2>") if walk_used else "k_system_pair<1>"
    cap_file = "r2_walk_items.json" if tables_used else "r2_walk_eval.json"
    hw = None
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", cap_file)))
        # instruction counts of a launch depend on the inputs only (same seeded workload); scale by the items
        scale = st["items"] / cap["items"]
        hw_flops = (2 * cap["dfma"] + cap["dmul"] + cap["dadd"]) * scale
        hw = {"fp64_thread_flops_per_launch": hw_flops, "tflops": hw_flops / (k_ms_avg * 1e-3) / 1e12,
              "frac_of_peak": hw_flops / (k_ms_avg * 1e-3) / 1e12 / fp64_peak,
              "pipe_fp64_pct_ncu": cap["pipe_fp64_pct"], "issue_active_pct_ncu": cap["issue_active_pct"],
              "warp_instructions_per_sample": cap["warp_inst"] * 32.0 / cap["items"] if cap.get("warp_inst") else None,
              "source": cap["source"]}
        traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * scale
    except Exception:
        pass
    if tables_used:
        # executed work: what the pipe did (ncu count); the A-table cannot describe a launch that replaces the
        # reference formula by table look-ups, so it is reported beside it as `algorithmic`
        achieved = hw["tflops"] if hw else None
        frac = hw["frac_of_peak"] if hw else None
    else:
        achieved = flops_direct / (k_ms_avg * 1e-3) / 1e12
        frac = achieved / fp64_peak
    roofline = {
        "bound": "fp64", "kernel": kernel, "achieved": achieved,
        "peak": fp64_peak, "unit": "TFLOP/s", "frac": frac, "traffic": traffic,
        "traffic_note": "DRAM bytes per launch (ncu, profiles/%s), scaled to this rank's items; algorithmic bytes: 24 B "
                        "per in-window sample (t, (2 r0 / sigma, 1 / sigma^2)) + 3.6 KB of tables per set + 24 B per "
                        "transit -- read through L2, which holds all of it" % cap_file,
        "hardware": hw,
        "algorithmic": {"tflops": flops_ref / (k_ms_avg * 1e-3) / 1e12,
                        "ratio_to_peak": flops_ref / (k_ms_avg * 1e-3) / 1e12 / fp64_peak,
                        "note": "SURVEY 8d per-unit figure (302 flop per Kepler solve + orbit, 746 per flux evaluation: the "
                                "reference formula) x the samples this launch answered / its duration.  Not a pipe "
                                "utilisation: the tables answer a sample with ~50 FP64 instructions"},
        "items": st["items"],
        "peak_source": "DFMA loop (jxp_peak_fma_f64) measured in this run; MEASURED_PEAKS.json has no FP64 figure",
        "kernel_ms": k_ms_avg, "plan_kernel_ms": plan_ms, "tables_kernel_ms": tables_ms,
        "kernel_ms_note": "CUDA events around the launch inside the C-ABI call (jxp_profile_events), on the same inputs with "
                          "the same L2 flush, in steps run right after the timed ones (the event pair inside the call costs "
                          "~14 us per step)",
        "kernel_share_of_step": k_ms_avg / ms_per_step,
        "executed": {"samples_visited": st["kepler"], "flux_evaluations": st["ld"], "from_tables": st["tab_evals"],
                     "direct_code_fix_list": st["n_fix"], "sets_with_tables": st["tab_sets"],
                     "transiting_sets_without_tables": st["no_tables"], "ld_inside_disk_code": st["fast"],
                     "points": (hi - lo) * n_times},
        "flops_convention": "frac: executed FP64 thread instructions of the launch (2 per DFMA, 1 per DMUL / DADD; ncu on "
                            "the same seeded inputs) over the live duration, against the measured DFMA peak.  Without "
                            "tables (JXP_DEV_NO_TABLES) the launch is k_walk_eval and frac is the A-table of the work it "
                            "executed (302 per Kepler solve + orbit, 746 / 394 per flux evaluation outside / inside the disk)",
        "hbm_peak_gbs": peaks.get("hbm_gbs"),
    }

    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline_leg(p, u_h, t_h, y_np, sigma)

    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": points_per_step / (py_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": py_ms,
                "path": ("light_curves.limb_dark.log_likelihood(System(Central()).add_body(**sampled elements), u)"
                         "(t, y, sigma)" if world == 1 else
                         "jaxoplanet_b200.distributed.log_likelihood_distributed(sampled elements, u, t, y, sigma)") +
                        ": NumPy in, NumPy out, System rebuilt every step (K0 derives the constants on the device)",
                "c_abi": {"value": points_per_step / (c_ms * 1e-3), "unit": "points/s", "ms_per_step": c_ms,
                          "path": "jxp_orbital_elements_f64 + jxp_system_light_curve_f64 from one pinned host arena "
                                  "(one H2D copy per step), pinned D2H of the log-likelihoods"}},
        "launch": ("CUDA graph replay of the captured C-ABI call" + (" + peer-store exchange" if world > 1 else "")
                   if use_graph else "stream-ordered launches through the C ABI"),
        "ms_per_step_stream_launches": eager_ms,
        "per_step_ms_rank0": per_step_ms, "weak": weak, "exchange": exchange, "flux_all_gather": flux_gather,
        "gpu_launches": len(launches) * args.steps,
        "gpu_launches_per_step": launches,
        "clocks": clocks, "secondary": secondary,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def secondary_configs(dev, lib, ev0, ev1, rank, world, fp64_peak, max_over_ranks, barrier):
    """Configs 1, 2, 4 and 5 at full size, their set / pair axis sharded over the ranks (no collective);
    best of 3 per rank, max over ranks, whole-job rates."""
    import torch

    from jaxoplanet_b200 import _array as A
    from jaxoplanet_b200 import _lib
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.distributed import shard_range
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

    out = {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def best_of(fn, n=3, warm=5):
        # (five untimed runs first: on freshly allocated output buffers the float K5 takes 6.7 ms for its first four
        # runs and 5.9 ms from the fifth on -- tools/k5_f32_order.py --, whatever ran before)
        best = 1e30
        barrier()
        for i in range(warm + n):
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            if i >= warm:
                best = min(best, ev0.elapsed_time(ev1))
        return max_over_ranks(best)

    # C2: batched Kepler solve, 1e8 (M, e) pairs
    n2 = 100_000_000
    lo, hi = shard_range(n2, rank, world)
    g = torch.Generator(device=dev)
    g.manual_seed(rank)
    M = torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g) * (2 * np.pi)
    e = torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g) * 0.99
    s_ = torch.empty_like(M)
    c_ = torch.empty_like(M)
    ms = best_of(lambda: _lib.call("jxp_kepler_f64", M.data_ptr(), 1, e.data_ptr(), 1, hi - lo, s_.data_ptr(),
                                   c_.data_ptr(), 0, 0, 0, A.stream_ptr()))
    out["c2_kepler_1e8"] = {
        "solves_per_s": n2 / (ms * 1e-3), "ms": ms,
        "fp64_frac": n2 * W_KEPLER / (ms * 1e-3) / 1e12 / (fp64_peak * world),
        "hbm_gbs": 32.0 * n2 / (ms * 1e-3) / 1e9,
        "hbm_frac": (32.0 * n2 / (ms * 1e-3) / 1e9) / (peaks["hbm_gbs"] * world) if peaks.get("hbm_gbs") else None}
    del M, e, s_, c_
    if world == 1:
        # K2: core.limb_dark.light_curve on 1e8 random in-transit (b, r) points
        rr = torch.rand(n2, dtype=torch.float64, device=dev) * 0.12 + 0.03
        bb = torch.rand(n2, dtype=torch.float64, device=dev) * (1 + rr)
        u2 = torch.tensor([0.3, 0.2], dtype=torch.float64, device=dev)
        f_ = torch.empty_like(bb)
        ms = best_of(lambda: _lib.call("jxp_limb_dark_f64", u2.data_ptr(), 2, bb.data_ptr(), 1, rr.data_ptr(), 1, n2,
                                       10, f_.data_ptr(), 0, 0, 0, A.stream_ptr()))
        n_in = int((bb < 1 - rr).sum())
        out["k2_limb_dark_1e8_in_transit"] = {
            "evaluations_per_s": n2 / (ms * 1e-3), "ms": ms,
            "fp64_frac": (n_in * W_LD_INSIDE + (n2 - n_in) * W_LD) / (ms * 1e-3) / 1e12 / fp64_peak,
            "fp64_frac_r1_convention": n2 * W_LD / (ms * 1e-3) / 1e12 / fp64_peak,
            "inside_disk_fraction": n_in / n2}
        del rr, bb, f_
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
    lo, hi = shard_range(1024, rank, world)
    bodies, n_sets, _ = pack_bodies(W.system_from({k: v[lo:hi] for k, v in p.items()}))
    bodies = bodies.to(dev).contiguous()
    ud = torch.as_tensor(np.ascontiguousarray(u[lo:hi])).to(dev)
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
    wk4 = (ctypes.c_int64 * 4)()
    _lib.call("jxp_system_walk_stats", ws.data_ptr(), n_sets, 3, n_times, ctypes.addressof(wk4), A.stream_ptr())
    out["c4_3planets_15x_supersampled"] = {
        "points_per_s": 1024 * n_times / (ms * 1e-3), "ms": ms,
        "inner_evaluations_per_s": 1024 * n_times * 45 / (ms * 1e-3),
        "rank0_executed": {"sub_samples_visited": int(stats[0]), "flux_evaluations": int(stats[1]),
                           "from_tables": int(wk4[0]), "set_body_units_with_tables": int(wk4[2])},
        # the reference formula's flops for the visited sub-samples over the time; NOT a pipe utilisation (the
        # per-set function tables answer an evaluation with ~50 FP64 instructions)
        "rank0_reference_formula_tflops": (int(stats[0]) * (W_KEPLER + W_ORBIT) + int(stats[1]) * W_LD) / (ms * 1e-3) / 1e12,
        "rank0_in_transit_fraction": float((fs != 0).double().mean())}
    del fs, ws, bodies
    torch.cuda.empty_cache()

    # C5: flux + 8 partials, 16384 sets x 5e4 times, f64 and f32
    theta, t5 = W.c5()
    lo, hi = shard_range(16384, rank, world)
    thd = torch.as_tensor(np.ascontiguousarray(theta[lo:hi])).to(dev)
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
        gbs = 16384 * nt5 * 9 * fl.element_size() / (ms * 1e-3) / 1e9
        out["c5_flux_plus_8_partials_" + nm] = {
            "points_per_s": 16384 * nt5 / (ms * 1e-3), "ms": ms, "store_gbs": gbs,
            "hbm_frac": gbs / (peaks["hbm_gbs"] * world) if peaks.get("hbm_gbs") else None}
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
        "points_per_s": 16384 * nt5 / (ms * 1e-3), "ms": ms,
        "path": "transit list (k_tl_build, k_tlg_eval, k_tlg_reduce)" if (ls5[0] > 0 and ls5[1] == 0)
                else "window-test kernel (k_transit_partials)",
        "rank0_list_items": int(ls5[0])}
    out["sharding"] = f"set / pair axis split over {world} rank(s), no collective; time = max over ranks"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time stream-ordered launches instead of a graph replay")
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
