"""ctypes wrapper of oracle/ref_c.c.  TEST INFRASTRUCTURE ONLY (CPU baseline for bench.py)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libref_c.so")
_lib = None


STAMP = os.path.join(HERE, "_build", "host_cpu.txt")


def _host_cpu() -> str:
    """Model name + ISA flags of this host: the library is built with -march=native, so a copy built on
    another machine (the builder container vs the GPU box) must not be loaded."""
    model, flags = "", ""
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("model name") and not model:
                    model = ln.split(":", 1)[1].strip()
                elif ln.startswith("flags") and not flags:
                    flags = " ".join(sorted(ln.split(":", 1)[1].split()))
                if model and flags:
                    break
    except OSError:
        pass
    import hashlib

    return f"{model} {hashlib.sha1(flags.encode()).hexdigest()[:12]}"


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "ref_c.c")
    mk = os.path.join(HERE, "Makefile")
    stale = (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(mk)))
    cpu = _host_cpu()
    try:
        same_host = open(STAMP).read().strip() == cpu
    except OSError:
        same_host = False
    if force or stale or not same_host:
        subprocess.run(["make", "-C", HERE, "-B"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        with open(STAMP, "w") as fh:
            fh.write(cpu + "\n")
    return LIB


def load():
    global _lib
    if _lib is None:
        build()  # (no-op when the library is current AND was built on this host)
        lib = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        lib.refc_max_threads.restype = C.c_int
        lib.refc_set_threads.argtypes = [C.c_int]
        lib.refc_kepler.argtypes = [dp, dp, C.c_int64, dp, dp]
        lib.refc_ld_quad.argtypes = [C.c_double, C.c_double, dp, dp, C.c_int64, dp]
        lib.refc_system.argtypes = [dp, C.c_int32, C.c_int32, dp, C.c_int64, dp, C.c_int64, dp, dp,
                                    C.c_int32, dp, dp, dp, C.c_int64, dp]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def max_threads() -> int:
    return load().refc_max_threads()


def use_all_cores() -> int:
    """Use every core this process may run on (launchers such as torchrun export
    OMP_NUM_THREADS=1, which would silently make the CPU baseline single-threaded)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    load().refc_set_threads(n)
    return max_threads()


def kepler(M, ecc):
    M = np.ascontiguousarray(M, dtype=np.float64)
    ecc = np.ascontiguousarray(np.broadcast_to(ecc, M.shape), dtype=np.float64)
    s = np.empty_like(M)
    c = np.empty_like(M)
    load().refc_kepler(_p(M), _p(ecc), M.size, _p(s), _p(c))
    return s, c


def ld_quad(u, b, r):
    b = np.ascontiguousarray(b, dtype=np.float64)
    r = np.ascontiguousarray(np.broadcast_to(r, b.shape), dtype=np.float64)
    out = np.empty_like(b)
    load().refc_ld_quad(float(u[0]), float(u[1]), _p(b), _p(r), b.size, _p(out))
    return out


def system(bodies, u, t, *, dt=None, w=None, y=None, sigma=None, want_flux=True):
    """bodies [n_sets, n_bodies, 12], u [n_sets, 2] or [2], t [n_times]."""
    bodies = np.ascontiguousarray(bodies, dtype=np.float64)
    n_sets, n_bodies, _ = bodies.shape
    u = np.ascontiguousarray(u, dtype=np.float64)
    u_stride = 2 if u.ndim == 2 else 0
    t = np.ascontiguousarray(t, dtype=np.float64)
    dt = np.zeros(1) if dt is None else np.ascontiguousarray(dt, dtype=np.float64)
    w = np.ones(1) if w is None else np.ascontiguousarray(w, dtype=np.float64)
    flux = np.empty((n_sets, t.size)) if want_flux else None
    ll = None
    s_stride = 0
    if y is not None:
        y = np.ascontiguousarray(y, dtype=np.float64)
        sigma = np.ascontiguousarray(np.atleast_1d(sigma), dtype=np.float64)
        s_stride = 0 if sigma.size == 1 else 1
        ll = np.empty(n_sets)
    load().refc_system(_p(bodies), n_sets, n_bodies, _p(u), u_stride, _p(t), t.size, _p(dt), _p(w),
                       dt.size, _p(flux), _p(y), _p(sigma), s_stride, _p(ll))
    return flux, ll
