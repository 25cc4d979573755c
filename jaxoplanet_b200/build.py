"""Build the sm_100a shared library in-tree (jaxoplanet_b200/lib/libjxp_b200.so).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU
box with the repo snapshot.  Usage: ``python -m jaxoplanet_b200.build [--force]``.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libjxp_b200.so")
SOURCES = ["jxp_kernels.cu", "jxp_partials.cu", "jxp_transit.cu", "jxp_walk.cu", "jxp_state.cu", "jxp_peer.cu"]
HEADERS = [
    os.path.join(CSRC, "jxp_scalar.cuh"),
    os.path.join(CSRC, "jxp_fastmath.cuh"),
    os.path.join(CSRC, "jxp_host.h"),
    os.path.join(CSRC, "jxp_math.cuh"),
    os.path.join(CSRC, "jxp_orbit.cuh"),
    os.path.join(CSRC, "jxp_list.cuh"),
    os.path.join(CSRC, "jxp_prep.cuh"),
    os.path.join(CSRC, "jxp_tables.cuh"),
    os.path.join(CSRC, "jxp_tables_data.inc"),
    os.path.join(ROOT, "include", "jaxoplanet_b200.h"),
]

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
    "--fmad=true",
    "-Xptxas",
    "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; jaxoplanet_b200 has no CPU fallback")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    log = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB_PATH + ".tmp", *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    with open(os.path.join(LIB_DIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
