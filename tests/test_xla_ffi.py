"""The XLA FFI boundary (jaxoplanet_b200/csrc/jxp_xla_ffi.cc + jaxoplanet_b200/jax_bindings.py).

JAX is not installable in the development image, so here the handler translation unit is compiled
against tests/xla_ffi_stub/ (same names and call signatures as xla/ffi/api/ffi.h, no behaviour): every
handler is type-checked against its binding and the library must link against libjxp_b200.so and export
one symbol per registered target.  Where ``import jax`` works (and a GPU is present) the real header is
used and the custom calls are run against the oracle.
"""

import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "xla_ffi_stub")
SRC = os.path.join(ROOT, "jaxoplanet_b200", "csrc", "jxp_xla_ffi.cc")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


def _lib_dir():
    from jaxoplanet_b200 import build as b

    b.build_library()
    return b.LIB_DIR


def test_ffi_handlers_type_check_link_and_export_every_target(tmp_path):
    from jaxoplanet_b200 import jax_bindings as jb

    out = str(tmp_path / "libjxp_b200_ffi_check.so")
    cmd = ["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-Wall", "-Werror", f"-I{STUB}", f"-I{CUDA_INC}", SRC,
           f"-L{_lib_dir()}", "-ljxp_b200", "-o", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    nm = subprocess.run(["nm", "-D", "--defined-only", out], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(re.findall(r" T (Jxp\w+)", nm))
    assert exported == set(jb.TARGETS), exported ^ set(jb.TARGETS)
    # every C entry point of the hot path is reachable from some handler
    und = subprocess.run(["nm", "-D", "--undefined-only", out], stdout=subprocess.PIPE, text=True, check=True).stdout
    used = set(re.findall(r" U (jxp_\w+)", und))
    for sym in ("jxp_kepler_f64", "jxp_kepler_f32", "jxp_limb_dark_f64", "jxp_limb_dark_f32",
                "jxp_system_light_curve_f64", "jxp_system_light_curve_f32", "jxp_system_workspace_bytes",
                "jxp_transit_partials_f64", "jxp_transit_partials_f32", "jxp_transit_workspace_bytes",
                "jxp_transit_partials_poly_f64", "jxp_transit_poly_workspace_bytes",
                "jxp_orbital_elements_f64", "jxp_orbit_state_f64", "jxp_orbit_state_f32",
                "jxp_transit_orbit_light_curve_f64", "jxp_transit_orbit_light_curve_f32",
                "jxp_interp_periodic_f64", "jxp_interp_periodic_f32", "jxp_peer_allgather_f64", "jxp_last_error"):
        assert sym in used, sym


def test_stub_rejects_a_handler_that_does_not_match_its_binding(tmp_path):
    """The stand-in header is only worth something if a wrong binding fails to compile."""
    src = open(SRC).read()
    good = "XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpKeplerF64, KeplerF64, JXP_STREAM.A64.A64.R64.R64.R64.R64);"
    assert good in src
    bad = str(tmp_path / "bad.cc")
    with open(bad, "w") as fh:
        fh.write(src.replace(good, good.replace(".R64.R64.R64.R64", ".R64.R64.R64"))
                 .replace('#include "../../include/jaxoplanet_b200.h"',
                          f'#include "{os.path.join(ROOT, "include", "jaxoplanet_b200.h")}"'))
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", f"-I{STUB}", f"-I{CUDA_INC}", bad],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode != 0 and "does not match its binding" in r.stdout


def test_jax_bindings_module_imports_without_jax_and_names_every_symbol():
    from jaxoplanet_b200 import jax_bindings as jb

    assert len(set(jb.TARGETS.values())) == len(jb.TARGETS) == 24
    if not jb.available():
        assert not hasattr(jb, "kepler")  # nothing pretends to work without JAX
        cmd = jb.ffi_compile_command("/nonexistent/include", "/tmp/x.so")
        assert cmd[0] == "g++" and any(a.endswith("jxp_xla_ffi.cc") for a in cmd)


@pytest.mark.gpu
def test_jax_custom_calls_match_oracle_when_jax_is_installed():
    jax = pytest.importorskip("jax")
    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp

    from jaxoplanet_b200 import jax_bindings as jb
    from oracle import ref_numpy as ref

    rng = np.random.default_rng(0)
    M = rng.uniform(-10, 10, 4096)
    e = rng.uniform(0, 0.9, 4096)
    s, c = jb.kepler(jnp.asarray(M), jnp.asarray(e))
    s0, c0 = ref.kepler(M, e)
    assert np.max(np.abs(np.asarray(s) - s0)) <= 1e-12 and np.max(np.abs(np.asarray(c) - c0)) <= 1e-12
    # jit + vmap + grad compose (kepler.py:59-79 rule, factors from the kernel)
    g = jax.jit(jax.vmap(jax.grad(lambda m, ee: jb.kepler(m, ee)[0])))(jnp.asarray(M[:64]), jnp.asarray(e[:64]))
    sd, cd = ref.kepler_jvp(M[:64], e[:64], np.ones(64), np.zeros(64))[1]
    assert np.max(np.abs(np.asarray(g) - sd)) <= 1e-9 * np.max(np.abs(sd))
    u = np.array([0.3, 0.2])
    b = np.linspace(0, 1.2, 1001)
    lc = jb.light_curve(jnp.asarray(u), jnp.asarray(b), 0.1)
    assert np.max(np.abs(np.asarray(lc) - ref.ld_light_curve(u, b, np.full_like(b, 0.1)))) <= 1e-12
