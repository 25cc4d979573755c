"""JAX side of the XLA FFI boundary (csrc/jxp_xla_ffi.cc): the hot-path callables of jaxoplanet as
``jax.ffi.ffi_call`` custom calls with ``custom_jvp`` rules fed by the kernels, so ``jit`` / ``vmap`` /
``grad`` and NumPyro compose as they do with the reference.

Reference callables (exoplanet-dev/jaxoplanet):
  kepler(M, ecc)                          src/jaxoplanet/core/kepler.py:14-79 (jit, custom_jvp, defjvp)
  light_curve(u, b, r, order)             src/jaxoplanet/core/limb_dark.py:19-47
  system_light_curve / system_log_likelihood   src/jaxoplanet/light_curves/limb_dark.py:15-62 (+ transforms.integrate)
  transit_light_curve(theta, t)           the same, differentiable w.r.t. (P, t0, e, omega, b, r, u1, u2) by the fused
                                          forward-mode kernel instead of JAX autodiff

This module imports ``jax`` at import time and therefore only loads where JAX is installed (it is not
in the image this repository is developed in: ``available()`` is False there and nothing else in the
package imports this module).  ``build_ffi()`` compiles the handler library against
``jax.ffi.include_dir()``; ``register()`` registers its symbols for platform "CUDA".
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FFI_SRC = os.path.join(HERE, "csrc", "jxp_xla_ffi.cc")
FFI_LIB = os.path.join(HERE, "lib", "libjxp_b200_ffi.so")

# exported handler symbol -> FFI target name
TARGETS = {
    "JxpKeplerF64": "jxp_kepler_f64", "JxpKeplerF32": "jxp_kepler_f32",
    "JxpLimbDarkF64": "jxp_limb_dark_f64", "JxpLimbDarkF32": "jxp_limb_dark_f32",
    "JxpSystemFluxF64": "jxp_system_flux_f64", "JxpSystemFluxF32": "jxp_system_flux_f32",
    "JxpSystemFluxSumF64": "jxp_system_flux_sum_f64", "JxpSystemFluxSumF32": "jxp_system_flux_sum_f32",
    "JxpSystemLogLikeF64": "jxp_system_loglike_f64", "JxpSystemLogLikeF32": "jxp_system_loglike_f32",
    "JxpTransitPartialsF64": "jxp_transit_partials_f64", "JxpTransitPartialsF32": "jxp_transit_partials_f32",
    "JxpTransitGradientF64": "jxp_transit_gradient_f64", "JxpTransitGradientF32": "jxp_transit_gradient_f32",
    "JxpTransitPartialsPolyF64": "jxp_transit_partials_poly_f64",
    "JxpTransitGradientPolyF64": "jxp_transit_gradient_poly_f64",
    "JxpOrbitalElementsF64": "jxp_orbital_elements_f64",
    "JxpPeerAllGatherF64": "jxp_peer_allgather_f64",
    "JxpOrbitStateF64": "jxp_orbit_state_f64", "JxpOrbitStateF32": "jxp_orbit_state_f32",
    "JxpTransitOrbitF64": "jxp_transit_orbit_f64", "JxpTransitOrbitF32": "jxp_transit_orbit_f32",
    "JxpInterpPeriodicF64": "jxp_interp_periodic_f64", "JxpInterpPeriodicF32": "jxp_interp_periodic_f32",
}


def available() -> bool:
    """True when JAX with the FFI API is importable (then the bindings below can be built and used)."""
    try:
        import jax  # noqa: F401
        import jax.ffi  # noqa: F401

        return True
    except Exception:
        return False


def ffi_compile_command(include_dir: str, out: str) -> list[str]:
    from jaxoplanet_b200.build import LIB_DIR

    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    return ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", f"-I{include_dir}", f"-I{cuda_inc}", FFI_SRC,
            f"-L{LIB_DIR}", "-ljxp_b200", "-Wl,-rpath,$ORIGIN", "-o", out]


def build_ffi(force: bool = False) -> str:
    """Compile csrc/jxp_xla_ffi.cc against the XLA FFI headers shipped with jaxlib."""
    import jax.ffi

    from jaxoplanet_b200.build import build_library

    build_library()
    if force or not os.path.exists(FFI_LIB) or os.path.getmtime(FFI_LIB) < os.path.getmtime(FFI_SRC):
        r = subprocess.run(ffi_compile_command(jax.ffi.include_dir(), FFI_LIB), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building {FFI_LIB} failed:\n{r.stdout}")
    return FFI_LIB


_registered = False


def register() -> None:
    """jax.ffi.register_ffi_target for every handler (platform "CUDA"); idempotent."""
    global _registered
    if _registered:
        return
    import jax.ffi

    lib = ctypes.CDLL(build_ffi())
    for symbol, target in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    _registered = True


if available():
    import jax
    import jax.numpy as jnp

    def _suffix(dtype) -> str:
        return "f64" if jnp.dtype(dtype) == jnp.float64 else "f32"

    def _expo(exposure_time, exp_order, num_samples, order, flags=0):
        return dict(exposure_time=np.float64(0.0 if exposure_time is None else exposure_time),
                    exp_order=np.int64(exp_order), num_samples=np.int64(num_samples), gl_order=np.int64(order),
                    flags=np.int64(flags))

    # ------------------------------------------------------------------ core.kepler (kepler.py:14-79)
    def _kepler_call(M, ecc):
        register()
        M, ecc = jnp.broadcast_arrays(jnp.asarray(M), jnp.asarray(ecc))
        out = jax.ShapeDtypeStruct(M.shape, M.dtype)
        return jax.ffi.ffi_call("jxp_kepler_" + _suffix(M.dtype), (out,) * 4, vmap_method="broadcast_all")(M, ecc)

    @jax.custom_jvp
    def _kepler(M, ecc):
        sinf, cosf, _, _ = _kepler_call(M, ecc)
        return sinf, cosf

    @_kepler.defjvp
    def _kepler_jvp(primals, tangents):
        # the rule of kepler.py:59-79 with its two factors from the kernel; linear in the tangents and written
        # in jnp, so JAX can transpose it for reverse mode
        M, ecc = primals
        M_dot, e_dot = tangents
        sinf, cosf, dfdM, dfde = _kepler_call(M, ecc)
        f_dot = M_dot * dfdM + e_dot * dfde
        return (sinf, cosf), (cosf * f_dot, -sinf * f_dot)

    @jax.jit
    def kepler(M, ecc):
        """Drop-in for jaxoplanet.core.kepler.kepler: (sin f, cos f) of the true anomaly."""
        return _kepler(M, ecc)

    # ------------------------------------------------------------------ core.limb_dark (limb_dark.py:19-47)
    def _limb_dark_call(u, b, r, order):
        register()
        u = jnp.asarray(u)
        assert u.ndim == 1, "u must be one-dimensional (core/limb_dark.py:40)"
        b, r = jnp.broadcast_arrays(jnp.asarray(b, dtype=u.dtype), jnp.asarray(r, dtype=u.dtype))
        one = jax.ShapeDtypeStruct(b.shape, b.dtype)
        svec = jax.ShapeDtypeStruct((u.shape[0] + 1,) + b.shape, b.dtype)
        return jax.ffi.ffi_call("jxp_limb_dark_" + _suffix(b.dtype), (one, one, one, svec),
                                vmap_method="sequential")(u, b, r, order=np.int64(order))

    def _greens(u):
        """g(u) of limb_dark.py:87-99, normalised as in :45 (plain jnp: differentiable, tiny)."""
        from scipy.special import binom

        u = jnp.concatenate([-jnp.ones(1, dtype=u.dtype), u])
        size = u.shape[0]
        i = np.arange(size)
        arg = binom(i[None, :], i[:, None]) @ jnp.asarray(u)
        p = (-1) ** (i + 1) * arg
        g = [jnp.zeros((), dtype=u.dtype)] * (size + 2)
        for n in range(size - 1, 1, -1):
            g[n] = p[n] / (n + 2) + g[n + 2]
        if size > 1:
            g[1] = p[1] + 3 * g[3]
        g[0] = p[0] + 2 * g[2]
        g = jnp.stack(g[:size])
        return g / (jnp.pi * (g[0] + g[1] / 1.5))

    @partial(jax.custom_jvp, nondiff_argnums=(3,))
    def _limb_dark(u, b, r, order):
        return _limb_dark_call(u, b, r, order)[0]

    @_limb_dark.defjvp
    def _limb_dark_jvp(order, primals, tangents):
        u, b, r = primals
        u_dot, b_dot, r_dot = tangents
        flux, dfdb, dfdr, svec = _limb_dark_call(u, b, r, order)
        # d flux / d u = s . dg/du (s does not depend on u); dg/du by JAX on the tiny transform
        _, g_dot = jax.jvp(_greens, (jnp.asarray(u),), (jnp.asarray(u_dot),))
        return flux, dfdb * b_dot + dfdr * r_dot + jnp.tensordot(g_dot, svec, axes=1)

    @partial(jax.jit, static_argnames=("order",))
    def light_curve(u, b, r, *, order: int = 10):
        """Drop-in for jaxoplanet.core.limb_dark.light_curve: flux - 1 of a limb-darkened star."""
        return _limb_dark(jnp.asarray(u), b, r, order)

    # ------------------------------------------------------------------ fused system light curve
    def orbital_elements(elements):
        """OrbitalBody.__init__ on the device: element records [n][16] (NaN = not given) -> body records
        [n][12] (keplerian.py:262-340)."""
        register()
        elements = jnp.asarray(elements, dtype=jnp.float64)
        out = jax.ShapeDtypeStruct(elements.shape[:-1] + (12,), jnp.float64)
        return jax.ffi.ffi_call("jxp_orbital_elements_f64", out, vmap_method="sequential")(elements)

    def peer_all_gather(local, *, offset, n_total, rank, world, bufs):
        """The exchange of the set-sharded path (one process per GPU): this rank's block of per-set values -> the
        (n_total,) vector on every rank, by peer stores over NVLink.  `bufs`: addresses of the ranks' exchange buffers
        as mapped into this process (jaxoplanet_b200.distributed.PeerGather sets them up; at most eight)."""
        register()
        addr = [int(b) for b in bufs] + [0] * (8 - len(bufs))
        out = jax.ShapeDtypeStruct((int(n_total),), jnp.float64)
        return jax.ffi.ffi_call("jxp_peer_allgather_f64", out, has_side_effect=True)(
            jnp.asarray(local, dtype=jnp.float64), offset=np.int64(offset), rank=np.int64(rank), world=np.int64(world),
            **{f"buf{i}": np.int64(a) for i, a in enumerate(addr)})

    def _system(kind, bodies, u, t, extra, out, dtype, exposure_time, exp_order, num_samples, order):
        register()
        bodies = jnp.asarray(bodies, dtype=jnp.float64)
        return jax.ffi.ffi_call(f"jxp_system_{kind}_" + _suffix(dtype), out, vmap_method="sequential")(
            bodies, jnp.asarray(u, dtype=jnp.float64), jnp.asarray(t, dtype=jnp.float64).reshape(-1), *extra,
            **_expo(exposure_time, exp_order, num_samples, order))

    def system_light_curve(bodies, u, t, *, dtype=jnp.float64, exposure_time=None, exp_order=0, num_samples=7,
                           order=10, per_body=True):
        """light_curves.limb_dark.light_curve(System, *u)(t) for a batch of parameter sets: bodies
        [n_sets][n_bodies][12] (orbital_elements / pack_bodies), u [n_u] or [n_sets][n_u], t any shape ->
        [n_sets] + t.shape + [n_bodies] (per_body) or [n_sets] + t.shape (summed over bodies)."""
        t = jnp.asarray(t)
        n_sets, n_bodies = bodies.shape[0], bodies.shape[1]
        if per_body:
            out = jax.ShapeDtypeStruct((n_sets, t.size, n_bodies), dtype)
            return _system("flux", bodies, u, t, (), out, dtype, exposure_time, exp_order, num_samples,
                           order).reshape((n_sets,) + t.shape + (n_bodies,))
        out = jax.ShapeDtypeStruct((n_sets, t.size), dtype)
        return _system("flux_sum", bodies, u, t, (), out, dtype, exposure_time, exp_order, num_samples,
                       order).reshape((n_sets,) + t.shape)

    def system_log_likelihood(bodies, u, t, y, sigma, *, dtype=jnp.float64, exposure_time=None, exp_order=0,
                              num_samples=7, order=10):
        """sum_t Normal(1 + sum over bodies of the light curve, sigma).log_prob(y) per parameter set, fused."""
        out = jax.ShapeDtypeStruct((bodies.shape[0],), dtype)
        extra = (jnp.asarray(y, dtype=dtype).reshape(-1), jnp.atleast_1d(jnp.asarray(sigma, dtype=dtype)).reshape(-1))
        return _system("loglike", bodies, u, t, extra, out, dtype, exposure_time, exp_order, num_samples, order)

    # ------------------------------------------------------------------ differentiable transit light curve
    def _partials_call(theta, star, t, dtype, expo):
        register()
        n_sets, n_t = theta.shape[0], t.shape[0]
        outs = (jax.ShapeDtypeStruct((n_sets, n_t), dtype), jax.ShapeDtypeStruct((n_sets, 8, n_t), dtype))
        return jax.ffi.ffi_call("jxp_transit_partials_" + _suffix(dtype), outs, vmap_method="sequential")(
            theta, star, t, **expo)

    @partial(jax.custom_jvp, nondiff_argnums=(3, 4, 5, 6, 7))
    def _transit_lc(theta, star, t, dtype, exposure_time, exp_order, num_samples, order):
        return _partials_call(theta, star, t, dtype, _expo(exposure_time, exp_order, num_samples, order))[0]

    @_transit_lc.defjvp
    def _transit_lc_jvp(dtype, exposure_time, exp_order, num_samples, order, primals, tangents):
        theta, star, t = primals
        theta_dot = tangents[0]
        flux, part = _partials_call(theta, star, t, dtype, _expo(exposure_time, exp_order, num_samples, order))
        # sum_k d flux / d theta_k * theta_dot_k: linear in the tangent, in jnp (transposable => jax.grad works)
        return flux, jnp.einsum("skt,sk->st", part, theta_dot.astype(part.dtype))

    def transit_light_curve(theta, t, *, central_mass=1.0, central_radius=1.0, dtype=jnp.float64,
                            exposure_time=None, exp_order=0, num_samples=7, order=10):
        """flux - 1 of single-planet systems, theta [n_sets][8] = (period, time_transit, eccentricity,
        omega_peri, impact_param, radius, u1, u2); differentiable w.r.t. theta through the fused forward-mode
        kernel (K5) -- what jax.grad of the reference evaluates by autodiff."""
        theta = jnp.asarray(theta, dtype=jnp.float64)
        single = theta.ndim == 1
        theta = theta.reshape(-1, 8)
        star = jnp.stack([jnp.asarray(central_mass, dtype=jnp.float64), jnp.asarray(central_radius, dtype=jnp.float64)])
        out = _transit_lc(theta, star, jnp.asarray(t, dtype=jnp.float64).reshape(-1), dtype, exposure_time,
                          exp_order, num_samples, order)
        return out[0] if single else out

    # ---- polynomial limb darkening of any length: theta [n_sets][6 + n_u] (double precision, order 10)
    def _poly_partials_call(theta, star, t):
        register()
        n_sets, n_theta, n_t = theta.shape[0], theta.shape[1], t.shape[0]
        outs = (jax.ShapeDtypeStruct((n_sets, n_t), jnp.float64), jax.ShapeDtypeStruct((n_sets, n_theta, n_t), jnp.float64))
        return jax.ffi.ffi_call("jxp_transit_partials_poly_f64", outs, vmap_method="sequential")(
            theta, star, t, gl_order=np.int64(10), flags=np.int64(0))

    @jax.custom_jvp
    def _transit_lc_poly(theta, star, t):
        return _poly_partials_call(theta, star, t)[0]

    @_transit_lc_poly.defjvp
    def _transit_lc_poly_jvp(primals, tangents):
        theta, star, t = primals
        flux, part = _poly_partials_call(theta, star, t)
        return flux, jnp.einsum("skt,sk->st", part, tangents[0].astype(part.dtype))

    def transit_light_curve_poly(theta, t, *, central_mass=1.0, central_radius=1.0):
        """As transit_light_curve for a polynomial limb-darkening law of any length: theta [n_sets][6 + n_u] = (period,
        time_transit, eccentricity, omega_peri, impact_param, radius, u_1 .. u_n), n_u <= 8; differentiable w.r.t.
        theta (the reference differentiates core.limb_dark.light_curve for any u)."""
        theta = jnp.asarray(theta, dtype=jnp.float64)
        single = theta.ndim == 1
        theta = theta.reshape(-1, theta.shape[-1])
        star = jnp.stack([jnp.asarray(central_mass, dtype=jnp.float64), jnp.asarray(central_radius, dtype=jnp.float64)])
        out = _transit_lc_poly(theta, star, jnp.asarray(t, dtype=jnp.float64).reshape(-1))
        return out[0] if single else out

    def _poly_gradient_call(theta, star, t, y, sigma):
        register()
        outs = (jax.ShapeDtypeStruct((theta.shape[0],), jnp.float64), jax.ShapeDtypeStruct(theta.shape, jnp.float64))
        return jax.ffi.ffi_call("jxp_transit_gradient_poly_f64", outs, vmap_method="sequential")(
            theta, star, t, y, sigma, gl_order=np.int64(10), flags=np.int64(0))

    @jax.custom_jvp
    def _transit_ll_poly(theta, star, t, y, sigma):
        return _poly_gradient_call(theta, star, t, y, sigma)[0]

    @_transit_ll_poly.defjvp
    def _transit_ll_poly_jvp(primals, tangents):
        ll, dll = _poly_gradient_call(*primals)
        return ll, jnp.sum(dll * tangents[0].astype(dll.dtype), axis=-1)

    def transit_log_likelihood_poly(theta, t, y, sigma, *, central_mass=1.0, central_radius=1.0):
        """Gaussian log-likelihood per parameter set for theta [n_sets][6 + n_u], differentiable w.r.t. theta."""
        theta = jnp.asarray(theta, dtype=jnp.float64)
        theta = theta.reshape(-1, theta.shape[-1])
        star = jnp.stack([jnp.asarray(central_mass, dtype=jnp.float64), jnp.asarray(central_radius, dtype=jnp.float64)])
        return _transit_ll_poly(theta, star, jnp.asarray(t, dtype=jnp.float64).reshape(-1),
                                jnp.asarray(y, dtype=jnp.float64).reshape(-1),
                                jnp.atleast_1d(jnp.asarray(sigma, dtype=jnp.float64)).reshape(-1))

    @partial(jax.custom_jvp, nondiff_argnums=(5, 6, 7, 8, 9))
    def _transit_ll(theta, star, t, y, sigma, dtype, exposure_time, exp_order, num_samples, order):
        register()
        outs = (jax.ShapeDtypeStruct((theta.shape[0],), dtype), jax.ShapeDtypeStruct((theta.shape[0], 8), dtype))
        return jax.ffi.ffi_call("jxp_transit_gradient_" + _suffix(dtype), outs, vmap_method="sequential")(
            theta, star, t, y, sigma, **_expo(exposure_time, exp_order, num_samples, order))[0]

    @_transit_ll.defjvp
    def _transit_ll_jvp(dtype, exposure_time, exp_order, num_samples, order, primals, tangents):
        theta, star, t, y, sigma = primals
        register()
        outs = (jax.ShapeDtypeStruct((theta.shape[0],), dtype), jax.ShapeDtypeStruct((theta.shape[0], 8), dtype))
        ll, dll = jax.ffi.ffi_call("jxp_transit_gradient_" + _suffix(dtype), outs, vmap_method="sequential")(
            theta, star, t, y, sigma, **_expo(exposure_time, exp_order, num_samples, order))
        return ll, jnp.sum(dll * tangents[0].astype(dll.dtype), axis=-1)

    def transit_log_likelihood(theta, t, y, sigma, *, central_mass=1.0, central_radius=1.0, dtype=jnp.float64,
                               exposure_time=None, exp_order=0, num_samples=7, order=10):
        """Gaussian log-likelihood per parameter set with its theta-gradient reduced inside the kernel (the
        gradient step NUTS takes): jax.grad of this costs one launch and stores nothing per sample.
        Differentiable w.r.t. theta only."""
        theta = jnp.asarray(theta, dtype=jnp.float64).reshape(-1, 8)
        star = jnp.stack([jnp.asarray(central_mass, dtype=jnp.float64), jnp.asarray(central_radius, dtype=jnp.float64)])
        return _transit_ll(theta, star, jnp.asarray(t, dtype=jnp.float64).reshape(-1),
                           jnp.asarray(y, dtype=dtype).reshape(-1),
                           jnp.atleast_1d(jnp.asarray(sigma, dtype=dtype)).reshape(-1), dtype, exposure_time,
                           exp_order, num_samples, order)
