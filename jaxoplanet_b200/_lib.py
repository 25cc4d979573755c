"""ctypes binding of include/jaxoplanet_b200.h.

There is no CPU fallback: if the shared library is missing the import of any compute
entry point raises, and every non-zero status from the C ABI becomes a ``JxpError``.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JXP_LIB_PATH") or os.path.join(HERE, "lib", "libjxp_b200.so")

JXP_BODY_NFIELDS = 12
JXP_NTHETA = 8
JXP_FLAG_NO_WINDOW = 1
JXP_FLAG_COUNT = 2
JXP_STATE_NO_INCLINATION = 1
# development / measurement options (jxp_set_dev_options): paths that return the same results
JXP_DEV_NO_WALK = 1
JXP_DEV_NO_PAIR = 2
JXP_DEV_WALK_ANY_SIZE = 4
JXP_DEV_NO_LD_CLASSES = 8
JXP_DEV_PROFILE_PLAN = 16
JXP_DEV_NO_TABLES = 32
JXP_DEV_PROFILE_TABLES = 64
JXP_DEV_PROFILE_ITEMS = 128
JXP_DEV_NO_BULK_FILL = 4096
JXP_DEV_BULK_FILL_SWAP = 8192


def JXP_DEV_CTA_WARPS(n):
    return (int(n) & 0xF) << 8


def JXP_DEV_TR_TPW(n):
    return (int(n) & 0x3) << 14


def JXP_DEV_TR_RUN(n):
    return (int(n) & 0x7F) << 16


JXP_TO_NFIELDS = 6
JXP_EL_NFIELDS = 16
(EL_PERIOD, EL_SEMIMAJOR, EL_TIME_TRANSIT, EL_TIME_PERI, EL_ECC, EL_OMEGA, EL_SIN_OMEGA, EL_COS_OMEGA,
 EL_IMPACT_PARAM, EL_INCLINATION, EL_MASS, EL_RADIUS, EL_CENTRAL_MASS, EL_CENTRAL_RADIUS, EL_PARALLAX,
 EL_RESERVED) = range(16)

(BODY_PERIOD, BODY_TIME_TRANSIT, BODY_TIME_REF, BODY_ECC, BODY_COS_OMEGA, BODY_SIN_OMEGA,
 BODY_COS_INCL, BODY_SIN_INCL, BODY_SEMIMAJOR, BODY_CENTRAL_RADIUS, BODY_RADIUS,
 BODY_RESERVED) = range(12)


class JxpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"jaxoplanet_b200 C ABI error {code}: {msg}")
        self.code = code


_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double
_sz = C.c_size_t

SIGNATURES = {
    "jxp_version": (C.c_int, []),
    "jxp_last_error": (C.c_char_p, []),
    "jxp_set_dev_options": (C.c_int, [C.c_uint32]),
    "jxp_get_dev_options": (C.c_uint32, []),
    "jxp_kepler_f64": (C.c_int, [_p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "jxp_kepler_f32": (C.c_int, [_p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "jxp_limb_dark_f64": (C.c_int, [_p, _i32, _p, _i64, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p]),
    "jxp_limb_dark_f32": (C.c_int, [_p, _i32, _p, _i64, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p]),
    "jxp_system_workspace_bytes": (_sz, [_i32, _i32, _i64]),
    "jxp_system_light_curve_f64": (
        C.c_int,
        [_p, _i32, _i32, _p, _i32, _i64, _p, _i64, _f64, _i32, _i32, _i32, _i32, _p, _p, _p, _p,
         _i64, _p, _p, _sz, _p],
    ),
    "jxp_system_light_curve_f32": (
        C.c_int,
        [_p, _i32, _i32, _p, _i32, _i64, _p, _i64, _f64, _i32, _i32, _i32, _i32, _p, _p, _p, _p,
         _i64, _p, _p, _sz, _p],
    ),
    "jxp_system_stats": (C.c_int, [_p, _i32, _i32, _i64, _p, _p]),
    "jxp_system_list_stats": (C.c_int, [_p, _i32, _i32, _i64, _p, _p]),
    "jxp_system_walk_stats": (C.c_int, [_p, _i32, _i32, _i64, _p, _p]),
    "jxp_profile_events": (C.c_int, [_p, _p]),
    "jxp_transit_workspace_bytes": (_sz, [_i32, _i64]),
    "jxp_transit_list_stats": (C.c_int, [_p, _i32, _i64, _p, _p]),
    "jxp_transit_partials_f64": (
        C.c_int,
        [_p, _i32, _p, _i64, _p, _i64, _f64, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _i64, _p, _p,
         _p, _sz, _p],
    ),
    "jxp_transit_poly_workspace_bytes": (_sz, [_i32, _i64]),
    "jxp_transit_partials_poly_f64": (
        C.c_int,
        [_p, _i32, _i32, _p, _i64, _p, _i64, _i32, _i32, _p, _p, _p, _p, _i64, _p, _p, _p, _sz, _p],
    ),
    "jxp_transit_partials_f32": (
        C.c_int,
        [_p, _i32, _p, _i64, _p, _i64, _f64, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _i64, _p, _p,
         _p, _sz, _p],
    ),
    "jxp_transit_orbit_light_curve_f64": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p, _i32, _p, _i64, _i32, _p, _p]),
    "jxp_transit_orbit_light_curve_f32": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p, _i32, _p, _i64, _i32, _p, _p]),
    "jxp_orbital_elements_f64": (C.c_int, [_p, _i64, _p, _p]),
    "jxp_orbital_elements_soa_f64": (C.c_int, [_p, _i64, _p, _p]),
    "jxp_orbit_state_f64": (C.c_int, [_p, _i32, _p, _i64, _p, _i64, _p, _p, _i64, _i32, _p, _p, _p]),
    "jxp_orbit_state_f32": (C.c_int, [_p, _i32, _p, _i64, _p, _i64, _p, _p, _i64, _i32, _p, _p, _p]),
    "jxp_interp_periodic_f64": (C.c_int, [_p, _p, _i32, _i32, _f64, _f64, _p, _i64, _p, _p]),
    "jxp_interp_periodic_f32": (C.c_int, [_p, _p, _i32, _i32, _f64, _f64, _p, _i64, _p, _p]),
    "jxp_peak_fma_f64": (C.c_int, [_p, _i32, _i32, _i64, _p]),
    "jxp_peak_fma_f32": (C.c_int, [_p, _i32, _i32, _i64, _p]),
    "jxp_peer_buffer_bytes": (_sz, [_i64]),
    "jxp_peer_alloc": (C.c_int, [_sz, _p, _p]),
    "jxp_peer_open": (C.c_int, [_p, _p]),
    "jxp_peer_close": (C.c_int, [_p]),
    "jxp_peer_free": (C.c_int, [_p]),
    "jxp_peer_allgather_f64": (C.c_int, [_p, _i64, _i64, _i64, _p, _i32, _i32, C.c_uint64, _p, _p]),
    "jxp_peer_status": (C.c_int, [_p, _p, _p]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m jaxoplanet_b200.build` "
            "(nvcc, sm_100a). jaxoplanet_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        msg = load().jxp_last_error()
        raise JxpError(code, msg.decode() if msg else "")


def call(name: str, *args):
    check(getattr(load(), name)(*args))


class dev_options:
    """``with dev_options(JXP_DEV_NO_WALK): ...`` -- OR the mask into the library's development options
    for the duration of the block (tests compare code paths that return the same results)."""

    def __init__(self, mask: int):
        self.mask = int(mask)

    def __enter__(self):
        lib = load()
        self.old = int(lib.jxp_get_dev_options())
        lib.jxp_set_dev_options(self.old | self.mask)
        return self

    def __exit__(self, *exc):
        load().jxp_set_dev_options(self.old)
        return False
