/*
 * jaxoplanet_b200 -- C ABI of the B200 (sm_100a) kernels for jaxoplanet's hot path.
 *
 * The reference (exoplanet-dev/jaxoplanet) is pure Python/JAX and has no native
 * interface; the boundary a drop-in has to honour is its Python API.  Each entry
 * point below names the reference callable it stands behind (paths relative to the
 * reference checkout).  The Python facade in jaxoplanet_b200/ keeps the reference
 * signatures and reaches these symbols through ctypes; an XLA FFI handler is a thin
 * argument-marshalling wrapper over the same symbols.
 *
 * Conventions
 *  - plain pointers and sizes only; every buffer is DEVICE memory owned by the caller
 *    unless the argument comment says "host".  The library never allocates or frees
 *    caller-visible memory and never synchronises the stream.
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *    Launches go to the calling thread's current CUDA device.  All entry points are
 *    re-entrant and CUDA-graph capturable.
 *  - return value: 0 (JXP_OK) or a negative JxpStatus; the message for the calling
 *    thread's last failure is available from jxp_last_error().
 *  - strides are in ELEMENTS and must be 0 (broadcast one value) or 1 (dense).
 *  - `_f64` computes in IEEE double; `_f32` computes the solve and the flux in float
 *    (times and the phase reduction stay double, see DESIGN.md "fp32 mode").
 */
#ifndef JAXOPLANET_B200_H_
#define JAXOPLANET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JXP_ABI_VERSION 1

typedef enum JxpStatus {
  JXP_OK = 0,
  JXP_ERR_NULL_POINTER = -1,
  JXP_ERR_BAD_SHAPE = -2,
  JXP_ERR_UNSUPPORTED = -3,
  JXP_ERR_CUDA = -4,
  JXP_ERR_WORKSPACE = -5
} JxpStatus;

int jxp_version(void);
/* Thread-local, NUL-terminated, valid until the next failing call on this thread. */
const char* jxp_last_error(void);

/* ------------------------------------------------------------------------------------
 * K1  jaxoplanet.core.kepler.kepler(M, ecc) -> (sin f, cos f)
 *     src/jaxoplanet/core/kepler.py:14-56 (primal), :59-79 (custom JVP)
 * M, ecc broadcast against each other (stride 0 or 1) to n elements.
 * Optional outputs (NULL to skip):
 *   E      the eccentric anomaly in [0, 2pi) the reference keeps internal (kepler.py:39-43)
 *   dfdM   (1 + e cos f)^2 / (1 - e^2)^{3/2}          } the two factors of the reference's
 *   dfde   (2 + e cos f) sin f / (1 - e^2)             } JVP rule: f_dot = M_dot dfdM + e_dot dfde,
 *                                                        d sin f = cos f f_dot, d cos f = -sin f f_dot
 * ---------------------------------------------------------------------------------- */
int jxp_kepler_f64(const double* M, int64_t M_stride, const double* ecc, int64_t ecc_stride,
                   int64_t n, double* sinf, double* cosf, double* E, double* dfdM,
                   double* dfde, void* stream);
int jxp_kepler_f32(const float* M, int64_t M_stride, const float* ecc, int64_t ecc_stride,
                   int64_t n, float* sinf, float* cosf, float* E, float* dfdM, float* dfde,
                   void* stream);

/* ------------------------------------------------------------------------------------
 * K2  jaxoplanet.core.limb_dark.light_curve(u, b, r, order=10) -> flux - 1
 *     src/jaxoplanet/core/limb_dark.py:19-47 (+ :50-196)
 * u: n_u coefficients (0 <= n_u <= JXP_MAX_LD_COEFFS), one shared set as in the
 * reference (u.ndim == 1, limb_dark.py:40).  b, r broadcast (stride 0 or 1) to n.
 * order: Gauss-Legendre nodes of the line integral (1..JXP_MAX_GL_ORDER).
 * Optional outputs (NULL to skip):
 *   dflux_db, dflux_dr   derivatives of the *discretised* formula, as JAX autodiff of the
 *                        reference gives them (incl. the zero_safe_sqrt rule utils.py:16-24)
 *   svec                 [n_u + 1][n] solution vector s (limb_dark.py:73-82), so that the
 *                        caller forms d flux/d u = s . dg/du
 * ---------------------------------------------------------------------------------- */
#define JXP_MAX_LD_COEFFS 8
#define JXP_MAX_GL_ORDER 64
int jxp_limb_dark_f64(const double* u, int32_t n_u, const double* b, int64_t b_stride,
                      const double* r, int64_t r_stride, int64_t n, int32_t order,
                      double* flux_m1, double* dflux_db, double* dflux_dr, double* svec,
                      void* stream);
int jxp_limb_dark_f32(const float* u, int32_t n_u, const float* b, int64_t b_stride,
                      const float* r, int64_t r_stride, int64_t n, int32_t order,
                      float* flux_m1, float* dflux_db, float* dflux_dr, float* svec,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * K3/K4  jaxoplanet.light_curves.limb_dark.light_curve(System, u)(t), optionally wrapped in
 *        light_curves.transforms.integrate(..., exposure_time, order, num_samples), for a
 *        BATCH of parameter sets (what the reference user writes as jax.vmap over
 *        parameters), optionally fused with the Gaussian log-likelihood of observed flux.
 *     src/jaxoplanet/light_curves/limb_dark.py:15-62
 *     src/jaxoplanet/orbits/keplerian.py:402-419, 537-637   (relative_position)
 *     src/jaxoplanet/light_curves/transforms.py:21-116      (integrate)
 *
 * bodies: [n_sets][n_bodies][JXP_BODY_NFIELDS] derived per-body constants, i.e. the
 *         fields OrbitalBody.__init__ computes (keplerian.py:262-340); ALWAYS double (a float
 *         period or epoch cannot place a transit to 1 ppm of flux).
 * u:      [n_sets or 1][n_u] limb-darkening coefficients (u_set_stride = n_u or 0),
 *         0 <= n_u <= JXP_MAX_LD_COEFFS (general polynomial law), ALWAYS double.
 * t:      [n_times] time stamps, ALWAYS double, shared by all sets.
 * exposure_time <= 0 or num_samples <= 0: no exposure integration.  exp_order in {0,1,2}.
 * Outputs (any may be NULL):
 *   flux      [n_sets][n_times][n_bodies]  per-body flux - 1, masked to 0 where z <= 0
 *             (the reference's return value, with a leading set axis)
 *   flux_sum  [n_sets][n_times]            sum over bodies of the above
 *   loglike   [n_sets]   sum_t Normal(1 + flux_sum, sigma).log_prob(y); needs y, sigma and a
 *             workspace of jxp_system_workspace_bytes(...) bytes
 * flags: JXP_FLAG_* below.
 * ---------------------------------------------------------------------------------- */
enum {
  JXP_BODY_PERIOD = 0,
  JXP_BODY_TIME_TRANSIT = 1,
  JXP_BODY_TIME_REF = 2,
  JXP_BODY_ECC = 3, /* < 0: circular body (eccentricity is None, keplerian.py:299-303) */
  JXP_BODY_COS_OMEGA = 4,
  JXP_BODY_SIN_OMEGA = 5,
  JXP_BODY_COS_INCL = 6,
  JXP_BODY_SIN_INCL = 7,
  JXP_BODY_SEMIMAJOR = 8,
  JXP_BODY_CENTRAL_RADIUS = 9,
  JXP_BODY_RADIUS = 10,
  JXP_BODY_RESERVED = 11,
  JXP_BODY_NFIELDS = 12
};

/* Evaluate every point in full (no transit-window pre-test).  Results are identical
 * either way; this exists for measurement (DESIGN.md "window test"). */
#define JXP_FLAG_NO_WINDOW 1

/* Bytes of caller-provided device workspace (256-byte aligned) for one call of that shape: the
 * derived records, 1 / sigma, partial sums and, for single-body systems, the tables of the transit
 * walk: a 256-byte record and a few dozen 8-byte (first position, first sample) segments
 * per set, a 16-byte entry and an 8-byte partial sum per chunk of in-window samples -- nothing per
 * sample.  Sized for 1/8 of the grid in transit and 384 transits per set on average; when that does
 * not fit, or the time stamps are not sorted, the same call falls back to the window-test kernels
 * on the device -- results agree to rounding either way. */
size_t jxp_system_workspace_bytes(int32_t n_sets, int32_t n_bodies, int64_t n_times);

int jxp_system_light_curve_f64(const double* bodies, int32_t n_sets, int32_t n_bodies,
                               const double* u, int32_t n_u, int64_t u_set_stride,
                               const double* t, int64_t n_times, double exposure_time,
                               int32_t exp_order, int32_t num_samples, int32_t gl_order,
                               int32_t flags, double* flux, double* flux_sum,
                               const double* y, const double* sigma, int64_t sigma_stride,
                               double* loglike, void* workspace, size_t workspace_bytes,
                               void* stream);
int jxp_system_light_curve_f32(const double* bodies, int32_t n_sets, int32_t n_bodies,
                               const double* u, int32_t n_u, int64_t u_set_stride,
                               const double* t, int64_t n_times, double exposure_time,
                               int32_t exp_order, int32_t num_samples, int32_t gl_order,
                               int32_t flags, float* flux, float* flux_sum, const float* y,
                               const float* sigma, int64_t sigma_stride, float* loglike,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Development / measurement options, process-wide (one atomic word; the library never reads the
 * environment).  They select between code paths that return the same results; tests use them to
 * compare the paths with each other.  Default 0. */
#define JXP_DEV_NO_WALK 1u        /* single-body calls: window-test kernels instead of the transit walk */
#define JXP_DEV_NO_PAIR 2u        /* log-likelihood: generic window-test kernel instead of k_system_pair */
#define JXP_DEV_WALK_ANY_SIZE 4u  /* flux outputs: no minimum problem size for the transit walk */
#define JXP_DEV_NO_LD_CLASSES 8u  /* jxp_limb_dark_f64: plain kernel instead of the class-sorted one */
#define JXP_DEV_PROFILE_PLAN 16u  /* jxp_profile_events brackets the plan launch instead of the evaluation */
#define JXP_DEV_PROFILE_TABLES 64u /* jxp_profile_events brackets the table launch (k_walk_tables) */
#define JXP_DEV_PROFILE_ITEMS 128u /* jxp_profile_events brackets the table-set evaluation (k_walk_items) */
#define JXP_DEV_NO_TABLES 32u     /* transit walk: every sample by the direct Kepler + quadrature code (no per-set tables) */
#define JXP_DEV_NO_BULK_FILL 4096u /* jxp_transit_partials: zero rows by warp stores instead of bulk copies from shared memory */
#define JXP_DEV_BULK_FILL_SWAP 8192u /* jxp_transit_partials: the other bulk-fill mode (f64: every thread issues, f32: one warp issues) */
#define JXP_DEV_TR_TPW(n) (((uint32_t)(n) & 0x3u) << 14)      /* jxp_transit_partials, bulk fill: warp tiles per warp (1..2) */
#define JXP_DEV_GET_TR_TPW(m) (((m) >> 14) & 0x3u)
#define JXP_DEV_CTA_WARPS(n) (((uint32_t)(n) & 0xfu) << 8)   /* window-test kernels: warps per CTA (1..4) */
#define JXP_DEV_GET_CTA_WARPS(m) (((m) >> 8) & 0xfu)
#define JXP_DEV_TR_RUN(n) (((uint32_t)(n) & 0x7fu) << 16)    /* jxp_transit_partials: warp tiles per run */
#define JXP_DEV_GET_TR_RUN(m) (((m) >> 16) & 0x7fu)
int jxp_set_dev_options(uint32_t mask);
uint32_t jxp_get_dev_options(void);

/* Measurement helper: when both events (cudaEvent_t as void*) are non-NULL, the calling
 * thread's subsequent jxp_system_light_curve_* calls record them on the launch stream
 * immediately before and after the dominant kernel (k_system), so that a harness can time that
 * kernel alone with CUDA events.  Pass NULL, NULL to switch off (the default). */
int jxp_profile_events(void* start_event, void* stop_event);

/* Measurement helper: how much work the LAST jxp_system_light_curve_* call on this workspace
 * actually executed -- stats_host[0] = Kepler solves, stats_host[1] = limb-darkening
 * evaluations (host memory, 2 x int64).  Copies device -> host and SYNCHRONISES `stream`;
 * bench.py uses it so that the roofline counts executed work, not skipped samples. */
int jxp_system_stats(const void* workspace, int32_t n_sets, int32_t n_bodies, int64_t n_times,
                     int64_t* stats_host, void* stream);

/* Measurement helper: which kernels the LAST single-body call on this workspace used.  With
 * sorted time stamps, one body and no exposure integration only the in-window samples are visited
 * (DESIGN.md "transit walk"); the choice is made on the device.
 * stats_host (4 x int64): [0] = in-window samples visited, [1] = 0 when the walk produced the
 * result (and [0] > 0), bit 0: time stamps not sorted, bit 1: the tables did not fit the workspace --
 * in both cases the window-test kernels produced it; [2] = samples on the sub-list estimated to have
 * the planet fully inside the stellar disk, [3] = in-transit samples that passed the b < 1 - r check
 * and took the constant-kappa flux code.  SYNCHRONISES `stream`. */
int jxp_system_list_stats(const void* workspace, int32_t n_sets, int32_t n_bodies,
                          int64_t n_times, int64_t* stats_host, void* stream);

/* Measurement helper: the per-set function tables of the transit walk (csrc/jxp_tables.cuh) in the LAST call on
 * this workspace.  stats_host[0] = flux evaluations answered by a table, [1] = samples of table sets that took the
 * direct code instead (k_walk_fix), [2] = (set, body) units whose tables were used, [3] = transiting units without tables.  Evaluations not counted in [0] took the direct Kepler + quadrature
 * code (jxp_system_stats counts all of them).  SYNCHRONISES `stream`. */
int jxp_system_walk_stats(const void* workspace, int32_t n_sets, int32_t n_bodies, int64_t n_times,
                          int64_t* stats_host, void* stream);

/* ------------------------------------------------------------------------------------
 * K5  flux and its forward-mode partials w.r.t. the 8 sampled parameters of a single
 *     transiting planet, theta = (period, time_transit, eccentricity, omega_peri,
 *     impact_param, radius, u1, u2): what jax.jvp / jax.grad of
 *         lambda theta: limb_dark_light_curve(System(Central(mass, radius)).add_body(
 *             period=, time_transit=, eccentricity=, omega_peri=, impact_param=, radius=),
 *             [u1, u2])(t)[..., 0]
 *     evaluates through keplerian.py:262-340 (constants), kepler.py:59-79 (Kepler JVP),
 *     limb_dark.py (autodiff of the discretised flux) -- SURVEY.md 3.5.
 * theta:   [n_sets][8];  star: [n_sets or 1][2] = (central mass, central radius),
 *          star_set_stride = 2 or 0.
 * Outputs (any may be NULL):
 *   flux      [n_sets][n_times]      flux - 1
 *   partials  [n_sets][8][n_times]   d(flux)/d(theta_k)
 *   loglike   [n_sets], dloglike [n_sets][8]   Gaussian log-likelihood of (y, sigma) and its
 *             gradient, reduced in-kernel (workspace: jxp_transit_workspace_bytes).  With flux
 *             and partials NULL, no exposure integration and sorted time stamps the in-window
 *             samples go on a work list in the workspace (80 bytes per item, room for 1/16 of
 *             the grid; DESIGN.md "gradient step on the transit list"); otherwise, or when the
 *             list does not fit, the window-test kernel does the same work -- decided on the device.
 * ---------------------------------------------------------------------------------- */
#define JXP_NTHETA 8
size_t jxp_transit_workspace_bytes(int32_t n_sets, int64_t n_times);
/* Measurement helper, the K5 counterpart of jxp_system_list_stats (same 4 x int64 layout; [3] = 0):
 * which kernels the LAST gradient-step call on this workspace used.  SYNCHRONISES `stream`. */
int jxp_transit_list_stats(const void* workspace, int32_t n_sets, int64_t n_times,
                           int64_t* stats_host, void* stream);

int jxp_transit_partials_f64(const double* theta, int32_t n_sets, const double* star,
                             int64_t star_set_stride, const double* t, int64_t n_times,
                             double exposure_time, int32_t exp_order, int32_t num_samples,
                             int32_t gl_order, int32_t flags, double* flux, double* partials,
                             const double* y, const double* sigma, int64_t sigma_stride,
                             double* loglike, double* dloglike, void* workspace,
                             size_t workspace_bytes, void* stream);
int jxp_transit_partials_f32(const double* theta, int32_t n_sets, const double* star,
                             int64_t star_set_stride, const double* t, int64_t n_times,
                             double exposure_time, int32_t exp_order, int32_t num_samples,
                             int32_t gl_order, int32_t flags, float* flux, float* partials,
                             const float* y, const float* sigma, int64_t sigma_stride,
                             float* loglike, float* dloglike, void* workspace,
                             size_t workspace_bytes, void* stream);

/* K5 for general polynomial limb darkening (any n_u <= JXP_MAX_LD_COEFFS coefficients): what jax.jvp / jax.grad of
 *   limb_dark_light_curve(System(...).add_body(...), u)(t)[..., 0]   (light_curves/limb_dark.py:15-62)
 * evaluate for a law of any length -- the reference differentiates core.limb_dark.light_curve for any u
 * (core/limb_dark.py:19-47, greens_basis_transform :87-99, p_integral's l >= 3 terms :149-153, 175-182).
 * theta:    [n_sets][6 + n_u] (period, time_transit, eccentricity, omega_peri, impact_param, radius, u_1 .. u_n)
 * partials: [n_sets][6 + n_u][n_times]; dloglike: [n_sets][6 + n_u].  Double precision, order 10, no exposure
 * integration; a plain one-sample-per-thread kernel (the quadratic entry points above are the tuned path). */
size_t jxp_transit_poly_workspace_bytes(int32_t n_sets, int64_t n_times);
int jxp_transit_partials_poly_f64(const double* theta, int32_t n_sets, int32_t n_u, const double* star,
                                  int64_t star_set_stride, const double* t, int64_t n_times,
                                  int32_t gl_order, int32_t flags, double* flux, double* partials,
                                  const double* y, const double* sigma, int64_t sigma_stride,
                                  double* loglike, double* dloglike, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K8  Positions and velocities of Keplerian bodies.
 *   reference: OrbitalBody.position / central_position / relative_position / velocity /
 *     central_velocity / relative_velocity / radial_velocity
 *     src/jaxoplanet/orbits/keplerian.py:366-532 over _get_true_anomaly :537-541,
 *     _get_position_and_velocity :590-637 and _rotate_vector :543-588
 *
 * bodies:  [n_sets][JXP_BODY_NFIELDS] as for jxp_system_light_curve_* (the semimajor, radius and
 *          central radius fields are not used here), ALWAYS double.
 * r_scale: [n_sets] (stride 1) or one value (stride 0): the factor of the orbital radius -- the
 *          `semimajor` argument of _get_position_and_velocity (keplerian.py:620-625, parallax included).
 * k_amp:   [n_sets] or one value: the velocity amplitude k0 (keplerian.py:604-618).
 * asc_node: [n_sets][2] cos and sin of the ascending node, or NULL (keplerian.py:578-588).
 * flags:   JXP_STATE_NO_INCLINATION: velocities rotated without the inclination (a given
 *          semi-amplitude, keplerian.py:634 include_inclination=False).
 * pos, vel: [3][n_sets][n_times] (x, y, z planes) or NULL; time stamps ALWAYS double.
 * ---------------------------------------------------------------------------------- */
#define JXP_STATE_NO_INCLINATION 1
int jxp_orbit_state_f64(const double* bodies, int32_t n_sets, const double* r_scale, int64_t r_stride,
                        const double* k_amp, int64_t k_stride, const double* asc_node, const double* t,
                        int64_t n_times, int32_t flags, double* pos, double* vel, void* stream);
int jxp_orbit_state_f32(const double* bodies, int32_t n_sets, const double* r_scale, int64_t r_stride,
                        const double* k_amp, int64_t k_stride, const double* asc_node, const double* t,
                        int64_t n_times, int32_t flags, float* pos, float* vel, void* stream);

/* ------------------------------------------------------------------------------------
 * K0  OrbitalBody.__init__(central, body): the per-set derived constants
 *     src/jaxoplanet/orbits/keplerian.py:262-340
 * elements: [n][JXP_EL_NFIELDS] (device) the sampled Body arguments of one body per row, NaN =
 * "not given" (None in the reference); exactly one of period / semimajor, at most one of
 * impact_param / inclination and of time_transit / time_peri, omega as an angle or as sin and cos,
 * eccentricity NaN = circular.  central_mass / central_radius: the resolved Central (:30-70).
 * bodies: [n][JXP_BODY_NFIELDS] (device) the records jxp_system_light_curve_* takes.
 * ---------------------------------------------------------------------------------- */
enum {
  JXP_EL_PERIOD = 0,
  JXP_EL_SEMIMAJOR = 1,
  JXP_EL_TIME_TRANSIT = 2,
  JXP_EL_TIME_PERI = 3,
  JXP_EL_ECC = 4,
  JXP_EL_OMEGA = 5,
  JXP_EL_SIN_OMEGA = 6,
  JXP_EL_COS_OMEGA = 7,
  JXP_EL_IMPACT_PARAM = 8,
  JXP_EL_INCLINATION = 9,
  JXP_EL_MASS = 10,
  JXP_EL_RADIUS = 11,
  JXP_EL_CENTRAL_MASS = 12,
  JXP_EL_CENTRAL_RADIUS = 13,
  JXP_EL_PARALLAX = 14,
  JXP_EL_RESERVED = 15,
  JXP_EL_NFIELDS = 16
};
int jxp_orbital_elements_f64(const double* elements, int64_t n, double* bodies, void* stream);
/* The same with the element records field-major, elements[JXP_EL_NFIELDS][n]: a host that holds one array
 * per sampled parameter (a dict of (n,) arrays) packs them with contiguous copies. */
int jxp_orbital_elements_soa_f64(const double* elements, int64_t n, double* bodies, void* stream);

/* ------------------------------------------------------------------------------------
 * K6  light_curves.limb_dark.light_curve(TransitOrbit | TTVOrbit, u)(t)
 *     src/jaxoplanet/orbits/transit.py:7-107 (linear motion across the disk, no Kepler solve),
 *     src/jaxoplanet/orbits/ttv.py:46-225 (searchsorted time warp around each observed transit),
 *     src/jaxoplanet/light_curves/limb_dark.py:49-60, src/jaxoplanet/core/limb_dark.py:19-196
 * planets: [n_planets][JXP_TO_NFIELDS] (device) -- the fields TransitOrbit.__init__ derives.
 * TTV tables (device; bin_edges == NULL means a plain TransitOrbit): for planet p the edges
 * bin_edges[bin_offsets[p] .. bin_offsets[p+1]) and the values
 * bin_values[bin_offsets[p] + p .. bin_offsets[p+1] + p + 1) (one more value than edges) are the
 * `_bin_edges` / `_bin_values` of ttv.py:21-33; ttv_t0[p] is the linear reference time (`t0`).
 * u_host: n_u limb-darkening coefficients in HOST memory (one law for all planets, as in the
 * reference).  flux: [n_times][n_planets] (t.shape + orbit.shape), flux - 1 per planet.
 * ---------------------------------------------------------------------------------- */
enum {
  JXP_TO_PERIOD = 0,
  JXP_TO_TIME_TRANSIT = 1,
  JXP_TO_SPEED = 2,
  JXP_TO_DURATION = 3,
  JXP_TO_IMPACT_PARAM = 4,
  JXP_TO_RADIUS_RATIO = 5,
  JXP_TO_NFIELDS = 6
};
int jxp_transit_orbit_light_curve_f64(const double* planets, int32_t n_planets,
                                      const double* bin_edges, const double* bin_values,
                                      const int32_t* bin_offsets, const double* ttv_t0,
                                      const double* u_host, int32_t n_u, const double* t,
                                      int64_t n_times, int32_t gl_order, double* flux, void* stream);
int jxp_transit_orbit_light_curve_f32(const double* planets, int32_t n_planets,
                                      const double* bin_edges, const double* bin_values,
                                      const int32_t* bin_offsets, const double* ttv_t0,
                                      const double* u_host, int32_t n_u, const double* t,
                                      int64_t n_times, int32_t gl_order, float* flux, void* stream);

/* ------------------------------------------------------------------------------------
 * K7  the per-time half of transforms.interpolate(func, period=, time_transit=, num_samples=,
 *     duration=)(t): src/jaxoplanet/light_curves/transforms.py:163-183, i.e.
 *     jnp.interp(wrapped(t), time_grid, flux_grid, period=period).
 * xp: [n] the time grid reduced mod period, sorted and extended by one wrapped point on each
 * side (what jnp.interp builds internally); fp: [n][n_cols] the pre-computed flux at those
 * points (n_cols = number of bodies); out: [n_times][n_cols].  All device pointers.
 * ---------------------------------------------------------------------------------- */
int jxp_interp_periodic_f64(const double* xp, const double* fp, int32_t n, int32_t n_cols,
                            double period, double time_transit, const double* t, int64_t n_times,
                            double* out, void* stream);
int jxp_interp_periodic_f32(const double* xp, const float* fp, int32_t n, int32_t n_cols,
                            double period, double time_transit, const double* t, int64_t n_times,
                            float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Measurement helper (not part of the reference surface): a dependent-free DFMA (or FFMA)
 * loop that saturates the FP64 (FP32) pipe; bench.py times it with CUDA events to obtain the
 * FP64/FP32 roofline denominator MEASURED_PEAKS.json lacks.  Executes
 * 2 * n_threads * iters * 16 flops and writes one value per thread to `sink`.
 * ---------------------------------------------------------------------------------- */
int jxp_peak_fma_f64(double* sink, int32_t n_blocks, int32_t n_threads, int64_t iters,
                     void* stream);
int jxp_peak_fma_f32(float* sink, int32_t n_blocks, int32_t n_threads, int64_t iters,
                     void* stream);

/* ----------------------------------------------------------------------------------
 * Peer-memory all-gather of per-set values between the ranks of one box (one process per GPU):
 * the exchange step of the set-sharded path (the reference has no multi-device
 * code -- under jax.pmap / sharding this is the all_gather XLA would insert after
 * src/jaxoplanet/light_curves/limb_dark.py:62).  8 bytes per parameter set is pure latency for a
 * collective library, so each rank stores its block into every rank's exchange buffer over
 * NVLink and signals with an epoch flag (csrc/jxp_peer.cu).
 *
 *   jxp_peer_buffer_bytes(n_total)       size of one rank's exchange buffer
 *   jxp_peer_alloc(bytes, &ptr, handle)  cudaMalloc + zero + 64-byte CUDA IPC handle to pass to the peers
 *   jxp_peer_open(handle, &ptr)          map a peer's buffer (cudaIpcOpenMemHandle, peer access enabled)
 *   jxp_peer_close / jxp_peer_free
 *   jxp_peer_allgather_f64(src, n_local, offset, n_total, bufs[world], rank, world, epoch, out, stream)
 *       src[0 .. n_local) -> element offset .. of every rank's buffer; returns (stream-ordered) when the
 *       blocks of all ranks for this epoch have arrived in bufs[rank].  Epochs count up from 1, the same
 *       on every rank; the gathered values of epoch e sit at bufs[rank] + 256 + (e & 1) * n_total * 8.
 *       epoch = 0 (on every rank): the kernel counts the epochs itself in its buffer, so the launch carries
 *       no argument that changes from call to call and can be replayed from a CUDA graph.
 *       out (optional): the n_total gathered values are also copied there by the same launch.
 *   jxp_peer_status(own_buf, status[2], stream)  status[0] = 0, or which peer a wait timed out on (5 s);
 *       status[1] = the last epoch this rank has started
 * ---------------------------------------------------------------------------------- */
size_t jxp_peer_buffer_bytes(int64_t n_total);
int jxp_peer_alloc(size_t bytes, void** ptr, void* handle_out);
int jxp_peer_open(const void* handle, void** ptr);
int jxp_peer_close(void* ptr);
int jxp_peer_free(void* ptr);
int jxp_peer_allgather_f64(const double* src, int64_t n_local, int64_t offset, int64_t n_total,
                           const uint64_t* bufs, int32_t rank, int32_t world, uint64_t epoch, double* out,
                           void* stream);
int jxp_peer_status(const void* own_buf, uint64_t* status_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JAXOPLANET_B200_H_ */
