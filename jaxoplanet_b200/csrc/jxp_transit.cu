// K6: light curve of a TransitOrbit / TTVOrbit (orbits parameterised by the transit signal).
//
// Reference path, fused into one pass over (time sample x planet):
//   TTVOrbit._warp_times            src/jaxoplanet/orbits/ttv.py:194-214   (searchsorted time warp)
//   TransitOrbit.relative_position  src/jaxoplanet/orbits/transit.py:93-107 (mod, linear motion)
//   light_curve(orbit, u)(t)        src/jaxoplanet/light_curves/limb_dark.py:49-60
//   core.limb_dark.light_curve      src/jaxoplanet/core/limb_dark.py:19-196
// Lanes are consecutive time samples, so in-transit samples (contiguous runs in time) fill whole
// warps and out-of-transit warps skip the flux code: no compaction queue is needed here.
#include "../../include/jaxoplanet_b200.h"

#include <cuda_runtime.h>

#include "jxp_host.h"
#include "jxp_math.cuh"

using namespace jxp;
using namespace jxph;

namespace {

template <typename T>
struct ToArgs {
  const double* planets;  // [n_planets][JXP_TO_NFIELDS]
  int n_planets;
  const double* bin_edges;   // concatenated over planets, or nullptr (no TTVs)
  const double* bin_values;  // concatenated, one more entry per planet than bin_edges
  const int* bin_offsets;    // [n_planets + 1] offsets into bin_edges
  const double* ttv_t0;      // [n_planets] linear reference transit time (ttv.py:206-208)
  const double* t;
  int64_t n_times;
  T g[JXP_MAX_LD_COEFFS + 1];
  int lmax;
  T* flux;  // [n_times][n_planets]
  GLTable tab;
};

// jnp.mod(a, p) for p > 0: exact fmod, result moved into [0, p)  (transit.py:102)
__device__ __forceinline__ double pymod(double a, double p) {
  double r = fmod(a, p);
  if (r != 0.0 && r < 0.0) r += p;
  return r;
}

template <typename T, int ORDER, bool HI>
__global__ void __launch_bounds__(256) k_transit_orbit(const __grid_constant__ ToArgs<T> a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_times; i += stride) {
    const double ti = a.t[i];
    for (int p = 0; p < a.n_planets; ++p) {
      const double* q = a.planets + (size_t)p * JXP_TO_NFIELDS;
      const double period = q[JXP_TO_PERIOD], t0 = q[JXP_TO_TIME_TRANSIT];
      double tw = ti;
      if (a.bin_edges) {
        // jnp.searchsorted(edges, t) (side="left"): first index with edges[idx] >= t
        const int lo0 = a.bin_offsets[p], hi0 = a.bin_offsets[p + 1];
        int lo = lo0, hi = hi0;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (a.bin_edges[mid] < ti) lo = mid + 1; else hi = mid;
        }
        const double dt = a.bin_values[(lo - lo0) + lo0 + p];  // values: one extra entry per planet
        tw = ti - (dt - a.ttv_t0[p]);
      }
      const double half = 0.5 * period;
      const double ref_time = t0 - half;
      const double dt = pymod(tw - ref_time, period) - half;
      T out = T(0);
      if (fabs(dt) < 0.5 * q[JXP_TO_DURATION]) {  // z > 0
        const double x = q[JXP_TO_SPEED] * dt;
        const double y = q[JXP_TO_IMPACT_PARAM];
        const T b = T(sqrt(x * x + y * y));  // central_radius = 1
        const T r = T(q[JXP_TO_RADIUS_RATIO]);
        T s0, s1, s2, shi[6];
        ld_solution<T, ORDER, HI>(b, r, a.lmax, &a.tab, s0, s1, s2, shi);
        T f = ld_combine(a.lmax < 2 ? a.lmax : 2, s0, s1, s2, a.g[0], a.g[1], a.g[2]);
        if (HI) {
#pragma unroll
          for (int n = 3; n <= 8; ++n)
            if (n <= a.lmax) f = jfma(shi[n - 3], a.g[n], f);
        }
        out = f;
      }
      a.flux[(size_t)i * a.n_planets + p] = out;
    }
  }
}

template <typename T>
int launch_transit_orbit(const double* planets, int32_t n_planets, const double* bin_edges,
                         const double* bin_values, const int32_t* bin_offsets, const double* ttv_t0,
                         const double* u_host, int32_t n_u, const double* t, int64_t n_times,
                         int32_t gl_order, T* flux, void* stream) {
  if (n_planets < 0 || n_times < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_orbit_light_curve: negative size");
  if (n_u < 0 || n_u > JXP_MAX_LD_COEFFS)
    return fail(JXP_ERR_UNSUPPORTED, "jxp_transit_orbit_light_curve: n_u = %d outside 0..%d", n_u, JXP_MAX_LD_COEFFS);
  if (gl_order < 1 || gl_order > JXP_MAX_GL_ORDER)
    return fail(JXP_ERR_UNSUPPORTED, "jxp_transit_orbit_light_curve: order = %d outside 1..%d", gl_order,
                JXP_MAX_GL_ORDER);
  if (n_planets == 0 || n_times == 0) return JXP_OK;
  if (!planets || !t || !flux || (n_u > 0 && !u_host))
    return fail(JXP_ERR_NULL_POINTER, "jxp_transit_orbit_light_curve: null planets/u/t/flux");
  if (bin_edges && (!bin_values || !bin_offsets || !ttv_t0))
    return fail(JXP_ERR_NULL_POINTER, "jxp_transit_orbit_light_curve: TTV tables need edges, values, offsets and t0");
  ToArgs<T> a;
  a.planets = planets;
  a.n_planets = n_planets;
  a.bin_edges = bin_edges;
  a.bin_values = bin_values;
  a.bin_offsets = bin_offsets;
  a.ttv_t0 = ttv_t0;
  a.t = t;
  a.n_times = n_times;
  a.lmax = n_u;
  a.flux = flux;
  a.tab.order = gl_order;
  if (gl_order != 10) gl_table_host(gl_order, a.tab);
  // Green's-basis coefficients on the host (limb_dark.py:42-45, 87-99): u is a host array here,
  // one law for all planets as in the reference
  {
    double g[JXP_MAX_LD_COEFFS + 1] = {0};
    const double pi = 3.141592653589793238462643383279502884;
    if (n_u == 0) {
      g[0] = 1.0 / pi;
    } else {
      const int size = n_u + 1;
      double pcoef[JXP_MAX_LD_COEFFS + 1], gg[JXP_MAX_LD_COEFFS + 5] = {0};
      for (int n = 0; n < size; ++n) {
        double arg = 0.0, binom = 1.0;
        for (int i = n; i < size; ++i) {
          const double ui = (i == 0) ? -1.0 : u_host[i - 1];
          arg += ui * binom;
          binom = binom * (double)(i + 1) / (double)(i + 1 - n);
        }
        pcoef[n] = (n % 2 == 0) ? -arg : arg;
      }
      for (int n = size - 1; n > 1; --n) gg[n] = pcoef[n] / (double)(n + 2) + gg[n + 2];
      if (size > 1) gg[1] = pcoef[1] + 3.0 * gg[3];
      gg[0] = pcoef[0] + 2.0 * gg[2];
      const double norm = pi * (gg[0] + gg[1] / 1.5);
      for (int n = 0; n < size; ++n) g[n] = gg[n] / norm;
    }
    for (int n = 0; n <= JXP_MAX_LD_COEFFS; ++n) a.g[n] = T(g[n]);
  }
  const int threads = 256;
  int64_t blocks = (n_times + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const bool hi = n_u > 2;
  if (gl_order == 10) {
    if (hi) k_transit_orbit<T, 10, true><<<(unsigned)blocks, threads, 0, st>>>(a);
    else k_transit_orbit<T, 10, false><<<(unsigned)blocks, threads, 0, st>>>(a);
  } else {
    if (hi) k_transit_orbit<T, 0, true><<<(unsigned)blocks, threads, 0, st>>>(a);
    else k_transit_orbit<T, 0, false><<<(unsigned)blocks, threads, 0, st>>>(a);
  }
  return check_launch("jxp_transit_orbit_light_curve");
}
}  // namespace

// K7: periodic linear interpolation of a pre-computed light curve, the per-time half of
// transforms.interpolate (src/jaxoplanet/light_curves/transforms.py:163-183) with the semantics of
// jnp.interp(x, xp, fp, period=period): xp is the host-prepared table (reduced mod period, sorted,
// extended by one wrapped point on each side), fp[n][n_cols] its values.
template <typename T>
__global__ void __launch_bounds__(256)
k_interp_periodic(const double* __restrict__ xp, const T* __restrict__ fp, int n, int n_cols, double period,
                  double time_transit, const double* __restrict__ t, int64_t n_times, T* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // jnp.interp's flat-segment guard: |dx| <= np.spacing(np.finfo(xp.dtype).eps), xp in float64
  const double eps = 4.930380657631324e-32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_times; i += stride) {
    // transforms.py:167-171, then the `x % period` of jnp.interp
    const double tw = pymod(t[i] - time_transit + 0.5 * period, period) + 0.5 * period + time_transit;
    const double x = pymod(tw, period);
    int lo = 0, hi = n;  // searchsorted(xp, x, side="right")
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (xp[mid] <= x) lo = mid + 1; else hi = mid;
    }
    int j = lo < 1 ? 1 : (lo > n - 1 ? n - 1 : lo);
    const double x0 = xp[j - 1], dx = xp[j] - x0, delta = x - x0;
    const bool flat = fabs(dx) <= eps;
    const double w = flat ? 0.0 : delta / dx;
    for (int c = 0; c < n_cols; ++c) {
      const T f0 = fp[(size_t)(j - 1) * n_cols + c], f1 = fp[(size_t)j * n_cols + c];
      out[(size_t)i * n_cols + c] = flat ? f0 : T(f0 + T(w) * (f1 - f0));
    }
  }
}

template <typename T>
static int launch_interp(const double* xp, const T* fp, int32_t n, int32_t n_cols, double period,
                         double time_transit, const double* t, int64_t n_times, T* out, void* stream) {
  if (n < 2 || n_cols < 1 || n_times < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_interp_periodic: bad sizes");
  if (!(period > 0.0)) return fail(JXP_ERR_BAD_SHAPE, "jxp_interp_periodic: period must be positive");
  if (n_times == 0) return JXP_OK;
  if (!xp || !fp || !t || !out) return fail(JXP_ERR_NULL_POINTER, "jxp_interp_periodic: null pointer");
  const int threads = 256;
  int64_t blocks = (n_times + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  k_interp_periodic<T><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(xp, fp, n, n_cols, period,
                                                                            time_transit, t, n_times, out);
  return check_launch("jxp_interp_periodic");
}

extern "C" int jxp_interp_periodic_f64(const double* xp, const double* fp, int32_t n, int32_t n_cols,
                                       double period, double time_transit, const double* t,
                                       int64_t n_times, double* out, void* stream) {
  return launch_interp<double>(xp, fp, n, n_cols, period, time_transit, t, n_times, out, stream);
}
extern "C" int jxp_interp_periodic_f32(const double* xp, const float* fp, int32_t n, int32_t n_cols,
                                       double period, double time_transit, const double* t,
                                       int64_t n_times, float* out, void* stream) {
  return launch_interp<float>(xp, fp, n, n_cols, period, time_transit, t, n_times, out, stream);
}

// K0: OrbitalBody.__init__ on the device -- the per-set derived constants of
// src/jaxoplanet/orbits/keplerian.py:262-340 (and Central's two-of-three rule :30-70 is the caller's:
// central mass and radius come in resolved).  One thread per (set, body); NaN marks "not given"
// (None in the reference).  Same operation order as the reference; libm calls are CUDA's (<= 1-2 ulp
// from XLA's / NumPy's).
// SOA: the element records come field-major, [JXP_EL_NFIELDS][n] (what a host packs with contiguous copies)
template <bool SOA>
__global__ void __launch_bounds__(128)
k_orbital_elements(const double* __restrict__ el, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double p[JXP_EL_NFIELDS];
#pragma unroll
  for (int k = 0; k < JXP_EL_NFIELDS; ++k) p[k] = SOA ? el[(int64_t)k * n + i] : el[i * JXP_EL_NFIELDS + k];
  const double pi = 3.141592653589793238462643383279502884;
  const double G = 2942.2062175044193;  // src/jaxoplanet/constants.py:2
  const double central_mass = p[JXP_EL_CENTRAL_MASS], central_radius = p[JXP_EL_CENTRAL_RADIUS];
  const double mass = p[JXP_EL_MASS];
  const double total_mass = isnan(mass) ? central_mass : mass + central_mass;  // :265-268
  const double mass_factor = G * total_mass;
  double period = p[JXP_EL_PERIOD], semimajor = p[JXP_EL_SEMIMAJOR];
  if (isnan(semimajor)) {
    semimajor = cbrt(mass_factor * (period * period) / (4.0 * (pi * pi)));  // :275
  } else {
    period = 2.0 * pi * semimajor * sqrt(semimajor / mass_factor);  // :281
  }
  double sw = p[JXP_EL_SIN_OMEGA], cw = p[JXP_EL_COS_OMEGA];
  if (!isnan(p[JXP_EL_OMEGA])) {
    sw = sin(p[JXP_EL_OMEGA]);  // :286-287
    cw = cos(p[JXP_EL_OMEGA]);
  }
  const double ecc = p[JXP_EL_ECC];
  const bool circular = isnan(ecc);
  double M0, incl_factor;
  if (circular) {
    M0 = 0.5 * pi;  // :302
    incl_factor = 1.0;
  } else {
    const double opsw = 1.0 + sw;
    const double E0 = 2.0 * atan2(sqrt(1.0 - ecc) * cw, sqrt(1.0 + ecc) * opsw);  // :308-311
    M0 = E0 - ecc * sin(E0);                                                       // :312
    const double ome2 = 1.0 - ecc * ecc;
    incl_factor = (1.0 + ecc * sw) / ome2;  // :315
  }
  const double dcosidb = incl_factor * central_radius / semimajor;  // :318
  double cos_i, sin_i;
  if (!isnan(p[JXP_EL_IMPACT_PARAM])) {
    cos_i = dcosidb * p[JXP_EL_IMPACT_PARAM];  // :321-322
    sin_i = sqrt(1.0 - cos_i * cos_i);
  } else if (!isnan(p[JXP_EL_INCLINATION])) {
    cos_i = cos(p[JXP_EL_INCLINATION]);  // :325-327
    sin_i = sin(p[JXP_EL_INCLINATION]);
  } else {
    cos_i = 0.0;  // :329-331
    sin_i = 1.0;
  }
  const double time_ref = -M0 * period / (2.0 * pi);  // :334
  double time_transit = p[JXP_EL_TIME_TRANSIT];
  if (isnan(time_transit)) time_transit = isnan(p[JXP_EL_TIME_PERI]) ? 0.0 : p[JXP_EL_TIME_PERI] - time_ref;  // :336-340
  double a_out = semimajor;
  if (!isnan(p[JXP_EL_PARALLAX])) a_out = semimajor * p[JXP_EL_PARALLAX] / 215.03215567054764;  // :620-625 (au in R_sun)
  double* o = out + i * JXP_BODY_NFIELDS;
  o[JXP_BODY_PERIOD] = period;
  o[JXP_BODY_TIME_TRANSIT] = time_transit;
  o[JXP_BODY_TIME_REF] = time_ref;
  o[JXP_BODY_ECC] = circular ? -1.0 : ecc;
  o[JXP_BODY_COS_OMEGA] = circular ? 1.0 : cw;
  o[JXP_BODY_SIN_OMEGA] = circular ? 0.0 : sw;
  o[JXP_BODY_COS_INCL] = cos_i;
  o[JXP_BODY_SIN_INCL] = sin_i;
  o[JXP_BODY_SEMIMAJOR] = a_out;
  o[JXP_BODY_CENTRAL_RADIUS] = central_radius;
  o[JXP_BODY_RADIUS] = isnan(p[JXP_EL_RADIUS]) ? 0.0 : p[JXP_EL_RADIUS];
  o[JXP_BODY_RESERVED] = 0.0;
}

extern "C" int jxp_orbital_elements_f64(const double* elements, int64_t n, double* bodies, void* stream) {
  if (n < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_orbital_elements: n = %lld < 0", (long long)n);
  if (n == 0) return JXP_OK;
  if (!elements || !bodies) return fail(JXP_ERR_NULL_POINTER, "jxp_orbital_elements: null pointer");
  k_orbital_elements<false><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(elements, n, bodies);
  return check_launch("jxp_orbital_elements");
}

extern "C" int jxp_orbital_elements_soa_f64(const double* elements, int64_t n, double* bodies, void* stream) {
  if (n < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_orbital_elements_soa: n = %lld < 0", (long long)n);
  if (n == 0) return JXP_OK;
  if (!elements || !bodies) return fail(JXP_ERR_NULL_POINTER, "jxp_orbital_elements_soa: null pointer");
  k_orbital_elements<true><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(elements, n, bodies);
  return check_launch("jxp_orbital_elements_soa");
}

extern "C" int jxp_transit_orbit_light_curve_f64(const double* planets, int32_t n_planets,
                                                 const double* bin_edges, const double* bin_values,
                                                 const int32_t* bin_offsets, const double* ttv_t0,
                                                 const double* u_host, int32_t n_u, const double* t,
                                                 int64_t n_times, int32_t gl_order, double* flux,
                                                 void* stream) {
  return launch_transit_orbit<double>(planets, n_planets, bin_edges, bin_values, bin_offsets, ttv_t0,
                                      u_host, n_u, t, n_times, gl_order, flux, stream);
}
extern "C" int jxp_transit_orbit_light_curve_f32(const double* planets, int32_t n_planets,
                                                 const double* bin_edges, const double* bin_values,
                                                 const int32_t* bin_offsets, const double* ttv_t0,
                                                 const double* u_host, int32_t n_u, const double* t,
                                                 int64_t n_times, int32_t gl_order, float* flux,
                                                 void* stream) {
  return launch_transit_orbit<float>(planets, n_planets, bin_edges, bin_values, bin_offsets, ttv_t0,
                                     u_host, n_u, t, n_times, gl_order, flux, stream);
}
