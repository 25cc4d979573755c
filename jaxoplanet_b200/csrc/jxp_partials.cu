// K5: flux and forward-mode partials w.r.t. theta = (period, time_transit, eccentricity,
// omega_peri, impact_param, radius, u1, u2) of a single transiting planet, optionally with
// the Gaussian log-likelihood and its gradient reduced in-kernel.
//
// The partials are those of the reference program differentiated by JAX:
//   OrbitalBody.__init__            src/jaxoplanet/orbits/keplerian.py:262-340  (dual numbers
//                                   over the same formulas, prep kernel, once per set)
//   M = 2 pi (t - t0 - t_ref) / P   keplerian.py:538                            (dual)
//   kepler custom JVP               src/jaxoplanet/core/kepler.py:59-79         (closed form)
//   position / b                    keplerian.py:590-637, light_curves/limb_dark.py:53 (dual)
//   flux(b, r)                      src/jaxoplanet/core/limb_dark.py            (dual over the
//                                   discretised formulas incl. zero_safe_sqrt's rule)
//   d flux / d u                    s . dg/du                                   (limb_dark.py:45-47)
#include "../../include/jaxoplanet_b200.h"

#include <cuda_runtime.h>

#include "jxp_host.h"
#include "jxp_orbit.cuh"

using namespace jxp;
using namespace jxph;

namespace {
constexpr int NO = 5;  // orbit directions: P, t0, e, omega, b
constexpr int TR_K = 2;
constexpr int TR_MAX_SETS = 16;

template <typename T>
struct BodyD {
  BodyC<T> c;                 // primal record (+ window), shared with the forward kernels
  Dual<double, NO> tref;      // d t_ref / d(P, e, omega)
  Dual<T, NO> cw, sw, ci, si, na_ome2;
  T inv_ome2;                 // 1 / (1 - e^2)
  T inv_ome2_32;              // (1 - e^2)^{-3/2}
  T dg[3][2];                 // d g_k / d u_j  (normalised Green's coefficients)
  SetC<T> sc;
};

struct TrLayout {
  size_t off_bodies, off_winv, off_scalars, off_partial, total;
  int n_tiles;
};

TrLayout tr_layout(int n_sets, int64_t n_times) {
  TrLayout L;
  const int tile = 128 * TR_K;
  L.n_tiles = (int)((n_times + tile - 1) / tile);
  size_t off = 0;
  L.off_bodies = off;
  off = align_up(off + sizeof(BodyD<double>) * (size_t)n_sets, 256);
  L.off_winv = off;
  off = align_up(off + sizeof(double) * (size_t)n_times, 256);
  L.off_scalars = off;
  off = align_up(off + 8 * sizeof(double), 256);
  L.off_partial = off;
  off = align_up(off + sizeof(double) * (size_t)n_sets * L.n_tiles * (1 + JXP_NTHETA), 256);
  L.total = off;
  return L;
}

template <typename T, int N>
__device__ __forceinline__ Dual<T, N> cast_dual(const Dual<double, N>& x) {
  Dual<T, N> r;
  r.v = T(x.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = T(x.d[i]);
  return r;
}
}  // namespace

// one thread per set: theta -> derived constants with their derivatives
template <typename T>
__global__ void k_tr_prep(const double* __restrict__ theta, int n_sets,
                          const double* __restrict__ star, int64_t star_stride,
                          BodyD<T>* __restrict__ out, double widen, double margin) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sets) return;
  typedef Dual<double, NO> D;
  const double* th = theta + (size_t)s * JXP_NTHETA;
  const double mstar = star[(size_t)s * star_stride + 0];
  const double rstar = star[(size_t)s * star_stride + 1];
  const D P = D::seed(th[0], 0);
  const D t0 = D::seed(th[1], 1);
  const D e = D::seed(th[2], 2);
  const D om = D::seed(th[3], 3);
  const D bimp = D::seed(th[4], 4);
  const double radius = th[5];
  const double pi = Num<double>::pi;

  // keplerian.py:272-276: a = cbrt(G M P^2 / (4 pi^2))
  const double G = 2942.2062175044193;
  const D x = (P * P) * (G * mstar / (4.0 * pi * pi));
  D a;
  a.v = cbrt(x.v);
#pragma unroll
  for (int i = 0; i < NO; ++i) a.d[i] = x.d[i] * a.v / (3.0 * x.v);
  D sw, cw;
  jsincos(om, &sw, &cw);
  // keplerian.py:305-315
  const D opsw = sw + 1.0;
  const D E0 = jatan2(jsqrt(1.0 - e) * cw, jsqrt(e + 1.0) * opsw) * 2.0;
  const D M0 = E0 - e * jsin(E0);
  const D ome2 = 1.0 - e * e;
  const D incl = (e * sw + 1.0) / ome2;
  // keplerian.py:318-322
  const D dcosidb = incl * rstar / a;
  const D ci = dcosidb * bimp;
  const D si = jsqrt(1.0 - ci * ci);
  // keplerian.py:334
  const D tref = -(M0 * P) / (2.0 * pi);

  BodyD<T> o;
  BodyC<T>& c = o.c;
  c.t0 = t0.v;
  c.tref = tref.v;
  c.period = P.v;
  c.t0ref = c.t0 + c.tref;
  c.inv_period = 1.0 / c.period;
  c.flags = 0;
  c.pad = 0;
  c.e = T(e.v);
  c.ome = T(1.0 - e.v);
  c.ope = T(1.0 + e.v);
  c.s1me2 = T(sqrt((1.0 - e.v) * (1.0 + e.v)));
  c.inv_ope = T(1.0 / (1.0 + e.v));
  c.cw = T(cw.v);
  c.sw = T(sw.v);
  c.ci = T(ci.v);
  c.si = T(si.v);
  const D na = -(a * ome2);
  c.na_ome2 = T(na.v);
  c.inv_rstar = T(1.0 / rstar);
  const double rr = radius / rstar;
  c.rr = T(rr);
  c.onepr = T(1.0 + fabs(rr));
  transit_window(c, a.v, rstar, e.v, widen, margin);
  o.tref = tref;
  o.cw = cast_dual<T, NO>(cw);
  o.sw = cast_dual<T, NO>(sw);
  o.ci = cast_dual<T, NO>(ci);
  o.si = cast_dual<T, NO>(si);
  o.na_ome2 = cast_dual<T, NO>(na);
  o.inv_ome2 = T(1.0 / ome2.v);
  o.inv_ome2_32 = T(1.0 / (ome2.v * sqrt(ome2.v)));
  {
    typedef Dual<double, 2> U;
    U g0, g1, g2;
    ld_greens<U>(2, U::seed(th[6], 0), U::seed(th[7], 1), g0, g1, g2);
    o.sc.g0 = T(g0.v);
    o.sc.g1 = T(g1.v);
    o.sc.g2 = T(g2.v);
    o.sc.lmax = 2;
    for (int j = 0; j < 2; ++j) {
      o.dg[0][j] = T(g0.d[j]);
      o.dg[1][j] = T(g1.d[j]);
      o.dg[2][j] = T(g2.d[j]);
    }
  }
  out[s] = o;
}

// One sample that passed the window test: flux - 1 and its 8 partials; false (outputs
// untouched) when the sample is not in transit.
template <typename T, int ORDER>
__device__ __forceinline__ bool eval_sample_partials(const BodyD<T>& bd,
                                                     const GLTable* __restrict__ tab, double tm,
                                                     T& flux, T* __restrict__ part) {
  const BodyC<T>& c = bd.c;
  typedef Dual<T, NO> D;
  // mean anomaly and its tangents (double): M = 2 pi ((t - t0) - t_ref) / P
  const double x = (tm - c.t0) - c.tref;
  const double M = (Num<double>::two_pi * x) / c.period;
  const double k = Num<double>::two_pi / c.period;
  T Md[NO];
#pragma unroll
  for (int i = 0; i < NO; ++i) Md[i] = T(-k * bd.tref.d[i]);
  Md[0] += T(-M / c.period);
  Md[1] += T(-k);
  const T Mr = T(mod_two_pi(M));
  T sf, cf, Edummy;
  kepler_reduced<T, false>(Mr, c.e, c.ome, c.ope, c.s1me2, c.inv_ope, sf, cf, Edummy);
  // kepler.py:64-79
  const T ecosf = c.e * cf;
  const T opc = T(1) + ecosf;
  const T dfdM = opc * opc * bd.inv_ome2_32;
  const T dfde = (T(2) + ecosf) * sf * bd.inv_ome2;
  D sfd, cfd;
  sfd.v = sf;
  cfd.v = cf;
#pragma unroll
  for (int i = 0; i < NO; ++i) {
    T fdot = Md[i] * dfdM;
    if (i == 2) fdot += dfde;
    sfd.d[i] = cf * fdot;
    cfd.d[i] = -sf * fdot;
  }
  const D ed = D::seed(c.e, 2);
  D b, Z;
  sky_position_g<D, T>(ed, bd.cw, bd.sw, bd.ci, bd.si, bd.na_ome2, c.inv_rstar, sfd, cfd, b, Z);
  if (!(Z.v > T(0)) || (b.v >= c.onepr)) return false;
  typedef Dual<T, 2> L;
  L s0, s1, s2;
  ld_solution<L, ORDER>(L::seed(b.v, 0), L::seed(c.rr, 1), 2, tab, s0, s1, s2);
  const L f = ld_combine(2, s0, s1, s2, L(bd.sc.g0), L(bd.sc.g1), L(bd.sc.g2));
  flux = f.v;
#pragma unroll
  for (int i = 0; i < NO; ++i) part[i] = f.d[0] * b.d[i];
  part[5] = f.d[1] * c.inv_rstar;
  part[6] = s0.v * bd.dg[0][0] + s1.v * bd.dg[1][0] + s2.v * bd.dg[2][0];
  part[7] = s0.v * bd.dg[0][1] + s1.v * bd.dg[1][1] + s2.v * bd.dg[2][1];
  return true;
}

template <typename T>
struct TrArgs {
  const BodyD<T>* bodies;
  const double* t;
  int64_t n_times;
  int n_sets, sets_per_cta, n_tiles, use_window;
  T* flux;      // [n_sets][n_times]
  T* partials;  // [n_sets][8][n_times]
  const T* y;
  const T* winv;
  double* partial;  // [n_sets][n_tiles][9]
  Stencil st;
  GLTable tab;
};

// Same CTA structure as k_system (jxp_kernels.cu): warp-owned rows of 32 samples in registers,
// sets of the group staged in shared memory, window rejection first, one instance of the
// in-transit code.  Every output row is one fully coalesced warp store.
template <typename T, int ORDER, bool LOGLIKE>
__global__ void __launch_bounds__(256, 2) k_transit_partials(const __grid_constant__ TrArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int K = TR_K;
  constexpr int NQ = 1 + JXP_NTHETA;
  const int nthreads = blockDim.x;
  const int nwarps = nthreads >> 5;
  const int tile = blockIdx.x % a.n_tiles;
  const int group = blockIdx.x / a.n_tiles;
  const int set_base = group * a.sets_per_cta;
  const int nsl = min(a.sets_per_cta, a.n_sets - set_base);
  BodyD<T>* sbody = reinterpret_cast<BodyD<T>*>(smem_raw);
  double* spart = reinterpret_cast<double*>(sbody + a.sets_per_cta);  // [nwarps][9]
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.bodies + set_base);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(sbody);
    const int nw = nsl * (int)(sizeof(BodyD<T>) / 8);
    for (int i = threadIdx.x; i < nw; i += nthreads) dst[i] = src[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx0 = ((int64_t)tile * nthreads + warp * 32) * K + lane;
  double tk[K];
  unsigned valid = 0;
  SpanK sp;
  sp.mono = true;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const bool ok = idx0 + 32 * k < a.n_times;
    tk[k] = ok ? a.t[idx0 + 32 * k] : (k > 0 ? tk[k - 1] : 0.0);
    valid |= ok ? (1u << k) : 0u;
    if (k > 0) sp.mono = sp.mono && (tk[k] >= tk[k - 1]);
  }
  sp.first = tk[0];
  sp.span = tk[K - 1] - tk[0];
  __syncthreads();
  const int nsub = a.st.n;
  const bool use_window = a.use_window != 0;

  for (int sl = 0; sl < nsl; ++sl) {
    const int set = set_base + sl;
    const BodyD<T>& bd = sbody[sl];
    T out[K][NQ];  // flux, 8 partials
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int q = 0; q < NQ; ++q) out[k][q] = T(0);
    unsigned hitmask = 0;
#pragma unroll 1
    for (int j = 0; j < nsub; ++j) {
      const double dtj = a.st.dt[j];
      unsigned mask = window_mask<T, K>(bd.c, tk, sp, dtj, valid, use_window);
      const T wj = T(a.st.w[j]);
#pragma unroll 1
      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        T f1 = T(0);
        T p1[JXP_NTHETA];
        if (eval_sample_partials<T, ORDER>(bd, &a.tab, pick(tk, k) + dtj, f1, p1)) {
          const T w = (nsub == 1) ? T(1) : wj;
          hitmask |= 1u << k;
#pragma unroll
          for (int kk = 0; kk < K; ++kk) {
            if (kk == k) {
              out[kk][0] = jfma(w, f1, out[kk][0]);
#pragma unroll
              for (int q = 0; q < JXP_NTHETA; ++q) out[kk][1 + q] = jfma(w, p1[q], out[kk][1 + q]);
            }
          }
        }
      }
    }
    if (a.flux) {
      T* dst = a.flux + (size_t)set * a.n_times + idx0;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (valid & (1u << k)) dst[32 * k] = out[k][0];
    }
    if (a.partials) {
#pragma unroll
      for (int q = 0; q < JXP_NTHETA; ++q) {
        T* dst = a.partials + ((size_t)set * JXP_NTHETA + q) * a.n_times + idx0;
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (valid & (1u << k)) dst[32 * k] = out[k][1 + q];
      }
    }
    if (LOGLIKE) {
      double acc9[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc9[q] = 0.0;
      if (hitmask) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (hitmask & (1u << k)) {
            const T w = a.winv[idx0 + 32 * k];
            const T r0 = (a.y[idx0 + 32 * k] - T(1)) * w;
            const T mw = out[k][0] * w;
            const T rm = r0 - mw;  // ((y - 1) - m) / sigma
            acc9[0] += (double)(-mw) * (double)(T(2) * r0 - mw);
            const double g = (double)(rm * w);  // d loglike / d m
#pragma unroll
            for (int q = 0; q < JXP_NTHETA; ++q) acc9[1 + q] += g * (double)out[k][1 + q];
          }
        }
      }
      __syncthreads();  // spart is reused by every set of the group
      if (__any_sync(0xffffffffu, hitmask != 0)) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
          for (int o = 16; o > 0; o >>= 1) acc9[q] += __shfl_down_sync(0xffffffffu, acc9[q], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) spart[warp * NQ + q] = acc9[q];
      }
      __syncthreads();
      if (threadIdx.x < NQ) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += spart[w * NQ + threadIdx.x];
        a.partial[((size_t)set * a.n_tiles + tile) * NQ + threadIdx.x] = s;
      }
    }
  }
}

// loglike[s] = -chi2/2 + const, dloglike[s][k] = sum of the per-tile gradient partials
template <typename T>
__global__ void k_tr_final(const double* __restrict__ partial, int n_tiles, int n_sets,
                           const double* __restrict__ scalars, T* __restrict__ loglike,
                           T* __restrict__ dloglike) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sets * (1 + JXP_NTHETA)) return;
  const int s = i / (1 + JXP_NTHETA), q = i % (1 + JXP_NTHETA);
  double acc = 0.0;
  for (int k = 0; k < n_tiles; ++k) acc += partial[((size_t)s * n_tiles + k) * (1 + JXP_NTHETA) + q];
  if (q == 0) {
    if (loglike) loglike[s] = T(-0.5 * (scalars[0] + acc) + scalars[1]);
  } else if (dloglike) {
    dloglike[(size_t)s * JXP_NTHETA + (q - 1)] = T(acc);
  }
}

extern "C" size_t jxp_transit_workspace_bytes(int32_t n_sets, int64_t n_times) {
  if (n_sets <= 0 || n_times < 0) return 0;
  return tr_layout(n_sets, n_times).total;
}

namespace jxph {
// launches k_loglike_prep<T> (instantiated in jxp_kernels.cu)
int launch_loglike_prep_f64(const double* y, const double* sigma, int64_t ss, int64_t n, double* winv,
                            double* scalars, cudaStream_t st);
int launch_loglike_prep_f32(const float* y, const float* sigma, int64_t ss, int64_t n, float* winv,
                            double* scalars, cudaStream_t st);
}  // namespace jxph

static int loglike_prep(const double* y, const double* s, int64_t ss, int64_t n, double* w, double* sc,
                        cudaStream_t st) {
  return launch_loglike_prep_f64(y, s, ss, n, w, sc, st);
}
static int loglike_prep(const float* y, const float* s, int64_t ss, int64_t n, float* w, double* sc,
                        cudaStream_t st) {
  return launch_loglike_prep_f32(y, s, ss, n, w, sc, st);
}

template <typename T>
static int launch_transit(const double* theta, int32_t n_sets, const double* star,
                          int64_t star_stride, const double* t, int64_t n_times,
                          double exposure_time, int32_t exp_order, int32_t num_samples,
                          int32_t gl_order, int32_t flags, T* flux, T* partials, const T* y,
                          const T* sigma, int64_t sigma_stride, T* loglike, T* dloglike,
                          void* workspace, size_t workspace_bytes, void* stream) {
  static_assert(sizeof(BodyD<T>) % 8 == 0, "record size");
  if (n_sets < 0 || n_times < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials: negative size");
  if (gl_order < 1 || gl_order > JXP_MAX_GL_ORDER)
    return fail(JXP_ERR_UNSUPPORTED, "jxp_transit_partials: order = %d outside 1..%d", gl_order,
                JXP_MAX_GL_ORDER);
  if (star_stride != 0 && star_stride != 2)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials: star_set_stride must be 0 or 2");
  if (!stride_ok(sigma_stride))
    return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials: sigma_stride must be 0 or 1");
  TrArgs<T> a;
  int rc = make_stencil(exposure_time, exp_order, num_samples, a.st);
  if (rc != JXP_OK) return rc;
  if (n_sets == 0 || n_times == 0) return JXP_OK;
  if (!theta || !star || !t) return fail(JXP_ERR_NULL_POINTER, "jxp_transit_partials: null theta/star/t");
  const bool want_ll = loglike || dloglike;
  if (want_ll && (!y || !sigma))
    return fail(JXP_ERR_NULL_POINTER, "jxp_transit_partials: loglike needs y and sigma");
  if (!flux && !partials && !want_ll)
    return fail(JXP_ERR_NULL_POINTER, "jxp_transit_partials: no output requested");
  const TrLayout L = tr_layout(n_sets, n_times);
  if (!workspace || workspace_bytes < L.total)
    return fail(JXP_ERR_WORKSPACE, "jxp_transit_partials: workspace of %zu bytes needed, %zu given",
                L.total, workspace_bytes);
  if (((uintptr_t)workspace & 255) != 0)
    return fail(JXP_ERR_WORKSPACE, "jxp_transit_partials: workspace must be 256-byte aligned");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  BodyD<T>* d_bodies = reinterpret_cast<BodyD<T>*>(ws + L.off_bodies);
  T* d_winv = reinterpret_cast<T*>(ws + L.off_winv);
  double* d_scalars = reinterpret_cast<double*>(ws + L.off_scalars);
  double* d_partial = reinterpret_cast<double*>(ws + L.off_partial);
  cudaStream_t st = (cudaStream_t)stream;
  const bool f32 = sizeof(T) == 4;

  k_tr_prep<T><<<(n_sets + 127) / 128, 128, 0, st>>>(theta, n_sets, star, star_stride, d_bodies,
                                                    f32 ? 1.0 + 2e-4 : 1.0 + 1e-9, 1e-7);
  rc = check_launch("jxp_transit_partials(prep)");
  if (rc != JXP_OK) return rc;
  if (want_ll) {
    rc = loglike_prep(y, sigma, sigma_stride, n_times, d_winv, d_scalars, st);
    if (rc != JXP_OK) return rc;
  }
  const int sms = sm_count();
  int threads = 256;
  int64_t n_tiles = (n_times + (int64_t)threads * TR_K - 1) / ((int64_t)threads * TR_K);
  if (n_tiles * n_sets < 2 * sms) {
    threads = 128;
    n_tiles = (n_times + (int64_t)threads * TR_K - 1) / ((int64_t)threads * TR_K);
  }
  int spc = (int)((n_tiles * (int64_t)n_sets) / ((int64_t)sms * 32));
  if (spc < 1) spc = 1;
  if (spc > TR_MAX_SETS) spc = TR_MAX_SETS;
  if (spc > n_sets) spc = n_sets;
  const int64_t n_groups = (n_sets + spc - 1) / spc;
  const int64_t n_ctas = n_tiles * n_groups;
  if (n_ctas > 2147483647LL) return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials: grid too large");
  a.bodies = d_bodies;
  a.t = t;
  a.n_times = n_times;
  a.n_sets = n_sets;
  a.sets_per_cta = spc;
  a.n_tiles = (int)n_tiles;
  a.use_window = (flags & JXP_FLAG_NO_WINDOW) ? 0 : 1;
  a.flux = flux;
  a.partials = partials;
  a.y = y;
  a.winv = d_winv;
  a.partial = d_partial;
  a.tab.order = gl_order;
  if (gl_order != 10) gl_table_host(gl_order, a.tab);
  const size_t smem = sizeof(BodyD<T>) * (size_t)spc + sizeof(double) * (1 + JXP_NTHETA) * (threads / 32);
  auto launch = [&](auto kern) -> int {
    kern<<<(unsigned)n_ctas, threads, smem, st>>>(a);
    return check_launch("jxp_transit_partials");
  };
  if (gl_order == 10)
    rc = want_ll ? launch(k_transit_partials<T, 10, true>) : launch(k_transit_partials<T, 10, false>);
  else
    rc = want_ll ? launch(k_transit_partials<T, 0, true>) : launch(k_transit_partials<T, 0, false>);
  if (rc != JXP_OK) return rc;
  if (want_ll) {
    const int n = n_sets * (1 + JXP_NTHETA);
    k_tr_final<T><<<(n + 127) / 128, 128, 0, st>>>(d_partial, (int)n_tiles, n_sets, d_scalars, loglike,
                                                  dloglike);
    rc = check_launch("jxp_transit_partials(final)");
  }
  return rc;
}

extern "C" int jxp_transit_partials_f64(const double* theta, int32_t n_sets, const double* star,
                                        int64_t star_set_stride, const double* t, int64_t n_times,
                                        double exposure_time, int32_t exp_order,
                                        int32_t num_samples, int32_t gl_order, int32_t flags,
                                        double* flux, double* partials, const double* y,
                                        const double* sigma, int64_t sigma_stride, double* loglike,
                                        double* dloglike, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  return launch_transit<double>(theta, n_sets, star, star_set_stride, t, n_times, exposure_time,
                                exp_order, num_samples, gl_order, flags, flux, partials, y, sigma,
                                sigma_stride, loglike, dloglike, workspace, workspace_bytes, stream);
}
extern "C" int jxp_transit_partials_f32(const double* theta, int32_t n_sets, const double* star,
                                        int64_t star_set_stride, const double* t, int64_t n_times,
                                        double exposure_time, int32_t exp_order,
                                        int32_t num_samples, int32_t gl_order, int32_t flags,
                                        float* flux, float* partials, const float* y,
                                        const float* sigma, int64_t sigma_stride, float* loglike,
                                        float* dloglike, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  return launch_transit<float>(theta, n_sets, star, star_set_stride, t, n_times, exposure_time,
                               exp_order, num_samples, gl_order, flags, flux, partials, y, sigma,
                               sigma_stride, loglike, dloglike, workspace, workspace_bytes, stream);
}
