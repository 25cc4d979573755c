// Kernels and C-ABI launchers of jaxoplanet_b200 (sm_100a).  See include/jaxoplanet_b200.h
// for the contract of every entry point and DESIGN.md for the kernel designs.
#include "../../include/jaxoplanet_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <type_traits>

#include "jxp_host.h"
#include "jxp_orbit.cuh"
#include "jxp_list.cuh"
#include "jxp_prep.cuh"

using namespace jxp;
using namespace jxph;

// ======================================================================================
// error plumbing
// ======================================================================================
namespace {
thread_local char g_err[512] = "";
thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
}

// development / measurement options (include/jaxoplanet_b200.h, JXP_DEV_*): one atomic word, read once
// per call; the library never reads the environment
static std::atomic<uint32_t> g_dev_options{0u};
extern "C" int jxp_set_dev_options(uint32_t mask) {
  g_dev_options.store(mask, std::memory_order_relaxed);
  return JXP_OK;
}
extern "C" uint32_t jxp_get_dev_options(void) { return g_dev_options.load(std::memory_order_relaxed); }
namespace jxph {
uint32_t dev_options() { return g_dev_options.load(std::memory_order_relaxed); }
}

extern "C" int jxp_profile_events(void* start_event, void* stop_event) {
  g_ev_start = (cudaEvent_t)start_event;
  g_ev_stop = (cudaEvent_t)stop_event;
  return JXP_OK;
}

namespace jxph {
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(JXP_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return JXP_OK;
}

bool stride_ok(int64_t s) { return s == 0 || s == 1; }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    cached_dev = dev;
  }
  return cached;
}

// Gauss-Legendre nodes/weights on the host for orders other than the built-in 10
// (the reference calls scipy.special.roots_legendre(order), limb_dark.py:155).
void gl_table_host(int order, GLTable& tab) {
  tab.order = order;
  const double pi = 3.141592653589793238462643383279502884;
  for (int i = 0; i < (order + 1) / 2; ++i) {
    double x = std::cos(pi * (i + 0.75) / (order + 0.5));
    double dp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = x;
      for (int k = 2; k <= order; ++k) {
        double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = pk;
      }
      if (order == 1) {
        p0 = 1.0;
        p1 = x;
      }
      dp = order * (x * p1 - p0) / (x * x - 1.0);
      double dx = p1 / dp;
      x -= dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    // recompute derivative at the converged node
    {
      double p0 = 1.0, p1 = x;
      for (int k = 2; k <= order; ++k) {
        double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = pk;
      }
      dp = order * (x * p1 - p0) / (x * x - 1.0);
    }
    double w = 2.0 / ((1.0 - x * x) * dp * dp);
    if (2 * i + 1 == order) x = 0.0;
    tab.x[i] = -x;
    tab.x[order - 1 - i] = x;
    tab.w[i] = w;
    tab.w[order - 1 - i] = w;
  }
}
}  // namespace jxph

extern "C" int jxp_version(void) { return JXP_ABI_VERSION; }
extern "C" const char* jxp_last_error(void) { return g_err; }

// ======================================================================================
// K1  Kepler
// ======================================================================================
#ifndef JXP_KEP_ILP
#define JXP_KEP_ILP 2
#endif
template <typename T, bool EXTRA>
__global__ void __launch_bounds__(256)
k_kepler(const T* __restrict__ M, int64_t Ms, const T* __restrict__ ecc, int64_t es, int64_t n,
         T* __restrict__ sinf_o, T* __restrict__ cosf_o, T* __restrict__ E_o,
         T* __restrict__ dfdM_o, T* __restrict__ dfde_o) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (!EXTRA && std::is_same<T, double>::value) {
    // JXP_KEP_ILP independent solves per thread and iteration (see kepler_reduced_n)
    constexpr int NI = JXP_KEP_ILP;
    for (; i + (NI - 1) * stride < n; i += NI * stride) {
      double Mr[NI], e[NI], ome[NI], ope[NI], s1[NI], iope[NI], sf[NI], cf[NI];
#pragma unroll
      for (int q = 0; q < NI; ++q) {
        const double m = M[(i + q * stride) * Ms];
        e[q] = ecc[(i + q * stride) * es];
        ome[q] = 1.0 - e[q];
        ope[q] = 1.0 + e[q];
        Mr[q] = mod_two_pi(m);
      }
#pragma unroll
      for (int q = 0; q < NI; ++q) s1[q] = jsqrt(ome[q] * ope[q]);
#pragma unroll
      for (int q = 0; q < NI; ++q) iope[q] = jrcp(ope[q]);
      kepler_reduced_n<NI, false, true>(Mr, e, ome, ope, s1, iope, sf, cf);  // one division for d3 -> d4 -> dE
#pragma unroll
      for (int q = 0; q < NI; ++q) {
        sinf_o[i + q * stride] = sf[q];
        cosf_o[i + q * stride] = cf[q];
      }
    }
  }
  for (; i < n; i += stride) {
    const T m = M[i * Ms];
    const T e = ecc[i * es];
    const T ome = T(1) - e, ope = T(1) + e;
    const T ome2 = ome * ope;
    const T s1me2 = jsqrt(ome2);
    const T inv_ope = jrcp(ope);
    const T Mr = mod_two_pi(m);
    T sf, cf, E;
    kepler_reduced<T, EXTRA>(Mr, e, ome, ope, s1me2, inv_ope, sf, cf, E);
    sinf_o[i] = sf;
    cosf_o[i] = cf;
    if (EXTRA) {
      if (E_o) E_o[i] = E;
      if (dfdM_o || dfde_o) {
        // kepler.py:64-77
        const T ecosf = e * cf;
        const T inv = jrcp(ome2);
        const T opc = T(1) + ecosf;
        if (dfdM_o) dfdM_o[i] = jdiv(opc * opc * inv, s1me2);
        if (dfde_o) dfde_o[i] = (T(2) + ecosf) * sf * inv;
      }
    }
  }
}

template <typename T>
static int launch_kepler(const T* M, int64_t Ms, const T* ecc, int64_t es, int64_t n, T* sinf,
                         T* cosf, T* E, T* dfdM, T* dfde, void* stream) {
  if (n < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_kepler: n = %lld < 0", (long long)n);
  if (n == 0) return JXP_OK;
  if (!M || !ecc || !sinf || !cosf) return fail(JXP_ERR_NULL_POINTER, "jxp_kepler: null M/ecc/sinf/cosf");
  if (!stride_ok(Ms) || !stride_ok(es))
    return fail(JXP_ERR_BAD_SHAPE, "jxp_kepler: strides must be 0 or 1");
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 8 * 4;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (E || dfdM || dfde)
    k_kepler<T, true><<<(unsigned)blocks, threads, 0, st>>>(M, Ms, ecc, es, n, sinf, cosf, E, dfdM, dfde);
  else
    k_kepler<T, false><<<(unsigned)blocks, threads, 0, st>>>(M, Ms, ecc, es, n, sinf, cosf, E, dfdM, dfde);
  return check_launch("jxp_kepler");
}

extern "C" int jxp_kepler_f64(const double* M, int64_t Ms, const double* ecc, int64_t es,
                              int64_t n, double* sinf, double* cosf, double* E, double* dfdM,
                              double* dfde, void* stream) {
  return launch_kepler<double>(M, Ms, ecc, es, n, sinf, cosf, E, dfdM, dfde, stream);
}
extern "C" int jxp_kepler_f32(const float* M, int64_t Ms, const float* ecc, int64_t es, int64_t n,
                              float* sinf, float* cosf, float* E, float* dfdM, float* dfde,
                              void* stream) {
  return launch_kepler<float>(M, Ms, ecc, es, n, sinf, cosf, E, dfdM, dfde, stream);
}

// ======================================================================================
// K2  limb-darkened flux (core)
// ======================================================================================
template <typename T>
struct LdArgs {
  const T* u;
  int n_u;
  const T* b;
  int64_t bs;
  const T* r;
  int64_t rs;
  int64_t n;
  T* flux;
  T* db;
  T* dr;
  T* svec;
  GLTable tab;
};

template <typename T, int ORDER, bool GRAD, bool HI>
__global__ void __launch_bounds__(256) k_limb_dark(const __grid_constant__ LdArgs<T> a) {
  T uu[JXP_MAX_LD_COEFFS], g[JXP_MAX_LD_COEFFS + 1];
#pragma unroll
  for (int k = 0; k < JXP_MAX_LD_COEFFS; ++k) uu[k] = (k < a.n_u) ? a.u[k] : T(0);
#pragma unroll
  for (int k = 0; k <= JXP_MAX_LD_COEFFS; ++k) g[k] = T(0);
  if (HI) {
    ld_greens_n<T>(a.n_u, uu, g);
  } else {
    ld_greens<T>(a.n_u, uu[0], uu[1], g[0], g[1], g[2]);
  }
  const int lmax = a.n_u;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const T b = a.b[i * a.bs];
    const T r = a.r[i * a.rs];
    if (GRAD) {
      typedef Dual<T, 2> D;
      D s0, s1, s2, shi[6];
      ld_solution<D, ORDER, HI>(D::seed(b, 0), D::seed(r, 1), lmax, &a.tab, s0, s1, s2, shi);
      D f = ld_combine(lmax < 2 ? lmax : 2, s0, s1, s2, D(g[0]), D(g[1]), D(g[2]));
      if (HI) {
#pragma unroll
        for (int n = 3; n <= 8; ++n)
          if (n <= lmax) f = f + shi[n - 3] * g[n];
      }
      a.flux[i] = f.v;
      if (a.db) a.db[i] = f.d[0];
      if (a.dr) a.dr[i] = f.d[1];
      if (a.svec) {
        a.svec[i] = s0.v;
        if (lmax >= 1) a.svec[a.n + i] = s1.v;
        if (lmax >= 2) a.svec[2 * a.n + i] = s2.v;
        if (HI) {
#pragma unroll
          for (int n = 3; n <= 8; ++n)
            if (n <= lmax) a.svec[(int64_t)n * a.n + i] = shi[n - 3].v;
        }
      }
    } else {
      T s0, s1, s2, shi[6];
      ld_solution<T, ORDER, HI, 5>(b, r, lmax, &a.tab, s0, s1, s2, shi);
      T f = ld_combine(lmax < 2 ? lmax : 2, s0, s1, s2, g[0], g[1], g[2]);
      if (HI) {
#pragma unroll
        for (int n = 3; n <= 8; ++n)
          if (n <= lmax) f = jfma(shi[n - 3], g[n], f);
      }
      a.flux[i] = f;
    }
  }
}

// K2 with the points sorted by what they need (double, quadratic law, order 10, no derivatives): a
// planet fully inside the disk (b < 1 - r, r < 1) takes ld_solution_inside_n (constant kappas and node
// cosines: 6.1e10 evaluations/s dense against 2.9e10 for the general mirror-pair code,
// tools/ubench/ld_bench.cu), everything else the general code.  Unlike the fused kernels K2 sees (b, r)
// in any order, so every WARP keeps two private queues of (b, r, index) in shared memory, one per class:
// it classifies 128 consecutive points at a time into them and, whenever a queue holds 64 entries,
// evaluates 64 points of that class, two per lane, from the top of the queue.  No block barriers, no
// partial warps before the very end, and the code a point gets depends on the point alone.
constexpr int LDC_SUB = 128;   // points classified per round
constexpr int LDC_QCAP = 192;  // per class: up to 63 left over + 128 new
__global__ void __launch_bounds__(128, 4) k_limb_dark_cls(const __grid_constant__ LdArgs<double> a) {
  __shared__ double s_b[4][2][LDC_QCAP], s_r[4][2][LDC_QCAP];
  __shared__ unsigned long long s_i[4][2][LDC_QCAP];
  double g0, g1, g2;
  ld_greens<double>(a.n_u, a.n_u > 0 ? a.u[0] : 0.0, a.n_u > 1 ? a.u[1] : 0.0, g0, g1, g2);
  const int lmax = a.n_u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  int qn[2] = {0, 0};
  // 64 entries (fewer only in the final flush) from the top of queue c
  auto drain = [&](int c) {
    const int take = qn[c] < 64 ? qn[c] : 64;
    const int first = qn[c] - take;
    bool have[2];
    unsigned long long i[2];
    double b[2], r[2], s0[2], s1[2], s2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      have[h] = 32 * h + lane < take;
      const int e = first + (have[h] ? 32 * h + lane : 0);
      b[h] = s_b[warp][c][e];
      r[h] = s_r[warp][c][e];
      i[h] = s_i[warp][c][e];
    }
    if (c == 0)
      ld_solution_inside_n<2>(b, r, lmax, s0, s1, s2);
    else
      ld_solution_n<2, 1>(b, r, lmax, s0, s1, s2);
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (have[h]) a.flux[i[h]] = ld_combine(lmax, s0[h], s1[h], s2[h], g0, g1, g2);
    qn[c] = first;
    __syncwarp();
  };
  const int64_t n_sub = (a.n + LDC_SUB - 1) / LDC_SUB;
  const int64_t nw = (int64_t)gridDim.x * 4;
  for (int64_t sub = (int64_t)blockIdx.x * 4 + warp; sub < n_sub; sub += nw) {
#pragma unroll
    for (int k = 0; k < LDC_SUB / 32; ++k) {
      const int64_t i = sub * LDC_SUB + 32 * k + lane;
      int cls = -1;
      double b0 = 0.0, r0 = 0.0;
      if (i < a.n) {
        b0 = a.b[i * a.bs];
        r0 = a.r[i * a.rs];
        const double b = fabs(b0), r = fabs(r0);
        cls = ((b < 1.0 - r) && (r < 1.0)) ? 0 : 1;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == c);
        if (cls == c) {
          const int e = qn[c] + __popc(m & lt);
          s_b[warp][c][e] = b0;
          s_r[warp][c][e] = r0;
          s_i[warp][c][e] = (unsigned long long)i;
        }
        qn[c] += __popc(m);
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int c = 0; c < 2; ++c)
      while (qn[c] >= 64) drain(c);
  }
#pragma unroll 1
  for (int c = 0; c < 2; ++c)
    while (qn[c] > 0) drain(c);
}

template <typename T>
static int launch_limb_dark(const T* u, int32_t n_u, const T* b, int64_t bs, const T* r,
                            int64_t rs, int64_t n, int32_t order, T* flux, T* db, T* dr, T* svec,
                            void* stream) {
  if (n < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_limb_dark: n = %lld < 0", (long long)n);
  if (n_u < 0 || n_u > JXP_MAX_LD_COEFFS)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_limb_dark: n_u = %d out of range", n_u);
  if (order < 1 || order > JXP_MAX_GL_ORDER)
    return fail(JXP_ERR_UNSUPPORTED, "jxp_limb_dark: order = %d outside 1..%d", order,
                JXP_MAX_GL_ORDER);
  if (n == 0) return JXP_OK;
  if (!b || !r || !flux || (n_u > 0 && !u))
    return fail(JXP_ERR_NULL_POINTER, "jxp_limb_dark: null u/b/r/flux");
  if (!stride_ok(bs) || !stride_ok(rs))
    return fail(JXP_ERR_BAD_SHAPE, "jxp_limb_dark: strides must be 0 or 1");
  LdArgs<T> a;
  a.u = u;
  a.n_u = n_u;
  a.b = b;
  a.bs = bs;
  a.r = r;
  a.rs = rs;
  a.n = n;
  a.flux = flux;
  a.db = db;
  a.dr = dr;
  a.svec = svec;
  a.tab.order = order;
  if (order != 10) gl_table_host(order, a.tab);
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 8 * 4;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const bool grad = db || dr || svec;
  const bool hi = n_u > 2;
  auto go = [&](auto kern) { kern<<<(unsigned)blocks, threads, 0, st>>>(a); };
  if (hi) {
    // general polynomial law: rolled node loop (order 10 from the constant bank, or the table)
    if (order == 10) {
      if (grad) go(k_limb_dark<T, 11, true, true>); else go(k_limb_dark<T, 11, false, true>);
    } else {
      if (grad) go(k_limb_dark<T, 0, true, true>); else go(k_limb_dark<T, 0, false, true>);
    }
  } else if (order == 10) {
    if (sizeof(T) == 8) {
      // double: the mirror-pair node loop (ORDER tag 11), 2.9e10 vs 2.0e10 evaluations/s
      if (grad) {
        go(k_limb_dark<T, 11, true, false>);
      } else if (n >= 8192 && !(dev_options() & JXP_DEV_NO_LD_CLASSES)) {
        if constexpr (sizeof(T) == 8) {
          const int64_t ctas = (n + 4 * LDC_SUB - 1) / (4 * LDC_SUB), capc = (int64_t)sm_count() * 4;
          k_limb_dark_cls<<<(unsigned)(ctas < capc ? ctas : capc), 128, 0, st>>>(a);
        }
      } else {
        go(k_limb_dark<T, 11, false, false>);
      }
    } else {
      if (grad) go(k_limb_dark<T, 10, true, false>); else go(k_limb_dark<T, 10, false, false>);
    }
  } else {
    if (grad) go(k_limb_dark<T, 0, true, false>); else go(k_limb_dark<T, 0, false, false>);
  }
  return check_launch("jxp_limb_dark");
}

extern "C" int jxp_limb_dark_f64(const double* u, int32_t n_u, const double* b, int64_t bs,
                                 const double* r, int64_t rs, int64_t n, int32_t order,
                                 double* flux, double* db, double* dr, double* svec, void* stream) {
  return launch_limb_dark<double>(u, n_u, b, bs, r, rs, n, order, flux, db, dr, svec, stream);
}
extern "C" int jxp_limb_dark_f32(const float* u, int32_t n_u, const float* b, int64_t bs,
                                 const float* r, int64_t rs, int64_t n, int32_t order, float* flux,
                                 float* db, float* dr, float* svec, void* stream) {
  return launch_limb_dark<float>(u, n_u, b, bs, r, rs, n, order, flux, db, dr, svec, stream);
}

// ======================================================================================
// K3/K4  fused system light curve (+ exposure integration, + Gaussian log-likelihood)
// ======================================================================================
namespace {
constexpr int SYS_K = SYS_K_H;    // time samples held in registers per thread
constexpr int SYS_MAX_SETS = 64;  // parameter sets per CTA (upper bound)
}  // namespace

namespace jxph {
// transforms.py:67-89
int make_stencil(double exposure_time, int order, int num_samples, Stencil& st) {
  if (!(exposure_time > 0.0) || num_samples <= 0) {
    st.n = 1;
    st.dt[0] = 0.0;
    st.w[0] = 1.0;
    st.dt_min = st.dt_max = 0.0;
    return JXP_OK;
  }
  int ns = num_samples;
  ns += 1 - ns % 2;
  if (ns > MAX_SUB - 1)
    return fail(JXP_ERR_UNSUPPORTED, "exposure integration: num_samples = %d > %d", ns, MAX_SUB - 1);
  st.n = ns;
  double sum = 0.0;
  if (order == 0) {
    // linspace(-0.5, 0.5, 2 ns + 1)[1:-1:2]
    const int m = 2 * ns + 1;
    const double step = 1.0 / (m - 1);
    for (int j = 0; j < ns; ++j) {
      const int idx = 2 * j + 1;
      st.dt[j] = (idx == m - 1) ? 0.5 : -0.5 + idx * step;
      st.w[j] = 1.0;
    }
  } else if (order == 1 || order == 2) {
    const double step = ns > 1 ? 1.0 / (ns - 1) : 0.0;
    for (int j = 0; j < ns; ++j) {
      st.dt[j] = (j == ns - 1 && ns > 1) ? 0.5 : -0.5 + j * step;
      if (order == 1)
        st.w[j] = (j == 0 || j == ns - 1) ? 1.0 : 2.0;
      else
        st.w[j] = (j == 0 || j == ns - 1) ? 1.0 : ((j % 2 == 1) ? 4.0 : 2.0);
    }
  } else {
    return fail(JXP_ERR_UNSUPPORTED,
                "The parameter 'order' in 'integrate_exposure_time' must be 0, 1, or 2");
  }
  for (int j = 0; j < ns; ++j) sum += st.w[j];
  st.dt_min = 0.0;
  st.dt_max = 0.0;
  for (int j = 0; j < ns; ++j) {
    st.dt[j] *= exposure_time;
    st.w[j] /= sum;
    st.dt_min = st.dt[j] < st.dt_min ? st.dt[j] : st.dt_min;
    st.dt_max = st.dt[j] > st.dt_max ? st.dt[j] : st.dt_max;
  }
  return JXP_OK;
}
}  // namespace jxph


template <typename T>
__global__ void k_sys_prep(const double* __restrict__ bodies, int n_sets, int n_bodies,
                           const double* __restrict__ u, int n_u, int64_t u_stride,
                           BodyC<T>* __restrict__ out_b, SetC<T>* __restrict__ out_s,
                           unsigned long long* __restrict__ counters, int* __restrict__ tile_ctr,
                           double widen, double margin) {
  sys_prep_body<T>(blockIdx.x * blockDim.x + threadIdx.x, bodies, n_sets, n_bodies, u, n_u, u_stride, out_b,
                   out_s, counters, tile_ctr, widen, margin);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_loglike_prep(const T* __restrict__ y, const T* __restrict__ sigma, int64_t ss, int64_t n,
               T* __restrict__ winv, double* __restrict__ scalars) {
  loglike_prep_body<T>(blockIdx.x, gridDim.x, y, sigma, ss, n, winv, scalars, nullptr);
}

// the sortedness flags alone (list path of calls without a likelihood): LL_NCTA CTAs of 256 threads
__global__ void __launch_bounds__(256)
k_tl_sorted_check(const double* __restrict__ t, int64_t n, double* __restrict__ scalars) {
  int bad = 0;
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = lo + chunk < n ? lo + chunk : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
    if (i + 1 < n && !(t[i + 1] >= t[i])) bad = 1;
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) scalars[LL_FLAG0 + blockIdx.x] = bad ? 1.0 : 0.0;
}

// k_sys_prep and k_loglike_prep in one launch (they are independent): the first n_prep CTAs derive
// the records, the next LL_NCTA the likelihood constants (and the sortedness flags for the list path)
template <typename T>
__global__ void __launch_bounds__(256)
k_sys_prep_ll(const double* __restrict__ bodies, int n_sets, int n_bodies, const double* __restrict__ u, int n_u,
              int64_t u_stride, BodyC<T>* __restrict__ out_b, SetC<T>* __restrict__ out_s,
              unsigned long long* __restrict__ counters, int* __restrict__ tile_ctr, double widen, double margin,
              int n_prep, const T* __restrict__ y, const T* __restrict__ sigma, int64_t ss, int64_t n,
              T* __restrict__ winv, double* __restrict__ scalars, const double* __restrict__ t) {
  if ((int)blockIdx.x < n_prep)
    sys_prep_body<T>(blockIdx.x * blockDim.x + threadIdx.x, bodies, n_sets, n_bodies, u, n_u, u_stride, out_b,
                     out_s, counters, tile_ctr, widen, margin);
  else
    loglike_prep_body<T>((int)blockIdx.x - n_prep, LL_NCTA, y, sigma, ss, n, winv, scalars, t);
}

namespace jxph {
template <typename T>
__global__ void __launch_bounds__(256)
k_loglike_prep_t(const T* __restrict__ y, const T* __restrict__ sigma, int64_t ss, int64_t n,
                 T* __restrict__ winv, double* __restrict__ scalars, const double* __restrict__ t) {
  loglike_prep_body<T>(blockIdx.x, gridDim.x, y, sigma, ss, n, winv, scalars, t);
}
int launch_loglike_prep_f64(const double* y, const double* sigma, int64_t ss, int64_t n, double* winv,
                            double* scalars, cudaStream_t st, const double* t_sorted_check) {
  if (t_sorted_check)
    k_loglike_prep_t<double><<<LL_NCTA, 256, 0, st>>>(y, sigma, ss, n, winv, scalars, t_sorted_check);
  else
    k_loglike_prep<double><<<LL_NCTA, 256, 0, st>>>(y, sigma, ss, n, winv, scalars);
  return check_launch("loglike prep");
}
int launch_loglike_prep_f32(const float* y, const float* sigma, int64_t ss, int64_t n, float* winv,
                            double* scalars, cudaStream_t st) {
  k_loglike_prep<float><<<LL_NCTA, 256, 0, st>>>(y, sigma, ss, n, winv, scalars);
  return check_launch("loglike prep");
}
}  // namespace jxph

// loglike[s] = -(chi0 + sum_tiles partial[tile][s]) / 2 + const.  Block = 32 sets x 8 tile
// slices: slice j adds tiles j, j + 8, ... in order, the 8 slice sums are then added in order
// (fixed summation order -> deterministic); loads are coalesced over sets.
template <typename T>
__global__ void __launch_bounds__(256)
k_loglike_final(const double* __restrict__ partial, int n_tiles, int n_sets,
                const double* __restrict__ scalars, T* __restrict__ loglike,
                const unsigned long long* __restrict__ tl_state = nullptr, unsigned long long tl_cap = 0) {
  __shared__ double sh[8][33];
  if (tl_state != nullptr && tl_active(tl_state, tl_cap)) return;  // the transit-list kernels did the work
  const int ls = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int s = blockIdx.x * 32 + ls;
  double acc = 0.0;
  if (s < n_sets)
    for (int k = slice; k < n_tiles; k += 8) acc += partial[(size_t)k * n_sets + s];
  sh[slice][ls] = acc;
  __syncthreads();
  if (slice == 0 && s < n_sets) {
    double chi0 = 0.0, cst = 0.0;
    for (int k = 0; k < LL_NCTA; ++k) {
      chi0 += scalars[LL_SCAL0 + 2 * k];
      cst += scalars[LL_SCAL0 + 2 * k + 1];
    }
    double tot = 0.0;
    for (int j = 0; j < 8; ++j) tot += sh[j][ls];
    loglike[s] = T(-0.5 * (chi0 + tot) + cst);
  }
}

template <typename T>
struct SysArgs {
  const BodyC<T>* bodies;
  const SetC<T>* sets;
  const double* t;
  int64_t n_times;
  int n_sets, n_bodies, sets_per_cta, n_groups, n_wtiles, use_window;
  int* tile_ctr;  // [1 + n_groups] work distribution (zeroed by k_sys_prep)
  T* flux;
  T* flux_sum;
  const T* y;
  const T* winv;
  double* partial;  // [n_wtiles][n_sets]
  unsigned long long* counters;  // [2]: Kepler solves, limb-darkening evaluations executed
  // transit-list path (nullptr: not used by this call)
  unsigned long long* tl_state;  // [0] items on the inside list, [1] TL_UNSORTED | TL_OVERFLOW, [2] on the edge list
  unsigned long long* tl_items;  // [tl_cap] (set << 32 | sample), overwritten by the chi^2 term; inside
                                 // list = [0, state[0]), edge list = [cap - state[2], cap)
  unsigned long long* tl_off;    // [n_sets][2] offset of the set's slice in the inside / edge list
  unsigned* tl_cnt;              // [n_sets][2] items of the set on the inside / edge list
  unsigned long long tl_cap;
  const double* scalars;
  double* loglike;
  // byte offsets of the regions of k_system's dynamic shared memory (computed once by the launcher: the
  // kernel re-derived them from sets_per_cta / n_bodies / warps at every use under register pressure --
  // 9 % of config 4's instructions)
  int so_sset, so_sacc, so_sq, so_sqb, so_macc, so_meq, so_maccE;
  Stencil st;
  GLTable tab;
};

// k_system -- fused light curve over (parameter set x time sample), warp-centric.
//
// CTA = NW warps x (group of <= SYS_MAX_SETS parameter sets).  The group's records are staged in
// shared memory once per CTA; after that the warps are independent (no block barriers).  A warp
// owns 32 * SYS_K consecutive time samples as SYS_K rows of 32 (lane = column), time stamps in
// registers.  Three stages per warp:
//  A  candidate sets.  Lanes take one SET each: the warp's whole time span [tmin, tmax] (widened
//     by the exposure stencil) is tested against each body's transit window; a ballot gives the
//     sets that can have any sample in transit.  Cost ~1/100 instruction per (set, sample).
//  B  for a candidate set, lanes take their own samples again and test them one by one; hits
//     are appended (ballot + popc compaction) to a per-warp queue of (set, row, lane) items.
//  C  the queue is drained 32 items at a time: EVERY lane runs the Kepler + position + flux code
//     on a different in-window sample (of possibly different sets), so the expensive code runs
//     with full warps and a single instance of it exists in the kernel.  Inside, a lane walks
//     its own (body, sub-sample) pairs that pass the window test, so lanes working on different
//     bodies still execute together.
// Outputs: flux / flux_sum rows are written as zeros by the owning lane for samples that were
// not queued, and by the processing lane for queued ones; the likelihood uses a segmented
// warp-shuffle reduction keyed by set into a per-warp accumulator, written as one partial per
// (set, warp tile) and summed in fixed order by k_loglike_final (deterministic).
// 5 CTAs of 128 threads per SM = 96 registers per thread: with 64 (8 CTAs) ptxas serialises the
// independent chains of the node loop to save registers and spills inside it; measured config 3
// 1.09 -> 1.00 ms, config 4 26.6 -> 19.8 ms (gpurun_out/variants*.log)
#ifndef JXP_SYS_MINB
#define JXP_SYS_MINB 5
#endif
constexpr int SYS_QCAP = 256;  // queue entries per warp
constexpr int SYS_MB = 8;       // bodies the multi variant of stage C supports

template <typename T>
__device__ __forceinline__ bool phase_in_window(const BodyC<T>& c, double tm) {
  const double ph = fma(tm - c.t0ref, c.inv_period, -c.phase_c);
  const double d = ph - rint(ph);
  return (d >= c.wlo) & (d <= c.whi);
}

template <typename T, int ORDER, bool LOGLIKE, bool DENSE, bool HI = false, int PU = 1>
__global__ void __launch_bounds__(128, JXP_SYS_MINB) k_system(const __grid_constant__ SysArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_pick[2];
  constexpr int K = SYS_K;
  const int nthreads = blockDim.x;
  const int nwarps = nthreads >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = a.n_bodies;
  const int n_groups = a.n_groups;
  if (a.tl_state != nullptr && tl_active(a.tl_state, a.tl_cap)) return;  // the transit-list kernels did the work

  BodyC<T>* sbody = reinterpret_cast<BodyC<T>*>(smem_raw);
  SetC<T>* sset = reinterpret_cast<SetC<T>*>(smem_raw + a.so_sset);
  double* sacc = reinterpret_cast<double*>(smem_raw + a.so_sacc) + warp * a.sets_per_cta;  // [nwarps][sets_per_cta]
  unsigned short* sq = reinterpret_cast<unsigned short*>(smem_raw + a.so_sq) + warp * SYS_QCAP;
  // dense variant: sky separation of the queued in-transit samples
  T* sqb = reinterpret_cast<T*>(smem_raw + a.so_sqb) + warp * SYS_QCAP;
  // multi variant (PU == 5): per-warp evaluation queue and per-(queued sample, body) sums
  constexpr bool MULTI = (PU == 5);
  T* macc = reinterpret_cast<T*>(smem_raw + a.so_macc) + warp * (32 * SYS_MB);
  unsigned short* meq = reinterpret_cast<unsigned short*>(smem_raw + a.so_meq) + warp * 64;
  // multi variant, double, quadratic law, order 10: the in-transit (sample, body, sub-sample) triples are
  // sorted by flux code after the position stage (see "class queues" below).  Two queues of (b, meta)
  // alias the dense variant's sqb (unused here); the edge class has its own per-(sample, body) sums.
  constexpr bool CLS = MULTI && !HI && ORDER == 11 && std::is_same<T, double>::value;
  double* cqb = reinterpret_cast<double*>(sqb);                // [2][64]
  unsigned* cqm = reinterpret_cast<unsigned*>(cqb + 2 * 64);   // [2][64]
  T* maccE = reinterpret_cast<T*>(smem_raw + a.so_maccE) + warp * (32 * nb);  // [32][nb]

  const int nsub = a.st.n;
  const double dtmin = a.st.dt_min, dtmax = a.st.dt_max;
  const bool want_zero = (a.flux != nullptr) || (a.flux_sum != nullptr);
  unsigned n_kep = 0, n_ld = 0;
  int* const group_ctr = a.tile_ctr;     // next unclaimed set group
  int* const tile_ctr = a.tile_ctr + 1;  // per group: next warp tile

  // Persistent CTAs.  Phase 1: claim whole set groups from a global counter.  Phase 2 (no
  // unclaimed group left): scan the groups cyclically from a home position and help with the
  // ones that still have warp tiles.  Inside a group every WARP pulls its next tile from the
  // group's counter, so a warp that drew light tiles simply processes more of them.
  int scan = -1;  // -1: claiming; >= 0: groups scanned so far in phase 2
  const int home = (int)(blockIdx.x % (unsigned)n_groups);
  for (int iter = 0;; ++iter) {
    int group;
    if (scan < 0) {
      if (threadIdx.x == 0) s_pick[iter & 1] = atomicAdd(group_ctr, 1);
      __syncthreads();
      group = s_pick[iter & 1];
      if (group >= n_groups) {
        scan = 0;
        continue;
      }
    } else {
      if (scan >= n_groups) break;
      // every thread looks at one group; the first one with tiles left is taken
      const int gi = scan + (int)threadIdx.x;
      bool has = false;
      if (gi < n_groups) {
        int g = home + gi;
        g -= (g >= n_groups) ? n_groups : 0;
        has = *reinterpret_cast<volatile int*>(tile_ctr + g) < a.n_wtiles;
      }
      if (threadIdx.x == 0) s_pick[iter & 1] = 0x7fffffff;
      __syncthreads();
      const unsigned hb = __ballot_sync(0xffffffffu, has);
      if (hb && lane == 0) atomicMin(&s_pick[iter & 1], warp * 32 + __ffs(hb) - 1);
      __syncthreads();
      const int first = s_pick[iter & 1];
      if (first == 0x7fffffff) {
        scan += nthreads;
        continue;
      }
      group = home + scan + first;
      group -= (group >= n_groups) ? n_groups : 0;
      scan += first + 1;
    }
    const int set_base = group * a.sets_per_cta;
    const int nsl = min(a.sets_per_cta, a.n_sets - set_base);
    {
      // stage the group's records (plain 8-byte copies; sizeof(BodyC), sizeof(SetC) % 8 == 0)
      const unsigned long long* src =
          reinterpret_cast<const unsigned long long*>(a.bodies + (size_t)set_base * nb);
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(sbody);
      const int nw = nsl * nb * (int)(sizeof(BodyC<T>) / 8);
      for (int i = threadIdx.x; i < nw; i += nthreads) dst[i] = src[i];
      const unsigned long long* src2 = reinterpret_cast<const unsigned long long*>(a.sets + set_base);
      unsigned long long* dst2 = reinterpret_cast<unsigned long long*>(sset);
      const int nw2 = nsl * (int)(sizeof(SetC<T>) / 8);
      for (int i = threadIdx.x; i < nw2; i += nthreads) dst2[i] = src2[i];
    }
    if (a.use_window == 0) {
      // dense mode: mark every staged body "always" so that the hot loops test one flag only
      __syncthreads();
      for (int i = threadIdx.x; i < nsl * nb; i += nthreads) sbody[i].flags |= BODY_ALWAYS;
    }
    __syncthreads();

    // ---- warp tiles of this group (no block barriers inside) -----------------------------
    int wnext = 0;
    if (lane == 0) wnext = atomicAdd(tile_ctr + group, 1);
    while (true) {
    const int wtile_i = __shfl_sync(0xffffffffu, wnext, 0);
    if (wtile_i >= a.n_wtiles) break;
    if (lane == 0) wnext = atomicAdd(tile_ctr + group, 1);  // in flight while this tile runs
    // this warp's samples: sample k of lane l is wbase + 32 k + l
    const int64_t wtile = wtile_i;
    const int64_t wbase = wtile * (32 * K);
    double tk[K];
    unsigned valid = 0;
    double tmin = 1e300, tmax = -1e300;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int64_t idx = wbase + 32 * k + lane;
      const bool ok = idx < a.n_times;
      tk[k] = ok ? a.t[idx] : 0.0;
      valid |= ok ? (1u << k) : 0u;
      if (ok) {
        tmin = fmin(tmin, tk[k]);
        tmax = fmax(tmax, tk[k]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tmin = fmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
      tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    const double span_lo = tmin + dtmin, span = (tmax + dtmax) - span_lo;
    int qn = 0;  // queue fill (warp-uniform)
    if (LOGLIKE) {
      for (int i = lane; i < a.sets_per_cta; i += 32) sacc[i] = 0.0;
      __syncwarp();
    }

  if constexpr (DENSE) {
    // ---- dense variant (JXP_FLAG_NO_WINDOW, one body, no exposure integration) ----------------
    // Every sample runs the Kepler solve + position directly from registers (full warps, like
    // K1); samples found in transit are compacted with their separation b into the per-warp
    // queue and the flux code then runs on 32 in-transit samples at a time.
    int sl = 0, k = 0;
    while (true) {
      bool done = false;
      while (qn <= SYS_QCAP - 32) {
        if (sl >= nsl) {
          done = true;
          break;
        }
        const BodyC<T>& c = sbody[sl];
        const double tm = pick(tk, k);
        const bool ok = (valid >> k) & 1u;
        T b = T(0), Z = T(-1);
        if (ok) {
          ++n_kep;
          position_stage(c, tm, b, Z);
        }
        const bool hit = ok && (Z > T(0)) && !(b >= c.onepr);
        const unsigned mk = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const int pos = qn + __popc(mk & ((1u << lane) - 1u));
          sq[pos] = (unsigned short)((sl << 7) | (k << 5) | lane);
          sqb[pos] = b;
        }
        qn += __popc(mk);
        if (want_zero && ok && !hit) {
          const int64_t idx = wbase + 32 * k + lane;
          if (a.flux) a.flux[(size_t)(set_base + sl) * a.n_times + idx] = T(0);
          if (a.flux_sum) a.flux_sum[(size_t)(set_base + sl) * a.n_times + idx] = T(0);
        }
        if (++k == K) {
          k = 0;
          ++sl;
        }
      }
      __syncwarp();
#pragma unroll 1
      for (int q0 = 0; q0 < qn; q0 += 32) {
        const bool have = q0 + lane < qn;
        const unsigned item = have ? sq[q0 + lane] : 0u;
        const int isl = (int)(item >> 7);
        const int64_t idx = wbase + 32 * (int)((item >> 5) & 3u) + (int)(item & 31u);
        T total = T(0);
        if (have) {
          const SetC<T> sc = sset[isl];
          ++n_ld;
          T s0, s1, s2;
          ld_solution<T, ORDER>(sqb[q0 + lane], sbody[isl].rr, sc.lmax, &a.tab, s0, s1, s2);
          total = ld_combine(sc.lmax, s0, s1, s2, sc.g0, sc.g1, sc.g2);
          if (a.flux) a.flux[(size_t)(set_base + isl) * a.n_times + idx] = total;
          if (a.flux_sum) a.flux_sum[(size_t)(set_base + isl) * a.n_times + idx] = total;
        }
        if (LOGLIKE) {
          double dchi = 0.0;
          if (have && total != T(0)) {
            const T w = a.winv[idx];
            const T r0 = (a.y[idx] - T(1)) * w;
            const T mw = total * w;
            dchi = (double)(-mw) * (double)(T(2) * r0 - mw);
          }
          const int key = have ? isl : -1 - lane;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, dchi, o);
            const int ku = __shfl_up_sync(0xffffffffu, key, o);
            if (lane >= o && ku == key) dchi += u;
          }
          const int kn = __shfl_down_sync(0xffffffffu, key, 1);
          if (have && (lane == 31 || kn != key)) sacc[isl] += dchi;
          __syncwarp();
        }
      }
      qn = 0;
      __syncwarp();
      if (done) break;
    }
  } else {
  // fill (stages A, B) / drain (stage C) loop; ONE instance of the drain code
  int c0 = 0, c0_cur = 0;
  unsigned visit = 0, cmask = 0;
  while (true) {
    bool done = false;
    while (qn <= SYS_QCAP - 32 * K) {  // room for one more set (<= 32 K hits)
      if (!visit) {
        if (c0 >= nsl) {
          done = true;
          break;
        }
        // ---- stage A: lane = set --------------------------------------------------------
        bool cand = false;
        const int sl = c0 + lane;
        if (sl < nsl) {
          for (int b = 0; b < nb; ++b) {
            const BodyC<T>& c = sbody[sl * nb + b];
            if (c.flags & BODY_ALWAYS) {
              cand = true;
            } else {
              const double ph0 = fma(span_lo - c.t0ref, c.inv_period, -c.phase_c);
              const double d0 = ph0 - rint(ph0);
              const double d1 = fma(span, c.inv_period, d0);
              // the span lies before this cycle's window, or between it and the next one
              const bool miss = (d1 < c.wlo) | ((d0 > c.whi) & (d1 < 1.0 + c.wlo));
              cand = cand || !miss;
            }
          }
        }
        cmask = __ballot_sync(0xffffffffu, cand);
        const int nchunk = min(32, nsl - c0);
        // in flux modes every set of the chunk is visited (zero fill)
        visit = want_zero ? ((nchunk == 32) ? 0xffffffffu : ((1u << nchunk) - 1u)) : cmask;
        c0_cur = c0;
        c0 += 32;
        continue;
      }
      // ---- stage B: lane = sample, one set -------------------------------------------------
      const int s_in = __ffs(visit) - 1;
      visit &= visit - 1;
      const int sl = c0_cur + s_in;
      unsigned hits = 0;
      if (cmask & (1u << s_in)) {
        for (int b = 0; b < nb; ++b) {
          const BodyC<T>& c = sbody[sl * nb + b];
          if (c.flags & BODY_ALWAYS) {
            hits = valid;
          } else {
            // widened by the exposure stencil: a superset of the samples with any sub-sample hit
            const double wlo = fma(-dtmax, c.inv_period, c.wlo), whi = fma(-dtmin, c.inv_period, c.whi);
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const double ph = fma(tk[k] - c.t0ref, c.inv_period, -c.phase_c);
              const double d = ph - rint(ph);
              hits |= ((d >= wlo) & (d <= whi)) ? (1u << k) : 0u;
            }
          }
        }
        hits &= valid;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const unsigned mk = __ballot_sync(0xffffffffu, (hits >> k) & 1u);
          if ((hits >> k) & 1u)
            sq[qn + __popc(mk & ((1u << lane) - 1u))] = (unsigned short)((sl << 7) | (k << 5) | lane);
          qn += __popc(mk);
        }
      }
      if (want_zero) {
        const int set = set_base + sl;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if ((valid & ~hits) & (1u << k)) {
            const int64_t idx = wbase + 32 * k + lane;
            if (a.flux) {
#pragma unroll 1
              for (int b = 0; b < nb; ++b) a.flux[((size_t)set * a.n_times + idx) * nb + b] = T(0);
            }
            if (a.flux_sum) a.flux_sum[(size_t)set * a.n_times + idx] = T(0);
          }
        }
      }
    }

    // ---- stage C: drain the queue, 32 in-window samples at a time ---------------------------
    // Only FULL groups of 32 are drained while more samples may still arrive; the remainder
    // (< 32 items) stays at the front of the queue, so partial warps run once per warp tile.
    __syncwarp();
    const int q_end = done ? qn : (qn & ~31);
#pragma unroll 1
    for (int q0 = 0; q0 < q_end; q0 += 32) {
      const bool have = q0 + lane < qn;
      const unsigned item = have ? sq[q0 + lane] : 0u;
      const int sl = (int)(item >> 7);
      const int k = (int)((item >> 5) & 3u);
      const int src = (int)(item & 31u);
      // fetch the sample's time stamp from the owning lane
      double tm0 = 0.0;
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        const double v = __shfl_sync(0xffffffffu, tk[kk], src);
        tm0 = (kk == k) ? v : tm0;
      }
      const int64_t idx = wbase + 32 * k + src;
      T total = T(0);
      if constexpr (MULTI) {
        // Several (body, sub-sample) evaluations per queued sample (multi-planet systems, exposure
        // integration).  Lanes own different numbers of them -- at the edge of an exposure-widened
        // window only a few sub-samples are in -- so a per-lane walk leaves most lanes idle
        // (17.9 of 32 active on config 4).  Instead the (sample, body, sub-sample) triples that
        // pass the window test are compacted into a second queue and evaluated 32 at a time, one
        // per lane; the weighted values are added into per-(sample, body) sums in shared memory
        // in queue order (= increasing sub-sample index, the order of the reference's dot product),
        // lanes with the same target taking turns by lane index: deterministic.
        const BodyC<T>* bodies = sbody + sl * nb;
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
          macc[lane * SYS_MB + b] = T(0);
          if (CLS) maccE[lane * nb + b] = T(0);
        }
        __syncwarp();
        const unsigned lt = (1u << lane) - 1u;
        int cqn[2] = {0, 0};  // class queue fills (warp-uniform)
        const int nsteps = nb * nsub;
        int en = 0, b = 0, j = 0;
        bool bpass = false, always = false;
#pragma unroll 1
        for (int step = 0; step <= nsteps; ++step) {
          if (step < nsteps) {
            if (j == 0) {
              // body-level test against the exposure-widened window, once per body
              bpass = false;
              always = false;
              if (have) {
                const BodyC<T>& c = bodies[b];
                always = (c.flags & BODY_ALWAYS) != 0;
                const double ph = fma(tm0 - c.t0ref, c.inv_period, -c.phase_c);
                const double d = ph - rint(ph);
                bpass = always || !((d < fma(-dtmax, c.inv_period, c.wlo)) | (d > fma(-dtmin, c.inv_period, c.whi)));
              }
              if (!__any_sync(0xffffffffu, bpass)) {
                step += nsub - 1;
                ++b;
                continue;
              }
            }
            const bool pass = bpass && (always || nsub == 1 || phase_in_window(bodies[b], tm0 + a.st.dt[j]));
            const unsigned mk = __ballot_sync(0xffffffffu, pass);
            if (pass) meq[en + __popc(mk & lt)] = (unsigned short)((b << 11) | (j << 5) | lane);
            en += __popc(mk);
            if (++j == nsub) {
              j = 0;
              ++b;
            }
            __syncwarp();
          }
          if (en >= 32 || (step == nsteps && (en > 0 || (CLS && (cqn[0] | cqn[1]) != 0)))) {
            const int nbatch = en < 32 ? en : 32;
            const bool ev = lane < nbatch;
            const unsigned it = ev ? meq[lane] : 0u;
            const int esrc = (int)(it & 31u), ej = (int)((it >> 5) & 63u), eb = (int)(it >> 11);
            const double tms = __shfl_sync(0xffffffffu, tm0, esrc);
            const int sls = __shfl_sync(0xffffffffu, sl, esrc);
            if constexpr (CLS) {
              // Class queues.  Position stage for the 32 triples; the ones in transit go to one of two
              // queues by the flux code they need -- planet fully inside the disk (ld_solution_inside_n:
              // constant kappas and node cosines, 2.1x the rate of the general code) or not -- and a
              // queue is evaluated 32 entries at a time, first in first out.  Each class adds into its
              // own per-(sample, body) sum in increasing sub-sample order whatever the grouping, and the
              // two sums are added at the end: the result does not depend on which samples share a drain
              // (windowed and dense evaluation stay bit-identical).
              int cls = -1;
              double bsep = 0.0;
              if (ev) {
                ++n_kep;
                const BodyC<T>& c = sbody[sls * nb + eb];
                const double Mr = mean_anomaly_reduced(c, tms + a.st.dt[ej]);
                double sf1[1], cf1[1];  // (sin f, cos f) D: the position needs no division (sky_position_xy)
                if (c.flags & BODY_CIRCULAR) {
                  jsincos(Mr, &sf1[0], &cf1[0]);
                } else {
                  const double Mr1[1] = {Mr}, e1[1] = {c.e}, ome1[1] = {c.ome}, ope1[1] = {c.ope}, s11[1] = {c.s1me2},
                               io1[1] = {c.inv_ope};
                  kepler_reduced_n<1, true, true>(Mr1, e1, ome1, ope1, s11, io1, sf1, cf1);
                }
                double Z;
                sky_position_xy(c, sf1[0], cf1[0], bsep, Z);
                if ((Z > 0.0) && !(bsep >= c.onepr)) {
                  const double ar = fabs(c.rr);
                  cls = ((bsep < 1.0 - ar) && (ar < 1.0)) ? 0 : 1;
                  ++n_ld;
                }
              }
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                const unsigned m = __ballot_sync(0xffffffffu, cls == cc);
                if (cls == cc) {
                  const int e = cqn[cc] + __popc(m & lt);
                  cqb[cc * 64 + e] = bsep;
                  cqm[cc * 64 + e] = (unsigned)esrc | ((unsigned)eb << 5) | ((unsigned)ej << 8) | ((unsigned)sls << 14);
                }
                cqn[cc] += __popc(m);
              }
              __syncwarp();
              const bool last = (step == nsteps) && (en - nbatch == 0);
#pragma unroll 1
              for (int cc = 0; cc < 2; ++cc) {
                while (cqn[cc] >= 32 || (last && cqn[cc] > 0)) {
                  const int nb2 = cqn[cc] < 32 ? cqn[cc] : 32;
                  const bool ev2 = lane < nb2;
                  const unsigned meta = cqm[cc * 64 + (ev2 ? lane : 0)];
                  const int tgt2 = ev2 ? (int)(meta & 255u) : (-1 - lane);  // (queued sample, body)
                  const int ej2 = (int)((meta >> 8) & 63u), sls2 = (int)(meta >> 14), eb2 = (int)((meta >> 5) & 7u);
                  double bb1[1] = {cqb[cc * 64 + (ev2 ? lane : 0)]}, rr1[1] = {(double)sbody[sls2 * nb + eb2].rr};
                  double s0[1], s1[1], s2[1];
                  const SetC<T>& sc2 = sset[sls2];
                  if (cc == 0)
                    ld_solution_inside_n<1>(bb1, rr1, sc2.lmax, s0, s1, s2);
                  else
                    ld_solution<double, ORDER, false, PU>(bb1[0], rr1[0], sc2.lmax, &a.tab, s0[0], s1[0], s2[0]);
                  const double v2 = ld_combine(sc2.lmax, s0[0], s1[0], s2[0], (double)sc2.g0, (double)sc2.g1, (double)sc2.g2);
                  T* acc2 = cc == 0 ? macc : maccE;
                  const int slot = ev2 ? ((int)(meta & 31u) * (cc == 0 ? SYS_MB : nb) + eb2) : 0;
                  const unsigned same = __match_any_sync(0xffffffffu, tgt2);
                  const int rank = __popc(same & lt);
                  const int maxr = (int)__reduce_max_sync(0xffffffffu, (unsigned)(ev2 ? rank : 0));
#pragma unroll 1
                  for (int r = 0; r <= maxr; ++r) {
                    if (ev2 && rank == r) acc2[slot] = jfma(T(a.st.w[ej2]), T(v2), acc2[slot]);
                    __syncwarp();
                  }
                  // first in, first out: move the rest of the queue to the front
                  const int rem2 = cqn[cc] - nb2;
                  const double kb = (lane < rem2) ? cqb[cc * 64 + 32 + lane] : 0.0;
                  const unsigned km = (lane < rem2) ? cqm[cc * 64 + 32 + lane] : 0u;
                  __syncwarp();
                  if (lane < rem2) {
                    cqb[cc * 64 + lane] = kb;
                    cqm[cc * 64 + lane] = km;
                  }
                  cqn[cc] = rem2;
                  __syncwarp();
                }
              }
            } else {
            T v = T(0);
            if (ev) {
              ++n_kep;
              v = eval_in_window<T, ORDER, HI, PU>(sbody[sls * nb + eb], sset[sls], &a.tab, tms + a.st.dt[ej], n_ld);
            }
            const int tgt = ev ? (esrc * SYS_MB + eb) : (-1 - lane);
            const unsigned same = __match_any_sync(0xffffffffu, tgt);
            const int rank = __popc(same & lt);
            const int maxr = (int)__reduce_max_sync(0xffffffffu, (unsigned)(ev ? rank : 0));
#pragma unroll 1
            for (int r = 0; r <= maxr; ++r) {
              if (ev && rank == r) macc[tgt] = (nsub == 1) ? v : jfma(T(a.st.w[ej]), v, macc[tgt]);
              __syncwarp();
            }
            }
            // keep what did not fit into this batch
            const int rem = en - nbatch;
            const unsigned short keep = (lane < rem) ? meq[32 + lane] : (unsigned short)0;
            __syncwarp();
            if (lane < rem) meq[lane] = keep;
            en = rem;
            __syncwarp();
          }
        }
        if (have) {
          const int set = set_base + sl;
#pragma unroll 1
          for (int bb = 0; bb < nb; ++bb) {
            const T accb = CLS ? macc[lane * SYS_MB + bb] + maccE[lane * nb + bb] : macc[lane * SYS_MB + bb];
            if (a.flux) a.flux[((size_t)set * a.n_times + idx) * nb + bb] = accb;
            total += accb;
          }
          if (a.flux_sum) a.flux_sum[(size_t)set * a.n_times + idx] = total;
        }
        __syncwarp();
      } else if (have) {
        const int set = set_base + sl;
        const SetC<T> sc = sset[sl];
        const BodyC<T>* bodies = sbody + sl * nb;
        int nxt_b = 0, nxt_j = 0, cur_b = 0;
        T accb = T(0);
        // walk this sample's (body, sub-sample) pairs that pass the window test: a body whose
        // exposure-widened window misses the sample is skipped with one test
#pragma unroll 1
        while (true) {
          int b = 0, j = 0;
          bool found = false;
          while (nxt_b < nb) {
            const BodyC<T>& c = bodies[nxt_b];
            const bool always = (c.flags & BODY_ALWAYS) != 0;
            if (nxt_j == 0 && !always) {
              const double ph = fma(tm0 - c.t0ref, c.inv_period, -c.phase_c);
              const double d = ph - rint(ph);
              if ((d < fma(-dtmax, c.inv_period, c.wlo)) | (d > fma(-dtmin, c.inv_period, c.whi))) {
                ++nxt_b;
                continue;
              }
            }
            b = nxt_b;
            j = nxt_j;
            if (++nxt_j == nsub) {
              nxt_j = 0;
              ++nxt_b;
            }
            if (always || nsub == 1 || phase_in_window(c, tm0 + a.st.dt[j])) {
              found = true;
              break;
            }
          }
          if (!found) break;
          if (b != cur_b) {
            if (a.flux)
              for (int bb = cur_b; bb < b; ++bb)
                a.flux[((size_t)set * a.n_times + idx) * nb + bb] = (bb == cur_b) ? accb : T(0);
            total += accb;
            accb = T(0);
            cur_b = b;
          }
          ++n_kep;
          const T v = eval_in_window<T, ORDER, HI, PU>(bodies[b], sc, &a.tab, tm0 + a.st.dt[j], n_ld);
          accb = (nsub == 1) ? v : jfma(T(a.st.w[j]), v, accb);
        }
        if (a.flux)
          for (int bb = cur_b; bb < nb; ++bb)
            a.flux[((size_t)set * a.n_times + idx) * nb + bb] = (bb == cur_b) ? accb : T(0);
        total += accb;
        if (a.flux_sum) a.flux_sum[(size_t)set * a.n_times + idx] = total;
      }
      if (LOGLIKE) {
        // (r0 - m w)^2 - r0^2 with r0 = (y - 1) w: the change of chi^2 w.r.t. the zero model
        double dchi = 0.0;
        if (have && total != T(0)) {
          const T w = a.winv[idx];
          const T r0 = (a.y[idx] - T(1)) * w;
          const T mw = total * w;
          dchi = (double)(-mw) * (double)(T(2) * r0 - mw);
        }
        // segmented inclusive scan keyed by set (queue entries of one set are contiguous)
        const int key = have ? sl : -1 - lane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double u = __shfl_up_sync(0xffffffffu, dchi, o);
          const int ku = __shfl_up_sync(0xffffffffu, key, o);
          if (lane >= o && ku == key) dchi += u;
        }
        const int kn = __shfl_down_sync(0xffffffffu, key, 1);
        if (have && (lane == 31 || kn != key)) sacc[sl] += dchi;
        __syncwarp();
      }
    }
    {
      const int rem = qn - q_end;  // 0 when done
      const unsigned short keep = (lane < rem) ? sq[q_end + lane] : (unsigned short)0;
      __syncwarp();
      if (lane < rem) sq[lane] = keep;
      qn = rem;
    }
    __syncwarp();
    if (done) break;
  }
  }  // !DENSE

    if (LOGLIKE) {
      for (int sl = lane; sl < nsl; sl += 32)
        a.partial[(size_t)wtile * a.n_sets + (set_base + sl)] = sacc[sl];
      __syncwarp();
    }
    }  // warp tiles
    __syncthreads();  // the staged records are replaced in the next iteration
  }  // set groups
  // executed-work counters (two atomics per warp)
  n_kep = __reduce_add_sync(0xffffffffu, n_kep);
  n_ld = __reduce_add_sync(0xffffffffu, n_ld);
  if (lane == 0) {
    atomicAdd(a.counters + 0, (unsigned long long)n_kep);
    atomicAdd(a.counters + 1, (unsigned long long)n_ld);
  }
}

// k_system_pair -- the log-likelihood of single-planet systems without exposure integration
// (config 3) with TWO queued samples per lane.
//
// Same persistent-CTA / set-group / warp-tile skeleton and the same stages A and B as k_system.
// What changes is stage C: the fused kernel is latency-bound, not FP64-issue bound (FP64 pipe 49 %
// active, `wait` the top stall at ~5 warps per scheduler), so every lane evaluates two queued
// samples at once through kepler_reduced_n<2> / ld_solution_n<2>, whose statements interleave the
// two dependency chains.  To keep 64-wide drains full, queue items carry the absolute sample index
// and what is left (< 64 items) is carried over to the warp's next tile; the last partial drain
// happens once per warp and set group instead of once per tile.
// Determinism with carried-over items: the chi^2 terms of one (tile, set) are contiguous in the
// queue; they are added SEQUENTIALLY in queue order by one lane, continuing from a running sum when
// a segment is split across drains, and the total is stored once, when the segment closes -- so
// the result does not depend on how the dynamic tile assignment happened to split the segments.
constexpr int SP_QCAP = 256;
#ifndef JXP_PAIR_MINB
#define JXP_PAIR_MINB 4
#endif
template <int PU>
__global__ void __launch_bounds__(128, JXP_PAIR_MINB) k_system_pair(const __grid_constant__ SysArgs<double> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_pick[2];
  using T = double;
  constexpr int K = SYS_K;
  const int nthreads = blockDim.x;
  const int nwarps = nthreads >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_groups = a.n_groups;
  BodyC<T>* sbody = reinterpret_cast<BodyC<T>*>(smem_raw);
  SetC<T>* sset = reinterpret_cast<SetC<T>*>(sbody + a.sets_per_cta);
  double* sdchi = reinterpret_cast<double*>(sset + a.sets_per_cta) + warp * 66;  // 64 terms + open sum
  unsigned* sq = reinterpret_cast<unsigned*>(reinterpret_cast<double*>(sset + a.sets_per_cta) + nwarps * 66) +
                 warp * (SP_QCAP + 64 + 2);
  unsigned* skey = sq + SP_QCAP;  // keys of the 64 items of the current drain, then the open key
  unsigned n_kep = 0, n_ld = 0;
  int* const group_ctr = a.tile_ctr;
  int* const tile_ctr = a.tile_ctr + 1;
  const unsigned lt = (1u << lane) - 1u;
  if (a.tl_state != nullptr && tl_active(a.tl_state, a.tl_cap)) return;  // the transit-list kernels did the work

  int scan = -1;
  const int home = (int)(blockIdx.x % (unsigned)n_groups);
  for (int iter = 0;; ++iter) {
    int group;
    if (scan < 0) {
      if (threadIdx.x == 0) s_pick[iter & 1] = atomicAdd(group_ctr, 1);
      __syncthreads();
      group = s_pick[iter & 1];
      if (group >= n_groups) {
        scan = 0;
        continue;
      }
    } else {
      if (scan >= n_groups) break;
      const int gi = scan + (int)threadIdx.x;
      bool has = false;
      if (gi < n_groups) {
        int g = home + gi;
        g -= (g >= n_groups) ? n_groups : 0;
        has = *reinterpret_cast<volatile int*>(tile_ctr + g) < a.n_wtiles;
      }
      if (threadIdx.x == 0) s_pick[iter & 1] = 0x7fffffff;
      __syncthreads();
      const unsigned hb = __ballot_sync(0xffffffffu, has);
      if (hb && lane == 0) atomicMin(&s_pick[iter & 1], warp * 32 + __ffs(hb) - 1);
      __syncthreads();
      const int first = s_pick[iter & 1];
      if (first == 0x7fffffff) {
        scan += nthreads;
        continue;
      }
      group = home + scan + first;
      group -= (group >= n_groups) ? n_groups : 0;
      scan += first + 1;
    }
    const int set_base = group * a.sets_per_cta;
    const int nsl = min(a.sets_per_cta, a.n_sets - set_base);
    {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.bodies + (size_t)set_base);
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(sbody);
      const int nw = nsl * (int)(sizeof(BodyC<T>) / 8);
      for (int i = threadIdx.x; i < nw; i += nthreads) dst[i] = src[i];
      const unsigned long long* src2 = reinterpret_cast<const unsigned long long*>(a.sets + set_base);
      unsigned long long* dst2 = reinterpret_cast<unsigned long long*>(sset);
      const int nw2 = nsl * (int)(sizeof(SetC<T>) / 8);
      for (int i = threadIdx.x; i < nw2; i += nthreads) dst2[i] = src2[i];
    }
    __syncthreads();

    int qn = 0;                       // queue fill (warp-uniform)
    unsigned open_key = 0xffffffffu;  // (tile, set) segment still open at the end of the last drain
    int wnext = 0;
    if (lane == 0) wnext = atomicAdd(tile_ctr + group, 1);
    // fill state (warp-uniform except tk / valid)
    bool have_tile = false, more = true;
    int64_t wbase = 0;
    double tk[K];
    unsigned valid = 0, visit = 0;
    double span_lo = 0.0, span = 0.0;
    int c0 = 0, c0_cur = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) tk[k] = 0.0;
    while (true) {
      // ---- fill: stages A and B until 64 items are queued or the group has no tile left -------
      while (more && qn < 64) {
        if (!have_tile) {
          const int wtile_i = __shfl_sync(0xffffffffu, wnext, 0);
          if (wtile_i >= a.n_wtiles) {
            more = false;
            break;
          }
          if (lane == 0) wnext = atomicAdd(tile_ctr + group, 1);  // in flight while this tile runs
          wbase = (int64_t)wtile_i * (32 * K);
          valid = 0;
          double tmin = 1e300, tmax = -1e300;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const int64_t idx = wbase + 32 * k + lane;
            const bool ok = idx < a.n_times;
            tk[k] = ok ? a.t[idx] : 0.0;
            valid |= ok ? (1u << k) : 0u;
            if (ok) {
              tmin = fmin(tmin, tk[k]);
              tmax = fmax(tmax, tk[k]);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            tmin = fmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
            tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
          }
          span_lo = tmin;
          span = tmax - tmin;
          // zero partial sums of this tile (closed segments overwrite theirs later)
          for (int sl = lane; sl < nsl; sl += 32) a.partial[(size_t)wtile_i * a.n_sets + (set_base + sl)] = 0.0;
          __syncwarp();
          c0 = 0;
          visit = 0;
          have_tile = true;
        }
        if (!visit) {
          if (c0 >= nsl) {
            have_tile = false;
            continue;
          }
          // stage A: lane = set
          bool cand = false;
          const int sla = c0 + lane;
          if (sla < nsl) {
            const BodyC<T>& c = sbody[sla];
            if (c.flags & BODY_ALWAYS) {
              cand = true;
            } else {
              const double ph0 = fma(span_lo - c.t0ref, c.inv_period, -c.phase_c);
              const double d0 = ph0 - rint(ph0);
              const double d1 = fma(span, c.inv_period, d0);
              cand = !((d1 < c.wlo) | ((d0 > c.whi) & (d1 < 1.0 + c.wlo)));
            }
          }
          visit = __ballot_sync(0xffffffffu, cand);
          c0_cur = c0;
          c0 += 32;
          continue;
        }
        // stage B: lane = sample, one candidate set
        const int s_in = __ffs(visit) - 1;
        visit &= visit - 1;
        const int sl = c0_cur + s_in;
        const BodyC<T>& c = sbody[sl];
        unsigned hits = 0;
        if (c.flags & BODY_ALWAYS) {
          hits = valid;
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const double ph = fma(tk[k] - c.t0ref, c.inv_period, -c.phase_c);
            const double d = ph - rint(ph);
            hits |= ((d >= c.wlo) & (d <= c.whi)) ? (1u << k) : 0u;
          }
          hits &= valid;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const unsigned mk = __ballot_sync(0xffffffffu, (hits >> k) & 1u);
          if ((hits >> k) & 1u)
            sq[qn + __popc(mk & lt)] = ((unsigned)sl << 25) | (unsigned)(wbase + 32 * k + lane);
          qn += __popc(mk);
        }
        __syncwarp();
      }
      if (qn == 0) break;  // no tile left and nothing queued

      // ---- stage C: 64 queued samples (fewer only at the very end of the group), two per lane ----
      const int cnt = qn < 64 ? qn : 64;
      bool have[2];
      double Mr[2], e[2], ome[2], ope[2], s1[2], iope[2], sf[2], cf[2], bq[2], rq[2], tot[2];
      int slq[2];
      int64_t idxq[2];
      bool circ = false;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        have[h] = 32 * h + lane < cnt;
        const unsigned item = have[h] ? sq[32 * h + lane] : sq[0];
        slq[h] = (int)(item >> 25);
        idxq[h] = (int64_t)(item & 0x1ffffffu);
        const BodyC<T>& c = sbody[slq[h]];
        Mr[h] = mean_anomaly_reduced(c, a.t[idxq[h]]);
        e[h] = c.e;
        ome[h] = c.ome;
        ope[h] = c.ope;
        s1[h] = c.s1me2;
        iope[h] = c.inv_ope;
        circ = circ || (c.flags & BODY_CIRCULAR);
      }
      kepler_reduced_n<2>(Mr, e, ome, ope, s1, iope, sf, cf);
      if (__any_sync(0xffffffffu, circ)) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (sbody[slq[h]].flags & BODY_CIRCULAR) jsincos(Mr[h], &sf[h], &cf[h]);  // keplerian.py:539-540
      }
      bool intr[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const BodyC<T>& c = sbody[slq[h]];
        double Z;
        sky_position(c, sf[h], cf[h], bq[h], Z);
        rq[h] = c.rr;
        intr[h] = have[h] && (Z > 0.0) && !(bq[h] >= c.onepr);
        n_kep += have[h] ? 1u : 0u;
        n_ld += intr[h] ? 1u : 0u;
      }
      {
        double s0[2], s1v[2], s2[2];
        const int lmax = sset[slq[0]].lmax;  // one law length per call
        ld_solution_n<2, PU>(bq, rq, lmax, s0, s1v, s2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const SetC<T>& sc = sset[slq[h]];
          tot[h] = intr[h] ? ld_combine(lmax, s0[h], s1v[h], s2[h], sc.g0, sc.g1, sc.g2) : 0.0;
        }
      }
      // chi^2 terms and (tile, set) keys to shared memory
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double dchi = 0.0;
        if (have[h] && tot[h] != 0.0) {
          const double w = a.winv[idxq[h]];
          const double r0 = (a.y[idxq[h]] - 1.0) * w;
          const double mw = tot[h] * w;
          dchi = (-mw) * (2.0 * r0 - mw);
        }
        sdchi[32 * h + lane] = dchi;
        skey[32 * h + lane] =
            have[h] ? (((unsigned)slq[h] << 25) | (unsigned)(idxq[h] / (32 * K))) : 0xfffffffeu;
      }
      if (lane == 0) skey[65] = 0xfffffffdu;  // "no segment left open by this drain" until one says so
      __syncwarp();
      // sequential sums per (tile, set) segment, in queue order, one lane per segment start; the
      // segment ends come from a ballot of the starts, so the loop is only load + add
      bool start[2];
      unsigned keyh[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 32 * h + lane;
        keyh[h] = skey[i];
        const unsigned prev = (i == 0) ? open_key : skey[i == 0 ? 0 : i - 1];
        start[h] = (i < cnt) && (i == 0 || keyh[h] != prev);
      }
      const unsigned long long starts = (unsigned long long)__ballot_sync(0xffffffffu, start[0]) |
                                        ((unsigned long long)__ballot_sync(0xffffffffu, start[1]) << 32);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = 32 * h + lane;
        if (start[h]) {
          const unsigned key = keyh[h];
          const unsigned long long rest = (i == 63) ? 0ull : (starts >> (i + 1));
          const int j_end = rest ? (i + __ffsll((long long)rest)) : cnt;
          double sum = (i == 0 && key == open_key) ? sdchi[64] : 0.0;
#pragma unroll 4
          for (int j = i; j < j_end; ++j) sum += sdchi[j];
          const size_t slot = (size_t)(key & 0x1ffffffu) * a.n_sets + (set_base + (int)(key >> 25));
          if (j_end < cnt) {
            a.partial[slot] = sum;  // closed: every term of this (tile, set) is in
          } else {
            sdchi[65] = sum;  // still open at the end of this drain
            skey[65] = key;
          }
          // the segment left open by the previous drain does not continue here: it closes as it was
          if (i == 0 && key != open_key && open_key != 0xffffffffu) {
            const size_t oslot =
                (size_t)(open_key & 0x1ffffffu) * a.n_sets + (set_base + (int)(open_key >> 25));
            a.partial[oslot] = sdchi[64];
          }
        }
      }
      __syncwarp();
      {
        const unsigned ok65 = skey[65];
        const bool still_open = (ok65 == skey[cnt - 1]);
        const double osum = sdchi[65];
        __syncwarp();
        open_key = still_open ? ok65 : 0xffffffffu;
        if (lane == 0) sdchi[64] = still_open ? osum : 0.0;
      }
      // shift the rest of the queue down
      const int rem = qn - cnt;
      unsigned keep[(SP_QCAP - 64) / 32];
#pragma unroll
      for (int m = 0; m < (SP_QCAP - 64) / 32; ++m) keep[m] = (32 * m + lane < rem) ? sq[cnt + 32 * m + lane] : 0u;
      __syncwarp();
#pragma unroll
      for (int m = 0; m < (SP_QCAP - 64) / 32; ++m)
        if (32 * m + lane < rem) sq[32 * m + lane] = keep[m];
      qn = rem;
      __syncwarp();
    }
    // close the last open segment of this group
    if (open_key != 0xffffffffu && lane == 0) {
      const size_t oslot = (size_t)(open_key & 0x1ffffffu) * a.n_sets + (set_base + (int)(open_key >> 25));
      a.partial[oslot] = sdchi[64];
    }
    __syncthreads();
  }
  n_kep = __reduce_add_sync(0xffffffffu, n_kep);
  n_ld = __reduce_add_sync(0xffffffffu, n_ld);
  if (lane == 0) {
    atomicAdd(a.counters + 0, (unsigned long long)n_kep);
    atomicAdd(a.counters + 1, (unsigned long long)n_ld);
  }
}


extern "C" size_t jxp_system_workspace_bytes(int32_t n_sets, int32_t n_bodies, int64_t n_times) {
  if (n_sets <= 0 || n_bodies <= 0 || n_times < 0) return 0;
  // sized for the wider (double) records so one query serves both precisions
  return sys_layout(n_sets, n_bodies, n_times, sizeof(BodyC<double>), sizeof(SetC<double>),
                    sizeof(double))
      .total;
}

template <typename T>
static int launch_system(const double* bodies, int32_t n_sets, int32_t n_bodies, const double* u,
                         int32_t n_u, int64_t u_stride, const double* t, int64_t n_times,
                         double exposure_time, int32_t exp_order, int32_t num_samples,
                         int32_t gl_order, int32_t flags, T* flux, T* flux_sum, const T* y,
                         const T* sigma, int64_t sigma_stride, T* loglike, void* workspace,
                         size_t workspace_bytes, void* stream) {
  static_assert(sizeof(BodyC<T>) % 8 == 0 && sizeof(SetC<T>) % 8 == 0, "record sizes");
  if (n_sets < 0 || n_bodies < 0 || n_times < 0)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_system_light_curve: negative size");
  if (n_u < 0 || n_u > JXP_MAX_LD_COEFFS)
    return fail(JXP_ERR_UNSUPPORTED, "jxp_system_light_curve: n_u = %d outside 0..%d", n_u,
                JXP_MAX_LD_COEFFS);
  if (gl_order < 1 || gl_order > JXP_MAX_GL_ORDER)
    return fail(JXP_ERR_UNSUPPORTED, "jxp_system_light_curve: order = %d outside 1..%d", gl_order,
                JXP_MAX_GL_ORDER);
  if (n_bodies > 64) return fail(JXP_ERR_UNSUPPORTED, "jxp_system_light_curve: n_bodies > 64");
  if (u_stride != 0 && u_stride != n_u)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_system_light_curve: u_set_stride must be 0 or n_u");
  if (!stride_ok(sigma_stride))
    return fail(JXP_ERR_BAD_SHAPE, "jxp_system_light_curve: sigma_stride must be 0 or 1");
  SysArgs<T> a;
  int rc = make_stencil(exposure_time, exp_order, num_samples, a.st);
  if (rc != JXP_OK) return rc;
  if (n_sets == 0 || n_bodies == 0 || n_times == 0) return JXP_OK;
  if (!bodies || !t || (n_u > 0 && !u))
    return fail(JXP_ERR_NULL_POINTER, "jxp_system_light_curve: null bodies/u/t");
  if (loglike && (!y || !sigma))
    return fail(JXP_ERR_NULL_POINTER, "jxp_system_light_curve: loglike needs y and sigma");
  if (!flux && !flux_sum && !loglike)
    return fail(JXP_ERR_NULL_POINTER, "jxp_system_light_curve: no output requested");
  const size_t need = jxp_system_workspace_bytes(n_sets, n_bodies, n_times);
  if (!workspace || workspace_bytes < need)
    return fail(JXP_ERR_WORKSPACE, "jxp_system_light_curve: workspace of %zu bytes needed, %zu given",
                need, workspace_bytes);
  if (((uintptr_t)workspace & 255) != 0)
    return fail(JXP_ERR_WORKSPACE, "jxp_system_light_curve: workspace must be 256-byte aligned");

  const SysLayout L = sys_layout(n_sets, n_bodies, n_times, sizeof(BodyC<double>),
                                 sizeof(SetC<double>), sizeof(double));
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  BodyC<T>* d_bodies = reinterpret_cast<BodyC<T>*>(ws + L.off_bodies);
  SetC<T>* d_sets = reinterpret_cast<SetC<T>*>(ws + L.off_sets);
  T* d_winv = reinterpret_cast<T*>(ws + L.off_winv);
  double* d_scalars = reinterpret_cast<double*>(ws + L.off_scalars);
  double* d_partial = reinterpret_cast<double*>(ws + L.off_partial);
  cudaStream_t st = (cudaStream_t)stream;

  const bool f32 = sizeof(T) == 4;
  // transit-walk path (jxp_walk.cu): eligibility is decided here, use on the device (sorted time stamps)
  const uint32_t dev = dev_options();
  const bool env_no_tl = (dev & JXP_DEV_NO_WALK) != 0, env_no_pair = (dev & JXP_DEV_NO_PAIR) != 0;
  const bool tl_any = sizeof(T) == 8 && n_u <= 2 && gl_order == 10 && !(flags & JXP_FLAG_NO_WINDOW) &&
                      n_times < 0x7ffffe00LL && !env_no_tl;
  const bool tl_ok = tl_any && n_bodies == 1 && a.st.n == 1;
  const bool use_tl = tl_ok && loglike && !flux && !flux_sum && !env_no_pair && n_times < (1LL << 25);
  // ... and of calls that want the flux itself (light_curve(orbit, u)(t) of a batch of single-planet systems)
  // (from 2^20 points: below that the extra launch costs more than the walk saves -- config 1, one set x
  // 1e5 samples, takes 39 us with the window-test kernel)
  const int64_t flux_min_points = (dev & JXP_DEV_WALK_ANY_SIZE) ? 0 : (1LL << 20);
  // ... of systems with several bodies and of exposure integration too (config 4): one plan + evaluation per body,
  // a lane walks the sub-samples of its sample in order
  const bool use_tl_flux = tl_any && n_bodies <= 16 && !loglike && (flux || flux_sum) &&
                           (int64_t)n_sets * n_times * n_bodies * a.st.n >= flux_min_points;
  const double widen = f32 ? 1.0 + 2e-4 : 1.0 + 1e-9;
  if (use_tl || use_tl_flux) {
    if constexpr (sizeof(T) == 8) {
      if (use_tl_flux) {
        // zero rows by the copy engine's fill, the in-transit samples scattered over them by k_walk_eval;
        // with unsorted time stamps (or a table that does not fit) k_system below rewrites every row
        cudaError_t em = cudaSuccess;
        if (flux_sum) em = cudaMemsetAsync(flux_sum, 0, sizeof(T) * (size_t)n_sets * (size_t)n_times, st);
        if (em == cudaSuccess && flux)
          em = cudaMemsetAsync(flux, 0, sizeof(T) * (size_t)n_sets * (size_t)n_times * (size_t)n_bodies, st);
        if (em != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_system_light_curve: memset: %s", cudaGetErrorString(em));
      }
      rc = launch_walk_f64(bodies, n_sets, n_bodies, u, n_u, u_stride, t, n_times, a.st, flux, flux_sum, y, sigma,
                           sigma_stride, loglike, ws, L, widen, 1e-7, st, g_ev_start, g_ev_stop,
                           (dev & JXP_DEV_PROFILE_ITEMS) ? 3 : ((dev & JXP_DEV_PROFILE_TABLES) ? 2 : ((dev & JXP_DEV_PROFILE_PLAN) ? 1 : 0)));
      if (rc != JXP_OK) return rc;
    }
  } else {
    const int n = n_sets * n_bodies;
    unsigned long long* counters = reinterpret_cast<unsigned long long*>(d_scalars + 2);
    int* tile_ctr = reinterpret_cast<int*>(ws + L.off_tilectr);
    if (loglike) {
      const int n_prep = (n + 255) / 256;
      k_sys_prep_ll<T><<<n_prep + LL_NCTA, 256, 0, st>>>(bodies, n_sets, n_bodies, u, n_u, u_stride, d_bodies,
                                                         d_sets, counters, tile_ctr, widen, 1e-7, n_prep, y,
                                                         sigma, sigma_stride, n_times, d_winv, d_scalars, nullptr);
    } else {
      k_sys_prep<T><<<(n + 127) / 128, 128, 0, st>>>(bodies, n_sets, n_bodies, u, n_u, u_stride, d_bodies,
                                                    d_sets, counters, tile_ctr, widen, 1e-7);
    }
    rc = check_launch("jxp_system_light_curve(prep)");
    if (rc != JXP_OK) return rc;
  }

  // launch geometry: persistent CTAs (as many as are resident at once) that pull set groups and
  // warp tiles from counters in the workspace; as many sets per group as keeps >= 8 groups per
  // SM worth of (group, CTA tile) work (amortises the record staging)
  const int sms = sm_count();
  int threads = 128;
  {  // development option: CTA size of the window-test kernels
    const int v = 32 * (int)JXP_DEV_GET_CTA_WARPS(dev);
    if (v >= 32 && v <= 128) threads = v;
  }
  const int64_t n_wtiles = (n_times + 32 * SYS_K - 1) / (32 * SYS_K);
  int64_t n_tiles = (n_wtiles + threads / 32 - 1) / (threads / 32);
  while (threads > 32 && n_tiles * ((n_sets + SYS_MAX_SETS - 1) / SYS_MAX_SETS) < 4 * sms) {
    threads /= 2;
    n_tiles = (n_wtiles + threads / 32 - 1) / (threads / 32);
  }
  int spc = SYS_MAX_SETS;
  while (spc > 32 && n_tiles * ((n_sets + spc - 1) / spc) < 8 * sms) spc -= 32;
  if (spc > n_sets) spc = n_sets;
  if (n_bodies * spc > 512) spc = (512 / n_bodies) > 0 ? (512 / n_bodies) : 1;
  // keep the staged records of a group under ~24 KB so that shared memory never limits residency
  // below the JXP_SYS_MINB CTAs per SM the register budget allows
  while (spc > 16 && sizeof(BodyC<T>) * (size_t)spc * n_bodies > 24 * 1024) spc -= 16;
  const int64_t n_groups = (n_sets + spc - 1) / spc;
  const int64_t n_units = n_tiles * n_groups;  // CTA-sized pieces of work
  if (n_wtiles > 2000000000LL) return fail(JXP_ERR_BAD_SHAPE, "jxp_system_light_curve: too many time samples");
  if (spc > 511) return fail(JXP_ERR_UNSUPPORTED, "jxp_system_light_curve: sets per CTA");

  a.bodies = d_bodies;
  a.sets = d_sets;
  a.t = t;
  a.n_times = n_times;
  a.n_sets = n_sets;
  a.n_bodies = n_bodies;
  a.sets_per_cta = spc;
  a.n_groups = (int)n_groups;
  a.tile_ctr = reinterpret_cast<int*>(ws + L.off_tilectr);
  a.n_wtiles = (int)n_wtiles;
  a.use_window = (flags & JXP_FLAG_NO_WINDOW) ? 0 : 1;
  a.flux = flux;
  a.flux_sum = flux_sum;
  a.y = y;
  a.winv = d_winv;
  a.partial = d_partial;
  a.counters = reinterpret_cast<unsigned long long*>(d_scalars + 2);
  a.tl_state = nullptr;
  a.tl_items = nullptr;
  a.tl_off = nullptr;
  a.tl_cnt = nullptr;
  a.tl_cap = 0;
  a.scalars = d_scalars;
  a.loglike = nullptr;
  a.tab.order = gl_order;
  if (gl_order != 10) gl_table_host(gl_order, a.tab);

  // multi variant of stage C: double, order 10, <= SYS_MB bodies, more than one evaluation per sample
  const bool multi = sizeof(T) == 8 && n_u <= 2 && gl_order == 10 && n_bodies <= SYS_MB &&
                     (int64_t)n_bodies * a.st.n > 1 && !(((flags & JXP_FLAG_NO_WINDOW) != 0) && n_bodies == 1 && a.st.n == 1);
  {
    const size_t nwp = threads / 32;
    size_t o = sizeof(BodyC<T>) * (size_t)spc * n_bodies;
    a.so_sset = (int)o;
    o += sizeof(SetC<T>) * (size_t)spc;
    a.so_sacc = (int)o;
    o += sizeof(double) * (size_t)spc * nwp;
    a.so_sq = (int)o;
    o += sizeof(unsigned short) * (size_t)SYS_QCAP * nwp;
    a.so_sqb = (int)o;
    o += sizeof(T) * (size_t)SYS_QCAP * nwp;
    a.so_macc = (int)o;
    o += sizeof(T) * 32 * SYS_MB * nwp;
    a.so_meq = (int)o;
    o += sizeof(unsigned short) * 64 * nwp;
    a.so_maccE = (int)o;  // regions from so_macc on exist in the multi variant only
  }
  const size_t smem = sizeof(BodyC<T>) * (size_t)spc * n_bodies + sizeof(SetC<T>) * (size_t)spc +
                      sizeof(double) * (size_t)spc * (threads / 32) +
                      sizeof(unsigned short) * (size_t)SYS_QCAP * (threads / 32) +
                      sizeof(T) * (size_t)SYS_QCAP * (threads / 32) +
                      (multi ? (sizeof(T) * 32 * SYS_MB + sizeof(unsigned short) * 64 + sizeof(T) * 32 * n_bodies) *
                                   (size_t)(threads / 32) : 0);
  if (smem > 200 * 1024) return fail(JXP_ERR_UNSUPPORTED, "jxp_system_light_curve: shared memory");
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    int per_sm = 0;
    cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (eo != cudaSuccess || per_sm < 1)
      return fail(JXP_ERR_CUDA, "jxp_system_light_curve: occupancy query: %s", cudaGetErrorString(eo));
    const int64_t resident = (int64_t)per_sm * sms;
    const unsigned n_ctas = (unsigned)(n_units < resident ? n_units : resident);
    if (g_ev_start) cudaEventRecord(g_ev_start, st);
    kern<<<n_ctas, threads, smem, st>>>(a);
    if (g_ev_stop) cudaEventRecord(g_ev_stop, st);
    return check_launch("jxp_system_light_curve");
  };
  const bool dense = (a.use_window == 0) && n_bodies == 1 && a.st.n == 1 && n_u <= 2;
  if constexpr (sizeof(T) == 8) {
    // log-likelihood of single-planet systems, no exposure integration: two samples per lane
    const bool pair = loglike && !flux && !flux_sum && n_bodies == 1 && a.st.n == 1 && n_u <= 2 &&
                      gl_order == 10 && a.use_window && n_times < (1LL << 25) && spc <= 64 && !env_no_pair;
    // the transit walk was launched in front (use_tl / use_tl_flux): the kernels below are then its
    // on-device fallback -- they return at once unless k_walk_eval left a flag in counters[3]
    // (unsorted / NaN time stamps, tables that do not fit)
    if (use_tl || use_tl_flux) {
      a.tl_state = reinterpret_cast<unsigned long long*>(d_scalars + 4);
      a.tl_cap = ~0ull;
    }
    if (pair) {
      const size_t smem2 = sizeof(BodyC<T>) * (size_t)spc + sizeof(SetC<T>) * (size_t)spc +
                           (size_t)(threads / 32) * (66 * sizeof(double) + (SP_QCAP + 64 + 2) * sizeof(unsigned));
#ifndef JXP_PAIR_PU
#define JXP_PAIR_PU 1
#endif

      auto kern = k_system_pair<JXP_PAIR_PU>;
      int per_sm = 0;
      cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem2);
      if (eo != cudaSuccess || per_sm < 1)
        return fail(JXP_ERR_CUDA, "jxp_system_light_curve: occupancy query: %s", cudaGetErrorString(eo));
      const int64_t resident = (int64_t)per_sm * sms;
      const unsigned n_ctas = (unsigned)(n_units < resident ? n_units : resident);
      // with the transit walk in front this launch is the fallback: the events stay on k_walk_eval
      if (g_ev_start && !a.tl_state) cudaEventRecord(g_ev_start, st);
      kern<<<n_ctas, threads, smem2, st>>>(a);
      if (g_ev_stop && !a.tl_state) cudaEventRecord(g_ev_stop, st);
      rc = check_launch("jxp_system_light_curve(pair)");
      if (rc != JXP_OK) return rc;
      k_loglike_final<T><<<(n_sets + 31) / 32, 256, 0, st>>>(d_partial, (int)n_wtiles, n_sets, d_scalars, loglike,
                                                              a.tl_state, a.tl_cap);
      return check_launch("jxp_system_light_curve(loglike final)");
    }
  }
  if (n_u > 2) {
    // general polynomial limb darkening: queue kernel with the l >= 3 terms compiled in
    if (gl_order == 10)
      rc = loglike ? launch(k_system<T, 11, true, false, true>) : launch(k_system<T, 11, false, false, true>);
    else
      rc = loglike ? launch(k_system<T, 0, true, false, true>) : launch(k_system<T, 0, false, false, true>);
  } else if (gl_order == 10) {
    if (dense)
      rc = loglike ? launch(k_system<T, 11, true, true>) : launch(k_system<T, 11, false, true>);
    else if (multi)
      // several flux evaluations per queued sample: the unrolled node loop pays (config 4)
      rc = loglike ? launch(k_system<T, 11, true, false, false, 5>)
                   : launch(k_system<T, 11, false, false, false, 5>);
    else
      rc = loglike ? launch(k_system<T, 11, true, false>) : launch(k_system<T, 11, false, false>);
  } else {
    rc = loglike ? launch(k_system<T, 0, true, false>) : launch(k_system<T, 0, false, false>);
  }
  if (rc != JXP_OK) return rc;

  if (loglike) {
    k_loglike_final<T><<<(n_sets + 31) / 32, 256, 0, st>>>(d_partial, (int)n_wtiles, n_sets,
                                                            d_scalars, loglike);
    rc = check_launch("jxp_system_light_curve(loglike final)");
  }
  return rc;
}

extern "C" int jxp_system_stats(const void* workspace, int32_t n_sets, int32_t n_bodies,
                                int64_t n_times, int64_t* stats_host, void* stream) {
  if (!workspace || !stats_host) return fail(JXP_ERR_NULL_POINTER, "jxp_system_stats: null pointer");
  if (n_sets <= 0 || n_bodies <= 0 || n_times <= 0)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_system_stats: empty problem");
  const SysLayout L = sys_layout(n_sets, n_bodies, n_times, sizeof(BodyC<double>),
                                 sizeof(SetC<double>), sizeof(double));
  const unsigned char* ws = static_cast<const unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(stats_host, ws + L.off_scalars + 2 * sizeof(double),
                                  2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_system_stats: %s", cudaGetErrorString(e));
  return JXP_OK;
}

extern "C" int jxp_system_list_stats(const void* workspace, int32_t n_sets, int32_t n_bodies,
                                     int64_t n_times, int64_t* stats_host, void* stream) {
  if (!workspace || !stats_host) return fail(JXP_ERR_NULL_POINTER, "jxp_system_list_stats: null pointer");
  if (n_sets <= 0 || n_bodies <= 0 || n_times <= 0)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_system_list_stats: empty problem");
  const SysLayout L = sys_layout(n_sets, n_bodies, n_times, sizeof(BodyC<double>),
                                 sizeof(SetC<double>), sizeof(double));
  const unsigned char* ws = static_cast<const unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  // counters[2..5]: samples on the inside lists, flags, samples on the edge lists, samples on the constant-kappa code
  int64_t raw[4] = {0, 0, 0, 0};
  cudaError_t e = cudaMemcpyAsync(raw, ws + L.off_scalars + 4 * sizeof(double), 4 * sizeof(int64_t),
                                  cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) {
    stats_host[0] = raw[0] + raw[2];
    stats_host[1] = raw[1];
    stats_host[2] = raw[0];
    stats_host[3] = raw[3];
  }
  if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_system_list_stats: %s", cudaGetErrorString(e));
  return JXP_OK;
}

extern "C" int jxp_system_walk_stats(const void* workspace, int32_t n_sets, int32_t n_bodies,
                                     int64_t n_times, int64_t* stats_host, void* stream) {
  if (!workspace || !stats_host) return fail(JXP_ERR_NULL_POINTER, "jxp_system_walk_stats: null pointer");
  if (n_sets <= 0 || n_bodies <= 0 || n_times <= 0)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_system_walk_stats: empty problem");
  const SysLayout L = sys_layout(n_sets, n_bodies, n_times, sizeof(BodyC<double>),
                                 sizeof(SetC<double>), sizeof(double));
  const unsigned char* ws = static_cast<const unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t raw[4] = {0, 0, 0, 0};
  cudaError_t e = cudaMemcpyAsync(raw, ws + L.off_scalars + WALK_TSTAT0 * sizeof(double), 4 * sizeof(int64_t),
                                  cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_system_walk_stats: %s", cudaGetErrorString(e));
  for (int k = 0; k < 4; ++k) stats_host[k] = raw[k];
  return JXP_OK;
}

extern "C" int jxp_system_light_curve_f64(const double* bodies, int32_t n_sets, int32_t n_bodies,
                                          const double* u, int32_t n_u, int64_t u_set_stride,
                                          const double* t, int64_t n_times, double exposure_time,
                                          int32_t exp_order, int32_t num_samples, int32_t gl_order,
                                          int32_t flags, double* flux, double* flux_sum,
                                          const double* y, const double* sigma,
                                          int64_t sigma_stride, double* loglike, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  return launch_system<double>(bodies, n_sets, n_bodies, u, n_u, u_set_stride, t, n_times,
                               exposure_time, exp_order, num_samples, gl_order, flags, flux,
                               flux_sum, y, sigma, sigma_stride, loglike, workspace,
                               workspace_bytes, stream);
}
extern "C" int jxp_system_light_curve_f32(const double* bodies, int32_t n_sets, int32_t n_bodies,
                                          const double* u, int32_t n_u, int64_t u_set_stride,
                                          const double* t, int64_t n_times, double exposure_time,
                                          int32_t exp_order, int32_t num_samples, int32_t gl_order,
                                          int32_t flags, float* flux, float* flux_sum,
                                          const float* y, const float* sigma, int64_t sigma_stride,
                                          float* loglike, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  return launch_system<float>(bodies, n_sets, n_bodies, u, n_u, u_set_stride, t, n_times,
                              exposure_time, exp_order, num_samples, gl_order, flags, flux,
                              flux_sum, y, sigma, sigma_stride, loglike, workspace, workspace_bytes,
                              stream);
}

// ======================================================================================
// FP64 / FP32 pipe peak (measurement helper)
// ======================================================================================
template <typename T>
__global__ void __launch_bounds__(1024) k_peak_fma(T* sink, int64_t iters) {
  T x = T(1.0) + T(1e-9) * T(threadIdx.x);
  const T y = T(1e-7);
  T a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  T b0 = x + 8, b1 = x + 9, b2 = x + 10, b3 = x + 11, b4 = x + 12, b5 = x + 13, b6 = x + 14,
    b7 = x + 15;
  const T m = T(0.999999);
  for (int64_t i = 0; i < iters; ++i) {
    a0 = jfma(a0, m, y); a1 = jfma(a1, m, y); a2 = jfma(a2, m, y); a3 = jfma(a3, m, y);
    a4 = jfma(a4, m, y); a5 = jfma(a5, m, y); a6 = jfma(a6, m, y); a7 = jfma(a7, m, y);
    b0 = jfma(b0, m, y); b1 = jfma(b1, m, y); b2 = jfma(b2, m, y); b3 = jfma(b3, m, y);
    b4 = jfma(b4, m, y); b5 = jfma(b5, m, y); b6 = jfma(b6, m, y); b7 = jfma(b7, m, y);
  }
  sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] =
      ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)) + ((b0 + b1) + (b2 + b3)) +
      ((b4 + b5) + (b6 + b7));
}

template <typename T>
static int launch_peak(T* sink, int32_t n_blocks, int32_t n_threads, int64_t iters, void* stream) {
  if (!sink) return fail(JXP_ERR_NULL_POINTER, "jxp_peak_fma: null sink");
  if (n_blocks <= 0 || n_threads <= 0 || n_threads > 1024 || iters < 0)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_peak_fma: bad launch shape");
  k_peak_fma<T><<<n_blocks, n_threads, 0, (cudaStream_t)stream>>>(sink, iters);
  return check_launch("jxp_peak_fma");
}
extern "C" int jxp_peak_fma_f64(double* sink, int32_t n_blocks, int32_t n_threads, int64_t iters,
                                void* stream) {
  return launch_peak<double>(sink, n_blocks, n_threads, iters, stream);
}
extern "C" int jxp_peak_fma_f32(float* sink, int32_t n_blocks, int32_t n_threads, int64_t iters,
                                void* stream) {
  return launch_peak<float>(sink, n_blocks, n_threads, iters, stream);
}
