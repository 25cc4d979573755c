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

#include <cstdlib>

#include "jxp_host.h"
#include "jxp_orbit.cuh"
#include "jxp_list.cuh"

using namespace jxp;
using namespace jxph;

namespace {
constexpr int NO = 5;  // orbit directions: P, t0, e, omega, b
constexpr int TR_K = 2;
constexpr int TR_MAX_SETS = 32;

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
  size_t off_tl_off, off_tl_cnt, off_tl_items, off_tl_vals;  // transit list of the gradient step
  unsigned long long tl_cap;
  int n_tiles;
};

TrLayout tr_layout(int n_sets, int64_t n_times) {
  TrLayout L;
  const int tile = 32 * TR_K;  // one warp tile
  L.n_tiles = (int)((n_times + tile - 1) / tile);
  size_t off = 0;
  L.off_bodies = off;
  off = align_up(off + sizeof(BodyD<double>) * (size_t)n_sets, 256);
  L.off_winv = off;
  off = align_up(off + sizeof(double) * (size_t)n_times, 256);
  L.off_scalars = off;
  off = align_up(off + 256 * sizeof(double), 256);
  L.off_partial = off;
  off = align_up(off + sizeof(double) * (size_t)n_sets * (L.n_tiles + 4) * (1 + JXP_NTHETA), 256);
  // transit list (log-likelihood + gradient without stored rows): an 8-byte item and 9 doubles (chi^2
  // term and the 8 gradient terms) per in-window sample, room for 1/16 of the grid
  L.off_tl_off = off;
  off = align_up(off + 2 * sizeof(unsigned long long) * (size_t)n_sets, 256);
  L.off_tl_cnt = off;
  off = align_up(off + 2 * sizeof(unsigned) * (size_t)n_sets, 256);
  L.tl_cap = (unsigned long long)n_sets * (unsigned long long)n_times / 16ull + 65536ull;
  if (L.tl_cap > (1ull << 26)) L.tl_cap = 1ull << 26;  // at most 5.4 GB of list; beyond it the call falls back
  L.off_tl_items = off;
  off = align_up(off + sizeof(unsigned long long) * (size_t)L.tl_cap, 256);
  L.off_tl_vals = off;
  off = align_up(off + sizeof(double) * (1 + JXP_NTHETA) * (size_t)L.tl_cap, 256);
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
// (general polynomial limb darkening, jxp_transit_partials_poly_f64: theta rows are 6 + n_u long and the normalised
// Green's coefficients with their derivatives w.r.t. u go to `poly`)
struct PolyD {
  double g[JXP_MAX_LD_COEFFS + 1];                       // g_n / norm, limb_dark.py:42-45, 87-99
  double dg[JXP_MAX_LD_COEFFS + 1][JXP_MAX_LD_COEFFS];   // d g_n / d u_j
};

template <typename T>
__global__ void k_tr_prep(const double* __restrict__ theta, int n_sets,
                          const double* __restrict__ star, int64_t star_stride,
                          BodyD<T>* __restrict__ out, double widen, double margin,
                          unsigned long long* __restrict__ tl_state, int theta_stride = JXP_NTHETA,
                          int n_u = 2, PolyD* __restrict__ poly = nullptr) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sets) return;
  if (s == 0 && tl_state) tl_state[0] = tl_state[1] = tl_state[2] = tl_state[3] = 0ull;  // transit list of this call
  typedef Dual<double, NO> D;
  const double* th = theta + (size_t)s * theta_stride;
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
  c.e = e.v;
  c.ome = 1.0 - e.v;
  c.ope = 1.0 + e.v;
  c.s1me2 = sqrt((1.0 - e.v) * (1.0 + e.v));
  c.inv_ope = 1.0 / (1.0 + e.v);
  c.cw = cw.v;
  c.sw = sw.v;
  c.ci = ci.v;
  c.si = si.v;
  const D na = -(a * ome2);
  c.na_ome2 = na.v;
  c.na = -a.v;
  c.inv_rstar = 1.0 / rstar;
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
  if (poly != nullptr) {
    typedef Dual<double, JXP_MAX_LD_COEFFS> U;
    U uu[JXP_MAX_LD_COEFFS], gg[JXP_MAX_LD_COEFFS + 1];
    for (int k = 0; k < JXP_MAX_LD_COEFFS; ++k) uu[k] = k < n_u ? U::seed(th[6 + k], k) : U(0.0);
    for (int k = 0; k <= JXP_MAX_LD_COEFFS; ++k) gg[k] = U(0.0);
    ld_greens_n<U>(n_u, uu, gg);
    PolyD pd;
    for (int k = 0; k <= JXP_MAX_LD_COEFFS; ++k) {
      pd.g[k] = gg[k].v;
      for (int j = 0; j < JXP_MAX_LD_COEFFS; ++j) pd.dg[k][j] = gg[k].d[j];
    }
    poly[s] = pd;
    o.sc.g0 = T(gg[0].v);
    o.sc.g1 = T(gg[1].v);
    o.sc.g2 = T(gg[2].v);
    for (int k = 0; k < 6; ++k) o.sc.ghi[k] = T(gg[3 + k].v);
    o.sc.lmax = n_u;
    for (int j = 0; j < 2; ++j) o.dg[0][j] = o.dg[1][j] = o.dg[2][j] = T(0);
  } else {
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
// on_inside_list (transit-list path): a sample on the inside list that has b < 1 - r itself takes
// ld_solution_inside (constant kappas and node cosines: no atan2, no node sin/cos, no derivative through
// them) -- decided per sample, never by its neighbours.
template <typename T, int ORDER>
__device__ __forceinline__ bool eval_sample_partials(const BodyD<T>& bd,
                                                     const GLTable* __restrict__ tab, double tm,
                                                     T& flux, T* __restrict__ part,
                                                     bool on_inside_list = false) {
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
  // the primal of the position stage is always double (BodyC, jxp_orbit.cuh); the tangents run in T
  const double Mr = mod_two_pi(M);
  double sf_p, cf_p, Edummy;
  kepler_reduced<double, false>(Mr, c.e, c.ome, c.ope, c.s1me2, c.inv_ope, sf_p, cf_p, Edummy);
  const T sf = T(sf_p), cf = T(cf_p);
  const T e_t = T(c.e), inv_rstar_t = T(c.inv_rstar);
  // kepler.py:64-79
  const T ecosf = e_t * cf;
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
  const D ed = D::seed(e_t, 2);
  D b, Z;
  sky_position_g<D, T>(ed, bd.cw, bd.sw, bd.ci, bd.si, bd.na_ome2, inv_rstar_t, sfd, cfd, b, Z);
  if (sizeof(T) < sizeof(double)) {
    // fp32 mode: (x0, y0) = r0 (cos f, sin f) are of order a / R* ~ 15 stellar radii and cancel to O(1) in
    // the rotation; in float that alone costs 1e-6 of a stellar radius in b.  The primal comes from the
    // double position stage, the float chain above supplies the tangents.
    double bp, Zp;
    sky_position(c, sf_p, cf_p, bp, Zp);
    b.v = T(bp);
    Z.v = T(Zp);
  }
  if (!(Z.v > T(0)) || (b.v >= c.onepr)) return false;
  typedef Dual<T, 2> L;
  L s0, s1, s2;
  const T ar = jabs(c.rr);
  if (ORDER == 11 && on_inside_list && (b.v < T(1) - ar) && (ar < T(1)))
    ld_solution_inside<L>(L::seed(b.v, 0), L::seed(c.rr, 1), s0, s1, s2);
  else
    ld_solution<L, ORDER>(L::seed(b.v, 0), L::seed(c.rr, 1), 2, tab, s0, s1, s2);
  const L f = ld_combine(2, s0, s1, s2, L(bd.sc.g0), L(bd.sc.g1), L(bd.sc.g2));
  flux = f.v;
#pragma unroll
  for (int i = 0; i < NO; ++i) part[i] = f.d[0] * b.d[i];
  part[5] = f.d[1] * inv_rstar_t;
  part[6] = s0.v * bd.dg[0][0] + s1.v * bd.dg[1][0] + s2.v * bd.dg[2][0];
  part[7] = s0.v * bd.dg[0][1] + s1.v * bd.dg[1][1] + s2.v * bd.dg[2][1];
  return true;
}

template <typename T>
struct TrArgs {
  const BodyD<T>* bodies;
  const double* t;
  int64_t n_times;
  int n_sets, sets_per_cta, n_tiles, n_wtiles, use_window;
  int run;  // consecutive warp tiles one warp processes (queued samples are carried from one to the next)
  int tpw;        // (bulk_fill) warp tiles per warp: the CTA owns nwarps * tpw consecutive warp tiles
  int bulk_fill;  // zero rows written by the copy engine (cp.async.bulk from a zeroed shared-memory tile), run == 1 only:
                  // 1 = every thread issues its share and the drains wait for the copies, 2 = warp 0 issues, results buffered
  T* flux;      // [n_sets][n_times]
  T* partials;  // [n_sets][8][n_times]
  const T* y;
  const T* winv;
  double* partial;  // [n_sets][n_tiles][9]
  // transit-list path of the gradient step (nullptr: not used by this call)
  unsigned long long* tl_state;
  unsigned long long* tl_items;
  unsigned long long* tl_off;
  unsigned* tl_cnt;
  unsigned long long tl_cap;
  double* tl_vals;  // [9][tl_cap]
  const double* scalars;
  T* loglike;
  T* dloglike;
  Stencil st;
  GLTable tab;
};

// Same warp-centric three-stage structure as k_system (jxp_kernels.cu): (A) lanes = sets, the
// warp's time span against each set's transit window; (B) lanes = samples for candidate sets,
// hits compacted into a per-warp queue, zero rows written for everything else; (C) the queue
// drained 32 in-window samples at a time through ONE instance of the dual-number code.
// Every zero row is one fully coalesced warp store (256 B in f64).
constexpr int TR_QCAP = 128;
constexpr int TR_RB = 160;  // buffered results per warp (bulk_fill): five full drains
constexpr int TR_TPW_MAX = 2;
// 3 CTAs/SM = 168 registers: with 128 the dual-number code spills (100 B stores / 124 B loads per
// thread) and the local-memory traffic competes with the 59 GB row-store stream (config 5: 11.7 ms;
// 96 registers 13.3 ms; 168 registers 11.15 ms; 200 registers, no spill at all, 12.4 ms)
#ifndef JXP_TR_MINB
#define JXP_TR_MINB 3
#endif
// the float instantiations need ~100 registers: 5 CTAs per SM keep more row stores in flight (config 5 in float:
// 6.60 ms at 3 or 4 CTAs per SM, 6.26 ms at 5)
#ifndef JXP_TR_MINB_F32
#define JXP_TR_MINB_F32 5
#endif
template <typename T, int ORDER, bool LOGLIKE>
__global__ void __launch_bounds__(128, sizeof(T) == 4 ? JXP_TR_MINB_F32 : JXP_TR_MINB)
k_transit_partials(const __grid_constant__ TrArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int K = TR_K;
  constexpr int NQ = 1 + JXP_NTHETA;
  if (a.tl_state != nullptr && tl_active(a.tl_state, a.tl_cap)) return;  // the transit-list kernels did the work
  const int nthreads = blockDim.x;
  const int nwarps = nthreads >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x % a.n_tiles;
  const int group = blockIdx.x / a.n_tiles;
  const int set_base = group * a.sets_per_cta;
  const int nsl = min(a.sets_per_cta, a.n_sets - set_base);
  BodyD<T>* sbody = reinterpret_cast<BodyD<T>*>(smem_raw);
  double* sacc_all = reinterpret_cast<double*>(sbody + a.sets_per_cta);  // [nwarps][spc][9]
  unsigned* sq_all = reinterpret_cast<unsigned*>(sacc_all + (LOGLIKE ? (size_t)nwarps * a.sets_per_cta * NQ : 0));
  double* sacc = sacc_all + (size_t)warp * a.sets_per_cta * NQ;
  unsigned* sq = sq_all + warp * TR_QCAP;
  // (bulk_fill) results of the drains that finish before the zero rows have landed: [NQ][TR_RB] values + TR_RB items per warp
  T* sres = reinterpret_cast<T*>(sq_all + (size_t)nwarps * TR_QCAP) + (size_t)warp * (NQ * TR_RB);
  unsigned* sitem = reinterpret_cast<unsigned*>(reinterpret_cast<T*>(sq_all + (size_t)nwarps * TR_QCAP) +
                                                (size_t)nwarps * (NQ * TR_RB)) + warp * TR_RB;
  int nbuf = 0;
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.bodies + set_base);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(sbody);
    const int nw = nsl * (int)(sizeof(BodyD<T>) / 8);
    for (int i = threadIdx.x; i < nw; i += nthreads) dst[i] = src[i];
    if (LOGLIKE)
      for (int i = lane; i < a.sets_per_cta * NQ; i += 32) sacc[i] = 0.0;
  }
  if (a.use_window == 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < nsl; i += nthreads) sbody[i].c.flags |= BODY_ALWAYS;
  }
  // this warp's run of consecutive warp tiles
  // Step r of the run handles warp tile (r * n_tiles + tile) * nwarps + warp: at every step the warps
  // of a CTA, and the CTAs running side by side, write ADJACENT pieces of each output row (the row
  // stores then stream through DRAM pages like a plain fill; with contiguous runs per warp the same
  // fill was 15 % slower).  Queued samples carry their absolute index, so the tiles of a run need
  // not be neighbours.
  const int64_t wrun = (int64_t)tile * nwarps + warp;
  const int nsub = a.st.n;
  const double dtmin = a.st.dt_min, dtmax = a.st.dt_max;
  const bool want_rows = (a.flux != nullptr) || (a.partials != nullptr);
  int qn = 0;
  // ---- zero rows by the copy engine ---------------------------------------------------------------
  // With one warp tile per warp (run == 1) the CTA owns `nwarps * 32 K` consecutive samples of every row of its
  // sets: each row piece (2 KB in f64, 1 KB in f32) is ONE bulk copy from a zeroed shared-memory tile, issued by
  // one thread -- 9 x 32 copies per CTA instead of 9 x 32 x 4 warp-wide stores.  The copy queue of the SM is the
  // bottleneck of this kernel (the issuing thread blocks while it is full: 37 % of the stall samples of a capture in
  // which every thread issued its share), so WARP 0 ALONE issues the copies while the other warps go straight on to
  // the window tests and the dual-number code; the warp tiles of the CTA are claimed dynamically (warp 0 joins when
  // its copies are queued).  The results of the drains are kept in shared memory (TR_RB per warp) and stored over
  // the zero rows once the copies are complete (fill_wait, then flush) -- normally at the end of the warp's tiles,
  // earlier only if the buffer fills up (sets that always transit).  With the in-kernel reduction (LOGLIKE) the
  // tiles stay with their warps: the order of the per-warp sums must not depend on timing.
  __shared__ __align__(128) unsigned char s_zero[4 * TR_TPW_MAX * 32 * TR_K * sizeof(T)];
  __shared__ int s_next;
  bool fill_pending = false;
  if (threadIdx.x == 0) s_next = 0;
  if (a.bulk_fill && want_rows) {
    for (int i = threadIdx.x; i < (int)(sizeof(s_zero) / 16); i += nthreads)
      reinterpret_cast<uint4*>(s_zero)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fill_pending = true;
  }
  __syncthreads();
  const bool decoupled = a.bulk_fill == 2;
  const bool dyn = fill_pending && decoupled && !LOGLIKE;
  if (fill_pending && (warp == 0 || !decoupled)) {
    const int64_t cbase = (int64_t)tile * (nwarps * a.tpw) * (32 * K);
    const int64_t clen = min((int64_t)(nwarps * a.tpw) * (32 * K), a.n_times - cbase);
    if (clen > 0) {
      const unsigned bytes = (unsigned)(clen * (int64_t)sizeof(T));
      const unsigned src = (unsigned)__cvta_generic_to_shared(s_zero);
      const int rps = (a.flux ? 1 : 0) + (a.partials ? JXP_NTHETA : 0);
      for (int rho = decoupled ? lane : (int)threadIdx.x; rho < nsl * rps; rho += decoupled ? 32 : nthreads) {
        const int sl = rho / rps;
        int q = rho - sl * rps;
        const size_t set = (size_t)(set_base + sl);
        T* dst;
        if (a.flux && q == 0) {
          dst = a.flux + set * a.n_times + cbase;
        } else {
          q -= a.flux ? 1 : 0;
          dst = a.partials + (set * JXP_NTHETA + q) * a.n_times + cbase;
        }
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                     : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  // every thread's copies complete (writes performed, not just the source read) before any lane of the CTA stores
  // an in-window sample over them; each warp passes here exactly once (before its first drain, or on its way out)
  auto fill_wait = [&]() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
    fill_pending = false;
  };
  auto flush = [&]() {
    __syncwarp();
    for (int e = lane; e < nbuf; e += 32) {
      const unsigned item = sitem[e];
      const size_t set = (size_t)(set_base + (int)(item >> 26));
      const size_t idx = (size_t)(item & 0x3ffffffu);
      if (a.flux) a.flux[set * a.n_times + idx] = sres[e];
      if (a.partials) {
#pragma unroll
        for (int q = 0; q < JXP_NTHETA; ++q) a.partials[(set * JXP_NTHETA + q) * a.n_times + idx] = sres[(1 + q) * TR_RB + e];
      }
    }
    nbuf = 0;
    __syncwarp();
  };

  // fill state
  bool have_tile = false, more = true;
  int r_next = 0;
  int64_t wbase = 0;
  double tk[K];
  unsigned valid = 0, visit = 0;
  double span_lo = 0.0, span = 0.0;
  int c0 = 0, c0_cur = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) tk[k] = 0.0;
  while (true) {
    // ---- fill: stages A and B until 32 samples are queued or the run has no tile left ------------
    while (more && qn < 32) {
      if (!have_tile) {
        if (a.bulk_fill) {  // the CTA's nwarps * tpw consecutive tiles: whoever is free (dyn), or warp + nwarps * r
          int wt = warp + nwarps * r_next;
          if (dyn) {
            if (lane == 0) wt = atomicAdd(&s_next, 1);
            wt = __shfl_sync(0xffffffffu, wt, 0);
          } else {
            ++r_next;
          }
          wbase = ((int64_t)tile * (nwarps * a.tpw) + wt) * (32 * K);
          if (wt >= nwarps * a.tpw || wbase >= a.n_times) {
            more = false;
            break;
          }
        } else {
          wbase = (((int64_t)r_next * a.n_tiles + tile) * nwarps + warp) * (32 * K);
          if (r_next >= a.run || wbase >= a.n_times) {
            more = false;
            break;
          }
          ++r_next;
        }
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
        span_lo = tmin + dtmin;
        span = (tmax + dtmax) - span_lo;
        // Zero rows first: the whole (sets of the group) x (flux + 8 partials) x (this tile's 64
        // samples) block is written as one 128-bit store per lane and row (512 contiguous bytes per
        // warp store in f64); stage C then overwrites the in-window samples (3 % of them on config 5).
        if (want_rows && !a.bulk_fill) {
          static_assert(TR_K == 2, "the zero fill assumes 2 samples per lane");
          struct alignas(2 * sizeof(T)) V2 {
            T x, y;
          };
          const int64_t s0 = wbase + 2 * lane;
          const bool two = s0 + 1 < a.n_times, one = s0 < a.n_times;
          const bool vec = ((a.n_times & 1) == 0) &&
                           ((reinterpret_cast<uintptr_t>(a.flux) & (2 * sizeof(T) - 1)) == 0) &&
                           ((reinterpret_cast<uintptr_t>(a.partials) & (2 * sizeof(T) - 1)) == 0);
          V2 z;
          z.x = T(0);
          z.y = T(0);
#pragma unroll 1
          for (int sl = 0; sl < nsl; ++sl) {
            const size_t set = (size_t)(set_base + sl);
            if (a.flux) {
              T* row = a.flux + set * a.n_times + s0;
              if (vec && two) {
                *reinterpret_cast<V2*>(row) = z;
              } else {
                if (one) row[0] = T(0);
                if (two) row[1] = T(0);
              }
            }
            if (a.partials) {
              T* row = a.partials + set * JXP_NTHETA * a.n_times + s0;
#pragma unroll
              for (int q = 0; q < JXP_NTHETA; ++q) {
                if (vec && two) {
                  *reinterpret_cast<V2*>(row) = z;
                } else {
                  if (one) row[0] = T(0);
                  if (two) row[1] = T(0);
                }
                row += a.n_times;
              }
            }
          }
          __syncwarp();  // orders these stores before stage C's stores of other lanes to the same rows
        }
        c0 = 0;
        visit = 0;
        have_tile = true;
      }
      if (!visit) {
        if (c0 >= nsl) {
          have_tile = false;
          continue;
        }
        // ---- stage A: lane = set ----------------------------------------------------------
        bool cand = false;
        const int sl = c0 + lane;
        if (sl < nsl) {
          const BodyC<T>& c = sbody[sl].c;
          if (c.flags & BODY_ALWAYS) {
            cand = true;
          } else {
            const double ph0 = fma(span_lo - c.t0ref, c.inv_period, -c.phase_c);
            const double d0 = ph0 - rint(ph0);
            const double d1 = fma(span, c.inv_period, d0);
            cand = !((d1 < c.wlo) | ((d0 > c.whi) & (d1 < 1.0 + c.wlo)));
          }
        }
        visit = __ballot_sync(0xffffffffu, cand);  // only candidate sets are looked at sample by sample
        c0_cur = c0;
        c0 += 32;
        continue;
      }
      // ---- stage B: lane = sample, one candidate set ------------------------------------------
      const int s_in = __ffs(visit) - 1;
      visit &= visit - 1;
      const int sl = c0_cur + s_in;
      unsigned hits = 0;
      const BodyC<T>& c = sbody[sl].c;
      if (c.flags & BODY_ALWAYS) {
        hits = valid;
      } else {
        const double wlo = fma(-dtmax, c.inv_period, c.wlo), whi = fma(-dtmin, c.inv_period, c.whi);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const double ph = fma(tk[k] - c.t0ref, c.inv_period, -c.phase_c);
          const double d = ph - rint(ph);
          hits |= ((d >= wlo) & (d <= whi)) ? (1u << k) : 0u;
        }
        hits &= valid;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const unsigned mk = __ballot_sync(0xffffffffu, (hits >> k) & 1u);
        if ((hits >> k) & 1u)
          sq[qn + __popc(mk & ((1u << lane) - 1u))] = ((unsigned)sl << 26) | (unsigned)(wbase + 32 * k + lane);
        qn += __popc(mk);
      }
      __syncwarp();
    }
    if (qn == 0) break;  // no tile left in the run and nothing queued
    if (fill_pending && !decoupled) fill_wait();

    // ---- stage C: drain -----------------------------------------------------------------------
    // 32 queued samples (fewer only at the very end of the run): queued samples carry their absolute
    // index and are carried from one tile of the run to the next, so the drains stay full
    __syncwarp();
    {
      const int cnt = qn < 32 ? qn : 32;
      const bool have = lane < cnt;
      const unsigned item = have ? sq[lane] : sq[0];
      const int sl = (int)(item >> 26);
      const int64_t idx = (int64_t)(item & 0x3ffffffu);
      const double tm0 = a.t[idx];
      T out[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) out[q] = T(0);
      bool any = false;
      if (have) {
        const BodyD<T>& bd = sbody[sl];
        const int set = set_base + sl;
#pragma unroll 1
        for (int j = 0; j < nsub; ++j) {
          const double tm = tm0 + a.st.dt[j];
          if (!(bd.c.flags & BODY_ALWAYS) && outside_window(bd.c, tm)) continue;
          T f1 = T(0);
          T p1[JXP_NTHETA];
          if (eval_sample_partials<T, ORDER>(bd, &a.tab, tm, f1, p1)) {
            const T w = (nsub == 1) ? T(1) : T(a.st.w[j]);
            out[0] = jfma(w, f1, out[0]);
#pragma unroll
            for (int q = 0; q < JXP_NTHETA; ++q) out[1 + q] = jfma(w, p1[q], out[1 + q]);
            any = true;
          }
        }
        if (!fill_pending) {
          if (a.flux) a.flux[(size_t)set * a.n_times + idx] = out[0];
          if (a.partials) {
#pragma unroll
            for (int q = 0; q < JXP_NTHETA; ++q)
              a.partials[((size_t)set * JXP_NTHETA + q) * a.n_times + idx] = out[1 + q];
          }
        }
      }
      if (fill_pending) {  // (warp-uniform)
        if (nbuf + 32 > TR_RB) {  // buffer full: wait for the zero rows, empty it, store directly from here on
          fill_wait();
          flush();
          if (have) {
            const int set = set_base + sl;
            if (a.flux) a.flux[(size_t)set * a.n_times + idx] = out[0];
            if (a.partials) {
#pragma unroll
              for (int q = 0; q < JXP_NTHETA; ++q)
                a.partials[((size_t)set * JXP_NTHETA + q) * a.n_times + idx] = out[1 + q];
            }
          }
        } else {
          if (have) {
            sitem[nbuf + lane] = item;
#pragma unroll
            for (int q = 0; q < NQ; ++q) sres[q * TR_RB + nbuf + lane] = out[q];
          }
          nbuf += cnt;
        }
      }
      if (LOGLIKE) {
        double acc9[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc9[q] = 0.0;
        if (have && any) {
          const T w = a.winv[idx];
          const T r0 = (a.y[idx] - T(1)) * w;
          const T mw = out[0] * w;
          const T rm = r0 - mw;  // ((y - 1) - m) / sigma
          acc9[0] = (double)(-mw) * (double)(T(2) * r0 - mw);
          const double g = (double)(rm * w);  // d loglike / d m
#pragma unroll
          for (int q = 0; q < JXP_NTHETA; ++q) acc9[1 + q] = g * (double)out[1 + q];
        }
        // segmented inclusive scan.  Entries of one set are contiguous within a tile, but with samples
        // carried over from earlier tiles of the run a set can occur in several separate pieces of a
        // batch, so the scan is keyed by the piece number (count of piece starts up to this lane).
        const int skey = have ? sl : -1 - lane;
        const int skey_prev = __shfl_up_sync(0xffffffffu, skey, 1);
        const unsigned starts = __ballot_sync(0xffffffffu, lane == 0 || skey != skey_prev);
        const int key = __popc(starts & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int ku = __shfl_up_sync(0xffffffffu, key, o);
          const bool take = lane >= o && ku == key;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const double u = __shfl_up_sync(0xffffffffu, acc9[q], o);
            if (take) acc9[q] += u;
          }
        }
        const int kn = __shfl_down_sync(0xffffffffu, key, 1);
        // With samples carried over from earlier tiles of the run the same set can end more than one
        // segment of a batch: those lanes take turns by lane index (= queue order, deterministic).
        const bool endl = have && (lane == 31 || kn != key);
        const unsigned same = __match_any_sync(0xffffffffu, endl ? sl : -1 - lane);
        const int rank = __popc(same & ((1u << lane) - 1u));
        const int maxr = (int)__reduce_max_sync(0xffffffffu, (unsigned)(endl ? rank : 0));
#pragma unroll 1
        for (int rr = 0; rr <= maxr; ++rr) {
          if (endl && rank == rr) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) sacc[sl * NQ + q] += acc9[q];
          }
          __syncwarp();
        }
      }
      // shift the rest of the queue down
      const int rem = qn - cnt;
      unsigned keep[(TR_QCAP - 32) / 32];
#pragma unroll
      for (int m = 0; m < (TR_QCAP - 32) / 32; ++m) keep[m] = (32 * m + lane < rem) ? sq[cnt + 32 * m + lane] : 0u;
      __syncwarp();
#pragma unroll
      for (int m = 0; m < (TR_QCAP - 32) / 32; ++m)
        if (32 * m + lane < rem) sq[32 * m + lane] = keep[m];
      qn = rem;
      __syncwarp();
    }
  }
  if (fill_pending) fill_wait();
  flush();
  if (LOGLIKE) {
    for (int i = lane; i < nsl * NQ; i += 32) {
      const int sl = i / NQ, q = i - sl * NQ;
      a.partial[((size_t)wrun * a.n_sets + (set_base + sl)) * NQ + q] = sacc[i];
    }
  }
}

// loglike[s] = -chi2/2 + const, dloglike[s][k] = sum of the per-tile gradient partials
// Block = 32 consecutive (set, q) entries x 8 tile slices; fixed summation order.
template <typename T>
__global__ void __launch_bounds__(256)
k_tr_final(const double* __restrict__ partial, int n_tiles, int n_sets,
           const double* __restrict__ scalars, T* __restrict__ loglike, T* __restrict__ dloglike,
           const unsigned long long* __restrict__ tl_state, unsigned long long tl_cap) {
  constexpr int NQ = 1 + JXP_NTHETA;
  __shared__ double sh[8][33];
  if (tl_state != nullptr && tl_active(tl_state, tl_cap)) return;  // the transit-list kernels did the work
  const int li = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + li;  // i = set * NQ + q
  const int n = n_sets * NQ;
  double acc = 0.0;
  if (i < n)
    for (int k = slice; k < n_tiles; k += 8) acc += partial[(size_t)k * n + i];  // coalesced over i
  sh[slice][li] = acc;
  __syncthreads();
  if (slice != 0 || i >= n) return;
  double tot = 0.0;
  for (int j = 0; j < 8; ++j) tot += sh[j][li];
  const int s = i / NQ, q = i % NQ;
  if (q == 0) {
    if (loglike) {
      double chi0 = 0.0, cst = 0.0;
      for (int k = 0; k < LL_NCTA_H; ++k) {
        chi0 += scalars[LL_SCAL0_H + 2 * k];
        cst += scalars[LL_SCAL0_H + 2 * k + 1];
      }
      loglike[s] = T(-0.5 * (chi0 + tot) + cst);
    }
  } else if (dloglike) {
    dloglike[(size_t)s * JXP_NTHETA + (q - 1)] = T(tot);
  }
}

// --------------------------------------------------------------------------------------
// Transit-list path of the gradient step (log-likelihood and its 8-gradient, no rows stored, no
// exposure integration, double, sorted time stamps) -- the K5 counterpart of k_tl_eval / k_tl_reduce in
// jxp_kernels.cu; k_tl_build (jxp_list.cuh) makes the two lists from the BodyC at the head of each
// BodyD record.  One item per lane (the dual-number code needs the registers); the chi^2 term and the
// 8 gradient terms of an item go to 9 planes of tl_vals at the item's slot, k_tlg_reduce adds each
// set's slices plane by plane in a fixed order.  Falls back to k_transit_partials / k_tr_final on the
// device exactly like the log-likelihood path (tl_active()).
// --------------------------------------------------------------------------------------
#ifndef JXP_TLG_MINB
#define JXP_TLG_MINB 4
#endif
__global__ void __launch_bounds__(128, JXP_TLG_MINB) k_tlg_eval(const __grid_constant__ TrArgs<double> a) {
  constexpr int NQ = 1 + JXP_NTHETA;
  if (!tl_active(a.tl_state, a.tl_cap)) return;
  const unsigned long long totA = a.tl_state[0], totB = a.tl_state[2];
  const unsigned long long nA = (totA + 31ull) >> 5, nB = (totB + 31ull) >> 5;
  const unsigned long long n_chunks = nA + nB;  // 32-item chunks of the inside list, then of the edge list
  const int lane = threadIdx.x & 31;
  const unsigned long long nw = (unsigned long long)gridDim.x * (blockDim.x >> 5);
  for (unsigned long long ch = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
       ch < n_chunks; ch += nw) {
    const bool on_inside_list = ch < nA;
    unsigned long long slot;
    bool have;
    if (on_inside_list) {
      slot = ch * 32ull + lane;
      have = slot < totA;
    } else {
      const unsigned long long p = (ch - nA) * 32ull + lane;
      slot = a.tl_cap - totB + p;
      have = p < totB;
    }
    double out[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = 0.0;
    if (have) {
      const unsigned long long item = a.tl_items[slot];
      const BodyD<double>& bd = a.bodies[(unsigned)(item >> 32)];
      const unsigned idx = (unsigned)item;
      double f1 = 0.0, p1[JXP_NTHETA];
      if (eval_sample_partials<double, 11>(bd, &a.tab, a.t[idx], f1, p1, on_inside_list)) {
        const double w = a.winv[idx];
        const double r0 = (a.y[idx] - 1.0) * w;
        const double mw = f1 * w;
        const double g = (r0 - mw) * w;  // d loglike / d model
        out[0] = (-mw) * (2.0 * r0 - mw);
#pragma unroll
        for (int q = 0; q < JXP_NTHETA; ++q) out[1 + q] = g * p1[q];
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) a.tl_vals[(size_t)q * a.tl_cap + slot] = out[q];
    }
  }
}

__global__ void __launch_bounds__(128) k_tlg_reduce(const __grid_constant__ TrArgs<double> a) {
  constexpr int NQ = 1 + JXP_NTHETA;
  if (!tl_active(a.tl_state, a.tl_cap)) return;
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= a.n_sets) return;
  const unsigned cntA = a.tl_cnt[2 * s], cntB = a.tl_cnt[2 * s + 1];
  const unsigned long long offA = a.tl_off[2 * s], offB = a.tl_cap - a.tl_off[2 * s + 1] - cntB;
  double tot[NQ];
#pragma unroll 1
  for (int q = 0; q < NQ; ++q) {
    const double* __restrict__ d = a.tl_vals + (size_t)q * a.tl_cap;
    double v = tl_slice_sum(d + offA, cntA, lane);
    v += tl_slice_sum(d + offB, cntB, lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    tot[q] = v;
  }
  if (lane == 0) {
    if (a.loglike) {
      double chi0 = 0.0, cst = 0.0;
      for (int k = 0; k < LL_NCTA_H; ++k) {
        chi0 += a.scalars[LL_SCAL0_H + 2 * k];
        cst += a.scalars[LL_SCAL0_H + 2 * k + 1];
      }
      a.loglike[s] = -0.5 * (chi0 + tot[0]) + cst;
    }
    if (a.dloglike) {
#pragma unroll
      for (int q = 0; q < JXP_NTHETA; ++q) a.dloglike[(size_t)s * JXP_NTHETA + q] = tot[1 + q];
    }
  }
}

extern "C" size_t jxp_transit_workspace_bytes(int32_t n_sets, int64_t n_times) {
  if (n_sets <= 0 || n_times < 0) return 0;
  return tr_layout(n_sets, n_times).total;
}

namespace jxph {
// launches k_loglike_prep<T> (instantiated in jxp_kernels.cu)
int launch_loglike_prep_f64(const double* y, const double* sigma, int64_t ss, int64_t n, double* winv,
                            double* scalars, cudaStream_t st, const double* t_sorted_check = nullptr);
int launch_loglike_prep_f32(const float* y, const float* sigma, int64_t ss, int64_t n, float* winv,
                            double* scalars, cudaStream_t st);
}  // namespace jxph

static int loglike_prep(const double* y, const double* s, int64_t ss, int64_t n, double* w, double* sc,
                        cudaStream_t st, const double* t_check) {
  return launch_loglike_prep_f64(y, s, ss, n, w, sc, st, t_check);
}
static int loglike_prep(const float* y, const float* s, int64_t ss, int64_t n, float* w, double* sc,
                        cudaStream_t st, const double*) {
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

  unsigned long long* tl_state = reinterpret_cast<unsigned long long*>(d_scalars + 4);
  // transit-list path (see k_tlg_eval): eligibility is decided here, use on the device
  const bool use_tl = sizeof(T) == 8 && want_ll && !flux && !partials && a.st.n == 1 && gl_order == 10 &&
                      !(flags & JXP_FLAG_NO_WINDOW) && n_times < 0xfffffe00LL && !(dev_options() & JXP_DEV_NO_WALK);
  k_tr_prep<T><<<(n_sets + 127) / 128, 128, 0, st>>>(theta, n_sets, star, star_stride, d_bodies,
                                                    f32 ? 1.0 + 2e-4 : 1.0 + 1e-9, 1e-7, tl_state);
  rc = check_launch("jxp_transit_partials(prep)");
  if (rc != JXP_OK) return rc;
  if (want_ll) {
    rc = loglike_prep(y, sigma, sigma_stride, n_times, d_winv, d_scalars, st, use_tl ? t : nullptr);
    if (rc != JXP_OK) return rc;
  }
  const int sms = sm_count();
  int threads = 128;
  const int64_t n_wtiles = (n_times + 32 * TR_K - 1) / (32 * TR_K);
  if (n_times >= (1LL << 26)) return fail(JXP_ERR_UNSUPPORTED, "jxp_transit_partials: n_times >= 2^26");
  // a warp processes `run` consecutive warp tiles and carries queued samples from one to the next
  // (full 32-wide drains); shorter runs when the problem is too small to fill the machine otherwise
  // measured on config 5 (run = 1 / 4 / 16): the arithmetic alone 6.8 / 5.4 / 4.9 ms, the row stores alone
  // 8.7 / 9.9 / 10.8 ms -- long runs keep the drains full, short ones keep the store stream smooth
  int run = (flux || partials) ? 1 : 16;
  {  // development option: warp tiles per run
    const int v = (int)JXP_DEV_GET_TR_RUN(dev_options());
    if (v >= 1 && v <= 64) run = v;
  }
  auto n_wruns_of = [&](int r) { return (n_wtiles + r - 1) / r; };
  while (run > 1 && n_wruns_of(run) / 4 * ((n_sets + TR_MAX_SETS - 1) / TR_MAX_SETS) < 8 * sms) run /= 2;
  const int64_t n_wruns = n_wruns_of(run);
  int64_t n_tiles = (n_wruns + threads / 32 - 1) / (threads / 32);
  while (threads > 32 && n_tiles * ((n_sets + TR_MAX_SETS - 1) / TR_MAX_SETS) < 4 * sms) {
    threads /= 2;
    n_tiles = (n_wruns + threads / 32 - 1) / (threads / 32);
  }
  int spc = TR_MAX_SETS;
  if (spc > n_sets) spc = n_sets;
  const int64_t n_groups = (n_sets + spc - 1) / spc;
  int64_t n_ctas = n_tiles * n_groups;
  if (n_ctas > 2147483647LL) return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials: grid too large");
  a.bodies = d_bodies;
  a.t = t;
  a.n_times = n_times;
  a.n_sets = n_sets;
  a.sets_per_cta = spc;
  a.n_tiles = (int)n_tiles;
  a.n_wtiles = (int)n_wtiles;
  a.run = run;
  // (bulk copies need 16-byte aligned row pieces: a tile starts at a multiple of 64 samples)
  // In double the copy queue is the bottleneck (59 GB on config 5) and the issuing threads block on it: mode 2
  // (one issuing warp, results buffered) 10.7 -> 10.0 ms; in float the arithmetic is, and the plain mode 1 is the
  // faster one (6.3 -> 5.8 ms).
  a.bulk_fill = (run == 1 && (flux || partials) && !(dev_options() & JXP_DEV_NO_BULK_FILL) &&
                 (n_times * (int64_t)sizeof(T)) % 16 == 0 && ((uintptr_t)flux & 15) == 0 && ((uintptr_t)partials & 15) == 0)
                    ? (((sizeof(T) == 8) != ((dev_options() & JXP_DEV_BULK_FILL_SWAP) != 0)) ? 2 : 1)
                    : 0;
  a.tpw = 1;
  if (a.bulk_fill) {
    // (development option: two warp tiles per warp -- 4 KB / 2 KB copies, queued samples carried from one tile to the
    // next -- measured slower on config 5: 10.13 against 9.96 ms in double, 6.0-6.8 against 5.84 ms in float)
    int tpw = 1;
    {
      const int v = (int)JXP_DEV_GET_TR_TPW(dev_options());
      if (v >= 1 && v <= TR_TPW_MAX) tpw = v;
    }
    const int nw = threads / 32;
    while (tpw > 1 && ((n_wtiles + nw * tpw - 1) / (nw * tpw)) * ((n_sets + TR_MAX_SETS - 1) / TR_MAX_SETS) < 8 * sms) --tpw;
    a.tpw = tpw;
    n_tiles = (n_wtiles + nw * tpw - 1) / (nw * tpw);
    a.n_tiles = (int)n_tiles;
    n_ctas = n_tiles * n_groups;
  }
  a.use_window = (flags & JXP_FLAG_NO_WINDOW) ? 0 : 1;
  a.flux = flux;
  a.partials = partials;
  a.y = y;
  a.winv = d_winv;
  a.partial = d_partial;
  a.tl_state = nullptr;
  a.tl_items = nullptr;
  a.tl_off = nullptr;
  a.tl_cnt = nullptr;
  a.tl_cap = 0;
  a.tl_vals = nullptr;
  a.scalars = d_scalars;
  a.loglike = loglike;
  a.dloglike = dloglike;
  a.tab.order = gl_order;
  if (gl_order != 10) gl_table_host(gl_order, a.tab);
  if constexpr (sizeof(T) == 8) {
    if (use_tl) {
      a.tl_state = tl_state;
      a.tl_items = reinterpret_cast<unsigned long long*>(ws + L.off_tl_items);
      a.tl_off = reinterpret_cast<unsigned long long*>(ws + L.off_tl_off);
      a.tl_cnt = reinterpret_cast<unsigned*>(ws + L.off_tl_cnt);
      a.tl_cap = L.tl_cap;
      a.tl_vals = reinterpret_cast<double*>(ws + L.off_tl_vals);
      TlBuildArgs ba;
      ba.bodies = reinterpret_cast<const unsigned char*>(d_bodies);
      ba.body_stride = sizeof(BodyD<T>);
      ba.n_sets = n_sets;
      ba.t = t;
      ba.n_times = n_times;
      ba.sorted_flags = d_scalars + LL_FLAG0_H;
      ba.tl_state = a.tl_state;
      ba.tl_items = a.tl_items;
      ba.tl_off = a.tl_off;
      ba.tl_cnt = a.tl_cnt;
      ba.tl_cap = a.tl_cap;
      k_tl_build<1><<<(n_sets + 3) / 4, 128, 0, st>>>(ba);
      rc = check_launch("jxp_transit_partials(transit list)");
      if (rc != JXP_OK) return rc;
      int per_sm = 0;
      cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tlg_eval, 128, 0);
      if (eo != cudaSuccess || per_sm < 1)
        return fail(JXP_ERR_CUDA, "jxp_transit_partials: occupancy query: %s", cudaGetErrorString(eo));
      // the list length is only known on the device: a resident grid, chunks strided over it
      const int64_t max_chunks = (int64_t)((L.tl_cap + 31ull) / 32ull) + 1;
      const int64_t want = (max_chunks + 3) / 4;
      const int64_t resident = (int64_t)per_sm * sms;
      k_tlg_eval<<<(unsigned)(want < resident ? want : resident), 128, 0, st>>>(a);
      rc = check_launch("jxp_transit_partials(transit list eval)");
      if (rc != JXP_OK) return rc;
      k_tlg_reduce<<<(n_sets + 3) / 4, 128, 0, st>>>(a);
      rc = check_launch("jxp_transit_partials(transit list reduce)");
      if (rc != JXP_OK) return rc;
    }
  }
  const size_t smem = sizeof(BodyD<T>) * (size_t)spc +
                      (want_ll ? sizeof(double) * (1 + JXP_NTHETA) * (size_t)spc * (threads / 32) : 0) +
                      sizeof(unsigned) * (size_t)TR_QCAP * (threads / 32) +
                      (a.bulk_fill == 2 ? (sizeof(T) * (1 + JXP_NTHETA) + sizeof(unsigned)) * (size_t)TR_RB * (threads / 32) : 0);
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<(unsigned)n_ctas, threads, smem, st>>>(a);
    return check_launch("jxp_transit_partials");
  };
  if (gl_order == 10)
    rc = want_ll ? launch(k_transit_partials<T, 11, true>) : launch(k_transit_partials<T, 11, false>);
  else
    rc = want_ll ? launch(k_transit_partials<T, 0, true>) : launch(k_transit_partials<T, 0, false>);
  if (rc != JXP_OK) return rc;
  if (want_ll) {
    const int n = n_sets * (1 + JXP_NTHETA);
    k_tr_final<T><<<(n + 31) / 32, 256, 0, st>>>(d_partial, (int)(n_tiles * (threads / 32)), n_sets, d_scalars, loglike,
                                                  dloglike, a.tl_state, a.tl_cap);
    rc = check_launch("jxp_transit_partials(final)");
  }
  return rc;
}


// --------------------------------------------------------------------------------------
// General polynomial limb darkening (any number of coefficients up to JXP_MAX_LD_COEFFS): flux and partials
// w.r.t. theta = (period, time_transit, eccentricity, omega_peri, impact_param, radius, u_1 .. u_n).
// The reference differentiates light_curve(u, b, r) for any u (core/limb_dark.py:19-47, 87-99, 140-182); here the
// orbit tangents are those of eval_sample_partials, d flux / d(b, r) comes from dual numbers over the full solution
// vector (the HI flux code: l = 3 .. n_u terms of p_integral), d flux / d u_j = s . d(g / norm) / d u_j.
// A plain kernel (one sample per thread, window test per sample, rows stored coalesced): the fused quadratic path
// above stays the tuned one.
// --------------------------------------------------------------------------------------
namespace {
constexpr int POLY_NP_MAX = 6 + JXP_MAX_LD_COEFFS;
constexpr int POLY_TILE = 256;

template <int ORDER>
__device__ __forceinline__ bool eval_sample_partials_poly(const BodyD<double>& bd, const PolyD& pd, int n_u,
                                                          const GLTable* __restrict__ tab, double tm, double& flux,
                                                          double* __restrict__ part) {
  const BodyC<double>& c = bd.c;
  typedef Dual<double, NO> D;
  const double x = (tm - c.t0) - c.tref;
  const double M = (Num<double>::two_pi * x) / c.period;
  const double k = Num<double>::two_pi / c.period;
  double Md[NO];
#pragma unroll
  for (int i = 0; i < NO; ++i) Md[i] = -k * bd.tref.d[i];
  Md[0] += -M / c.period;
  Md[1] += -k;
  const double Mr = mod_two_pi(M);
  double sf, cf, Edummy;
  kepler_reduced<double, false>(Mr, c.e, c.ome, c.ope, c.s1me2, c.inv_ope, sf, cf, Edummy);
  // kepler.py:64-79
  const double ecosf = c.e * cf;
  const double opc = 1.0 + ecosf;
  const double dfdM = opc * opc * bd.inv_ome2_32;
  const double dfde = (2.0 + ecosf) * sf * bd.inv_ome2;
  D sfd, cfd;
  sfd.v = sf;
  cfd.v = cf;
#pragma unroll
  for (int i = 0; i < NO; ++i) {
    double fdot = Md[i] * dfdM;
    if (i == 2) fdot += dfde;
    sfd.d[i] = cf * fdot;
    cfd.d[i] = -sf * fdot;
  }
  const D ed = D::seed(c.e, 2);
  D b, Z;
  sky_position_g<D, double>(ed, bd.cw, bd.sw, bd.ci, bd.si, bd.na_ome2, c.inv_rstar, sfd, cfd, b, Z);
  if (!(Z.v > 0.0) || (b.v >= c.onepr)) return false;
  typedef Dual<double, 2> L;
  L sv[JXP_MAX_LD_COEFFS + 1];
  for (int n = 0; n <= JXP_MAX_LD_COEFFS; ++n) sv[n] = L(0.0);
  ld_solution<L, ORDER, true>(L::seed(b.v, 0), L::seed(c.rr, 1), n_u, tab, sv[0], sv[1], sv[2], sv + 3);
  // flux = s . g - 1 (limb_dark.py:45-47); only s_0 .. s_max(n_u, 0) enter
  double f = -1.0, fb = 0.0, fr = 0.0;
  for (int n = 0; n <= n_u; ++n) {
    f = fma(sv[n].v, pd.g[n], f);
    fb = fma(sv[n].d[0], pd.g[n], fb);
    fr = fma(sv[n].d[1], pd.g[n], fr);
  }
  flux = f;
#pragma unroll
  for (int i = 0; i < NO; ++i) part[i] = fb * b.d[i];
  part[5] = fr * c.inv_rstar;
  for (int j = 0; j < n_u; ++j) {
    double acc = 0.0;
    for (int n = 0; n <= n_u; ++n) acc = fma(sv[n].v, pd.dg[n][j], acc);
    part[6 + j] = acc;
  }
  return true;
}

struct PolyArgs {
  const BodyD<double>* bodies;
  const PolyD* poly;
  const double* t;
  int64_t n_times;
  int n_sets, n_u, n_tiles, use_window;
  double* flux;      // [n_sets][n_times] or nullptr
  double* partials;  // [n_sets][6 + n_u][n_times] or nullptr
  const double* y;
  const double* winv;
  double* partial;   // [n_sets][n_tiles][1 + 6 + n_u] per-tile sums (log-likelihood calls)
  GLTable tab;
};

// one sample per thread, POLY_TILE consecutive samples of one set per CTA
template <int ORDER, bool LOGLIKE>
__global__ void __launch_bounds__(POLY_TILE) k_poly_rows(const __grid_constant__ PolyArgs a) {
  __shared__ double s_red[POLY_TILE / 32];
  const int set = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * POLY_TILE + threadIdx.x;
  const int np = 6 + a.n_u;
  const BodyD<double>& bd = a.bodies[set];
  double flux = 0.0, part[POLY_NP_MAX];
  for (int k = 0; k < POLY_NP_MAX; ++k) part[k] = 0.0;
  bool hit = false;
  if (i < a.n_times) {
    const double tm = a.t[i];
    if (!a.use_window || (bd.c.flags & BODY_ALWAYS) || !outside_window(bd.c, tm))
      hit = eval_sample_partials_poly<ORDER>(bd, a.poly[set], a.n_u, &a.tab, tm, flux, part);
    if (!hit) flux = 0.0;
    if (a.flux) a.flux[(size_t)set * a.n_times + i] = flux;
    if (a.partials)
      for (int k = 0; k < np; ++k) a.partials[((size_t)set * np + k) * a.n_times + i] = hit ? part[k] : 0.0;
  }
  if (LOGLIKE) {
    // per-tile sums in a fixed order: chi^2 term, then (y - 1 - f) / sigma^2 . d f / d theta_k
    const bool any_hit = __syncthreads_or(hit ? 1 : 0) != 0;
    double w = 0.0, res = 0.0;
    if (i < a.n_times) {
      w = a.winv[i];
      res = ((a.y[i] - 1.0) - flux) * w;
    }
    for (int q = 0; q <= (any_hit ? np : 0); ++q) {
      double v = q == 0 ? res * res : (hit ? res * w * part[q - 1] : 0.0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < POLY_TILE / 32; ++k) tot += s_red[k];
        a.partial[((size_t)set * a.n_tiles + blockIdx.x) * (1 + POLY_NP_MAX) + q] = tot;
      }
      __syncthreads();
    }
    if (!any_hit && threadIdx.x < np)
      a.partial[((size_t)set * a.n_tiles + blockIdx.x) * (1 + POLY_NP_MAX) + 1 + threadIdx.x] = 0.0;
  }
}

// per (set, quantity): the tile sums added in tile order -> loglike, dloglike
__global__ void __launch_bounds__(128) k_poly_final(const double* __restrict__ partial, int n_sets, int n_tiles,
                                                    int np, const double* __restrict__ scalars,
                                                    double* __restrict__ loglike, double* __restrict__ dloglike) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_sets * (1 + np)) return;
  const int set = idx / (1 + np), q = idx % (1 + np);
  const double* p = partial + (size_t)set * n_tiles * (1 + POLY_NP_MAX) + q;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k = 0; k < n_tiles; k += 4)
    for (int m = 0; m < 4; ++m)
      if (k + m < n_tiles) acc[m] += p[(size_t)(k + m) * (1 + POLY_NP_MAX)];
  const double tot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  if (q == 0) {
    if (loglike) {
      double cst = 0.0;  // sum_t (-log sigma - log(2 pi) / 2): the per-CTA sums of the likelihood prep, in order
      for (int k = 0; k < LL_NCTA_H; ++k) cst += scalars[LL_SCAL0_H + 2 * k + 1];
      loglike[set] = -0.5 * tot + cst;
    }
  } else if (dloglike) {
    dloglike[(size_t)set * np + (q - 1)] = tot;
  }
}

struct PolyLayout {
  size_t off_bodies, off_poly, off_winv, off_scalars, off_partial, total;
  int n_tiles;
};
PolyLayout poly_layout(int n_sets, int64_t n_times) {
  PolyLayout L;
  L.n_tiles = (int)((n_times + POLY_TILE - 1) / POLY_TILE);
  size_t off = 0;
  L.off_bodies = off;
  off = align_up(off + sizeof(BodyD<double>) * (size_t)n_sets, 256);
  L.off_poly = off;
  off = align_up(off + sizeof(PolyD) * (size_t)n_sets, 256);
  L.off_winv = off;
  off = align_up(off + sizeof(double) * (size_t)n_times, 256);
  L.off_scalars = off;
  off = align_up(off + 256 * sizeof(double), 256);
  L.off_partial = off;
  off = align_up(off + sizeof(double) * (size_t)n_sets * L.n_tiles * (1 + POLY_NP_MAX), 256);
  L.total = off;
  return L;
}
}  // namespace

extern "C" size_t jxp_transit_poly_workspace_bytes(int32_t n_sets, int64_t n_times) {
  if (n_sets <= 0 || n_times <= 0) return 0;
  return poly_layout(n_sets, n_times).total;
}

extern "C" int jxp_transit_partials_poly_f64(const double* theta, int32_t n_sets, int32_t n_u, const double* star,
                                             int64_t star_stride, const double* t, int64_t n_times,
                                             int32_t gl_order, int32_t flags, double* flux, double* partials,
                                             const double* y, const double* sigma, int64_t sigma_stride,
                                             double* loglike, double* dloglike, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  if (n_sets < 0 || n_times < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials_poly: negative size");
  if (n_u < 0 || n_u > JXP_MAX_LD_COEFFS)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials_poly: n_u = %d outside 0..%d", n_u, JXP_MAX_LD_COEFFS);
  if (gl_order != 10) return fail(JXP_ERR_UNSUPPORTED, "jxp_transit_partials_poly: order must be 10");
  if (star_stride != 0 && star_stride != 2)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials_poly: star_set_stride must be 0 or 2");
  if (!stride_ok(sigma_stride))
    return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_partials_poly: sigma_stride must be 0 or 1");
  if (n_sets == 0 || n_times == 0) return JXP_OK;
  if (!theta || !star || !t) return fail(JXP_ERR_NULL_POINTER, "jxp_transit_partials_poly: null theta/star/t");
  const bool want_ll = loglike || dloglike;
  if (want_ll && (!y || !sigma))
    return fail(JXP_ERR_NULL_POINTER, "jxp_transit_partials_poly: loglike needs y and sigma");
  if (!flux && !partials && !want_ll)
    return fail(JXP_ERR_NULL_POINTER, "jxp_transit_partials_poly: no output requested");
  const PolyLayout L = poly_layout(n_sets, n_times);
  if (!workspace || workspace_bytes < L.total)
    return fail(JXP_ERR_WORKSPACE, "jxp_transit_partials_poly: workspace of %zu bytes needed, %zu given", L.total,
                workspace_bytes);
  if (((uintptr_t)workspace & 255) != 0)
    return fail(JXP_ERR_WORKSPACE, "jxp_transit_partials_poly: workspace must be 256-byte aligned");
  if (n_sets > 65535) return fail(JXP_ERR_UNSUPPORTED, "jxp_transit_partials_poly: more than 65535 sets per call");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  PolyArgs a;
  a.bodies = reinterpret_cast<BodyD<double>*>(ws + L.off_bodies);
  a.poly = reinterpret_cast<PolyD*>(ws + L.off_poly);
  a.t = t;
  a.n_times = n_times;
  a.n_sets = n_sets;
  a.n_u = n_u;
  a.n_tiles = L.n_tiles;
  a.use_window = (flags & JXP_FLAG_NO_WINDOW) ? 0 : 1;
  a.flux = flux;
  a.partials = partials;
  a.y = y;
  a.winv = reinterpret_cast<double*>(ws + L.off_winv);
  a.partial = reinterpret_cast<double*>(ws + L.off_partial);
  a.tab.order = 10;
  double* d_scalars = reinterpret_cast<double*>(ws + L.off_scalars);
  k_tr_prep<double><<<(n_sets + 127) / 128, 128, 0, st>>>(theta, n_sets, star, star_stride,
                                                         const_cast<BodyD<double>*>(a.bodies), 1.0 + 1e-9, 1e-7,
                                                         nullptr, 6 + n_u, n_u, const_cast<PolyD*>(a.poly));
  int rc = check_launch("jxp_transit_partials_poly(prep)");
  if (rc != JXP_OK) return rc;
  if (want_ll) {
    rc = loglike_prep(y, sigma, sigma_stride, n_times, const_cast<double*>(a.winv), d_scalars, st, nullptr);
    if (rc != JXP_OK) return rc;
  }
  const dim3 grid((unsigned)L.n_tiles, (unsigned)n_sets);
  if (want_ll)
    k_poly_rows<10, true><<<grid, POLY_TILE, 0, st>>>(a);
  else
    k_poly_rows<10, false><<<grid, POLY_TILE, 0, st>>>(a);
  rc = check_launch("jxp_transit_partials_poly(rows)");
  if (rc != JXP_OK) return rc;
  if (want_ll) {
    const int np = 6 + n_u;
    k_poly_final<<<(n_sets * (1 + np) + 127) / 128, 128, 0, st>>>(a.partial, n_sets, L.n_tiles, np, d_scalars, loglike,
                                                                 dloglike);
    rc = check_launch("jxp_transit_partials_poly(final)");
  }
  return rc;
}

extern "C" int jxp_transit_list_stats(const void* workspace, int32_t n_sets, int64_t n_times,
                                      int64_t* stats_host, void* stream) {
  if (!workspace || !stats_host) return fail(JXP_ERR_NULL_POINTER, "jxp_transit_list_stats: null pointer");
  if (n_sets <= 0 || n_times <= 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_transit_list_stats: empty problem");
  const TrLayout L = tr_layout(n_sets, n_times);
  const unsigned char* ws = static_cast<const unsigned char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t raw[4] = {0, 0, 0, 0};  // inside-list items, flags, edge-list items, unused
  cudaError_t e = cudaMemcpyAsync(raw, ws + L.off_scalars + 4 * sizeof(double), 4 * sizeof(int64_t),
                                  cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_transit_list_stats: %s", cudaGetErrorString(e));
  stats_host[0] = raw[0] + raw[2];
  stats_host[1] = raw[1] | ((unsigned long long)(raw[0] + raw[2]) > L.tl_cap ? 2 : 0);
  stats_host[2] = raw[0];
  stats_host[3] = 0;
  return JXP_OK;
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
