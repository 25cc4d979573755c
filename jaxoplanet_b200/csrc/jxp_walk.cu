// Transit walk: the light curve / Gaussian log-likelihood of single-planet parameter sets over
// sorted time stamps (the headline case, BASELINE config 3) without ever looking at a sample that
// cannot be in transit and without materialising a work list.
//
// Reference path: light_curves/limb_dark.py:15-62 over orbits/keplerian.py:402-419, 537-637 and
// core/kepler.py:14-108, core/limb_dark.py:19-196 -- the reference evaluates the full formula at
// every sample and masks (limb_dark.py:58); out of transit the result is exactly 0.
//
//   k_walk_plan  one warp per parameter set: OrbitalBody constants -> BodyC / SetC (for the
//                fallback kernels) and a 32-double set record; the transits that intersect
//                [t_first, t_last] -> sample index ranges by search on the sorted time stamps
//                (the same rigorous window as the window-test kernels) -> a few dozen
//                (first position, first sample) SEGMENTS per set, on two lists: samples estimated
//                to have the planet fully inside the disk, and ingress / egress samples.  Positions
//                on a list are cut into CHUNKS of 32 N; every chunk gets a 16-byte entry (set, first
//                position, first segment).  The same launch derives the set-independent likelihood
//                terms (LL_NCTA extra CTAs).
//   k_walk_eval  a flat map over the chunk entries, N items per lane.  A chunk belongs to ONE set, so
//                everything that depends on the set is warp-uniform: the record is staged once per
//                chunk in shared memory, per-item work is t -> M -> Kepler -> separation -> flux ->
//                chi^2 term.  Chunks of the inside list take the constant-kappa flux code.  The chunk's
//                terms are added by a fixed shuffle tree into one partial; the warp that completes a
//                set's last chunk adds the set's partials in a fixed order and writes the
//                log-likelihood: bit-reproducible, no float atomics, no separate reduction launch.
// Unsorted / NaN time stamps or a table that does not fit are detected on the device; the kernels
// then return at once and the window-test kernels launched behind them do the work.
#include "../../include/jaxoplanet_b200.h"

#include <cuda_runtime.h>

#include "jxp_host.h"
#include "jxp_list.cuh"
#include "jxp_orbit.cuh"
#include "jxp_prep.cuh"

using namespace jxp;
using namespace jxph;

namespace jxp {

// fields of the 32-double set record
enum : int {
  WK_T0 = 0, WK_TREF, WK_PERIOD, WK_INVP, WK_E, WK_OME, WK_OPE, WK_S1ME2, WK_INVOPE,
  WK_CW, WK_SW, WK_K1, WK_K2, WK_ZS, WK_RR, WK_ONEPR, WK_G0, WK_G1, WK_G2, WK_FLAGS,
  WK_NFIELDS
};
static_assert(WK_NFIELDS <= WALK_REC, "record size");
constexpr unsigned WKF_CIRCULAR = 1u, WKF_SLOW_STARTER = 2u;

struct WalkSet {
  unsigned seg_off;   // first segment of the inside list; the edge list follows its sentinel
  unsigned n_tr;      // transits (inside list: n_tr segments, edge list: 2 n_tr)
  unsigned cntA, cntB;    // positions on the inside / edge list
  unsigned chA, chB;      // first chunk slot on the inside list (from the bottom) / edge list (from the top)
  unsigned n_chunks;      // chunks of both lists
  unsigned done;          // chunks finished (k_walk_eval)
};
static_assert(sizeof(WalkSet) == 32, "WalkSet");

struct WalkArgs {
  // inputs of the call
  const double* bodies;  // [n_sets][JXP_BODY_NFIELDS]
  const double* u;
  int n_u;
  int64_t u_stride;
  const double* t;
  int64_t n_times;
  int n_sets;
  const double* y;
  const double* sigma;
  int64_t sigma_stride;
  // workspace
  BodyC<double>* out_b;
  SetC<double>* out_s;
  unsigned long long* counters;
  int* tile_ctr;
  double* winv;
  double* scalars;
  double* rec;            // [n_sets][WALK_REC]
  WalkSet* wset;          // [n_sets]
  unsigned long long* seg;   // (cum | start << 32) segments
  uint4* chunk;           // chunk entries: inside list from 0 up, edge list from cap down
  double* part;           // per-chunk partial sums (same slots)
  unsigned long long* state;  // [0] segments allocated, [1] inside chunks, [2] edge chunks, [3] flags,
                              // [4] positions on the inside lists, [5] on the edge lists (zeroed per call)
  unsigned long long cap_seg, cap_chunk;
  int n_plan_ctas;
  double widen, margin;
  // outputs
  double* flux;
  double* flux_sum;
  double* loglike;
};

constexpr unsigned long long WK_OVERFLOW = 2ull;

// --------------------------------------------------------------------------------------------
// plan
// --------------------------------------------------------------------------------------------
template <int CH, bool LL>
__global__ void __launch_bounds__(256) k_walk_plan(const __grid_constant__ WalkArgs a) {
  if ((int)blockIdx.x >= a.n_plan_ctas) {
    const int cta = (int)blockIdx.x - a.n_plan_ctas;
    if (LL) {
      loglike_prep_body<double>(cta, LL_NCTA, a.y, a.sigma, a.sigma_stride, a.n_times, a.winv, a.scalars, a.t);
    } else {
      int bad = 0;
      const int64_t n = a.n_times;
      const int64_t chunk = (n + LL_NCTA - 1) / LL_NCTA;
      const int64_t lo = (int64_t)cta * chunk;
      const int64_t hi = lo + chunk < n ? lo + chunk : n;
      for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
        if (i + 1 < n && !(a.t[i + 1] >= a.t[i])) bad = 1;
      bad = __syncthreads_or(bad);
      if (threadIdx.x == 0) a.scalars[LL_FLAG0 + cta] = bad ? 1.0 : 0.0;
    }
    return;
  }
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= a.n_sets) return;
  if (lane == 0)
    sys_prep_body<double>(s, a.bodies, a.n_sets, 1, a.u, a.n_u, a.u_stride, a.out_b, a.out_s, a.counters,
                          a.tile_ctr, a.widen, a.margin);
  __syncwarp();
  const BodyC<double>& c = a.out_b[s];
  if (lane == 0) {
    const SetC<double>& sc = a.out_s[s];
    double* r = a.rec + (size_t)s * WALK_REC;
    r[WK_T0] = c.t0;
    r[WK_TREF] = c.tref;
    r[WK_PERIOD] = c.period;
    r[WK_INVP] = 1.0 / c.period;
    r[WK_E] = c.e;
    r[WK_OME] = c.ome;
    r[WK_OPE] = c.ope;
    r[WK_S1ME2] = c.s1me2;
    r[WK_INVOPE] = c.inv_ope;
    r[WK_CW] = c.cw;
    r[WK_SW] = c.sw;
    r[WK_K1] = c.na * c.inv_rstar;         // x1 / R* = k1 (cw xD - sw yD)
    r[WK_K2] = c.ci * c.na * c.inv_rstar;  // Y / R*  = k2 (sw xD + cw yD)
    r[WK_ZS] = -c.si * c.na;               // Z = zs (sw xD + cw yD)
    r[WK_RR] = c.rr;
    r[WK_ONEPR] = c.onepr;
    r[WK_G0] = sc.g0;
    r[WK_G1] = sc.g1;
    r[WK_G2] = sc.g2;
    unsigned f = (c.flags & BODY_CIRCULAR) ? WKF_CIRCULAR : 0u;
    if (!(c.ome > 1e-5)) f |= WKF_SLOW_STARTER;
    f |= (unsigned)sc.lmax << 8;
    r[WK_FLAGS] = __longlong_as_double((long long)f);
#pragma unroll
    for (int k = WK_NFIELDS; k < WALK_REC; ++k) r[k] = 0.0;
  }

  // transits that intersect [t_first, t_last] (as k_tl_build, jxp_list.cuh)
  const unsigned n = (unsigned)a.n_times;
  const double P = c.period, t0ref = c.t0ref, phc = c.phase_c, wlo = c.wlo, whi = c.whi;
  const double t_first = a.t[0], t_last = a.t[n - 1];
  const double inv_dt = (t_last > t_first) ? (double)(n - 1) / (t_last - t_first) : 0.0;
  bool all = (c.flags & BODY_ALWAYS) != 0;
  long long n_tr = 1;
  double k_first = 0.0;
  if (!all) {
    if (!(wlo <= whi)) {
      n_tr = 0;  // never in transit
    } else {
      const double pa = (t_first - t0ref) * c.inv_period - phc;
      const double pb = (t_last - t0ref) * c.inv_period - phc;
      const double lo_d = floor(fmin(pa, pb) - whi) - 1.0;
      const double hi_d = ceil(fmax(pa, pb) - wlo) + 1.0;
      if (!(hi_d - lo_d < 4.0 * (double)n + 64.0)) {
        all = true;  // more transits than samples (or not finite): take every sample
      } else {
        k_first = lo_d;
        n_tr = (long long)(hi_d - lo_d) + 1;
      }
    }
  }
  const double w_in = all ? 0.0 : c.win_in;
  auto interval = [&](long long k, unsigned& lo, unsigned& hi, unsigned& ilo, unsigned& ihi) {
    lo = hi = ilo = ihi = 0;
    if (k >= n_tr) return;
    if (all) {
      hi = n;
      return;
    }
    const double base = phc + (k_first + (double)k);
    const double ta = fma(base + wlo, P, t0ref), tb = fma(base + whi, P, t0ref);
    lo = tl_search<false>(a.t, n, fmin(ta, tb), t_first, inv_dt);
    hi = tl_search<true>(a.t, n, fmax(ta, tb), t_first, inv_dt);
    if (hi < lo) hi = lo;  // unsorted time stamps: anything in range will do, the call falls back
    ilo = ihi = lo;
    if (w_in > 0.0 && hi - lo >= 8u) {
      const double tc = fma(base - w_in, P, t0ref), td = fma(base + w_in, P, t0ref);
      unsigned i0 = tl_search<false>(a.t, n, fmin(tc, td), t_first, inv_dt);
      unsigned i1 = tl_search<true>(a.t, n, fmax(tc, td), t_first, inv_dt);
      i0 = i0 < lo ? lo : (i0 > hi ? hi : i0);
      i1 = i1 < i0 ? i0 : (i1 > hi ? hi : i1);
      if (i1 - i0 >= 8u) {
        ilo = i0;
        ihi = i1;
      }
    }
  };

  // segments: inside list [seg0, seg0 + n_tr] (sentinel last), edge list [seg0 + n_tr + 1, seg0 + 3 n_tr + 1]
  const unsigned long long n_seg = 3ull * (unsigned long long)n_tr + 2ull;
  unsigned long long seg0 = 0;
  if (lane == 0) seg0 = atomicAdd(a.state, n_seg);
  seg0 = __shfl_sync(0xffffffffu, seg0, 0);
  WalkSet ws;
  ws.seg_off = (unsigned)seg0;
  ws.n_tr = (unsigned)n_tr;
  ws.cntA = ws.cntB = ws.chA = ws.chB = ws.n_chunks = ws.done = 0u;
  if (seg0 + n_seg > a.cap_seg) {
    if (lane == 0) {
      atomicOr(a.state + 3, WK_OVERFLOW);
      a.wset[s] = ws;
    }
    return;
  }
  unsigned long long* segA = a.seg + seg0;
  unsigned long long* segB = segA + n_tr + 1;
  unsigned long long cumA = 0, cumB = 0;  // running totals (64-bit: a list may exceed 2^32 only in theory)
  for (long long k0 = 0; k0 < n_tr; k0 += 32) {
    unsigned lo, hi, il, ih;
    interval(k0 + lane, lo, hi, il, ih);
    const unsigned cA = ih - il, cB0 = il - lo, cB1 = hi - ih;
    unsigned inA = cA, inB = cB0 + cB1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned vA = __shfl_up_sync(0xffffffffu, inA, o);
      const unsigned vB = __shfl_up_sync(0xffffffffu, inB, o);
      if (lane >= o) {
        inA += vA;
        inB += vB;
      }
    }
    if (k0 + lane < n_tr) {
      const unsigned eA = (unsigned)cumA + (inA - cA), eB = (unsigned)cumB + (inB - cB0 - cB1);
      segA[k0 + lane] = (unsigned long long)eA | ((unsigned long long)il << 32);
      segB[2 * (k0 + lane)] = (unsigned long long)eB | ((unsigned long long)lo << 32);
      segB[2 * (k0 + lane) + 1] = (unsigned long long)(eB + cB0) | ((unsigned long long)ih << 32);
    }
    cumA += __shfl_sync(0xffffffffu, inA, 31);
    cumB += __shfl_sync(0xffffffffu, inB, 31);
  }
  if (cumA > 0xfffffff0ull || cumB > 0xfffffff0ull) {  // positions are 32-bit
    if (lane == 0) {
      atomicOr(a.state + 3, WK_OVERFLOW);
      a.wset[s] = ws;
    }
    return;
  }
  if (lane == 0) {
    segA[n_tr] = cumA;  // sentinels: first position past the list
    segB[2 * n_tr] = cumB;
  }
  unsigned nchA = (unsigned)((cumA + CH - 1) / CH), nchB = (unsigned)((cumB + CH - 1) / CH);
  // a set without any in-window sample still needs one (empty) chunk: the warp that finishes a set's
  // last chunk writes its log-likelihood
  const bool dummy = LL && (nchA + nchB == 0u);
  if (dummy) nchA = 1u;
  unsigned long long chA = 0, chB = 0;
  if (lane == 0) {
    chA = atomicAdd(a.state + 1, (unsigned long long)nchA);
    chB = atomicAdd(a.state + 2, (unsigned long long)nchB);
  }
  chA = __shfl_sync(0xffffffffu, chA, 0);
  chB = __shfl_sync(0xffffffffu, chB, 0);
  ws.cntA = (unsigned)cumA;
  ws.cntB = (unsigned)cumB;
  ws.chA = (unsigned)chA;
  ws.chB = (unsigned)chB;
  ws.n_chunks = nchA + nchB;
  if (chA + nchA > a.cap_chunk || chB + nchB > a.cap_chunk) {
    if (lane == 0) {
      atomicOr(a.state + 3, WK_OVERFLOW);
      ws.n_chunks = 0u;
      a.wset[s] = ws;
    }
    return;
  }
  if (lane == 0) {
    a.wset[s] = ws;
    atomicAdd(a.state + 4, cumA);
    atomicAdd(a.state + 5, cumB);
  }
  __syncwarp();
  // chunk entries: lane = segment; segment k covers positions [cum_k, cum_{k+1}) and so the chunks whose
  // first position falls in that range
  uint4* entA = a.chunk + chA;
  uint4* entB = a.chunk + (a.cap_chunk - chB - nchB);
  if (dummy) {
    if (lane == 0) entA[0] = make_uint4((unsigned)s, 0u, (unsigned)seg0, 1u);  // valid = 0, 1 segment left
    return;
  }
#pragma unroll 1
  for (int list = 0; list < 2; ++list) {
    const unsigned long long* sg = list ? segB : segA;
    const unsigned nsg = list ? 2u * (unsigned)n_tr : (unsigned)n_tr;
    const unsigned cnt = list ? (unsigned)cumB : (unsigned)cumA;
    uint4* ent = list ? entB : entA;
    const unsigned segbase = (unsigned)(sg - a.seg);
    for (unsigned k = lane; k < nsg; k += 32) {
      const unsigned c0 = (unsigned)sg[k], c1 = (unsigned)sg[k + 1];
      const unsigned rem = nsg - k + 1u;  // segments from k on, sentinel included
      for (unsigned ci = (c0 + CH - 1) / CH; ci * (unsigned)CH < c1; ++ci) {
        const unsigned p0 = ci * (unsigned)CH;
        const unsigned valid = cnt - p0 < (unsigned)CH ? cnt - p0 : (unsigned)CH;
        ent[ci] = make_uint4((unsigned)s, p0, segbase + k,
                             ((unsigned)list << 31) | (valid << 8) | (rem < 255u ? rem : 255u));
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// eval
// --------------------------------------------------------------------------------------------
#ifndef JXP_WALK_N
#define JXP_WALK_N 2
#endif
#ifndef JXP_WALK_MINB
#define JXP_WALK_MINB 4
#endif

// Kepler solve for N mean anomalies of ONE orbit (e, 1 - e, 1 + e, sqrt(1 - e^2), 1 / (1 + e) warp-uniform):
// core/kepler.py:28-56, 82-108 with the re-arrangements of kepler_reduced_n<N, XY, ONE_DIV> (jxp_math.cuh);
// returns (sin f, cos f) D, D = (1 - e) c^2 + (1 + e) s^2.
template <int N>
__device__ __forceinline__ void kepler_xy_u(const double (&Mr)[N], double e, double ome, double ope,
                                            double s1me2, double inv_ope, bool slow, double (&yD)[N],
                                            double (&xD)[N]) {
  const double pi = Num<double>::pi;
  bool high[N];
  double M[N], E0[N], hs[N], hc[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    high[q] = Mr[q] > pi;
    M[q] = high[q] ? Num<double>::two_pi - Mr[q] : Mr[q];
  }
  if (!slow) {
    const float ef = (float)e, omef = (float)ome, iopef = (float)inv_ope;
#pragma unroll
    for (int q = 0; q < N; ++q) E0[q] = (double)kepler_starter_t<float>((float)M[q], ef, omef, iopef);
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) E0[q] = kepler_starter_t<double>(M[q], e, ome, inv_ope);
  }
  // E0 / 2 lies in [0, pi / 2] (to the starter's 1e-6): no range reduction
#pragma unroll
  for (int q = 0; q < N; ++q) fsincos_q1(0.5 * E0[q], &hs[q], &hc[q]);
  double f0[N], num[N], den[N], dE[N];
  if (!slow) {
    // d3 -> d4 -> dE of kepler.py:101-107 carried as fractions: one division (see kepler_reduced_n)
#pragma unroll
    for (int q = 0; q < N; ++q) {
      const double sinE = 2.0 * hs[q] * hc[q];
      const double cE = 2.0 * hs[q] * hs[q];
      const double sE = E0[q] - sinE;
      f0[q] = fma(e, sE, fma(E0[q], ome, -M[q]));
      const double f1 = fma(e, cE, ome);
      const double f2 = e * sinE;
      const double f3 = 1.0 - f1;
      const double N3 = -f0[q] * f1;
      const double D3 = fma(-0.5 * f0[q], f2, f1 * f1);
      const double D32 = D3 * D3;
      const double hf2 = 0.5 * f2, f36 = f3 * (1.0 / 6.0);
      const double D4 = fma(N3, fma(hf2, D3, N3 * f36), f1 * D32);
      const double N4 = -f0[q] * D32;
      const double D42 = D4 * D4;
      const double N42 = N4 * N4;
      den[q] = fma(D42, fma(f1, D4, hf2 * N4), N42 * fma(f36, D4, -(f2 * (1.0 / 24.0)) * N4));
      num[q] = -f0[q] * (D42 * D4);
    }
#pragma unroll
    for (int q = 0; q < N; ++q) dE[q] = fdiv(num[q], den[q]);
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) {
      const double sinE = 2.0 * hs[q] * hc[q];
      const double cE = 2.0 * hs[q] * hs[q];
      const double sE = E0[q] - sinE;
      f0[q] = fma(e, sE, fma(E0[q], ome, -M[q]));
      const double f1 = fma(e, cE, ome);
      const double f2 = e * sinE;
      const double f3 = 1.0 - f1;
      const double d3 = fdiv(-f0[q] * f1, fma(-0.5 * f0[q], f2, f1 * f1));
      const double d4 = fdiv(-f0[q], fma(d3 * d3, f3 * (1.0 / 6.0), fma(0.5 * d3, f2, f1)));
      const double d42 = d4 * d4;
      dE[q] = fdiv(-f0[q], fma(-d42 * d4, f2 * (1.0 / 24.0), fma(d42, f3 * (1.0 / 6.0), fma(0.5 * d4, f2, f1))));
    }
  }
  const double two_s = 2.0 * s1me2;
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double h = 0.5 * dE[q];
    const double h2 = h * h;
    const double sh = fma(-h * h2, 1.0 / 6.0, h);
    const double ch = fma(h2, fma(h2, 1.0 / 24.0, -0.5), 1.0);
    const double s = fma(hs[q], ch, hc[q] * sh);
    const double c = fma(hc[q], ch, -hs[q] * sh);
    const double sf = two_s * s * c;
    yD[q] = high[q] ? -sf : sf;
    xD[q] = fma(ome, c * c, -ope * (s * s));
  }
}

struct WalkEvalArgs {
  const double* rec;
  const WalkSet* wset;
  const unsigned long long* seg;
  const uint4* chunk;
  double* part;
  const unsigned long long* state;
  unsigned long long cap_chunk;
  const double* scalars;
  const double* t;
  int64_t n_times;
  const double* y;
  const double* winv;
  double* flux;
  double* flux_sum;
  double* loglike;
  WalkSet* wset_rw;
  unsigned long long* counters;
};

// Sortedness flags of the plan launch (one per likelihood-prep CTA) and the table overflow flag.  Thread 0
// of the grid leaves the verdict in counters[3] for the window-test kernels launched behind this one (they
// return at once when it is 0) and for jxp_system_list_stats.
__device__ __forceinline__ bool walk_eval_active(const WalkEvalArgs& a, int lane) {
  const bool bad = a.scalars[LL_FLAG0 + lane] != 0.0 || a.scalars[LL_FLAG0 + 32 + lane] != 0.0;
  const bool unsorted = __any_sync(0xffffffffu, bad);
  const bool over = a.state[3] != 0ull || a.state[1] + a.state[2] > a.cap_chunk;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.counters[2] = a.state[4];  // in-window samples on the inside / edge lists (jxp_system_list_stats)
    a.counters[4] = a.state[5];
    a.counters[3] = (unsorted ? 1ull : 0ull) | (over ? 2ull : 0ull);
  }
  return !unsorted && !over;
}

template <int N, bool FLUX>
__global__ void __launch_bounds__(128, JXP_WALK_MINB) k_walk_eval(const __grid_constant__ WalkEvalArgs a) {
  constexpr int CH = 32 * N;
  __shared__ double s_rec[4][WALK_REC];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (!walk_eval_active(a, lane)) return;
  const unsigned nA = (unsigned)a.state[1], nB = (unsigned)a.state[2];  // <= cap_chunk <= 2^27
  const unsigned n_chunks = nA + nB;
  const unsigned nw = gridDim.x * 4u;
  unsigned ch = blockIdx.x * 4u + wid;
  if (ch >= n_chunks) return;
  const unsigned topB = (unsigned)a.cap_chunk - nB;  // slot of the first edge-list chunk
  auto slot_of = [&](unsigned c) { return c < nA ? c : topB + (c - nA); };
  unsigned n_kep = 0, n_ld = 0, n_fast = 0;
  double* rec = s_rec[wid];

  // software pipeline: entry two chunks ahead, record + segment window one chunk ahead
  uint4 ent = a.chunk[slot_of(ch)];
  double rec_v = a.rec[(size_t)ent.x * WALK_REC + lane];
  unsigned long long seg_v = a.seg[ent.z + (unsigned)(lane < (int)(ent.w & 255u) ? lane : (int)(ent.w & 255u) - 1)];
  uint4 ent_n = ent;
  if (ch + nw < n_chunks) ent_n = a.chunk[slot_of(ch + nw)];
  for (; ch < n_chunks; ch += nw) {
    // ---- this chunk's staged inputs; start the loads of the next one
    const unsigned set = ent.x, p0 = ent.y, segk = ent.z, meta = ent.w;
    const unsigned valid = (meta >> 8) & 0xffffu, nrem = meta & 255u;
    const bool inside_list = (meta >> 31) == 0u;
    __syncwarp();
    rec[lane] = rec_v;
    const unsigned long long seg_c = seg_v;
    const bool more = ch + nw < n_chunks;
    if (more) {
      ent = ent_n;
      rec_v = a.rec[(size_t)ent.x * WALK_REC + lane];
      const int nr = (int)(ent.w & 255u);
      seg_v = a.seg[ent.z + (unsigned)(lane < nr ? lane : nr - 1)];
      if (ch + 2u * nw < n_chunks) ent_n = a.chunk[slot_of(ch + 2u * nw)];
    }
    __syncwarp();

    // ---- positions -> sample indices: largest j with cum[j] <= p among the staged 32 segments
    const unsigned cumw = (lane < (int)nrem) ? (unsigned)seg_c : 0xffffffffu;
    const unsigned startw = (unsigned)(seg_c >> 32);
    bool have[N];
    unsigned idx[N];
#pragma unroll
    for (int h = 0; h < N; ++h) {
      const unsigned p = p0 + 32u * h + lane;
      have[h] = 32u * h + lane < valid;
      int j = 0;
#pragma unroll
      for (int st = 16; st > 0; st >>= 1) {
        const unsigned v = __shfl_sync(0xffffffffu, cumw, j + st);
        if (v <= p) j += st;
      }
      unsigned cj = __shfl_sync(0xffffffffu, cumw, j);
      unsigned sj = __shfl_sync(0xffffffffu, startw, j);
      if (j == 31 && nrem > 32u && have[h]) {
        // more than 32 segments inside one chunk (runs of one or two samples): walk the table itself; the
        // list's sentinel (first position past the list, > p for a valid position) ends the search
        unsigned k = segk + 31u;
        while ((unsigned)a.seg[k + 1u] <= p) ++k;
        const unsigned long long sv = a.seg[k];
        cj = (unsigned)sv;
        sj = (unsigned)(sv >> 32);
      }
      idx[h] = have[h] ? sj + (p - cj) : 0u;
    }
    double tm[N], yv[N], wv[N];
#pragma unroll
    for (int h = 0; h < N; ++h) tm[h] = a.t[idx[h]];
    if (!FLUX) {
#pragma unroll
      for (int h = 0; h < N; ++h) {
        yv[h] = a.y[idx[h]];
        wv[h] = a.winv[idx[h]];
      }
    }
    double tot[N];
#pragma unroll
    for (int h = 0; h < N; ++h) tot[h] = 0.0;
    if (valid != 0u) {
      // ---- t -> mean anomaly (keplerian.py:538, kepler.py:31), the reference's operation order; the
      // division by the period is a multiplication by fl(1 / P) with one residual correction
      const double P = rec[WK_PERIOD], invP = rec[WK_INVP], t0 = rec[WK_T0], tref = rec[WK_TREF];
      const unsigned flags = (unsigned)__double_as_longlong(rec[WK_FLAGS]);
      const int lmax = (int)(flags >> 8);
      double Mr[N];
#pragma unroll
      for (int h = 0; h < N; ++h) {
        const double x = (tm[h] - t0) - tref;
        const double num = Num<double>::two_pi * x;
        const double q = num * invP;
        const double r = fma(-P, q, num);
        Mr[h] = mod_two_pi(fma(r, invP, q));
      }
      double yD[N], xD[N];
      if (flags & WKF_CIRCULAR) {
#pragma unroll
        for (int h = 0; h < N; ++h) fsincos(Mr[h], &yD[h], &xD[h]);  // keplerian.py:539-540
      } else {
        kepler_xy_u<N>(Mr, rec[WK_E], rec[WK_OME], rec[WK_OPE], rec[WK_S1ME2], rec[WK_INVOPE],
                       (flags & WKF_SLOW_STARTER) != 0u, yD, xD);
      }
      // ---- sky position (keplerian.py:568-576, 630-632; light_curves/limb_dark.py:53): with
      // (x0, y0) = -a (xD, yD) the constants -a / R*, cos i and -sin i are folded into k1, k2, zs
      const double cw = rec[WK_CW], sw = rec[WK_SW], k1 = rec[WK_K1], k2 = rec[WK_K2], zs = rec[WK_ZS];
      const double rr = rec[WK_RR], onepr = rec[WK_ONEPR];
      const double ar = fabs(rr);
      const bool r_small = ar < 1.0;
      double bq[N], rq[N];
      bool intr[N], fast[N], want_gen = false, want_fast = false;
#pragma unroll
      for (int h = 0; h < N; ++h) {
        const double u1 = fma(cw, xD[h], -sw * yD[h]);
        const double u2 = fma(sw, xD[h], cw * yD[h]);
        const double v1 = k1 * u1, v2 = k2 * u2;
        bq[h] = fsqrt(fma(v1, v1, v2 * v2));
        rq[h] = rr;
        intr[h] = have[h] && (zs * u2 > 0.0) && !(bq[h] >= onepr);
        n_kep += have[h] ? 1u : 0u;
        n_ld += intr[h] ? 1u : 0u;
        // the flux code an item gets depends on the item alone: inside list AND b < 1 - r
        fast[h] = inside_list && (bq[h] < 1.0 - ar) && r_small;
        n_fast += (intr[h] && fast[h]) ? 1u : 0u;
        want_gen = want_gen || (intr[h] && !fast[h]);
        want_fast = want_fast || (intr[h] && fast[h]);
      }
      want_gen = __any_sync(0xffffffffu, want_gen);
      want_fast = __any_sync(0xffffffffu, want_fast);
      const double g0 = rec[WK_G0], g1 = rec[WK_G1], g2 = rec[WK_G2];
      if (want_fast) {
        double s0[N], s1v[N], s2[N];
        ld_solution_inside_n<N>(bq, rq, lmax, s0, s1v, s2);  // constant node cosines, no atan2
#pragma unroll
        for (int h = 0; h < N; ++h)
          if (intr[h] && fast[h]) tot[h] = ld_combine(lmax, s0[h], s1v[h], s2[h], g0, g1, g2);
      }
      if (want_gen) {
        double s0[N], s1v[N], s2[N];
        ld_solution_n<N, 1>(bq, rq, lmax, s0, s1v, s2);
#pragma unroll
        for (int h = 0; h < N; ++h)
          if (intr[h] && !fast[h]) tot[h] = ld_combine(lmax, s0[h], s1v[h], s2[h], g0, g1, g2);
      }
    }
    if (FLUX) {
      // rows were zeroed before the launch; positions of different items are disjoint
#pragma unroll
      for (int h = 0; h < N; ++h) {
        if (have[h] && tot[h] != 0.0) {
          const size_t o = (size_t)set * (size_t)a.n_times + idx[h];
          if (a.flux_sum) a.flux_sum[o] = tot[h];
          if (a.flux) a.flux[o] = tot[h];
        }
      }
    } else {
      // chi^2 = sum_t ((y - 1) / sigma)^2 + sum_items [ (r0 - m / sigma)^2 - r0^2 ]
      double csum = 0.0;
#pragma unroll
      for (int h = 0; h < N; ++h) {
        double dchi = 0.0;
        if (have[h] && tot[h] != 0.0) {
          const double r0 = (yv[h] - 1.0) * wv[h];
          const double mw = tot[h] * wv[h];
          dchi = (-mw) * (2.0 * r0 - mw);
        }
        csum += dchi;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
      // the warp that finishes a set's last chunk adds the set's partials in a fixed order
      const WalkSet wsv = a.wset[set];
      unsigned prev = 0;
      if (lane == 0) {
        a.part[slot_of(ch)] = csum;
        __threadfence();
        prev = atomicAdd(&a.wset_rw[set].done, 1u);
      }
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (prev + 1u == wsv.n_chunks) {
        __threadfence();
        const unsigned nchA = wsv.n_chunks - (unsigned)((wsv.cntB + CH - 1) / CH);
        const unsigned nchB = wsv.n_chunks - nchA;
        const double* pA = a.part + wsv.chA;
        const double* pB = a.part + (a.cap_chunk - wsv.chB - nchB);
        double acc = 0.0;
        for (unsigned i = lane; i < nchA; i += 32) acc += __ldcg(pA + i);
        for (unsigned i = lane; i < nchB; i += 32) acc += __ldcg(pB + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        // set-independent terms: the LL_NCTA pairs of the likelihood prep added in order
        static_assert(LL_NCTA == 64, "two pairs per lane");
        const double2 pa = reinterpret_cast<const double2*>(a.scalars + LL_SCAL0)[lane];
        const double2 pb = reinterpret_cast<const double2*>(a.scalars + LL_SCAL0)[32 + lane];
        double chi0 = 0.0, cst = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          chi0 += __shfl_sync(0xffffffffu, pa.x, k);
          cst += __shfl_sync(0xffffffffu, pa.y, k);
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          chi0 += __shfl_sync(0xffffffffu, pb.x, k);
          cst += __shfl_sync(0xffffffffu, pb.y, k);
        }
        if (lane == 0) a.loglike[set] = -0.5 * (chi0 + acc) + cst;
      }
    }
  }
  n_kep = __reduce_add_sync(0xffffffffu, n_kep);
  n_ld = __reduce_add_sync(0xffffffffu, n_ld);
  n_fast = __reduce_add_sync(0xffffffffu, n_fast);
  if (lane == 0) {
    atomicAdd(a.counters + 0, (unsigned long long)n_kep);
    atomicAdd(a.counters + 1, (unsigned long long)n_ld);
    atomicAdd(a.counters + 5, (unsigned long long)n_fast);
  }
}

}  // namespace jxp

namespace jxph {

// Launches of the transit-walk path (double, one body per set, no exposure integration, quadratic law,
// order 10).  prep: plan launch (records, segments, chunk entries, likelihood terms / sortedness flags);
// then the evaluation.  The caller launches the window-test kernels behind it as the on-device fallback.
int launch_walk_f64(const double* bodies, int n_sets, const double* u, int n_u, int64_t u_stride,
                    const double* t, int64_t n_times, double* flux, double* flux_sum, const double* y,
                    const double* sigma, int64_t sigma_stride, double* loglike, unsigned char* ws,
                    const SysLayout& L, double widen, double margin, cudaStream_t st,
                    cudaEvent_t ev_start, cudaEvent_t ev_stop, int ev_which) {
  constexpr int N = JXP_WALK_N;
  constexpr int CH = 32 * N;
  double* d_scalars = reinterpret_cast<double*>(ws + L.off_scalars);
  WalkArgs a;
  a.bodies = bodies;
  a.u = u;
  a.n_u = n_u;
  a.u_stride = u_stride;
  a.t = t;
  a.n_times = n_times;
  a.n_sets = n_sets;
  a.y = y;
  a.sigma = sigma;
  a.sigma_stride = sigma_stride;
  a.out_b = reinterpret_cast<BodyC<double>*>(ws + L.off_bodies);
  a.out_s = reinterpret_cast<SetC<double>*>(ws + L.off_sets);
  a.counters = reinterpret_cast<unsigned long long*>(d_scalars + 2);
  a.tile_ctr = reinterpret_cast<int*>(ws + L.off_tilectr);
  a.winv = reinterpret_cast<double*>(ws + L.off_winv);
  a.scalars = d_scalars;
  a.rec = reinterpret_cast<double*>(ws + L.off_wk_rec);
  a.wset = reinterpret_cast<WalkSet*>(ws + L.off_wk_set);
  a.seg = reinterpret_cast<unsigned long long*>(ws + L.off_wk_seg);
  a.chunk = reinterpret_cast<uint4*>(ws + L.off_wk_chunk);
  a.part = reinterpret_cast<double*>(ws + L.off_wk_part);
  a.state = reinterpret_cast<unsigned long long*>(d_scalars + WALK_STATE0);
  a.cap_seg = L.wk_cap_seg;
  a.cap_chunk = L.wk_cap_chunk;
  a.n_plan_ctas = (n_sets + 7) / 8;
  a.widen = widen;
  a.margin = margin;
  a.flux = flux;
  a.flux_sum = flux_sum;
  a.loglike = loglike;
  cudaError_t e = cudaMemsetAsync(a.state, 0, 8 * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk): memset: %s", cudaGetErrorString(e));
  if (ev_start && ev_which == 1) cudaEventRecord(ev_start, st);
  if (loglike)
    k_walk_plan<CH, true><<<a.n_plan_ctas + LL_NCTA_H, 256, 0, st>>>(a);
  else
    k_walk_plan<CH, false><<<a.n_plan_ctas + LL_NCTA_H, 256, 0, st>>>(a);
  if (ev_stop && ev_which == 1) cudaEventRecord(ev_stop, st);
  int rc = check_launch("jxp_system_light_curve(walk plan)");
  if (rc != JXP_OK) return rc;

  WalkEvalArgs b;
  b.rec = a.rec;
  b.wset = a.wset;
  b.wset_rw = a.wset;
  b.seg = a.seg;
  b.chunk = a.chunk;
  b.part = a.part;
  b.state = a.state;
  b.cap_chunk = a.cap_chunk;
  b.scalars = d_scalars;
  b.t = t;
  b.n_times = n_times;
  b.y = y;
  b.winv = a.winv;
  b.flux = flux;
  b.flux_sum = flux_sum;
  b.loglike = loglike;
  b.counters = a.counters;
  static thread_local int per_sm_ll = 0, per_sm_fx = 0;
  int& per_sm = loglike ? per_sm_ll : per_sm_fx;
  if (per_sm == 0) {
    cudaError_t eo = loglike ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_walk_eval<N, false>, 128, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_walk_eval<N, true>, 128, 0);
    if (eo != cudaSuccess || per_sm < 1) {
      per_sm = 0;
      return fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk): occupancy query: %s", cudaGetErrorString(eo));
    }
  }
  // the number of chunks is only known on the device: a resident grid, chunks strided over it
  unsigned long long bound = (unsigned long long)n_sets * ((unsigned long long)n_times / CH + 3ull);
  if (bound > L.wk_cap_chunk) bound = L.wk_cap_chunk;
  const long long want = (long long)((bound + 3ull) / 4ull);
  const long long resident = (long long)per_sm * sm_count();
  const unsigned grid = (unsigned)(want < resident ? want : resident);
  if (ev_start && ev_which == 0) cudaEventRecord(ev_start, st);
  if (loglike)
    k_walk_eval<N, false><<<grid, 128, 0, st>>>(b);
  else
    k_walk_eval<N, true><<<grid, 128, 0, st>>>(b);
  if (ev_stop && ev_which == 0) cudaEventRecord(ev_stop, st);
  return check_launch("jxp_system_light_curve(walk eval)");
}

}  // namespace jxph
