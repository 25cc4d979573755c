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
  WK_T0 = 0, WK_TREF, WK_PERIOD, WK_INVP, WK_E, WK_OME, WK_OPE, WK_S1ME2, WK_CW, WK_SW, WK_K1, WK_K2,
  WK_ZS, WK_ONEPR, WK_G0, WK_G1, WK_G2, WK_RR, WK_INVOPE, WK_FLAGS,
  WK_NFIELDS
};
static_assert(WK_T0 % 2 == 0 && WK_PERIOD % 2 == 0 && WK_E % 2 == 0 && WK_OPE % 2 == 0 && WK_CW % 2 == 0 &&
                  WK_K1 % 2 == 0 && WK_ZS % 2 == 0 && WK_G0 % 2 == 0,
              "pairs read as double2");
static_assert(WK_NFIELDS <= WALK_REC, "record size");

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
                              // [4] positions on the inside lists, [5] on the edge lists, [6], [7] claim counters of
                              // k_walk_eval, [8] likelihood-prep CTAs done (all zeroed per call); scalars
                              // [WALK_STATE0 + 10], [+ 11]: the set-independent likelihood terms
  unsigned long long cap_seg, cap_chunk;
  int n_plan_ctas;
  double widen, margin;
  // several bodies per set: one plan + evaluation per body, this launch's body and the stride of the records
  int n_bodies, body;
  // exposure integration (light_curves/transforms.py:67-116): extreme offsets of the stencil (<= 0, >= 0); a
  // sample is on the list when ANY of its sub-samples can be in transit, on the inside list when ALL of them
  // are estimated to have the planet inside the disk
  double dt_lo, dt_hi;
  unsigned min_in;  // shortest run of samples worth a segment of its own on the inside list
  // outputs
  double* flux;
  double* flux_sum;
  double* loglike;
};

constexpr unsigned long long WK_OVERFLOW = 2ull, WK_UNSUPPORTED = 4ull;

// --------------------------------------------------------------------------------------------
// plan
// --------------------------------------------------------------------------------------------
template <int CH, bool LL>
__global__ void __launch_bounds__(256, 4) k_walk_plan(const __grid_constant__ WalkArgs a) {
  if ((int)blockIdx.x >= a.n_plan_ctas) {
    const int cta = (int)blockIdx.x - a.n_plan_ctas;
    if (LL) {
      loglike_prep_body<double>(cta, LL_NCTA, a.y, a.sigma, a.sigma_stride, a.n_times, a.winv, a.scalars, a.t);
      // the last of these CTAs to finish adds the LL_NCTA pairs in order: the set-independent terms
      if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.state + 8, 1ull) == (unsigned long long)(LL_NCTA - 1)) {
          __threadfence();
          double chi0 = 0.0, cst = 0.0;
          for (int k = 0; k < LL_NCTA; ++k) {
            chi0 += __ldcg(a.scalars + LL_SCAL0 + 2 * k);
            cst += __ldcg(a.scalars + LL_SCAL0 + 2 * k + 1);
          }
          a.scalars[WALK_STATE0 + 10] = chi0;
          a.scalars[WALK_STATE0 + 11] = cst;
        }
      }
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
  const int unit = s * a.n_bodies + a.body;  // (set, body) record
  if (lane == 0)
    sys_prep_body<double>(unit, a.bodies, a.n_sets, a.n_bodies, a.u, a.n_u, a.u_stride, a.out_b, a.out_s,
                          a.counters, a.tile_ctr, a.widen, a.margin);
  __syncwarp();
  const BodyC<double>& c = a.out_b[unit];
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
    const unsigned f = (unsigned)sc.lmax << 8;
    r[WK_FLAGS] = __longlong_as_double((long long)f);
    // 1 - e <= 1e-5: the float starter and the one-division correction are not good there (and such a body
    // has no provable transit window anyway): the whole call goes to the window-test kernels
    if (!(c.ome > 1e-5)) atomicOr(a.state + 3, WK_UNSUPPORTED);
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
      const double pa = ((t_first + a.dt_lo) - t0ref) * c.inv_period - phc;
      const double pb = ((t_last + a.dt_hi) - t0ref) * c.inv_period - phc;
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
  // First index i in [0, n] with !(t[i] < x) (lower bound) or !(t[i] <= x) (UPPER) on sorted t: the four
  // samples around the uniform-grid estimate are loaded at once and decide it when the boundary lies among
  // them (one round trip on a regular cadence); anything else goes to the bracketing search of jxp_list.cuh.
  auto search4 = [&](double x0, double x1, double x2, double x3, unsigned (&r)[4]) {
    const double xs[4] = {x0, x1, x2, x3};
    unsigned base[4];
    double v[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double gd = (xs[q] - t_first) * inv_dt;
      gd = fmin(fmax(gd, 0.0), (double)n);
      const unsigned g = (unsigned)gd;
      base[q] = n <= 4u ? 0u : (g < 2u ? 0u : (g - 2u > n - 4u ? n - 4u : g - 2u));
#pragma unroll
      for (int m = 0; m < 4; ++m) v[q][m] = a.t[base[q] + ((unsigned)m < n ? (unsigned)m : n - 1u)];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool upper = (q & 1) != 0;  // x0, x2: lower bounds; x1, x3: upper bounds
      const unsigned span = n < 4u ? n : 4u;
      unsigned cnt = 0;
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if ((unsigned)m < span) cnt += (upper ? (v[q][m] <= xs[q]) : (v[q][m] < xs[q])) ? 1u : 0u;
      const bool decided = (cnt > 0u || base[q] == 0u) && (cnt < span || base[q] + span == n);
      if (decided)
        r[q] = base[q] + cnt;
      else
        r[q] = upper ? tl_search<true>(a.t, n, xs[q], t_first, inv_dt) : tl_search<false>(a.t, n, xs[q], t_first, inv_dt);
    }
  };
  auto interval = [&](long long k, unsigned& lo, unsigned& hi, unsigned& ilo, unsigned& ihi) {
    lo = hi = ilo = ihi = 0;
    if (k >= n_tr) return;
    if (all) {
      hi = n;
      return;
    }
    const double base = phc + (k_first + (double)k);
    const double ta = fma(base + wlo, P, t0ref), tb = fma(base + whi, P, t0ref);
    const double tc = fma(base - w_in, P, t0ref), td = fma(base + w_in, P, t0ref);
    // exposure integration: the sample at t stands for the sub-samples t + dt_j, dt_lo <= dt_j <= dt_hi
    const double x_in0 = fmin(tc, td) - a.dt_lo, x_in1 = fmax(tc, td) - a.dt_hi;
    const bool want_in = w_in > 0.0 && x_in1 > x_in0;
    unsigned r[4];
    search4(fmin(ta, tb) - a.dt_hi, fmax(ta, tb) - a.dt_lo, x_in0, x_in1, r);
    lo = r[0];
    hi = r[1];
    if (hi < lo) hi = lo;  // unsorted time stamps: anything in range will do, the call falls back
    ilo = ihi = lo;
    if (want_in && hi - lo >= a.min_in) {
      unsigned i0 = r[2], i1 = r[3];
      i0 = i0 < lo ? lo : (i0 > hi ? hi : i0);
      i1 = i1 < i0 ? i0 : (i1 > hi ? hi : i1);
      if (i1 - i0 >= a.min_in) {
        ilo = i0;
        ihi = i1;
      }
    }
  };

  // segments (non-empty runs only): inside list at [seg0, seg0 + nsA] (sentinel last), edge list at
  // [seg0 + n_tr + 1, seg0 + n_tr + 1 + nsB]; room is taken for the most there can be
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
  // running totals, packed (positions << 8 | segments of this block): a list may exceed 2^32 only in theory
  unsigned long long cumA = 0, cumB = 0;
  unsigned nsA = 0, nsB = 0;
  for (long long k0 = 0; k0 < n_tr; k0 += 32) {
    unsigned lo, hi, il, ih;
    interval(k0 + lane, lo, hi, il, ih);
    const unsigned cA = ih - il, cB0 = il - lo, cB1 = hi - ih;
    const unsigned long long mineA = ((unsigned long long)cA << 8) | (cA ? 1u : 0u);
    const unsigned long long mineB = ((unsigned long long)(cB0 + cB1) << 8) | ((cB0 ? 1u : 0u) + (cB1 ? 1u : 0u));
    unsigned long long inA = mineA, inB = mineB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long vA = __shfl_up_sync(0xffffffffu, inA, o);
      const unsigned long long vB = __shfl_up_sync(0xffffffffu, inB, o);
      if (lane >= o) {
        inA += vA;
        inB += vB;
      }
    }
    {
      const unsigned long long exA = inA - mineA, exB = inB - mineB;
      const unsigned eA = (unsigned)(cumA + (exA >> 8)), eB = (unsigned)(cumB + (exB >> 8));
      unsigned jA = nsA + (unsigned)(exA & 255u), jB = nsB + (unsigned)(exB & 255u);
      if (cA) segA[jA] = (unsigned long long)eA | ((unsigned long long)il << 32);
      if (cB0) segB[jB++] = (unsigned long long)eB | ((unsigned long long)lo << 32);
      if (cB1) segB[jB] = (unsigned long long)(eB + cB0) | ((unsigned long long)ih << 32);
    }
    const unsigned long long tA = __shfl_sync(0xffffffffu, inA, 31), tB = __shfl_sync(0xffffffffu, inB, 31);
    cumA += tA >> 8;
    cumB += tB >> 8;
    nsA += (unsigned)(tA & 255u);
    nsB += (unsigned)(tB & 255u);
  }
  if (cumA > 0xfffffff0ull || cumB > 0xfffffff0ull) {  // positions are 32-bit
    if (lane == 0) {
      atomicOr(a.state + 3, WK_OVERFLOW);
      a.wset[s] = ws;
    }
    return;
  }
  if (lane == 0) {
    segA[nsA] = cumA;  // sentinels: first position past the list
    segB[nsB] = cumB;
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
    const unsigned nsg = list ? nsB : nsA;
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
// core/kepler.py:28-56, 82-108 with the re-arrangements of kepler_reduced_n<N, XY, ONE_DIV> (jxp_math.cuh):
// Markley starter in float (1 - e > 1e-5: the plan sends anything else to the window-test kernels), half
// angle from a degree-8 sin / cos pair without range reduction, the three nested corrections carried as
// fractions (one division), half-angle rotation by dE / 2.  Returns (sin f, cos f) D, D = (1 - e) c^2 +
// (1 + e) s^2.  A circular body (keplerian.py:539-540: f = M) is the e = 0 case: the correction then lands
// on E = M exactly (f1 = 1, f2 = f3 = 0) and the half-angle products give (sin M, cos M).
template <int N>
__device__ __forceinline__ void kepler_xy_u(const double (&Mr)[N], double e, double ome, double ope,
                                            double s1me2, double inv_ope, double (&yD)[N], double (&xD)[N]) {
  const double pi = Num<double>::pi;
  bool high[N];
  double M[N], E0[N], hs[N], hc[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    high[q] = Mr[q] > pi;
    M[q] = high[q] ? Num<double>::two_pi - Mr[q] : Mr[q];
  }
  const float ef = (float)e, omef = (float)ome, iopef = (float)inv_ope;
#pragma unroll
  for (int q = 0; q < N; ++q) E0[q] = (double)kepler_starter_t<float>((float)M[q], ef, omef, iopef);
  // E0 / 2 lies in [0, pi / 2] (to the starter's 1e-6): no range reduction
#pragma unroll
  for (int q = 0; q < N; ++q) fsincos_q1(0.5 * E0[q], &hs[q], &hc[q]);
  double num[N], den[N], dE[N];
  // d3 -> d4 -> dE of kepler.py:101-107 carried as fractions: one division (see kepler_reduced_n)
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double sinE = 2.0 * hs[q] * hc[q];
    const double cE = 2.0 * hs[q] * hs[q];
    const double sE = E0[q] - sinE;
    const double f0 = fma(e, sE, fma(E0[q], ome, -M[q]));
    const double f1 = fma(e, cE, ome);
    const double f2 = e * sinE;
    const double f3 = 1.0 - f1;
    const double N3 = -f0 * f1;
    const double D3 = fma(-0.5 * f0, f2, f1 * f1);
    const double D32 = D3 * D3;
    const double hf2 = 0.5 * f2, f36 = f3 * (1.0 / 6.0);
    const double D4 = fma(N3, fma(hf2, D3, N3 * f36), f1 * D32);
    const double N4 = -f0 * D32;
    const double D42 = D4 * D4;
    const double N42 = N4 * N4;
    den[q] = fma(D42, fma(f1, D4, hf2 * N4), N42 * fma(f36, D4, -(f2 * (1.0 / 24.0)) * N4));
    num[q] = -f0 * (D42 * D4);
  }
#pragma unroll
  // (dE is a correction of the float starter, |dE| <~ 1e-5: its reciprocal needs 1e-8, one second-order step gives 1e-12)
  for (int q = 0; q < N; ++q) dE[q] = num[q] * frcp_node(den[q]);
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
  unsigned long long* state_rw;  // [6], [7]: claim counters of the inside / edge list (zeroed per call)
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
  int n_bodies, body;  // flux outputs: this launch's column; body > 0 adds into flux_sum and into the counters
  Stencil st;          // exposure integration (EXPO instantiation): offsets and weights of the sub-samples
};

// Sortedness flags of the plan launch (one per likelihood-prep CTA) and the table overflow flag.  Thread 0
// of the grid leaves the verdict in counters[3] for the window-test kernels launched behind this one (they
// return at once when it is 0) and for jxp_system_list_stats.
__device__ __forceinline__ bool walk_eval_active(const WalkEvalArgs& a, int lane) {
  const bool bad = a.scalars[LL_FLAG0 + lane] != 0.0 || a.scalars[LL_FLAG0 + 32 + lane] != 0.0;
  const bool unsorted = __any_sync(0xffffffffu, bad);
  const bool over = a.state[3] != 0ull || a.state[1] + a.state[2] > a.cap_chunk;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long verdict = (unsorted ? 1ull : 0ull) | (over ? 2ull : 0ull);
    const unsigned long long visited = (a.state[4] + a.state[5]) * (unsigned long long)a.st.n;
    if (a.body == 0) {
      a.counters[2] = a.state[4];  // in-window samples on the inside / edge lists (jxp_system_list_stats)
      a.counters[4] = a.state[5];
      a.counters[3] = verdict;
      if (!verdict) a.counters[0] = visited;  // Kepler solves = (sub-)samples visited
    } else {  // the launches of one call run one after the other on the stream: plain read-modify-write
      a.counters[2] += a.state[4];
      a.counters[4] += a.state[5];
      a.counters[3] |= verdict;
      if (!verdict) a.counters[0] += visited;
    }
  }
  return !unsorted && !over;
}

// positions p0 + 32 h + lane of a chunk -> sample indices, from the chunk's segment window `win` (shared
// memory; entry j = segment segk + j: first position | first sample << 32, first positions strictly
// increasing, the list's sentinel = first position past the list).  The window holds 32 N segments: every
// segment is at least one position long, so no chunk of 32 N positions needs more.  Every segment that
// starts inside the chunk sets one bit of a position mask (warp-wide OR); a position's segment is the
// number of such starts at or below it.  No loop, no branch.
template <int N>
__device__ __forceinline__ void walk_indices(const uint4& ent, const unsigned long long* __restrict__ win,
                                             int lane, unsigned (&idx)[N], bool (&have)[N]) {
  const unsigned p0 = ent.y;
  const unsigned valid = (ent.w >> 8) & 0xffffu, nrem = ent.w & 255u;
  unsigned bits[N];
#pragma unroll
  for (int h = 0; h < N; ++h) bits[h] = 0u;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const unsigned j = 32u * n + lane;
    const unsigned off = (j >= 1u && j < nrem) ? (unsigned)win[j] - p0 : 0xffffffffu;  // >= 1 for real ones
#pragma unroll
    for (int h = 0; h < N; ++h) bits[h] |= (off >> 5) == (unsigned)h ? (1u << (off & 31u)) : 0u;
  }
  unsigned below = 0;
#pragma unroll
  for (int h = 0; h < N; ++h) {
    const unsigned mask = __reduce_or_sync(0xffffffffu, bits[h]);
    have[h] = 32u * h + lane < valid;
    const unsigned j = below + __popc(mask & (0xffffffffu >> (31 - lane)));
    below += __popc(mask);
    const unsigned long long sv = win[j];
    idx[h] = have[h] ? (unsigned)(sv >> 32) + (p0 + 32u * h + lane - (unsigned)sv) : 0u;
  }
}

// The set's partial sums added in a fixed order by ONE lane -> loglike[set] (eight interleaved sums over the
// inside-list chunks then the edge-list chunks, combined pairwise: the order depends on the set alone, so
// which warp or lane does it changes nothing).
template <int CH>
__device__ __forceinline__ void walk_finish_set(const WalkEvalArgs& a, unsigned set) {
  __threadfence();
  const WalkSet wsv = a.wset[set];
  const unsigned nchB = (wsv.cntB + CH - 1) / CH;
  const unsigned nchA = wsv.n_chunks - nchB;
  const double* pA = a.part + wsv.chA;
  const double* pB = a.part + ((unsigned)a.cap_chunk - wsv.chB - nchB);
  double acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.0;
  for (unsigned i = 0; i < nchA; i += 8) {
    double v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = i + m < nchA ? __ldcg(pA + i + m) : 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] += v[m];
  }
  for (unsigned i = 0; i < nchB; i += 8) {
    double v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = i + m < nchB ? __ldcg(pB + i + m) : 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] += v[m];
  }
  const double tot = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  a.loglike[set] = -0.5 * (a.scalars[WALK_STATE0 + 10] + tot) + a.scalars[WALK_STATE0 + 11];
}

// ---- asynchronous global -> shared copies (LDGSTS): staged inputs cost no registers while they fly
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
               "l"(__cvta_generic_to_global(gmem))
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
               "l"(__cvta_generic_to_global(gmem))
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Flux of core/limb_dark.py:19-84 for N points of ONE set (radius ratio r warp-uniform) with the planet
// fully inside the disk: kappa0 = pi, kappa1 = 0 and kite_area = 0 exactly, constant node cosines (see
// ld_solution_inside_n, jxp_math.cuh -- same formulas and summation order).  The caller guarantees
// 2e-6 <= r < 1 and b < 1 - r - 4e-15; then every node's 1 - z^2 = r^2 + b^2 - 2 b r c lies in
// [1.68e-3 r^2, (b + r)^2] (|c| <= 0.99916), i.e. strictly inside [10 eps, 1 - 10 eps], so the two guards of
// limb_dark.py:163-170 never fire and are not evaluated.  Everything that depends on r alone is computed
// once, and the closed-form terms come after the node loop so that little is live inside it.
template <int N>
__device__ __forceinline__ void ld_inside_u(const double (&b)[N], double r, int lmax, double g0, double g1,
                                            double g2, double (&f)[N]) {
  const double pi = Num<double>::pi;
  const double r2 = r * r;
  double s1[N];
  if (lmax >= 1) {
    // z^2 = 1 - (r^2 + b^2 - 2 b r c) as ONE fma per node, (1 - r^2 - b^2) + (2 b r) c: the same quantity to an
    // ulp of 1 (the reference rounds r^2 + b^2 - 2 b r c to the same absolute precision before subtracting).
    // The node sum  sum_i w_i (r - b c_i) t_i,  t_i = (1 - z_i^3) / (1 - z_i^2),  is accumulated as
    // r sum_i w_i t_i - b sum_i (w_i c_i) t_i: the c_i are constants here, so every node costs two fmas with
    // a constant operand instead of (r - b c_i), a product and a share of the pair sum (rounding: ~3e-16 of
    // a sum of order 1, i.e. ~1e-17 of flux).
    double two_br[N], omr2b2[N], S1[N], S2[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
      two_br[q] = b[q] * r * 2.0;
      omr2b2[q] = 1.0 - fma(b[q], b[q], r2);
      S1[q] = 0.0;
      S2[q] = 0.0;
    }
#pragma unroll
    for (int p = 0; p < 5; ++p) {
      const double wp = c_gl10[1][5 + p];
      const double cm = c_gl10[2][5 + p], cq = c_gl10[2][4 - p];  // nodes +x_p, -x_p
      const double wcm = c_gl10_wc[5 + p], wcq = c_gl10_wc[4 - p];
      double z2[2 * N], t[2 * N];
#pragma unroll
      for (int m = 0; m < 2 * N; ++m) z2[m] = fma(two_br[m >> 1], (m & 1) ? cq : cm, omr2b2[m >> 1]);
#pragma unroll
      for (int m = 0; m < 2 * N; ++m) t[m] = fnode_factor(z2[m]);  // (1 - z^3) / (1 - z^2)
#pragma unroll
      for (int q = 0; q < N; ++q) {
        S1[q] = fma(wp, t[2 * q] + t[2 * q + 1], S1[q]);
        S2[q] = fma(wcm, t[2 * q], fma(wcq, t[2 * q + 1], S2[q]));
      }
    }
    double acc[N];
#pragma unroll
    for (int q = 0; q < N; ++q) acc[q] = fma(-b[q], S2[q], r * S1[q]);
    const double k1 = (pi * 0.5) * (r * (2.0 / 3.0));
#pragma unroll
    for (int q = 0; q < N; ++q) s1[q] = fma(-k1, acc[q], pi * (2.0 / 3.0));
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) s1[q] = 0.0;
  }
  // s0s2, limb_dark.py:111-137 with kappa0 = pi, kappa1 = 0, area = 0
  const double s0l = (1.0 - r2) * pi;
  const double s0s = pi - r2 * pi;
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double bpr = b[q] + r;
    const double onembpr2 = (1.0 + bpr) * (1.0 - bpr);
    const double eta2 = r2 * fma(b[q] * b[q], 2.0, r2) * 0.5;
    const double delta = b[q] * r * 4.0;
    const bool large = onembpr2 + delta > delta;
    const double s2l = s0l * 2.0 + (eta2 - 0.5) * (4.0 * pi);
    const double s2s = s0s * 2.0 + (eta2 * pi * 2.0 - pi) * 2.0;
    f[q] = ld_combine(lmax, large ? s0l : s0s, s1[q], large ? s2l : s2s, g0, g1, g2);
  }
}

constexpr unsigned WK_RMAX = 8;   // chunks per claim (upper bound)
constexpr unsigned WK_LOG = 32;   // claims a warp remembers before it signals their completion
template <int N>
struct WalkWarpSmem {
  double rec[2][WALK_REC];            // set record of the chunk being evaluated / of the next one
  unsigned long long seg[2][32 * N];  // segment window of the next chunk / of the one after
  uint4 ent[3][WK_RMAX];              // chunk entries of the current claim, the next one, the one after
  double tb[32 * N];                  // time stamps of the next chunk
  double yb[2][32 * N];               // observed flux, 1 / sigma of this chunk / of the next one
  double wb[2][32 * N];
  unsigned ib[2][32 * N];             // sample indices (0xffffffff: padding)
  unsigned long long log[WK_LOG];     // finished claims: first slot | chunks << 32
};

// EXPO (flux outputs only): every item is evaluated at its n_sub sub-sample times t + dt_j and the weighted
// sum sum_j w_j f(t + dt_j) is formed in increasing j, the order of the reference's dot product
// (light_curves/transforms.py:113-116).
template <int N, bool FLUX, bool EXPO = false>
__global__ void __launch_bounds__(128, JXP_WALK_MINB) k_walk_eval(const __grid_constant__ WalkEvalArgs a) {
  static_assert(!EXPO || FLUX, "exposure integration on the walk path: flux outputs only");
  constexpr int CH = 32 * N;
  static_assert(CH <= 255, "segments left in a window are an 8-bit field of the chunk entry");
  __shared__ __align__(16) WalkWarpSmem<N> s_all[4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (!walk_eval_active(a, lane)) return;
  WalkWarpSmem<N>& sm = s_all[wid];
  const unsigned nA = (unsigned)a.state[1], nB = (unsigned)a.state[2];  // <= cap_chunk <= 2^27
  const unsigned nw = gridDim.x * 4u;
  const unsigned topB = (unsigned)a.cap_chunk - nB;  // slot of the first edge-list chunk
  if (nA + nB == 0u) return;

  // ---- work distribution.  A CLAIM is up to WK_RMAX consecutive chunk entries of one list (mostly one
  // set), taken from two global counters (inside list first, then the dearer edge list); the size shrinks
  // as the remaining work does (guided self-scheduling), so the warps finish within one chunk of each other.
  constexpr unsigned long long CL_VALID = 1ull << 63, CL_B = 1ull << 62;
  bool a_done = false;         // lane 0 only
  unsigned posA = 0, posB = 0;  // lane 0 only: where this warp last saw the counters
  auto claim_size = [&](unsigned remaining) {
    const unsigned r = remaining / (nw * 2u);
    return r < 1u ? 1u : (r > WK_RMAX ? WK_RMAX : r);
  };
  auto claim_issue = [&]() -> unsigned long long {  // lane 0 asks; the answer is read a claim later
    unsigned long long r = 0ull;
    if (lane == 0) {
      if (!a_done) {
        const unsigned R = claim_size((posA < nA ? nA - posA : 0u) + nB);
        const unsigned s = (unsigned)atomicAdd(a.state_rw + 6, (unsigned long long)R);
        if (s < nA) {
          const unsigned n = nA - s < R ? nA - s : R;
          r = CL_VALID | ((unsigned long long)n << 32) | s;
          posA = s + n;
        } else {
          a_done = true;
        }
      }
      if (r == 0ull) {
        const unsigned R = claim_size(posB < nB ? nB - posB : 0u);
        const unsigned s = (unsigned)atomicAdd(a.state_rw + 7, (unsigned long long)R);
        if (s < nB) {
          const unsigned n = nB - s < R ? nB - s : R;
          r = CL_VALID | CL_B | ((unsigned long long)n << 32) | s;
          posB = s + n;
        }
      }
    }
    return r;
  };
  auto claim_decode = [&](unsigned long long raw, unsigned& slot0, unsigned& cnt) -> bool {
    raw = __shfl_sync(0xffffffffu, raw, 0);
    slot0 = (unsigned)raw + ((raw & CL_B) ? topB : 0u);
    cnt = (unsigned)(raw >> 32) & 0xffffu;
    return (raw & CL_VALID) != 0ull;
  };
  auto stage_entries = [&](int buf, unsigned slot0, unsigned cnt) {
    if ((unsigned)lane < cnt) cp_async16(&sm.ent[buf][lane], a.chunk + slot0 + lane);
  };
  auto stage_rec_seg = [&](int buf, const uint4& e) {  // the chunk's set record and its segment window
    cp_async8(&sm.rec[buf][lane], a.rec + (size_t)e.x * WALK_REC + lane);
    const unsigned nr = e.w & 255u;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const unsigned j = 32u * n + lane;
      cp_async8(&sm.seg[buf][j], a.seg + e.z + (j < nr ? j : nr - 1u));
    }
  };
  auto stage_samples = [&](int buf, const uint4& e, int segbuf) {  // indices -> t, y, 1 / sigma of the chunk
    unsigned idx[N];
    bool have[N];
    walk_indices<N>(e, sm.seg[segbuf], lane, idx, have);
#pragma unroll
    for (int h = 0; h < N; ++h) {
      cp_async8(&sm.tb[32 * h + lane], a.t + idx[h]);
      if (!FLUX) {
        cp_async8(&sm.yb[buf][32 * h + lane], a.y + idx[h]);
        cp_async8(&sm.wb[buf][32 * h + lane], a.winv + idx[h]);
      }
      sm.ib[buf][32 * h + lane] = have[h] ? idx[h] : 0xffffffffu;
    }
  };

  unsigned n_ld = 0, n_fast = 0;  // per lane: limb-darkening evaluations, of which with the constant-kappa code
  unsigned nlog = 0;
  // completion of the logged claims: one fence for all their partial sums, then lane i counts the finished
  // chunks of claim i into their sets; the lane that completes a set adds the set's partials in a fixed
  // order (bit-reproducible: who does it changes nothing)
  auto flush_log = [&]() {
    __syncwarp();
    __threadfence();
    if ((unsigned)lane < nlog) {
      const unsigned long long lg = sm.log[lane];
      const unsigned slot0 = (unsigned)lg, cnt = (unsigned)(lg >> 32);
      unsigned sets[WK_RMAX + 1];
#pragma unroll
      for (unsigned k = 0; k < WK_RMAX; ++k) sets[k] = k < cnt ? a.chunk[slot0 + k].x : 0xffffffffu;
      sets[WK_RMAX] = 0xffffffffu;
      unsigned run = 0;
#pragma unroll
      for (unsigned k = 0; k < WK_RMAX; ++k) {
        if (k < cnt) {
          ++run;
          if (sets[k + 1u] != sets[k]) {  // (entries past the claim hold 0xffffffff)
            const unsigned prev = atomicAdd(&a.wset_rw[sets[k]].done, run);
            if (prev + run == a.wset[sets[k]].n_chunks) walk_finish_set<CH>(a, sets[k]);
            run = 0;
          }
        }
      }
    }
    nlog = 0;
    __syncwarp();
  };

  // ---- prologue: the first two claims, the first chunk's inputs (latencies exposed once per warp)
  unsigned cur_slot0, cur_cnt, nxt_slot0 = 0, nxt_cnt = 0, nn_slot0 = 0, nn_cnt = 0;
  if (!claim_decode(claim_issue(), cur_slot0, cur_cnt)) return;
  bool nxt_valid = claim_decode(claim_issue(), nxt_slot0, nxt_cnt);
  bool nn_valid = false;
  unsigned long long raw_nn = nxt_valid ? claim_issue() : 0ull;  // the claim after next: read a claim later
  stage_entries(0, cur_slot0, cur_cnt);
  if (nxt_valid) stage_entries(1, nxt_slot0, nxt_cnt);
  cp_async_wait_all();
  __syncwarp();
  stage_rec_seg(0, sm.ent[0][0]);
  cp_async_wait_all();
  __syncwarp();
  stage_samples(0, sm.ent[0][0], 0);
  int eb = 0, rb = 0, pb = 0;
  unsigned k = 0;
  double pend_sum = 0.0;        // the previous chunk's chi^2 terms, one per lane: reduced while this chunk's
  unsigned pend_slot = 0;       // Kepler code runs (the shuffle tree is pure latency)
  bool pend = false;

  for (;;) {
    // ================= top: this chunk's inputs have landed; the next chunk's record / segments start to fly
    cp_async_wait_all();
    __syncwarp();
    const uint4 e = sm.ent[eb][k];
    const bool same_claim = k + 1u < cur_cnt;
    const bool has_next_chunk = same_claim || nxt_valid;
    const int eb_next = eb == 2 ? 0 : eb + 1;
    // (without a next chunk this one is staged again: harmless, and the code stays free of branches)
    const uint4 en = same_claim ? sm.ent[eb][k + 1u] : (nxt_valid ? sm.ent[eb_next][0] : e);
    stage_rec_seg(rb ^ 1, en);
    const double* rec = sm.rec[rb];
    const double2* rec2 = reinterpret_cast<const double2*>(rec);
    const unsigned set = e.x;
    const unsigned valid = (e.w >> 8) & 0xffffu;
    const bool inside_list = (e.w >> 31) == 0u;
    if (!FLUX) {
      // the previous chunk's partial sum: fixed shuffle tree, one store
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pend_sum += __shfl_xor_sync(0xffffffffu, pend_sum, o);
      if (pend && lane == 0) a.part[pend_slot] = pend_sum;
    }
    // ---- t -> mean anomaly (keplerian.py:538, kepler.py:31), the reference's operation order; the
    // division by the period is a multiplication by fl(1 / P) with one residual correction
    const double2 tt = rec2[WK_T0 / 2], pp = rec2[WK_PERIOD / 2];
    const int lmax = (int)((unsigned)__double_as_longlong(rec[WK_FLAGS]) >> 8);
    double tsamp[N], wsum[N];
#pragma unroll
    for (int h = 0; h < N; ++h) {
      tsamp[h] = sm.tb[32 * h + lane];
      wsum[h] = 0.0;
    }
    const int n_sub = EXPO ? a.st.n : 1;
#pragma unroll 1
    for (int j = 0; j < n_sub; ++j) {
    double Mr[N];
#pragma unroll
    for (int h = 0; h < N; ++h) {
      const double x = ((EXPO ? tsamp[h] + a.st.dt[j] : tsamp[h]) - tt.x) - tt.y;
      const double num = Num<double>::two_pi * x;
      const double q = num * pp.y;
      const double r = fma(-pp.x, q, num);
      Mr[h] = mod_two_pi(fma(r, pp.y, q));
    }
    double yD[N], xD[N];
#if defined(JXP_WALK_SKIP) && (JXP_WALK_SKIP & 1)
    // measurement only: no Kepler solve
#pragma unroll
    for (int h = 0; h < N; ++h) {
      yD[h] = Mr[h] * 1e-3;
      xD[h] = 1.0 - Mr[h] * 1e-4;
    }
#else
    {
      const double2 ee = rec2[WK_E / 2], oo = rec2[WK_OPE / 2];
      kepler_xy_u<N>(Mr, ee.x, ee.y, oo.x, oo.y, rec[WK_INVOPE], yD, xD);
    }
#endif
    // ---- sky position (keplerian.py:568-576, 630-632; light_curves/limb_dark.py:53): with
    // (x0, y0) = -a (xD, yD) the constants -a / R*, cos i and -sin i are folded into k1, k2, zs
    const double2 ws2 = rec2[WK_CW / 2], kk = rec2[WK_K1 / 2], zo = rec2[WK_ZS / 2];
    const double rr = rec[WK_RR];
    const double ar = fabs(rr);
    const bool r_ok = (ar < 1.0) && (ar >= 2e-6);  // what ld_inside_u needs of the set ...
    const double b_lim = (1.0 - ar) - 4e-15;       // ... and of the sample
    double tot[N], bq[N];
    bool intr[N], fast[N], want_gen = false, want_fast = false;
#pragma unroll
    for (int h = 0; h < N; ++h) {
      const double u1 = fma(ws2.x, xD[h], -ws2.y * yD[h]);
      const double u2 = fma(ws2.y, xD[h], ws2.x * yD[h]);
      const double v1 = kk.x * u1, v2 = kk.y * u2;
      bq[h] = fsqrt(fma(v1, v1, v2 * v2));
      intr[h] = (32u * h + lane < valid) && (zo.x * u2 > 0.0) && !(bq[h] >= zo.y);
      // the flux code an item gets depends on the item alone: inside list AND b safely below 1 - r
      fast[h] = inside_list && (bq[h] < b_lim) && r_ok;
      n_ld += intr[h] ? 1u : 0u;
      n_fast += (intr[h] && fast[h]) ? 1u : 0u;
      want_gen = want_gen || (intr[h] && !fast[h]);
      want_fast = want_fast || (intr[h] && fast[h]);
      tot[h] = 0.0;
    }
    want_gen = __any_sync(0xffffffffu, want_gen);
    want_fast = __any_sync(0xffffffffu, want_fast);
    // ================= mid: the next chunk's segment window has landed -> its samples start to fly
    if (j == 0) {
      cp_async_wait_all();
      __syncwarp();
      stage_samples(pb ^ 1, en, rb ^ 1);
    }
    // ================= flux (core/limb_dark.py:19-196)
    const double2 gg = rec2[WK_G0 / 2];
    const double g2 = rec[WK_G2];
#if defined(JXP_WALK_SKIP) && (JXP_WALK_SKIP & 2)
    // measurement only: no flux code
#pragma unroll
    for (int h = 0; h < N; ++h) tot[h] = intr[h] ? bq[h] * 1e-3 : 0.0;
#else
    if (want_fast) {
      double f[N];
      ld_inside_u<N>(bq, ar, lmax, gg.x, gg.y, g2, f);  // constant node cosines, no atan2, no guards
#pragma unroll
      for (int h = 0; h < N; ++h)
        if (intr[h] && fast[h]) tot[h] = f[h];
    }
    if (want_gen) {
      // (the node guards of limb_dark.py:163-170 stay in: a guard-free variant with a rare-path fallback measured
      // no faster -- the compares and selects run on otherwise idle pipes)
      double rq[N], s0[N], s1v[N], s2[N];
#pragma unroll
      for (int h = 0; h < N; ++h) rq[h] = rr;
      ld_solution_n<N, 1>(bq, rq, lmax, s0, s1v, s2);
#pragma unroll
      for (int h = 0; h < N; ++h)
        if (intr[h] && !fast[h]) tot[h] = ld_combine(lmax, s0[h], s1v[h], s2[h], gg.x, gg.y, g2);
    }
#endif
    if (EXPO) {
#pragma unroll
      for (int h = 0; h < N; ++h) wsum[h] = fma(a.st.w[j], tot[h], wsum[h]);
    } else {
#pragma unroll
      for (int h = 0; h < N; ++h) wsum[h] = tot[h];
    }
    }  // sub-samples
    double tot[N];
#pragma unroll
    for (int h = 0; h < N; ++h) tot[h] = wsum[h];
    // ================= end: outputs of this chunk
    if (FLUX) {
      // rows were zeroed before the first launch of the call; positions of different items are disjoint, and
      // the launches of the bodies of one call follow one another on the stream (flux_sum: body order)
#pragma unroll
      for (int h = 0; h < N; ++h) {
        const unsigned ix = sm.ib[pb][32 * h + lane];
        if (ix != 0xffffffffu && tot[h] != 0.0) {
          const size_t o = (size_t)set * (size_t)a.n_times + ix;
          if (a.flux_sum) a.flux_sum[o] = (a.body > 0) ? a.flux_sum[o] + tot[h] : tot[h];
          if (a.flux) a.flux[o * (size_t)a.n_bodies + (size_t)a.body] = tot[h];
        }
      }
    } else {
      // chi^2 = sum_t ((y - 1) / sigma)^2 + sum_items [ (r0 - m / sigma)^2 - r0^2 ]
      double csum = 0.0;
#pragma unroll
      for (int h = 0; h < N; ++h) {
        const double wv = sm.wb[pb][32 * h + lane];
        const double r0 = (sm.yb[pb][32 * h + lane] - 1.0) * wv;
        const double mw = tot[h] * wv;
        csum += (tot[h] != 0.0) ? (-mw) * (2.0 * r0 - mw) : 0.0;
      }
      pend_sum = csum;
      pend_slot = cur_slot0 + k;
      pend = true;
    }
    // ================= advance; claims are looked up two ahead (the atomic's answer is read a claim later)
    rb ^= 1;
    pb ^= 1;
    if (k == 0u && nxt_valid) {
      nn_valid = claim_decode(raw_nn, nn_slot0, nn_cnt);
      if (nn_valid) {
        stage_entries(eb_next == 2 ? 0 : eb_next + 1, nn_slot0, nn_cnt);
        raw_nn = claim_issue();
      }
    }
    if (same_claim) {
      ++k;
      continue;
    }
    if (!FLUX) {
      if (lane == 0) sm.log[nlog] = (unsigned long long)cur_slot0 | ((unsigned long long)cur_cnt << 32);
      if (++nlog == WK_LOG || !nxt_valid) {
        // the last chunk's partial sum has to be in memory before its claim is reported
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pend_sum += __shfl_xor_sync(0xffffffffu, pend_sum, o);
        if (pend && lane == 0) a.part[pend_slot] = pend_sum;
        pend = false;
        pend_sum = 0.0;
        flush_log();
      }
    }
    if (!nxt_valid) break;
    cur_slot0 = nxt_slot0;
    cur_cnt = nxt_cnt;
    nxt_slot0 = nn_slot0;
    nxt_cnt = nn_cnt;
    nxt_valid = nn_valid;
    nn_valid = false;
    eb = eb_next;
    k = 0;
  }
  n_ld = __reduce_add_sync(0xffffffffu, n_ld);
  n_fast = __reduce_add_sync(0xffffffffu, n_fast);
  if (lane == 0) {
    atomicAdd(a.counters + 1, (unsigned long long)n_ld);
    atomicAdd(a.counters + 5, (unsigned long long)n_fast);
  }
}

}  // namespace jxp

namespace jxph {

// Launches of the transit-walk path (double, quadratic law, order 10).  Per body of the system: a plan launch
// (records, segments, chunk entries, likelihood terms / sortedness flags), then the evaluation; the bodies of
// one call follow one another on the stream and share the tables.  The caller launches the window-test
// kernels behind it as the on-device fallback.
int launch_walk_f64(const double* bodies, int n_sets, int n_bodies, const double* u, int n_u, int64_t u_stride,
                    const double* t, int64_t n_times, const Stencil& stencil, double* flux, double* flux_sum,
                    const double* y, const double* sigma, int64_t sigma_stride, double* loglike,
                    unsigned char* ws, const SysLayout& L, double widen, double margin, cudaStream_t st,
                    cudaEvent_t ev_start, cudaEvent_t ev_stop, int ev_which) {
  constexpr int N = JXP_WALK_N;
  constexpr int CH = 32 * N;
  const bool expo = stencil.n > 1 || stencil.dt[0] != 0.0 || stencil.w[0] != 1.0;
  if (loglike && (expo || n_bodies != 1))
    return fail(JXP_ERR_UNSUPPORTED, "jxp_system_light_curve(walk): log-likelihood of one body without exposure integration only");
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
  a.n_bodies = n_bodies;
  a.dt_lo = expo ? stencil.dt_min : 0.0;
  a.dt_hi = expo ? stencil.dt_max : 0.0;
  // with exposure integration a sample stands for n_sub evaluations: even a single all-inside sample pays
  a.min_in = expo ? (stencil.n >= 8 ? 1u : 2u) : 8u;
  a.flux = flux;
  a.flux_sum = flux_sum;
  a.loglike = loglike;

  WalkEvalArgs b;
  b.rec = a.rec;
  b.wset = a.wset;
  b.wset_rw = a.wset;
  b.seg = a.seg;
  b.chunk = a.chunk;
  b.part = a.part;
  b.state = a.state;
  b.state_rw = a.state;
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
  b.n_bodies = n_bodies;
  b.st = stencil;
  if (!expo) {
    b.st.n = 1;
    b.st.dt[0] = 0.0;
    b.st.w[0] = 1.0;
  }
  static thread_local int per_sm_ll = 0, per_sm_fx = 0, per_sm_ex = 0;
  int& per_sm = loglike ? per_sm_ll : (expo ? per_sm_ex : per_sm_fx);
  if (per_sm == 0) {
    cudaError_t eo = loglike ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_walk_eval<N, false>, 128, 0)
                     : expo  ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_walk_eval<N, true, true>, 128, 0)
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

  for (int body = 0; body < n_bodies; ++body) {
    a.body = body;
    b.body = body;
    cudaError_t e = cudaMemsetAsync(a.state, 0, 10 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk): memset: %s", cudaGetErrorString(e));
    if (ev_start && ev_which == 1 && body == 0) cudaEventRecord(ev_start, st);
    if (loglike)
      k_walk_plan<CH, true><<<a.n_plan_ctas + LL_NCTA_H, 256, 0, st>>>(a);
    else
      k_walk_plan<CH, false><<<a.n_plan_ctas + LL_NCTA_H, 256, 0, st>>>(a);
    if (ev_stop && ev_which == 1 && body == n_bodies - 1) cudaEventRecord(ev_stop, st);
    int rc = check_launch("jxp_system_light_curve(walk plan)");
    if (rc != JXP_OK) return rc;
    if (ev_start && ev_which == 0 && body == 0) cudaEventRecord(ev_start, st);
    if (loglike)
      k_walk_eval<N, false><<<grid, 128, 0, st>>>(b);
    else if (expo)
      k_walk_eval<N, true, true><<<grid, 128, 0, st>>>(b);
    else
      k_walk_eval<N, true><<<grid, 128, 0, st>>>(b);
    if (ev_stop && ev_which == 0 && body == n_bodies - 1) cudaEventRecord(ev_stop, st);
    rc = check_launch("jxp_system_light_curve(walk eval)");
    if (rc != JXP_OK) return rc;
  }
  return JXP_OK;
}

}  // namespace jxph
