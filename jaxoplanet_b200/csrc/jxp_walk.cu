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
//   per-set function tables (jxp_tables.cuh): with many samples per set the in-window samples are not evaluated
//                with the reference formula at all.  k_walk_plan also fits b^2 as a series in the mean anomaly
//                across the set's transit window (lanes = 32 Chebyshev nodes, node values from the direct code);
//                k_walk_tables (side stream, beside the plan) fits the flux F(b) of the set's (r, u) on pieces
//                graded towards the contact points.  A set with tables whose transits span >= 12 samples becomes
//                ITEMS (one per transit; groups of WK_GROUP per set) for
//   k_walk_items one warp walks a group's transits as one coalesced sample stream: mean anomaly -> b^2 (Horner) ->
//                piece of the F table -> flux (Horner) -> chi^2 term; two partial sums per group, added in order
//                by the lane that completes the set.  Samples the tables do not answer go on a list for
//   k_walk_fix   the direct quadrature on that list; chi^2 terms added in fixed point (order-independent).
//                Sets with tables that stay on the chunk lists (exposure stencils, short transits) read the same
//                tables in k_walk_eval; sets without tables (large r, wide windows) take the direct code there.
// Unsorted / NaN time stamps or a table that does not fit are detected on the device; the kernels
// then return at once and the window-test kernels launched behind them do the work.
#include "../../include/jaxoplanet_b200.h"

#include <cuda_runtime.h>

#include "jxp_host.h"
#include "jxp_list.cuh"
#include "jxp_orbit.cuh"
#include "jxp_prep.cuh"
#include "jxp_tables.cuh"

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
  double* tab;      // [n_sets][TB_DOUBLES] function tables (jxp_tables.cuh); nullptr: not used
  uint4* items;     // (set, first sample, samples, -) per slot of a transit of a set with tables; nullptr: no item sets
  unsigned long long cap_items;
  long long* fixsum;  // [n_sets][2] (zeroed for the item sets here)
  double2* aw;        // [n_times] (2 (y - 1) / sigma^2, 1 / sigma^2) for k_walk_items, or nullptr
  // outputs
  double* flux;
  double* flux_sum;
  double* loglike;
};

constexpr unsigned long long WK_OVERFLOW = 2ull, WK_UNSUPPORTED = 4ull;
constexpr unsigned WK_ITEM = 128;  // samples per item slot aimed at
constexpr unsigned WK_GROUP = 8;   // items per group: the unit k_walk_items claims and walks as one sample stream

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

// --------------------------------------------------------------------------------------------
// plan
// --------------------------------------------------------------------------------------------
// MINB: CTAs per SM the register budget is cut for -- 4 (64 registers, some spilled) when it takes that many to have
// every set's warp resident at once, 2 (128 registers) for smaller batches: the launch is a latency chain per set
template <int CH, bool LL, int MINB>
__global__ void __launch_bounds__(256, MINB) k_walk_plan(const __grid_constant__ WalkArgs a) {
  if ((int)blockIdx.x >= a.n_plan_ctas) {
    const int cta = (int)blockIdx.x - a.n_plan_ctas;
    if (LL) {
      loglike_prep_body<double>(cta, LL_NCTA, a.y, a.sigma, a.sigma_stride, a.n_times, a.winv, a.scalars, a.t, a.aw);
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
  // the records of the CTA's sets: one lane per set of the first warp (the derivation is a ~1500-instruction
  // serial chain; as lane 0 of every set's own warp it occupied eight warps' issue slots for one lane's work each)
  if (threadIdx.x < (blockDim.x >> 5)) {
    const int s2 = blockIdx.x * (blockDim.x >> 5) + (int)threadIdx.x;
    if (s2 < a.n_sets)
      sys_prep_body<double>(s2 * a.n_bodies + a.body, a.bodies, a.n_sets, a.n_bodies, a.u, a.n_u, a.u_stride, a.out_b,
                            a.out_s, a.counters, a.tile_ctr, a.widen, a.margin);
    __threadfence_block();
  }
  __syncthreads();
  if (s >= a.n_sets) return;
  const int unit = s * a.n_bodies + a.body;  // (set, body) record
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
  // ---- b^2 series of the set (jxp_tables.cuh): lanes = the 32 Chebyshev nodes of the transit window (widened by
  // the exposure stencil); node values from the same Kepler / sky-position code the direct path uses
  bool tab_set = false;  // this set is evaluated from its function tables by k_walk_items (whole transits) ...
  bool tab_ok = false;   // ... or, from the chunk lists, by k_walk_eval (exposure stencils, short transits)
  if (a.tab != nullptr) {
    double* T = a.tab + (size_t)s * TB_DOUBLES;
    const double dlo = wlo + a.dt_lo * c.inv_period, dhi = whi + a.dt_hi * c.inv_period;
    const double half = 0.5 * (dhi - dlo), mid = 0.5 * (dhi + dlo);
    bool okb = !all && n_tr > 0 && c.ome > 1e-5 && c.period > 0.0 && half > 0.0 && half <= 0.1;
    const double Mc = mod_two_pi(Num<double>::two_pi * phc);
    double f = 0.0;
    {
      const double d = fma(half, d_tb_node32[lane], mid);
      double Mn[1] = {mod_two_pi(fma(Num<double>::two_pi, d, Mc))}, yD[1], xD[1];
      kepler_xy_u<1>(Mn, c.e, c.ome, c.ope, c.s1me2, c.inv_ope, yD, xD);
      const double u1 = fma(c.cw, xD[0], -c.sw * yD[0]);
      const double u2 = fma(c.sw, xD[0], c.cw * yD[0]);
      const double k1 = c.na * c.inv_rstar, k2 = c.ci * c.na * c.inv_rstar, zs = -c.si * c.na;
      const double v1 = k1 * u1, v2 = k2 * u2;
      f = fma(v1, v1, v2 * v2);
      okb = okb && (zs * u2 > 0.0) && isfinite(f);  // in front of the star over the whole window
    }
    okb = __all_sync(0xffffffffu, okb);
    // node values -> Chebyshev coefficients (lane k sums over the nodes) -> degree: the last coefficient above the
    // rounding noise of the node values, at most 23 (eight coefficients of evidence behind it)
    double ck = 0.0, fm = fabs(f);
    for (int i = 0; i < 32; ++i) ck = fma(d_tb_dctT32[i][lane], __shfl_sync(0xffffffffu, f, i), ck);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fm = fmax(fm, __shfl_xor_sync(0xffffffffu, fm, o));
    // (node values carry a few 1e-15 of rounding noise, a / R* times that of the Kepler solve: the degree is where
    // three coefficients in a row are first below that level -- isolated noise above it further up is dropped
    // with the rest -- provided nothing beyond stands clearly above the noise, i.e. the series has converged)
    const unsigned big = __ballot_sync(0xffffffffu, fabs(ck) > 3e-15 * fm);
    const unsigned loud = __ballot_sync(0xffffffffu, fabs(ck) > 2e-14 * fm);
    const unsigned quiet3 = ~big & (~big >> 1) & (~big >> 2) & 0x3fffffffu;  // k, k + 1, k + 2 all small
    const int DB = quiet3 ? __ffs(quiet3) - 2 : 31;                        // (first such k) - 1; >= -1
    okb = okb && DB >= 1 && DB <= 23 && (loud >> (DB + 1)) == 0u;
    // ... -> monomial coefficients (lane j; T_k has the parity of k)
    double mj = 0.0;
    for (int k = 31; k >= 0; --k) {
      const double c_k = __shfl_sync(0xffffffffu, ck, k);
      if (k <= DB && k >= lane && ((k - lane) & 1) == 0) mj = fma(d_tb_c2mT[k][lane], c_k, mj);
    }
    if (lane < TB_B2N) T[TB_B2 + lane] = (okb && lane <= DB) ? mj : 0.0;
    // routing: tables pay when a transit spans enough samples to fill a warp's round and every evaluation is a
    // sample (no exposure stencil: those stay on the chunk lists, whose positions pack short transits densely)
    {
      const double wsamp = (whi - wlo) * P * inv_dt;  // samples per window at the mean cadence
      // (the same test k_walk_tables makes -- that launch runs beside this one; a set whose flux fits do not
      // converge there, r near the end of the validated range, gets all-invalid pieces and its samples take the
      // direct code through the fix list)
      const SetC<double>& sc = a.out_s[s];
      const double ar = fabs((double)c.rr);
      const bool fin = ar >= TB_R_MIN && ar <= TB_R_MAX && isfinite(sc.g0) && isfinite(sc.g1) && isfinite(sc.g2);
      tab_ok = okb && fin;
      tab_set = tab_ok && a.items != nullptr && wsamp >= 12.0 && n_tr < (1ll << 22);
    }
    if (lane == 0) {
      double* H = T + TB_HDR;
      unsigned long long* ts = reinterpret_cast<unsigned long long*>(a.scalars + WALK_TSTAT0);
      if (tab_ok) atomicAdd(ts + 2, 1ull);
      else if (n_tr > 0 && !all) atomicAdd(ts + 3, 1ull);  // transiting sets without tables
      H[H_MC] = Mc;
      H[H_ZSC] = 1.0 / (Num<double>::two_pi * half);
      H[H_ZOFF] = -mid / half;
      H[H_BINFO] = __longlong_as_double((long long)((tab_ok ? TB_VALID : 0u) | (unsigned)(DB | 1)));
      H[12] = H[13] = H[14] = H[15] = 0.0;
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

  // ---- a set with tables: its transits become ITEMS (set, first sample, samples) for k_walk_items -- m slots per
  // transit (m from the window length at the mean cadence, so that a slot holds <= ~WK_ITEM samples), contiguous
  // per set and in time order: the set's partial sums are added in that order
  if (tab_set) {
    const double wsamp = (whi - wlo) * P * inv_dt;
    const unsigned m = 1u + (unsigned)(fmin(wsamp, 1e6) / (double)WK_ITEM);
    const unsigned long long n_it = (unsigned long long)n_tr * m;
    // (items are handed out in GROUPS of WK_GROUP; a set's slice starts on a group boundary and is padded with empty
    // items, so a group belongs to one set and holds the same items whatever else is in the call)
    const unsigned long long n_pad = (n_it + WK_GROUP - 1ull) / WK_GROUP * WK_GROUP;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.state + 12, n_pad);
    base = __shfl_sync(0xffffffffu, base, 0);
    WalkSet wi;
    wi.seg_off = 0u;
    wi.n_tr = (unsigned)n_tr;
    wi.cntA = (unsigned)(2ull * (n_pad / WK_GROUP));  // half-groups of the set (one partial sum each) ...
    wi.chA = (unsigned)(2ull * (base / WK_GROUP));    // ... from this one on
    wi.cntB = 1u;                            // (marks an item set)
    wi.chB = wi.n_chunks = wi.done = 0u;
    if (base + n_pad > a.cap_items || n_pad > 0xfffffff0ull) {
      if (lane == 0) {
        atomicOr(a.state + 3, WK_OVERFLOW);
        wi.cntA = 0u;
        a.wset[s] = wi;
      }
      return;
    }
    for (unsigned long long k = n_it + lane; k < n_pad; k += 32) a.items[base + k] = make_uint4((unsigned)s, 0u, 0u, 0u);
    unsigned long long npos = 0;
    for (long long k0 = 0; k0 < n_tr; k0 += 32) {
      unsigned lo, hi, il, ih;
      interval(k0 + lane, lo, hi, il, ih);
      if (k0 + lane < n_tr) {
        const unsigned cnt = hi - lo, per = (cnt + m - 1u) / m;
        uint4* it = a.items + base + (unsigned long long)(k0 + lane) * m;
        for (unsigned j = 0; j < m; ++j) {
          const unsigned s0 = j * per < cnt ? j * per : cnt;
          const unsigned cj = cnt - s0 < per ? cnt - s0 : per;
          it[j] = make_uint4((unsigned)s, lo + s0, cj, 0u);
        }
        npos += cnt;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) npos += __shfl_xor_sync(0xffffffffu, npos, o);
    if (lane == 0) {
      a.wset[s] = wi;
      a.fixsum[2 * (size_t)s] = 0ll;
      a.fixsum[2 * (size_t)s + 1] = 0ll;
      atomicAdd(a.state + 5, npos);
    }
    return;
  }

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
  const double* tab;   // function tables (jxp_tables.cuh) or nullptr: chunks of sets that have them take two Horner
                       // chains per (sub-)sample instead of the Kepler solve and the quadrature
  unsigned long long* tstat;  // [0]: evaluations answered by a table
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
  double tabs[2][TB_DOUBLES];         // function tables of two sets (a set's tables travel when the set changes)
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
  unsigned tset0 = 0xffffffffu, tset1 = 0xffffffffu;  // sets whose tables sm.tabs[0], [1] hold
  int tcur = 0, tnext = 0;                            // table buffer of this chunk / of the next one
  auto stage_entries = [&](int buf, unsigned slot0, unsigned cnt) {
    if ((unsigned)lane < cnt) cp_async16(&sm.ent[buf][lane], a.chunk + slot0 + lane);
  };
  auto stage_rec_seg = [&](int buf, const uint4& e) {  // the chunk's set record and its segment window
    cp_async8(&sm.rec[buf][lane], a.rec + (size_t)e.x * WALK_REC + lane);
    if (a.tab != nullptr) {
      tnext = e.x == (tcur ? tset1 : tset0) ? tcur : tcur ^ 1;
      if ((tnext ? tset1 : tset0) != e.x) {
        const double2* T = reinterpret_cast<const double2*>(a.tab + (size_t)e.x * TB_DOUBLES);
        double2* dst = reinterpret_cast<double2*>(sm.tabs[tnext]);
#pragma unroll
        for (int i = 0; i < (TB_DOUBLES / 2 + 31) / 32; ++i)
          if (32 * i + lane < TB_DOUBLES / 2) cp_async16(dst + 32 * i + lane, T + 32 * i + lane);
        if (tnext) tset1 = e.x; else tset0 = e.x;
      }
    }
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
  unsigned n_tab = 0;             // ... and of which answered by the set's tables
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
  tcur = tnext;
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
    unsigned finfo = 0u, binfo = 0u;
    if (a.tab != nullptr) {
      finfo = (unsigned)__double_as_longlong(sm.tabs[tcur][TB_HDR + H_FINFO]);
      binfo = (unsigned)__double_as_longlong(sm.tabs[tcur][TB_HDR + H_BINFO]);
    }
    const bool use_tab = ((finfo & binfo) & TB_VALID) != 0u;
    double tsamp[N], wsum[N];
#pragma unroll
    for (int h = 0; h < N; ++h) {
      tsamp[h] = sm.tb[32 * h + lane];
      wsum[h] = 0.0;
    }
    const int n_sub = EXPO ? a.st.n : 1;
#pragma unroll 1
    for (int j = 0; j < n_sub; ++j) {
    double Mr[N];  // the mean anomaly; reduced to [0, 2 pi) below where the Kepler solve wants it
#pragma unroll
    for (int h = 0; h < N; ++h) {
      const double x = ((EXPO ? tsamp[h] + a.st.dt[j] : tsamp[h]) - tt.x) - tt.y;
      const double num = Num<double>::two_pi * x;
      const double q = num * pp.y;
      const double r = fma(-pp.x, q, num);
      Mr[h] = fma(r, pp.y, q);
    }
    double tot[N];
    bool intr[N];
    if (use_tab) {
      // ================= the set has function tables (jxp_tables.cuh): b^2 from its series in the mean anomaly, the
      // flux from the piece of its F(b) table the (sub-)sample falls on; a piece that is missing or did not converge
      // (NaN) sends the sample to the direct flux code below
      const double* th = sm.tabs[tcur];
      const double* H = th + TB_HDR;
      const double xs = H[H_XS], xe = H[H_XE], mc = H[H_MC], zsc = H[H_ZSC], zoff = H[H_ZOFF];
      const int j0 = (int)(finfo & 0xffu), db = (int)(binfo & 0xffu);
      double zv[N], b2[N];
#pragma unroll
      for (int h = 0; h < N; ++h) {
        // (reduced next to the conjunction's mean anomaly instead of into [0, 2 pi): see k_walk_items)
        const double kk = rint((Mr[h] - mc) * Num<double>::inv_two_pi);
        const double d = fma(-kk, Num<double>::two_pi, Mr[h]) - mc;
        zv[h] = fma(d, zsc, zoff);
        b2[h] = 0.0;
      }
      auto b2_step = [&](int k) {
        const double2 c1 = *reinterpret_cast<const double2*>(th + TB_B2 + k - 1);  // (m_{k-1}, m_k)
        const double2 c0 = *reinterpret_cast<const double2*>(th + TB_B2 + k - 3);  // (m_{k-3}, m_{k-2})
#pragma unroll
        for (int h = 0; h < N; ++h) {
          b2[h] = fma(fma(b2[h], zv[h], c1.y), zv[h], c1.x);
          b2[h] = fma(fma(b2[h], zv[h], c0.y), zv[h], c0.x);
        }
      };
      if (db > 11) {
#pragma unroll 1
        for (int k = db | 3; k > 11; k -= 4) b2_step(k);
      }
      b2_step(11);
      b2_step(7);
      b2_step(3);
      bool ins[N], want_edge = false;
      double q[N];
      int pc[N];
#pragma unroll
      for (int h = 0; h < N; ++h) {
        intr[h] = (32u * h + lane < valid) && (b2[h] < xe);
        q[h] = xs - b2[h];
        ins[h] = q[h] > 0.0;
        const int jb = tb_binade(q[h]);
        pc[h] = jb < j0 ? j0 : (jb >= TB_NB ? TB_P_BAD : jb);
        want_edge = want_edge || (intr[h] && !ins[h]);
        n_ld += intr[h] ? 1u : 0u;
      }
      if (__any_sync(0xffffffffu, want_edge)) {
        const double ti1 = H[H_TI1], ti2 = H[H_TI2], ti3 = H[H_TI3], te1 = H[H_TE1];
#pragma unroll
        for (int h = 0; h < N; ++h) {
          const bool ing = b2[h] < 1.0;
          const double w = fsqrt_pos(fmax(ing ? -q[h] : xe - b2[h], 1e-300));
          const int pe = ing ? TB_P_ING + (w >= ti1 ? 1 : 0) + (w >= ti2 ? 1 : 0) + (w >= ti3 ? 1 : 0)
                             : TB_P_EGR + (w >= te1 ? 1 : 0);
          pc[h] = ins[h] ? pc[h] : pe;
          q[h] = ins[h] ? q[h] : w;
        }
      }
      const double2* P[N];
      double v[N], acc[N];
#pragma unroll
      for (int h = 0; h < N; ++h) {
        P[h] = reinterpret_cast<const double2*>(th + TB_PIECE0 + (intr[h] ? pc[h] : TB_P_ZERO) * TB_PSTRIDE);
        const double2 so = *P[h];
        v[h] = fma(q[h], so.x, so.y);
        acc[h] = 0.0;
      }
#pragma unroll
      for (int k = TB_NC / 2; k >= 1; --k) {
#pragma unroll
        for (int h = 0; h < N; ++h) {
          const double2 cc = P[h][k];  // (m_{2k-2}, m_{2k-1})
          acc[h] = fma(fma(acc[h], v[h], cc.y), v[h], cc.x);
        }
      }
      bool need[N], need_any = false;
#pragma unroll
      for (int h = 0; h < N; ++h) {
        need[h] = !(acc[h] == acc[h]);
        need_any = need_any || need[h];
        tot[h] = need[h] ? 0.0 : acc[h];
        n_tab += (intr[h] && !need[h]) ? 1u : 0u;
      }
      // ================= mid: the next chunk's segment window has landed -> its samples start to fly
      if (j == 0) {
        cp_async_wait_all();
        __syncwarp();
        stage_samples(pb ^ 1, en, rb ^ 1);
      }
      if (__any_sync(0xffffffffu, need_any)) {
        const double2 gg = rec2[WK_G0 / 2];
        const double g2 = rec[WK_G2];
        double bq[N], rq[N], s0[N], s1v[N], s2[N];
#pragma unroll
        for (int h = 0; h < N; ++h) {
          bq[h] = fsqrt(fmax(b2[h], 0.0));
          rq[h] = rec[WK_RR];
        }
        ld_solution_n<N, 1>(bq, rq, lmax, s0, s1v, s2);
#pragma unroll
        for (int h = 0; h < N; ++h)
          if (need[h]) tot[h] = ld_combine(lmax, s0[h], s1v[h], s2[h], gg.x, gg.y, g2);
      }
    } else {
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
#pragma unroll
        for (int h = 0; h < N; ++h) Mr[h] = mod_two_pi(Mr[h]);
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
      double bq[N];
      bool fast[N], want_gen = false, want_fast = false;
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
    }
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
    tcur = tnext;
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
  n_tab = __reduce_add_sync(0xffffffffu, n_tab);
  if (lane == 0) {
    atomicAdd(a.counters + 1, (unsigned long long)n_ld);
    atomicAdd(a.counters + 5, (unsigned long long)n_fast);
    if (n_tab) atomicAdd(a.tstat, (unsigned long long)n_tab);
  }
}


// --------------------------------------------------------------------------------------------
// items: the sets with function tables (jxp_tables.cuh)
// --------------------------------------------------------------------------------------------
struct WalkFix {  // a sample of a table set that its tables do not answer
  unsigned set, idx;
  double b2;
};
static_assert(sizeof(WalkFix) == 16, "WalkFix");
constexpr double WK_FIX_SCALE = 17592186044416.0;  // 2^44: the fractions of the chi^2 terms of fixed samples are summed in fixed point

struct WalkItemArgs {
  const double* rec;
  const double* tab;
  const uint4* items;
  double* part;               // per-item partial sums
  const WalkSet* wset;
  WalkSet* wset_rw;
  const unsigned long long* state;
  unsigned long long* state_rw;  // [13]: item claim counter, [14]: fix entries, [15]: CTAs of k_walk_fix done
  unsigned long long cap_chunk, cap_fix;
  WalkFix* fix;
  long long* fixsum;          // [n_sets][2] fixed-point sums (integer parts, fractions) of the chi^2 terms of a set's fixed samples
  const double* scalars;
  const double* t;
  int64_t n_times;
  const double* y;
  const double* winv;
  const double2* aw;
  double* flux;
  double* flux_sum;
  double* loglike;
  unsigned long long* counters;
  unsigned long long* tstat;
  int n_sets, n_bodies, body;
};

// the verdict of walk_eval_active, read only (k_walk_eval of the same call writes it down)
__device__ __forceinline__ bool walk_items_active(const WalkItemArgs& a, int lane) {
  const bool bad = a.scalars[LL_FLAG0 + lane] != 0.0 || a.scalars[LL_FLAG0 + 32 + lane] != 0.0;
  const bool unsorted = __any_sync(0xffffffffu, bad);
  const bool over = a.state[3] != 0ull || a.state[1] + a.state[2] > a.cap_chunk;
  return !unsorted && !over;
}

// The partial sums of an item set added in group order by ONE lane -> loglike[set] (eight interleaved sums,
// combined pairwise: the order depends on the set alone).
__device__ __forceinline__ void walk_finish_items(const WalkItemArgs& a, unsigned set) {
  __threadfence();
  const WalkSet wsv = a.wset[set];
  const double* p = a.part + wsv.chA;
  const unsigned n = wsv.cntA;
  double acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.0;
  for (unsigned i = 0; i < n; i += 8) {
    double v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = i + m < n ? __ldcg(p + i + m) : 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] += v[m];
  }
  const double tot = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  a.loglike[set] = -0.5 * (a.scalars[WALK_STATE0 + 10] + tot) + a.scalars[WALK_STATE0 + 11];
}

struct WalkItemSmem {
  double tabs[TB_DOUBLES];
  double rec[WALK_REC];
  double nb_b[32];     // samples of the current half-group next to the contact point: separation ...
  unsigned nb_ix[32];  // ... and sample index (evaluated together behind the half-group's rounds)
};

// The inside-the-disk flux code for ONE sample: the rare path of k_walk_items for the samples right next to the
// contact point b = 1 - r (dx below the last binade of the F table: 2e-4 of the samples).  They are parked in
// shared memory during the rounds of a half-group and evaluated together behind them, one per lane, through this
// call (not inline: the walk keeps its registers) -- instead of going through a list in HBM and a kernel of its
// own behind the walk (k_walk_fix: 10-17 us at the end of every step, for 2 714 samples on config 3).
__device__ __noinline__ double walk_inside_one(double b, double r, int lmax, double g0, double g1, double g2) {
  double bb[1] = {b}, f[1];
  ld_inside_u<1>(bb, r, lmax, g0, g1, g2, f);
  return f[0];
}
// Measured on config 3 (step, ms; tools/gpu_ab.sh): 3 samples per lane at 4 / 5 / 6 CTAs per SM (128 / 96 / 80
// registers) 0.2727 / 0.2748 / 0.2818 (2 samples per lane at 6), 4 samples per lane 0.281 (4 CTAs) / 0.292 (5).
#ifndef JXP_ITEMS_MINB
#define JXP_ITEMS_MINB 4
#endif
#ifndef JXP_ITEMS_NN
#define JXP_ITEMS_NN 3  // samples per lane and full round
#endif

// One warp walks whole transits: an item is a run of consecutive samples of one transit of one set, so t, y and
// 1 / sigma are read coalesced, 64 samples per round (two per lane; 32 in an item's last round when that is all
// it needs), and everything that depends on the set sits in shared memory (loaded when the set changes: a claim is
// WI_CLAIM consecutive items, mostly of one set).  Per sample: mean anomaly in the reference's operation order ->
// offset from the conjunction -> b^2 from the set's series -> the flux from the piece of the set's F table the
// sample falls on -> chi^2 term.  A sample whose piece is missing or invalid (NaN) goes on the fix list
// (k_walk_fix: the direct quadrature code; a few 1e-4 of the samples, next to the contact point b = 1 - r) and
// counts as zero here.  The lane sums of an item are added by a fixed shuffle tree into the item's partial sum; the
// lane that completes a set adds the set's partial sums in item order: bit-reproducible.
template <bool FLUX>
__global__ void __launch_bounds__(128, JXP_ITEMS_MINB) k_walk_items(const __grid_constant__ WalkItemArgs a) {
  __shared__ __align__(16) WalkItemSmem s_all[4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (!walk_items_active(a, lane)) return;
  const unsigned n_items = (unsigned)a.state[12];
  if (n_items == 0u) return;
  WalkItemSmem& sm = s_all[wid];
  const double* th = sm.tabs;
  const double* rec = sm.rec;
    unsigned cur_set = 0xffffffffu;
  unsigned n_ld = 0, n_inl = 0;
  // constants of the current set (registers; re-read when the set changes)
  double xs = 0.0, xe = 0.0, mc = 0.0, zsc = 0.0, zoff = 0.0, t0 = 0.0, tref = 0.0, per = 0.0, iper = 0.0;
  double ti1 = 0.0, ti2 = 0.0, ti3 = 0.0, te1 = 0.0;
  int j0 = 0, db = 0;
  // claims: a GROUP of WK_GROUP consecutive items of one set (~500 samples), or -- when there are fewer groups than
  // two per warp -- half of one: a group's rounds are split in two halves with a partial sum each, whoever walks
  // them, so the claim size changes nothing in the result.  The answer of the atomic is read a claim later.
  const unsigned n_groups = n_items / WK_GROUP;
  const unsigned n_half = 2u * n_groups;
  const unsigned R = n_groups >= 2u * gridDim.x * 4u ? 2u : 1u;  // halves per claim
  unsigned long long nxt = 0ull;
  if (lane == 0) nxt = atomicAdd(a.state_rw + 13, (unsigned long long)R);
  for (;;) {
    const unsigned hg0 = (unsigned)__shfl_sync(0xffffffffu, nxt, 0);
    if (hg0 >= n_half) break;
    if (lane == 0) nxt = atomicAdd(a.state_rw + 13, (unsigned long long)R);
    const unsigned grp = hg0 >> 1;
    uint4 its = make_uint4(0u, 0u, 0u, 0u);
    if ((unsigned)lane < WK_GROUP) its = __ldg(a.items + (size_t)grp * WK_GROUP + lane);
    const unsigned set = __shfl_sync(0xffffffffu, its.x, 0);
    // the group's items as ONE sample stream: position p of the stream belongs to the item k with P_k <= p < P_k+1
    // (P = exclusive prefix sums of the items' lengths, lanes 0..7), so every round is full but the last
    unsigned pin = its.z;
#pragma unroll
    for (int o = 1; o < (int)WK_GROUP; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, pin, o);
      if (lane >= o) pin += v;
    }
    const unsigned pex = pin - its.z;
    const unsigned total = __shfl_sync(0xffffffffu, pin, WK_GROUP - 1);
    unsigned Pj[WK_GROUP];
#pragma unroll
    for (int j = 1; j < (int)WK_GROUP; ++j) Pj[j] = __shfl_sync(0xffffffffu, pex, j);
    const unsigned lo_first = __shfl_sync(0xffffffffu, its.y, 0);
    const unsigned dofs = its.y - pex;  // (lanes 0..7) sample index minus stream position, constant along an item
    {
      if (set != cur_set) {
        __syncwarp();
        const double2* T = reinterpret_cast<const double2*>(a.tab + (size_t)set * TB_DOUBLES);
        double2* dst = reinterpret_cast<double2*>(sm.tabs);
#pragma unroll
        for (int i = 0; i < (TB_DOUBLES / 2 + 31) / 32; ++i)
          if (32 * i + lane < TB_DOUBLES / 2) dst[32 * i + lane] = __ldg(T + 32 * i + lane);
        sm.rec[lane] = __ldg(a.rec + (size_t)set * WALK_REC + lane);
        cur_set = set;
        __syncwarp();
        const double* H = th + TB_HDR;
        xs = H[H_XS]; xe = H[H_XE]; ti1 = H[H_TI1]; ti2 = H[H_TI2]; ti3 = H[H_TI3]; te1 = H[H_TE1];
        mc = H[H_MC]; zsc = H[H_ZSC]; zoff = H[H_ZOFF];
        j0 = (int)((unsigned)__double_as_longlong(H[H_FINFO]) & 0xffu);
        db = (int)((unsigned)__double_as_longlong(H[H_BINFO]) & 0xffu);
        t0 = rec[WK_T0]; tref = rec[WK_TREF]; per = rec[WK_PERIOD]; iper = rec[WK_INVP];
      }
      double csum = 0.0;
      unsigned n_nb = 0;  // samples parked in sm.nb_* (warp-uniform)
      // one round: NN samples per lane from position r0 of the group's stream on
      auto round = [&](auto nn_tag, unsigned r0) {
        constexpr int NN = decltype(nn_tag)::value;
        bool have[NN];
        unsigned ix[NN];
        double ts[NN];
        double2 awv[NN];
#pragma unroll
        for (int h = 0; h < NN; ++h) {
          const unsigned p = r0 + 32u * h + lane;
          have[h] = p < total;
          // k = #{j : P_j <= p} (the P_j ascend): three compares of a binary search instead of seven
          static_assert(WK_GROUP == 8, "binary search over eight items");
          unsigned k = p >= Pj[4] ? 4u : 0u;
          k += p >= (k ? Pj[6] : Pj[2]) ? 2u : 0u;
          const unsigned m1 = (k & 4u) ? ((k & 2u) ? Pj[7] : Pj[5]) : ((k & 2u) ? Pj[3] : Pj[1]);
          k += p >= m1 ? 1u : 0u;
          const unsigned d_k = __shfl_sync(0xffffffffu, dofs, k);
          ix[h] = have[h] ? p + d_k : lo_first;  // (padding lanes read a valid sample and contribute zero)
          ts[h] = __ldg(a.t + ix[h]);
          if (!FLUX) awv[h] = __ldg(a.aw + ix[h]);
        }
        // ---- t -> mean anomaly (keplerian.py:538, kepler.py:31) in the reference's operation order.  The
        // reference reduces M to [0, 2 pi) by subtracting k fl(2 pi) (exactly: jxp_math.cuh mod_two_pi); here k is
        // chosen so that the remainder lands next to the conjunction's mean anomaly instead -- the same number up
        // to one exact step of fl(2 pi) -- and the offset from the conjunction is one more subtraction.
        double zv[NN], b2[NN];
#pragma unroll
        for (int h = 0; h < NN; ++h) {
          const double x = (ts[h] - t0) - tref;
          const double num = Num<double>::two_pi * x;
          const double q = num * iper;
          const double r = fma(-per, q, num);
          const double M = fma(r, iper, q);
          const double kk = rint((M - mc) * Num<double>::inv_two_pi);
          const double d = fma(-kk, Num<double>::two_pi, M) - mc;
          zv[h] = fma(d, zsc, zoff);
          b2[h] = 0.0;
        }
        // b^2 from the series, four coefficients per step (the degree is padded to 4 k + 3 with zeros); degree
        // <= 11 (most sets) without a loop
        auto b2_step = [&](int k) {
          const double2 c1 = *reinterpret_cast<const double2*>(th + TB_B2 + k - 1);  // (m_{k-1}, m_k)
          const double2 c0 = *reinterpret_cast<const double2*>(th + TB_B2 + k - 3);  // (m_{k-3}, m_{k-2})
#pragma unroll
          for (int h = 0; h < NN; ++h) {
            b2[h] = fma(fma(b2[h], zv[h], c1.y), zv[h], c1.x);
            b2[h] = fma(fma(b2[h], zv[h], c0.y), zv[h], c0.x);
          }
        };
        if (db > 11) {
#pragma unroll 1
          for (int k = db | 3; k > 11; k -= 4) b2_step(k);
        }
        b2_step(11);
        b2_step(7);
        b2_step(3);
        // ---- the piece of the F table: binade of dx = (1 - r)^2 - b^2 inside the disk, else the ingress / egress
        // pieces in sqrt(b^2 - (1 - r)^2) / sqrt((1 + r)^2 - b^2)
        // (the edge code runs for the 32-sample slices that have an ingress / egress lane only: a lane inside the
        // disk keeps its binade piece either way, a lane out of transit reads the all-zero piece)
        bool intr[NN], ins[NN];
        double q[NN];
        int pc[NN];
#pragma unroll
        for (int h = 0; h < NN; ++h) {
          intr[h] = have[h] && (b2[h] < xe);
          q[h] = xs - b2[h];
          ins[h] = q[h] > 0.0;
          const int jb = tb_binade(q[h]);
          pc[h] = jb < j0 ? j0 : (jb >= TB_NB ? TB_P_BAD : jb);
          n_ld += intr[h] ? 1u : 0u;
        }
        {
#pragma unroll
          for (int h = 0; h < NN; ++h) {
            if (!__any_sync(0xffffffffu, intr[h] && !ins[h])) continue;
            const bool ing = b2[h] < 1.0;
            const double w = fsqrt_pos(fmax(ing ? -q[h] : xe - b2[h], 1e-300));
            const int pe = ing ? TB_P_ING + (w >= ti1 ? 1 : 0) + (w >= ti2 ? 1 : 0) + (w >= ti3 ? 1 : 0)
                               : TB_P_EGR + (w >= te1 ? 1 : 0);
            pc[h] = ins[h] ? pc[h] : pe;
            q[h] = ins[h] ? q[h] : w;
          }
        }
        // (a sample that is not in transit reads the all-zero piece: F = 0 without a select)
        const double2* P[NN];
        double v[NN], acc[NN];
#pragma unroll
        for (int h = 0; h < NN; ++h) {
          P[h] = reinterpret_cast<const double2*>(th + TB_PIECE0 + (intr[h] ? pc[h] : TB_P_ZERO) * TB_PSTRIDE);
          const double2 so = *P[h];
          v[h] = fma(q[h], so.x, so.y);
          acc[h] = 0.0;
        }
#pragma unroll
        for (int k = TB_NC / 2; k >= 1; --k) {
#pragma unroll
          for (int h = 0; h < NN; ++h) {
            const double2 cc = P[h][k];  // (m_{2k-2}, m_{2k-1})
            acc[h] = fma(fma(acc[h], v[h], cc.y), v[h], cc.x);
          }
        }
        // a NaN (missing / invalid piece): onto the fix list; zero here
        bool need[NN], need_any = false;
#pragma unroll
        for (int h = 0; h < NN; ++h) {
          need[h] = !(acc[h] == acc[h]);
          need_any = need_any || need[h];
        }
        if (__any_sync(0xffffffffu, need_any)) {
          // inside the disk, below the last binade, safely below the contact point (what ld_inside_u needs: the
          // same conditions as in k_walk_eval): the direct inside-the-disk code here and now
          const double ar = fabs(rec[WK_RR]);
          const bool r_ok = (ar < 1.0) && (ar >= 2e-6);
          const double b_lim = (1.0 - ar) - 4e-15;
#pragma unroll
          for (int h = 0; h < NN; ++h) {
            const double bq = fsqrt(fmax(b2[h], 0.0));
            const bool ff = need[h] && ins[h] && pc[h] == TB_P_BAD && r_ok && bq < b_lim;
            const unsigned mf = __ballot_sync(0xffffffffu, ff);
            if (mf) {
              const unsigned slot = n_nb + __popc(mf & ((1u << lane) - 1u));
              if (ff && slot < 32u) {  // (a 33rd sample of one half-group takes the fix list)
                sm.nb_b[slot] = bq;
                sm.nb_ix[slot] = ix[h];
                acc[h] = 0.0;
                need[h] = false;
              }
              n_nb = min(n_nb + (unsigned)__popc(mf), 32u);
            }
          }
          // what is left (a piece that did not converge): onto the fix list
#pragma unroll
          for (int h = 0; h < NN; ++h) {
            const unsigned mn = __ballot_sync(0xffffffffu, need[h]);
            unsigned long long fb = 0ull;
            if (lane == 0 && mn) fb = atomicAdd(a.state_rw + 14, (unsigned long long)__popc(mn));
            fb = __shfl_sync(0xffffffffu, fb, 0);
            if (need[h]) {
              const unsigned long long slot = fb + __popc(mn & ((1u << lane) - 1u));
              if (slot < a.cap_fix) {
                WalkFix f;
                f.set = set;
                f.idx = ix[h];
                f.b2 = b2[h];
                a.fix[slot] = f;
              }
            }
            acc[h] = need[h] ? 0.0 : acc[h];
          }
        }
        if (FLUX) {
#pragma unroll
          for (int h = 0; h < NN; ++h) {
            if (acc[h] != 0.0) {
              const size_t o = (size_t)set * (size_t)a.n_times + ix[h];
              if (a.flux_sum) a.flux_sum[o] = (a.body > 0) ? a.flux_sum[o] + acc[h] : acc[h];
              if (a.flux) a.flux[o * (size_t)a.n_bodies + (size_t)a.body] = acc[h];
            }
          }
        } else {
          // chi^2 = sum_t ((y - 1) / sigma)^2 + sum_items [ (r0 - m / sigma)^2 - r0^2 ], the bracket = m (m / sigma^2 -
          // 2 r0 / sigma) (exactly 0 for m = 0)
#pragma unroll
          for (int h = 0; h < NN; ++h) csum = fma(acc[h], fma(acc[h], awv[h].y, -awv[h].x), csum);
        }
      };
      constexpr unsigned RS = 32u * JXP_ITEMS_NN;  // samples of a full round
      const unsigned nfull = total / RS, rem = total - RS * nfull;
      const unsigned nrounds = nfull + (rem ? 1u : 0u), nsplit = (nrounds + 1u) / 2u;
      for (unsigned half = hg0 & 1u; half < (hg0 & 1u) + R; ++half) {
        csum = 0.0;
        const unsigned rend = half ? nrounds : nsplit;
        for (unsigned r = half ? nsplit : 0u; r < rend; ++r) {
          if (r < nfull)
            round(std::integral_constant<int, JXP_ITEMS_NN>{}, RS * r);
#if JXP_ITEMS_NN >= 4
          else if (rem > 96u)
            round(std::integral_constant<int, 4>{}, RS * r);
#endif
          else if (rem > 64u)
            round(std::integral_constant<int, 3>{}, RS * r);
          else if (rem > 32u)
            round(std::integral_constant<int, 2>{}, RS * r);
          else
            round(std::integral_constant<int, 1>{}, RS * r);
        }
        if (n_nb) {  // the half-group's samples next to the contact point: one per lane, in the order they were parked
          __syncwarp();
          if ((unsigned)lane < n_nb) {
            const unsigned ixn = sm.nb_ix[lane];
            const int lmax = (int)((unsigned)__double_as_longlong(rec[WK_FLAGS]) >> 8);
            const double F = walk_inside_one(sm.nb_b[lane], fabs(rec[WK_RR]), lmax, rec[WK_G0], rec[WK_G1], rec[WK_G2]);
            if (FLUX) {
              if (F != 0.0) {
                const size_t o = (size_t)set * (size_t)a.n_times + ixn;
                if (a.flux_sum) a.flux_sum[o] = (a.body > 0) ? a.flux_sum[o] + F : F;
                if (a.flux) a.flux[o * (size_t)a.n_bodies + (size_t)a.body] = F;
              }
            } else {
              const double2 awn = __ldg(a.aw + ixn);
              csum = fma(F, fma(F, awn.y, -awn.x), csum);
            }
          }
          n_inl += n_nb;  // (counted once per warp: see the end of the kernel)
          n_nb = 0;
          __syncwarp();
        }
        if (!FLUX) {
          // the half-group's partial sum (fixed shuffle tree); the claim is counted into its set below, and the lane
          // that completes a set adds the set's partial sums in order
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
          if (lane == 0) a.part[2u * grp + half] = csum;
        }
      }
      if (!FLUX && lane == 0) {  // one fence and one count for the claim's partial sums
        __threadfence();
        const unsigned prev = atomicAdd(&a.wset_rw[set].done, R);
        if (prev + R == a.wset[set].cntA) walk_finish_items(a, set);
      }
    }
  }
  n_ld = __reduce_add_sync(0xffffffffu, n_ld);
  if (lane == 0) atomicAdd(a.tstat, (unsigned long long)(n_ld - n_inl));  // (k_walk_fix takes its samples off again)
  if (lane == 0 && n_inl) atomicAdd(a.tstat + 1, (unsigned long long)n_inl);  // (answered by the direct code)
  if (lane == 0) atomicAdd(a.counters + 1, (unsigned long long)n_ld);
}

// The samples of table sets the tables did not answer: the direct flux code (core/limb_dark.py:19-196), one
// sample per thread.  Flux outputs are stored; chi^2 terms are added to the set's sum in FIXED POINT (2^-40: exact
// integer adds, so the order the list was filled in does not matter), and the last CTA to finish folds the sums
// into the log-likelihoods the item kernel wrote: bit-reproducible.
template <bool FLUX>
__global__ void __launch_bounds__(256) k_walk_fix(const __grid_constant__ WalkItemArgs a) {
  const int lane = threadIdx.x & 31;
  if (!walk_items_active(a, lane)) return;
  unsigned long long n_fix = a.state[14];
  if (n_fix == 0ull) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(a.tstat, 0ull - (n_fix < a.cap_fix ? n_fix : a.cap_fix));
    atomicAdd(a.tstat + 1, n_fix);
  }
  if (n_fix > a.cap_fix) {  // the list did not fit: the window-test kernels behind this one redo the call
    if (blockIdx.x == 0 && threadIdx.x == 0) a.counters[3] |= 2ull;
    return;
  }
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_fix;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const WalkFix f = a.fix[i];
    const double* rec = a.rec + (size_t)f.set * WALK_REC;
    const int lmax = (int)((unsigned)__double_as_longlong(rec[WK_FLAGS]) >> 8);
    double bq[1] = {fsqrt(fmax(f.b2, 0.0))}, rq[1] = {rec[WK_RR]}, s0[1], s1[1], s2[1];
    ld_solution_n<1, 1>(bq, rq, lmax, s0, s1, s2);
    const double F = ld_combine(lmax, s0[0], s1[0], s2[0], rec[WK_G0], rec[WK_G1], rec[WK_G2]);
    if (FLUX) {
      if (F != 0.0) {
        const size_t o = (size_t)f.set * (size_t)a.n_times + f.idx;
        if (a.flux_sum) a.flux_sum[o] = (a.body > 0) ? a.flux_sum[o] + F : F;
        if (a.flux) a.flux[o * (size_t)a.n_bodies + (size_t)a.body] = F;
      }
    } else {
      const double wv = a.winv[f.idx];
      const double r0v = (a.y[f.idx] - 1.0) * wv;
      const double mw = F * wv;
      const double term = (F != 0.0) ? (-mw) * (2.0 * r0v - mw) : 0.0;
      // integer part and fraction (2^-44) separately: exact adds, no overflow below ~1e6 samples per set
      const double hi = rint(term);
      const long long qh = __double2ll_rn(hi), ql = __double2ll_rn((term - hi) * WK_FIX_SCALE);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.fixsum + 2 * (size_t)f.set), (unsigned long long)qh);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.fixsum + 2 * (size_t)f.set + 1), (unsigned long long)ql);
    }
  }
  if (!FLUX) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.state_rw + 15, 1ull) == (unsigned long long)(gridDim.x - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      // (eight sets per thread and step, every load of a step issued before the first use: two round trips per step)
      const longlong2* fs = reinterpret_cast<const longlong2*>(a.fixsum);
      for (int s0 = threadIdx.x; s0 < a.n_sets; s0 += 8 * (int)blockDim.x) {
        longlong2 v[8];
        unsigned isitem[8];
        double ll[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int s = s0 + u * (int)blockDim.x;
          const bool in = s < a.n_sets;
          v[u] = in ? __ldcg(fs + s) : make_longlong2(0ll, 0ll);
          isitem[u] = in ? __ldcg(&a.wset[s].cntB) : 0u;
          ll[u] = in ? __ldcg(a.loglike + s) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int s = s0 + u * (int)blockDim.x;
          if ((v[u].x != 0ll || v[u].y != 0ll) && isitem[u] == 1u)
            a.loglike[s] = fma(-0.5, (double)v[u].x + (double)v[u].y * (1.0 / WK_FIX_SCALE), ll[u]);
        }
      }
    }
  }
}

}  // namespace jxp

namespace jxph {


// one warp per set when the sets fill the machine, else one CTA per set (k_walk_tables)
static void launch_tables(const jxp::TabArgs& ta, int n_sets, cudaStream_t st) {
  if (n_sets >= 8 * sm_count())
    jxp::k_walk_tables<1><<<(n_sets + 3) / 4, 128, 0, st>>>(ta);
  else
    jxp::k_walk_tables<4><<<n_sets, 128, 0, st>>>(ta);
}

// A side stream per (host thread, device) for work that may run beside the caller's stream (fork / join by events).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream* side_stream() {
  static thread_local SideStream tab[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = tab[dev];
  if (s.stream == nullptr) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) {
      s.stream = nullptr;
      return nullptr;
    }
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      cudaStreamDestroy(s.stream);
      s.stream = nullptr;
      return nullptr;
    }
  }
  return &s;
}

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
  const bool use_tables = (dev_options() & JXP_DEV_NO_TABLES) == 0;
  a.tab = use_tables ? reinterpret_cast<double*>(ws + L.off_wk_tab) : nullptr;
  a.items = use_tables && !expo ? reinterpret_cast<uint4*>(ws + L.off_wk_items) : nullptr;
  a.cap_items = L.wk_cap_items;
  a.fixsum = reinterpret_cast<long long*>(ws + L.off_wk_fixsum);
  a.aw = (a.items != nullptr && loglike) ? reinterpret_cast<double2*>(ws + L.off_wk_aw) : nullptr;
  TabArgs ta;
  ta.bodies = bodies;
  ta.n_sets = n_sets;
  ta.n_bodies = n_bodies;
  ta.u = u;
  ta.n_u = n_u;
  ta.u_stride = u_stride;
  ta.tab = a.tab;

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
  b.tab = a.tab;
  b.tstat = reinterpret_cast<unsigned long long*>(d_scalars + WALK_TSTAT0);
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
    cudaEvent_t join_ev = nullptr;
    // (slots 10, 11 hold the likelihood terms the plan launch writes; 12, 13: item allocation / claims; 14, 15: k_walk_fix)
    // (the table statistics sit right behind the state block: one fill for both at the first body)
    static_assert(WALK_TSTAT0 == WALK_STATE0 + 16, "state block and table statistics are contiguous");
    cudaError_t e = cudaMemsetAsync(a.state, 0, (body == 0 ? 20 : 16) * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk): memset: %s", cudaGetErrorString(e));
    if (ev_start && (ev_which == 1 || ev_which == 2) && body == 0) cudaEventRecord(ev_start, st);
    bool forked = false;
    if (use_tables) {
      // the flux tables depend on (r, u) alone: built beside the plan on a side stream (fork / join by events, which
      // a stream capture follows too); serial when the table launch itself is being timed
      ta.body = body;
      SideStream* ss = ev_which == 2 ? nullptr : side_stream();
      if (ss != nullptr && cudaEventRecord(ss->fork, st) == cudaSuccess &&
          cudaStreamWaitEvent(ss->stream, ss->fork, 0) == cudaSuccess) {
        launch_tables(ta, n_sets, ss->stream);
        forked = true;
      } else {
        launch_tables(ta, n_sets, st);
      }
      int rct = check_launch("jxp_system_light_curve(walk tables)");
      if (rct == JXP_OK && forked && cudaEventRecord(ss->join, ss->stream) != cudaSuccess)
        rct = fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk tables): event record");
      if (rct != JXP_OK) return rct;
      if (forked) join_ev = ss->join;
    }
    if (ev_stop && ev_which == 2 && body == n_bodies - 1) cudaEventRecord(ev_stop, st);
    {
      const int n_cta = a.n_plan_ctas + LL_NCTA_H;
      const int minb = n_cta <= 2 * sm_count() ? 2 : (n_cta <= 3 * sm_count() ? 3 : 4);
      if (loglike) {
        if (minb == 2)
          k_walk_plan<CH, true, 2><<<n_cta, 256, 0, st>>>(a);
        else if (minb == 3)
          k_walk_plan<CH, true, 3><<<n_cta, 256, 0, st>>>(a);
        else
          k_walk_plan<CH, true, 4><<<n_cta, 256, 0, st>>>(a);
      } else {
        if (minb == 2)
          k_walk_plan<CH, false, 2><<<n_cta, 256, 0, st>>>(a);
        else if (minb == 3)
          k_walk_plan<CH, false, 3><<<n_cta, 256, 0, st>>>(a);
        else
          k_walk_plan<CH, false, 4><<<n_cta, 256, 0, st>>>(a);
      }
    }
    if (ev_stop && ev_which == 1 && body == n_bodies - 1) cudaEventRecord(ev_stop, st);
    int rc = check_launch("jxp_system_light_curve(walk plan)");
    if (rc != JXP_OK) return rc;
    // (both evaluation kernels read the tables)
    if (join_ev != nullptr && cudaStreamWaitEvent(st, join_ev, 0) != cudaSuccess)
      return fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk tables): join");
    join_ev = nullptr;
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
    if (a.items != nullptr) {
      WalkItemArgs c;
      c.rec = a.rec;
      c.tab = a.tab;
      c.items = a.items;
      c.part = reinterpret_cast<double*>(ws + L.off_wk_ipart);
      c.wset = a.wset;
      c.wset_rw = a.wset;
      c.state = a.state;
      c.state_rw = a.state;
      c.cap_chunk = a.cap_chunk;
      c.scalars = d_scalars;
      c.t = t;
      c.n_times = n_times;
      c.y = y;
      c.winv = a.winv;
      c.aw = a.aw;
      c.flux = flux;
      c.flux_sum = flux_sum;
      c.loglike = loglike;
      c.counters = a.counters;
      c.tstat = reinterpret_cast<unsigned long long*>(d_scalars + WALK_TSTAT0);
      c.n_bodies = n_bodies;
      c.body = body;
      c.cap_fix = L.wk_cap_fix;
      c.fix = reinterpret_cast<WalkFix*>(ws + L.off_wk_fix);
      c.fixsum = a.fixsum;
      c.n_sets = n_sets;
      static thread_local int it_per_sm[2] = {0, 0};
      int& ips = it_per_sm[loglike ? 0 : 1];
      if (ips == 0) {
        cudaError_t eo = loglike ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ips, k_walk_items<false>, 128, 0)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ips, k_walk_items<true>, 128, 0);
        if (eo != cudaSuccess || ips < 1) {
          ips = 0;
          return fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk items): occupancy query: %s", cudaGetErrorString(eo));
        }
      }
      if (join_ev != nullptr && cudaStreamWaitEvent(st, join_ev, 0) != cudaSuccess)
        return fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk tables): join");
      join_ev = nullptr;
      const unsigned igrid = (unsigned)(ips * sm_count());
      const unsigned fgrid = (unsigned)((n_sets + 3) / 4 < 148 ? (n_sets + 3) / 4 : 148);
      if (ev_start && ev_which == 3 && body == 0) cudaEventRecord(ev_start, st);
      if (loglike)
        k_walk_items<false><<<igrid, 128, 0, st>>>(c);
      else
        k_walk_items<true><<<igrid, 128, 0, st>>>(c);
      if (ev_stop && ev_which == 3 && body == n_bodies - 1) cudaEventRecord(ev_stop, st);
      if (loglike)
        k_walk_fix<false><<<fgrid, 256, 0, st>>>(c);
      else
        k_walk_fix<true><<<fgrid, 256, 0, st>>>(c);
      rc = check_launch("jxp_system_light_curve(walk items)");
      if (rc != JXP_OK) return rc;
    }
    if (join_ev != nullptr && cudaStreamWaitEvent(st, join_ev, 0) != cudaSuccess)
      return fail(JXP_ERR_CUDA, "jxp_system_light_curve(walk tables): join");
  }
  return JXP_OK;
}

}  // namespace jxph
