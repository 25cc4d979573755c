// Per-set function tables of the transit walk (jxp_walk.cu).
//
// The reference evaluates, at every sample, a Kepler solve (core/kepler.py:14-108), the sky position
// (orbits/keplerian.py:537-637) and the limb-darkened flux with a ten-node quadrature
// (core/limb_dark.py:19-196).  For ONE parameter set both are functions of a single variable:
//   * the squared separation b^2 is a smooth function of the mean anomaly across the transit window (the same for
//     every transit of the set), and
//   * the flux F is a function of b alone (r and the limb-darkening law are fixed) -- analytic for b < 1 - r with
//     branch points of the quadrature's square roots clustered just beyond b = 1 - r, and analytic in the square
//     root of the distance to either contact point for 1 - r < b < 1 + r.
// A series with many time samples per set (config 3: 1e5 samples, ~3000 of them in transit, per set) therefore pays
// for tabulating both once per set: the plan fits b^2(M) (32 Chebyshev nodes = the 32 lanes of the set's plan warp,
// adaptive degree <= 21) and k_walk_tables fits F on pieces graded towards the contact points (degree 17 each:
// binades of dx = (1 - r)^2 - b^2 inside the disk, four pieces in sqrt(b^2 - (1 - r)^2) and two in
// sqrt((1 + r)^2 - b^2) for ingress / egress).  The node values come from the SAME device code the direct path
// uses (kepler_xy_u, ld_solution_n), so the tables reproduce the reference formula, quadrature error included,
// to the rounding noise of that code (~3e-16 absolute in flux, measured); a fit whose last Chebyshev coefficients
// are not down at that level is marked invalid and its samples take the direct code.  The fits are converted to
// the monomial basis (safe: accepted fits decay by >= 5x per degree, faster than the basis change grows) so that
// a sample costs one Horner chain per table: ~14 + 18 fused multiply-adds instead of ~75 + 250..470 FP64
// instructions.  Nothing here decides a value on its own: a sample whose piece is missing, invalid or out of range
// is evaluated by the direct code in the same kernel.
#pragma once

#include "jxp_host.h"
#include "jxp_math.cuh"

namespace jxp {

#include "jxp_tables_data.inc"

constexpr int TB_NC = 18;             // coefficients per flux piece (degree 17)
constexpr int TB_NB = 12;             // inside-the-disk pieces: dx in [2^-(j+1), 2^-j), j < TB_NB
constexpr int TB_P_ING = TB_NB;       // 4 ingress-side pieces (1 - r <= b < 1)
constexpr int TB_P_EGR = TB_NB + 4;   // 2 egress-side pieces (1 <= b < 1 + r)
constexpr int TB_P_BAD = TB_NB + 6;   // never valid
constexpr int TB_P_ZERO = TB_NB + 7;  // F = 0 (samples out of transit)
constexpr int TB_NPIECE = TB_NB + 8;
constexpr int TB_PSTRIDE = 20;        // scale, offset, 18 monomial coefficients
constexpr int TB_B2 = 0;              // b^2 series: 24 monomial coefficients
constexpr int TB_B2N = 24;
constexpr int TB_HDR = 24;            // 16 header doubles
constexpr int TB_PIECE0 = 40;
constexpr int TB_DOUBLES = TB_PIECE0 + TB_NPIECE * TB_PSTRIDE;
static_assert(TB_DOUBLES == jxph::WALK_TAB, "workspace carving (jxp_host.h) and table layout disagree");
static_assert(TB_PIECE0 % 2 == 0 && TB_PSTRIDE % 2 == 0 && TB_DOUBLES % 2 == 0, "pieces are read as double2");
// header fields (k_walk_tables: XS .. FINFO, k_walk_plan: MC .. BINFO)
enum : int { H_XS = 0, H_XE, H_TI1, H_TI2, H_TI3, H_TE1, H_FINFO, H_R7, H_MC, H_ZSC, H_ZOFF, H_BINFO };
constexpr double TB_ING1 = 0.05, TB_ING2 = 0.2, TB_ING3 = 0.5, TB_EGR1 = 0.65;  // piece boundaries, fractions of the variable's range
constexpr unsigned TB_VALID = 0x80000000u;
constexpr double TB_R_MIN = 1e-3, TB_R_MAX = 0.5;  // radius ratios the tables are built for

__device__ __forceinline__ int tb_binade(double dx) {  // dx in [2^-(j+1), 2^-j) -> j (dx > 0, normal)
  return 1022 - ((__double2hiint(dx) >> 20) & 0x7ff);
}

struct TabArgs {
  const double* bodies;  // [n_sets][n_bodies][JXP_BODY_NFIELDS]
  int n_sets, n_bodies, body;
  const double* u;
  int n_u;
  int64_t u_stride;
  double* tab;  // [n_sets][TB_DOUBLES]
};

__device__ __forceinline__ double tb_pow2(int e) {  // 2^e, -1022 <= e <= 1023
  return __hiloint2double((1023 + e) << 20, 0);
}

// The flux pieces of (r, u) and their header: WPS warps per parameter set (1 when there are sets enough to fill the
// machine, 4 -- one CTA per set -- when there are not: the work of a set is a ~30 us chain for one warp).
constexpr int TB_WARPS = 4;
// 7 CTAs per SM (72 registers): the 4096 warps of config 3 are resident at once instead of in 1.7 waves of ~30 us
// chains; step 0.2713 -> 0.2675 ms on 4096 sets, 0.0900 -> 0.0865 ms on 512 (4 / 5 / 6 / 8 / 10 CTAs per SM:
// 0.2713 / 0.2708 / 0.2716 / 0.2690 / 0.2711; tools/gpu_ab.sh)
#ifndef JXP_TAB_MINB
#define JXP_TAB_MINB 7
#endif
template <int WPS>
__global__ void __launch_bounds__(32 * TB_WARPS, JXP_TAB_MINB) k_walk_tables(const __grid_constant__ TabArgs a) {
  constexpr int SPC = TB_WARPS / WPS;  // sets per CTA
  __shared__ double s_f[SPC][TB_NPIECE][TB_NC];  // node values
  __shared__ double s_c[SPC][TB_NPIECE][TB_NC];  // Chebyshev coefficients
  __shared__ double s_q[SPC][TB_NPIECE][2];      // q-range of a piece: q = dx inside the disk,
                                                  // sqrt(b^2 - xs) / sqrt(xe - b^2) for ingress / egress
  __shared__ double s_dct[TB_NC][TB_NC], s_c2m[TB_NC][TB_NC];
  __shared__ unsigned char s_pl[SPC][TB_NPIECE];
  __shared__ int s_good[SPC];
  for (int i = threadIdx.x; i < TB_NC * TB_NC; i += blockDim.x) {
    s_dct[i / TB_NC][i % TB_NC] = d_tb_dctT18[i / TB_NC][i % TB_NC];
    s_c2m[i / TB_NC][i % TB_NC] = d_tb_c2mT[i / TB_NC][i % TB_NC];
  }
  const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) / WPS, sub = (threadIdx.x >> 5) % WPS;
  const int sraw = blockIdx.x * SPC + w;
  const bool live_set = sraw < a.n_sets;
  const int s = live_set ? sraw : a.n_sets - 1;  // (a CTA's spare warps recompute the last set and write nothing)
  if (threadIdx.x < SPC) s_good[threadIdx.x] = 1;
  const double* p = a.bodies + ((size_t)s * a.n_bodies + a.body) * JXP_BODY_NFIELDS;
  double* T = a.tab + (size_t)s * TB_DOUBLES;
  const double rstar = p[JXP_BODY_CENTRAL_RADIUS];
  const double rr = fabs(p[JXP_BODY_RADIUS] / rstar);
  const double* us = a.u + (size_t)s * a.u_stride;
  double g0, g1, g2;
  ld_greens<double>(a.n_u, a.n_u > 0 ? us[0] : 0.0, a.n_u > 1 ? us[1] : 0.0, g0, g1, g2);
  const int lmax = a.n_u;
  const bool ok = rr >= TB_R_MIN && rr <= TB_R_MAX && isfinite(g0) && isfinite(g1) && isfinite(g2);
  const double xs = (1.0 - rr) * (1.0 - rr), xe = (1.0 + rr) * (1.0 + rr);
  const double Wi = fsqrt(1.0 - xs), We = fsqrt(xe - 1.0);
  const int j0 = ok ? tb_binade(xs) : 0;
  // smallest separation the set's transits reach, estimated from the conjunction (keplerian.py:318-327) and
  // taken generously: pieces that lie entirely below it are not built (a sample that lands on one anyway takes
  // the direct code)
  double bcut = 0.0;
  {
    double e = p[JXP_BODY_ECC];
    if (e < 0.0) e = 0.0;
    const double sw = p[JXP_BODY_SIN_OMEGA], ci = p[JXP_BODY_COS_INCL], si = p[JXP_BODY_SIN_INCL];
    const double opec = si >= 0.0 ? 1.0 + e * sw : 1.0 - e * sw;
    const double b0 = fabs(fdiv(p[JXP_BODY_SEMIMAJOR] * (1.0 - e * e) * ci, opec * rstar));
    bcut = 0.98 * b0 - 0.005;
    if (!(bcut > 0.0)) bcut = 0.0;  // (NaN: build everything)
  }
  // lane = piece: its q-range and whether any sample can land on it (every warp of the set computes the same)
  bool active = false;
  double qlo = 0.0, qhi = 0.0;
  if (lane < TB_P_BAD) {
    double bhi2;
    if (lane < TB_NB) {
      qlo = tb_pow2(-(lane + 1));
      qhi = lane == j0 ? xs : tb_pow2(-lane);
      bhi2 = xs - qlo;
    } else if (lane < TB_P_EGR) {
      const int k = lane - TB_P_ING;
      qlo = Wi * (k == 0 ? 0.0 : (k == 1 ? TB_ING1 : (k == 2 ? TB_ING2 : TB_ING3)));
      qhi = Wi * (k == 0 ? TB_ING1 : (k == 1 ? TB_ING2 : (k == 2 ? TB_ING3 : 1.0)));
      bhi2 = fma(qhi, qhi, xs);
    } else {
      const int k = lane - TB_P_EGR;
      qlo = We * (k == 0 ? 0.0 : TB_EGR1);
      qhi = We * (k == 0 ? TB_EGR1 : 1.0);
      bhi2 = fma(-qlo, qlo, xe);
    }
    active = ok && !(lane < TB_NB && lane < j0) && (bhi2 > bcut * bcut);
  }
  const unsigned amask = __ballot_sync(0xffffffffu, active);
  const int npl = __popc(amask);
  if (sub == 0) {
    if (lane < TB_P_BAD) {
      s_q[w][lane][0] = qlo;
      s_q[w][lane][1] = qhi;
    }
    if (active) s_pl[w][__popc(amask & ((1u << lane) - 1u))] = (unsigned char)lane;
  }
  __syncthreads();
  // node values: the flattened (piece, node) list, two per lane and 64 per round -- the inside-the-disk pieces
  // first (they come first in the list) with the constant-kappa code of the direct path, then the rest with the
  // general code; a round never mixes the two.  The set's warps take the rounds in turn.
  const int n_nodes = npl * TB_NC;
  const int n_in = __popc(amask & ((1u << TB_NB) - 1u)) * TB_NC;
  const int r_in = (n_in + 63) / 64, r_all = r_in + (n_nodes - n_in + 63) / 64;
  for (int r = sub; r < r_all; r += WPS) {
    const bool inside = r < r_in;
    const int f0 = inside ? 64 * r : n_in + 64 * (r - r_in);
    const int fend = inside ? n_in : n_nodes;
    double bb[2];
    int pi[2], ni[2];
    bool live[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int f = f0 + 32 * h + lane;
      live[h] = f < fend;
      pi[h] = live[h] ? f / TB_NC : 0;
      ni[h] = live[h] ? f % TB_NC : 0;
      const int pid = s_pl[w][pi[h]];
      const double lo = s_q[w][pid][0], hi = s_q[w][pid][1];
      const double q = fma(0.5 * (hi - lo), d_tb_node18[ni[h]], 0.5 * (hi + lo));
      double b2 = pid < TB_NB ? xs - q : (pid < TB_P_EGR ? fma(q, q, xs) : fma(-q, q, xe));
      if (!(b2 > 0.0)) b2 = 0.0;
      bb[h] = fsqrt(b2);
    }
    double F[2];
    if (inside) {
      ld_inside_u<2>(bb, rr, lmax, g0, g1, g2, F);
    } else {
      double rq[2] = {rr, rr}, s0[2], s1[2], s2[2];
      ld_solution_n<2, 1>(bb, rq, lmax, s0, s1, s2);
#pragma unroll
      for (int h = 0; h < 2; ++h) F[h] = ld_combine(lmax, s0[h], s1[h], s2[h], g0, g1, g2);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (live[h]) s_f[w][pi[h]][ni[h]] = F[h];
  }
  __syncthreads();
  // node values -> Chebyshev coefficients (entries <= 1 / 9: backward stable) ...
  for (int f = sub * 32 + lane; f < n_nodes; f += 32 * WPS) {
    const int pi = f / TB_NC, k = f % TB_NC;
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < TB_NC; ++i) acc = fma(s_dct[i][k], s_f[w][pi][i], acc);
    s_c[w][pi][k] = acc;
  }
  __syncthreads();
  // ... -> monomial coefficients (integer matrix; T_k has the parity of k), highest degree first.  A fit whose
  // last two Chebyshev coefficients are not down at the rounding noise of the node values is marked invalid.
  const double tol = 4e-16 + 2e-14 * rr * rr;
  bool all_good = true;
  for (int f = sub * 32 + lane; f < n_nodes; f += 32 * WPS) {
    const int pi = f / TB_NC, j = f % TB_NC;
    const int pid = s_pl[w][pi];
    double acc = 0.0;
    for (int k = TB_NC - 1 - ((TB_NC - 1 - j) & 1); k >= j; k -= 2) acc = fma(s_c2m[k][j], s_c[w][pi][k], acc);
    double* P = T + TB_PIECE0 + pid * TB_PSTRIDE;
    if (live_set) P[2 + j] = acc;
    if (j == 0) {
      const double lo = s_q[w][pid][0], hi = s_q[w][pid][1];
      const double tail = fmax(fabs(s_c[w][pi][TB_NC - 1]), fabs(s_c[w][pi][TB_NC - 2]));
      const bool good = tail <= tol;  // (NaN: not good)
      all_good = all_good && good;
      const double inv = fdiv(1.0, hi - lo);
      if (live_set) {
        P[0] = good ? 2.0 * inv : nan("");
        P[1] = -(hi + lo) * inv;
      }
    }
  }
  // a set with a piece that did not converge is not evaluated from tables at all (large r / r close to the end of
  // the validated range: the direct code is the safe and, with many samples on that piece, the faster choice)
  if (!__all_sync(0xffffffffu, all_good) && lane == 0) atomicAnd(&s_good[w], 0);
  __syncthreads();
  all_good = s_good[w] != 0;
  if (!live_set || sub != 0) return;
  // pieces that were not built: NaN scale, zero coefficients (the all-zero piece: F = 0)
  for (int f = lane; f < TB_NPIECE * TB_PSTRIDE; f += 32) {
    const int pid = f / TB_PSTRIDE, k = f % TB_PSTRIDE;
    if (!((amask >> pid) & 1u)) T[TB_PIECE0 + f] = (k == 0 && pid != TB_P_ZERO) ? nan("") : 0.0;
  }
  if (!all_good)
    for (int pid = lane; pid < TB_P_BAD; pid += 32) T[TB_PIECE0 + pid * TB_PSTRIDE] = nan("");
  if (lane == 0) {
    double* H = T + TB_HDR;
    H[H_XS] = xs;
    H[H_XE] = xe;
    H[H_TI1] = TB_ING1 * Wi;
    H[H_TI2] = TB_ING2 * Wi;
    H[H_TI3] = TB_ING3 * Wi;
    H[H_TE1] = TB_EGR1 * We;
    H[H_FINFO] = __longlong_as_double((long long)(((ok && all_good) ? TB_VALID : 0u) | (unsigned)j0));
    H[H_R7] = 0.0;
  }
}

}  // namespace jxp
