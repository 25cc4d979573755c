// Per-body constants and the per-time-sample evaluation of the fused light-curve path:
//   t -> mean anomaly -> Kepler -> sky position -> b -> limb-darkened flux
// src/jaxoplanet/orbits/keplerian.py:262-340, 537-637; light_curves/limb_dark.py:49-60.
#pragma once

#include "jxp_math.cuh"

namespace jxp {

enum : int { BODY_CIRCULAR = 1, BODY_ALWAYS = 2 };

// Device-side record of one (parameter set, body), produced once per call by the prep
// kernel and broadcast to a CTA through shared memory.  Everything the POSITION stage reads
// (t -> mean anomaly -> Kepler solve -> sky position -> b) stays double in the fp32 mode: a float
// angle of order pi carries 2.4e-7 rad, times a / R* ~ 15 that is 4e-6 of a stellar radius in b and
// several ppm of flux on ingress (measured on BASELINE config 3: 6.8 ppm).  Only in-window samples
// reach that stage, so the fp32 mode keeps its speed where it has it: the flux code (DESIGN.md
// "fp32 mode").
template <typename T>
struct BodyC {
  double t0;          // time_transit
  double tref;        // time_ref
  double period;
  double t0ref;       // t0 + tref               } approximate phase for the
  double inv_period;  // 1 / period              } transit-window pre-test
  double phase_c;     // window centre [cycles]  }
  double wlo, whi;    // window edges relative to the centre [cycles], margins included
  double win_in;      // estimated half-width [cycles] of the part of the transit with the planet fully
                      // inside the disk, slightly short; 0: none.  Only groups work, never decides a value.
  double e, ome, ope, s1me2, inv_ope;
  double cw, sw, ci, si;
  double na_ome2;  // -a (1 - e^2)
  double na;       // -a
  double inv_rstar;
  T rr, onepr;
  int flags;
  int pad;
};

template <typename T>
struct SetC {
  T g0, g1, g2;  // normalised Green's coefficients
  T ghi[6];      // g_3 .. g_8 (general polynomial limb darkening, used by the HI kernels only)
  int lmax;
};

// Mean anomaly at true anomaly f (monotone map used only for the window edges).  The window is widened by
// 1e-9 (relative) and 1e-7 cycles, so the branch-free primitives of jxp_fastmath.cuh (<= 1e-15) serve here:
// the CUDA library versions made the per-set preparation a ~10 us serial chain.
__device__ inline double mean_from_true(double f, double sq1me, double sq1pe, double e) {
  double sh, ch;
  fsincos(0.5 * f, &sh, &ch);
  const double E = 2.0 * fatan2(sq1me * sh, sq1pe * ch);
  double sE, cE;
  fsincos(E, &sE, &cE);
  return E - e * sE;
}

// Conservative transit window in orbital phase.  With theta = f + omega - theta_c the angle
// from mid-transit, the sky-projected separation obeys
//     (b R*)^2 = r0^2 (cos^2 i + sin^2 i sin^2 theta)        (keplerian.py:568-576, 630)
// so a sample can only be in transit (z > 0 and b < 1 + r) if
//     sin^2 theta < ((1 + r)^2 R*^2 / r_min^2 - cos^2 i) / sin^2 i =: S,
// r_min a lower bound of |r0| over the samples considered.  Two passes: r_min = a (1 - e)
// gives a first interval; |r0| >= a (1 - e^2) / (1 + e max cos f) over THAT interval gives the
// final one.  S <= 0: the body never transits (empty window).  The true-anomaly interval is
// mapped to mean anomaly; `widen` (relative, on (1 + r) R*) and `margin` (cycles) cover
// rounding (and the fp32 mode's coarser b).  Anything not provably narrow is flagged
// BODY_ALWAYS (no skipping).
template <typename T>
__device__ inline void transit_window(BodyC<T>& c, double a, double rstar, double e,
                                      double widen, double margin) {
  c.flags |= BODY_ALWAYS;
  c.phase_c = 0.0;
  c.wlo = -1.0;
  c.whi = 1.0;
  c.win_in = 0.0;
  const double sw = (double)c.sw, cw = (double)c.cw;
  const double ci = (double)c.ci, si = (double)c.si;
  const double rho = fsqrt(sw * sw + cw * cw);
  if (!(e >= 0.0) || !(e < 0.99) || !(rho > 0.0) || !(si * si > 1e-6)) return;
  const double reach = (1.0 + fabs((double)c.rr)) * rstar * widen;  // (1 + r) R*
  const double omega = fatan2(sw, cw);
  const double theta_c = (si >= 0.0) ? Num<double>::half_pi : 3.0 * Num<double>::half_pi;
  const double fc = theta_c - omega;
  const double inv_si2 = fdiv(1.0, si * si);
  double sfc, cfc;
  fsincos(fc, &sfc, &cfc);
  double rmin = rho * a * (1.0 - e);
  double delta = 0.0, sd = 0.0, cd = 1.0;
  for (int pass = 0; pass < 2; ++pass) {
    const double q = fdiv(reach, rmin);
    const double S = (q * q - ci * ci) * inv_si2;
    if (!(S < 0.9)) return;  // wide (or NaN): evaluate everything
    if (S <= 0.0) {          // never in transit: empty window
      if (!isfinite(c.inv_period) || !isfinite(c.t0ref)) return;
      c.wlo = 1.0;
      c.whi = -1.0;
      c.flags &= ~BODY_ALWAYS;
      return;
    }
    sd = fsqrt(S);                              // sin(delta)
    cd = fsqrt(1.0 - S);                        // cos(delta), delta in (0, 1.25)
    delta = fatan2(sd, cd);                     // asin(sqrt(S))
    if (pass == 0) {
      const double fcr = fc - Num<double>::two_pi * rint(fc * Num<double>::inv_two_pi);
      // max of cos(fc - delta), cos(fc + delta) = cos fc cos delta + |sin fc| sin delta
      const double cmax = (fabs(fcr) <= delta) ? 1.0 : fma(cfc, cd, fabs(sfc) * sd);
      rmin = fdiv(rho * a * (1.0 - e * e), 1.0 + e * cmax) * (1.0 - 1e-12);
    }
  }
  const double sq1me = fsqrt(1.0 - e), sq1pe = fsqrt(1.0 + e);
  const double Mc = mean_from_true(fc, sq1me, sq1pe, e);
  double lo = mean_from_true(fc - delta, sq1me, sq1pe, e) - Mc;
  double hi = mean_from_true(fc + delta, sq1me, sq1pe, e) - Mc;
  // (the differences are multiples of 2 pi away from the wanted branch by at most a few turns)
  for (int it = 0; it < 8 && lo > 0.0; ++it) lo -= Num<double>::two_pi;
  for (int it = 0; it < 8 && hi < 0.0; ++it) hi += Num<double>::two_pi;
  lo = lo * Num<double>::inv_two_pi - margin;
  hi = hi * Num<double>::inv_two_pi + margin;
  if (!(lo > -0.45) || !(hi < 0.45) || !(lo <= 0.0) || !(hi >= 0.0)) return;
  if (!isfinite(Mc) || !isfinite(c.inv_period) || !isfinite(c.t0ref)) return;
  c.phase_c = Mc * Num<double>::inv_two_pi;
  c.wlo = lo;
  c.whi = hi;
  c.flags &= ~BODY_ALWAYS;
  // Fully-inside part (b < 1 - r) around mid-transit, estimated with straight-line motion at the
  // conjunction speed: sqrt((1 - r)^2 - b0^2) / v, v = 2 pi (a / R*) (1 + e cos f_c) / sqrt(1 - e^2)
  // stellar radii per cycle.  The list path uses it to put those samples on a list of their own so
  // that whole warps take the constant-node flux code; every warp still checks b < 1 - r itself.
  {
    const double opec = 1.0 + e * cfc;
    const double ome2 = 1.0 - e * e;
    const double r0c = fdiv(rho * a * ome2, opec);
    const double b0 = fdiv(r0c * fabs(ci), rstar);
    const double rr = fabs((double)c.rr);
    const double v = fdiv(Num<double>::two_pi * fdiv(rho * a, rstar) * opec, fsqrt(ome2));
    const double q2 = (1.0 - rr) * (1.0 - rr) - b0 * b0;
    if (rr < 1.0 && b0 < 1.0 - rr && q2 > 0.0 && v > 0.0 && isfinite(v)) {
#ifndef JXP_WIN_IN_FACTOR
#define JXP_WIN_IN_FACTOR 0.97
#endif
      const double half = fdiv(JXP_WIN_IN_FACTOR * fsqrt(q2), v);
      if (half > 0.0 && half < hi && -half > lo) c.win_in = half;
    }
  }
}

// true when the sample at time tm cannot be in transit
template <typename T>
__device__ __forceinline__ bool outside_window(const BodyC<T>& c, double tm) {
  const double ph = fma(tm - c.t0ref, c.inv_period, -c.phase_c);
  const double d = ph - rint(ph);
  return (d < c.wlo) | (d > c.whi);
}

// Mean anomaly reduced to [0, 2pi), same operation order as keplerian.py:538 and
// kepler.py:31 (always in double).
template <typename T>
__device__ __forceinline__ double mean_anomaly_reduced(const BodyC<T>& c, double tm) {
  const double x = (tm - c.t0) - c.tref;
  const double M = (Num<double>::two_pi * x) / c.period;
  return mod_two_pi(M);
}

// Sky-projected separation (in R*) and the z > 0 test for one sample.  S is T or Dual<T, N>.
template <typename S, typename B>
__device__ __forceinline__ void sky_position_g(const S& e, const S& cw, const S& sw, const S& ci,
                                               const S& si, const S& na_ome2, B inv_rstar,
                                               const S& sf, const S& cf, S& b, S& Z) {
  const S r0 = jdiv(na_ome2, e * cf + B(1));  // keplerian.py:630
  const S x0 = r0 * cf, y0 = r0 * sf;
  const S x1 = cw * x0 - sw * y0;  // :568-569
  const S y1 = sw * x0 + cw * y0;
  const S Y = ci * y1;  // :575-576
  Z = -(si * y1);
  b = jsqrt(x1 * x1 + Y * Y) * inv_rstar;  // light_curves/limb_dark.py:53
}

// The same from (yD, xD) = (sin f, cos f) D of kepler_reduced_n<N, true>: (x0, y0) = -a (xD, yD).
template <typename T>
__device__ __forceinline__ void sky_position_xy(const BodyC<T>& c, double yD, double xD, double& b, double& Z) {
  const double x0 = c.na * xD, y0 = c.na * yD;
  const double x1 = jfma(c.cw, x0, -c.sw * y0);  // keplerian.py:568-569
  const double y1 = jfma(c.sw, x0, c.cw * y0);
  const double Y = c.ci * y1;                    // :575-576
  Z = -c.si * y1;
  b = jsqrt(jfma(x1, x1, Y * Y)) * c.inv_rstar;  // light_curves/limb_dark.py:53
}

template <typename T>
__device__ __forceinline__ void sky_position(const BodyC<T>& c, double sf, double cf, double& b, double& Z) {
  const double r0 = jdiv(c.na_ome2, jfma(c.e, cf, 1.0));  // keplerian.py:630
  const double x0 = r0 * cf, y0 = r0 * sf;
  const double x1 = jfma(c.cw, x0, -c.sw * y0);  // :568-569
  const double y1 = jfma(c.sw, x0, c.cw * y0);
  const double Y = c.ci * y1;                    // :575-576
  Z = -c.si * y1;
  b = jsqrt(jfma(x1, x1, Y * Y)) * c.inv_rstar;  // light_curves/limb_dark.py:53
}

// The position stage of one sample: t -> M -> (sin f, cos f) -> (b, Z).  Always double (see BodyC); the fp32
// mode rounds b and Z once, at the end.
template <typename T>
__device__ __forceinline__ void position_stage(const BodyC<T>& c, double tm, T& b, T& Z) {
  const double Mr = mean_anomaly_reduced(c, tm);
  double sf, cf, Edummy;
  if (c.flags & BODY_CIRCULAR) {
    jsincos(Mr, &sf, &cf);  // keplerian.py:539-540
  } else {
    kepler_reduced<double, false>(Mr, c.e, c.ome, c.ope, c.s1me2, c.inv_ope, sf, cf, Edummy);
  }
  double bd, Zd;
  sky_position(c, sf, cf, bd, Zd);
  b = T(bd);
  Z = T(Zd);
}

// Register-array element by a runtime index (select chain; keeps the array in registers).
template <typename V, int K>
__device__ __forceinline__ V pick(const V (&a)[K], int k) {
  V r = a[0];
#pragma unroll
  for (int i = 1; i < K; ++i) r = (k == i) ? a[i] : r;
  return r;
}

// flux - 1 of one body at one time (0 when not in transit); the caller has already applied
// the transit-window pre-test.
template <typename T, int ORDER, bool HI = false, int PU = 1>
__device__ __forceinline__ T eval_in_window(const BodyC<T>& c, const SetC<T>& sc,
                                            const GLTable* __restrict__ tab, double tm,
                                            unsigned& n_ld) {
  T b, Z;
  position_stage(c, tm, b, Z);
  if (!(Z > T(0)) || (b >= c.onepr)) return T(0);  // limb_dark.py:58 mask, :60 no_occ
  ++n_ld;
  T s0, s1, s2;
  if (HI) {
    T shi[6];
    ld_solution<T, ORDER, true>(b, c.rr, sc.lmax, tab, s0, s1, s2, shi);
    T f = ld_combine(2, s0, s1, s2, sc.g0, sc.g1, sc.g2);
#pragma unroll
    for (int n = 3; n <= 8; ++n)
      if (n <= sc.lmax) f = jfma(shi[n - 3], sc.ghi[n - 3], f);
    return f;
  }
  ld_solution<T, ORDER, false, PU>(b, c.rr, sc.lmax, tab, s0, s1, s2);
  return ld_combine(sc.lmax, s0, s1, s2, sc.g0, sc.g1, sc.g2);
}

// Per-thread view of the K consecutive time samples a thread owns: used to reject all K
// against a body's transit window with one test when the samples are time-ordered.
struct SpanK {
  double first;  // t of the first valid sample
  double span;   // t_last - t_first (>= 0 when monotone)
  bool mono;     // samples non-decreasing
};

// Candidate mask (bit k set: sample k may be in transit) for one (body, sub-sample offset).
template <typename T, int K>
__device__ __forceinline__ unsigned window_mask(const BodyC<T>& c, const double (&tk)[K],
                                                const SpanK& sp, double dtj, unsigned valid,
                                                bool use_window) {
  if (!use_window || (c.flags & BODY_ALWAYS)) return valid;
  const double t0ref = c.t0ref, invP = c.inv_period, phc = c.phase_c, wlo = c.wlo, whi = c.whi;
  if (sp.mono) {
    const double ph0 = fma((sp.first + dtj) - t0ref, invP, -phc);
    const double d0 = ph0 - rint(ph0);
    const double d1 = fma(sp.span, invP, d0);
    // whole interval before this cycle's window, or between this window and the next one
    if ((d1 < wlo) | ((d0 > whi) & (d1 < 1.0 + wlo))) return 0u;
  }
  unsigned mask = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double ph = fma((tk[k] + dtj) - t0ref, invP, -phc);
    const double d = ph - rint(ph);
    mask |= ((d >= wlo) & (d <= whi)) ? (1u << k) : 0u;
  }
  return mask & valid;
}

}  // namespace jxp
