// Per-body constants and the per-time-sample evaluation of the fused light-curve path:
//   t -> mean anomaly -> Kepler -> sky position -> b -> limb-darkened flux
// src/jaxoplanet/orbits/keplerian.py:262-340, 537-637; light_curves/limb_dark.py:49-60.
#pragma once

#include "jxp_math.cuh"

namespace jxp {

enum : int { BODY_CIRCULAR = 1, BODY_ALWAYS = 2 };

// Device-side record of one (parameter set, body), produced once per call by the prep
// kernel and broadcast to a CTA through shared memory.  Time-like fields stay double in
// the fp32 mode (DESIGN.md "fp32 mode").
template <typename T>
struct BodyC {
  double t0;          // time_transit
  double tref;        // time_ref
  double period;
  double t0ref;       // t0 + tref               } approximate phase for the
  double inv_period;  // 1 / period              } transit-window pre-test
  double phase_c;     // window centre [cycles]  }
  double wlo, whi;    // window edges relative to the centre [cycles], margins included
  T e, ome, ope, s1me2, inv_ope;
  T cw, sw, ci, si;
  T na_ome2;  // -a (1 - e^2)
  T inv_rstar, rr, onepr;
  int flags;
  int pad;
};

template <typename T>
struct SetC {
  T g0, g1, g2;  // normalised Green's coefficients
  int lmax;
};

// Mean anomaly at true anomaly f (monotone map used only for the window edges).
__device__ inline double mean_from_true(double f, double e) {
  double sh, ch;
  sincos(0.5 * f, &sh, &ch);
  double E = 2.0 * atan2(sqrt(1.0 - e) * sh, sqrt(1.0 + e) * ch);
  return E - e * sin(E);
}

// Conservative transit window in orbital phase.  A sample can only be in transit
// (z > 0 and b < 1 + r) if |cos(f + omega)| < (1 + r) R* / (a (1 - e)), because
// b R* >= |X| = |r0 cos(f + omega)| and |r0| >= a (1 - e).  The true-anomaly interval is
// mapped to mean anomaly and widened by `widen` (relative, on the bound) and `margin`
// (cycles).  Anything not provably narrow is flagged BODY_ALWAYS (no skipping).
template <typename T>
__device__ inline void transit_window(BodyC<T>& c, double a, double rstar, double e,
                                      double widen, double margin) {
  c.flags |= BODY_ALWAYS;
  c.phase_c = 0.0;
  c.wlo = -1.0;
  c.whi = 1.0;
  const double sw = (double)c.sw, cw = (double)c.cw;
  const double rho = sqrt(sw * sw + cw * cw);
  const double bound = (1.0 + fabs((double)c.rr)) * rstar / (a * (1.0 - e) * rho) * widen;
  if (!(bound < 0.95) || !(bound > 0.0)) return;
  if (!(e >= 0.0) || !(e < 0.99)) return;
  const double delta = asin(bound);
  const double omega = atan2(sw, cw);
  const double theta_c = ((double)c.si >= 0.0) ? Num<double>::half_pi : 3.0 * Num<double>::half_pi;
  const double fc = theta_c - omega;
  const double Mc = mean_from_true(fc, e);
  double lo = mean_from_true(fc - delta, e) - Mc;
  double hi = mean_from_true(fc + delta, e) - Mc;
  while (lo > 0.0) lo -= Num<double>::two_pi;
  while (hi < 0.0) hi += Num<double>::two_pi;
  lo = lo * Num<double>::inv_two_pi - margin;
  hi = hi * Num<double>::inv_two_pi + margin;
  if (!(lo > -0.45) || !(hi < 0.45)) return;
  if (!isfinite(Mc) || !isfinite(c.inv_period) || !isfinite(c.t0ref)) return;
  c.phase_c = Mc * Num<double>::inv_two_pi;
  c.wlo = lo;
  c.whi = hi;
  c.flags &= ~BODY_ALWAYS;
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

template <typename T>
__device__ __forceinline__ void sky_position(const BodyC<T>& c, T sf, T cf, T& b, T& Z) {
  const T r0 = jdiv(c.na_ome2, jfma(c.e, cf, T(1)));  // keplerian.py:630
  const T x0 = r0 * cf, y0 = r0 * sf;
  const T x1 = jfma(c.cw, x0, -c.sw * y0);       // :568-569
  const T y1 = jfma(c.sw, x0, c.cw * y0);
  const T Y = c.ci * y1;                         // :575-576
  Z = -c.si * y1;
  b = jsqrt(jfma(x1, x1, Y * Y)) * c.inv_rstar;  // light_curves/limb_dark.py:53
}

// Register-array element by a runtime index (select chain; keeps the array in registers).
template <typename V, int K>
__device__ __forceinline__ V pick(const V (&a)[K], int k) {
  V r = a[0];
#pragma unroll
  for (int i = 1; i < K; ++i) r = (k == i) ? a[i] : r;
  return r;
}

// flux - 1 of one body at one time (0 when not in transit).
template <typename T, int ORDER>
__device__ __forceinline__ T eval_sample(const BodyC<T>& c, const SetC<T>& sc,
                                         const GLTable* __restrict__ tab, double tm,
                                         bool use_window) {
  if (use_window && !(c.flags & BODY_ALWAYS) && outside_window(c, tm)) return T(0);
  const T Mr = T(mean_anomaly_reduced(c, tm));
  T sf, cf, Edummy;
  if (c.flags & BODY_CIRCULAR) {
    jsincos(Mr, &sf, &cf);  // keplerian.py:539-540
  } else {
    kepler_reduced<T, false>(Mr, c.e, c.ome, c.ope, c.s1me2, c.inv_ope, sf, cf, Edummy);
  }
  T b, Z;
  sky_position(c, sf, cf, b, Z);
  if (!(Z > T(0)) || (b >= c.onepr)) return T(0);  // limb_dark.py:58 mask, :60 no_occ
  T s0, s1, s2;
  ld_solution<T, ORDER>(b, c.rr, sc.lmax, tab, s0, s1, s2);
  return ld_combine(sc.lmax, s0, s1, s2, sc.g0, sc.g1, sc.g2);
}

}  // namespace jxp
