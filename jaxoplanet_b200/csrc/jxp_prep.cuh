// Per-call preparation shared by the translation units: per-(set, body) derived constants
// (OrbitalBody.__init__, src/jaxoplanet/orbits/keplerian.py:262-340 -> BodyC / SetC) and the
// set-independent terms of the Gaussian log-likelihood.
#pragma once

#include "../../include/jaxoplanet_b200.h"
#include "jxp_host.h"
#include "jxp_orbit.cuh"

namespace jxp {
using jxph::LL_NCTA_H;
constexpr int LL_NCTA = jxph::LL_NCTA_H;
constexpr int LL_SCAL0 = jxph::LL_SCAL0_H;
constexpr int LL_FLAG0 = jxph::LL_FLAG0_H;  // per-CTA "time stamps not sorted" flags (list path)

// one thread per (set, body): derived constants -> BodyC (+ transit window), u -> SetC
template <typename T>
__device__ __noinline__ void sys_prep_body(int i, const double* __restrict__ bodies, int n_sets, int n_bodies,
                           const double* __restrict__ u, int n_u, int64_t u_stride,
                           BodyC<T>* __restrict__ out_b, SetC<T>* __restrict__ out_s,
                           unsigned long long* __restrict__ counters, int* __restrict__ tile_ctr,
                           double widen, double margin) {
  if (i >= n_sets * n_bodies) return;
  if (i <= n_sets) tile_ctr[i] = 0;  // work distribution state of k_system (n_sets + 1 entries)
  if (i == 0 && n_sets * n_bodies <= n_sets) tile_ctr[n_sets] = 0;
  const double* p = bodies + (size_t)i * JXP_BODY_NFIELDS;
  BodyC<T> c;
  c.t0 = p[JXP_BODY_TIME_TRANSIT];
  c.tref = p[JXP_BODY_TIME_REF];
  c.period = p[JXP_BODY_PERIOD];
  c.t0ref = c.t0 + c.tref;
  c.inv_period = 1.0 / c.period;
  double e = p[JXP_BODY_ECC];
  c.flags = 0;
  c.pad = 0;
  if (e < 0.0) {
    c.flags |= BODY_CIRCULAR;
    e = 0.0;
  }
  const double a = p[JXP_BODY_SEMIMAJOR];
  const double rstar = p[JXP_BODY_CENTRAL_RADIUS];
  c.e = e;
  c.ome = 1.0 - e;
  c.ope = 1.0 + e;
  c.s1me2 = sqrt((1.0 - e) * (1.0 + e));
  c.inv_ope = 1.0 / (1.0 + e);
  c.cw = p[JXP_BODY_COS_OMEGA];
  c.sw = p[JXP_BODY_SIN_OMEGA];
  c.ci = p[JXP_BODY_COS_INCL];
  c.si = p[JXP_BODY_SIN_INCL];
  c.na_ome2 = -a * (1.0 - e * e);
  c.na = -a;
  c.inv_rstar = 1.0 / rstar;
  const double rr = p[JXP_BODY_RADIUS] / rstar;
  c.rr = T(rr);
  c.onepr = T(1.0 + fabs(rr));
  transit_window(c, a, rstar, e, widen, margin);
  out_b[i] = c;
  if (i == 0) {  // executed-work counters of this call (jxp_system_stats)
    counters[0] = 0ull;
    counters[1] = 0ull;
    counters[2] = 0ull;  // transit list: items on the inside list
    counters[3] = 0ull;  // transit list: TL_UNSORTED | TL_OVERFLOW (non-zero: list path off)
    counters[4] = 0ull;  // transit list: items on the edge list
    counters[5] = 0ull;  // transit list: chunks evaluated with the inside-the-disk flux code
  }
  if (i % n_bodies == 0) {
    const int s = i / n_bodies;
    const double* us = u + (size_t)s * u_stride;
    double gg[JXP_MAX_LD_COEFFS + 1], uu[JXP_MAX_LD_COEFFS];
    for (int k = 0; k <= JXP_MAX_LD_COEFFS; ++k) gg[k] = 0.0;
    for (int k = 0; k < JXP_MAX_LD_COEFFS; ++k) uu[k] = (k < n_u) ? us[k] : 0.0;
    if (n_u > 2)
      ld_greens_n<double>(n_u, uu, gg);
    else
      ld_greens<double>(n_u, uu[0], uu[1], gg[0], gg[1], gg[2]);
    SetC<T> sc;
    sc.g0 = T(gg[0]);
    sc.g1 = T(gg[1]);
    sc.g2 = T(gg[2]);
    for (int k = 0; k < 6; ++k) sc.ghi[k] = T(gg[3 + k]);
    sc.lmax = n_u;
    out_s[s] = sc;
  }
}

// winv[t] = 1 / sigma[t]; per-CTA partial sums of ((y - 1) winv)^2 and of
// (-log sigma - log(2 pi) / 2) into scalars[LL_SCAL0 + 2 cta .. ]; LL_NCTA CTAs over contiguous
// chunks, fixed summation order (k_loglike_final / k_tr_final add the LL_NCTA pairs in order).
// CTA `cta` of `ncta` (256 threads); t != nullptr: also flag time stamps that are not sorted
template <typename T>
__device__ __forceinline__ void
loglike_prep_body(int cta, int ncta, const T* __restrict__ y, const T* __restrict__ sigma, int64_t ss, int64_t n,
                  T* __restrict__ winv, double* __restrict__ scalars, const double* __restrict__ t,
                  double2* __restrict__ aw = nullptr) {
  __shared__ double sh0[8], sh1[8];
  double a0 = 0.0, a1 = 0.0;
  int bad = 0;
  const double half_log_2pi = 0.918938533204672741780329736405617639;
  const int64_t chunk = (n + ncta - 1) / ncta;
  const int64_t lo = (int64_t)cta * chunk;
  const int64_t hi = lo + chunk < n ? lo + chunk : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const T sg = sigma[i * ss];
    const T w = T(1) / sg;
    winv[i] = w;
    const double r0 = (double)((y[i] - T(1)) * w);
    // (2 r0 / sigma, 1 / sigma^2): the chi^2 term of a model value m is m (m / sigma^2 - 2 r0 / sigma) (k_walk_items)
    if (aw != nullptr) aw[i] = make_double2(2.0 * r0 * (double)w, (double)w * (double)w);
    a0 += r0 * r0;
    a1 -= log((double)sg) + half_log_2pi;
    // the transit-list path needs non-decreasing, NaN-free time stamps
    if (t != nullptr && i + 1 < n && !(t[i + 1] >= t[i])) bad = 1;
  }
  if (t != nullptr) {
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) scalars[LL_FLAG0 + cta] = bad ? 1.0 : 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_down_sync(0xffffffffu, a0, o);
    a1 += __shfl_down_sync(0xffffffffu, a1, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh0[w] = a0;
    sh1[w] = a1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      t0 += sh0[k];
      t1 += sh1[k];
    }
    scalars[LL_SCAL0 + 2 * cta] = t0;
    scalars[LL_SCAL0 + 2 * cta + 1] = t1;
  }
}

}  // namespace jxp
