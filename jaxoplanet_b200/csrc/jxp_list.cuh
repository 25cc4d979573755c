// Transit list: the in-window samples of single-planet parameter sets as an explicit work list
// (DESIGN.md "transit list").  Shared by the log-likelihood path (jxp_kernels.cu) and the
// log-likelihood + gradient path (jxp_partials.cu).
#pragma once

#include "jxp_host.h"
#include "jxp_orbit.cuh"

namespace jxp {

// tl_state[1] of the transit-list path (tl_state[0], [2] = items on the inside / edge list); zeroed by
// k_sys_prep.  The two lists share one buffer (inside from the bottom, edge from the top).
constexpr unsigned long long TL_UNSORTED = 1ull, TL_OVERFLOW = 2ull;
__device__ __forceinline__ bool tl_active(const unsigned long long* __restrict__ st, unsigned long long cap) {
  return st[1] == 0ull && st[0] + st[2] <= cap;
}

struct TlBuildArgs {
  const unsigned char* bodies;  // records that start with a BodyC<double>
  size_t body_stride;           // bytes from one record to the next
  int n_sets;
  const double* t;              // sorted time stamps
  int64_t n_times;
  const double* sorted_flags;   // [64] per-CTA "not sorted" flags of the likelihood prep
  unsigned long long* tl_state;  // [0] items on the inside list, [1] TL_* flags, [2] items on the edge list
  unsigned long long* tl_items;  // [tl_cap] (set << 32 | sample): inside list from 0 up, edge list from the top down
  unsigned long long* tl_off;    // [n_sets][2] offset of the set's slice in the inside / edge list
  unsigned* tl_cnt;              // [n_sets][2] items of the set on the inside / edge list
  unsigned long long tl_cap;
};

// First index i in [0, n] with !pred(t[i]), pred(v) = (v < x) (lower bound) or (v <= x) (upper
// bound), on sorted t.  The estimate from a uniform grid through the end points is bracketed by
// doubling, then bisected: ~4 dependent loads on a regular cadence instead of log2(n).
template <bool UPPER>
__device__ __forceinline__ unsigned tl_search(const double* __restrict__ t, unsigned n, double x,
                                              double t_first, double inv_dt) {
  auto pred = [&](unsigned i) {
    const double v = t[i];
    return UPPER ? (v <= x) : (v < x);
  };
  double gd = (x - t_first) * inv_dt;
  gd = fmin(fmax(gd, 0.0), (double)n);
  const unsigned g = (unsigned)gd;
  unsigned w = 2;
  unsigned lo = g > w ? g - w : 0u;
  while (lo > 0u && !pred(lo - 1u)) {
    w = (w < 0x40000000u) ? (w << 1) : 0xffffffffu;
    lo = g > w ? g - w : 0u;
  }
  w = 2;
  unsigned hi = (n - g > w) ? g + w : n;
  while (hi < n && pred(hi)) {
    w = (w < 0x40000000u) ? (w << 1) : 0xffffffffu;
    hi = (n - g > w) ? g + w : n;
  }
  while (lo < hi) {
    const unsigned mid = lo + ((hi - lo) >> 1);
    if (pred(mid)) lo = mid + 1u; else hi = mid;
  }
  return lo;
}

// TAG: one instantiation per translation unit (the kernel lives in a header).
template <int TAG>
__global__ void __launch_bounds__(128) k_tl_build(const __grid_constant__ TlBuildArgs a) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= a.n_sets) return;
  {  // sortedness flags of k_sys_prep_ll (one per CTA of the likelihood prep)
    const bool bad = a.sorted_flags[lane] != 0.0 || a.sorted_flags[32 + lane] != 0.0;
    if (__any_sync(0xffffffffu, bad)) {
      if (s == 0 && lane == 0) atomicOr(a.tl_state + 1, TL_UNSORTED);
      return;
    }
  }
  const BodyC<double>& c = *reinterpret_cast<const BodyC<double>*>(a.bodies + (size_t)s * a.body_stride);
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
  // transit k -> sample range [lo, hi) in the window and the part [ilo, ihi) of it estimated to have the
  // planet fully inside the disk (at least 8 samples, else empty)
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
  // the first 64 transits stay in registers between the counting and the writing pass
  unsigned lo0, hi0, il0, ih0, lo1, hi1, il1, ih1;
  interval(lane, lo0, hi0, il0, ih0);
  interval(32 + lane, lo1, hi1, il1, ih1);
  unsigned long long cntA = (unsigned long long)(ih0 - il0) + (ih1 - il1);
  unsigned long long cntB = (unsigned long long)(hi0 - lo0) + (hi1 - lo1) - cntA;
  for (long long k0 = 64; k0 < n_tr; k0 += 32) {
    unsigned lo, hi, il, ih;
    interval(k0 + lane, lo, hi, il, ih);
    cntA += ih - il;
    cntB += (hi - lo) - (ih - il);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cntA += __shfl_xor_sync(0xffffffffu, cntA, o);
    cntB += __shfl_xor_sync(0xffffffffu, cntB, o);
  }
  unsigned long long offA = 0, offB = 0;
  if (lane == 0) {
    offA = atomicAdd(a.tl_state, cntA);
    offB = atomicAdd(a.tl_state + 2, cntB);
    a.tl_off[2 * s] = offA;
    a.tl_off[2 * s + 1] = offB;
    a.tl_cnt[2 * s] = (unsigned)cntA;
    a.tl_cnt[2 * s + 1] = (unsigned)cntB;
  }
  offA = __shfl_sync(0xffffffffu, offA, 0);
  offB = __shfl_sync(0xffffffffu, offB, 0);
  // no write may leave the buffer; whether the two lists ran into each other is only known when every
  // set has allocated: tl_active() checks the totals in the kernels that follow
  if (offA + cntA > a.tl_cap || offB + cntB > a.tl_cap) {
    if (lane == 0) atomicOr(a.tl_state + 1, TL_OVERFLOW);
    return;
  }
  unsigned long long* outA = a.tl_items + offA;
  unsigned long long* outB = a.tl_items + (a.tl_cap - offB - cntB);
  const unsigned long long tag = (unsigned long long)s << 32;
  for (long long k0 = 0; k0 < n_tr; k0 += 32) {
    unsigned lo, hi, il, ih;
    if (k0 == 0) {
      lo = lo0; hi = hi0; il = il0; ih = ih0;
    } else if (k0 == 32) {
      lo = lo1; hi = hi1; il = il1; ih = ih1;
    } else {
      interval(k0 + lane, lo, hi, il, ih);
    }
    const unsigned cA = ih - il, cB = (hi - lo) - cA;
    unsigned inA = cA, inB = cB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned vA = __shfl_up_sync(0xffffffffu, inA, o);
      const unsigned vB = __shfl_up_sync(0xffffffffu, inB, o);
      if (lane >= o) {
        inA += vA;
        inB += vB;
      }
    }
    const unsigned exA = inA - cA, exB = inB - cB;
    unsigned nz = __ballot_sync(0xffffffffu, hi != lo);
    while (nz) {
      const int j = __ffs(nz) - 1;
      nz &= nz - 1;
      const unsigned lj = __shfl_sync(0xffffffffu, lo, j), hj = __shfl_sync(0xffffffffu, hi, j);
      const unsigned ilj = __shfl_sync(0xffffffffu, il, j), ihj = __shfl_sync(0xffffffffu, ih, j);
      const unsigned eAj = __shfl_sync(0xffffffffu, exA, j), eBj = __shfl_sync(0xffffffffu, exB, j);
      const unsigned nAj = ihj - ilj, nB0 = ilj - lj, nBj = (hj - lj) - nAj;
      for (unsigned i = lane; i < nAj; i += 32) outA[eAj + i] = tag | (unsigned long long)(ilj + i);
      for (unsigned i = lane; i < nBj; i += 32)
        outB[eBj + i] = tag | (unsigned long long)(i < nB0 ? lj + i : ihj + (i - nB0));
    }
    outA += __shfl_sync(0xffffffffu, inA, 31);
    outB += __shfl_sync(0xffffffffu, inB, 31);
  }
}


// One slice of per-item terms added in a fixed order: 8 independent lane-strided sums, software-pipelined so
// that the loads of round r + 1 are in flight while round r is added; returns this lane's share.
__device__ __forceinline__ double tl_slice_sum(const double* __restrict__ d, unsigned cnt, int lane) {
  double acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.0;
  const unsigned full = cnt & ~255u;
  unsigned i = lane;
  if (i < full) {
    double v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = __ldcs(d + i + 32 * m);
#pragma unroll 1
    for (i += 256; i < full; i += 256) {
      double w[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) w[m] = __ldcs(d + i + 32 * m);
#pragma unroll
      for (int m = 0; m < 8; ++m) acc[m] += v[m];
#pragma unroll
      for (int m = 0; m < 8; ++m) v[m] = w[m];
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] += v[m];
  }
#pragma unroll
  for (int m = 0; m < 8; ++m)
    if (i + 32 * m < cnt) acc[m] += d[i + 32 * m];  // i + 224 < full + 256: no wrap for cnt < 2^32 - 256
  return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

}  // namespace jxp
