// Branch-free FP64 primitives for the hot loops: reciprocal, division, square root and
// sine/cosine on a bounded range.  The CUDA library versions carry slow-path calls
// (denormals, huge arguments) that are never taken here but cost instruction-cache
// footprint and registers; the fused kernels are instruction-fetch bound when those are
// inlined ten times per sample (profiles/r1_c3_first.md).  Each routine is a hardware seed
// (MUFU.RCP64H / MUFU.RSQ64H, ~2^-20) refined with FMAs to <= 1 ulp.
//
// Domain: finite, normal-range inputs (what the hot path produces); zeros are handled where
// the formulas can produce them (sqrt(0) = 0).  NaN in -> NaN out.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace jxp {

__device__ __forceinline__ double rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

// 1 / x, ~1 ulp
__device__ __forceinline__ double frcp(double x) {
  double y = rcp_seed(x);
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);  // third-order step: y (1 + e + e^2)
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// a / b, <= 1 ulp (residual-corrected)
__device__ __forceinline__ double fdiv(double a, double b) {
  double y = rcp_seed(b);
  double e = fma(-b, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  double q = a * y;
  double r = fma(-b, q, a);
  return fma(r, y, q);
}

// sqrt(x) for x >= 0 (0 -> 0, tiny/denormal -> 0), <= 1 ulp
__device__ __forceinline__ double fsqrt(double x) {
  double y = rsqrt_seed(x);
  double g = x * y;
  double h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  double p = fma(r, 1.5, 1.0) * r;  // r + 1.5 r^2
  g = fma(g, p, g);
  h = fma(h, p, h);
  double d = fma(-g, g, x);
  g = fma(d, h, g);
  return (x > 2.3e-308) ? g : ((x >= 0.0) ? 0.0 : nan(""));
}

// sqrt(x) for x known to be positive and normal (no zero / negative handling)
__device__ __forceinline__ double fsqrt_pos(double x) {
  double y = rsqrt_seed(x);
  double g = x * y;
  double h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  double p = fma(r, 1.5, 1.0) * r;
  g = fma(g, p, g);
  h = fma(h, p, h);
  double d = fma(-g, g, x);
  return fma(d, h, g);
}

__device__ __constant__ double c_sin[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03,
                                           -1.98412698298579493134e-04, 2.75573137070700676789e-06,
                                           -2.50507602534068634195e-08, 1.58969099521155010221e-10};
__device__ __constant__ double c_cosk[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03,
                                            2.48015872894767294178e-05, -2.75573143513906633035e-07,
                                            2.08757232129817482790e-09, -1.13596475577881948265e-11};

// sin and cos for |x| <~ 1e4: two-term Cody-Waite reduction by pi/2 and the classic
// minimax kernels on [-pi/4, pi/4] (coefficients: Sun fdlibm k_sin.c / k_cos.c, public domain).
__device__ __forceinline__ void fsincos(double x, double* sp, double* cp) {
  const double two_over_pi = 6.36619772367581382433e-01;
  const double pio2_hi = 1.57079632679489655800e+00;
  const double pio2_lo = 6.12323399573676603587e-17;
  const double magic = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, two_over_pi, magic);
  const int q = __double2loint(t);
  const double qd = t - magic;
  double r = fma(-qd, pio2_hi, x);
  r = fma(-qd, pio2_lo, r);
  const double z = r * r;
  double ps = c_sin[5];
#pragma unroll
  for (int k = 4; k >= 0; --k) ps = fma(ps, z, c_sin[k]);
  const double s = fma(r * z, ps, r);
  double pc = c_cosk[5];
#pragma unroll
  for (int k = 4; k >= 0; --k) pc = fma(pc, z, c_cosk[k]);
  const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
  const bool swap = q & 1;
  double so = swap ? c : s;
  double co = swap ? s : c;
  if (q & 2) so = -so;
  if ((q + 1) & 2) co = -co;
  *sp = so;
  *cp = co;
}

__device__ __forceinline__ double fcos(double x) {
  double s, c;
  fsincos(x, &s, &c);
  return c;
}

// Taylor coefficients (-1)^k / (2k)! of cos, in the constant bank: a DFMA takes them as a direct
// operand, whereas 64-bit immediates cost two extra moves each under register pressure.
__device__ __constant__ double c_cos0pi[15] = {
    1.0,
    -0.5,
    0.041666666666666664,
    -0.001388888888888889,
    2.48015873015873e-05,
    -2.755731922398589e-07,
    2.08767569878681e-09,
    -1.1470745597729725e-11,
    4.779477332387385e-14,
    -1.5619206968586225e-16,
    4.110317623312165e-19,
    -8.896791392450574e-22,
    1.6117375710961184e-24,
    -2.4795962632247976e-27,
    3.279889237069838e-30};

// cos(x) for 0 <= x <= pi (the quadrature node angles): even Taylor polynomial in x^2 through
// x^28 (remainder < 3e-18 at pi), no range reduction and no selects; absolute error ~1e-15
// at x ~ pi from the alternating partial sums, far below what the flux needs (1e-12).
__device__ __forceinline__ double fcos_0pi(double x) {
  // even/odd split in z = x^2: two independent Horner chains of depth 7 instead of one of 14
  const double z = x * x;
  const double z2 = z * z;
  double pa = c_cos0pi[14];
  double pb = c_cos0pi[13];
#pragma unroll
  for (int k = 12; k >= 0; k -= 2) pa = fma(pa, z2, c_cos0pi[k]);
#pragma unroll
  for (int k = 11; k >= 1; k -= 2) pb = fma(pb, z2, c_cos0pi[k]);
  return fma(pb, z, pa);
}

// cos(a) = Pc(a^2) and sin(a) = a Ps(a^2) for |a| <= pi/2 + 1e-6: degree-8 near-minimax fits
// (Chebyshev interpolation in mpmath; truncation 2.5e-18 / 2e-19, evaluated error <= 2.3e-16).
__device__ __constant__ double c_pc[9] = {1.0,
                                          -0.4999999999999997,
                                          0.04166666666666388,
                                          -0.0013888888888772988,
                                          2.480158727741477e-05,
                                          -2.755731639101142e-07,
                                          2.087656183919406e-09,
                                          -1.1462901882009371e-11,
                                          4.608976789597796e-14};
__device__ __constant__ double c_ps[9] = {1.0,
                                          -0.16666666666666666,
                                          0.008333333333333186,
                                          -0.00019841269841208676,
                                          2.7557319211229526e-06,
                                          -2.5052106890561953e-08,
                                          1.6058940878096299e-10,
                                          -7.643026547544204e-13,
                                          2.7215748296140716e-15};

// sin and cos of a first-quadrant-sized angle, no reduction, two independent Horner chains
__device__ __forceinline__ void fsincos_q1(double a, double* sp, double* cp) {
  const double z = a * a;
  double pc = c_pc[8], ps = c_ps[8];
#pragma unroll
  for (int k = 7; k >= 0; --k) {
    pc = fma(pc, z, c_pc[k]);
    ps = fma(ps, z, c_ps[k]);
  }
  *cp = pc;
  *sp = a * ps;
}

// sqrt(x), x positive and normal: third-order step on the 2^-20 seed (error ~3 e^3 ~ 2^-58 before
// the last rounding, i.e. <= ~0.51 ulp); no residual correction
__device__ __forceinline__ double fsqrt_fast(double x) {
  const double y = rsqrt_seed(x);
  const double g = x * y;
  const double h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  const double p = fma(r, 1.5, 1.0) * r;
  return fma(g, p, g);
}

// 1 / x to ~2^-60 (third-order step on the seed), no residual correction
__device__ __forceinline__ double frcp_fast(double x) {
  const double y = rcp_seed(x);
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  return fma(y, e, y);
}

// The quadrature-node versions of sqrt and 1 / x: z = sqrt(z^2) and 1 / (1 + z) inside the factor
// 1 + z^2 / (1 + z) of one node term of limb_dark.py:160-172.  With y the 2^-20 seed of 1 / sqrt(x), g = x y
// and the residual e = 1 - y g = -2 eps - eps^2:  sqrt(x) = g (1 + e / 2 + 3 e^2 / 8 + O(e^3)).
//   order 3: g + (g e) (1/2 + 3 e / 8)  -- 5 FP64 instructions, <= ~0.51 ulp like fsqrt_fast (which takes 6)
//   order 2: g + (g e) / 2              -- 4 instructions, relative error 1.5 eps^2 <= 1.4e-12
// and for the reciprocal, e = 1 - x y:  order 3: y + y (e + e^2) (3 instructions), order 2: y + y e (2; eps^2).
#ifndef JXP_NODE_SQRT_ORDER
#define JXP_NODE_SQRT_ORDER 3
#endif
#ifndef JXP_NODE_RCP_ORDER
#define JXP_NODE_RCP_ORDER 3
#endif
__device__ __forceinline__ double fsqrt_node(double x) {
  const double y = rsqrt_seed(x);
  const double g = x * y;
  const double e = fma(-y, g, 1.0);
  const double ge = g * e;
#if JXP_NODE_SQRT_ORDER == 3
  return fma(ge, fma(e, 0.375, 0.5), g);
#else
  return fma(ge, 0.5, g);
#endif
}
__device__ __forceinline__ double frcp_node(double x) {
  const double y = rcp_seed(x);
  double e = fma(-x, y, 1.0);
#if JXP_NODE_RCP_ORDER == 3
  e = fma(e, e, e);
#endif
  return fma(y, e, y);
}

// t = (1 - z^3) / (1 - z^2) = 1 + z^2 / (1 + z) from x = z^2 in (0, 1]: the node factor of limb_dark.py:160-172.
// Both hardware seeds are issued at the start -- the reciprocal's from 1 + g with g = x y the UNREFINED root
// (off by 2^-20, which the third-order step absorbs: (1.5e-6)^3 ~ 3e-18) -- so that neither MUFU latency sits
// behind the square root's refinement on the critical path.
#ifndef JXP_NODE_EARLY_SEED
#define JXP_NODE_EARLY_SEED 0
#endif
__device__ __forceinline__ double fnode_factor(double x) {
#if JXP_NODE_EARLY_SEED
  const double y = rsqrt_seed(x);
  const double g = x * y;
  const double v = rcp_seed(1.0 + g);
  const double e = fma(-y, g, 1.0);
  const double ge = g * e;
#if JXP_NODE_SQRT_ORDER == 3
  const double z = fma(ge, fma(e, 0.375, 0.5), g);
#else
  const double z = fma(ge, 0.5, g);
#endif
  double e2 = fma(-(z + 1.0), v, 1.0);
  e2 = fma(e2, e2, e2);
  const double w = fma(v, e2, v);
  return fma(x, w, 1.0);
#else
  const double z = fsqrt_node(x);
  return fma(x, frcp_node(z + 1.0), 1.0);
#endif
}

__device__ __constant__ double c_atan[11] = {
    -0.33333333333333120084,
    0.19999999999940620353,
    -0.14285714279223825339,
    0.11111110743605271685,
    -0.090908967710695170919,
    0.076920446486449617822,
    -0.066629439708755771522,
    0.058468205483397593716,
    -0.050348390947738729638,
    0.037958485922477833544,
    -0.01779789458017953007};

// atan2(y, x), any signs, |result error| ~ 1e-16.  One division: t = min/max of (|y|, |x|) is
// reduced to |t'| <= tan(pi/8) with t' = (min - max) / (min + max) and an offset of pi/4, then
// t' * P(t'^2) with a 12-term near-minimax P (error 1.3e-18, fitted with mpmath).
__device__ __forceinline__ double fatan2(double y, double x) {
  const double ay = fabs(y), ax = fabs(x);
  const bool swap = ay > ax;
  const double mn = swap ? ax : ay, mx = swap ? ay : ax;
  const bool big = mn > 0.41421356237309503 * mx;
  const double num = big ? mn - mx : mn;
  const double den = big ? mn + mx : mx;
  double t = fdiv(num, den);
  t = (den == 0.0) ? 0.0 : t;  // atan2(0, 0) = 0
  const double s = t * t;
  double p = c_atan[10];
#pragma unroll
  for (int k = 9; k >= 0; --k) p = fma(p, s, c_atan[k]);
  double a = fma(t * s, p, t);
  a = big ? a + 0.78539816339744830962 : a;
  a = swap ? 1.57079632679489661923 - a : a;
  a = (x < 0.0) ? 3.14159265358979323846 - a : a;
  return copysign(a, y);
}

}  // namespace jxp
