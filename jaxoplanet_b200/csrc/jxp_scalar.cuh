// Scalar abstraction for the device math: the same templated formulas are instantiated
// with a plain floating type (primal kernels) and with Dual<T, N> (forward-mode partials).
// The Dual rules are the JAX differentiation rules of the primitives the reference uses
// (abs: sign(0) = 0, maximum/minimum: ties split 1/2, where: taken branch,
// zero_safe_sqrt: src/jaxoplanet/utils.py:16-24), so that the partials equal what
// jax.jvp of the reference produces for the discretised formulas.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "jxp_fastmath.cuh"

namespace jxp {

template <typename T>
struct Num;

template <>
struct Num<double> {
  static constexpr double pi = 3.141592653589793238462643383279502884;
  static constexpr double two_pi = 6.283185307179586476925286766559005768;
  static constexpr double half_pi = 1.570796326794896619231321691639751442;
  static constexpr double inv_two_pi = 0.159154943091895335768883763372514362;
  static constexpr double eps = 2.220446049250313e-16;
  static constexpr double third = 0.333333333333333333333333333333333333;
};
template <>
struct Num<float> {
  static constexpr float pi = 3.14159265358979323846f;
  static constexpr float two_pi = 6.28318530717958647692f;
  static constexpr float half_pi = 1.57079632679489661923f;
  static constexpr float inv_two_pi = 0.15915494309189533577f;
  static constexpr float eps = 1.1920928955078125e-07f;
  static constexpr float third = 0.33333333333333333333f;
};

// ---- primitive wrappers on plain types -------------------------------------------
__device__ __forceinline__ double jfma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float jfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double jsqrt(double x) { return fsqrt(x); }
__device__ __forceinline__ float jsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double jzsqrt(double x) { return fsqrt(x); }
__device__ __forceinline__ float jzsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double jcbrt(double x) { return cbrt(x); }
// cube root to float accuracy (the Kepler starter only needs ~1e-6)
__device__ __forceinline__ double jcbrt_lo(double x) { return (double)cbrtf((float)x); }
__device__ __forceinline__ float jcbrt_lo(float x) { return cbrtf(x); }
// sqrt of a value known to be positive and normal; clamp-at-zero that the primal path may skip
__device__ __forceinline__ double jsqrt_pos(double x) { return fsqrt_pos(x); }
__device__ __forceinline__ float jsqrt_pos(float x) { return sqrtf(x); }
__device__ __forceinline__ double jclamp0_lazy(double x) { return x; }
__device__ __forceinline__ float jclamp0_lazy(float x) { return x; }
// x^(n/2) for x >= 0, 3 <= n <= 8 (general polynomial limb darkening)
template <typename T>
__device__ __forceinline__ T jpow_half_m1(T x, int n) {  // x^(n/2 - 1)
  const T sq = jsqrt(x);
  T p = (n & 1) ? sq : x;       // n = 3: x^0.5, n = 4: x
  for (int k = (n & 1) ? 3 : 4; k < n; k += 2) p = p * x;
  return p;
}
__device__ __forceinline__ double jpow_half(double x, int n) { return jpow_half_m1(x, n) * x; }
__device__ __forceinline__ float jpow_half(float x, int n) { return jpow_half_m1(x, n) * x; }
// cos of a quadrature node angle in [0, pi]
__device__ __forceinline__ double jcos_node(double x) { return fcos_0pi(x); }
__device__ __forceinline__ float jcos_node(float x) { return cosf(x); }
__device__ __forceinline__ float jcbrt(float x) { return cbrtf(x); }
__device__ __forceinline__ double jabs(double x) { return fabs(x); }
__device__ __forceinline__ float jabs(float x) { return fabsf(x); }
__device__ __forceinline__ double jmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float jmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double jmin(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float jmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double jatan2(double y, double x) { return fatan2(y, x); }
__device__ __forceinline__ float jatan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double jcos(double x) { return fcos(x); }
__device__ __forceinline__ float jcos(float x) { return cosf(x); }
__device__ __forceinline__ double jsin(double x) {
  double s_, c_;
  fsincos(x, &s_, &c_);
  return s_;
}
__device__ __forceinline__ float jsin(float x) { return sinf(x); }
__device__ __forceinline__ void jsincos(double x, double* s, double* c) { fsincos(x, s, c); }
__device__ __forceinline__ void jsincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ double jdiv(double a, double b) { return fdiv(a, b); }
__device__ __forceinline__ float jdiv(float a, float b) { return a / b; }
__device__ __forceinline__ double jrcp(double x) { return frcp(x); }
__device__ __forceinline__ float jrcp(float x) { return 1.0f / x; }
__device__ __forceinline__ double jfloor(double x) { return floor(x); }
__device__ __forceinline__ float jfloor(float x) { return floorf(x); }
__device__ __forceinline__ double jsel(bool c, double a, double b) { return c ? a : b; }
__device__ __forceinline__ float jsel(bool c, float a, float b) { return c ? a : b; }
__device__ __forceinline__ double val(double x) { return x; }
__device__ __forceinline__ float val(float x) { return x; }

template <typename S>
struct BaseOf {
  using type = S;
};

// ---- forward-mode dual numbers ----------------------------------------------------
template <typename T, int N>
struct Dual {
  T v;
  T d[N];
  __device__ __forceinline__ Dual() {}
  __device__ __forceinline__ Dual(T x) : v(x) {
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = T(0);
  }
  __device__ __forceinline__ static Dual seed(T x, int k) {
    Dual r(x);
    r.d[k] = T(1);
    return r;
  }
};
template <typename T, int N>
struct BaseOf<Dual<T, N>> {
  using type = T;
};

#define JXP_DUAL_TPL template <typename T, int N>
#define JXP_D Dual<T, N>
#define JXP_FOR _Pragma("unroll") for (int i = 0; i < N; ++i)

JXP_DUAL_TPL __device__ __forceinline__ T val(const JXP_D& x) { return x.v; }

JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator-(const JXP_D& a) {
  JXP_D r;
  r.v = -a.v;
  JXP_FOR r.d[i] = -a.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator+(const JXP_D& a, const JXP_D& b) {
  JXP_D r;
  r.v = a.v + b.v;
  JXP_FOR r.d[i] = a.d[i] + b.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator-(const JXP_D& a, const JXP_D& b) {
  JXP_D r;
  r.v = a.v - b.v;
  JXP_FOR r.d[i] = a.d[i] - b.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator*(const JXP_D& a, const JXP_D& b) {
  JXP_D r;
  r.v = a.v * b.v;
  JXP_FOR r.d[i] = jfma(a.d[i], b.v, a.v * b.d[i]);
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator/(const JXP_D& a, const JXP_D& b) {
  JXP_D r;
  T inv = jrcp(b.v);
  r.v = jdiv(a.v, b.v);
  JXP_FOR r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
// mixed with plain scalars
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator+(const JXP_D& a, T b) {
  JXP_D r = a;
  r.v = a.v + b;
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator+(T a, const JXP_D& b) { return b + a; }
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator-(const JXP_D& a, T b) {
  JXP_D r = a;
  r.v = a.v - b;
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator-(T a, const JXP_D& b) {
  JXP_D r;
  r.v = a - b.v;
  JXP_FOR r.d[i] = -b.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator*(const JXP_D& a, T b) {
  JXP_D r;
  r.v = a.v * b;
  JXP_FOR r.d[i] = a.d[i] * b;
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator*(T a, const JXP_D& b) { return b * a; }
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator/(const JXP_D& a, T b) {
  T inv = jrcp(b);
  return a * inv;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jdiv(const JXP_D& a, const JXP_D& b) { return a / b; }
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jdiv(const JXP_D& a, T b) { return a / b; }
JXP_DUAL_TPL __device__ __forceinline__ JXP_D operator/(T a, const JXP_D& b) {
  JXP_D r;
  T inv = jrcp(b.v);
  r.v = a * inv;
  T f = -r.v * inv;
  JXP_FOR r.d[i] = f * b.d[i];
  return r;
}

JXP_DUAL_TPL __device__ __forceinline__ JXP_D jfma(const JXP_D& a, const JXP_D& b,
                                                   const JXP_D& c) {
  return a * b + c;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jsqrt(const JXP_D& x) {
  JXP_D r;
  r.v = jsqrt(x.v);
  T f = T(0.5) * jrcp(r.v);
  JXP_FOR r.d[i] = x.d[i] * f;
  return r;
}
// zero_safe_sqrt, src/jaxoplanet/utils.py:16-24
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jzsqrt(const JXP_D& x) {
  JXP_D r;
  r.v = jsqrt(x.v);
  T denom = (x.v < T(10) * Num<T>::eps) ? T(1) : x.v;
  T f = T(0.5) * r.v * jrcp(denom);
  JXP_FOR r.d[i] = x.d[i] * f;
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jabs(const JXP_D& x) {
  JXP_D r;
  r.v = jabs(x.v);
  T s = (x.v > T(0)) ? T(1) : ((x.v < T(0)) ? T(-1) : T(0));
  JXP_FOR r.d[i] = s * x.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jmax(const JXP_D& a, const JXP_D& b) {
  JXP_D r;
  r.v = jmax(a.v, b.v);
  T wa = (a.v > b.v) ? T(1) : ((a.v == b.v) ? T(0.5) : T(0));
  T wb = T(1) - wa;
  JXP_FOR r.d[i] = wa * a.d[i] + wb * b.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jmin(const JXP_D& a, const JXP_D& b) {
  JXP_D r;
  r.v = jmin(a.v, b.v);
  T wa = (a.v < b.v) ? T(1) : ((a.v == b.v) ? T(0.5) : T(0));
  T wb = T(1) - wa;
  JXP_FOR r.d[i] = wa * a.d[i] + wb * b.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jatan2(const JXP_D& y, const JXP_D& x) {
  JXP_D r;
  r.v = jatan2(y.v, x.v);
  T inv = jrcp(x.v * x.v + y.v * y.v);
  T fy = x.v * inv, fx = -y.v * inv;
  JXP_FOR r.d[i] = fy * y.d[i] + fx * x.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jcos(const JXP_D& x) {
  JXP_D r;
  T s, c;
  jsincos(x.v, &s, &c);
  r.v = c;
  JXP_FOR r.d[i] = -s * x.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jcos_node(const JXP_D& x) { return jcos(x); }
// d x^(n/2) = (n/2) x^(n/2 - 1) dx   (jnp.power rule; finite at x = 0 for n >= 3)
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jpow_half(const JXP_D& x, int n) {
  JXP_D r;
  const T m1 = jpow_half_m1(x.v, n);
  r.v = m1 * x.v;
  const T f = T(0.5) * T(n) * m1;
  JXP_FOR r.d[i] = f * x.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jsqrt_pos(const JXP_D& x) { return jzsqrt(x); }
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jclamp0_lazy(const JXP_D& x) {
  return jmax(JXP_D(T(0)), x);
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jsin(const JXP_D& x) {
  JXP_D r;
  T s, c;
  jsincos(x.v, &s, &c);
  r.v = s;
  JXP_FOR r.d[i] = c * x.d[i];
  return r;
}
JXP_DUAL_TPL __device__ __forceinline__ void jsincos(const JXP_D& x, JXP_D* so, JXP_D* co) {
  T s, c;
  jsincos(x.v, &s, &c);
  so->v = s;
  co->v = c;
  JXP_FOR {
    so->d[i] = c * x.d[i];
    co->d[i] = -s * x.d[i];
  }
}
JXP_DUAL_TPL __device__ __forceinline__ JXP_D jsel(bool c, const JXP_D& a, const JXP_D& b) {
  return c ? a : b;
}

#undef JXP_DUAL_TPL
#undef JXP_D
#undef JXP_FOR

}  // namespace jxp
