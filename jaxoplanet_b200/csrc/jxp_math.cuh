// Device math of the hot path: Kepler solve and polynomial(quadratic)-limb-darkened flux.
// Written from the published algorithms (Markley 1995 starter + one high-order
// correction; Agol, Luger & Foreman-Mackey 2020 solution vector with the reference's
// Gauss-Legendre line integral), re-arranged for the GPU FP64 pipe.  Where an expression
// is re-arranged relative to the reference the comment says why it stays inside the
// parity tolerance (DESIGN.md "numerics").
#pragma once
#include <type_traits>

#include "jxp_scalar.cuh"

namespace jxp {

// ------------------------------------------------------------------------------------
// M mod 2*pi into [0, 2pi), exact like the reference's `M % (2 * jnp.pi)`
// (src/jaxoplanet/core/kepler.py:31): k = floor(M / 2pi) from a multiply, then ONE fma
// M - k * fl(2pi).  The fma result is exact because the true remainder is a multiple of
// ulp(fl(2pi))/2 below 8 and therefore representable; a wrong k (M within an ulp of a
// multiple) is fixed by one exact +-2pi step.
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T mod_two_pi(T M) {
  T k = jfloor(M * Num<T>::inv_two_pi);
  T r = jfma(-k, Num<T>::two_pi, M);
  if (r < T(0)) r += Num<T>::two_pi;
  if (r >= Num<T>::two_pi) r -= Num<T>::two_pi;
  return r;
}

// Markley starter E0(M, e) on M in [0, pi], kepler.py:82-92, with the two divisions merged.
// The single correction that follows is 5th order, so the starter only has to be good to
// ~1e-4: the double-precision kernels evaluate it in FLOAT (FP32 pipe, otherwise idle) for
// 1 - e > 1e-5 -- measured final |E - E_true| identical to the all-double path down to
// 1 - e = 1e-6 (DESIGN.md "numerics") -- and in double beyond that.
template <typename T>
__device__ __forceinline__ T kepler_starter_t(T M, T e, T ome, T inv_ope) {
  const T pi = Num<T>::pi;
  const T k_a = T(3.0 * 3.141592653589793238462643383279502884 /
                  (3.141592653589793238462643383279502884 -
                   6.0 / 3.141592653589793238462643383279502884));
  const T k_b = T(1.6 / (3.141592653589793238462643383279502884 -
                         6.0 / 3.141592653589793238462643383279502884));
  const T M2 = M * M;
  const T alpha = jfma(k_b * (pi - M), inv_ope, k_a);
  const T d = jfma(alpha, e, T(3) * ome);
  const T alphad = alpha * d;
  const T r = jfma(T(3) * alphad, d - ome, M2) * M;
  const T q = jfma(T(2) * alphad, ome, -M2);
  const T q2 = q * q;
  T w = jcbrt_lo(jabs(r) + jsqrt(jfma(q2, q, r * r)));
  w = w * w;
  const T W = jfma(w, w + q, q2);
  return jdiv(jfma(T(2) * r, w, M * W), W * d);
}
__device__ __forceinline__ float kepler_starter(float M, float e, float ome, float inv_ope) {
  return kepler_starter_t<float>(M, e, ome, inv_ope);
}
__device__ __forceinline__ double kepler_starter(double M, double e, double ome, double inv_ope) {
  if (ome > 1e-5)
    return (double)kepler_starter_t<float>((float)M, (float)e, (float)ome, (float)inv_ope);
  return kepler_starter_t<double>(M, e, ome, inv_ope);
}

// ------------------------------------------------------------------------------------
// Kepler solve for one (M, e):  src/jaxoplanet/core/kepler.py:28-56, 82-108.
//   in : Mr in [0, 2pi) (already reduced), e, ome = 1 - e, ope = 1 + e,
//        s1me2 = sqrt((1 - e)(1 + e)), inv_ope = 1 / (1 + e)
//   out: sin f, cos f and (optionally) E in [0, 2pi)
// Same starter and the same single correction as the reference.  Re-arrangements:
//  * the starter's two divisions are merged into one (starter accuracy is irrelevant to
//    the result: the correction is 5th order, starter error <= 4.4e-4);
//  * sin E0, 1 - cos E0 come from sincos(E0/2): 2 s c and 2 s^2 (no cancellation);
//  * d3 uses one division: -f0 f1 / (f1^2 - f0 f2 / 2);
//  * instead of tan(E/2) sqrt((1+e)/(1-e)) -> (sin f, cos f) the half angle is advanced by
//    dE/2 with a 3-term Taylor rotation (|dE/2| <= 2.2e-4, truncation < 1e-24) and
//        sin f = 2 sqrt(1-e^2) s c / D,  cos f = ((1-e) c^2 - (1+e) s^2) / D,
//        D = (1-e) c^2 + (1+e) s^2        (s, c = sin, cos of E/2)
//    which is the same function (multiply tan-half-angle formulas through by (1-e) c^2),
//    has no cancellation in D and is finite at E = pi where tan(E/2) overflows.
// ------------------------------------------------------------------------------------
template <typename T, bool WANT_E>
__device__ __forceinline__ void kepler_reduced(T Mr, T e, T ome, T ope, T s1me2, T inv_ope,
                                               T& sinf, T& cosf, T& E) {
  const T pi = Num<T>::pi;
  const bool high = Mr > pi;
  const T M = high ? Num<T>::two_pi - Mr : Mr;

  // starter, kepler.py:82-92
  const T E0 = kepler_starter(M, e, ome, inv_ope);

  // refine, kepler.py:95-108
  T hs, hc;
  jsincos(T(0.5) * E0, &hs, &hc);
  const T sinE = T(2) * hs * hc;
  const T cE = T(2) * hs * hs;  // 1 - cos E0
  const T sE = E0 - sinE;
  const T f0 = jfma(e, sE, jfma(E0, ome, -M));
  const T f1 = jfma(e, cE, ome);
  const T f2 = e * sinE;
  const T f3 = T(1) - f1;
  const T d3 = jdiv(-f0 * f1, jfma(T(-0.5) * f0, f2, f1 * f1));
  const T d4 = jdiv(-f0, jfma(d3 * d3, f3 * T(1.0 / 6.0), jfma(T(0.5) * d3, f2, f1)));
  const T d42 = d4 * d4;
  const T dE = jdiv(-f0, jfma(-d42 * d4, f2 * T(1.0 / 24.0),
                               jfma(d42, f3 * T(1.0 / 6.0), jfma(T(0.5) * d4, f2, f1))));

  // advance the half angle by dE/2
  const T h = T(0.5) * dE;
  const T h2 = h * h;
  const T sh = jfma(-h * h2, T(1.0 / 6.0), h);
  const T ch = jfma(h2, jfma(h2, T(1.0 / 24.0), T(-0.5)), T(1));
  const T s = jfma(hs, ch, hc * sh);
  const T c = jfma(hc, ch, -hs * sh);

  const T ss = s * s, cc = c * c;
  const T a = ome * cc, b = ope * ss;
  const T invD = jrcp(a + b);
  const T sf = T(2) * s1me2 * s * c * invD;
  sinf = high ? -sf : sf;
  cosf = (a - b) * invD;
  if (WANT_E) {
    const T E1 = E0 + dE;
    E = high ? Num<T>::two_pi - E1 : E1;
  }
}

// N independent solves per thread, written statement by statement over the N items so that the
// instruction stream interleaves N dependency chains.  Why: the solve is one long chain (three
// dependent divisions) and a warp's DEPENDENT DFMAs issue 8 cycles apart; measured on B200
// (tools/ubench/dfma_lat.cu) a stream with one chain per warp saturates at 67 % of the FP64 pipe
// however many warps are resident, two chains reach 80 %, four 92 %.  Same arithmetic as
// kepler_reduced, operation for operation.
// XY: return (sin f, cos f) times D = (1-e) c^2 + (1+e) s^2 instead (no division).  The position
// needs exactly these: 1 + e cos f = (1 - e^2) (c^2 + s^2) / D, so r0 = -a (1 - e^2) / (1 + e cos f)
// = -a D and (r0 cos f, r0 sin f) = -a (cos f D, sin f D) -- keplerian.py:630-632 without its division
// and without the 1 / D of the half-angle formulas (c^2 + s^2 = 1 to rounding).
template <int N, bool XY = false, bool ONE_DIV = false>
__device__ __forceinline__ void kepler_reduced_n(const double (&Mr)[N], const double (&e)[N],
                                                 const double (&ome)[N], const double (&ope)[N],
                                                 const double (&s1me2)[N], const double (&inv_ope)[N],
                                                 double (&sinf)[N], double (&cosf)[N]) {
  const double pi = Num<double>::pi;
  bool high[N], slow = false;
  double M[N], E0[N], hs[N], hc[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    high[q] = Mr[q] > pi;
    M[q] = high[q] ? Num<double>::two_pi - Mr[q] : Mr[q];
    slow = slow || !(ome[q] > 1e-5);
  }
#pragma unroll
  for (int q = 0; q < N; ++q)
    E0[q] = (double)kepler_starter_t<float>((float)M[q], (float)e[q], (float)ome[q], (float)inv_ope[q]);
  if (slow) {
#pragma unroll
    for (int q = 0; q < N; ++q)
      if (!(ome[q] > 1e-5)) E0[q] = kepler_starter_t<double>(M[q], e[q], ome[q], inv_ope[q]);
  }
#pragma unroll
  for (int q = 0; q < N; ++q) jsincos(0.5 * E0[q], &hs[q], &hc[q]);
  double f0[N], f1[N], f2[N], f3[N], d3[N], d4[N], dE[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double sinE = 2.0 * hs[q] * hc[q];
    const double cE = 2.0 * hs[q] * hs[q];
    const double sE = E0[q] - sinE;
    f0[q] = fma(e[q], sE, fma(E0[q], ome[q], -M[q]));
    f1[q] = fma(e[q], cE, ome[q]);
    f2[q] = e[q] * sinE;
    f3[q] = 1.0 - f1[q];
  }
  if (ONE_DIV && !slow) {
    // the three nested corrections d3 -> d4 -> dE of kepler.py:101-107 carried as fractions, ONE
    // division instead of three dependent ones (each a MUFU seed + a Newton chain, ~60 cycles of
    // latency): d3 = N3 / D3, d4 = -f0 D3^2 / Q4 = N4 / D4, dE = -f0 D4^3 / Q.  Same function; with
    // 1 - e > 1e-5 (the float-starter range) f1 >= 1e-5 and D4^3 >= 1e-75: no underflow.
#pragma unroll
    for (int q = 0; q < N; ++q) {
      const double N3 = -f0[q] * f1[q];
      const double D3 = fma(-0.5 * f0[q], f2[q], f1[q] * f1[q]);
      const double D32 = D3 * D3;
      const double hf2 = 0.5 * f2[q], f36 = f3[q] * (1.0 / 6.0);
      const double D4 = fma(N3, fma(hf2, D3, N3 * f36), f1[q] * D32);
      const double N4 = -f0[q] * D32;
      const double D42 = D4 * D4;
      const double N42 = N4 * N4;
      // Q = f1 D4^3 + f2/2 N4 D4^2 + f3/6 N4^2 D4 - f2/24 N4^3
      const double Q = fma(D42, fma(f1[q], D4, hf2 * N4), N42 * fma(f36, D4, -(f2[q] * (1.0 / 24.0)) * N4));
      d3[q] = D42 * D4;  // numerator factor D4^3 (d3 itself is not needed any more)
      d4[q] = Q;
    }
#pragma unroll
    for (int q = 0; q < N; ++q) dE[q] = jdiv(-f0[q] * d3[q], d4[q]);
  } else {
#pragma unroll
  for (int q = 0; q < N; ++q) d3[q] = jdiv(-f0[q] * f1[q], fma(-0.5 * f0[q], f2[q], f1[q] * f1[q]));
#pragma unroll
  for (int q = 0; q < N; ++q)
    d4[q] = jdiv(-f0[q], fma(d3[q] * d3[q], f3[q] * (1.0 / 6.0), fma(0.5 * d3[q], f2[q], f1[q])));
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double d42 = d4[q] * d4[q];
    dE[q] = jdiv(-f0[q], fma(-d42 * d4[q], f2[q] * (1.0 / 24.0),
                              fma(d42, f3[q] * (1.0 / 6.0), fma(0.5 * d4[q], f2[q], f1[q]))));
  }
  }
  double s[N], c[N], a[N], b[N], invD[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double h = 0.5 * dE[q];
    const double h2 = h * h;
    const double sh = fma(-h * h2, 1.0 / 6.0, h);
    const double ch = fma(h2, fma(h2, 1.0 / 24.0, -0.5), 1.0);
    s[q] = fma(hs[q], ch, hc[q] * sh);
    c[q] = fma(hc[q], ch, -hs[q] * sh);
    a[q] = ome[q] * (c[q] * c[q]);
    b[q] = ope[q] * (s[q] * s[q]);
  }
  if (XY) {
#pragma unroll
    for (int q = 0; q < N; ++q) {
      const double sf = 2.0 * s1me2[q] * s[q] * c[q];
      sinf[q] = high[q] ? -sf : sf;
      cosf[q] = a[q] - b[q];
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < N; ++q) invD[q] = jrcp(a[q] + b[q]);
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double sf = 2.0 * s1me2[q] * s[q] * c[q] * invD[q];
    sinf[q] = high[q] ? -sf : sf;
    cosf[q] = (a[q] - b[q]) * invD[q];
  }
}

// ------------------------------------------------------------------------------------
// Limb darkening, src/jaxoplanet/core/limb_dark.py.  S is T or Dual<T, N>.
// ------------------------------------------------------------------------------------

// Gauss-Legendre nodes.  ORDER == 10 (the reference default, limb_dark.py:20) is compiled
// in: scipy.special.roots_legendre(10) (limb_dark.py:155) as exact float64 literals, plus
// cos(fl(fl(pi/2 * x_i) + pi/2)) for the kappa0 == pi case (planet inside the disk), where the
// node angles do not depend on (b, r) -- the values the reference's own argument rounding gives.
// ORDER == 0 selects a runtime table (any order up to 64) passed by the host.
struct GLTable {
  int order;
  double x[64];
  double w[64];
};

template <int ORDER>
struct GL;

// constant-bank copies for the rolled node loop of the fused kernels
__device__ __constant__ double c_gl10[3][10] = {
    {-0.9739065285171717, -0.8650633666889844, -0.6794095682990244, -0.4333953941292472,
     -0.14887433898163116, 0.14887433898163116, 0.4333953941292472, 0.6794095682990244,
     0.8650633666889844, 0.9739065285171717},
    {0.06667134430868714, 0.14945134915058053, 0.21908636251598224, 0.26926671930999674,
     0.2955242247147533, 0.2955242247147533, 0.26926671930999674, 0.21908636251598224,
     0.14945134915058053, 0.06667134430868714},
    {0.9991601288170098, 0.9776208824732407, 0.8758595017700398, 0.6293961480326326,
     0.23172567069845093, -0.23172567069845082, -0.6293961480326326, -0.8758595017700398,
     -0.9776208824732407, -0.9991601288170098}};

// w_i c_i for the constant node cosines of a planet fully inside the disk (c_gl10[1][i] * c_gl10[2][i])
__device__ __constant__ double c_gl10_wc[10] = {
    0.06667134430868714 * 0.9991601288170098,  0.14945134915058053 * 0.9776208824732407,
    0.21908636251598224 * 0.8758595017700398,  0.26926671930999674 * 0.6293961480326326,
    0.2955242247147533 * 0.23172567069845093,  0.2955242247147533 * -0.23172567069845082,
    0.26926671930999674 * -0.6293961480326326, 0.21908636251598224 * -0.8758595017700398,
    0.14945134915058053 * -0.9776208824732407, 0.06667134430868714 * -0.9991601288170098};

// ORDER == 11 is a tag: order 10 with a rolled (partially unrolled) node loop reading the tables
// from the constant bank -- small code for the fused kernels, where instruction fetch matters.
template <>
struct GL<11> {
  // no constant-node shortcut: in the fused kernels a warp mixes inside-the-disk and
  // ingress/egress samples, and a per-lane shortcut makes the warp execute BOTH loop versions
  // (profiles/r1_c3_v3_queue.md); one uniform loop with the 17-instruction node cosine is cheaper
  static constexpr bool has_cpi = false;
  __device__ __forceinline__ static int n(const GLTable*) { return 10; }
  __device__ __forceinline__ static double x(int i, const GLTable*) { return c_gl10[0][i]; }
  __device__ __forceinline__ static double w(int i, const GLTable*) { return c_gl10[1][i]; }
  __device__ __forceinline__ static double cpi(int i) { return c_gl10[2][i]; }
};

template <>
struct GL<10> {
  static constexpr bool has_cpi = true;
  __device__ __forceinline__ static int n(const GLTable*) { return 10; }
  __device__ __forceinline__ static double x(int i, const GLTable*) {
    constexpr double X[10] = {-0.9739065285171717,  -0.8650633666889844, -0.6794095682990244,
                              -0.4333953941292472,  -0.14887433898163116, 0.14887433898163116,
                              0.4333953941292472,   0.6794095682990244,  0.8650633666889844,
                              0.9739065285171717};
    return X[i];
  }
  __device__ __forceinline__ static double w(int i, const GLTable*) {
    constexpr double W[10] = {0.06667134430868714, 0.14945134915058053, 0.21908636251598224,
                              0.26926671930999674, 0.2955242247147533,  0.2955242247147533,
                              0.26926671930999674, 0.21908636251598224, 0.14945134915058053,
                              0.06667134430868714};
    return W[i];
  }
  __device__ __forceinline__ static double cpi(int i) {
    constexpr double C[10] = {0.9991601288170098,  0.9776208824732407,  0.8758595017700398,
                              0.6293961480326326,  0.23172567069845093, -0.23172567069845082,
                              -0.6293961480326326, -0.8758595017700398, -0.9776208824732407,
                              -0.9991601288170098};
    return C[i];
  }
};

template <>
struct GL<0> {
  static constexpr bool has_cpi = false;
  __device__ __forceinline__ static int n(const GLTable* t) { return t->order; }
  __device__ __forceinline__ static double x(int i, const GLTable* t) { return t->x[i]; }
  __device__ __forceinline__ static double w(int i, const GLTable* t) { return t->w[i]; }
  __device__ __forceinline__ static double cpi(int) { return 0.0; }
};

// kite_area, limb_dark.py:187-196
template <typename S>
__device__ __forceinline__ S kite_area(S a, S b, S c) {
  using B = typename BaseOf<S>::type;
  S t0 = jmin(a, b), t1 = jmax(a, b);
  a = t0;
  b = t1;
  t0 = jmin(b, c);
  t1 = jmax(b, c);
  b = t0;
  c = t1;
  t0 = jmin(a, b);
  t1 = jmax(a, b);
  a = t0;
  b = t1;
  S sq = (a + (b + c)) * (c - (a - b)) * (c + (a - b)) * (a + (b - c));
  return jzsqrt(jmax(S(B(0)), sq));
}

// One node of the l = 1 line integral (limb_dark.py:160-172): (1 - z^3) / (1 - z^2) given 1 - z^2, and
// whether the reference forces the node's term to 0.  (1 - z^3) / (1 - z^2) is evaluated as
// 1 + z^2 / (1 + z): the same function without the cancellation in 1 - z^3.
//   double: the reference's guards (z^2 < 10 eps or 1 - z^2 < 10 eps: term = 0).
//   fp32 MODE: the reference's guards are at 10 eps OF THE DTYPE (utils.get_dtype_eps); in float they zero
//   every node with z^2 < 1.2e-6, i.e. the whole integral for a planet within ~6e-7 of the outer contact --
//   1e-4 of flux (measured on BASELINE config 3: 118 ppm) where the double evaluation of the same formula
//   gives ~1e-11.  The fp32 mode answers to the DOUBLE reference (1 ppm), so it keeps the double guards'
//   meaning: they fire on a set of measure zero, and in between the term is finite and smooth.  Where float
//   cannot resolve z^2 (or 1 - z^2) the factor is taken at its limit (1 at the limb, 3 / 2 at the centre:
//   off by <= 1.2e-6 of a term that is itself <= 1e-3 there) with zero derivative; no term is zeroed.
template <typename S>
__device__ __forceinline__ S ld_node_factor(const S& omz2, bool& dead) {
  using B = typename BaseOf<S>::type;
  const B eps10 = B(10) * Num<B>::eps;
  if constexpr (std::is_same<B, float>::value) {
    const bool at_limb = !(val(omz2) <= B(1) - eps10);
    const bool at_centre = !(val(omz2) >= eps10);
    dead = false;
    const S z2s = jsel(at_limb || at_centre, S(B(1)), B(1) - omz2);
    const S z = jsqrt_pos(z2s);
    const S t = jdiv(z2s, z + B(1)) + B(1);
    return jsel(at_limb, S(B(1)), t);  // (at the centre z2s = 1 gives the limit 3 / 2)
  } else {
    dead = !((val(omz2) >= eps10) && (val(omz2) <= B(1) - eps10));
    const S z2s = jsel(dead, S(B(1)), B(1) - omz2);
    const S z = jsqrt_pos(z2s);
    return jdiv(z2s, z + B(1)) + B(1);
  }
}

// Solution vector (s0, s1, s2) for l_max <= 2, limb_dark.py:50-84 with kappas :102-108,
// s0s2 :111-137 and the l = 1 term of p_integral :140-173.
//   lmax 0: only s0 is meaningful; lmax 1: s0, s1; lmax 2: all.
//   HI: also the l = 3..lmax terms (limb_dark.py:149-153, 175-182) into shi[0..lmax-3].
//   PU: unroll factor of the mirror-pair node loop of the fused double kernels (1: smallest code,
//       matters when the hot code would otherwise outgrow the 32 KB instruction cache; 5: most ILP).
template <typename S, int ORDER, bool HI = false, int PU = 1>
__device__ __forceinline__ void ld_solution(S b, S r, int lmax, const GLTable* __restrict__ tab,
                                            S& s0, S& s1, S& s2, S* __restrict__ shi = nullptr) {
  using B = typename BaseOf<S>::type;
  const B pi = Num<B>::pi;
  const B eps10 = B(10) * Num<B>::eps;
  b = jabs(b);
  r = jabs(r);

  // kappas(b, r) is evaluated on the ORIGINAL b (limb_dark.py:58)
  const S b2o = b * b;
  const S factor = (r - B(1)) * (r + B(1));
  const bool cond_k = (val(b) > fabs(double(B(1) - val(r)))) && (val(b) < B(1) + val(r));
  S area = S(B(0));
  S kappa0, kappa1;
  if (cond_k) {
    area = kite_area(r, b, S(B(1)));
    kappa0 = jatan2(area, b2o + factor);
    kappa1 = jatan2(area, b2o - factor);
  } else {
    // area == 0 (constant): atan2(+0, x) is 0 for x >= +0 and pi for x < 0, derivative 0; a NaN b or
    // r has to stay a NaN as it does through the reference's atan2 (the comparisons above are all
    // false for it, and s0 of the uniform-disk term would otherwise come out finite)
    const B k0x = val(b2o) + val(factor), k1x = val(b2o) - val(factor);
    kappa0 = S((k0x != k0x) ? k0x : ((k0x < B(0)) ? pi : B(0)));
    kappa1 = S((k1x != k1x) ? k1x : ((k1x < B(0)) ? pi : B(0)));
  }

  const bool no_occ = val(b) >= B(1) + val(r);
  const bool full_occ = B(1) + val(b) <= val(r);
  const bool cond = no_occ || full_occ;
  if (cond) {
    // limb_dark.py:63-82 with b_ = 1: s0 = pi | 0, s2 = 0, P = 0
    s0 = S(no_occ ? pi : B(0));
    if (full_occ) s0 = S(B(0));
    s2 = S(B(0));
    s1 = -(kappa1 - pi) * B(2.0 / 3.0);
    if (HI) {
#pragma unroll
      for (int n = 0; n < 6; ++n) shi[n] = S(B(0));
    }
    return;
  }

  const S b2 = b2o;
  const S r2 = r * r;

  // s0s2, limb_dark.py:111-137
  {
    const S bpr = b + r;
    const S onembpr2 = (B(1) + bpr) * (B(1) - bpr);
    const S eta2 = r2 * (r2 + b2 * B(2)) * B(0.5);
    const S delta = b * r * B(4);
    const bool large = val(onembpr2) + val(delta) > val(delta);
    if (large) {
      s0 = (B(1) - r2) * pi;
      s2 = s0 * B(2) + (eta2 - B(0.5)) * (B(4) * pi);
    } else {
      const S Alens = kappa1 + r2 * kappa0 - area * B(0.5);
      s0 = pi - Alens;
      s2 = s0 * B(2) +
           (-(pi - kappa1) + eta2 * kappa0 * B(2) - area * (B(1) + r2 * B(5) + b2) * B(0.25)) *
               B(2);
    }
  }
  if (lmax < 1) {
    s1 = S(B(0));
    return;
  }

  // l = 1 term of p_integral, limb_dark.py:140-173, 184
  const S rng = kappa0 * B(0.5);
  const S two_br = b * r * B(2);
  const S r2pb2 = r2 + b2;
  const S two_r_3 = r * B(2.0 / 3.0);
  S acc = S(B(0));
  // l >= 3 (limb_dark.py:149-153): k^2 with the "factor" hack for b r -> 0
  S acch[6];
  S k2 = S(B(0)), factor4 = S(B(1));
  bool k2_cond = false;
  if (HI) {
#pragma unroll
    for (int n = 0; n < 6; ++n) acch[n] = S(B(0));
    factor4 = b * r * B(4);
    k2_cond = val(factor4) < eps10;
    factor4 = jsel(k2_cond, S(B(1)), factor4);
    k2 = jmax(S(B(0)), jdiv(B(1) - r2 - b2 + two_br, factor4));
  }
  if constexpr (ORDER == 11 && !HI && std::is_same<S, double>::value) {
    // fused kernels, double: the 10 nodes as 5 mirror pairs +-x_p.  With h = kappa0 / 2 the node
    // angles are h (1 +- x_p), so cos = cos h cos(h x_p) -+ sin h sin(h x_p): one short sin/cos
    // polynomial pair (|h x_p| <= 1.53, no reduction) serves two nodes.  sqrt and the division use
    // the uncorrected third-order steps (<= ~1 ulp) and 2 r / 3 is applied once after the loop.
    double sh, ch;
    fsincos_q1(rng, &sh, &ch);
    double accp = 0.0;
#pragma unroll PU
    for (int p = 0; p < 5; ++p) {
      const double xp = c_gl10[0][5 + p], wp = c_gl10[1][5 + p];
      double sa, ca;
      fsincos_q1(rng * xp, &sa, &ca);
      const double t1 = ch * ca;
      const double cm = fma(-sh, sa, t1);  // node +x_p: angle h + h x_p
      const double cq = fma(sh, sa, t1);   // node -x_p: angle h - h x_p
      double pair = 0.0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double c = q ? cq : cm;
        const double omz2 = fma(-two_br, c, r2pb2);
        const bool dead = !((omz2 >= eps10) && (omz2 <= 1.0 - eps10));
        const double z2s = dead ? 1.0 : 1.0 - omz2;
        const double z = fsqrt_node(z2s);
        const double t = fma(z2s, frcp_node(z + 1.0), 1.0);  // (1 - z^3) / (1 - z^2)
        const double res = fma(-b, c, r) * t;
        pair += dead ? 0.0 : res;
      }
      accp = fma(wp, pair, accp);
    }
    acc = accp * two_r_3;
  } else {
  // one node of limb_dark.py:160-172 (and :149-153, 175-182 when HI), given c = cos of its angle
  auto node_term = [&](const S& c, B wi) {
    // max(0, .) only matters for the guard below (a negative value is < 10 eps and the term is
    // forced to 0 either way), so the primal path skips it; the dual path keeps it for the
    // derivative.  z^2 < 10 eps  <=>  1 - z^2 > 1 - 10 eps exactly.
    const S omz2 = jclamp0_lazy(r2pb2 - two_br * c);
    // the term is 2 r (r - b c) (1 - z^3) / (3 (1 - z^2)), guards in ld_node_factor
    bool dead;
    const S tf = ld_node_factor(omz2, dead);
    S res = two_r_3 * (r - b * c) * tf;
    res = jsel(dead, S(B(0)), res);
    acc = acc + res * wi;
    if (HI) {
      // sin^2((phi + rng) / 2) = (1 - cos(phi + kappa0 / 2)) / 2   (limb_dark.py:161-162)
      const S sn2 = (B(1) - c) * B(0.5);
      const S f0 = jmax(S(B(0)), jsel(k2_cond, B(1) - r2, factor4 * (k2 - sn2)));
      const S common = two_r_3 * B(3) * (r - b + b * sn2 * B(2));  // 2 r (r - b + 2 b s^2)
#pragma unroll
      for (int n = 3; n <= 8; ++n)
        if (n <= lmax) acch[n - 3] = acch[n - 3] + jpow_half(f0, n) * common * wi;
    }
  };
  if constexpr (ORDER == 11 && !HI && !std::is_same<S, B>::value && std::is_same<B, double>::value) {
    // dual numbers, fused kernels (K5): the mirror-pair form gives sin AND cos of both node angles
    // from the same four products, which is exactly what d cos(angle) = -sin(angle) d angle needs;
    // d angle = (1 +- x_p) d(kappa0 / 2).  The rest of the node term stays on the generic dual rules.
    double sh, ch;
    fsincos_q1(val(rng), &sh, &ch);
#pragma unroll 1
    for (int p = 0; p < 5; ++p) {
      const double xp = c_gl10[0][5 + p], wp = c_gl10[1][5 + p];
      double sa, ca;
      fsincos_q1(val(rng) * xp, &sa, &ca);
      const double t1 = ch * ca, t2 = sh * sa, u1 = sh * ca, u2 = ch * sa;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double cosv = q ? t1 + t2 : t1 - t2;                 // angle h -+ ... : q = 0 is node +x_p
        const double msin = q ? -(u1 - u2) * (1.0 - xp) : -(u1 + u2) * (1.0 + xp);
        S c = rng * msin;  // derivative part: -sin(angle) (1 +- x_p) d rng
        c.v = cosv;
        node_term(c, wp);
      }
    }
  } else {
  const bool const_nodes = GL<ORDER>::has_cpi && !cond_k && (val(kappa0) != B(0));
  const int order = GL<ORDER>::n(tab);
#pragma unroll(ORDER == 10 ? 10 : 2)
  for (int i = 0; i < (ORDER == 10 ? 10 : order); ++i) {
    const B xi = B(GL<ORDER>::x(i, tab));
    const B wi = B(GL<ORDER>::w(i, tab));
    S c;
    if (const_nodes)
      c = S(B(GL<ORDER>::cpi(i)));
    else
      c = jcos_node(rng * xi + kappa0 * B(0.5));
    node_term(c, wi);
  }
  }
  }
  if (HI) {
#pragma unroll
    for (int n = 0; n < 6; ++n) shi[n] = -(rng * acch[n]);
  }
  const S P0 = rng * acc;
  s1 = -P0 - (kappa1 - pi) * B(2.0 / 3.0);
}

// Branch-free solution vector for N independent (b, r) points per thread, double, order 10,
// l_max <= 2: the same formulas as ld_solution with every `where` of limb_dark.py kept as a
// select (the reference evaluates both sides too), written statement by statement over the N
// items so that N dependency chains interleave (see kepler_reduced_n for why that matters).
// atan2(+0, x) is exactly 0 or pi, so the kappa selects of the scalar version are not needed.
template <int N, int PU>
__device__ __forceinline__ void ld_solution_n(const double (&b_in)[N], const double (&r_in)[N],
                                              int lmax, double (&s0)[N], double (&s1)[N],
                                              double (&s2)[N]) {
  const double pi = Num<double>::pi;
  const double eps10 = 10.0 * Num<double>::eps;
  double b[N], r[N], area[N], k0[N], k1[N], b2[N], r2[N];
  bool no_occ[N], full_occ[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    b[q] = fabs(b_in[q]);
    r[q] = fabs(r_in[q]);
  }
  // kite_area(r, b, 1) on the original b, limb_dark.py:187-196 (sorted Heron form)
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const bool cond_k = (b[q] > fabs(1.0 - r[q])) && (b[q] < 1.0 + r[q]);
    double x = r[q], y = b[q], z = 1.0, t;
    t = (x < y) ? x : y; y = (x < y) ? y : x; x = t;  // x <= y
    t = (y < z) ? y : z; z = (y < z) ? z : y; y = t;  // z = max
    t = (x < y) ? x : y; y = (x < y) ? y : x; x = t;  // x <= y <= z
    const double sq = (x + (y + z)) * (z - (x - y)) * (z + (x - y)) * (x + (y - z));
    area[q] = cond_k ? sq : 0.0;
  }
#pragma unroll
  for (int q = 0; q < N; ++q) area[q] = fsqrt(fmax(area[q], 0.0));
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double b2o = b[q] * b[q];
    const double factor = (r[q] - 1.0) * (r[q] + 1.0);
    k0[q] = b2o + factor;
    k1[q] = b2o - factor;
  }
#pragma unroll
  for (int q = 0; q < N; ++q) k0[q] = fatan2(area[q], k0[q]);
#pragma unroll
  for (int q = 0; q < N; ++q) k1[q] = fatan2(area[q], k1[q]);
#pragma unroll
  for (int q = 0; q < N; ++q) {
    no_occ[q] = b[q] >= 1.0 + r[q];
    full_occ[q] = 1.0 + b[q] <= r[q];
    b[q] = (no_occ[q] || full_occ[q]) ? 1.0 : b[q];  // b_ of limb_dark.py:63
    b2[q] = b[q] * b[q];
    r2[q] = r[q] * r[q];
  }
  // s0s2, limb_dark.py:111-137
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double bpr = b[q] + r[q];
    const double onembpr2 = (1.0 + bpr) * (1.0 - bpr);
    const double eta2 = r2[q] * (r2[q] + b2[q] * 2.0) * 0.5;
    const double delta = b[q] * r[q] * 4.0;
    const bool large = onembpr2 + delta > delta;
    const double s0l = (1.0 - r2[q]) * pi;
    const double s2l = s0l * 2.0 + (eta2 - 0.5) * (4.0 * pi);
    const double Alens = k1[q] + r2[q] * k0[q] - area[q] * 0.5;
    const double s0s = pi - Alens;
    const double s2s = s0s * 2.0 + (-(pi - k1[q]) + eta2 * k0[q] * 2.0 -
                                    area[q] * (1.0 + r2[q] * 5.0 + b2[q]) * 0.25) * 2.0;
    const double s0v = large ? s0l : s0s;
    const double s2v = large ? s2l : s2s;
    s0[q] = no_occ[q] ? pi : (full_occ[q] ? 0.0 : s0v);
    s2[q] = (no_occ[q] || full_occ[q]) ? 0.0 : s2v;
  }
  if (lmax < 1) {
#pragma unroll
    for (int q = 0; q < N; ++q) s1[q] = 0.0;
    return;
  }
  // l = 1 term of p_integral (limb_dark.py:140-173, 184) over the 5 mirror pairs of nodes
  double sh[N], ch[N], rng[N], two_br[N], omr2b2[N], acc[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    rng[q] = k0[q] * 0.5;
    two_br[q] = b[q] * r[q] * 2.0;
    omr2b2[q] = 1.0 - (r2[q] + b2[q]);
    acc[q] = 0.0;
  }
#pragma unroll
  for (int q = 0; q < N; ++q) fsincos_q1(rng[q], &sh[q], &ch[q]);
#pragma unroll PU
  for (int p = 0; p < 5; ++p) {
    const double xp = c_gl10[0][5 + p], wp = c_gl10[1][5 + p];
    double c[2 * N], z2s[2 * N], t[2 * N];
    bool dead[2 * N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
      double sa, ca;
      fsincos_q1(rng[q] * xp, &sa, &ca);
      const double t1 = ch[q] * ca;
      c[2 * q] = fma(-sh[q], sa, t1);
      c[2 * q + 1] = fma(sh[q], sa, t1);
    }
    // z^2 = 1 - (r^2 + b^2 - 2 b r c) as ONE fma, (1 - r^2 - b^2) + (2 b r) c -- the same quantity to an ulp
    // of 1, which is the precision the reference's own 1 - omz2 has; the guards 1 - z^2 < 10 eps and
    // z^2 < 10 eps (limb_dark.py:163-170) are read off z^2
#pragma unroll
    for (int m = 0; m < 2 * N; ++m) {
      const double z2 = fma(two_br[m >> 1], c[m], omr2b2[m >> 1]);
      dead[m] = !((z2 <= 1.0 - eps10) && (z2 >= eps10));
      z2s[m] = dead[m] ? 1.0 : z2;
    }
#pragma unroll
    for (int m = 0; m < 2 * N; ++m) t[m] = fnode_factor(z2s[m]);
#pragma unroll
    for (int q = 0; q < N; ++q) {
      const double ra = fma(-b[q], c[2 * q], r[q]) * t[2 * q];
      const double rb = fma(-b[q], c[2 * q + 1], r[q]) * t[2 * q + 1];
      acc[q] = fma(wp, (dead[2 * q] ? 0.0 : ra) + (dead[2 * q + 1] ? 0.0 : rb), acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double P0 = (no_occ[q] || full_occ[q]) ? 0.0 : rng[q] * (acc[q] * (r[q] * (2.0 / 3.0)));
    s1[q] = -P0 - (k1[q] - pi) * (2.0 / 3.0);
  }
}

// ld_solution (l_max = 2, order 10) for a point with the planet fully inside the disk, b < 1 - r and
// r < 1 (the caller guarantees it), S = T or Dual<T, N>: kite_area = 0, kappa0 = pi and kappa1 = 0 are
// CONSTANTS there (the general code takes the same branch: `cond_k` false, derivative 0), so the
// node angles pi (1 +- x_p) / 2 and their cosines are constants too (GL<10>::cpi) and carry no
// derivative.  Node order of the mirror-pair loops (+x_p, -x_p for p = 0..4).
template <typename S>
__device__ __forceinline__ void ld_solution_inside(S b, S r, S& s0, S& s1, S& s2) {
  using B = typename BaseOf<S>::type;
  const B pi = Num<B>::pi;
  const B eps10 = B(10) * Num<B>::eps;
  b = jabs(b);
  r = jabs(r);
  const S b2 = b * b;
  const S r2 = r * r;
  {
    const S bpr = b + r;
    const S onembpr2 = (B(1) + bpr) * (B(1) - bpr);
    const S eta2 = r2 * (r2 + b2 * B(2)) * B(0.5);
    const S delta = b * r * B(4);
    const bool large = val(onembpr2) + val(delta) > val(delta);
    if (large) {
      s0 = (B(1) - r2) * pi;
      s2 = s0 * B(2) + (eta2 - B(0.5)) * (B(4) * pi);
    } else {
      const S Alens = r2 * pi;  // kappa1 + r^2 kappa0 - area / 2
      s0 = pi - Alens;
      s2 = s0 * B(2) + (eta2 * (pi * B(2)) - pi) * B(2);
    }
  }
  const S two_br = b * r * B(2);
  const S r2pb2 = r2 + b2;
  const S two_r_3 = r * B(2.0 / 3.0);
  S acc = S(B(0));
#pragma unroll 1
  for (int p = 0; p < 5; ++p) {
    const B wi = B(c_gl10[1][5 + p]);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const B c = B(q ? c_gl10[2][4 - p] : c_gl10[2][5 + p]);
      const S omz2 = jclamp0_lazy(r2pb2 - two_br * c);
      bool dead;
      const S tf = ld_node_factor(omz2, dead);
      S res = two_r_3 * (r - b * c) * tf;
      res = jsel(dead, S(B(0)), res);
      acc = acc + res * wi;
    }
  }
  const S P0 = acc * (pi * B(0.5));
  s1 = -P0 + pi * B(2.0 / 3.0);
}

// ld_solution_n for points with the planet fully inside the disk, b < 1 - r and r < 1 (the caller
// guarantees it): there kite_area = 0, kappa0 = atan2(+0, b^2 + r^2 - 1) = pi and kappa1 =
// atan2(+0, b^2 - r^2 + 1) = 0 exactly, so the two atan2, the kite area and the ten node cosines
// drop out: the node angles are pi (1 +- x_p) / 2, constants (c_gl10[2], the values the reference's
// own argument rounding produces, as in K2's GL<10>::cpi).  Same pairing and summation order as the
// general version; the `large` select of s0s2 (limb_dark.py:118) is kept.
template <int N>
__device__ __forceinline__ void ld_solution_inside_n(const double (&b_in)[N], const double (&r_in)[N],
                                                     int lmax, double (&s0)[N], double (&s1)[N],
                                                     double (&s2)[N]) {
  const double pi = Num<double>::pi;
  const double eps10 = 10.0 * Num<double>::eps;
  double b[N], r[N], b2[N], r2[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    b[q] = fabs(b_in[q]);
    r[q] = fabs(r_in[q]);
    b2[q] = b[q] * b[q];
    r2[q] = r[q] * r[q];
  }
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double bpr = b[q] + r[q];
    const double onembpr2 = (1.0 + bpr) * (1.0 - bpr);
    const double eta2 = r2[q] * (r2[q] + b2[q] * 2.0) * 0.5;
    const double delta = b[q] * r[q] * 4.0;
    const bool large = onembpr2 + delta > delta;
    const double s0l = (1.0 - r2[q]) * pi;
    const double s2l = s0l * 2.0 + (eta2 - 0.5) * (4.0 * pi);
    const double Alens = 0.0 + r2[q] * pi - 0.0 * 0.5;  // kappa1 + r^2 kappa0 - area / 2
    const double s0s = pi - Alens;
    const double s2s = s0s * 2.0 + (-(pi - 0.0) + eta2 * pi * 2.0 - 0.0 * (1.0 + r2[q] * 5.0 + b2[q]) * 0.25) * 2.0;
    s0[q] = large ? s0l : s0s;
    s2[q] = large ? s2l : s2s;
  }
  if (lmax < 1) {
#pragma unroll
    for (int q = 0; q < N; ++q) s1[q] = 0.0;
    return;
  }
  double two_br[N], r2pb2[N], acc[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    two_br[q] = b[q] * r[q] * 2.0;
    r2pb2[q] = r2[q] + b2[q];
    acc[q] = 0.0;
  }
#pragma unroll
  for (int p = 0; p < 5; ++p) {
    const double wp = c_gl10[1][5 + p];
    const double cm = c_gl10[2][5 + p], cq = c_gl10[2][4 - p];  // nodes +x_p, -x_p
    double z2s[2 * N], z[2 * N], t[2 * N];
    bool dead[2 * N];
#pragma unroll
    for (int m = 0; m < 2 * N; ++m) {
      const double omz2 = fma(-two_br[m >> 1], (m & 1) ? cq : cm, r2pb2[m >> 1]);
      dead[m] = !((omz2 >= eps10) && (omz2 <= 1.0 - eps10));
      z2s[m] = dead[m] ? 1.0 : 1.0 - omz2;
    }
#pragma unroll
    for (int m = 0; m < 2 * N; ++m) t[m] = fnode_factor(z2s[m]);
#pragma unroll
    for (int q = 0; q < N; ++q) {
      const double ra = fma(-b[q], cm, r[q]) * t[2 * q];
      const double rb = fma(-b[q], cq, r[q]) * t[2 * q + 1];
      acc[q] = fma(wp, (dead[2 * q] ? 0.0 : ra) + (dead[2 * q + 1] ? 0.0 : rb), acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const double P0 = (pi * 0.5) * (acc[q] * (r[q] * (2.0 / 3.0)));
    s1[q] = -P0 - (0.0 - pi) * (2.0 / 3.0);
  }
}

// Normalised Green's-basis coefficients for n_u <= 2, limb_dark.py:42-45, 87-99.
template <typename S>
__device__ __forceinline__ void ld_greens(int n_u, S u1, S u2, S& g0, S& g1, S& g2) {
  using B = typename BaseOf<S>::type;
  const B pi = Num<B>::pi;
  if (n_u == 0) {
    g0 = S(B(1) / pi);
    g1 = S(B(0));
    g2 = S(B(0));
    return;
  }
  if (n_u == 1) {
    g0 = B(1) - u1;
    g1 = u1;
    g2 = S(B(0));
  } else {
    g2 = u2 * B(-0.25);
    g1 = u1 + u2 * B(2);
    g0 = (B(1) - u1 - u2) + g2 * B(2);
  }
  const S norm = (g0 + g1 / B(1.5)) * pi;
  g0 = g0 / norm;
  g1 = g1 / norm;
  g2 = g2 / norm;
}

// Normalised Green's-basis coefficients for any n_u <= 8 (g[0..n_u]), limb_dark.py:42-45, 87-99.
template <typename S>
__device__ inline void ld_greens_n(int n_u, const S* __restrict__ u, S* __restrict__ g) {
  using B = typename BaseOf<S>::type;
  const B pi = Num<B>::pi;
  if (n_u == 0) {
    g[0] = S(B(1) / pi);
    return;
  }
  const int size = n_u + 1;  // u_ext = (-1, u_0, ..)
  S p[11], gg[13];
  for (int n = 0; n < 13; ++n) gg[n] = S(B(0));
  for (int n = 0; n < size; ++n) {
    // arg_n = sum_i binom(i, n) u_ext[i]
    S arg = S(B(0));
    B binom = B(1);  // C(n, n)
    for (int i = n; i < size; ++i) {
      const S ui = (i == 0) ? S(B(-1)) : u[i - 1];
      arg = arg + ui * binom;
      binom = binom * B(i + 1) / B(i + 1 - n);  // C(i + 1, n)
    }
    p[n] = (n % 2 == 0) ? -arg : arg;  // (-1)^(n + 1)
  }
  for (int n = size - 1; n > 1; --n) gg[n] = p[n] / B(n + 2) + gg[n + 2];
  gg[1] = p[1] + gg[3] * B(3);
  gg[0] = p[0] + gg[2] * B(2);
  const S norm = (gg[0] + gg[1] / B(1.5)) * pi;
  for (int n = 0; n < size; ++n) g[n] = gg[n] / norm;
}

template <typename S>
__device__ __forceinline__ S ld_combine(int lmax, const S& s0, const S& s1, const S& s2,
                                        const S& g0, const S& g1, const S& g2) {
  using B = typename BaseOf<S>::type;
  S f = s0 * g0;
  if (lmax >= 1) f = f + s1 * g1;
  if (lmax >= 2) f = f + s2 * g2;
  return f - B(1);
}

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

}  // namespace jxp
