/* C/OpenMP restatement of the jaxoplanet hot path.  TEST INFRASTRUCTURE ONLY: the CPU
 * baseline that bench.py times (cpu_baseline / --impl reference); never linked into the
 * product.  Same formulas, operation order and "compute everything, then select" structure
 * as the reference (which evaluates the full limb-darkening formula at every sample and
 * masks afterwards: light_curves/limb_dark.py:56-58, core/limb_dark.py:60-71).
 *
 *   kepler        src/jaxoplanet/core/kepler.py:28-56, 82-108
 *   limb dark     src/jaxoplanet/core/limb_dark.py:39-47, 50-84, 102-196
 *   orbit         src/jaxoplanet/orbits/keplerian.py:537-637
 *   light curve   src/jaxoplanet/light_curves/limb_dark.py:49-60
 *   integrate     src/jaxoplanet/light_curves/transforms.py:67-116
 * Checked against oracle/ref_numpy.py in tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846
#define EPS 2.220446049250313e-16

void refc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int refc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline double pymod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
  return r;
}

/* kepler.py:82-92 */
static inline double starter(double M, double ecc, double ome) {
  double M2 = M * M;
  double alpha = 3 * PI / (PI - 6 / PI);
  alpha += 1.6 / (PI - 6 / PI) * (PI - M) / (1 + ecc);
  double d = 3 * ome + alpha * ecc;
  double alphad = alpha * d;
  double r = (3 * alphad * (d - ome) + M2) * M;
  double q = 2 * alphad * ome - M2;
  double q2 = q * q;
  double w = cbrt(fabs(r) + sqrt(q2 * q + r * r));
  w = w * w;
  return (2 * r * w / (w * w + w * q + q2) + M) / d;
}

/* kepler.py:95-108 */
static inline double refine(double M, double ecc, double ome, double E) {
  double sE = E - sin(E);
  double cE = 1 - cos(E);
  double f_0 = ecc * sE + E * ome - M;
  double f_1 = ecc * cE + ome;
  double f_2 = ecc * (E - sE);
  double f_3 = 1 - f_1;
  double d_3 = -f_0 / (f_1 - 0.5 * f_0 * f_2 / f_1);
  double d_4 = -f_0 / (f_1 + 0.5 * d_3 * f_2 + (d_3 * d_3) * f_3 / 6);
  double d_42 = d_4 * d_4;
  double dE = -f_0 / (f_1 + 0.5 * d_4 * f_2 + d_4 * d_4 * f_3 / 6 - d_42 * d_4 * f_2 / 24);
  return E + dE;
}

/* kepler.py:28-56 */
static inline void kepler1(double M, double ecc, double* sinf_, double* cosf_) {
  M = pymod(M, 2 * PI);
  int high = M > PI;
  if (high) M = 2 * PI - M;
  double ome = 1 - ecc;
  double E = starter(M, ecc, ome);
  E = refine(M, ecc, ome, E);
  if (high) E = 2 * PI - E;
  double tan_half_f = sqrt((1 + ecc) / (1 - ecc)) * tan(0.5 * E);
  double tan2_half_f = tan_half_f * tan_half_f;
  double denom = 1 / (1 + tan2_half_f);
  *sinf_ = 2 * tan_half_f * denom;
  *cosf_ = (1 - tan2_half_f) * denom;
}

void refc_kepler(const double* M, const double* ecc, int64_t n, double* sinf_, double* cosf_) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) kepler1(M[i], ecc[i], &sinf_[i], &cosf_[i]);
}

static const double GL_X[10] = {-0.9739065285171717,  -0.8650633666889844, -0.6794095682990244,
                                -0.4333953941292472,  -0.14887433898163116, 0.14887433898163116,
                                0.4333953941292472,   0.6794095682990244,  0.8650633666889844,
                                0.9739065285171717};
static const double GL_W[10] = {0.06667134430868714, 0.14945134915058053, 0.21908636251598224,
                                0.26926671930999674, 0.2955242247147533,  0.2955242247147533,
                                0.26926671930999674, 0.21908636251598224, 0.14945134915058053,
                                0.06667134430868714};

/* limb_dark.py:187-196 */
static inline double kite_area(double a, double b, double c) {
  double t;
  if (a > b) { t = a; a = b; b = t; }
  if (b > c) { t = b; b = c; c = t; }
  if (a > b) { t = a; a = b; b = t; }
  double sq = (a + (b + c)) * (c - (a - b)) * (c + (a - b)) * (a + (b - c));
  return sqrt(fmax(0.0, sq));
}

/* limb_dark.py:50-84 for l_max = 2, then s . g - 1 (limb_dark.py:47) */
static inline double ld_quad1(double b, double r, double g0, double g1, double g2) {
  b = fabs(b);
  r = fabs(r);
  /* kappas :102-108 */
  double b2o = b * b;
  double factor = (r - 1) * (r + 1);
  int cond_k = (b > fabs(1 - r)) && (b < 1 + r);
  double area = cond_k ? kite_area(r, b, 1.0) : 0.0;
  double kappa0 = atan2(area, b2o + factor);
  double kappa1 = atan2(area, b2o - factor);

  int no_occ = b >= 1 + r;
  int full_occ = 1 + b <= r;
  int cond = no_occ || full_occ;
  double b_ = cond ? 1.0 : b;
  double b2 = b_ * b_;
  double r2 = r * r;

  /* s0s2 :111-137 (both branches evaluated, then selected) */
  double bpr = b_ + r;
  double onembpr2 = (1 + bpr) * (1 - bpr);
  double eta2 = 0.5 * r2 * (r2 + 2 * b2);
  double s0_lrg = PI * (1 - r2);
  double s2_lrg = 2 * s0_lrg + 4 * PI * (eta2 - 0.5);
  double Alens = kappa1 + r2 * kappa0 - area * 0.5;
  double s0_sml = PI - Alens;
  double s2_sml =
      2 * s0_sml + 2 * (-(PI - kappa1) + 2 * eta2 * kappa0 - 0.25 * area * (1 + 5 * r2 + b2));
  double delta = 4 * b_ * r;
  int lrg = (onembpr2 + delta) > delta;
  double s0 = lrg ? s0_lrg : s0_sml;
  double s2 = lrg ? s2_lrg : s2_sml;
  if (no_occ) s0 = PI;
  if (full_occ) s0 = 0.0;
  if (cond) s2 = 0.0;

  /* p_integral :140-173, 184 */
  double rng = 0.5 * kappa0;
  double acc = 0.0;
  for (int i = 0; i < 10; ++i) {
    double phi = rng * GL_X[i];
    double c = cos(phi + 0.5 * kappa0);
    double omz2 = fmax(0.0, r2 + b2 - 2 * b_ * r * c);
    double z2 = 1 - omz2;
    if (z2 < 10 * EPS) z2 = 1.0;
    double z3 = z2 * sqrt(z2);
    int cnd = omz2 < 10 * EPS;
    if (cnd) omz2 = 1.0;
    double result = 2 * r * (r - b_ * c) * (1 - z3) / (3 * omz2);
    if (cnd) result = 0.0;
    acc += result * GL_W[i];
  }
  double P0 = cond ? 0.0 : rng * acc;
  double s1 = -P0 - 2 * (kappa1 - PI) / 3;
  return (s0 * g0 + s1 * g1 + s2 * g2) - 1;
}

static inline void greens_quad(double u1, double u2, double* g0, double* g1, double* g2) {
  double a2 = -u2 / 4;
  double a1 = u1 + 2 * u2;
  double a0 = (1 - u1 - u2) + 2 * a2;
  double norm = PI * (a0 + a1 / 1.5);
  *g0 = a0 / norm;
  *g1 = a1 / norm;
  *g2 = a2 / norm;
}

void refc_ld_quad(double u1, double u2, const double* b, const double* r, int64_t n, double* out) {
  double g0, g1, g2;
  greens_quad(u1, u2, &g0, &g1, &g2);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) out[i] = ld_quad1(b[i], r[i], g0, g1, g2);
}

/* one body, one time: light_curves/limb_dark.py:49-60 over keplerian.py:537-637.
 * body record = the 12 fields of include/jaxoplanet_b200.h (JXP_BODY_*). */
static inline double body_flux(const double* p, double g0, double g1, double g2, double t) {
  double period = p[0], tt = p[1], tref = p[2], ecc = p[3];
  double cw = p[4], sw = p[5], ci = p[6], si = p[7], a = p[8], rstar = p[9], rad = p[10];
  double M = 2 * PI * ((t - tt) - tref) / period;
  double sf, cf, r0 = -a;
  if (ecc < 0) {
    sf = sin(M);
    cf = cos(M);
  } else {
    kepler1(M, ecc, &sf, &cf);
    r0 *= (1 - ecc * ecc) / (1 + ecc * cf);
  }
  double x = r0 * cf, y = r0 * sf, x1, y1;
  if (ecc < 0) {
    x1 = x;
    y1 = y;
  } else {
    x1 = cw * x - sw * y;
    y1 = sw * x + cw * y;
  }
  double Y = ci * y1, Z = -si * y1;
  double b = sqrt(x1 * x1 + Y * Y) / rstar;
  double r = rad / rstar;
  double lc = ld_quad1(b, r, g0, g1, g2); /* evaluated at every sample, like the reference */
  return Z > 0 ? lc : 0.0;
}

/* flux_sum[set][t] = sum over bodies of the exposure-integrated per-body flux - 1;
 * loglike[set] = sum_t Normal(1 + flux_sum, sigma).log_prob(y).  Either output may be NULL.
 * dt/w: exposure stencil (n_sub entries; n_sub = 1, dt = 0, w = 1 for none). */
void refc_system(const double* bodies, int32_t n_sets, int32_t n_bodies, const double* u,
                 int64_t u_stride, const double* t, int64_t n_times, const double* dt,
                 const double* w, int32_t n_sub, double* flux_sum, const double* y,
                 const double* sigma, int64_t sigma_stride, double* loglike) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int32_t s = 0; s < n_sets; ++s) {
    double g0, g1, g2;
    greens_quad(u[s * u_stride], u[s * u_stride + 1], &g0, &g1, &g2);
    double ll = 0.0;
    for (int64_t i = 0; i < n_times; ++i) {
      double total = 0.0;
      for (int32_t k = 0; k < n_bodies; ++k) {
        const double* p = bodies + ((size_t)s * n_bodies + k) * 12;
        double acc = 0.0;
        for (int32_t j = 0; j < n_sub; ++j) acc += w[j] * body_flux(p, g0, g1, g2, t[i] + dt[j]);
        total += acc;
      }
      if (flux_sum) flux_sum[(size_t)s * n_times + i] = total;
      if (loglike) {
        double sg = sigma[i * sigma_stride];
        double resid = (y[i] - (1 + total)) / sg;
        ll += -0.5 * resid * resid - log(sg) - 0.5 * log(2 * PI);
      }
    }
    if (loglike) loglike[s] = ll;
  }
}
