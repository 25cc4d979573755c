// K8: positions and velocities of Keplerian bodies over (parameter set x time sample).
//
// Reference path, one pass instead of a chain of array operations:
//   OrbitalBody._get_true_anomaly         src/jaxoplanet/orbits/keplerian.py:537-541
//   OrbitalBody._get_position_and_velocity                         keplerian.py:590-637
//   OrbitalBody._rotate_vector                                      keplerian.py:543-588
// behind position / central_position / relative_position / velocity / central_velocity /
// relative_velocity / radial_velocity (keplerian.py:366-532): those differ only in the scalar that
// multiplies the orbital radius (`r_scale`: -a, -a M*/M_tot, a m/M_tot, times parallax / au) and in
// the velocity amplitude (`k_amp`: k0 of keplerian.py:604-618), which the caller derives per set.
#include "../../include/jaxoplanet_b200.h"

#include <cuda_runtime.h>

#include "jxp_host.h"
#include "jxp_math.cuh"

using namespace jxp;
using namespace jxph;

namespace {

template <typename T>
struct StateArgs {
  const double* bodies;  // [n_sets][JXP_BODY_NFIELDS]
  int n_sets;
  const double* r_scale;  // [n_sets] or [1]
  int64_t r_stride;
  const double* k_amp;  // [n_sets] or [1]
  int64_t k_stride;
  const double* asc;  // [n_sets][2] (cos, sin of the ascending node) or nullptr
  const double* t;
  int64_t n_times;
  int flags;
  T* pos;  // [3][n_sets][n_times] or nullptr
  T* vel;  // [3][n_sets][n_times] or nullptr
};

template <typename T>
__global__ void __launch_bounds__(256) k_orbit_state(const __grid_constant__ StateArgs<T> a) {
  const int64_t total = (int64_t)a.n_sets * a.n_times;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const size_t plane = (size_t)total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int s = (int)(i / a.n_times);
    const double tm = a.t[i - (int64_t)s * a.n_times];
    const double* p = a.bodies + (size_t)s * JXP_BODY_NFIELDS;
    // mean anomaly, the reference's operation order (keplerian.py:538), always in double
    const double x = (tm - p[JXP_BODY_TIME_TRANSIT]) - p[JXP_BODY_TIME_REF];
    const double M = (Num<double>::two_pi * x) / p[JXP_BODY_PERIOD];
    const T Mr = T(mod_two_pi(M));
    const double ed = p[JXP_BODY_ECC];
    const bool circular = ed < 0.0;
    T sf, cf;
    T e = T(0);
    if (circular) {
      jsincos(Mr, &sf, &cf);  // keplerian.py:539-540
    } else {
      e = T(ed);
      const T ome = T(1.0 - ed), ope = T(1.0 + ed);
      T Edummy;
      kepler_reduced<T, false>(Mr, e, ome, ope, T(sqrt((1.0 - ed) * (1.0 + ed))), T(1.0 / (1.0 + ed)), sf, cf,
                               Edummy);
    }
    const T cw = circular ? T(1) : T(p[JXP_BODY_COS_OMEGA]), sw = circular ? T(0) : T(p[JXP_BODY_SIN_OMEGA]);
    const T ci = T(p[JXP_BODY_COS_INCL]), si = T(p[JXP_BODY_SIN_INCL]);
    T co = T(1), so = T(0);
    if (a.asc) {
      co = T(a.asc[2 * (size_t)s]);
      so = T(a.asc[2 * (size_t)s + 1]);
    }
    if (a.pos) {
      // keplerian.py:620-632: r0 (1 - e^2) / (1 + e cos f), then the rotation :568-588
      T r0 = T(a.r_scale[(size_t)s * a.r_stride]);
      if (!circular) r0 = r0 * jdiv(T(1) - e * e, jfma(e, cf, T(1)));
      const T x0 = r0 * cf, y0 = r0 * sf;
      const T x1 = cw * x0 - sw * y0, y1 = sw * x0 + cw * y0;
      const T x2 = x1, y2 = ci * y1, Z = -si * y1;
      a.pos[i] = a.asc ? co * x2 - so * y2 : x2;
      a.pos[plane + i] = a.asc ? so * x2 + co * y2 : y2;
      a.pos[2 * plane + i] = Z;
    }
    if (a.vel) {
      // keplerian.py:626-629, 634
      const T k0 = T(a.k_amp[(size_t)s * a.k_stride]);
      const T v1 = -k0 * sf, v2 = k0 * (cf + e);
      const T x1 = cw * v1 - sw * v2, y1 = sw * v1 + cw * v2;
      T x2 = x1, y2 = y1, Z = -y1;
      if (!(a.flags & JXP_STATE_NO_INCLINATION)) {
        y2 = ci * y1;
        Z = -si * y1;
      }
      a.vel[i] = a.asc ? co * x2 - so * y2 : x2;
      a.vel[plane + i] = a.asc ? so * x2 + co * y2 : y2;
      a.vel[2 * plane + i] = Z;
    }
  }
}

template <typename T>
int launch_state(const double* bodies, int32_t n_sets, const double* r_scale, int64_t r_stride,
                 const double* k_amp, int64_t k_stride, const double* asc_node, const double* t,
                 int64_t n_times, int32_t flags, T* pos, T* vel, void* stream) {
  if (n_sets < 0 || n_times < 0) return fail(JXP_ERR_BAD_SHAPE, "jxp_orbit_state: negative size");
  if (!stride_ok(r_stride) || !stride_ok(k_stride))
    return fail(JXP_ERR_BAD_SHAPE, "jxp_orbit_state: strides must be 0 or 1");
  if (n_sets == 0 || n_times == 0) return JXP_OK;
  if (!bodies || !t) return fail(JXP_ERR_NULL_POINTER, "jxp_orbit_state: null bodies / t");
  if (!pos && !vel) return fail(JXP_ERR_NULL_POINTER, "jxp_orbit_state: no output requested");
  if ((pos && !r_scale) || (vel && !k_amp))
    return fail(JXP_ERR_NULL_POINTER, "jxp_orbit_state: positions need r_scale, velocities k_amp");
  StateArgs<T> a;
  a.bodies = bodies;
  a.n_sets = n_sets;
  a.r_scale = r_scale;
  a.r_stride = r_stride;
  a.k_amp = k_amp;
  a.k_stride = k_stride;
  a.asc = asc_node;
  a.t = t;
  a.n_times = n_times;
  a.flags = flags;
  a.pos = pos;
  a.vel = vel;
  const int64_t total = (int64_t)n_sets * n_times;
  const int64_t want = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  k_orbit_state<T><<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("jxp_orbit_state");
}

}  // namespace

extern "C" int jxp_orbit_state_f64(const double* bodies, int32_t n_sets, const double* r_scale,
                                   int64_t r_stride, const double* k_amp, int64_t k_stride,
                                   const double* asc_node, const double* t, int64_t n_times, int32_t flags,
                                   double* pos, double* vel, void* stream) {
  return launch_state<double>(bodies, n_sets, r_scale, r_stride, k_amp, k_stride, asc_node, t, n_times, flags,
                              pos, vel, stream);
}
extern "C" int jxp_orbit_state_f32(const double* bodies, int32_t n_sets, const double* r_scale,
                                   int64_t r_stride, const double* k_amp, int64_t k_stride,
                                   const double* asc_node, const double* t, int64_t n_times, int32_t flags,
                                   float* pos, float* vel, void* stream) {
  return launch_state<float>(bodies, n_sets, r_scale, r_stride, k_amp, k_stride, asc_node, t, n_times, flags, pos,
                             vel, stream);
}
