// XLA FFI custom-call handlers over the C ABI of include/jaxoplanet_b200.h: what lets
// jax.jit / jax.vmap / jax.custom_jvp (and so NumPyro) compose with the sm_100a kernels.
//
// Each handler is argument marshalling only: shapes from the XLA buffers -> the sizes of the C
// entry point, the stream XLA hands over (ffi::PlatformStream<cudaStream_t>), the workspace from
// ffi::ScratchAllocator, static options as attributes.  Nothing here allocates, synchronises or
// keeps state, so the handlers are safe under XLA's concurrent executor threads and inside CUDA
// graphs (command buffers).
//
// Reference callables behind the handlers (file:line in exoplanet-dev/jaxoplanet):
//   JxpKepler*            core.kepler.kepler + its defjvp            src/jaxoplanet/core/kepler.py:14-79
//   JxpLimbDark*          core.limb_dark.light_curve                 src/jaxoplanet/core/limb_dark.py:19-47
//   JxpSystemFlux* / JxpSystemFluxSum* / JxpSystemLogLike*
//                         light_curves.limb_dark.light_curve(System, *u)(t) (+ transforms.integrate,
//                         + Normal(1 + lc, sigma).log_prob(y).sum())  src/jaxoplanet/light_curves/limb_dark.py:15-62,
//                                                                    light_curves/transforms.py:21-116
//   JxpTransitPartials* / JxpTransitGradient* (+ ...PolyF64: polynomial limb darkening of any length)
//                         jax.jvp / jax.grad of the above             orbits/keplerian.py:262-340, core/kepler.py:59-79
//   JxpOrbitalElementsF64 OrbitalBody.__init__                        src/jaxoplanet/orbits/keplerian.py:262-340
//   JxpOrbitState*        OrbitalBody.position / velocity families    src/jaxoplanet/orbits/keplerian.py:366-637
//   JxpPeerAllGatherF64   (no counterpart: the all_gather XLA inserts behind light_curves/limb_dark.py:62 when the
//                         parameter sets are sharded over devices) -- peer stores over NVLink, csrc/jxp_peer.cu
//   JxpTransitOrbit*      light_curve(TransitOrbit | TTVOrbit, *u)(t) src/jaxoplanet/orbits/transit.py:93-107, ttv.py:194-225
//   JxpInterpPeriodic*    transforms.interpolate(...)(t)              src/jaxoplanet/light_curves/transforms.py:163-183
//
// Build (jaxoplanet_b200/jax_bindings.py::build_ffi does this when `import jax` works):
//   g++ -std=c++17 -O2 -shared -fPIC -I"$(python -c 'import jax; print(jax.ffi.include_dir())')"
//       -I/usr/local/cuda/include -Iinclude jaxoplanet_b200/csrc/jxp_xla_ffi.cc
//       -Ljaxoplanet_b200/lib -ljxp_b200 -Wl,-rpath,'$ORIGIN' -o jaxoplanet_b200/lib/libjxp_b200_ffi.so
// JAX is not installable in the image this repository is developed in: there the translation unit is
// compiled against tests/xla_ffi_stub/ (the same class and macro names with the same call signatures,
// no behaviour), which checks every handler against its binding; the C ABI below it is what the parity
// tests exercise.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <optional>
#include <type_traits>

#include "../../include/jaxoplanet_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

template <ffi::DataType DT>
using Buf = ffi::Buffer<DT>;
template <ffi::DataType DT>
using Out = ffi::Result<ffi::Buffer<DT>>;

inline ffi::Error status(int rc) {
  if (rc == JXP_OK) return ffi::Error::Success();
  const ffi::ErrorCode code = rc == JXP_ERR_CUDA        ? ffi::ErrorCode::kInternal
                              : rc == JXP_ERR_WORKSPACE ? ffi::ErrorCode::kResourceExhausted
                              : rc == JXP_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                          : ffi::ErrorCode::kInvalidArgument;
  return ffi::Error(code, jxp_last_error());
}

inline ffi::Error bad(const char* msg) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, msg); }

template <typename B>
inline int64_t dim(const B& b, int k) {
  const auto d = b.dimensions();
  const int n = static_cast<int>(d.size());
  if (k < 0) k += n;
  return (k >= 0 && k < n) ? static_cast<int64_t>(d[k]) : 1;
}
template <typename B>
inline int rank(const B& b) {
  return static_cast<int>(b.dimensions().size());
}

// The workspace of a call, from XLA's scratch allocator (freed by XLA when the handler returns its
// stream work; the launchers only enqueue).
inline ffi::Error workspace(ffi::ScratchAllocator& scratch, size_t bytes, void** ws) {
  *ws = nullptr;
  if (bytes == 0) return ffi::Error::Success();
  std::optional<void*> p = scratch.Allocate(bytes, 256);
  if (!p.has_value() || *p == nullptr)
    return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaxoplanet_b200: scratch allocation failed");
  *ws = *p;
  return ffi::Error::Success();
}

// ---------------------------------------------------------------------------------------------
// K1  kepler(M, ecc) -> sin f, cos f, and the two factors of the reference's JVP rule
//     (vmap_method = "broadcast_all": M and ecc arrive with equal shapes)
// ---------------------------------------------------------------------------------------------
template <ffi::DataType DT, typename T, typename Fn>
ffi::Error kepler_impl(Fn fn, cudaStream_t stream, Buf<DT> M, Buf<DT> ecc, Out<DT> sinf, Out<DT> cosf, Out<DT> dfdM,
                       Out<DT> dfde) {
  const int64_t n = static_cast<int64_t>(M.element_count());
  const int64_t ne = static_cast<int64_t>(ecc.element_count());
  if (ne != n && ne != 1) return bad("jxp_kepler: M and ecc must have equal shapes (or ecc one element)");
  return status(fn(M.typed_data(), 1, ecc.typed_data(), ne == 1 ? 0 : 1, n, sinf->typed_data(), cosf->typed_data(),
                   static_cast<T*>(nullptr), dfdM->typed_data(), dfde->typed_data(), stream));
}
ffi::Error KeplerF64(cudaStream_t s, Buf<ffi::F64> M, Buf<ffi::F64> e, Out<ffi::F64> sf, Out<ffi::F64> cf,
                     Out<ffi::F64> dM, Out<ffi::F64> de) {
  return kepler_impl<ffi::F64, double>(jxp_kepler_f64, s, M, e, sf, cf, dM, de);
}
ffi::Error KeplerF32(cudaStream_t s, Buf<ffi::F32> M, Buf<ffi::F32> e, Out<ffi::F32> sf, Out<ffi::F32> cf,
                     Out<ffi::F32> dM, Out<ffi::F32> de) {
  return kepler_impl<ffi::F32, float>(jxp_kepler_f32, s, M, e, sf, cf, dM, de);
}

// ---------------------------------------------------------------------------------------------
// K2  core.limb_dark.light_curve(u, b, r, order) -> flux - 1, d/db, d/dr, solution vector [n_u + 1][n]
// ---------------------------------------------------------------------------------------------
template <ffi::DataType DT, typename Fn>
ffi::Error limb_dark_impl(Fn fn, cudaStream_t stream, Buf<DT> u, Buf<DT> b, Buf<DT> r, Out<DT> flux, Out<DT> dfdb,
                          Out<DT> dfdr, Out<DT> svec, int64_t order) {
  const int64_t n = static_cast<int64_t>(b.element_count());
  const int64_t nr = static_cast<int64_t>(r.element_count());
  const int64_t n_u = static_cast<int64_t>(u.element_count());
  if (rank(u) > 1) return bad("jxp_limb_dark: u must be one-dimensional (core/limb_dark.py:40)");
  if (nr != n && nr != 1) return bad("jxp_limb_dark: b and r must have equal shapes (or r one element)");
  if (static_cast<int64_t>(svec->element_count()) != (n_u + 1) * n) return bad("jxp_limb_dark: svec must be [n_u + 1][n]");
  return status(fn(u.typed_data(), static_cast<int32_t>(n_u), b.typed_data(), 1, r.typed_data(), nr == 1 ? 0 : 1, n,
                   static_cast<int32_t>(order), flux->typed_data(), dfdb->typed_data(), dfdr->typed_data(),
                   svec->typed_data(), stream));
}
ffi::Error LimbDarkF64(cudaStream_t s, Buf<ffi::F64> u, Buf<ffi::F64> b, Buf<ffi::F64> r, Out<ffi::F64> f,
                       Out<ffi::F64> db, Out<ffi::F64> dr, Out<ffi::F64> sv, int64_t order) {
  return limb_dark_impl<ffi::F64>(jxp_limb_dark_f64, s, u, b, r, f, db, dr, sv, order);
}
ffi::Error LimbDarkF32(cudaStream_t s, Buf<ffi::F32> u, Buf<ffi::F32> b, Buf<ffi::F32> r, Out<ffi::F32> f,
                       Out<ffi::F32> db, Out<ffi::F32> dr, Out<ffi::F32> sv, int64_t order) {
  return limb_dark_impl<ffi::F32>(jxp_limb_dark_f32, s, u, b, r, f, db, dr, sv, order);
}

// ---------------------------------------------------------------------------------------------
// K3/K4  light_curve(System, *u)(t) for a batch of parameter sets
//   bodies [n_sets][n_bodies][12] f64, u [n_sets or 1][n_u] f64, t [n_times] f64
// ---------------------------------------------------------------------------------------------
struct SystemShape {
  int32_t n_sets, n_bodies, n_u;
  int64_t u_stride, n_times;
};
inline ffi::Error system_shape(const Buf<ffi::F64>& bodies, const Buf<ffi::F64>& u, const Buf<ffi::F64>& t,
                               SystemShape* s) {
  if (rank(bodies) != 3 || dim(bodies, 2) != JXP_BODY_NFIELDS)
    return bad("jxp_system_light_curve: bodies must be [n_sets][n_bodies][12]");
  s->n_sets = static_cast<int32_t>(dim(bodies, 0));
  s->n_bodies = static_cast<int32_t>(dim(bodies, 1));
  s->n_u = static_cast<int32_t>(dim(u, -1));
  if (rank(u) == 0) s->n_u = 0;
  const int64_t u_sets = rank(u) >= 2 ? dim(u, 0) : 1;
  if (u_sets != 1 && u_sets != s->n_sets) return bad("jxp_system_light_curve: u must be [n_u] or [n_sets][n_u]");
  s->u_stride = u_sets == 1 ? 0 : s->n_u;
  s->n_times = static_cast<int64_t>(t.element_count());
  return ffi::Error::Success();
}

struct Exposure {
  double time;
  int64_t order, num_samples, gl_order, flags;
};

template <typename T, typename Fn>
ffi::Error system_call(Fn fn, cudaStream_t stream, ffi::ScratchAllocator& scratch, const Buf<ffi::F64>& bodies,
                       const Buf<ffi::F64>& u, const Buf<ffi::F64>& t, const Exposure& ex, T* flux, T* flux_sum,
                       const T* y, const T* sigma, int64_t sigma_stride, T* loglike) {
  SystemShape sh;
  if (ffi::Error e = system_shape(bodies, u, t, &sh); e.failure()) return e;
  const size_t wsb = jxp_system_workspace_bytes(sh.n_sets, sh.n_bodies, sh.n_times);
  void* ws = nullptr;
  if (ffi::Error e = workspace(scratch, wsb, &ws); e.failure()) return e;
  return status(fn(bodies.typed_data(), sh.n_sets, sh.n_bodies, u.typed_data(), sh.n_u, sh.u_stride, t.typed_data(),
                   sh.n_times, ex.time, static_cast<int32_t>(ex.order), static_cast<int32_t>(ex.num_samples),
                   static_cast<int32_t>(ex.gl_order), static_cast<int32_t>(ex.flags), flux, flux_sum, y, sigma,
                   sigma_stride, loglike, ws, wsb, stream));
}

#define JXP_EXPO_PARAMS double exposure_time, int64_t exp_order, int64_t num_samples, int64_t gl_order, int64_t flags
#define JXP_EXPO_ARGS \
  Exposure { exposure_time, exp_order, num_samples, gl_order, flags }

// per-body flux [n_sets][n_times][n_bodies]: the reference's return value with a leading set axis
ffi::Error SystemFluxF64(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> bodies, Buf<ffi::F64> u,
                         Buf<ffi::F64> t, Out<ffi::F64> flux, JXP_EXPO_PARAMS) {
  return system_call<double>(jxp_system_light_curve_f64, s, scratch, bodies, u, t, JXP_EXPO_ARGS, flux->typed_data(),
                             nullptr, nullptr, nullptr, 0, nullptr);
}
ffi::Error SystemFluxF32(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> bodies, Buf<ffi::F64> u,
                         Buf<ffi::F64> t, Out<ffi::F32> flux, JXP_EXPO_PARAMS) {
  return system_call<float>(jxp_system_light_curve_f32, s, scratch, bodies, u, t, JXP_EXPO_ARGS, flux->typed_data(),
                            nullptr, nullptr, nullptr, 0, nullptr);
}
// flux summed over bodies [n_sets][n_times]
ffi::Error SystemFluxSumF64(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> bodies, Buf<ffi::F64> u,
                            Buf<ffi::F64> t, Out<ffi::F64> flux_sum, JXP_EXPO_PARAMS) {
  return system_call<double>(jxp_system_light_curve_f64, s, scratch, bodies, u, t, JXP_EXPO_ARGS, nullptr,
                             flux_sum->typed_data(), nullptr, nullptr, 0, nullptr);
}
ffi::Error SystemFluxSumF32(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> bodies, Buf<ffi::F64> u,
                            Buf<ffi::F64> t, Out<ffi::F32> flux_sum, JXP_EXPO_PARAMS) {
  return system_call<float>(jxp_system_light_curve_f32, s, scratch, bodies, u, t, JXP_EXPO_ARGS, nullptr,
                            flux_sum->typed_data(), nullptr, nullptr, 0, nullptr);
}
// fused Gaussian log-likelihood [n_sets]; sigma one element or [n_times]
template <ffi::DataType DT, typename T, typename Fn>
ffi::Error loglike_impl(Fn fn, cudaStream_t s, ffi::ScratchAllocator& scratch, const Buf<ffi::F64>& bodies,
                        const Buf<ffi::F64>& u, const Buf<ffi::F64>& t, const Buf<DT>& y, const Buf<DT>& sigma,
                        Out<DT>& loglike, const Exposure& ex) {
  const int64_t n_times = static_cast<int64_t>(t.element_count());
  const int64_t n_sig = static_cast<int64_t>(sigma.element_count());
  if (static_cast<int64_t>(y.element_count()) != n_times) return bad("jxp_system_light_curve: y must have the shape of t");
  if (n_sig != 1 && n_sig != n_times) return bad("jxp_system_light_curve: sigma must be a scalar or have the shape of t");
  return system_call<T>(fn, s, scratch, bodies, u, t, ex, static_cast<T*>(nullptr), static_cast<T*>(nullptr),
                        y.typed_data(), sigma.typed_data(), n_sig == 1 ? 0 : 1, loglike->typed_data());
}
ffi::Error SystemLogLikeF64(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> bodies, Buf<ffi::F64> u,
                            Buf<ffi::F64> t, Buf<ffi::F64> y, Buf<ffi::F64> sigma, Out<ffi::F64> loglike,
                            JXP_EXPO_PARAMS) {
  return loglike_impl<ffi::F64, double>(jxp_system_light_curve_f64, s, scratch, bodies, u, t, y, sigma, loglike,
                                        JXP_EXPO_ARGS);
}
ffi::Error SystemLogLikeF32(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> bodies, Buf<ffi::F64> u,
                            Buf<ffi::F64> t, Buf<ffi::F32> y, Buf<ffi::F32> sigma, Out<ffi::F32> loglike,
                            JXP_EXPO_PARAMS) {
  return loglike_impl<ffi::F32, float>(jxp_system_light_curve_f32, s, scratch, bodies, u, t, y, sigma, loglike,
                                       JXP_EXPO_ARGS);
}

// ---------------------------------------------------------------------------------------------
// K5  flux + forward-mode partials w.r.t. theta = (P, t0, e, omega, b, r, u1, u2)
//   theta [n_sets][8] f64, star [n_sets or 1][2] f64, t [n_times] f64
// ---------------------------------------------------------------------------------------------
struct TransitShape {
  int32_t n_sets;
  int64_t star_stride, n_times;
};
inline ffi::Error transit_shape(const Buf<ffi::F64>& theta, const Buf<ffi::F64>& star, const Buf<ffi::F64>& t,
                                TransitShape* s) {
  if (rank(theta) != 2 || dim(theta, 1) != JXP_NTHETA) return bad("jxp_transit_partials: theta must be [n_sets][8]");
  s->n_sets = static_cast<int32_t>(dim(theta, 0));
  const int64_t ns = static_cast<int64_t>(star.element_count());
  if (ns != 2 && ns != 2 * static_cast<int64_t>(s->n_sets))
    return bad("jxp_transit_partials: star must be [2] or [n_sets][2] (central mass, central radius)");
  s->star_stride = ns == 2 ? 0 : 2;
  s->n_times = static_cast<int64_t>(t.element_count());
  return ffi::Error::Success();
}

template <typename T, typename Fn>
ffi::Error transit_call(Fn fn, cudaStream_t stream, ffi::ScratchAllocator& scratch, const Buf<ffi::F64>& theta,
                        const Buf<ffi::F64>& star, const Buf<ffi::F64>& t, const Exposure& ex, T* flux, T* partials,
                        const T* y, const T* sigma, int64_t sigma_stride, T* loglike, T* dloglike) {
  TransitShape sh;
  if (ffi::Error e = transit_shape(theta, star, t, &sh); e.failure()) return e;
  const size_t wsb = jxp_transit_workspace_bytes(sh.n_sets, sh.n_times);
  void* ws = nullptr;
  if (ffi::Error e = workspace(scratch, wsb, &ws); e.failure()) return e;
  return status(fn(theta.typed_data(), sh.n_sets, star.typed_data(), sh.star_stride, t.typed_data(), sh.n_times, ex.time,
                   static_cast<int32_t>(ex.order), static_cast<int32_t>(ex.num_samples),
                   static_cast<int32_t>(ex.gl_order), static_cast<int32_t>(ex.flags), flux, partials, y, sigma,
                   sigma_stride, loglike, dloglike, ws, wsb, stream));
}
// flux [n_sets][n_times] and partials [n_sets][8][n_times]: the primal and the JVP factors of the custom_jvp
ffi::Error TransitPartialsF64(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> theta, Buf<ffi::F64> star,
                              Buf<ffi::F64> t, Out<ffi::F64> flux, Out<ffi::F64> partials, JXP_EXPO_PARAMS) {
  return transit_call<double>(jxp_transit_partials_f64, s, scratch, theta, star, t, JXP_EXPO_ARGS, flux->typed_data(),
                              partials->typed_data(), nullptr, nullptr, 0, nullptr, nullptr);
}
ffi::Error TransitPartialsF32(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> theta, Buf<ffi::F64> star,
                              Buf<ffi::F64> t, Out<ffi::F32> flux, Out<ffi::F32> partials, JXP_EXPO_PARAMS) {
  return transit_call<float>(jxp_transit_partials_f32, s, scratch, theta, star, t, JXP_EXPO_ARGS, flux->typed_data(),
                             partials->typed_data(), nullptr, nullptr, 0, nullptr, nullptr);
}
// the gradient step NUTS needs: log-likelihood [n_sets] and its gradient [n_sets][8], nothing stored per sample
template <ffi::DataType DT, typename T, typename Fn>
ffi::Error gradient_impl(Fn fn, cudaStream_t s, ffi::ScratchAllocator& scratch, const Buf<ffi::F64>& theta,
                         const Buf<ffi::F64>& star, const Buf<ffi::F64>& t, const Buf<DT>& y, const Buf<DT>& sigma,
                         Out<DT>& loglike, Out<DT>& dloglike, const Exposure& ex) {
  const int64_t n_times = static_cast<int64_t>(t.element_count());
  const int64_t n_sig = static_cast<int64_t>(sigma.element_count());
  if (static_cast<int64_t>(y.element_count()) != n_times) return bad("jxp_transit_partials: y must have the shape of t");
  if (n_sig != 1 && n_sig != n_times) return bad("jxp_transit_partials: sigma must be a scalar or have the shape of t");
  return transit_call<T>(fn, s, scratch, theta, star, t, ex, static_cast<T*>(nullptr), static_cast<T*>(nullptr),
                         y.typed_data(), sigma.typed_data(), n_sig == 1 ? 0 : 1, loglike->typed_data(),
                         dloglike->typed_data());
}
ffi::Error TransitGradientF64(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> theta, Buf<ffi::F64> star,
                              Buf<ffi::F64> t, Buf<ffi::F64> y, Buf<ffi::F64> sigma, Out<ffi::F64> loglike,
                              Out<ffi::F64> dloglike, JXP_EXPO_PARAMS) {
  return gradient_impl<ffi::F64, double>(jxp_transit_partials_f64, s, scratch, theta, star, t, y, sigma, loglike,
                                         dloglike, JXP_EXPO_ARGS);
}
ffi::Error TransitGradientF32(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> theta, Buf<ffi::F64> star,
                              Buf<ffi::F64> t, Buf<ffi::F32> y, Buf<ffi::F32> sigma, Out<ffi::F32> loglike,
                              Out<ffi::F32> dloglike, JXP_EXPO_PARAMS) {
  return gradient_impl<ffi::F32, float>(jxp_transit_partials_f32, s, scratch, theta, star, t, y, sigma, loglike,
                                        dloglike, JXP_EXPO_ARGS);
}

// K5 for polynomial limb darkening of any length: theta [n_sets][6 + n_u] (n_u from the shape), double precision
inline ffi::Error transit_poly_call(cudaStream_t stream, ffi::ScratchAllocator& scratch, const Buf<ffi::F64>& theta,
                                    const Buf<ffi::F64>& star, const Buf<ffi::F64>& t, int64_t gl_order, int64_t flags,
                                    double* flux, double* partials, const double* y, const double* sigma,
                                    int64_t sigma_stride, double* loglike, double* dloglike) {
  if (rank(theta) != 2 || dim(theta, 1) < 6 || dim(theta, 1) > 6 + JXP_MAX_LD_COEFFS)
    return bad("jxp_transit_partials_poly: theta must be [n_sets][6 + n_u], 0 <= n_u <= 8");
  const int32_t n_sets = static_cast<int32_t>(dim(theta, 0));
  const int32_t n_u = static_cast<int32_t>(dim(theta, 1)) - 6;
  const int64_t ns = static_cast<int64_t>(star.element_count());
  if (ns != 2 && ns != 2 * static_cast<int64_t>(n_sets))
    return bad("jxp_transit_partials_poly: star must be [2] or [n_sets][2] (central mass, central radius)");
  const int64_t n_times = static_cast<int64_t>(t.element_count());
  const size_t wsb = jxp_transit_poly_workspace_bytes(n_sets, n_times);
  void* ws = nullptr;
  if (ffi::Error e = workspace(scratch, wsb, &ws); e.failure()) return e;
  return status(jxp_transit_partials_poly_f64(theta.typed_data(), n_sets, n_u, star.typed_data(), ns == 2 ? 0 : 2,
                                              t.typed_data(), n_times, static_cast<int32_t>(gl_order),
                                              static_cast<int32_t>(flags), flux, partials, y, sigma, sigma_stride,
                                              loglike, dloglike, ws, wsb, stream));
}
ffi::Error TransitPartialsPolyF64(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> theta, Buf<ffi::F64> star,
                                  Buf<ffi::F64> t, Out<ffi::F64> flux, Out<ffi::F64> partials, int64_t gl_order,
                                  int64_t flags) {
  return transit_poly_call(s, scratch, theta, star, t, gl_order, flags, flux->typed_data(), partials->typed_data(),
                           nullptr, nullptr, 0, nullptr, nullptr);
}
ffi::Error TransitGradientPolyF64(cudaStream_t s, ffi::ScratchAllocator scratch, Buf<ffi::F64> theta, Buf<ffi::F64> star,
                                  Buf<ffi::F64> t, Buf<ffi::F64> y, Buf<ffi::F64> sigma, Out<ffi::F64> loglike,
                                  Out<ffi::F64> dloglike, int64_t gl_order, int64_t flags) {
  const int64_t n_times = static_cast<int64_t>(t.element_count());
  const int64_t n_sig = static_cast<int64_t>(sigma.element_count());
  if (static_cast<int64_t>(y.element_count()) != n_times)
    return bad("jxp_transit_partials_poly: y must have the shape of t");
  if (n_sig != 1 && n_sig != n_times) return bad("jxp_transit_partials_poly: sigma must be a scalar or have the shape of t");
  return transit_poly_call(s, scratch, theta, star, t, gl_order, flags, nullptr, nullptr, y.typed_data(),
                           sigma.typed_data(), n_sig == 1 ? 0 : 1, loglike->typed_data(), dloglike->typed_data());
}

// ---------------------------------------------------------------------------------------------
// K0  OrbitalBody.__init__: element records [n][16] -> body records [n][12]
// ---------------------------------------------------------------------------------------------
ffi::Error OrbitalElementsF64(cudaStream_t s, Buf<ffi::F64> elements, Out<ffi::F64> bodies) {
  if (dim(elements, -1) != JXP_EL_NFIELDS) return bad("jxp_orbital_elements: elements must be [n][16]");
  const int64_t n = static_cast<int64_t>(elements.element_count()) / JXP_EL_NFIELDS;
  if (static_cast<int64_t>(bodies->element_count()) != n * JXP_BODY_NFIELDS)
    return bad("jxp_orbital_elements: bodies must be [n][12]");
  return status(jxp_orbital_elements_f64(elements.typed_data(), n, bodies->typed_data(), s));
}

// ---------------------------------------------------------------------------------------------
// The exchange of the set-sharded path: local [n_local] -> gathered [n_total] on every rank.  The exchange buffers
// (jxp_peer_alloc / jxp_peer_open, one per rank, set up once by the host program) come in as attributes: addresses
// as mapped into THIS process, up to eight ranks.  The epoch is counted on the device (argument 0), so the call has
// no state on the host and sits in a command buffer like any other.
// ---------------------------------------------------------------------------------------------
ffi::Error PeerAllGatherF64(cudaStream_t s, Buf<ffi::F64> local, Out<ffi::F64> gathered, int64_t offset, int64_t rank,
                            int64_t world, int64_t buf0, int64_t buf1, int64_t buf2, int64_t buf3, int64_t buf4,
                            int64_t buf5, int64_t buf6, int64_t buf7) {
  if (world < 1 || world > 8) return bad("jxp_peer_allgather: 1..8 ranks through this handler");
  const uint64_t bufs[8] = {(uint64_t)buf0, (uint64_t)buf1, (uint64_t)buf2, (uint64_t)buf3,
                            (uint64_t)buf4, (uint64_t)buf5, (uint64_t)buf6, (uint64_t)buf7};
  return status(jxp_peer_allgather_f64(local.typed_data(), static_cast<int64_t>(local.element_count()), offset,
                                       static_cast<int64_t>(gathered->element_count()), bufs, (int32_t)rank,
                                       (int32_t)world, 0, gathered->typed_data(), s));
}

// ---------------------------------------------------------------------------------------------
// K8  positions and velocities: bodies [n_sets][12], r_scale / k_amp [n_sets] or [1],
//     asc_node [n_sets][2] (cos, sin; pass has_asc_node = 0 and any [n_sets][2] array when the
//     ascending node is not given), t [n_times] -> pos, vel [3][n_sets][n_times]
// ---------------------------------------------------------------------------------------------
template <ffi::DataType DT, typename Fn>
ffi::Error state_impl(Fn fn, cudaStream_t s, const Buf<ffi::F64>& bodies, const Buf<ffi::F64>& r_scale,
                      const Buf<ffi::F64>& k_amp, const Buf<ffi::F64>& asc, const Buf<ffi::F64>& t, Out<DT>& pos,
                      Out<DT>& vel, int64_t has_asc_node, int64_t flags) {
  if (dim(bodies, -1) != JXP_BODY_NFIELDS) return bad("jxp_orbit_state: bodies must be [n_sets][12]");
  const int32_t n_sets = static_cast<int32_t>(bodies.element_count() / JXP_BODY_NFIELDS);
  const int64_t n_times = static_cast<int64_t>(t.element_count());
  const int64_t nr = static_cast<int64_t>(r_scale.element_count()), nk = static_cast<int64_t>(k_amp.element_count());
  if ((nr != 1 && nr != n_sets) || (nk != 1 && nk != n_sets)) return bad("jxp_orbit_state: r_scale / k_amp must be [n_sets] or one element");
  if (has_asc_node && static_cast<int64_t>(asc.element_count()) != 2 * static_cast<int64_t>(n_sets))
    return bad("jxp_orbit_state: asc_node must be [n_sets][2]");
  return status(fn(bodies.typed_data(), n_sets, r_scale.typed_data(), nr == 1 ? 0 : 1, k_amp.typed_data(),
                   nk == 1 ? 0 : 1, has_asc_node ? asc.typed_data() : nullptr, t.typed_data(), n_times,
                   static_cast<int32_t>(flags), pos->typed_data(), vel->typed_data(), s));
}
ffi::Error OrbitStateF64(cudaStream_t s, Buf<ffi::F64> bodies, Buf<ffi::F64> r_scale, Buf<ffi::F64> k_amp,
                         Buf<ffi::F64> asc, Buf<ffi::F64> t, Out<ffi::F64> pos, Out<ffi::F64> vel, int64_t has_asc_node,
                         int64_t flags) {
  return state_impl<ffi::F64>(jxp_orbit_state_f64, s, bodies, r_scale, k_amp, asc, t, pos, vel, has_asc_node, flags);
}
ffi::Error OrbitStateF32(cudaStream_t s, Buf<ffi::F64> bodies, Buf<ffi::F64> r_scale, Buf<ffi::F64> k_amp,
                         Buf<ffi::F64> asc, Buf<ffi::F64> t, Out<ffi::F32> pos, Out<ffi::F32> vel, int64_t has_asc_node,
                         int64_t flags) {
  return state_impl<ffi::F32>(jxp_orbit_state_f32, s, bodies, r_scale, k_amp, asc, t, pos, vel, has_asc_node, flags);
}

// ---------------------------------------------------------------------------------------------
// K6  light_curve(TransitOrbit | TTVOrbit, *u)(t): planets [n_planets][6]; TTV tables as in the header
//     (has_ttv = 0: plain TransitOrbit, the table arguments are ignored).  The C entry point takes the
//     limb-darkening coefficients from HOST memory; under XLA they are static: attributes u0..u7, n_u.
// ---------------------------------------------------------------------------------------------
template <ffi::DataType DT, typename Fn>
ffi::Error transit_orbit_impl(Fn fn, cudaStream_t s, const Buf<ffi::F64>& planets, const Buf<ffi::F64>& bin_edges,
                              const Buf<ffi::F64>& bin_values, const Buf<ffi::S32>& bin_offsets,
                              const Buf<ffi::F64>& ttv_t0, const Buf<ffi::F64>& t, Out<DT>& flux, int64_t has_ttv,
                              int64_t gl_order, int64_t n_u, const double (&uc)[JXP_MAX_LD_COEFFS]) {
  if (dim(planets, -1) != JXP_TO_NFIELDS) return bad("jxp_transit_orbit_light_curve: planets must be [n_planets][6]");
  if (n_u < 0 || n_u > JXP_MAX_LD_COEFFS) return bad("jxp_transit_orbit_light_curve: 0 <= n_u <= 8");
  const int32_t n_planets = static_cast<int32_t>(planets.element_count() / JXP_TO_NFIELDS);
  return status(fn(planets.typed_data(), n_planets, has_ttv ? bin_edges.typed_data() : nullptr,
                   has_ttv ? bin_values.typed_data() : nullptr, has_ttv ? bin_offsets.typed_data() : nullptr,
                   has_ttv ? ttv_t0.typed_data() : nullptr, uc, static_cast<int32_t>(n_u), t.typed_data(),
                   static_cast<int64_t>(t.element_count()), static_cast<int32_t>(gl_order), flux->typed_data(), s));
}
#define JXP_U_PARAMS double u0, double u1, double u2, double u3, double u4, double u5, double u6, double u7
#define JXP_U_ARRAY \
  { u0, u1, u2, u3, u4, u5, u6, u7 }
ffi::Error TransitOrbitF64(cudaStream_t s, Buf<ffi::F64> planets, Buf<ffi::F64> bin_edges, Buf<ffi::F64> bin_values,
                           Buf<ffi::S32> bin_offsets, Buf<ffi::F64> ttv_t0, Buf<ffi::F64> t, Out<ffi::F64> flux,
                           int64_t has_ttv, int64_t gl_order, int64_t n_u, JXP_U_PARAMS) {
  const double uc[JXP_MAX_LD_COEFFS] = JXP_U_ARRAY;
  return transit_orbit_impl<ffi::F64>(jxp_transit_orbit_light_curve_f64, s, planets, bin_edges, bin_values, bin_offsets,
                                      ttv_t0, t, flux, has_ttv, gl_order, n_u, uc);
}
ffi::Error TransitOrbitF32(cudaStream_t s, Buf<ffi::F64> planets, Buf<ffi::F64> bin_edges, Buf<ffi::F64> bin_values,
                           Buf<ffi::S32> bin_offsets, Buf<ffi::F64> ttv_t0, Buf<ffi::F64> t, Out<ffi::F32> flux,
                           int64_t has_ttv, int64_t gl_order, int64_t n_u, JXP_U_PARAMS) {
  const double uc[JXP_MAX_LD_COEFFS] = JXP_U_ARRAY;
  return transit_orbit_impl<ffi::F32>(jxp_transit_orbit_light_curve_f32, s, planets, bin_edges, bin_values, bin_offsets,
                                      ttv_t0, t, flux, has_ttv, gl_order, n_u, uc);
}

// ---------------------------------------------------------------------------------------------
// K7  transforms.interpolate(...)(t): xp [n] f64, fp [n][n_cols], t [n_times] -> out [n_times][n_cols]
// ---------------------------------------------------------------------------------------------
template <ffi::DataType DT, typename Fn>
ffi::Error interp_impl(Fn fn, cudaStream_t s, const Buf<ffi::F64>& xp, const Buf<DT>& fp, const Buf<ffi::F64>& t,
                       Out<DT>& out, double period, double time_transit) {
  const int32_t n = static_cast<int32_t>(xp.element_count());
  if (n <= 0 || fp.element_count() % static_cast<size_t>(n) != 0) return bad("jxp_interp_periodic: fp must be [n][n_cols]");
  const int32_t n_cols = static_cast<int32_t>(fp.element_count() / static_cast<size_t>(n));
  return status(fn(xp.typed_data(), fp.typed_data(), n, n_cols, period, time_transit, t.typed_data(),
                   static_cast<int64_t>(t.element_count()), out->typed_data(), s));
}
ffi::Error InterpPeriodicF64(cudaStream_t s, Buf<ffi::F64> xp, Buf<ffi::F64> fp, Buf<ffi::F64> t, Out<ffi::F64> out,
                             double period, double time_transit) {
  return interp_impl<ffi::F64>(jxp_interp_periodic_f64, s, xp, fp, t, out, period, time_transit);
}
ffi::Error InterpPeriodicF32(cudaStream_t s, Buf<ffi::F64> xp, Buf<ffi::F32> fp, Buf<ffi::F64> t, Out<ffi::F32> out,
                             double period, double time_transit) {
  return interp_impl<ffi::F32>(jxp_interp_periodic_f32, s, xp, fp, t, out, period, time_transit);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// bindings (the exported symbols jax.ffi.register_ffi_target takes through jax.ffi.pycapsule)
// ---------------------------------------------------------------------------------------------
#define JXP_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define JXP_STREAM_SCRATCH JXP_STREAM.Ctx<ffi::ScratchAllocator>()
#define JXP_EXPO_ATTRS                                                                                     \
  Attr<double>("exposure_time").Attr<int64_t>("exp_order").Attr<int64_t>("num_samples").Attr<int64_t>("gl_order") \
      .Attr<int64_t>("flags")
#define A64 Arg<ffi::Buffer<ffi::F64>>()
#define A32 Arg<ffi::Buffer<ffi::F32>>()
#define R64 Ret<ffi::Buffer<ffi::F64>>()
#define R32 Ret<ffi::Buffer<ffi::F32>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpKeplerF64, KeplerF64, JXP_STREAM.A64.A64.R64.R64.R64.R64);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpKeplerF32, KeplerF32, JXP_STREAM.A32.A32.R32.R32.R32.R32);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpLimbDarkF64, LimbDarkF64,
                              JXP_STREAM.A64.A64.A64.R64.R64.R64.R64.Attr<int64_t>("order"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpLimbDarkF32, LimbDarkF32,
                              JXP_STREAM.A32.A32.A32.R32.R32.R32.R32.Attr<int64_t>("order"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpSystemFluxF64, SystemFluxF64, JXP_STREAM_SCRATCH.A64.A64.A64.R64.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpSystemFluxF32, SystemFluxF32, JXP_STREAM_SCRATCH.A64.A64.A64.R32.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpSystemFluxSumF64, SystemFluxSumF64, JXP_STREAM_SCRATCH.A64.A64.A64.R64.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpSystemFluxSumF32, SystemFluxSumF32, JXP_STREAM_SCRATCH.A64.A64.A64.R32.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpSystemLogLikeF64, SystemLogLikeF64,
                              JXP_STREAM_SCRATCH.A64.A64.A64.A64.A64.R64.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpSystemLogLikeF32, SystemLogLikeF32,
                              JXP_STREAM_SCRATCH.A64.A64.A64.A32.A32.R32.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpTransitPartialsF64, TransitPartialsF64,
                              JXP_STREAM_SCRATCH.A64.A64.A64.R64.R64.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpTransitPartialsF32, TransitPartialsF32,
                              JXP_STREAM_SCRATCH.A64.A64.A64.R32.R32.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpTransitGradientF64, TransitGradientF64,
                              JXP_STREAM_SCRATCH.A64.A64.A64.A64.A64.R64.R64.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpTransitGradientF32, TransitGradientF32,
                              JXP_STREAM_SCRATCH.A64.A64.A64.A32.A32.R32.R32.JXP_EXPO_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpTransitPartialsPolyF64, TransitPartialsPolyF64,
                              JXP_STREAM_SCRATCH.A64.A64.A64.R64.R64.Attr<int64_t>("gl_order").Attr<int64_t>("flags"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpTransitGradientPolyF64, TransitGradientPolyF64,
                              JXP_STREAM_SCRATCH.A64.A64.A64.A64.A64.R64.R64.Attr<int64_t>("gl_order")
                                  .Attr<int64_t>("flags"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpOrbitalElementsF64, OrbitalElementsF64, JXP_STREAM.A64.R64);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpPeerAllGatherF64, PeerAllGatherF64,
                              JXP_STREAM.A64.R64.Attr<int64_t>("offset").Attr<int64_t>("rank").Attr<int64_t>("world")
                                  .Attr<int64_t>("buf0").Attr<int64_t>("buf1").Attr<int64_t>("buf2").Attr<int64_t>("buf3")
                                  .Attr<int64_t>("buf4").Attr<int64_t>("buf5").Attr<int64_t>("buf6").Attr<int64_t>("buf7"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpOrbitStateF64, OrbitStateF64,
                              JXP_STREAM.A64.A64.A64.A64.A64.R64.R64.Attr<int64_t>("has_asc_node").Attr<int64_t>("flags"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpOrbitStateF32, OrbitStateF32,
                              JXP_STREAM.A64.A64.A64.A64.A64.R32.R32.Attr<int64_t>("has_asc_node").Attr<int64_t>("flags"));
#define JXP_U_ATTRS                                                                                               \
  Attr<double>("u0").Attr<double>("u1").Attr<double>("u2").Attr<double>("u3").Attr<double>("u4").Attr<double>("u5") \
      .Attr<double>("u6").Attr<double>("u7")
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpTransitOrbitF64, TransitOrbitF64,
                              JXP_STREAM.A64.A64.A64.Arg<ffi::Buffer<ffi::S32>>().A64.A64.R64.Attr<int64_t>("has_ttv")
                                  .Attr<int64_t>("gl_order").Attr<int64_t>("n_u").JXP_U_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpTransitOrbitF32, TransitOrbitF32,
                              JXP_STREAM.A64.A64.A64.Arg<ffi::Buffer<ffi::S32>>().A64.A64.R32.Attr<int64_t>("has_ttv")
                                  .Attr<int64_t>("gl_order").Attr<int64_t>("n_u").JXP_U_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpInterpPeriodicF64, InterpPeriodicF64,
                              JXP_STREAM.A64.A64.A64.R64.Attr<double>("period").Attr<double>("time_transit"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxpInterpPeriodicF32, InterpPeriodicF32,
                              JXP_STREAM.A64.A32.A64.R32.Attr<double>("period").Attr<double>("time_transit"));
