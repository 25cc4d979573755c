// Host-side helpers shared by the translation units of libjxp_b200 (not part of the ABI).
#pragma once
#include <cstddef>
#include <cstdint>

#include "jxp_math.cuh"

namespace jxph {
constexpr int MAX_SUB = 64;
struct Stencil {
  int n;
  double dt_min, dt_max;  // extreme offsets (<= 0, >= 0)
  double dt[MAX_SUB];
  double w[MAX_SUB];
};
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
bool stride_ok(int64_t s);
int sm_count();
void gl_table_host(int order, jxp::GLTable& tab);
int make_stencil(double exposure_time, int order, int num_samples, Stencil& st);
constexpr int LL_NCTA_H = 64;   // CTAs of k_loglike_prep (jxp_kernels.cu)
constexpr int LL_SCAL0_H = 8;   // first per-CTA partial in the scalars block
constexpr int LL_FLAG0_H = LL_SCAL0_H + 2 * LL_NCTA_H;  // per-CTA "time stamps not sorted" flags
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace jxph
