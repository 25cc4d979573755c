// Host-side helpers shared by the translation units of libjxp_b200 (not part of the ABI).
#pragma once
#include <cstddef>
#include <cstdint>

#include <cuda_runtime.h>

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
uint32_t dev_options();  // jxp_set_dev_options mask (JXP_DEV_*)
void gl_table_host(int order, jxp::GLTable& tab);
int make_stencil(double exposure_time, int order, int num_samples, Stencil& st);
constexpr int LL_NCTA_H = 64;   // CTAs of k_loglike_prep (jxp_kernels.cu)
constexpr int LL_SCAL0_H = 8;   // first per-CTA partial in the scalars block
constexpr int LL_FLAG0_H = LL_SCAL0_H + 2 * LL_NCTA_H;  // per-CTA "time stamps not sorted" flags
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- workspace carving of jxp_system_light_curve, shared by the size query and the launchers ----
constexpr int SYS_K_H = 4;          // time samples per thread of k_system (smallest tile = 32 * SYS_K_H)
constexpr int WALK_REC = 32;        // doubles per record of the transit-walk path (jxp_walk.cu)
constexpr int WALK_STATE0 = 200;    // scalars[200..]: segment / chunk allocation counters of the walk path
constexpr int WALK_TAB = 440;       // doubles per set of the walk path's function tables (jxp_tables.cuh)
constexpr int WALK_TSTAT0 = 216;    // scalars[216..219]: table statistics of the last call (jxp_system_walk_stats)
struct SysLayout {
  size_t off_bodies, off_sets, off_winv, off_scalars, off_partial, off_tilectr, total;
  size_t off_tl_off, off_tl_cnt, off_tl_items;  // transit-list path (single-body systems)
  unsigned long long tl_cap;                     // items the list can hold
  // transit-walk path: records, per-set bookkeeping, segment table, chunk table, per-chunk partial sums
  size_t off_wk_rec, off_wk_set, off_wk_seg, off_wk_chunk, off_wk_part, off_wk_tab, off_wk_items, off_wk_ipart, off_wk_fix, off_wk_fixsum, off_wk_aw;
  unsigned long long wk_cap_seg, wk_cap_chunk, wk_cap_items, wk_cap_fix;
  int tile, n_tiles;
};

inline SysLayout sys_layout(int n_sets, int n_bodies, int64_t n_times, size_t body_sz, size_t set_sz,
                            size_t elem_sz) {
  SysLayout L;
  L.tile = 32 * SYS_K_H;  // smallest tile any launch configuration uses (one warp per CTA)
  L.n_tiles = (int)((n_times + L.tile - 1) / L.tile);
  size_t off = 0;
  L.off_bodies = off;
  off = align_up(off + body_sz * (size_t)n_sets * n_bodies, 256);
  L.off_sets = off;
  off = align_up(off + set_sz * (size_t)n_sets, 256);
  L.off_winv = off;
  off = align_up(off + elem_sz * (size_t)n_times, 256);
  L.off_scalars = off;
  off = align_up(off + 256 * sizeof(double), 256);
  L.off_partial = off;
  off = align_up(off + sizeof(double) * (size_t)n_sets * L.n_tiles, 256);
  L.off_tilectr = off;  // [0]: next unclaimed set group, [1 + g]: next warp tile of group g
  off = align_up(off + sizeof(int) * ((size_t)n_sets + 1), 256);
  // transit list: one 8-byte slot per in-window (set, sample); room for 1/8 of the grid (a list
  // that would not fit makes the call fall back to the window-test kernels, on the device)
  L.off_tl_off = L.off_tl_cnt = L.off_tl_items = off;
  L.tl_cap = 0;
#ifdef JXP_KEEP_OLD_LIST
  if (n_bodies == 1) {
    L.off_tl_off = off;  // per set: first item on the inside list, on the edge list
    off = align_up(off + 2 * sizeof(unsigned long long) * (size_t)n_sets, 256);
    L.off_tl_cnt = off;  // per set: items on the inside list, on the edge list
    off = align_up(off + 2 * sizeof(unsigned) * (size_t)n_sets, 256);
    L.off_tl_items = off;
    L.tl_cap = (unsigned long long)n_sets * (unsigned long long)n_times / 8ull + 65536ull;
    if (L.tl_cap > (1ull << 28)) L.tl_cap = 1ull << 28;  // at most 2 GB of list; beyond it the call falls back
    off = align_up(off + sizeof(unsigned long long) * (size_t)L.tl_cap, 256);
  }
#endif
  // transit walk (jxp_walk.cu): nothing per item -- a record and a few dozen (start, count) segments per
  // (set, body), one 16-byte entry and one partial sum per chunk of in-window samples.  Room for 1/8 of
  // the grid in chunks of 32 and 384 segments per (set, body) on average; a call that needs more falls
  // back to the window-test kernels, on the device.
  {
    const size_t nb = (size_t)n_sets * n_bodies;
    L.off_wk_rec = off;
    off = align_up(off + sizeof(double) * WALK_REC * nb, 256);
    L.off_wk_set = off;
    off = align_up(off + 32 * nb, 256);
    // segments: up to 3 per transit and body (per-body launches share the table).  384 per (set, body) covers
    // long baselines of single planets; a short-period planet on a long 30-min baseline (config 4: 2e5 samples
    // = 4167 days, ~2300 transits of the inner planet) has one transit per ~40 samples, hence the n_times term
    L.wk_cap_seg = 384ull * nb + 4096ull + (unsigned long long)n_sets * (unsigned long long)n_times / 24ull;
    L.off_wk_seg = off;
    off = align_up(off + 8 * (size_t)L.wk_cap_seg, 256);
    L.wk_cap_chunk = (unsigned long long)n_sets * (unsigned long long)n_times / 256ull + 2ull * nb + 1024ull;
    if (L.wk_cap_chunk > (1ull << 27)) L.wk_cap_chunk = 1ull << 27;
    L.off_wk_chunk = off;
    off = align_up(off + 16 * (size_t)L.wk_cap_chunk, 256);
    L.off_wk_part = off;
    off = align_up(off + 8 * (size_t)L.wk_cap_chunk, 256);
    // function tables: per set (the per-body launches of one call reuse them), 2.9 KB each
    L.off_wk_tab = off;
    off = align_up(off + sizeof(double) * WALK_TAB * (size_t)n_sets, 256);
    // items of the sets with tables: 16 bytes + one partial sum per slot of a transit (only transits of >= 12
    // samples become items; room for 1/21 of the grid in such transits -- more falls back, on the device)
    L.wk_cap_items = 64ull * (unsigned long long)n_sets + (unsigned long long)n_sets * (unsigned long long)n_times / 256ull + 4096ull;
    if (L.wk_cap_items > (1ull << 27)) L.wk_cap_items = 1ull << 27;
    L.off_wk_items = off;
    off = align_up(off + 16 * (size_t)L.wk_cap_items, 256);
    L.off_wk_ipart = off;
    off = align_up(off + 8 * (size_t)L.wk_cap_items, 256);
    // samples of table sets that the tables do not answer (k_walk_fix): 16 bytes each, one fixed-point sum per set
    L.wk_cap_fix = 64ull * (unsigned long long)n_sets + 65536ull;
    L.off_wk_fix = off;
    off = align_up(off + 16 * (size_t)L.wk_cap_fix, 256);
    L.off_wk_fixsum = off;
    off = align_up(off + 16 * (size_t)n_sets, 256);
    L.off_wk_aw = off;  // (2 (y - 1) / sigma^2, 1 / sigma^2) per sample
    off = align_up(off + 16 * (size_t)n_times, 256);
  }
  L.total = off;
  return L;
}
// transit walk (jxp_walk.cu); ev_which: 0 = profile events around the evaluation, 1 = around the plan
int launch_walk_f64(const double* bodies, int n_sets, int n_bodies, const double* u, int n_u, int64_t u_stride,
                    const double* t, int64_t n_times, const Stencil& stencil, double* flux, double* flux_sum,
                    const double* y, const double* sigma, int64_t sigma_stride, double* loglike,
                    unsigned char* ws, const SysLayout& L, double widen, double margin, cudaStream_t st,
                    cudaEvent_t ev_start, cudaEvent_t ev_stop, int ev_which);
}  // namespace jxph
