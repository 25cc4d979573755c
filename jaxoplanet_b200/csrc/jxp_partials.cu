// K5 placeholder (replaced by the forward-mode partials kernel).
#include "../../include/jaxoplanet_b200.h"
extern "C" size_t jxp_transit_workspace_bytes(int32_t, int64_t) { return 0; }
extern "C" int jxp_transit_partials_f64(const double*, int32_t, const double*, int64_t, const double*, int64_t, double, int32_t, int32_t, int32_t, int32_t, double*, double*, const double*, const double*, int64_t, double*, double*, void*, size_t, void*) { return JXP_ERR_UNSUPPORTED; }
extern "C" int jxp_transit_partials_f32(const double*, int32_t, const double*, int64_t, const double*, int64_t, double, int32_t, int32_t, int32_t, int32_t, float*, float*, const float*, const float*, int64_t, float*, float*, void*, size_t, void*) { return JXP_ERR_UNSUPPORTED; }
