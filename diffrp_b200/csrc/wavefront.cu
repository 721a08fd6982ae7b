// wavefront.cu -- fused wavefront path tracer (placeholder; filled in next)
#include "internal.h"
void drp_free_workspace(BvhHandle*) {}
extern "C" int drp_render(uint64_t, const drp_scene_t*, const drp_render_params_t*, float*, void*) { drp_set_error("drp_render: not implemented"); return DRP_ERR_INVALID; }
extern "C" int drp_finalize(const float*, int32_t, int32_t, int32_t, float*, float*, float*, float*, float*, float*, void*) { drp_set_error("drp_finalize: not implemented"); return DRP_ERR_INVALID; }
extern "C" int drp_render_stats(uint64_t, drp_render_stats_t*) { return DRP_ERR_INVALID; }
