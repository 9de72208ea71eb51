// ivs_fused.cu -- placeholder while the fused kernel is being brought up
#include "ivs_internal.h"
ivs_status ivs_fused_plan_create(ivs_ctx* ctx, ivs_sim*, ivs_fused_plan*) {
    return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "fused step kernel not built yet: pass IVS_SIM_UNFUSED");
}
void ivs_fused_plan_destroy(ivs_fused_plan*) {}
ivs_status ivs_fused_enqueue_step(ivs_ctx* ctx, ivs_fused_plan*, ivs_sim*) {
    return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "fused step kernel not built yet");
}
