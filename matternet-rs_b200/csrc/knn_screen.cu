// knn_screen.cu -- tensor-core distance screen (placeholder until the tcgen05 kernel lands).
#include "common.cuh"

int32_t sfb_knn_screened(sfb_ctx* ctx, const sfb_mat*, const double*, const sfb_knn_params*, uint64_t, uint64_t, sfb_knn*) {
    return sfb_fail(ctx, SFB_EUNSUPPORTED, "tensor-core screen not built yet");
}
