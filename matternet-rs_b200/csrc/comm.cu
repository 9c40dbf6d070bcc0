// comm.cu -- multi-GPU exchange steps (placeholder until NCCL is wired).
#include "common.cuh"

extern "C" int32_t sfb_comm_unique_id(uint8_t*) { return SFB_EUNSUPPORTED; }
extern "C" int32_t sfb_comm_init(sfb_ctx* ctx, const uint8_t*, int32_t, int32_t) { return sfb_fail(ctx, SFB_EUNSUPPORTED, "NCCL not wired yet"); }
extern "C" int32_t sfb_knn_allgather(sfb_ctx* ctx, const sfb_knn*, uint64_t, sfb_knn**) { return sfb_fail(ctx, SFB_EUNSUPPORTED, "NCCL not wired yet"); }
extern "C" int32_t sfb_lambda_allgather(sfb_ctx* ctx, const sfb_csr*, const sfb_mat*, uint64_t, uint64_t, const sfb_lambda_params*, double*, double*) { return sfb_fail(ctx, SFB_EUNSUPPORTED, "NCCL not wired yet"); }
extern "C" int32_t sfb_comm_barrier(sfb_ctx* ctx) { return sfb_fail(ctx, SFB_EUNSUPPORTED, "NCCL not wired yet"); }
