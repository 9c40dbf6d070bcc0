// comm.cu -- the exchange steps of the row-sharded build (one process per GPU, NCCL over NVLink).
//
// The reference is single-process (rayon threads; SURVEY.md section 5 "distributed comm backend:
// none").  Sharding query rows across GPUs leaves two real exchanges on the path:
//   1. all-gather of the kNN lists (idx u32, dist f64, count u32) before symmetrisation, because the
//      reverse edges of a row live in other ranks' lists (src_legacy/laplacian.rs:305-345);
//   2. min / max all-reduce + all-gather of lambda, because normalise_lambdas is a global min-max
//      (src_legacy/core.rs:1341-1355).
// libnccl is loaded with dlopen on first use so that single-GPU users carry no NCCL dependency.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>

#include <new>

#include "common.cuh"

int32_t sfb_knn_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t k, sfb_knn** out);
int32_t sfb_lambda_device(sfb_ctx* ctx, const sfb_csr* L, const double* x_dev, uint64_t n, uint32_t f,
                          const sfb_lambda_params* prm, double* d_lambda, double* d_disp, const double* tau_in, double* d_mm);
int32_t sfb_normalise_device(sfb_ctx* ctx, double* d_lambda, uint64_t n, const double* d_mm);

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
            api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
            if (api.GetUniqueId && api.CommInitRank && api.AllGather && api.AllReduce && api.GetErrorString && api.GroupStart && api.GroupEnd) api.handle = h;
        }
    }
    return api.handle ? &api : nullptr;
}

#define SFB_NCCL(ctx, call)                                                                              \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) return sfb_fail((ctx), SFB_ENCCL, "%s: %s", #call, nccl()->GetErrorString(r__)); \
    } while (0)

}  // namespace

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in the ABI");

extern "C" int32_t sfb_comm_unique_id(uint8_t id[128]) {
    if (!id) return SFB_EINVAL;
    NcclApi* n = nccl();
    if (!n) return SFB_ENCCL;
    ncclUniqueId u;
    if (n->GetUniqueId(&u) != ncclSuccess) return SFB_ENCCL;
    memcpy(id, &u, 128);
    return SFB_OK;
}

extern "C" int32_t sfb_comm_init(sfb_ctx* ctx, const uint8_t id[128], int32_t rank, int32_t world) {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return sfb_fail(ctx, SFB_EINVAL, "bad rank/world");
    NcclApi* n = nccl();
    if (!n) return sfb_fail(ctx, SFB_ENCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
    SFB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t comm;
    SFB_NCCL(ctx, n->CommInitRank(&comm, world, u, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank; ctx->world = world;
    return SFB_OK;
}

void sfb_comm_destroy(sfb_ctx* ctx) {
    if (ctx && ctx->nccl_comm && nccl() && nccl()->CommDestroy) { nccl()->CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
}

int32_t sfb_comm_allreduce_sum_f64(sfb_ctx* ctx, double* buf, size_t n) {
    if (ctx->world == 1) return SFB_OK;
    if (!ctx->nccl_comm) return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised");
    StageTimer tc(ctx, &ctx->times.ms_comm);
    SFB_NCCL(ctx, nccl()->AllReduce(buf, buf, n, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return SFB_OK;
}

int32_t sfb_comm_allreduce_sum_f32(sfb_ctx* ctx, float* buf, size_t n) {
    if (ctx->world == 1) return SFB_OK;
    if (!ctx->nccl_comm) return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised");
    SFB_NCCL(ctx, nccl()->AllReduce(buf, buf, n, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return SFB_OK;
}

// in-place all-gather of equal byte slots: rank r's slot is base + r * bytes_per_rank
int32_t sfb_comm_allgather_bytes(sfb_ctx* ctx, void* base, size_t bytes_per_rank) {
    if (ctx->world == 1) return SFB_OK;
    if (!ctx->nccl_comm) return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised");
    StageTimer tc(ctx, &ctx->times.ms_comm);
    SFB_NCCL(ctx, nccl()->AllGather((const char*)base + (size_t)ctx->rank * bytes_per_rank, base, bytes_per_rank, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return SFB_OK;
}
int32_t sfb_comm_allreduce_max_u64(sfb_ctx* ctx, unsigned long long* buf, size_t n) {
    if (ctx->world == 1) return SFB_OK;
    if (!ctx->nccl_comm) return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised");
    StageTimer tc(ctx, &ctx->times.ms_comm);
    SFB_NCCL(ctx, nccl()->AllReduce(buf, buf, n, ncclUint64, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return SFB_OK;
}

extern "C" int32_t sfb_comm_barrier(sfb_ctx* ctx) {
    if (!ctx) return SFB_EINVAL;
    if (ctx->world == 1) { SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return SFB_OK; }
    if (!ctx->nccl_comm) return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised");
    DevBuf b;
    SFB_CUDA(ctx, b.alloc(8));
    SFB_CUDA(ctx, cudaMemsetAsync(b.p, 0, 8, ctx->stream));
    SFB_NCCL(ctx, nccl()->AllReduce(b.p, b.p, 1, ncclInt32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}

// Every rank holds rows [rank*S, min((rank+1)*S, total)) with S = ceil(total / world); shards are
// padded to S rows so a plain ncclAllGather (equal counts) assembles rows [0, world*S) in order.
extern "C" int32_t sfb_knn_allgather(sfb_ctx* ctx, const sfb_knn* shard, uint64_t total_rows, sfb_knn** out) {
    if (!ctx || !shard || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    const uint64_t world = (uint64_t)ctx->world, S = (total_rows + world - 1) / world;
    const uint64_t lo = (uint64_t)ctx->rank * S, hi = lo + S < total_rows ? lo + S : total_rows;
    if (shard->q_begin != (lo < total_rows ? lo : shard->q_begin) || shard->rows != (hi > lo ? hi - lo : 0) || shard->total != total_rows)
        return sfb_fail(ctx, SFB_EINVAL, "rank %d must hold rows [%llu, %llu) of %llu", ctx->rank, (unsigned long long)lo, (unsigned long long)hi, (unsigned long long)total_rows);
    const uint32_t k = shard->k;
    SFB_TRY(sfb_knn_alloc(ctx, world * S, k, out));
    sfb_knn* g = *out;
    g->rows = total_rows; g->total = total_rows; g->q_begin = 0; g->stats = shard->stats;
    if (world == 1) {
        SFB_CUDA(ctx, cudaMemcpyAsync(g->idx, shard->idx, sizeof(uint32_t) * shard->rows * k, cudaMemcpyDeviceToDevice, ctx->stream));
        SFB_CUDA(ctx, cudaMemcpyAsync(g->dist, shard->dist, sizeof(double) * shard->rows * k, cudaMemcpyDeviceToDevice, ctx->stream));
        SFB_CUDA(ctx, cudaMemcpyAsync(g->cnt, shard->cnt, sizeof(uint32_t) * shard->rows, cudaMemcpyDeviceToDevice, ctx->stream));
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SFB_OK;
    }
    if (!ctx->nccl_comm) { sfb_knn_free(g); *out = nullptr; return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised"); }
    // in-place all-gather: each rank's send buffer is its own slot of the receive buffer
    uint32_t* my_idx = g->idx + lo * k; double* my_dist = g->dist + lo * k; uint32_t* my_cnt = g->cnt + lo;
    SFB_CUDA(ctx, cudaMemsetAsync(my_idx, 0xFF, sizeof(uint32_t) * S * k, ctx->stream));
    SFB_CUDA(ctx, cudaMemsetAsync(my_dist, 0, sizeof(double) * S * k, ctx->stream));
    SFB_CUDA(ctx, cudaMemsetAsync(my_cnt, 0, sizeof(uint32_t) * S, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(my_idx, shard->idx, sizeof(uint32_t) * shard->rows * k, cudaMemcpyDeviceToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(my_dist, shard->dist, sizeof(double) * shard->rows * k, cudaMemcpyDeviceToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(my_cnt, shard->cnt, sizeof(uint32_t) * shard->rows, cudaMemcpyDeviceToDevice, ctx->stream));
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    // the three lists in one grouped call: one launch, one pass over the NVLink rings
    StageTimer tc(ctx, &ctx->times.ms_comm);
    SFB_NCCL(ctx, nccl()->GroupStart());
    ncclResult_t r1 = nccl()->AllGather(my_idx, g->idx, S * k, ncclUint32, comm, ctx->stream);
    ncclResult_t r2 = nccl()->AllGather(my_dist, g->dist, S * k, ncclFloat64, comm, ctx->stream);
    ncclResult_t r3 = nccl()->AllGather(my_cnt, g->cnt, S, ncclUint32, comm, ctx->stream);
    SFB_NCCL(ctx, nccl()->GroupEnd());
    if (r1 != ncclSuccess || r2 != ncclSuccess || r3 != ncclSuccess) return sfb_fail(ctx, SFB_ENCCL, "ncclAllGather of the kNN lists failed");
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}

// The corpus over NVLink instead of PCIe: every rank uploads only its own row shard and the full matrix is
// assembled on every GPU by one all-gather (north_star: "the corpus is ring-passed or replicated over NVLink").
extern "C" int32_t sfb_mat_allgather_rows(sfb_ctx* ctx, const sfb_mat* shard, uint64_t total_rows, sfb_mat** out) {
    if (!ctx || !shard || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    const uint64_t world = (uint64_t)ctx->world, S = (total_rows + world - 1) / world;
    const uint64_t lo = (uint64_t)ctx->rank * S < total_rows ? (uint64_t)ctx->rank * S : total_rows;
    const uint64_t hi = lo + S < total_rows ? lo + S : total_rows;
    if (shard->rows != hi - lo) return sfb_fail(ctx, SFB_EINVAL, "rank %d must hold rows [%llu, %llu) of %llu", ctx->rank, (unsigned long long)lo, (unsigned long long)hi, (unsigned long long)total_rows);
    if (world > 1 && !ctx->nccl_comm) return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised");
    sfb_mat* m = new (std::nothrow) sfb_mat();
    if (!m) return SFB_ENOMEM;
    m->ctx = ctx; m->rows = total_rows; m->cols = shard->cols;
    const size_t row_bytes = sizeof(double) * shard->cols;
    cudaError_t e = sfb_dev_alloc(ctx, (void**)&m->d, world * S * row_bytes);
    if (e != cudaSuccess) { delete m; return sfb_fail(ctx, SFB_ENOMEM, "all-gathered matrix: %s", cudaGetErrorString(e)); }
    double* mine = m->d + lo * shard->cols;
    StageTimer t(ctx, &ctx->times.ms_h2d);
    int32_t st = SFB_OK;
    if (cudaMemcpyAsync(mine, shard->d, shard->rows * row_bytes, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "shard copy failed");
    if (st == SFB_OK && world > 1) {
        ncclResult_t r = nccl()->AllGather(m->d + (uint64_t)ctx->rank * S * shard->cols, m->d, S * shard->cols, ncclFloat64, (ncclComm_t)ctx->nccl_comm, ctx->stream);
        if (r != ncclSuccess) st = sfb_fail(ctx, SFB_ENCCL, "ncclAllGather: %s", nccl()->GetErrorString(r));
    }
    if (st == SFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "matrix all-gather failed");
    if (st != SFB_OK) { sfb_mat_free(m); return st; }
    *out = m;
    return SFB_OK;
}

// x holds this rank's rows [row0, row0 + x->rows) of the total_rows items (same ceil-split as above).
extern "C" int32_t sfb_lambda_allgather(sfb_ctx* ctx, const sfb_csr* L, const sfb_mat* x, uint64_t row0, uint64_t total_rows,
                                        const sfb_lambda_params* prm, double* out_lambda, double* stats) {
    if (!ctx || !L || !x || !prm || !out_lambda) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    const uint64_t world = (uint64_t)ctx->world, S = (total_rows + world - 1) / world;
    const uint64_t lo = (uint64_t)ctx->rank * S, hi = lo + S < total_rows ? lo + S : total_rows;
    if (row0 != lo || x->rows != hi - lo) return sfb_fail(ctx, SFB_EINVAL, "rank %d must hold rows [%llu, %llu)", ctx->rank, (unsigned long long)lo, (unsigned long long)hi);
    if (world > 1 && !ctx->nccl_comm) return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised");
    double h_mm[2] = {0.0, 0.0};
    StageTimer t(ctx, &ctx->times.ms_lambda);
    DevBuf all, mm;
    SFB_CUDA(ctx, all.alloc(sizeof(double) * world * S));
    SFB_CUDA(ctx, mm.alloc(sizeof(double) * 2));
    double* mine = all.as<double>() + lo;
    if (x->rows < S) SFB_CUDA(ctx, cudaMemsetAsync(mine + x->rows, 0, sizeof(double) * (S - x->rows), ctx->stream));
    ctx->lambda_sharded = true;    // CORE_F32SEM: the energy total is all-reduced inside (every rank makes the same calls)
    int32_t st = sfb_lambda_device(ctx, L, x->d, x->rows, x->cols, prm, mine, nullptr, nullptr, mm.as<double>());
    ctx->lambda_sharded = false;
    SFB_TRY(st);
    if (world > 1) {   // global min / max(0, .) (core.rs:1345-1346), on the device values: no host round trip
        StageTimer tc(ctx, &ctx->times.ms_comm);
        ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
        SFB_NCCL(ctx, nccl()->GroupStart());
        ncclResult_t r1 = nccl()->AllReduce(mm.p, mm.p, 1, ncclFloat64, ncclMin, comm, ctx->stream);
        ncclResult_t r2 = nccl()->AllReduce(mm.as<double>() + 1, mm.as<double>() + 1, 1, ncclFloat64, ncclMax, comm, ctx->stream);
        SFB_NCCL(ctx, nccl()->GroupEnd());
        if (r1 != ncclSuccess || r2 != ncclSuccess) return sfb_fail(ctx, SFB_ENCCL, "min / max all-reduce failed");
    }
    if (prm->normalise_minmax) SFB_TRY(sfb_normalise_device(ctx, mine, x->rows, mm.as<double>()));
    if (world > 1) SFB_NCCL(ctx, nccl()->AllGather(mine, all.p, S, ncclFloat64, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    t.stop();
    StageTimer t2(ctx, &ctx->times.ms_d2h);
    SFB_CUDA(ctx, cudaMemcpyAsync(out_lambda, all.p, sizeof(double) * total_rows, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(h_mm, mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, ctx->stream));
    t2.stop();
    if (stats) { stats[0] = h_mm[0]; stats[1] = h_mm[1]; stats[2] = (h_mm[1] - h_mm[0]) > 1e-9 ? (h_mm[1] - h_mm[0]) : 1e-9; }
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}
