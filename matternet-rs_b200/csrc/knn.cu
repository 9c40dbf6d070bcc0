// knn.cu -- sfb_knn_build: dispatch between the exact f64 path and the tensor-core screen.
#include <math.h>

#include <new>

#include "common.cuh"

int32_t sfb_knn_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t k, sfb_knn** out);

int32_t sfb_knn_dense(sfb_ctx* ctx, const double* xd, uint32_t m, uint64_t kd, int metric, uint32_t k, double eps,
                      uint64_t q_begin, uint64_t nq, uint32_t* out_idx, double* out_dist, uint32_t* out_cnt, int collective);
bool sfb_dense_shape(uint64_t nodes, uint64_t dims);
int32_t sfb_transpose_device(sfb_ctx* ctx, const double* a, uint64_t rows, uint64_t cols, double* b);

// `x` holds the nodes as ROWS (columns_are_nodes == false) or as COLUMNS (true: x is dims x nodes, the
// untransposed item matrix of GraphFactory::build_laplacian_matrix_from_k_cluster, graph.rs:193-216).
static int32_t knn_build_any(sfb_ctx* ctx, const sfb_mat* x, bool columns_are_nodes, const sfb_knn_params* p, sfb_knn** out, int collective) {
    if (!ctx || !x || !p || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    const uint64_t nodes = columns_are_nodes ? x->cols : x->rows, dims = columns_are_nodes ? x->rows : x->cols;
    // the reference asserts n >= 2 && d >= 2 (src_legacy/laplacian.rs:130-135)
    if (nodes < 2) return sfb_fail(ctx, SFB_EINVAL, "need at least 2 nodes (got %llu)", (unsigned long long)nodes);
    if (dims > 0xFFFFFFFFull && !sfb_dense_shape(nodes, dims)) return sfb_fail(ctx, SFB_EUNSUPPORTED, "dimension count must fit u32");
    if (p->metric < SFB_METRIC_COSINE || p->metric > SFB_METRIC_L2SQ) return sfb_fail(ctx, SFB_EINVAL, "unknown metric %d", p->metric);
    if (p->k == 0 || p->k > 128) return sfb_fail(ctx, SFB_EUNSUPPORTED, "k must be in 1..128 (got %u)", p->k);
    if (isnan(p->eps)) return sfb_fail(ctx, SFB_EINVAL, "eps is NaN");
    uint64_t q_begin = p->q_begin, q_end = p->q_end ? p->q_end : nodes;
    if (q_begin >= q_end || q_end > nodes) return sfb_fail(ctx, SFB_EINVAL, "bad query shard [%llu, %llu)", (unsigned long long)q_begin, (unsigned long long)q_end);
    const uint64_t nq = q_end - q_begin;

    SFB_TRY(sfb_knn_alloc(ctx, nq, p->k, out));
    sfb_knn* g = *out;
    g->q_begin = q_begin;
    g->total = nodes;
    g->stats = sfb_knn_stats{};
    g->stats.rows = nq;

    HostTrace tr(ctx, "knn_build");
    StageTimer total(ctx, &ctx->times.ms_knn);
    int32_t st = SFB_OK;
    int screen = p->screen;
    const bool dense = sfb_dense_shape(nodes, dims) && (screen == SFB_SCREEN_AUTO || screen == SFB_SCREEN_EXACT_F64);
    if (dense) {
        // feature-graph shape: exact f64 Gram tiles straight from the dims-major matrix
        StageTimer tf(ctx, nullptr);
        DevBuf tmp;
        const double* xd = x->d;
        if (!columns_are_nodes) {
            cudaError_t e = (sfb_tls_ctx = ctx, tmp.alloc(sizeof(double) * nodes * dims));
            if (e != cudaSuccess) st = sfb_fail(ctx, SFB_ENOMEM, "dims-major copy: %s", cudaGetErrorString(e));
            else st = sfb_transpose_device(ctx, x->d, nodes, dims, tmp.as<double>());
            xd = tmp.as<double>();
        }
        if (st == SFB_OK) st = sfb_knn_dense(ctx, xd, (uint32_t)nodes, dims, p->metric, p->k, p->eps, q_begin, nq, g->idx, g->dist, g->cnt, collective);
        if (st == SFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "dense kNN failed");
        g->stats.ms_fallback = tf.stop();
        g->stats.rows_fallback = nq;
        g->stats.screen_used = SFB_SCREEN_EXACT_F64;
    } else {
        // many nodes: rows must be the nodes
        sfb_mat view = *x;
        DevBuf tmp;
        if (columns_are_nodes) {
            cudaError_t e = (sfb_tls_ctx = ctx, tmp.alloc(sizeof(double) * nodes * dims));
            if (e != cudaSuccess) st = sfb_fail(ctx, SFB_ENOMEM, "node-major copy: %s", cudaGetErrorString(e));
            else st = sfb_transpose_device(ctx, x->d, dims, nodes, tmp.as<double>());
            view.d = tmp.as<double>(); view.rows = nodes; view.cols = (uint32_t)dims; view.owns = false;
        }
        DevBuf norms;
        const bool shared_work = ctx->knn_collective && ctx->world > 1;
        const uint64_t nw = shared_work ? (uint64_t)ctx->world : 1, S = (nodes + nw - 1) / nw;
        if (st == SFB_OK) {
            cudaError_t e = (sfb_tls_ctx = ctx, norms.alloc(sizeof(double) * nw * S));
            if (e != cudaSuccess) st = sfb_fail(ctx, SFB_ENOMEM, "norms: %s", cudaGetErrorString(e));
        }
        if (st == SFB_OK && p->metric == SFB_METRIC_COSINE) {
            if (!shared_work) st = sfb_row_norms(ctx, &view, norms.as<double>());
            else {   // the norms of this rank's rows, all-gathered
                const uint64_t r0 = (uint64_t)ctx->rank * S < nodes ? (uint64_t)ctx->rank * S : nodes, r1 = r0 + S < nodes ? r0 + S : nodes;
                if (r1 > r0) { sfb_mat part = view; part.d = view.d + r0 * view.cols; part.rows = r1 - r0; st = sfb_row_norms(ctx, &part, norms.as<double>() + r0); }
                if (st == SFB_OK) st = sfb_comm_allgather_bytes(ctx, norms.p, (size_t)S * sizeof(double));
            }
        }
        tr.mark("norms");
        if (screen == SFB_SCREEN_AUTO) {
            // the screen pays off once the pair count dwarfs its fixed costs and k' stays small
            bool big = nodes >= 4096 && dims >= 32 && p->k <= 64;
            screen = big ? SFB_SCREEN_F16 : SFB_SCREEN_EXACT_F64;
        }
        if (st == SFB_OK) {
            if (screen == SFB_SCREEN_EXACT_F64) {
                StageTimer tf(ctx, nullptr);
                st = sfb_knn_exact(ctx, &view, norms.as<double>(), p->metric, p->k, p->eps, nullptr, nq, q_begin, g->idx, g->dist, g->cnt);
                g->stats.ms_fallback = tf.stop();
                g->stats.rows_fallback = nq;
                g->stats.screen_used = SFB_SCREEN_EXACT_F64;
            } else {
                sfb_knn_params pp = *p;
                pp.screen = screen;
                st = sfb_knn_screened(ctx, &view, norms.as<double>(), &pp, q_begin, q_end, g);
            }
        }
        if (st == SFB_OK) {
            cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
            if (e2 != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "kNN: %s", cudaGetErrorString(e2));
        }
    }
    total.stop();
    tr.mark("rest");
    if (st != SFB_OK) { sfb_knn_free(g); *out = nullptr; }
    return st;
}

extern "C" int32_t sfb_knn_build(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* p, sfb_knn** out) {
    return knn_build_any(ctx, x, false, p, out, 0);
}

// The COLLECTIVE form for a row-sharded build: every rank passes the same matrix and its own query shard.
extern "C" int32_t sfb_knn_build_sharded(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* p, sfb_knn** out) {
    if (!ctx) return SFB_EINVAL;
    if (ctx->world > 1 && !ctx->nccl_comm) return sfb_fail(ctx, SFB_ENCCL, "communicator not initialised");
    ctx->knn_collective = true;
    const int32_t st = knn_build_any(ctx, x, false, p, out, 0);
    ctx->knn_collective = false;
    return st;
}

extern "C" int32_t sfb_knn_build_columns(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* p, sfb_knn** out) {
    return knn_build_any(ctx, x, true, p, out, 0);
}

extern "C" int32_t sfb_knn_build_columns_sharded(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* p, sfb_knn** out) {
    return knn_build_any(ctx, x, true, p, out, 1);
}


// ---- feature graph hidden behind the item screen ----------------------------------------------------------------
// The Gram-tile kernel of the feature graph is latency-bound (one dependent f64 add chain per pair: ~27 ms at
// N = 1M whatever the number of tiles, so it does not shrink with more GPUs) and needs almost nothing of an SM: 64
// threads and 8 KB of shared memory.  _begin registers it; the screen launcher fires it on a side stream right after
// the persistent screen kernel is enqueued, so its CTAs run beside the screen's; _end joins, all-reduces (sharded),
// and selects.  If no screen runs before _end, the job simply runs there.
int32_t sfb_gram_launch(sfb_ctx* ctx, cudaStream_t stream, const double* xd, uint32_t m, uint64_t kd, int metric, double* g,
                        uint32_t gt, uint32_t t0, uint32_t t1, bool small_smem);
uint32_t sfb_gram_tile_edge(const sfb_ctx* ctx, uint32_t m, int collective);
void sfb_gram_tile_range(const sfb_ctx* ctx, uint32_t m, uint32_t gt, int collective, uint32_t* t0, uint32_t* t1);
int32_t sfb_gram_finish(sfb_ctx* ctx, double* g, uint32_t m, int metric, uint32_t k, double eps, uint64_t q_begin, uint64_t nq,
                        uint32_t* out_idx, double* out_dist, uint32_t* out_cnt, int collective);

struct sfb_pending {
    sfb_ctx* ctx = nullptr;
    const sfb_mat* x = nullptr;
    sfb_knn_params p{};
    int collective = 0;
    bool dense = false, launched = false;
    bool deferred = false;   // no side stream (or a tiny problem): _end runs the tiles inline
    bool after_screen = false;   // fired when the screen kernel has FINISHED: runs beside the rescore, the exchanges and the item Laplacian
    uint32_t gt = 16;        // pair-tile edge (knn_exact.cu: sfb_gram_tile_edge)
    double* g = nullptr;   // 2 * m * m doubles
};

void sfb_side_job_fire(sfb_ctx* ctx) {
    sfb_pending* pd = ctx->side_job;
    if (!pd) return;
    ctx->side_job = nullptr;
    if (!ctx->side) return;   // _begin could not create the side stream: _end runs the job inline
    const uint32_t m = pd->x->cols;
    uint32_t t0, t1;
    sfb_gram_tile_range(ctx, m, pd->gt, pd->collective, &t0, &t1);
    // co-resident: side_fork was recorded by _begin (after the matrix upload and the memset of g) -- NOT here: an event
    // recorded now would sit behind the screen kernel that was just enqueued.  after_screen: that is exactly what is wanted.
    if (pd->after_screen) cudaEventRecord(ctx->side_fork, ctx->stream);
    cudaStreamWaitEvent(ctx->side, ctx->side_fork, 0);
    if (sfb_gram_launch(ctx, ctx->side, pd->x->d, m, pd->x->rows, pd->p.metric, pd->g, pd->gt, t0, t1, !pd->after_screen) != SFB_OK) return;
    cudaEventRecord(ctx->side_done, ctx->side);
    pd->launched = true;
}

extern "C" int32_t sfb_knn_build_columns_begin(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* p, int32_t sharded, sfb_pending** out) {
    if (!ctx || !x || !p || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    if (ctx->side_job) return sfb_fail(ctx, SFB_EINVAL, "a feature-graph build is already pending on this context");
    sfb_pending* pd = new (std::nothrow) sfb_pending();
    if (!pd) return SFB_ENOMEM;
    pd->ctx = ctx; pd->x = x; pd->p = *p; pd->collective = sharded ? 1 : 0;
    const uint64_t nodes = x->cols, dims = x->rows;
    pd->dense = nodes >= 2 && sfb_dense_shape(nodes, dims) && (p->screen == SFB_SCREEN_AUTO || p->screen == SFB_SCREEN_EXACT_F64) &&
                p->k >= 1 && p->k <= 128 && p->metric >= SFB_METRIC_COSINE && p->metric <= SFB_METRIC_L2SQ;
    if (pd->dense) {
        const size_t bytes = sizeof(double) * 2 * (size_t)nodes * nodes;
        if (sfb_dev_alloc(ctx, (void**)&pd->g, bytes) != cudaSuccess) { delete pd; return sfb_fail(ctx, SFB_ENOMEM, "Gram buffer"); }
        if (pd->collective && ctx->world > 1) cudaMemsetAsync(pd->g, 0, bytes / 2, ctx->stream);
        if (!ctx->side && cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) == cudaSuccess) {
            cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ctx->side_done, cudaEventDisableTiming);
        }
        // Two ways to keep the chains off the critical path, both on the side stream, both fired by the next screen launch:
        //   co-resident   warps BESIDE the screen's CTAs, operands in a register ring (gram_warp_kernel, 8 x 8 pair tiles, at most
        //                 two warps per SM): ~50 ns per fold step there, and the screen does not notice (tools/gram_probe.py: one
        //                 rank's share of C2 at 8 GPUs rides a 98 ms screen for -0.5 ms) -- the choice while tiles per warp x N
        //                 steps finish before the screen does;
        //   after-screen  the stand-alone shared-memory kernel, released by an event behind the screen kernel: it runs beside the
        //                 rescore, the fallback, the list exchange and the item Laplacian, which leave the FP64 pipe idle.
        uint32_t t0, t1;
        const char* gt_forced = getenv("SFB_GRAM_GT");
        const uint32_t gt_co = gt_forced ? sfb_gram_tile_edge(ctx, (uint32_t)nodes, pd->collective) : 8u;
        sfb_gram_tile_range(ctx, (uint32_t)nodes, gt_co, pd->collective, &t0, &t1);
        const double tiles_per_warp = ceil((double)(t1 - t0) / (2.0 * (double)ctx->sm_count));
        const double co_s = tiles_per_warp * (double)dims * (gt_co == 8 ? 60e-9 : 140e-9);
        const double screen_s = 2.0 * (double)dims * (double)dims * (double)nodes / (double)(ctx->world > 0 ? ctx->world : 1) / 1.1e15;
        bool co = co_s < 0.9 * screen_s;
        if (const char* e = getenv("SFB_GRAM_MODE")) { if (e[0] == 'c') co = true; else if (e[0] == 'a') co = false; }
        pd->gt = co ? gt_co : sfb_gram_tile_edge(ctx, (uint32_t)nodes, pd->collective);
        // chains of less than half a millisecond are not worth a side slot: _end runs them inline (and such builds may be stacked)
        const bool side = ctx->side && dims >= 4096 && (double)dims * 30e-9 > 0.5e-3 && !getenv("SFB_NO_SIDE_STREAM");
        if (side) {
            pd->after_screen = !co;
            if (co) cudaEventRecord(ctx->side_fork, ctx->stream);
            ctx->side_job = pd;   // fired by the next screen launch, or by _end
        }
        pd->deferred = !side;
    }
    *out = pd;
    return SFB_OK;
}

extern "C" int32_t sfb_knn_build_columns_end(sfb_ctx* ctx, sfb_pending* pd, sfb_knn** out) {
    if (!ctx || !pd || !out || pd->ctx != ctx) return sfb_fail(ctx, SFB_EINVAL, "bad pending handle");
    *out = nullptr;
    int32_t st = SFB_OK;
    if (!pd->dense) {
        st = knn_build_any(ctx, pd->x, true, &pd->p, out, pd->collective);   // argument errors and non-feature shapes: the plain path
    } else {
        const uint32_t m = pd->x->cols;
        const uint32_t kk = pd->p.k;
        uint64_t q_begin = pd->p.q_begin, q_end = pd->p.q_end ? pd->p.q_end : m;
        if (q_begin >= q_end || q_end > m || isnan(pd->p.eps)) st = sfb_fail(ctx, SFB_EINVAL, "bad query shard or eps");
        if (ctx->side_job == pd || pd->deferred) {   // no screen ran in between, or not worth hiding: run the Gram tiles now, on the main stream
            ctx->side_job = nullptr;
            uint32_t t0, t1;
            pd->gt = sfb_gram_tile_edge(ctx, m, pd->collective);   // the stand-alone tiling
            sfb_gram_tile_range(ctx, m, pd->gt, pd->collective, &t0, &t1);
            if (st == SFB_OK) st = sfb_gram_launch(ctx, ctx->stream, pd->x->d, m, pd->x->rows, pd->p.metric, pd->g, pd->gt, t0, t1, false);
        } else if (pd->launched) {
            cudaStreamWaitEvent(ctx->stream, ctx->side_done, 0);
        } else if (st == SFB_OK) st = sfb_fail(ctx, SFB_ECUDA, "the side launch of the Gram tiles failed");
        if (st == SFB_OK) st = sfb_knn_alloc(ctx, q_end - q_begin, kk, out);
        if (st == SFB_OK) {
            sfb_knn* g = *out;
            g->q_begin = q_begin; g->total = m; g->stats = sfb_knn_stats{};
            g->stats.rows = q_end - q_begin; g->stats.rows_fallback = g->stats.rows; g->stats.screen_used = SFB_SCREEN_EXACT_F64;
            StageTimer t(ctx, &ctx->times.ms_knn);
            st = sfb_gram_finish(ctx, pd->g, m, pd->p.metric, kk, pd->p.eps, q_begin, q_end - q_begin, g->idx, g->dist, g->cnt, pd->collective);
            if (st == SFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "feature-graph selection failed");
            g->stats.ms_fallback = t.stop();
            if (st != SFB_OK) { sfb_knn_free(g); *out = nullptr; }
        }
        if (ctx->side) cudaStreamSynchronize(ctx->side);
        sfb_dev_free(ctx, pd->g);
    }
    delete pd;
    return st;
}
