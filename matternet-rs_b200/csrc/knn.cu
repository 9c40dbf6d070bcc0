// knn.cu -- sfb_knn_build: dispatch between the exact f64 path and the tensor-core screen.
#include <math.h>

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
        if (st == SFB_OK) {
            cudaError_t e = (sfb_tls_ctx = ctx, norms.alloc(sizeof(double) * nodes));
            if (e != cudaSuccess) st = sfb_fail(ctx, SFB_ENOMEM, "norms: %s", cudaGetErrorString(e));
        }
        if (st == SFB_OK && p->metric == SFB_METRIC_COSINE) st = sfb_row_norms(ctx, &view, norms.as<double>());
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

extern "C" int32_t sfb_knn_build_columns(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* p, sfb_knn** out) {
    return knn_build_any(ctx, x, true, p, out, 0);
}

extern "C" int32_t sfb_knn_build_columns_sharded(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* p, sfb_knn** out) {
    return knn_build_any(ctx, x, true, p, out, 1);
}
