// knn.cu -- sfb_knn_build: dispatch between the exact f64 path and the tensor-core screen.
#include <math.h>

#include "common.cuh"

int32_t sfb_knn_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t k, sfb_knn** out);

extern "C" int32_t sfb_knn_build(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* p, sfb_knn** out) {
    if (!ctx || !x || !p || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    // the reference asserts n >= 2 && d >= 2 (src_legacy/laplacian.rs:130-135)
    if (x->rows < 2) return sfb_fail(ctx, SFB_EINVAL, "need at least 2 rows (got %llu)", (unsigned long long)x->rows);
    if (p->metric < SFB_METRIC_COSINE || p->metric > SFB_METRIC_L2SQ) return sfb_fail(ctx, SFB_EINVAL, "unknown metric %d", p->metric);
    if (p->k == 0 || p->k > 128) return sfb_fail(ctx, SFB_EUNSUPPORTED, "k must be in 1..128 (got %u)", p->k);
    if (isnan(p->eps)) return sfb_fail(ctx, SFB_EINVAL, "eps is NaN");
    uint64_t q_begin = p->q_begin, q_end = p->q_end ? p->q_end : x->rows;
    if (q_begin >= q_end || q_end > x->rows) return sfb_fail(ctx, SFB_EINVAL, "bad query shard [%llu, %llu)", (unsigned long long)q_begin, (unsigned long long)q_end);
    const uint64_t nq = q_end - q_begin;

    SFB_TRY(sfb_knn_alloc(ctx, nq, p->k, out));
    sfb_knn* g = *out;
    g->q_begin = q_begin;
    g->total = x->rows;
    g->stats = sfb_knn_stats{};
    g->stats.rows = nq;

    StageTimer total(ctx, &ctx->times.ms_knn);
    DevBuf norms;
    sfb_tls_ctx = ctx;
    cudaError_t e = norms.alloc(sizeof(double) * x->rows);
    if (e != cudaSuccess) { sfb_knn_free(g); *out = nullptr; return sfb_fail(ctx, SFB_ENOMEM, "norms: %s", cudaGetErrorString(e)); }
    int32_t st = SFB_OK;
    if (p->metric == SFB_METRIC_COSINE) st = sfb_row_norms(ctx, x, norms.as<double>());

    int screen = p->screen;
    if (screen == SFB_SCREEN_AUTO) {
        // the screen pays off once the pair count dwarfs its fixed costs and k' stays small
        bool big = x->rows >= 4096 && x->cols >= 32 && p->k <= 64;
        screen = big ? SFB_SCREEN_F16 : SFB_SCREEN_EXACT_F64;
    }
    if (st == SFB_OK) {
        if (screen == SFB_SCREEN_EXACT_F64) {
            StageTimer tf(ctx, nullptr);
            st = sfb_knn_exact(ctx, x, norms.as<double>(), p->metric, p->k, p->eps, nullptr, nq, q_begin, g->idx, g->dist, g->cnt);
            g->stats.ms_fallback = tf.stop();
            g->stats.rows_fallback = nq;
            g->stats.screen_used = SFB_SCREEN_EXACT_F64;
        } else {
            sfb_knn_params pp = *p;
            pp.screen = screen;
            st = sfb_knn_screened(ctx, x, norms.as<double>(), &pp, q_begin, q_end, g);
        }
    }
    if (st == SFB_OK) {
        cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
        if (e2 != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "kNN: %s", cudaGetErrorString(e2));
    }
    total.stop();
    if (st != SFB_OK) { sfb_knn_free(g); *out = nullptr; }
    return st;
}
