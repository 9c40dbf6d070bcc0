// pipeline.cu -- reference-shaped one-shot entry points: host buffers in, host buffers out.
//   sfb_build_laplacian_matrix  = build_laplacian_matrix          (src_legacy/laplacian.rs:122-180)
//   sfb_compute_taumode_lambdas = compute_taumode_lambdas_parallel (src_legacy/taumode.rs:117-214)
//                                 + update_lambdas -> normalise_lambdas (core.rs:1427-1443,1341-1355)
#include <math.h>

#include "common.cuh"

extern "C" int32_t sfb_build_laplacian_matrix(sfb_ctx* ctx, const double* items, uint64_t nodes, uint32_t dims,
                                              const sfb_graph_params* gp, int32_t screen, sfb_csr** out) {
    if (!ctx || !items || !gp || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    // assert!(n >= 2 && d >= 2) (laplacian.rs:130-135)
    if (nodes < 2 || dims < 2) return sfb_fail(ctx, SFB_EINVAL, "items should be at least of shape (2,2): (%u,%llu)", dims, (unsigned long long)nodes);
    if (gp->normalise) return sfb_fail(ctx, SFB_EUNSUPPORTED, "normalise=true (StandardScaler pre-scaling, laplacian.rs:147-156) is out of scope");
    if (gp->topk == 0) return sfb_fail(ctx, SFB_EINVAL, "topk must be >= 1");
    sfb_mat* x = nullptr; sfb_knn* g = nullptr; sfb_adj* a = nullptr;
    int32_t st = sfb_mat_from_host(ctx, items, nodes, dims, &x);
    if (st == SFB_OK) {
        sfb_knn_params kp{};
        kp.metric = SFB_METRIC_COSINE;
        kp.k = gp->topk < nodes - 1 ? gp->topk : (uint32_t)(nodes - 1);
        kp.eps = gp->eps; kp.screen = screen; kp.allow_fallback = 1;
        st = sfb_knn_build(ctx, x, &kp, &g);
    }
    if (st == SFB_OK) {
        sfb_adj_params ap{gp->p, gp->sigma, -1};
        st = sfb_adjacency_build(ctx, g, &ap, &a, nullptr);
    }
    if (st == SFB_OK) {
        sfb_lap_params lp{0, 0.0};
        st = sfb_laplacian_build(ctx, a, &lp, out);
    }
    if (st == SFB_OK && gp->sparsity_check) {
        // graph.rs:232-240: panic when sparsity > 0.95
        double sparsity = 1.0 - (double)(*out)->nnz / ((double)nodes * (double)nodes);
        if (sparsity > 0.95) {
            sfb_csr_free(*out); *out = nullptr;
            st = sfb_fail(ctx, SFB_EINVAL, "Resulting laplacian matrix is too sparse %.6f", sparsity);
        }
    }
    sfb_adj_free(a); sfb_knn_free(g); sfb_mat_free(x);
    return st;
}

extern "C" int32_t sfb_compute_taumode_lambdas(sfb_ctx* ctx, const sfb_csr* L, const double* items, uint64_t n_items,
                                               uint32_t n_features, int32_t tau_mode, double tau_value, double* out_lambdas) {
    if (!ctx || !L || !items || !out_lambdas) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    sfb_mat* x = nullptr;
    SFB_TRY(sfb_mat_from_host(ctx, items, n_items, n_features, &x));
    sfb_lambda_params lp{SFB_LAMBDA_LEGACY_TAUMODE, tau_mode, tau_value, 1};
    int32_t st = sfb_lambda(ctx, L, x, &lp, out_lambdas, nullptr, nullptr);
    if (st == SFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "lambda D2H failed");
    sfb_mat_free(x);
    return st;
}
