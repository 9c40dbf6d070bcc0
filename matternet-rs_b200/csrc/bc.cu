// bc.cu -- successor Stage C: Bhattacharyya-coefficient kNN over the feature nodes (f32 semantics).
//
// Reference: LaplacianStage::execute / compute_bhattacharyya_weights / build_laplacian_flat
// (surfface-core/src/laplacian.rs:135-219,254-298,312-394) and bhattacharyya_coefficient
// (surfface-core/src/distance.rs:260-290).  The reference pulls the [C, F] centroid state to the CPU and
// scans all F^2 pairs per feature on rayon threads; here one CTA owns a feature node, its threads own the
// candidate nodes j (coalesced over the row-major state: a[c*F + j]) and run the per-pair left fold over the C
// centroids in f32 with separately rounded operations.  BC is not GEMM-form (a log and a ratio per centroid and
// pair), so this stays on the FP32 / SFU pipes: F^2 * C evaluations.  The top-k (BC desc, j asc) reuses the
// dense-key selection of knn_exact.cu; max-symmetrisation and the normalised Laplacian reuse laplacian.cu.
// log / exp: CUDA's logf / expf and the host libm's differ in the last bit, which would let a transcendental decide a
// neighbour ORDER differently on the two sides.  Both are therefore evaluated here in f64 with + - * / only, in a fixed
// order (the atanh series of synth.cuh; Taylor after a k*ln2 reduction), and rounded to f32 -- the same sequence as the
// CPU restatement the tests check against, so indices and weights agree with it bit for bit; against glibc (the reference on
// Linux) the values agree except for < 0.1 % of arguments that land 1 ulp apart (tests/test_oracle_kat.py).
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "synth.cuh"

int32_t sfb_adj_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t k, sfb_adj** out);
int32_t sfb_dense_select(sfb_ctx* ctx, const double* keys, uint32_t m, uint64_t q_begin, uint64_t nq, uint32_t k, double eps,
                         uint32_t* out_idx, double* out_dist, uint32_t* out_cnt);

namespace {

__device__ __forceinline__ float det_logf(float a) {
    if (a != a || a < 0.0f) return __int_as_float(0x7FC00000);
    if (a == 0.0f) return -INFINITY;
    if (isinf(a)) return INFINITY;
    return __double2float_rn(synth_log((double)a));   // every positive f32, subnormals included, is a normal f64
}
__device__ __forceinline__ float det_expf(float xf) {
    if (xf != xf) return xf;
    const double x = (double)xf;
    if (x > 100.0) return INFINITY;
    if (x < -120.0) return 0.0f;
    const double kf = rint(__dmul_rn(x, 1.4426950408889634));
    const double r = __dadd_rn(__dadd_rn(x, -__dmul_rn(kf, 6.93147180369123816490e-01)), -__dmul_rn(kf, 1.90821492927058770002e-10));
    double p = 1.0 / 87178291200.0;
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 6227020800.0);  p = __dadd_rn(__dmul_rn(p, r), 1.0 / 479001600.0); p = __dadd_rn(__dmul_rn(p, r), 1.0 / 39916800.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 3628800.0);     p = __dadd_rn(__dmul_rn(p, r), 1.0 / 362880.0);    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 40320.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 5040.0);        p = __dadd_rn(__dmul_rn(p, r), 1.0 / 720.0);       p = __dadd_rn(__dmul_rn(p, r), 1.0 / 120.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 24.0);          p = __dadd_rn(__dmul_rn(p, r), 1.0 / 6.0);         p = __dadd_rn(__dmul_rn(p, r), 0.5);
    p = __dadd_rn(__dmul_rn(p, r), 1.0);                 p = __dadd_rn(__dmul_rn(p, r), 1.0);
    const long long k = (long long)kf;
    const double sc = __longlong_as_double((k + 1023) << 52);
    return __double2float_rn(__dmul_rn(p, sc));
}

// keys[i][j] = -BC(i, j) (ascending key = descending affinity), +inf where j == i or BC <= thr
__global__ void __launch_bounds__(128) bc_keys_kernel(const float* __restrict__ means, const float* __restrict__ vars, uint32_t c,
                                                      uint32_t f, float reg, float thr, double* __restrict__ keys) {
    extern __shared__ float sm[];
    float* mu_i = sm;        // [c]
    float* v_i = sm + c;     // [c], floored
    const uint32_t i = blockIdx.x;
    for (uint32_t cc = threadIdx.x; cc < c; cc += blockDim.x) {
        mu_i[cc] = means[(size_t)cc * f + i];
        v_i[cc] = fmaxf(vars[(size_t)cc * f + i], reg);   // f32::max ignores NaN, as fmaxf does
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < f; j += blockDim.x) {
        double key = INFINITY;
        if (j != i) {
            float db = 0.0f;
            for (uint32_t cc = 0; cc < c; ++cc) {
                const float vi = v_i[cc], vj = fmaxf(__ldg(vars + (size_t)cc * f + j), reg);
                const float v_sum = __fadd_rn(vi, vj);
                const float dm = __fadd_rn(mu_i[cc], -__ldg(means + (size_t)cc * f + j));
                const float mean_term = __fdiv_rn(__fmul_rn(dm, dm), __fmul_rn(4.0f, v_sum));
                const float log_term = __fmul_rn(0.5f, det_logf(__fdiv_rn(v_sum, __fmul_rn(2.0f, __fsqrt_rn(__fmul_rn(vi, vj))))));
                db = __fadd_rn(db, __fadd_rn(mean_term, log_term));
            }
            float bc = det_expf(-db);
            bc = bc < 0.0f ? 0.0f : (bc > 1.0f ? 1.0f : bc);
            if (bc > thr) key = -(double)bc;
        }
        keys[(size_t)i * f + j] = key;
    }
}

__global__ void bc_to_adj_kernel(const double* __restrict__ dist, const uint32_t* __restrict__ idx, uint64_t n, double* __restrict__ w) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) w[g] = idx[g] == SFB_IDX_NONE ? 0.0 : -dist[g];
}

}  // namespace

extern "C" int32_t sfb_bc_adjacency_build(sfb_ctx* ctx, const float* means, const float* variances, uint32_t n_centroids,
                                          uint32_t n_features, uint32_t k, float variance_regularizer, float weight_threshold,
                                          sfb_adj** out) {
    if (!ctx || !means || !variances || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    const uint32_t c = n_centroids, f = n_features;
    if (c == 0 || f < 2) return sfb_fail(ctx, SFB_EINVAL, "need at least 2 features and 1 centroid (got %u, %u)", f, c);
    if (f > 16384) return sfb_fail(ctx, SFB_EUNSUPPORTED, "feature graphs above 16384 nodes need the tiled path");
    if ((size_t)c * 8 > ctx->smem_optin) return sfb_fail(ctx, SFB_EUNSUPPORTED, "too many centroids for one shared-memory profile");
    uint32_t kk = k < f - 1 ? k : f - 1;   // k.min(f - 1), laplacian.rs:260
    if (kk == 0 || kk > 128) return sfb_fail(ctx, SFB_EUNSUPPORTED, "k_neighbors must be in 1..128 (got %u)", k);
    StageTimer t(ctx, &ctx->times.ms_knn);
    DevBuf dm, dv, keys, dist;
    const size_t sz = (size_t)c * f * sizeof(float);
    SFB_CUDA(ctx, dm.alloc(sz));
    SFB_CUDA(ctx, dv.alloc(sz));
    SFB_CUDA(ctx, keys.alloc(sizeof(double) * (size_t)f * f));
    SFB_CUDA(ctx, dist.alloc(sizeof(double) * (size_t)f * kk));
    SFB_CUDA(ctx, cudaMemcpyAsync(dm.p, means, sz, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(dv.p, variances, sz, cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = (size_t)c * 2 * sizeof(float);
    SFB_CUDA(ctx, cudaFuncSetAttribute(bc_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bc_keys_kernel<<<f, 128, smem, ctx->stream>>>(dm.as<float>(), dv.as<float>(), c, f, variance_regularizer, weight_threshold, keys.as<double>());
    SFB_LAUNCH_CHECK(ctx);
    SFB_TRY(sfb_adj_alloc(ctx, f, kk, out));
    sfb_adj* a = *out;
    int32_t st = sfb_dense_select(ctx, keys.as<double>(), f, 0, f, kk, DBL_MAX, a->idx, dist.as<double>(), a->cnt);
    if (st == SFB_OK) {
        bc_to_adj_kernel<<<div_up((uint64_t)f * kk, 256), 256, 0, ctx->stream>>>(dist.as<double>(), a->idx, (uint64_t)f * kk, a->w);
        ctx->times.kernel_launches++;
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "Bhattacharyya adjacency failed");
    }
    if (st != SFB_OK) { sfb_adj_free(a); *out = nullptr; }
    return st;
}

// LaplacianStage::execute in one call (surfface-core/src/laplacian.rs:135-219): host state in, CSR handle out.
// degrees (f floats, may be NULL) receives the degree vector of the symmetrised graph (laplacian.rs:333-340).
extern "C" int32_t sfb_laplacian_stage_execute(sfb_ctx* ctx, const float* means, const float* variances, uint32_t n_centroids,
                                               uint32_t n_features, const sfb_laplacian_config* cfg, sfb_csr** out, float* degrees) {
    if (!cfg || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    sfb_adj* a = nullptr;
    SFB_TRY(sfb_bc_adjacency_build(ctx, means, variances, n_centroids, n_features, cfg->k_neighbors, cfg->variance_regularizer,
                                   cfg->weight_threshold, &a));
    int32_t st = SFB_OK;
    if (degrees) {
        // the unnormalised Laplacian of the same graph carries the degrees on its diagonal
        sfb_csr* lu = nullptr;
        sfb_lap_params up{0, (double)cfg->weight_threshold};
        st = sfb_laplacian_build(ctx, a, &up, &lu);
        if (st == SFB_OK) {
            const uint64_t f = n_features;
            std::string hb;
            hb.resize(sizeof(uint64_t) * (f + 1) + (sizeof(uint32_t) + sizeof(double)) * (lu->nnz ? lu->nnz : 1));
            uint64_t* ip = reinterpret_cast<uint64_t*>(&hb[0]);
            double* dv = reinterpret_cast<double*>(ip + f + 1);
            uint32_t* ix = reinterpret_cast<uint32_t*>(dv + (lu->nnz ? lu->nnz : 1));
            st = sfb_csr_copy(ctx, lu, ip, ix, dv);
            if (st == SFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "degree fetch failed");
            if (st == SFB_OK)
                for (uint64_t r = 0; r < f; ++r) {
                    degrees[r] = 0.0f;
                    for (uint64_t e = ip[r]; e < ip[r + 1]; ++e) if (ix[e] == r) degrees[r] = (float)dv[e];
                }
        }
        sfb_csr_free(lu);
    }
    if (st == SFB_OK) {
        sfb_lap_params lp{cfg->normalize ? 1 : 0, (double)cfg->weight_threshold};
        st = sfb_laplacian_build(ctx, a, &lp, out);
    }
    sfb_adj_free(a);
    return st;
}
