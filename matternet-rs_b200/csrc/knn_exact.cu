// knn_exact.cu -- brute-force kNN in the reference's f64 arithmetic.
//
// Computes exactly what src_legacy/tests/test_helpers.rs:77-125 computes (cosine) and
// src_legacy/energymaps.rs:875-892 / surfface-core/src/mst.rs:312-403 (L2): every pair's sum is a
// left fold over the dimension index with separately rounded multiply and add (__dmul_rn /
// __dadd_rn are never contracted into an FMA), sqrt and division are IEEE, so each distance has the
// same bits as the CPU restatement and the (distance, index) order is the same total order.
//
// Used (a) as the SFB_SCREEN_EXACT_F64 path, (b) as the fallback for rows the tensor-core screen
// cannot certify.  FP64-pipe bound: 2*nq*M*K flops.
#include <math.h>

#include "common.cuh"
#include "topk_list.cuh"

namespace {

constexpr int TQ = 64, TC = 64, KC = 16, TP = 66, THREADS = 256;

// 32 rows per warp through a padded shared tile: coalesced loads, per-lane left fold.
__global__ void row_norms_kernel(const double* __restrict__ x, uint64_t m, uint32_t kd, double* __restrict__ norms) {
    __shared__ double tile[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t row0 = ((uint64_t)blockIdx.x * 4 + w) * 32;
    if (row0 >= m) return;
    double acc = 0.0;
    for (uint32_t d0 = 0; d0 < kd; d0 += 32) {
        for (int r = 0; r < 32; ++r) {
            uint64_t g = row0 + r;
            tile[w][r][lane] = (g < m && d0 + lane < kd) ? x[g * kd + d0 + lane] : 0.0;
        }
        __syncwarp();
        uint32_t lim = kd - d0 < 32 ? kd - d0 : 32;
        for (uint32_t d = 0; d < lim; ++d) { double v = tile[w][lane][d]; acc = __dadd_rn(acc, __dmul_rn(v, v)); }
        __syncwarp();
    }
    if (row0 + lane < m) norms[row0 + lane] = __dsqrt_rn(acc);
}

struct ExactArgs {
    const double* x; const double* norms; uint64_t m; uint32_t kd; int metric; uint32_t k; double eps;
    const uint32_t* query_rows; uint64_t nq; uint64_t q_begin; uint32_t csplits; uint32_t tiles_per_split;
    uint32_t* out_idx; double* out_dist; uint32_t* out_cnt;
};

template <bool COS>
__global__ void __launch_bounds__(THREADS) knn_exact_kernel(ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* qs = reinterpret_cast<double*>(smem_raw);          // [KC][TP]
    double* cs = qs + KC * TP;                                 // [KC][TP]
    double* keys = cs + KC * TP;                               // [TQ][TC+1]
    double* qn = keys + TQ * (TC + 1);                         // [TQ]
    double* cn = qn + TQ;                                      // [TC]
    double* ld = cn + TC;                                      // [TQ][k]
    uint32_t* li = reinterpret_cast<uint32_t*>(ld + (size_t)TQ * a.k);  // [TQ][k]
    uint32_t* gq = li + (size_t)TQ * a.k;                      // [TQ] global query row (or NONE)
    uint32_t* lcnt = gq + TQ;                                  // [TQ]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const uint64_t q0 = (uint64_t)blockIdx.x * TQ;

    if (tid < TQ) {
        uint64_t qi = q0 + tid;
        uint32_t g = SFB_IDX_NONE;
        if (qi < a.nq) g = a.query_rows ? a.query_rows[qi] : (uint32_t)(a.q_begin + qi);
        gq[tid] = g;
        lcnt[tid] = 0;
        qn[tid] = (COS && g != SFB_IDX_NONE) ? a.norms[g] : 0.0;
    }
    __syncthreads();

    const uint64_t n_tiles = (a.m + TC - 1) / TC;
    const uint64_t t_begin = (uint64_t)blockIdx.y * a.tiles_per_split;
    uint64_t t_end = t_begin + a.tiles_per_split;
    if (t_end > n_tiles) t_end = n_tiles;

    for (uint64_t tile = t_begin; tile < t_end; ++tile) {
        const uint64_t c0 = tile * TC;
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

        for (uint32_t d0 = 0; d0 < a.kd; d0 += KC) {
            __syncthreads();
            {
                const int r = tid >> 2, cg = (tid & 3) * 4;
                uint32_t g = gq[r];
                const double* qsrc = a.x + (uint64_t)g * a.kd + d0 + cg;
                uint64_t cr = c0 + r;
                const double* csrc = a.x + cr * a.kd + d0 + cg;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bool din = d0 + cg + j < a.kd;
                    qs[(cg + j) * TP + r] = (din && g != SFB_IDX_NONE) ? __ldg(qsrc + j) : 0.0;
                    cs[(cg + j) * TP + r] = (din && cr < a.m) ? __ldg(csrc + j) : 0.0;
                }
            }
            if (d0 == 0 && tid < TC) cn[tid] = (COS && c0 + tid < a.m) ? a.norms[c0 + tid] : 0.0;
            __syncthreads();
#pragma unroll
            for (int d = 0; d < KC; ++d) {
                double q[4], c[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { q[i] = qs[d * TP + ty * 4 + i]; c[i] = cs[d * TP + tx * 4 + i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (COS) {
                            acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(q[i], c[j]));
                        } else {
                            double t = __dadd_rn(q[i], -c[j]);
                            acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(t, t));
                        }
                    }
            }
        }
        // keys for this tile
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double key;
                if (COS) {
                    double denom = __dmul_rn(qn[ty * 4 + i], cn[tx * 4 + j]);
                    double cosv = 0.0;
                    if (denom > 1e-12) {
                        cosv = __ddiv_rn(acc[i][j], denom);
                        if (cosv < -1.0) cosv = -1.0; else if (cosv > 1.0) cosv = 1.0;
                    }
                    double rect = cosv > 0.0 ? cosv : 0.0;
                    key = __dadd_rn(1.0, -rect);
                } else {
                    key = a.metric == SFB_METRIC_L2 ? __dsqrt_rn(acc[i][j]) : acc[i][j];
                }
                keys[(ty * 4 + i) * (TC + 1) + tx * 4 + j] = key;
            }
        __syncthreads();
        // selection: warp w owns query rows w*8 .. w*8+7
        for (int rr = 0; rr < 8; ++rr) {
            const int r = warp * 8 + rr;
            const uint32_t g = gq[r];
            if (g == SFB_IDX_NONE) continue;
            double* rld = ld + (size_t)r * a.k;
            uint32_t* rli = li + (size_t)r * a.k;
            uint32_t c = lcnt[r];
            double key0 = keys[r * (TC + 1) + lane], key1 = keys[r * (TC + 1) + lane + 32];
            uint64_t j0 = c0 + lane, j1 = c0 + lane + 32;
            double thr_d = c == a.k ? rld[a.k - 1] : INFINITY;
            uint32_t thr_i = c == a.k ? rli[a.k - 1] : SFB_IDX_NONE;
            bool p0 = j0 < a.m && (uint32_t)j0 != g && key0 <= a.eps && topk_key_less(key0, (uint32_t)j0, thr_d, thr_i);
            bool p1 = j1 < a.m && (uint32_t)j1 != g && key1 <= a.eps && topk_key_less(key1, (uint32_t)j1, thr_d, thr_i);
            uint32_t b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
            while (b0 | b1) {
                int src; double kd_; uint32_t jj;
                if (b0) { src = __ffs(b0) - 1; b0 &= b0 - 1; kd_ = __shfl_sync(0xffffffffu, key0, src); jj = (uint32_t)(c0 + src); }
                else    { src = __ffs(b1) - 1; b1 &= b1 - 1; kd_ = __shfl_sync(0xffffffffu, key1, src); jj = (uint32_t)(c0 + src + 32); }
                warp_list_insert(rld, rli, c, a.k, kd_, jj, lane);
            }
            if (lane == 0) lcnt[r] = c;
        }
        // keys / lists are re-read only after the next tile's __syncthreads
    }
    __syncthreads();
    // write the (partial) lists
    for (int rr = 0; rr < 8; ++rr) {
        const int r = warp * 8 + rr;
        uint64_t qi = q0 + r;
        if (qi >= a.nq) continue;
        uint32_t c = lcnt[r];
        size_t o = ((size_t)qi * a.csplits + blockIdx.y) * a.k;
        for (uint32_t t = lane; t < a.k; t += 32) {
            a.out_idx[o + t] = t < c ? li[(size_t)r * a.k + t] : SFB_IDX_NONE;
            a.out_dist[o + t] = t < c ? ld[(size_t)r * a.k + t] : INFINITY;
        }
        if (lane == 0) a.out_cnt[(size_t)qi * a.csplits + blockIdx.y] = c;
    }
}

// merge csplits partial lists per query: one warp per query row
__global__ void knn_merge_kernel(const uint32_t* __restrict__ pidx, const double* __restrict__ pdist,
                                 const uint32_t* __restrict__ pcnt, uint64_t nq, uint32_t csplits, uint32_t k,
                                 uint32_t* __restrict__ out_idx, double* __restrict__ out_dist, uint32_t* __restrict__ out_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* ld = reinterpret_cast<double*>(smem_raw) + (size_t)w * k;
    uint32_t* li = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)wpb * k) + (size_t)w * k;
    uint64_t qi = (uint64_t)blockIdx.x * wpb + w;
    if (qi >= nq) return;
    uint32_t c = 0;
    for (uint32_t s = 0; s < csplits; ++s) {
        size_t o = ((size_t)qi * csplits + s) * k;
        uint32_t pc = pcnt[(size_t)qi * csplits + s];
        for (uint32_t t = 0; t < pc; ++t) {
            double d = pdist[o + t]; uint32_t j = pidx[o + t];
            if (c == k && !topk_key_less(d, j, ld[k - 1], li[k - 1])) break;  // partial lists are sorted
            warp_list_insert(ld, li, c, k, d, j, lane);
        }
    }
    for (uint32_t t = lane; t < k; t += 32) {
        out_idx[qi * k + t] = t < c ? li[t] : SFB_IDX_NONE;
        out_dist[qi * k + t] = t < c ? ld[t] : INFINITY;
    }
    if (lane == 0) out_cnt[qi] = c;
}

// ---- few nodes, very long rows (the feature graph: D nodes x N dims, graph.rs:214-216) ----------
// The 64x64 tile kernel above would launch a handful of CTAs, each walking millions of dimensions.
// Here every (i, j) pair gets its own thread: 16 x 16 pair tiles, the dimension streamed through
// shared memory 32 at a time (coalesced 256-byte row segments), one left fold per pair.  For a
// self-kNN only tiles on or above the diagonal are computed and mirrored (products commute, so
// key(i,j) and key(j,i) have the same bits).  FP64-pipe bound: nodes^2 * dims multiply-adds.
template <bool COS>
__global__ void __launch_bounds__(256) knn_dense_keys_kernel(const double* __restrict__ x, const double* __restrict__ norms,
                                                             uint32_t m, uint32_t kd, int metric, double* __restrict__ keys) {
    __shared__ double qs[16][33], cs[16][33];
    const uint32_t ti = blockIdx.y, tj = blockIdx.x;
    if (tj < ti) return;  // mirrored
    const int tid = threadIdx.x, qi = tid >> 4, cj = tid & 15;
    const uint32_t gi = ti * 16 + qi, gj = tj * 16 + cj;
    double acc = 0.0;
    for (uint32_t d0 = 0; d0 < kd; d0 += 32) {
        __syncthreads();
        for (int e = tid; e < 16 * 32; e += 256) {
            int r = e >> 5, d = e & 31;
            uint32_t ri = ti * 16 + r, rj = tj * 16 + r;
            qs[r][d] = (ri < m && d0 + d < kd) ? __ldg(x + (uint64_t)ri * kd + d0 + d) : 0.0;
            cs[r][d] = (rj < m && d0 + d < kd) ? __ldg(x + (uint64_t)rj * kd + d0 + d) : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int d = 0; d < 32; ++d) {
            if (COS) acc = __dadd_rn(acc, __dmul_rn(qs[qi][d], cs[cj][d]));
            else { double t = __dadd_rn(qs[qi][d], -cs[cj][d]); acc = __dadd_rn(acc, __dmul_rn(t, t)); }
        }
    }
    if (gi >= m || gj >= m) return;
    double key;
    if (COS) {
        double denom = __dmul_rn(norms[gi], norms[gj]), cosv = 0.0;
        if (denom > 1e-12) { cosv = __ddiv_rn(acc, denom); if (cosv < -1.0) cosv = -1.0; else if (cosv > 1.0) cosv = 1.0; }
        double rect = cosv > 0.0 ? cosv : 0.0;
        key = __dadd_rn(1.0, -rect);
    } else key = metric == SFB_METRIC_L2 ? __dsqrt_rn(acc) : acc;
    keys[(uint64_t)gi * m + gj] = key;
    keys[(uint64_t)gj * m + gi] = key;
}

// one warp per query row: scan the dense key row, keep the (distance, index) top-k
__global__ void knn_dense_select_kernel(const double* __restrict__ keys, uint32_t m, uint64_t q_begin, uint64_t nq, uint32_t k,
                                        double eps, uint32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                        uint32_t* __restrict__ out_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* ld = reinterpret_cast<double*>(smem_raw) + (size_t)w * k;
    uint32_t* li = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)wpb * k) + (size_t)w * k;
    const uint64_t qi = (uint64_t)blockIdx.x * wpb + w;
    if (qi >= nq) return;
    const uint32_t g = (uint32_t)(q_begin + qi);
    uint32_t c = 0;
    for (uint32_t j0 = 0; j0 < m; j0 += 32) {
        uint32_t j = j0 + lane;
        double key = j < m ? keys[(uint64_t)g * m + j] : INFINITY;
        double td = c == k ? ld[k - 1] : INFINITY;
        uint32_t ti = c == k ? li[k - 1] : SFB_IDX_NONE;
        bool pass = j < m && j != g && key <= eps && topk_key_less(key, j, td, ti);
        uint32_t bal = __ballot_sync(0xffffffffu, pass);
        while (bal) {
            int src = __ffs(bal) - 1; bal &= bal - 1;
            double kd_ = __shfl_sync(0xffffffffu, key, src);
            warp_list_insert(ld, li, c, k, kd_, j0 + src, lane);
        }
    }
    for (uint32_t t = lane; t < k; t += 32) {
        out_idx[qi * k + t] = t < c ? li[t] : SFB_IDX_NONE;
        out_dist[qi * k + t] = t < c ? ld[t] : INFINITY;
    }
    if (lane == 0) out_cnt[qi] = c;
}

size_t exact_smem_bytes(uint32_t k) {
    return sizeof(double) * (2 * KC * TP + TQ * (TC + 1) + TQ + TC + (size_t)TQ * k) + sizeof(uint32_t) * ((size_t)TQ * k + 2 * TQ);
}

}  // namespace

int32_t sfb_row_norms(sfb_ctx* ctx, const sfb_mat* x, double* norms) {
    row_norms_kernel<<<div_up(x->rows, 128), 128, 0, ctx->stream>>>(x->d, x->rows, x->cols, norms);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

int32_t sfb_knn_exact(sfb_ctx* ctx, const sfb_mat* x, const double* norms, int metric, uint32_t k, double eps,
                      const uint32_t* query_rows, uint64_t nq, uint64_t q_begin, uint32_t* out_idx, double* out_dist,
                      uint32_t* out_cnt) {
    if (nq == 0) return SFB_OK;
    if (k == 0 || k > 128) return sfb_fail(ctx, SFB_EUNSUPPORTED, "k must be in 1..128 (got %u)", k);
    // feature-graph shape: few nodes with very long rows -> one thread per pair over a dense key matrix
    if (!query_rows && x->rows <= 4096 && (uint64_t)x->cols >= 8ull * x->rows) {
        const uint32_t m = (uint32_t)x->rows;
        DevBuf keys;
        SFB_CUDA(ctx, keys.alloc(sizeof(double) * (size_t)m * m));
        dim3 grid(div_up(m, 16), div_up(m, 16));
        if (metric == SFB_METRIC_COSINE) knn_dense_keys_kernel<true><<<grid, 256, 0, ctx->stream>>>(x->d, norms, m, x->cols, metric, keys.as<double>());
        else knn_dense_keys_kernel<false><<<grid, 256, 0, ctx->stream>>>(x->d, norms, m, x->cols, metric, keys.as<double>());
        SFB_LAUNCH_CHECK(ctx);
        const int wpb = 4;
        size_t ssm = (size_t)wpb * k * (sizeof(double) + sizeof(uint32_t));
        knn_dense_select_kernel<<<div_up(nq, wpb), wpb * 32, ssm, ctx->stream>>>(keys.as<double>(), m, q_begin, nq, k, eps, out_idx, out_dist, out_cnt);
        SFB_LAUNCH_CHECK(ctx);
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SFB_OK;
    }
    const uint64_t q_tiles = (nq + TQ - 1) / TQ, n_tiles = (x->rows + TC - 1) / TC;
    uint64_t want = 2ull * ctx->sm_count;
    uint32_t csplits = 1;
    if (q_tiles < want) {
        csplits = (uint32_t)((want + q_tiles - 1) / q_tiles);
        uint64_t max_splits = (n_tiles + 7) / 8;  // at least 8 corpus tiles per split
        if (csplits > max_splits) csplits = (uint32_t)max_splits;
        if (csplits > 65535) csplits = 65535;
        if (csplits < 1) csplits = 1;
    }
    uint32_t tiles_per_split = (uint32_t)((n_tiles + csplits - 1) / csplits);
    csplits = (uint32_t)((n_tiles + tiles_per_split - 1) / tiles_per_split);

    DevBuf pidx, pdist, pcnt;
    ExactArgs a{x->d, norms, x->rows, x->cols, metric, k, eps, query_rows, nq, q_begin, csplits, tiles_per_split,
                out_idx, out_dist, out_cnt};
    if (csplits > 1) {
        SFB_CUDA(ctx, pidx.alloc(sizeof(uint32_t) * nq * csplits * k));
        SFB_CUDA(ctx, pdist.alloc(sizeof(double) * nq * csplits * k));
        SFB_CUDA(ctx, pcnt.alloc(sizeof(uint32_t) * nq * csplits));
        a.out_idx = pidx.as<uint32_t>(); a.out_dist = pdist.as<double>(); a.out_cnt = pcnt.as<uint32_t>();
    }
    size_t smem = exact_smem_bytes(k);
    dim3 grid((unsigned)q_tiles, csplits);
    if (metric == SFB_METRIC_COSINE) {
        SFB_CUDA(ctx, cudaFuncSetAttribute(knn_exact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_exact_kernel<true><<<grid, THREADS, smem, ctx->stream>>>(a);
    } else {
        SFB_CUDA(ctx, cudaFuncSetAttribute(knn_exact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_exact_kernel<false><<<grid, THREADS, smem, ctx->stream>>>(a);
    }
    SFB_LAUNCH_CHECK(ctx);
    if (csplits > 1) {
        const int wpb = 4;
        size_t msmem = (size_t)wpb * k * (sizeof(double) + sizeof(uint32_t));
        knn_merge_kernel<<<div_up(nq, wpb), wpb * 32, msmem, ctx->stream>>>(pidx.as<uint32_t>(), pdist.as<double>(),
                                                                             pcnt.as<uint32_t>(), nq, csplits, k, out_idx,
                                                                             out_dist, out_cnt);
        SFB_LAUNCH_CHECK(ctx);
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // partial buffers die with this scope
    }
    return SFB_OK;
}
