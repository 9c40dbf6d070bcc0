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

    // a fallback call brings a handful of rows: thread groups whose four query rows are all padding skip the FP64 work
    // (they still stage tiles and meet the barriers)
    const bool rows_live = q0 + (uint64_t)ty * 4 < a.nq;
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
            if (rows_live)
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

// ---- a handful of query rows (the rows the screen could not certify: a few per million) ---------------------------
// The 64 x 64 tile kernel above leaves all but one thread group idle for such a call and stages its operands with
// synchronous loads: 4 ms for three rows at 1M x 384, latency-bound.  Here a thread owns one CORPUS row of a 128-row tile
// and carries the chains of all NQ query rows (NQ x 2 FP64 instructions per dimension on every lane), the corpus chunk
// (16 dimensions x 128 rows, row stride 17 doubles: conflict-free for lane = row) and the matching query chunk come
// through a cp.async ring that runs ahead across tile boundaries, and the selection is the same thresholded warp insert.
// Same left folds, same bits.  Reads the corpus once: HBM-bound from about four rows down, FP64-bound above.
constexpr int FEW_TC = 128, FEW_KC = 16, FEW_CS = 17, FEW_NST = 4;

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 8 : 0;   // src-size 0: zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

template <int NQ> __host__ __device__ constexpr size_t few_smem_doubles(uint32_t k) {
    return (size_t)FEW_NST * (FEW_TC * FEW_CS + FEW_KC * NQ) + (size_t)NQ * (FEW_TC + 1) + NQ + (size_t)NQ * k;
}

template <bool COS, int NQ>
__global__ void __launch_bounds__(FEW_TC) knn_exact_few_kernel(ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* cs = reinterpret_cast<double*>(smem_raw);                       // [NST][TC][CS]
    double* qs = cs + (size_t)FEW_NST * FEW_TC * FEW_CS;                   // [NST][KC][NQ]
    double* keys = qs + (size_t)FEW_NST * FEW_KC * NQ;                     // [NQ][TC + 1]
    double* qn = keys + (size_t)NQ * (FEW_TC + 1);                         // [NQ]
    double* ld = qn + NQ;                                                  // [NQ][k]
    uint32_t* li = reinterpret_cast<uint32_t*>(ld + (size_t)NQ * a.k);     // [NQ][k]
    uint32_t* gq = li + (size_t)NQ * a.k;                                  // [NQ]
    uint32_t* lcnt = gq + NQ;                                              // [NQ]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < NQ) {
        uint32_t g = SFB_IDX_NONE;
        if ((uint64_t)tid < a.nq) g = a.query_rows ? a.query_rows[tid] : (uint32_t)(a.q_begin + tid);
        gq[tid] = g; lcnt[tid] = 0;
        qn[tid] = (COS && g != SFB_IDX_NONE) ? a.norms[g] : 0.0;
    }
    __syncthreads();
    const uint64_t n_tiles = (a.m + FEW_TC - 1) / FEW_TC;
    const uint64_t t_begin = (uint64_t)blockIdx.x * a.tiles_per_split;
    const uint64_t t_end = t_begin + a.tiles_per_split < n_tiles ? t_begin + a.tiles_per_split : n_tiles;
    const uint32_t n_kc = (a.kd + FEW_KC - 1) / FEW_KC;
    const uint64_t n_steps = t_end > t_begin ? (t_end - t_begin) * n_kc : 0;   // (tile, chunk) pairs, in order

    auto stage = [&](uint64_t s) {
        if (s < n_steps) {
            const int buf = (int)(s % FEW_NST);
            const uint64_t tile = t_begin + s / n_kc;
            const uint32_t d0 = (uint32_t)(s % n_kc) * FEW_KC;
            double* cb = cs + (size_t)buf * FEW_TC * FEW_CS;
            // 128 rows x 16 dimensions, 8 bytes per copy: lanes walk a row's 128-byte run
#pragma unroll 4
            for (int e = tid; e < FEW_TC * FEW_KC; e += FEW_TC) {
                const int r = e / FEW_KC, d = e % FEW_KC;
                const uint64_t row = tile * FEW_TC + r;
                const bool v = row < a.m && d0 + d < a.kd;
                cp_async8(cb + r * FEW_CS + d, v ? a.x + row * a.kd + d0 + d : a.x, v);
            }
#pragma unroll
            for (int e = tid; e < NQ * FEW_KC; e += FEW_TC) {   // the query chunk, [dimension][query]
                const int sq = e / FEW_KC, sd = e % FEW_KC;
                const uint32_t sg = gq[sq];
                const bool v = sg != SFB_IDX_NONE && d0 + sd < a.kd;
                cp_async8(qs + (size_t)buf * FEW_KC * NQ + sd * NQ + sq, v ? a.x + (uint64_t)sg * a.kd + d0 + sd : a.x, v);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
#pragma unroll
    for (int c = 0; c < FEW_NST - 1; ++c) stage((uint64_t)c);
    for (uint64_t s = 0; s < n_steps; ++s) {
        const int buf = (int)(s % FEW_NST);
        stage(s + FEW_NST - 1);   // the slot consumed in the previous iteration (barrier at its end)
        asm volatile("cp.async.wait_group %0;" ::"n"(FEW_NST - 1) : "memory");
        __syncthreads();
        const double* cb = cs + (size_t)buf * FEW_TC * FEW_CS + tid * FEW_CS;
        const double* qb = qs + (size_t)buf * FEW_KC * NQ;
#pragma unroll
        for (int d = 0; d < FEW_KC; ++d) {
            const double c = cb[d];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (COS) acc[q] = __dadd_rn(acc[q], __dmul_rn(qb[d * NQ + q], c));
                else { const double t = __dadd_rn(qb[d * NQ + q], -c); acc[q] = __dadd_rn(acc[q], __dmul_rn(t, t)); }
            }
        }
        if ((s + 1) % n_kc == 0) {
            // end of a tile: keys, then the selection (warp w owns query rows w, w + 4, ...)
            const uint64_t c0 = (t_begin + s / n_kc) * FEW_TC, crow = c0 + tid;
            const double cn = (COS && crow < a.m) ? a.norms[crow] : 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                double key;
                if (COS) {
                    const double denom = __dmul_rn(qn[q], cn);
                    double cosv = 0.0;
                    if (denom > 1e-12) { cosv = __ddiv_rn(acc[q], denom); if (cosv < -1.0) cosv = -1.0; else if (cosv > 1.0) cosv = 1.0; }
                    const double rect = cosv > 0.0 ? cosv : 0.0;
                    key = __dadd_rn(1.0, -rect);
                } else key = a.metric == SFB_METRIC_L2 ? __dsqrt_rn(acc[q]) : acc[q];
                keys[q * (FEW_TC + 1) + tid] = key;
                acc[q] = 0.0;
            }
            __syncthreads();
            for (int q = warp; q < NQ; q += FEW_TC / 32) {
                const uint32_t g = gq[q];
                if (g == SFB_IDX_NONE) continue;
                double* rld = ld + (size_t)q * a.k;
                uint32_t* rli = li + (size_t)q * a.k;
                uint32_t c = lcnt[q];
#pragma unroll 1
                for (int part = 0; part < FEW_TC / 32; ++part) {
                    const double key = keys[q * (FEW_TC + 1) + part * 32 + lane];
                    const uint64_t j = c0 + part * 32 + lane;
                    const double thr_d = c == a.k ? rld[a.k - 1] : INFINITY;
                    const uint32_t thr_i = c == a.k ? rli[a.k - 1] : SFB_IDX_NONE;
                    const bool pr = j < a.m && (uint32_t)j != g && key <= a.eps && topk_key_less(key, (uint32_t)j, thr_d, thr_i);
                    uint32_t b = __ballot_sync(0xffffffffu, pr);
                    while (b) {
                        const int src = __ffs(b) - 1; b &= b - 1;
                        warp_list_insert(rld, rli, c, a.k, __shfl_sync(0xffffffffu, key, src), (uint32_t)(c0 + part * 32 + src), lane);
                    }
                }
                if (lane == 0) lcnt[q] = c;
            }
        }
        __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int q = warp; q < NQ; q += FEW_TC / 32) {
        if ((uint64_t)q >= a.nq) continue;
        const uint32_t c = lcnt[q];
        const size_t o = ((size_t)q * a.csplits + blockIdx.x) * a.k;
        for (uint32_t t = lane; t < a.k; t += 32) {
            a.out_idx[o + t] = t < c ? li[(size_t)q * a.k + t] : SFB_IDX_NONE;
            a.out_dist[o + t] = t < c ? ld[(size_t)q * a.k + t] : INFINITY;
        }
        if (lane == 0) a.out_cnt[(size_t)q * a.csplits + blockIdx.x] = c;
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
// Input layout: DIMS-MAJOR, xd[n * m + i] = node i at dimension n -- i.e. the item matrix itself (N x D);
// the reference transposes it first (graph.rs:216), here the transposed copy is never needed.
// Every (i, j) pair is one left fold over the N dimensions with separately rounded multiply and add,
// so each sum has the bits of the CPU restatement; the pairs are the parallelism.  A CTA of 64 threads
// owns a 16 x 16 pair tile (thread = 2 x 2 pairs: four independent FP64 chains), streams the two
// 16-column strips of xd through shared memory 64 dimensions at a time (cp.async, double buffered)
// and writes the raw sums G[i][j] (= G[j][i]: products commute) for tiles on or above the diagonal.
// The diagonal G[i][i] is the squared norm, the same left fold as row_norms_kernel.
// FP64-pipe bound: m^2/2 * N multiply + add pairs; at m = 384 the 300 tiles fill 148 SMs.
// GCH (template) = dimensions per staged chunk: 64 standalone (32 KB of shared memory), 16 when the kernel runs beside a
// resident screen CTA (8 KB fit next to its 216 of 227 KB)

__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gsrc, int bytes, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? bytes : 0;   // src-size 0: the destination is zero-filled
    if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

// GTILE: pair-tile edge.  16: thread = 2 x 2 pairs (four chains per thread).  8: thread = 1 pair -- four times as many,
// shorter-lived CTAs; used when a rank's share of the tiles would otherwise leave most SMs without a chain (every chain
// is N steps long whatever the tile count: with G ranks the 16-edge tiling gives each rank 300 / G CTAs of two warps).
// NST: stages of the cp.async ring.  A chunk is consumed in GCH * (GTILE / 8)^2 * 4 FP64-pipe cycles per warp -- a few
// hundred at most -- while a global -> shared copy takes ~1000: with two stages the kernel waited for memory four steps out
// of five (27 ns per fold step standalone, 45 ns in the 16-dimension variant).  NST - 1 chunks are kept in flight.
template <bool COS, int VEC, int GCH, int GTILE, int NST>  // VEC doubles per cp.async: 2 when m is even (16-byte aligned strips), else 1
__global__ void __launch_bounds__(64) gram_tile_kernel(const double* __restrict__ xd, uint32_t m, uint64_t kd, double* __restrict__ G, uint32_t tile0,
                                                       uint32_t n_tiles) {
    constexpr int R = GTILE / 8;
    __shared__ __align__(16) double sa[NST][GCH][GTILE], sb[NST][GCH][GTILE];
    const uint32_t T = (m + GTILE - 1) / GTILE;
    // a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...: beside the screen kernel the grid is capped at one CTA per SM
    for (uint32_t tt = blockIdx.x; tt < n_tiles; tt += gridDim.x) {
    // decode the upper-triangular tile index
    uint32_t ti = 0, rem = tt + tile0;
    while (rem >= T - ti) { rem -= T - ti; ++ti; }
    const uint32_t tj = ti + rem;
    const int tid = threadIdx.x, ty = tid >> 3, tx = tid & 7;
    const uint32_t ci = ti * GTILE, cj = tj * GTILE;
    const uint64_t n_chunks = (kd + GCH - 1) / GCH;

    auto stage = [&](uint64_t c) {   // chunk c -> ring slot c % NST; one commit group per call, empty past the end
        if (c < n_chunks) {
            const int buf = (int)(c % NST);
            const uint64_t n0 = c * GCH;
            constexpr int PIECES = GTILE / VEC;            // pieces per strip row
            for (int e = tid; e < GCH * PIECES; e += 64) {
                const int r = e / PIECES, pc = (e % PIECES) * VEC;
                const uint64_t n = n0 + r;
                const bool rv = n < kd;
                const double* src = xd + (rv ? n : 0) * m;
                const bool va = rv && ci + pc + VEC <= m, vb = rv && cj + pc + VEC <= m;
                cp_async_zfill(&sa[buf][r][pc], va ? src + ci + pc : xd, VEC * 8, va);
                cp_async_zfill(&sb[buf][r][pc], vb ? src + cj + pc : xd, VEC * 8, vb);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double acc[R][R];
#pragma unroll
    for (int di = 0; di < R; ++di)
#pragma unroll
        for (int dj = 0; dj < R; ++dj) acc[di][dj] = 0.0;
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) stage((uint64_t)c);
    for (uint64_t c = 0; c < n_chunks; ++c) {
        const int buf = (int)(c % NST);
        stage(c + NST - 1);   // its slot was consumed in the previous iteration (barrier at its end)
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 1) : "memory");
        __syncthreads();
        const uint64_t left = kd - c * GCH;
        const int lim = left < (uint64_t)GCH ? (int)left : GCH;
#pragma unroll 8
        for (int n = 0; n < lim; ++n) {
            double av[R], bv[R];
#pragma unroll
            for (int d = 0; d < R; ++d) { av[d] = sa[buf][n][R * ty + d]; bv[d] = sb[buf][n][R * tx + d]; }
#pragma unroll
            for (int di = 0; di < R; ++di)
#pragma unroll
                for (int dj = 0; dj < R; ++dj) {
                    if (COS) acc[di][dj] = __dadd_rn(acc[di][dj], __dmul_rn(av[di], bv[dj]));
                    else { const double t = __dadd_rn(av[di], -bv[dj]); acc[di][dj] = __dadd_rn(acc[di][dj], __dmul_rn(t, t)); }
                }
        }
        __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const uint32_t i0 = ci + R * ty, j0 = cj + R * tx;
#pragma unroll
    for (int di = 0; di < R; ++di)
#pragma unroll
        for (int dj = 0; dj < R; ++dj) {
            const uint32_t i = i0 + di, j = j0 + dj;
            if (i < m && j < m) { G[(uint64_t)i * m + j] = acc[di][dj]; G[(uint64_t)j * m + i] = acc[di][dj]; }
        }
    __syncthreads();
    }
}

// ---- the same chains without shared memory: one WARP per pair tile, operands in a register ring -----------------------
// A chain is N dependent FP64 adds whatever the tile shape, so the kernel's time is N x (time per fold step); the
// shared-memory kernel above pays a cp.async wait and two CTA barriers per chunk and can look ahead only as far as its
// ring is deep -- 48 steps in the 8 KB that fit beside a resident screen CTA, where a global load takes several
// microseconds (measured: ~100 ns per step co-resident, 17-26 ms of an 8-GPU C2 step exposed).  Here a warp owns a
// GT x GT tile (lane = (ty, tx) of a 4 x 8 grid, RA x RB chains per lane) and streams its two column strips straight
// into REGISTERS: a round is 8 fold steps, its 8 x GT + 8 x GT operands are RA + 2 RB coalesced loads (each lane holds a
// different element), D - 1 rounds are in flight, and a step fetches its operands from the holding lanes by shuffle.  No
// shared memory (any number of warps co-reside with the screen CTA: its 192 threads x 168 registers leave half the
// register file), no barriers, 8 (D - 1) steps of look-ahead (88 at GT = 8).
template <bool COS, int GT>
__global__ void __launch_bounds__(32, 12) gram_warp_kernel(const double* __restrict__ xd, uint32_t m, uint64_t kd, double* __restrict__ G,
                                                       uint32_t tile0, uint32_t n_tiles) {
    constexpr int RA = GT / 4, RB = GT / 8;      // rows ty + 4 da, columns tx + 8 db
    constexpr int NB = 2 * RB;                   // b loads per round: steps 0-3 and 4-7 of each 8-column group
    constexpr int D = GT == 8 ? 12 : 6;          // rounds in flight (D * (RA + NB) doubles of ring per lane)
    const int lane = threadIdx.x, ty = lane >> 3, tx = lane & 7;
    const uint32_t T = (m + GT - 1) / GT;
    const uint64_t n_rounds = (kd + 7) / 8;
    for (uint32_t tt = blockIdx.x; tt < n_tiles; tt += gridDim.x) {
        uint32_t ti = 0, rem = tt + tile0;
        while (rem >= T - ti) { rem -= T - ti; ++ti; }
        const uint32_t ci = ti * GT, cj = (ti + rem) * GT;
        // this lane's slots of a round: a[q] = x[n0 + lane / 4][ci + lane % 4 + 4 q], b[q] = x[n0 + lane / 8 + 4 (q & 1)][cj + lane % 8 + 8 (q / 2)]
        const uint32_t a_s = lane >> 2, a_c = ci + (lane & 3), b_s = lane >> 3, b_c = cj + (lane & 7);
        const double* pa = xd + (uint64_t)a_s * m + a_c;      // both advance 8 rows per round
        const double* pb = xd + (uint64_t)b_s * m + b_c;
        const uint64_t m4 = 4ull * m, m8 = 8ull * m;
        uint32_t colmask = 0;                                  // bit q: a column in range, bit 8 + g: b column group in range
#pragma unroll
        for (int q = 0; q < RA; ++q) colmask |= (a_c + 4 * q < m ? 1u : 0u) << q;
#pragma unroll
        for (int g = 0; g < RB; ++g) colmask |= (b_c + 8 * g < m ? 1u : 0u) << (8 + g);
        uint64_t n_next = 0;                                   // first row of the next round to load
        double ra[D][RA], rb[D][NB];
        auto load_round = [&](double (&a)[RA], double (&b)[NB]) {
            const bool va = n_next + a_s < kd, vb0 = n_next + b_s < kd, vb1 = n_next + b_s + 4 < kd;
#pragma unroll
            for (int q = 0; q < RA; ++q) a[q] = (va && (colmask >> q & 1u)) ? __ldg(pa + 4 * q) : 0.0;
#pragma unroll
            for (int q = 0; q < NB; ++q) b[q] = (((q & 1) ? vb1 : vb0) && (colmask >> (8 + (q >> 1)) & 1u)) ? __ldg(pb + (q & 1) * m4 + 8 * (q >> 1)) : 0.0;
            pa += m8; pb += m8; n_next += 8;
        };
        double acc[RA][RB];
#pragma unroll
        for (int da = 0; da < RA; ++da)
#pragma unroll
            for (int db = 0; db < RB; ++db) acc[da][db] = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) load_round(ra[d], rb[d]);
        // rounds past the end hold zeros: x + 0 * 0 = x, the chain's bits do not change
        for (uint64_t base = 0; base < n_rounds; base += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    double av[RA], bv[RB];
#pragma unroll
                    for (int da = 0; da < RA; ++da) av[da] = __shfl_sync(0xffffffffu, ra[d][da], s * 4 + ty);
#pragma unroll
                    for (int db = 0; db < RB; ++db) bv[db] = __shfl_sync(0xffffffffu, rb[d][2 * db + (s >> 2)], (s & 3) * 8 + tx);
#pragma unroll
                    for (int da = 0; da < RA; ++da)
#pragma unroll
                        for (int db = 0; db < RB; ++db) {
                            if (COS) acc[da][db] = __dadd_rn(acc[da][db], __dmul_rn(av[da], bv[db]));
                            else { const double t = __dadd_rn(av[da], -bv[db]); acc[da][db] = __dadd_rn(acc[da][db], __dmul_rn(t, t)); }
                        }
                }
                load_round(ra[d], rb[d]);   // refill the slot just consumed: D - 1 rounds stay in flight
            }
        }
#pragma unroll
        for (int da = 0; da < RA; ++da)
#pragma unroll
            for (int db = 0; db < RB; ++db) {
                const uint32_t i = ci + ty + 4 * da, j = cj + tx + 8 * db;
                if (i < m && j < m) { G[(uint64_t)i * m + j] = acc[da][db]; G[(uint64_t)j * m + i] = acc[da][db]; }
            }
    }
}

// raw sums G -> distance keys, written behind G (cosine: the norms are the square roots of the diagonal)
template <bool COS>
__global__ void gram_keys_kernel(double* __restrict__ G, uint32_t m, int metric, double* __restrict__ norms) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= m) return;
    const double acc = G[(uint64_t)i * m + j];
    double key;
    if (COS) {
        const double ni = __dsqrt_rn(G[(uint64_t)i * m + i]), nj = __dsqrt_rn(G[(uint64_t)j * m + j]);
        double denom = __dmul_rn(ni, nj), cosv = 0.0;
        if (denom > 1e-12) { cosv = __ddiv_rn(acc, denom); if (cosv < -1.0) cosv = -1.0; else if (cosv > 1.0) cosv = 1.0; }
        const double rect = cosv > 0.0 ? cosv : 0.0;
        key = __dadd_rn(1.0, -rect);
        if (i == 0 && norms) norms[j] = nj;
    } else key = metric == SFB_METRIC_L2 ? __dsqrt_rn(acc) : acc;
    G[(uint64_t)m * m + (uint64_t)i * m + j] = key;
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

// top-k (key asc, index asc) of every row of a dense m x m key matrix (shared with the Bhattacharyya graph, bc.cu)
int32_t sfb_dense_select(sfb_ctx* ctx, const double* keys, uint32_t m, uint64_t q_begin, uint64_t nq, uint32_t k, double eps,
                         uint32_t* out_idx, double* out_dist, uint32_t* out_cnt) {
    const int wpb = 4;
    size_t ssm = (size_t)wpb * k * (sizeof(double) + sizeof(uint32_t));
    knn_dense_select_kernel<<<div_up(nq, wpb), wpb * 32, ssm, ctx->stream>>>(keys, m, q_begin, nq, k, eps, out_idx, out_dist, out_cnt);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

bool sfb_dense_shape(uint64_t nodes, uint64_t dims) { return nodes <= 4096 && dims >= 8ull * nodes; }

// kNN over few nodes with very long rows, from the DIMS-MAJOR matrix xd[kd][m] (see gram_tile_kernel).
int32_t sfb_comm_allreduce_sum_f64(sfb_ctx* ctx, double* buf, size_t n);

// Pair-tile edge for this call: 16 (thread = 2 x 2 pairs), or 8 (thread = one pair) when the 16-edge tiling would give this
// rank fewer CTAs than a quarter of its SMs -- the sharded builds at 8 GPUs.  SFB_GRAM_GT=8 / 16 forces it (tests, A/B).
// 32 (thread = 4 x 4 pairs) when even that tiling gives every SM four CTAs: with thousands of nodes (C4: 3072) the chains are
// plentiful and the kernel is bound by the FP64 pipe, where 16 chains per thread halve the shared-memory reads per multiply-add.
uint32_t sfb_gram_tile_edge(const sfb_ctx* ctx, uint32_t m, int collective) {
    if (const char* e = getenv("SFB_GRAM_GT")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) return (uint32_t)v; }
    const uint32_t T = (m + 15) / 16, tiles = T * (T + 1) / 2;
    const uint32_t world = collective && ctx->world > 1 ? (uint32_t)ctx->world : 1u;
    const uint32_t T32 = (m + 31) / 32, tiles32 = T32 * (T32 + 1) / 2;
    if (tiles32 / world >= 4u * (uint32_t)ctx->sm_count) return 32u;
    return (tiles + world - 1) / world * 4 <= (uint32_t)ctx->sm_count ? 8u : 16u;
}

// raw pair sums of tiles [t0, t1) into g (m x m doubles) on `stream`; small_smem: the 8 KB variant that co-resides
// with the screen kernel
int32_t sfb_gram_launch(sfb_ctx* ctx, cudaStream_t stream, const double* xd, uint32_t m, uint64_t kd, int metric, double* g,
                        uint32_t gt, uint32_t t0, uint32_t t1, bool small_smem) {
    if (t1 <= t0) return SFB_OK;
    const uint32_t nt = t1 - t0;
    // beside a screen: the register-ring kernel (hidden completely at one rank's share of C2 at 8 GPUs, where the 8 KB ring left
    // 17-32 ms exposed); alone: the shared-memory ring (14 against 24 ns per fold step).  SFB_GRAM_SMEM / SFB_GRAM_REGS force one.
    if (gt != 32 && (small_smem ? !getenv("SFB_GRAM_SMEM") : getenv("SFB_GRAM_REGS") != nullptr)) {
        // one warp per tile; beside the screen kernel at most two warps per SM walk the tiles (the screen CTA's registers must
        // still fit wherever the block scheduler puts them)
        const uint32_t cap = 2u * (uint32_t)ctx->sm_count, wgrid = small_smem && nt > cap ? cap : nt;
        const bool c = metric == SFB_METRIC_COSINE;
        if (gt == 8) { if (c) gram_warp_kernel<true, 8><<<wgrid, 32, 0, stream>>>(xd, m, kd, g, t0, nt); else gram_warp_kernel<false, 8><<<wgrid, 32, 0, stream>>>(xd, m, kd, g, t0, nt); }
        else { if (c) gram_warp_kernel<true, 16><<<wgrid, 32, 0, stream>>>(xd, m, kd, g, t0, nt); else gram_warp_kernel<false, 16><<<wgrid, 32, 0, stream>>>(xd, m, kd, g, t0, nt); }
        SFB_LAUNCH_CHECK(ctx);
        return SFB_OK;
    }
    const bool cos = metric == SFB_METRIC_COSINE, even = (m & 1u) == 0 && (reinterpret_cast<uintptr_t>(xd) & 15u) == 0;
    // beside the screen kernel: at most one CTA per SM, so that a CTA placed before the screen's never keeps the
    // screen's 216 KB from fitting (two of them would)
    const uint32_t grid = small_smem && nt > (uint32_t)ctx->sm_count ? (uint32_t)ctx->sm_count : nt;
#define SFB_GRAM(C_, V_, G_, T_, S_) gram_tile_kernel<C_, V_, G_, T_, S_><<<grid, 64, 0, stream>>>(xd, m, kd, g, t0, nt)
#define SFB_GRAM_CV(G_, T_, S_)                                                               \
    do {                                                                                      \
        if (cos) { if (even) SFB_GRAM(true, 2, G_, T_, S_); else SFB_GRAM(true, 1, G_, T_, S_); }   \
        else { if (even) SFB_GRAM(false, 2, G_, T_, S_); else SFB_GRAM(false, 1, G_, T_, S_); }     \
    } while (0)
    // shared memory: 2 * NST * GCH * GTILE * 8 bytes -- 8 KB beside the screen (its CTA leaves ~11 KB), 32 KB standalone
    if (gt == 8) { if (small_smem) SFB_GRAM_CV(16, 8, 4); else SFB_GRAM_CV(32, 8, 8); }
    else if (gt == 32) SFB_GRAM_CV(16, 32, 3);   // 24 KB; never the choice beside a screen
    else { if (small_smem) SFB_GRAM_CV(16, 16, 2); else SFB_GRAM_CV(32, 16, 4); }
#undef SFB_GRAM_CV
#undef SFB_GRAM
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

void sfb_gram_tile_range(const sfb_ctx* ctx, uint32_t m, uint32_t gt, int collective, uint32_t* t0, uint32_t* t1) {
    const uint32_t T = (m + gt - 1) / gt, tiles = T * (T + 1) / 2;
    *t0 = 0; *t1 = tiles;
    if (collective && ctx->world > 1) {
        const uint32_t per = (tiles + (uint32_t)ctx->world - 1) / (uint32_t)ctx->world;
        *t0 = (uint32_t)ctx->rank * per; if (*t0 > tiles) *t0 = tiles;
        *t1 = *t0 + per < tiles ? *t0 + per : tiles;
    }
}

// g: [0, m*m) raw sums (complete on this rank, or this rank's tile share when `collective`), [m*m, 2*m*m) scratch for
// the keys.  All-reduces the sums if needed, turns them into distance keys and selects the top-k of the query rows.
int32_t sfb_gram_finish(sfb_ctx* ctx, double* g, uint32_t m, int metric, uint32_t k, double eps, uint64_t q_begin, uint64_t nq,
                        uint32_t* out_idx, double* out_dist, uint32_t* out_cnt, int collective) {
    if (collective && ctx->world > 1) SFB_TRY(sfb_comm_allreduce_sum_f64(ctx, g, (size_t)m * m));
    dim3 kgrid(div_up(m, 128), m);
    if (metric == SFB_METRIC_COSINE) gram_keys_kernel<true><<<kgrid, 128, 0, ctx->stream>>>(g, m, metric, nullptr);
    else gram_keys_kernel<false><<<kgrid, 128, 0, ctx->stream>>>(g, m, metric, nullptr);
    SFB_LAUNCH_CHECK(ctx);
    return sfb_dense_select(ctx, g + (size_t)m * m, m, q_begin, nq, k, eps, out_idx, out_dist, out_cnt);
}

// `collective` != 0: every rank of the communicator calls this with the same matrix; the pair tiles are split
// across the ranks and the raw sums all-reduced (each entry is non-zero on exactly one rank: x + 0 is exact).
int32_t sfb_knn_dense(sfb_ctx* ctx, const double* xd, uint32_t m, uint64_t kd, int metric, uint32_t k, double eps,
                      uint64_t q_begin, uint64_t nq, uint32_t* out_idx, double* out_dist, uint32_t* out_cnt, int collective) {
    if (nq == 0) return SFB_OK;
    if (k == 0 || k > 128) return sfb_fail(ctx, SFB_EUNSUPPORTED, "k must be in 1..128 (got %u)", k);
    DevBuf g;  // [0, m*m): raw sums, [m*m, 2*m*m): keys
    SFB_CUDA(ctx, g.alloc(sizeof(double) * 2 * (size_t)m * m));
    uint32_t t0, t1;
    const uint32_t gt = sfb_gram_tile_edge(ctx, m, collective);
    sfb_gram_tile_range(ctx, m, gt, collective, &t0, &t1);
    if (collective && ctx->world > 1) SFB_CUDA(ctx, cudaMemsetAsync(g.p, 0, sizeof(double) * (size_t)m * m, ctx->stream));
    SFB_TRY(sfb_gram_launch(ctx, ctx->stream, xd, m, kd, metric, g.as<double>(), gt, t0, t1, false));
    return sfb_gram_finish(ctx, g.as<double>(), m, metric, k, eps, q_begin, nq, out_idx, out_dist, out_cnt, collective);
}

int32_t sfb_knn_exact(sfb_ctx* ctx, const sfb_mat* x, const double* norms, int metric, uint32_t k, double eps,
                      const uint32_t* query_rows, uint64_t nq, uint64_t q_begin, uint32_t* out_idx, double* out_dist,
                      uint32_t* out_cnt) {
    if (nq == 0) return SFB_OK;
    if (k == 0 || k > 128) return sfb_fail(ctx, SFB_EUNSUPPORTED, "k must be in 1..128 (got %u)", k);
    if (nq <= 16 && x->rows >= 16384 && !getenv("SFB_EXACT_NO_FEW")) {
        // a handful of rows against a long corpus: knn_exact_few_kernel, the corpus split over every SM, then the merge
        const uint64_t n_tiles_f = (x->rows + FEW_TC - 1) / FEW_TC;
        const int nqt = nq <= 4 ? 4 : (nq <= 8 ? 8 : 16);
        const size_t smem = (nqt == 4 ? few_smem_doubles<4>(k) : nqt == 8 ? few_smem_doubles<8>(k) : few_smem_doubles<16>(k)) * sizeof(double) +
                            sizeof(uint32_t) * ((size_t)nqt * k + 2 * nqt);
        int per_sm = (int)((ctx->smem_optin ? ctx->smem_optin : 232448) / (smem + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm >= 1) {
            uint64_t want = (uint64_t)ctx->sm_count * per_sm;
            if (want > n_tiles_f) want = n_tiles_f;
            const uint32_t tps = (uint32_t)((n_tiles_f + want - 1) / want);
            const uint32_t splits = (uint32_t)((n_tiles_f + tps - 1) / tps);
            DevBuf pidx, pdist, pcnt;
            SFB_CUDA(ctx, pidx.alloc(sizeof(uint32_t) * nq * splits * k));
            SFB_CUDA(ctx, pdist.alloc(sizeof(double) * nq * splits * k));
            SFB_CUDA(ctx, pcnt.alloc(sizeof(uint32_t) * nq * splits));
            ExactArgs a{x->d, norms, x->rows, x->cols, metric, k, eps, query_rows, nq, q_begin, splits, tps,
                        pidx.as<uint32_t>(), pdist.as<double>(), pcnt.as<uint32_t>()};
            const bool cos = metric == SFB_METRIC_COSINE;
#define SFB_FEW(C_, Q_)                                                                                                        \
    do {                                                                                                                       \
        SFB_CUDA(ctx, cudaFuncSetAttribute(knn_exact_few_kernel<C_, Q_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        knn_exact_few_kernel<C_, Q_><<<splits, FEW_TC, smem, ctx->stream>>>(a);                                                \
    } while (0)
            if (cos) { if (nqt == 4) SFB_FEW(true, 4); else if (nqt == 8) SFB_FEW(true, 8); else SFB_FEW(true, 16); }
            else { if (nqt == 4) SFB_FEW(false, 4); else if (nqt == 8) SFB_FEW(false, 8); else SFB_FEW(false, 16); }
#undef SFB_FEW
            SFB_LAUNCH_CHECK(ctx);
            const int wpb = 4;
            const size_t msmem = (size_t)wpb * k * (sizeof(double) + sizeof(uint32_t));
            knn_merge_kernel<<<div_up(nq, wpb), wpb * 32, msmem, ctx->stream>>>(pidx.as<uint32_t>(), pdist.as<double>(), pcnt.as<uint32_t>(), nq, splits, k,
                                                                                 out_idx, out_dist, out_cnt);
            SFB_LAUNCH_CHECK(ctx);
            SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // partial buffers die with this scope
            return SFB_OK;
        }
    }
    const uint64_t q_tiles = (nq + TQ - 1) / TQ, n_tiles = (x->rows + TC - 1) / TC;
    // one query tile (a fallback of a few rows): only one or two warps per CTA carry FP64 work, so spread the corpus
    // over four CTAs per SM instead of two
    uint64_t want = (q_tiles == 1 ? 4ull : 2ull) * ctx->sm_count;
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
