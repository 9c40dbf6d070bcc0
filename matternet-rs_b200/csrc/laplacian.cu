// laplacian.cu -- kernel weights, sparsification, symmetrisation and CSR Laplacian assembly.
//
// Reference: src_legacy/laplacian.rs:231-290 (weights + inline sparsification), :297-348
// (symmetrise), :351-419 + :161 (L = D - W as CSR); src_legacy/sparsification.rs:32-113 (SF-GRASS);
// surfface-core/src/laplacian.rs:333-372,209-219 (normalised L_sym).
//
// The reference symmetrises with an O(M*E) DashMap scan and assembles through a sequential TriMat
// fill.  Here: reverse edges are bucketed by destination with a counting sort (histogram, exclusive
// scan, scatter), then each row is handled by one warp (one block for rows longer than 256): forward
// and reverse entries are bitonic-sorted by column in shared memory, duplicates merged (max), the
// degree is a left fold in ascending column order (same order as the reference), and the CSR row is
// written with coalesced stores.  HBM-bound: reads M*k*12 B of lists, writes (M+1)*8 + nnz*12 B.
#include <math.h>

#include <new>

#include "common.cuh"

int32_t sfb_adj_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t k, sfb_adj** out);

namespace {

constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ double kernel_weight(double d, double sigma, double p, int pmode) {
    double r = __ddiv_rn(d, sigma);
    // (d/sigma)^p: p = 1 and p = 2 are exact roundings; other exponents go through pow (<= 2 ulp)
    double t = pmode == 1 ? r : (pmode == 2 ? __dmul_rn(r, r) : pow(r, p));
    return __ddiv_rn(1.0, __dadd_rn(1.0, t));
}

__device__ __forceinline__ bool score_before(double sa, uint32_t ja, double sb, uint32_t jb) {
    return sa > sb || (sa == sb && ja < jb);  // (score desc, j asc)
}

__global__ void sum_u32_kernel(const uint32_t* __restrict__ v, uint64_t n, unsigned long long* out) {
    unsigned long long s = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s += v[i];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// One warp per row, E = ceil(k/32) entries per lane.
// MODE 0: weights from distances (+ inline sparsification when `sparsify`)  laplacian.rs:245-290
// MODE 1: SF-GRASS on existing weights                                       sparsification.rs:63-101
template <int E, int MODE>
__global__ void adjacency_kernel(const uint32_t* in_idx, const double* in_val,
                                 const uint32_t* in_cnt, const uint32_t* deg, uint64_t m,
                                 uint32_t k, double p, double sigma, int pmode, int sparsify, double ratio,
                                 uint32_t* out_idx, double* out_w, uint32_t* out_cnt) {
    const int lane = threadIdx.x & 31;
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= m) return;
    const uint32_t c = in_cnt[i];
    uint32_t j[E]; double w[E], sc[E]; bool keep[E];
    uint32_t len = 0;
#pragma unroll
    for (int u = 0; u < E; ++u) {
        uint32_t t = lane + 32 * u;
        bool valid = t < c;
        j[u] = valid ? in_idx[i * k + t] : SFB_IDX_NONE;
        double v = valid ? in_val[i * k + t] : 0.0;
        w[u] = MODE == 0 ? (valid ? kernel_weight(v, sigma, p, pmode) : 0.0) : v;
        keep[u] = valid && (MODE == 1 || w[u] > 1e-12);
        len += __popc(__ballot_sync(FULL, keep[u]));
    }
    bool select = MODE == 1 ? (len > 0) : (sparsify && len > 2);
    uint32_t keep_count = len;
    if (select) {
        if (MODE == 0) { keep_count = len / 2; if (keep_count < 1) keep_count = 1; }
        else {
            keep_count = (uint32_t)ceil(__dmul_rn((double)len, ratio));
            if (keep_count < 1) keep_count = 1;
            if (keep_count > len) keep_count = len;
        }
        const uint32_t di = MODE == 0 ? c : deg[i];
#pragma unroll
        for (int u = 0; u < E; ++u) {
            uint32_t dj = keep[u] ? (MODE == 0 ? in_cnt[j[u]] : deg[j[u]]) : 0u;
            sc[u] = __dmul_rn(w[u], __dsqrt_rn((double)((unsigned long long)di * (unsigned long long)dj)));
        }
    }
    uint32_t pos[E];
    if (select) {
        // rank of every kept entry in (score desc, j asc)
#pragma unroll
        for (int u = 0; u < E; ++u) pos[u] = 0;
#pragma unroll
        for (int v = 0; v < E; ++v)
            for (int src = 0; src < 32; ++src) {
                double s2 = __shfl_sync(FULL, sc[v], src);
                uint32_t j2 = __shfl_sync(FULL, j[v], src);
                bool k2 = __shfl_sync(FULL, (int)keep[v], src);
                if (!k2) continue;
#pragma unroll
                for (int u = 0; u < E; ++u) pos[u] += score_before(s2, j2, sc[u], j[u]) ? 1u : 0u;
            }
    } else {
        // compact in the original (distance asc) order
        uint32_t base = 0;
#pragma unroll
        for (int u = 0; u < E; ++u) {
            uint32_t b = __ballot_sync(FULL, keep[u]);
            pos[u] = base + __popc(b & ((1u << lane) - 1u));
            base += __popc(b);
        }
    }
    // pad first, then place the kept entries (a row is owned by this warp only)
#pragma unroll
    for (int u = 0; u < E; ++u) {
        uint32_t t = lane + 32 * u;
        if (t < k && t >= keep_count) { out_idx[i * k + t] = SFB_IDX_NONE; out_w[i * k + t] = 0.0; }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < E; ++u)
        if (keep[u] && pos[u] < keep_count) { out_idx[i * k + pos[u]] = j[u]; out_w[i * k + pos[u]] = w[u]; }
    if (lane == 0) out_cnt[i] = keep_count;
}

// ---- symmetrise -------------------------------------------------------------------------------
__global__ void rev_count_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ cnt, uint64_t m, uint32_t k,
                                 uint32_t* __restrict__ rev_cnt) {
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * k) return;
    uint64_t i = gid / k; uint32_t t = (uint32_t)(gid % k);
    if (t >= cnt[i]) return;
    uint32_t j = idx[gid];
    if (j != (uint32_t)i) atomicAdd(&rev_cnt[j], 1u);
}
__global__ void rev_scatter_kernel(const uint32_t* __restrict__ idx, const double* __restrict__ w, const uint32_t* __restrict__ cnt,
                                   uint64_t m, uint32_t k, const uint64_t* __restrict__ rev_off, uint32_t* __restrict__ fill,
                                   uint32_t* __restrict__ rev_src, double* __restrict__ rev_w) {
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * k) return;
    uint64_t i = gid / k; uint32_t t = (uint32_t)(gid % k);
    if (t >= cnt[i]) return;
    uint32_t j = idx[gid];
    if (j == (uint32_t)i) return;
    uint64_t o = rev_off[j] + atomicAdd(&fill[j], 1u);
    rev_src[o] = (uint32_t)i; rev_w[o] = w[gid];
}
__global__ void row_len_kernel(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ rev_cnt, uint64_t m,
                               uint32_t* __restrict__ len, uint32_t* __restrict__ max_len) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t l = 0;
    if (i < m) { l = cnt[i] + rev_cnt[i]; len[i] = l; }
    for (int o = 16; o; o >>= 1) l = max(l, __shfl_down_sync(FULL, l, o));
    if ((threadIdx.x & 31) == 0 && l) atomicMax(max_len, l);
}

// Bitonic sort of P (power of two) (col, w) pairs in shared memory by (col asc, w desc), by NT
// cooperating threads (a warp: sync = __syncwarp; a block: __syncthreads).
template <bool BLOCK>
__device__ __forceinline__ void group_sync() { if (BLOCK) __syncthreads(); else __syncwarp(); }

template <bool BLOCK>
__device__ void bitonic_sort_pairs(uint32_t* col, double* w, uint32_t P, uint32_t tid, uint32_t nt) {
    for (uint32_t size = 2; size <= P; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            group_sync<BLOCK>();
            for (uint32_t t = tid; t < P / 2; t += nt) {
                uint32_t lo = 2 * t - (t & (stride - 1));  // index with the `stride` bit clear
                uint32_t hi = lo + stride;
                bool up = (lo & size) == 0;
                uint32_t ca = col[lo], cb = col[hi];
                double wa = w[lo], wb = w[hi];
                bool a_after_b = ca > cb || (ca == cb && wa < wb);
                if (a_after_b == up) { col[lo] = cb; col[hi] = ca; w[lo] = wb; w[hi] = wa; }
            }
        }
    group_sync<BLOCK>();
}

// Pass A: per row, merge forward + reverse entries, sort by column, dedupe (max), write the unique
// neighbours to tmp at tmp_off[i], the unique count to ulen[i], the degree (left fold, ascending
// column) to deg[i].  BLOCK = false: one warp per row, rows with len <= CAP; BLOCK = true: one block
// per row from `row_list`.
template <bool BLOCK, uint32_t CAP>
__global__ void lap_merge_rows_kernel(const uint32_t* __restrict__ a_idx, const double* __restrict__ a_w,
                                      const uint32_t* __restrict__ a_cnt, uint32_t k, const uint64_t* __restrict__ rev_off,
                                      const uint32_t* __restrict__ rev_src, const double* __restrict__ rev_w,
                                      const uint32_t* __restrict__ len, const uint64_t* __restrict__ tmp_off, uint64_t m,
                                      const uint32_t* __restrict__ row_list, uint32_t n_list, uint32_t warp_cap,
                                      uint32_t* __restrict__ tmp_col, double* __restrict__ tmp_w, uint32_t* __restrict__ ulen,
                                      double* __restrict__ deg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t nt = BLOCK ? blockDim.x : 32;
    const uint32_t tid = BLOCK ? threadIdx.x : (threadIdx.x & 31);
    const uint32_t groups = BLOCK ? 1 : blockDim.x >> 5, grp = BLOCK ? 0 : threadIdx.x >> 5;
    double* sw = reinterpret_cast<double*>(smem_raw) + (size_t)grp * CAP;
    uint32_t* scol = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)groups * CAP) + (size_t)grp * CAP;
    __shared__ uint32_t s_count;

    uint64_t i;
    if (BLOCK) { if (blockIdx.x >= n_list) return; i = row_list[blockIdx.x]; }
    else { i = (uint64_t)blockIdx.x * groups + grp; if (i >= m) return; }
    const uint32_t l = len[i];
    if (!BLOCK && l > warp_cap) return;  // long rows go to the block kernel
    if (l == 0) { if (tid == 0) { ulen[i] = 0; deg[i] = 0.0; } return; }
    uint32_t P = 2; while (P < l) P <<= 1;
    const uint32_t fc = a_cnt[i];
    const uint64_t ro = rev_off[i];
    for (uint32_t t = tid; t < P; t += nt) {
        uint32_t c = SFB_IDX_NONE; double w = 0.0;
        if (t < fc) { c = a_idx[i * k + t]; w = a_w[i * k + t]; if (c == (uint32_t)i) c = SFB_IDX_NONE; }
        else if (t < l) { c = rev_src[ro + (t - fc)]; w = rev_w[ro + (t - fc)]; }
        scol[t] = c; sw[t] = w;
    }
    bitonic_sort_pairs<BLOCK>(scol, sw, P, tid, nt);
    // heads of runs of equal column; sorted (col asc, w desc) => the head carries the max weight
    const uint64_t to = tmp_off[i];
    uint32_t base = 0;
    if (BLOCK) { if (tid == 0) s_count = 0; __syncthreads(); }
    for (uint32_t t0 = 0; t0 < P; t0 += nt) {
        uint32_t t = t0 + tid;
        bool head = t < P && scol[t] != SFB_IDX_NONE && (t == 0 || scol[t] != scol[t - 1]);
        uint32_t b = __ballot_sync(FULL, head);
        uint32_t pos;
        if (BLOCK) {
            // per-warp slots reserved in order: warps of the block handle ascending t ranges, so a
            // block-wide ordered scan is needed; do it with one shared counter per 32-chunk serially
            __shared__ uint32_t warp_base[32];
            if ((tid & 31) == 0) warp_base[tid >> 5] = __popc(b);
            __syncthreads();
            uint32_t pre = 0;
            for (uint32_t wv = 0; wv < (tid >> 5); ++wv) pre += warp_base[wv];
            uint32_t tot = 0;
            for (uint32_t wv = 0; wv < (nt >> 5); ++wv) tot += warp_base[wv];
            pos = s_count + pre + __popc(b & ((1u << (tid & 31)) - 1u));
            __syncthreads();
            if (tid == 0) s_count += tot;
            __syncthreads();
        } else {
            pos = base + __popc(b & ((1u << tid) - 1u));
            base += __popc(b);
        }
        if (head) { tmp_col[to + pos] = scol[t]; tmp_w[to + pos] = sw[t]; }
    }
    group_sync<BLOCK>();
    const uint32_t u = BLOCK ? s_count : base;
    if (tid == 0) {
        // degree: left fold over the unique neighbours in ascending column order (laplacian.rs:371)
        double s = 0.0;
        uint32_t prev = SFB_IDX_NONE;
        for (uint32_t t = 0; t < P; ++t) {
            uint32_t c = scol[t];
            if (c == SFB_IDX_NONE) break;
            if (c != prev) s = __dadd_rn(s, sw[t]);
            prev = c;
        }
        ulen[i] = u; deg[i] = s;
    }
}

// rows longer than `warp_cap`: compact their indices
__global__ void long_rows_kernel(const uint32_t* __restrict__ len, uint64_t m, uint32_t warp_cap, uint32_t* __restrict__ list,
                                 uint32_t* __restrict__ n_list) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m && len[i] > warp_cap) list[atomicAdd(n_list, 1u)] = (uint32_t)i;
}

// Pass B: CSR row lengths.
__global__ void csr_row_nnz_kernel(const uint32_t* __restrict__ ulen, const double* __restrict__ deg,
                                   const uint64_t* __restrict__ tmp_off, const uint32_t* __restrict__ tmp_col,
                                   const double* __restrict__ tmp_w, uint64_t m, int normalised, double thr,
                                   uint32_t* __restrict__ row_nnz) {
    const int lane = threadIdx.x & 31;
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= m) return;
    const uint32_t u = ulen[i];
    if (!normalised) { if (lane == 0) row_nnz[i] = u + 1; return; }
    const double di = deg[i];
    uint32_t n = 0;
    if (di > thr) {
        const uint64_t to = tmp_off[i];
        for (uint32_t t0 = 0; t0 < u; t0 += 32) {
            uint32_t t = t0 + lane;
            bool ok = false;
            if (t < u) {
                double dj = deg[tmp_col[to + t]];
                if (dj > thr) { double v = -__ddiv_rn(tmp_w[to + t], __dsqrt_rn(__dmul_rn(di, dj))); ok = fabs(v) > 1e-9; }
            }
            n += __popc(__ballot_sync(FULL, ok));
        }
        n += 1;  // diagonal 1.0
    }
    if (lane == 0) row_nnz[i] = n;
}

// Pass C: emit CSR rows (one warp per row, any length), diagonal inserted in column order.
__global__ void csr_emit_kernel(const uint32_t* __restrict__ ulen, const double* __restrict__ deg,
                                const uint64_t* __restrict__ tmp_off, const uint32_t* __restrict__ tmp_col,
                                const double* __restrict__ tmp_w, uint64_t m, int normalised, double thr,
                                const uint64_t* __restrict__ indptr, uint32_t* __restrict__ indices, double* __restrict__ data) {
    const int lane = threadIdx.x & 31;
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= m) return;
    const uint32_t u = ulen[i];
    const double di = deg[i];
    const uint64_t to = tmp_off[i];
    const uint64_t o = indptr[i];
    if (normalised && !(di > thr)) return;  // isolated node: empty row (surfface-core laplacian.rs:349-353)
    const double diag_val = normalised ? 1.0 : di;
    // phase 1: kept entries left of the diagonal
    uint32_t n_left = 0;
    for (uint32_t t0 = 0; t0 < u; t0 += 32) {
        uint32_t t = t0 + lane;
        bool left = false;
        if (t < u) {
            uint32_t c = tmp_col[to + t];
            bool ok = true;
            if (normalised) {
                double dj = deg[c];
                ok = dj > thr && fabs(__ddiv_rn(tmp_w[to + t], __dsqrt_rn(__dmul_rn(di, dj)))) > 1e-9;
            }
            left = ok && c < (uint32_t)i;
        }
        n_left += __popc(__ballot_sync(FULL, left));
    }
    if (lane == 0) { indices[o + n_left] = (uint32_t)i; data[o + n_left] = diag_val; }
    // phase 2: place the kept entries; those right of the diagonal shift by one
    uint32_t done = 0;
    for (uint32_t t0 = 0; t0 < u; t0 += 32) {
        uint32_t t = t0 + lane;
        uint32_t c = SFB_IDX_NONE; double v = 0.0; bool ok = false;
        if (t < u) {
            c = tmp_col[to + t];
            if (!normalised) { v = -tmp_w[to + t]; ok = true; }
            else {
                double dj = deg[c];
                if (dj > thr) { v = -__ddiv_rn(tmp_w[to + t], __dsqrt_rn(__dmul_rn(di, dj))); ok = fabs(v) > 1e-9; }
            }
        }
        uint32_t okb = __ballot_sync(FULL, ok);
        if (ok) {
            uint64_t dst = o + done + __popc(okb & ((1u << lane) - 1u)) + (c > (uint32_t)i ? 1u : 0u);
            indices[dst] = c; data[dst] = v;
        }
        done += __popc(okb);
    }
}

// ---- SpMV (graph.rs:464-501): y_r = sum over the row in CSR order, one thread per row so the
// fold order is the reference's ----------------------------------------------------------------
__global__ void spmv_rows_kernel(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                                 const double* __restrict__ data, uint64_t m, const double* __restrict__ x, double* __restrict__ y) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double s = 0.0;
    for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) s = __dadd_rn(s, __dmul_rn(data[e], x[indices[e]]));
    y[r] = s;
}
// num = sum x_r*y_r, den = sum x_r^2 with a fixed-shape tree (deterministic)
__global__ void dot2_kernel(const double* __restrict__ x, const double* __restrict__ y, uint64_t m, double* __restrict__ partial) {
    __shared__ double sn[256], sd[256];
    double n = 0.0, d = 0.0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (uint64_t)gridDim.x * blockDim.x) {
        n += x[r] * y[r]; d += x[r] * x[r];
    }
    sn[threadIdx.x] = n; sd[threadIdx.x] = d;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) { sn[threadIdx.x] += sn[threadIdx.x + o]; sd[threadIdx.x] += sd[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sn[0]; partial[2 * blockIdx.x + 1] = sd[0]; }
}

template <int MODE>
int32_t launch_adjacency(sfb_ctx* ctx, const uint32_t* in_idx, const double* in_val, const uint32_t* in_cnt,
                         const uint32_t* deg, uint64_t m, uint32_t k, double p, double sigma, int pmode, int sparsify,
                         double ratio, uint32_t* out_idx, double* out_w, uint32_t* out_cnt) {
    unsigned blocks = div_up(m * 32, 256);
#define SFB_ADJ_CASE(E)                                                                                               \
    adjacency_kernel<E, MODE><<<blocks, 256, 0, ctx->stream>>>(in_idx, in_val, in_cnt, deg, m, k, p, sigma, pmode,     \
                                                                sparsify, ratio, out_idx, out_w, out_cnt)
    if (k <= 32) SFB_ADJ_CASE(1);
    else if (k <= 64) SFB_ADJ_CASE(2);
    else SFB_ADJ_CASE(4);
#undef SFB_ADJ_CASE
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

}  // namespace

extern "C" int32_t sfb_adjacency_build(sfb_ctx* ctx, const sfb_knn* g, const sfb_adj_params* prm, sfb_adj** out,
                                       int32_t* sparsified) {
    if (!ctx || !g || !prm || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (g->rows != g->total || g->q_begin != 0)
        return sfb_fail(ctx, SFB_EINVAL, "adjacency needs the kNN lists of all %llu nodes (all-gather the shards first)", (unsigned long long)g->total);
    if (!(prm->sigma > 0.0) || !isfinite(prm->p)) return sfb_fail(ctx, SFB_EINVAL, "sigma must be > 0 and p finite");
    if (g->k > 128) return sfb_fail(ctx, SFB_EUNSUPPORTED, "k <= 128");
    SFB_TRY(sfb_adj_alloc(ctx, g->rows, g->k, out));
    sfb_adj* a = *out;
    StageTimer t(ctx, &ctx->times.ms_adjacency);
    int sparsify = prm->sparsify;
    if (sparsify < 0) {
        // mean degree > 10 (laplacian.rs:231-232), degree = #kNN entries with i != j && d <= eps
        DevBuf tot;
        SFB_CUDA(ctx, tot.alloc(8));
        SFB_CUDA(ctx, cudaMemsetAsync(tot.p, 0, 8, ctx->stream));
        sum_u32_kernel<<<2 * ctx->sm_count, 256, 0, ctx->stream>>>(g->cnt, g->rows, tot.as<unsigned long long>());
        SFB_LAUNCH_CHECK(ctx);
        unsigned long long h = 0;
        SFB_CUDA(ctx, cudaMemcpyAsync(&h, tot.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        sparsify = ((double)h / (double)g->rows) > 10.0 ? 1 : 0;
    }
    int pmode = prm->p == 1.0 ? 1 : (prm->p == 2.0 ? 2 : 0);
    int32_t st = launch_adjacency<0>(ctx, g->idx, g->dist, g->cnt, nullptr, g->rows, g->k, prm->p, prm->sigma, pmode, sparsify,
                                     0.0, a->idx, a->w, a->cnt);
    if (st == SFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "adjacency kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (st != SFB_OK) { sfb_adj_free(a); *out = nullptr; return st; }
    if (sparsified) *sparsified = sparsify;
    return SFB_OK;
}

extern "C" int32_t sfb_sparsify_sfgrass(sfb_ctx* ctx, sfb_adj* a, double ratio, int32_t* applied) {
    if (!ctx || !a) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (isnan(ratio)) return sfb_fail(ctx, SFB_EINVAL, "ratio is NaN");
    ratio = ratio < 0.1 ? 0.1 : (ratio > 1.0 ? 1.0 : ratio);  // sparsification.rs:26-29
    StageTimer t(ctx, &ctx->times.ms_adjacency);
    DevBuf tot, deg;
    SFB_CUDA(ctx, tot.alloc(8));
    SFB_CUDA(ctx, cudaMemsetAsync(tot.p, 0, 8, ctx->stream));
    sum_u32_kernel<<<2 * ctx->sm_count, 256, 0, ctx->stream>>>(a->cnt, a->rows, tot.as<unsigned long long>());
    SFB_LAUNCH_CHECK(ctx);
    unsigned long long h = 0;
    SFB_CUDA(ctx, cudaMemcpyAsync(&h, tot.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (((double)h / (double)a->rows) < 10.0) { if (applied) *applied = 0; return SFB_OK; }  // :42-52
    SFB_CUDA(ctx, deg.alloc(sizeof(uint32_t) * a->rows));
    SFB_CUDA(ctx, cudaMemcpyAsync(deg.p, a->cnt, sizeof(uint32_t) * a->rows, cudaMemcpyDeviceToDevice, ctx->stream));
    // rows are read into registers before being rewritten, neighbours' degrees come from the snapshot
    SFB_TRY(launch_adjacency<1>(ctx, a->idx, a->w, a->cnt, deg.as<uint32_t>(), a->rows, a->k, 0.0, 1.0, 0, 1, ratio, a->idx, a->w, a->cnt));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (applied) *applied = 1;
    return SFB_OK;
}

extern "C" int32_t sfb_laplacian_build(sfb_ctx* ctx, const sfb_adj* a, const sfb_lap_params* prm, sfb_csr** out) {
    if (!ctx || !a || !prm || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    *out = nullptr;
    const uint64_t m = a->rows; const uint32_t k = a->k;
    StageTimer timer(ctx, &ctx->times.ms_laplacian);
    constexpr uint32_t WARP_CAP = 256, BLOCK_CAP = 8192;

    DevBuf rev_cnt, fill, rev_off, len, tmp_off, max_len;
    SFB_CUDA(ctx, rev_cnt.alloc(sizeof(uint32_t) * m));
    SFB_CUDA(ctx, fill.alloc(sizeof(uint32_t) * m));
    SFB_CUDA(ctx, rev_off.alloc(sizeof(uint64_t) * (m + 1)));
    SFB_CUDA(ctx, len.alloc(sizeof(uint32_t) * m));
    SFB_CUDA(ctx, tmp_off.alloc(sizeof(uint64_t) * (m + 1)));
    SFB_CUDA(ctx, max_len.alloc(2 * sizeof(uint32_t)));
    SFB_CUDA(ctx, cudaMemsetAsync(rev_cnt.p, 0, sizeof(uint32_t) * m, ctx->stream));
    SFB_CUDA(ctx, cudaMemsetAsync(fill.p, 0, sizeof(uint32_t) * m, ctx->stream));
    SFB_CUDA(ctx, cudaMemsetAsync(max_len.p, 0, 2 * sizeof(uint32_t), ctx->stream));

    rev_count_kernel<<<div_up(m * k, 256), 256, 0, ctx->stream>>>(a->idx, a->cnt, m, k, rev_cnt.as<uint32_t>());
    SFB_LAUNCH_CHECK(ctx);
    SFB_TRY(sfb_scan_exclusive_u64(ctx, rev_cnt.as<uint32_t>(), m, rev_off.as<uint64_t>()));
    uint64_t n_rev = 0;
    SFB_CUDA(ctx, cudaMemcpyAsync(&n_rev, rev_off.as<uint64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

    DevBuf rev_src, rev_w;
    SFB_CUDA(ctx, rev_src.alloc(sizeof(uint32_t) * n_rev));
    SFB_CUDA(ctx, rev_w.alloc(sizeof(double) * n_rev));
    rev_scatter_kernel<<<div_up(m * k, 256), 256, 0, ctx->stream>>>(a->idx, a->w, a->cnt, m, k, rev_off.as<uint64_t>(),
                                                                    fill.as<uint32_t>(), rev_src.as<uint32_t>(), rev_w.as<double>());
    SFB_LAUNCH_CHECK(ctx);
    row_len_kernel<<<div_up(m, 256), 256, 0, ctx->stream>>>(a->cnt, rev_cnt.as<uint32_t>(), m, len.as<uint32_t>(), max_len.as<uint32_t>());
    SFB_LAUNCH_CHECK(ctx);
    SFB_TRY(sfb_scan_exclusive_u64(ctx, len.as<uint32_t>(), m, tmp_off.as<uint64_t>()));
    uint64_t n_tmp = 0; uint32_t h_max_len = 0;
    SFB_CUDA(ctx, cudaMemcpyAsync(&n_tmp, tmp_off.as<uint64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(&h_max_len, max_len.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_max_len > BLOCK_CAP)
        return sfb_fail(ctx, SFB_EUNSUPPORTED, "a node has %u incident directed edges; this build sorts rows of up to %u", h_max_len, BLOCK_CAP);

    DevBuf tmp_col, tmp_w, ulen, deg, row_nnz;
    SFB_CUDA(ctx, tmp_col.alloc(sizeof(uint32_t) * n_tmp));
    SFB_CUDA(ctx, tmp_w.alloc(sizeof(double) * n_tmp));
    SFB_CUDA(ctx, ulen.alloc(sizeof(uint32_t) * m));
    SFB_CUDA(ctx, deg.alloc(sizeof(double) * m));
    SFB_CUDA(ctx, row_nnz.alloc(sizeof(uint32_t) * m));

    {
        const int groups = 8;
        size_t smem = (size_t)groups * WARP_CAP * (sizeof(double) + sizeof(uint32_t));
        auto kern = lap_merge_rows_kernel<false, WARP_CAP>;
        SFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<div_up(m, groups), groups * 32, smem, ctx->stream>>>(a->idx, a->w, a->cnt, k, rev_off.as<uint64_t>(), rev_src.as<uint32_t>(),
                                                                    rev_w.as<double>(), len.as<uint32_t>(), tmp_off.as<uint64_t>(), m, nullptr, 0,
                                                                    WARP_CAP, tmp_col.as<uint32_t>(), tmp_w.as<double>(), ulen.as<uint32_t>(), deg.as<double>());
        SFB_LAUNCH_CHECK(ctx);
    }
    if (h_max_len > WARP_CAP) {
        DevBuf list;
        SFB_CUDA(ctx, list.alloc(sizeof(uint32_t) * m));
        uint32_t* n_list_d = max_len.as<uint32_t>() + 1;
        long_rows_kernel<<<div_up(m, 256), 256, 0, ctx->stream>>>(len.as<uint32_t>(), m, WARP_CAP, list.as<uint32_t>(), n_list_d);
        SFB_LAUNCH_CHECK(ctx);
        uint32_t n_list = 0;
        SFB_CUDA(ctx, cudaMemcpyAsync(&n_list, n_list_d, 4, cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        size_t smem = (size_t)BLOCK_CAP * (sizeof(double) + sizeof(uint32_t));
        auto kern = lap_merge_rows_kernel<true, BLOCK_CAP>;
        SFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n_list, 256, smem, ctx->stream>>>(a->idx, a->w, a->cnt, k, rev_off.as<uint64_t>(), rev_src.as<uint32_t>(), rev_w.as<double>(),
                                                 len.as<uint32_t>(), tmp_off.as<uint64_t>(), m, list.as<uint32_t>(), n_list, WARP_CAP,
                                                 tmp_col.as<uint32_t>(), tmp_w.as<double>(), ulen.as<uint32_t>(), deg.as<double>());
        SFB_LAUNCH_CHECK(ctx);
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    csr_row_nnz_kernel<<<div_up(m * 32, 256), 256, 0, ctx->stream>>>(ulen.as<uint32_t>(), deg.as<double>(), tmp_off.as<uint64_t>(),
                                                                     tmp_col.as<uint32_t>(), tmp_w.as<double>(), m, prm->normalised,
                                                                     prm->weight_threshold, row_nnz.as<uint32_t>());
    SFB_LAUNCH_CHECK(ctx);

    sfb_csr* L = new (std::nothrow) sfb_csr();
    if (!L) return SFB_ENOMEM;
    L->ctx = ctx; L->rows = m;
    L->symmetric = 1;   // union / max symmetrisation, L_ij and L_ji are the same expression of (w, d_i, d_j): symmetric bit for bit
    if (sfb_dev_alloc(ctx, (void**)&L->indptr, sizeof(uint64_t) * (m + 1)) != cudaSuccess) { sfb_csr_free(L); return sfb_fail(ctx, SFB_ENOMEM, "indptr"); }
    int32_t st = sfb_scan_exclusive_u64(ctx, row_nnz.as<uint32_t>(), m, L->indptr);
    if (st != SFB_OK) { sfb_csr_free(L); return st; }
    cudaMemcpyAsync(&L->nnz, L->indptr + m, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (sfb_dev_alloc(ctx, (void**)&L->indices, sizeof(uint32_t) * (L->nnz ? L->nnz : 1)) != cudaSuccess ||
        sfb_dev_alloc(ctx, (void**)&L->data, sizeof(double) * (L->nnz ? L->nnz : 1)) != cudaSuccess) {
        sfb_csr_free(L);
        return sfb_fail(ctx, SFB_ENOMEM, "CSR arrays (%llu nnz)", (unsigned long long)L->nnz);
    }
    csr_emit_kernel<<<div_up(m * 32, 256), 256, 0, ctx->stream>>>(ulen.as<uint32_t>(), deg.as<double>(), tmp_off.as<uint64_t>(),
                                                                  tmp_col.as<uint32_t>(), tmp_w.as<double>(), m, prm->normalised,
                                                                  prm->weight_threshold, L->indptr, L->indices, L->data);
    ctx->times.kernel_launches++;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { sfb_csr_free(L); return sfb_fail(ctx, SFB_ECUDA, "CSR emit: %s", cudaGetErrorString(e)); }
    *out = L;
    return SFB_OK;
}

extern "C" int32_t sfb_spmv(sfb_ctx* ctx, const sfb_csr* L, const double* x, double* y) {
    if (!ctx || !L || !x || !y) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    DevBuf dx, dy;
    SFB_CUDA(ctx, dx.alloc(sizeof(double) * L->rows));
    SFB_CUDA(ctx, dy.alloc(sizeof(double) * L->rows));
    SFB_CUDA(ctx, cudaMemcpyAsync(dx.p, x, sizeof(double) * L->rows, cudaMemcpyHostToDevice, ctx->stream));
    spmv_rows_kernel<<<div_up(L->rows, 128), 128, 0, ctx->stream>>>(L->indptr, L->indices, L->data, L->rows, dx.as<double>(), dy.as<double>());
    SFB_LAUNCH_CHECK(ctx);
    SFB_CUDA(ctx, cudaMemcpyAsync(y, dy.p, sizeof(double) * L->rows, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}

extern "C" int32_t sfb_rayleigh_quotient(sfb_ctx* ctx, const sfb_csr* L, const double* x, double* out) {
    if (!ctx || !L || !x || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    const int nb = 64;
    DevBuf dx, dy, part;
    SFB_CUDA(ctx, dx.alloc(sizeof(double) * L->rows));
    SFB_CUDA(ctx, dy.alloc(sizeof(double) * L->rows));
    SFB_CUDA(ctx, part.alloc(sizeof(double) * 2 * nb));
    SFB_CUDA(ctx, cudaMemcpyAsync(dx.p, x, sizeof(double) * L->rows, cudaMemcpyHostToDevice, ctx->stream));
    spmv_rows_kernel<<<div_up(L->rows, 128), 128, 0, ctx->stream>>>(L->indptr, L->indices, L->data, L->rows, dx.as<double>(), dy.as<double>());
    SFB_LAUNCH_CHECK(ctx);
    dot2_kernel<<<nb, 256, 0, ctx->stream>>>(dx.as<double>(), dy.as<double>(), L->rows, part.as<double>());
    SFB_LAUNCH_CHECK(ctx);
    double h[2 * nb];
    SFB_CUDA(ctx, cudaMemcpyAsync(h, part.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double num = 0.0, den = 0.0;
    for (int b = 0; b < nb; ++b) { num += h[2 * b]; den += h[2 * b + 1]; }
    *out = den > 1e-12 ? num / den : 0.0;  // graph.rs:447-453 (no max(0,.) here)
    return SFB_OK;
}
