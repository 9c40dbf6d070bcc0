// laplacian.cu -- kernel weights, sparsification, symmetrisation and CSR Laplacian assembly.
//
// Reference: src_legacy/laplacian.rs:231-290 (weights + inline sparsification), :297-348
// (symmetrise), :351-419 + :161 (L = D - W as CSR); src_legacy/sparsification.rs:32-113 (SF-GRASS);
// surfface-core/src/laplacian.rs:333-372,209-219 (normalised L_sym).
//
// The reference symmetrises with an O(M*E) DashMap scan and assembles through a sequential TriMat
// fill.  Here: reverse edges are bucketed by destination with a counting sort (histogram, one-pass
// look-back scan, scatter), then each row is handled by one warp: forward and reverse entries are sorted
// by column (in registers by a shuffle network for rows of up to 32 edges, in shared memory up to 256, by
// a block -- in global memory if need be -- for hub rows of any length), duplicates merged (max), the
// degree is a left fold in ascending column order (same order as the reference), and the CSR rows are
// staged per block and written with 16-byte stores.  Seven launches, one host round trip (the exact nnz).
// HBM-bound: reads M*k*12 B of lists, writes (M+1)*8 + nnz*12 B.
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
// Reverse edges are bucketed by destination with a counting sort: count (atomics), one-pass scan, scatter.  The scatter
// takes its slots by counting rev_cnt back DOWN (the order inside a bucket is arbitrary either way: the per-row sort
// fixes it), which leaves rev_cnt zeroed for the next call's scratch and needs no second counter array.
// [r_begin, r_end): the rows this call owns (all rows, or this rank's shard of a row-sharded build).
__global__ void rev_count_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ cnt, uint64_t m, uint32_t k,
                                 uint64_t r_begin, uint64_t r_end, uint32_t* __restrict__ rev_cnt) {
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * k) return;
    uint64_t i = gid / k; uint32_t t = (uint32_t)(gid % k);
    if (t >= cnt[i]) return;
    uint32_t j = idx[gid];
    if (j != (uint32_t)i && j >= r_begin && j < r_end) atomicAdd(&rev_cnt[j - r_begin], 1u);
}
__global__ void rev_scatter_kernel(const uint32_t* __restrict__ idx, const double* __restrict__ w, const uint32_t* __restrict__ cnt,
                                   uint64_t m, uint32_t k, uint64_t r_begin, uint64_t r_end, const uint64_t* __restrict__ rev_off,
                                   uint32_t* __restrict__ rev_cnt, uint32_t* __restrict__ rev_src, double* __restrict__ rev_w) {
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m * k) return;
    uint64_t i = gid / k; uint32_t t = (uint32_t)(gid % k);
    if (t >= cnt[i]) return;
    uint32_t j = idx[gid];
    if (j == (uint32_t)i || j < r_begin || j >= r_end) return;
    uint64_t o = rev_off[j - r_begin] + (atomicSub(&rev_cnt[j - r_begin], 1u) - 1u);
    rev_src[o] = (uint32_t)i; rev_w[o] = w[gid];
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

// Work descriptor shared by the row kernels.  Row i (global index) of the owned range has its forward list at
// a_idx[i * k ..], its reverse bucket at rev_off[i - r_begin], and its scratch segment for the merged, de-duplicated
// neighbours at tmp[(i - r_begin) * k + rev_off[i - r_begin] ..] (capacity k + bucket length: no second scan).
struct MergeArgs {
    const uint32_t* a_idx; const double* a_w; const uint32_t* a_cnt; uint32_t k;
    const uint64_t* rev_off; const uint32_t* rev_src; const double* rev_w;
    uint64_t r_begin, n_rows;
    uint32_t* tmp_col; double* tmp_w; uint32_t* ulen; double* deg; uint32_t* row_nnz;   // row_nnz: ulen + 1 (null when normalised)
    uint32_t* long_list; uint32_t* n_long;   // rows the warp kernel leaves to the block kernel
    // final != 0 (unnormalised form): the scratch segment receives the finished CSR row -- neighbours as -w with the diagonal
    // (the degree) inserted in column order, ulen + 1 entries -- and the emit pass is a plain segmented copy
    int final;
    int slots64;   // node indices below 2^26: rows of 33..64 edges are sorted in registers too (two entries per lane)
};
// scratch segment of local row li: capacity k + 1 + bucket length
__device__ __forceinline__ uint64_t tmp_offset(const MergeArgs& a, uint64_t li, uint64_t ro) { return li * (a.k + 1) + ro; }

// Pass A, rows of up to 32 incident edges (the common case: k forward + about as many reverse after sparsification):
// one warp per row, ONE entry per lane, sorted in registers by (column, slot) with a 15-step shuffle network; a column
// that comes from both directions keeps the larger weight (reference: identical for a symmetric metric, max otherwise).
// Rows of 33..256 edges are sorted in shared memory by the same warp; longer ones go to the block kernel's list.
template <uint32_t CAP, bool KEY32>
__global__ void __launch_bounds__(256) lap_merge_rows_kernel(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, grp = threadIdx.x >> 5, groups = blockDim.x >> 5;
    double* sw = reinterpret_cast<double*>(smem_raw) + (size_t)grp * CAP;
    uint32_t* scol = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)groups * CAP) + (size_t)grp * CAP;
    const uint64_t li = (uint64_t)blockIdx.x * groups + grp;
    if (li >= a.n_rows) return;
    const uint64_t i = a.r_begin + li;
    const uint32_t fc = a.a_cnt[i];
    const uint64_t ro = a.rev_off[li];
    const uint32_t rl = (uint32_t)(a.rev_off[li + 1] - ro);
    const uint32_t l = fc + rl;
    const uint64_t to = tmp_offset(a, li, ro);
    if (l == 0) {
        if (lane == 0) {
            a.ulen[li] = 0; a.deg[li] = 0.0;
            if (a.row_nnz) a.row_nnz[li] = 1;
            if (a.final) { a.tmp_col[to] = (uint32_t)i; a.tmp_w[to] = 0.0; }   // the diagonal is stored even when it is 0
        }
        return;
    }
    if (l > CAP) { if (lane == 0) a.long_list[atomicAdd(a.n_long, 1u)] = (uint32_t)li; return; }
    uint32_t u = 0, n_left = 0;   // unique neighbours; those left of the diagonal
    bool in_regs = l <= 32;
    if (in_regs) {
        uint32_t c = SFB_IDX_NONE; double w = 0.0;
        if (lane < fc) { c = a.a_idx[i * a.k + lane]; w = a.a_w[i * a.k + lane]; if (c == (uint32_t)i) c = SFB_IDX_NONE; }
        else if (lane < l) { c = a.rev_src[ro + (lane - fc)]; w = a.rev_w[ro + (lane - fc)]; }
        uint32_t col, slot;
        if (KEY32) {   // node indices below 2^27: (column, slot) fits one 32-bit key, one shuffle and one min / max per step
            uint32_t key = c == SFB_IDX_NONE ? 0xFFFFFFFFu : (c << 5 | lane);
#pragma unroll
            for (uint32_t size = 2; size <= 32; size <<= 1)
#pragma unroll
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    const uint32_t other = __shfl_xor_sync(FULL, key, stride);
                    const bool take_min = ((lane & size) == 0) == ((lane & stride) == 0);
                    key = take_min ? min(key, other) : max(key, other);
                }
            col = key == 0xFFFFFFFFu ? SFB_IDX_NONE : key >> 5; slot = key & 31u;
        } else {
            unsigned long long key = ((unsigned long long)c << 32) | lane;
#pragma unroll
            for (uint32_t size = 2; size <= 32; size <<= 1)
#pragma unroll
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(FULL, key, stride);
                    const bool take_min = ((lane & size) == 0) == ((lane & stride) == 0);
                    key = take_min ? (other < key ? other : key) : (other > key ? other : key);
                }
            col = (uint32_t)(key >> 32); slot = (uint32_t)key & 31u;
        }
        const double ws = __shfl_sync(FULL, w, (int)slot);
        const uint32_t col_prev = __shfl_up_sync(FULL, col, 1), col_prev2 = __shfl_up_sync(FULL, col, 2), col_next = __shfl_down_sync(FULL, col, 1);
        const double w_next = __shfl_down_sync(FULL, ws, 1);
        const bool valid = col != SFB_IDX_NONE;
        // a column normally occurs at most twice (once forward, once reverse); lists from the host may repeat one more often
        if (__any_sync(FULL, valid && lane >= 2 && col == col_prev && col == col_prev2)) in_regs = false;
        else {
            const bool head = valid && (lane == 0 || col != col_prev);
            const double wh = (lane < 31 && col_next == col && w_next > ws) ? w_next : ws;   // duplicates -> max
            const uint32_t hb = __ballot_sync(FULL, head);
            const uint32_t pos = __popc(hb & ((1u << lane) - 1u));
            u = __popc(hb);
            n_left = __popc(__ballot_sync(FULL, head && col < (uint32_t)i));
            if (head) {
                sw[pos] = wh;
                if (a.final) { const uint32_t d = pos + (col > (uint32_t)i ? 1u : 0u); a.tmp_col[to + d] = col; a.tmp_w[to + d] = -wh; }
                else { a.tmp_col[to + pos] = col; a.tmp_w[to + pos] = wh; }
            }
            __syncwarp();
        }
    }
    if (KEY32 && !in_regs && l <= 64 && a.slots64) {
        // 33..64 edges (k = 64 graphs after sparsification): two entries per lane, element e = lane (+ 32), one 64-key bitonic
        // network: stride 32 is a compare-exchange between the lane's own two keys, smaller strides are shuffles
        uint32_t key[2]; double w[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t t = lane + 32 * h;
            uint32_t c = SFB_IDX_NONE; w[h] = 0.0;
            if (t < fc) { c = a.a_idx[i * a.k + t]; w[h] = a.a_w[i * a.k + t]; if (c == (uint32_t)i) c = SFB_IDX_NONE; }
            else if (t < l) { c = a.rev_src[ro + (t - fc)]; w[h] = a.rev_w[ro + (t - fc)]; }
            key[h] = c == SFB_IDX_NONE ? 0xFFFFFFFFu : (c << 6 | t);
        }
#pragma unroll
        for (uint32_t size = 2; size <= 64; size <<= 1)
#pragma unroll
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                if (stride == 32) {   // partner of element lane is element lane + 32; size = 64: ascending
                    const uint32_t lo = min(key[0], key[1]), hi = max(key[0], key[1]);
                    key[0] = lo; key[1] = hi;
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t e = lane + 32 * h;
                        const uint32_t other = __shfl_xor_sync(FULL, key[h], stride);
                        const bool take_min = ((e & size) == 0) == ((e & stride) == 0);
                        key[h] = take_min ? min(key[h], other) : max(key[h], other);
                    }
                }
            }
        uint32_t col[2]; double ws[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            col[h] = key[h] == 0xFFFFFFFFu ? SFB_IDX_NONE : key[h] >> 6;
            const uint32_t slot = key[h] & 63u;
            const double wa = __shfl_sync(FULL, w[0], (int)(slot & 31u)), wb = __shfl_sync(FULL, w[1], (int)(slot & 31u));
            ws[h] = slot < 32u ? wa : wb;
        }
        // neighbours in sorted order: element e - 1 / e - 2 / e + 1 of (lane, h)
        const uint32_t c0_up1 = __shfl_up_sync(FULL, col[0], 1), c1_up1 = __shfl_up_sync(FULL, col[1], 1);
        const uint32_t c0_up2 = __shfl_up_sync(FULL, col[0], 2), c1_up2 = __shfl_up_sync(FULL, col[1], 2);
        const uint32_t c0_last = __shfl_sync(FULL, col[0], 31), c0_last2 = __shfl_sync(FULL, col[0], 30);
        const uint32_t c0_dn = __shfl_down_sync(FULL, col[0], 1), c1_dn = __shfl_down_sync(FULL, col[1], 1), c1_first = __shfl_sync(FULL, col[1], 0);
        const double w0_dn = __shfl_down_sync(FULL, ws[0], 1), w1_dn = __shfl_down_sync(FULL, ws[1], 1), w1_first = __shfl_sync(FULL, ws[1], 0);
        const uint32_t prev[2] = {lane ? c0_up1 : SFB_IDX_NONE, lane ? c1_up1 : c0_last};
        const uint32_t prev2[2] = {lane >= 2 ? c0_up2 : SFB_IDX_NONE, lane >= 2 ? c1_up2 : (lane == 1 ? c0_last : c0_last2)};
        const uint32_t next[2] = {lane < 31 ? c0_dn : c1_first, lane < 31 ? c1_dn : SFB_IDX_NONE};
        const double wnext[2] = {lane < 31 ? w0_dn : w1_first, w1_dn};
        bool triple = false, head[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool valid = col[h] != SFB_IDX_NONE;
            const bool has_prev = h == 1 || lane > 0, has_prev2 = h == 1 || lane > 1;
            triple = triple || (valid && has_prev2 && col[h] == prev[h] && col[h] == prev2[h]);
            head[h] = valid && (!has_prev || col[h] != prev[h]);
        }
        if (!__any_sync(FULL, triple)) {
            const uint32_t hb0 = __ballot_sync(FULL, head[0]), hb1 = __ballot_sync(FULL, head[1]);
            const uint32_t lt = (1u << lane) - 1u;
            const uint32_t pos[2] = {(uint32_t)__popc(hb0 & lt), (uint32_t)(__popc(hb0) + __popc(hb1 & lt))};
            u = __popc(hb0) + __popc(hb1);
            n_left = __popc(__ballot_sync(FULL, head[0] && col[0] < (uint32_t)i)) + __popc(__ballot_sync(FULL, head[1] && col[1] < (uint32_t)i));
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (head[h]) {
                    const bool dup = next[h] == col[h] && !(h == 1 && lane == 31);
                    const double wh = (dup && wnext[h] > ws[h]) ? wnext[h] : ws[h];   // duplicates -> max
                    sw[pos[h]] = wh;
                    if (a.final) { const uint32_t d = pos[h] + (col[h] > (uint32_t)i ? 1u : 0u); a.tmp_col[to + d] = col[h]; a.tmp_w[to + d] = -wh; }
                    else { a.tmp_col[to + pos[h]] = col[h]; a.tmp_w[to + pos[h]] = wh; }
                }
            __syncwarp();
            in_regs = true;
        }
    }
    if (!in_regs) {
        uint32_t P = 2; while (P < l) P <<= 1;
        for (uint32_t t = lane; t < P; t += 32) {
            uint32_t c = SFB_IDX_NONE; double w = 0.0;
            if (t < fc) { c = a.a_idx[i * a.k + t]; w = a.a_w[i * a.k + t]; if (c == (uint32_t)i) c = SFB_IDX_NONE; }
            else if (t < l) { c = a.rev_src[ro + (t - fc)]; w = a.rev_w[ro + (t - fc)]; }
            scol[t] = c; sw[t] = w;
        }
        bitonic_sort_pairs<false>(scol, sw, P, lane, 32);
        // heads of runs of equal column; sorted (col asc, w desc) => the head carries the max weight
        uint32_t base = 0;
        for (uint32_t t0 = 0; t0 < P; t0 += 32) {
            const uint32_t t = t0 + lane;
            uint32_t c = SFB_IDX_NONE; double w = 0.0; bool head = false;
            if (t < P) { c = scol[t]; w = sw[t]; head = c != SFB_IDX_NONE && (t == 0 || c != scol[t - 1]); }
            const uint32_t b = __ballot_sync(FULL, head);
            const uint32_t pos = base + __popc(b & ((1u << lane) - 1u));
            n_left += __popc(__ballot_sync(FULL, head && c < (uint32_t)i));
            __syncwarp();
            if (head) {
                sw[pos] = w;   // compacted in place for the fold below (pos <= t)
                if (a.final) { const uint32_t d = pos + (c > (uint32_t)i ? 1u : 0u); a.tmp_col[to + d] = c; a.tmp_w[to + d] = -w; }
                else { a.tmp_col[to + pos] = c; a.tmp_w[to + pos] = w; }
            }
            __syncwarp();
            base += __popc(b);
        }
        u = base;
    }
    // degree: left fold over the unique neighbours in ascending column order (laplacian.rs:371)
    if (lane == 0) {
        double s = 0.0;
        for (uint32_t t = 0; t < u; ++t) s = __dadd_rn(s, sw[t]);
        a.ulen[li] = u; a.deg[li] = s;
        if (a.row_nnz) a.row_nnz[li] = u + 1;
        if (a.final) { a.tmp_col[to + n_left] = (uint32_t)i; a.tmp_w[to + n_left] = s; }
    }
}

// Pass A for the long rows (hubs: in-degree is unbounded).  One block per listed row, any length: the entries are
// gathered into the row's scratch segment in global memory, sorted there by an ascending-only bitonic network (first
// step of every merge mirrored, so a padding element -- virtual index >= len, key +inf -- never has to move below a real
// one and the network works on lengths that are not powers of two), then de-duplicated and folded in place.  Rows
// that fit shared memory (<= SMEM_CAP) are sorted there instead.
template <uint32_t SMEM_CAP>
__global__ void __launch_bounds__(1024) lap_merge_long_kernel(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sw = reinterpret_cast<double*>(smem_raw);
    uint32_t* scol = reinterpret_cast<uint32_t*>(sw + SMEM_CAP);
    __shared__ uint32_t s_count, s_left, s_warp[32], s_wleft[32];
    const uint32_t tid = threadIdx.x, nt = blockDim.x, n_long = *a.n_long;
    for (uint32_t item = blockIdx.x; item < n_long; item += gridDim.x) {
        const uint64_t li = a.long_list[item], i = a.r_begin + li;
        const uint32_t fc = a.a_cnt[i];
        const uint64_t ro = a.rev_off[li];
        const uint32_t rl = (uint32_t)(a.rev_off[li + 1] - ro), l = fc + rl;
        const uint64_t to = tmp_offset(a, li, ro);
        uint32_t* gcol = a.tmp_col + to; double* gw = a.tmp_w + to;
        const bool in_smem = l <= SMEM_CAP;
        // global workspace: the row's own segment, shifted up by one slot (capacity k + 1 + bucket >= l + 1), so that the
        // compaction below -- which may move an entry one slot to the right of its rank to make room for the diagonal --
        // never writes a slot it has not read yet
        uint32_t* col = in_smem ? scol : gcol + 1; double* wv = in_smem ? sw : gw + 1;
        uint32_t P = 2; while (P < l) P <<= 1;
        for (uint32_t t = tid; t < (in_smem ? P : l); t += nt) {
            uint32_t c = SFB_IDX_NONE; double w = 0.0;
            if (t < fc) { c = a.a_idx[i * a.k + t]; w = a.a_w[i * a.k + t]; if (c == (uint32_t)i) c = SFB_IDX_NONE; }
            else if (t < l) { c = a.rev_src[ro + (t - fc)]; w = a.rev_w[ro + (t - fc)]; }
            col[t] = c; wv[t] = w;
        }
        __syncthreads();
        // ascending-only bitonic network over P virtual slots; slots >= len hold (+inf) and are never touched
        const uint32_t len = in_smem ? P : l;
        for (uint32_t size = 2; size <= P; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t t = tid; t < P / 2; t += nt) {
                    uint32_t lo, hi;
                    if (stride == size >> 1) { const uint32_t blk = t / stride, off = t % stride; lo = blk * size + off; hi = blk * size + size - 1 - off; }   // mirrored first step
                    else { lo = 2 * t - (t & (stride - 1)); hi = lo + stride; }
                    if (hi < len) {
                        const uint32_t ca = col[lo], cb = col[hi];
                        const double wa = wv[lo], wb = wv[hi];
                        if (ca > cb || (ca == cb && wa < wb)) { col[lo] = cb; col[hi] = ca; wv[lo] = wb; wv[hi] = wa; }
                    }
                }
                __syncthreads();
            }
        }
        // heads of runs (col asc, w desc: the head carries the max weight), compacted in order, chunk by chunk
        if (tid == 0) { s_count = 0; s_left = 0; }
        __syncthreads();
        for (uint32_t t0 = 0; t0 < l; t0 += nt) {
            const uint32_t t = t0 + tid;
            uint32_t c = SFB_IDX_NONE; double w = 0.0; bool head = false;
            if (t < l) { c = col[t]; w = wv[t]; head = c != SFB_IDX_NONE && (t == 0 || c != col[t - 1]); }
            const uint32_t b = __ballot_sync(FULL, head), bl = __ballot_sync(FULL, head && c < (uint32_t)i);
            if ((tid & 31) == 0) { s_warp[tid >> 5] = __popc(b); s_wleft[tid >> 5] = __popc(bl); }
            __syncthreads();   // every read of this chunk is done: the compacted writes below land at or before it
            uint32_t pre = 0, tot = 0, totl = 0;
            for (uint32_t wq = 0; wq < (nt >> 5); ++wq) { const uint32_t v = s_warp[wq]; if (wq < (tid >> 5)) pre += v; tot += v; totl += s_wleft[wq]; }
            const uint32_t pos = s_count + pre + __popc(b & ((1u << (tid & 31)) - 1u));
            if (head) {
                if (a.final) { const uint32_t d = pos + (c > (uint32_t)i ? 1u : 0u); gcol[d] = c; gw[d] = -w; }
                else { gcol[pos] = c; gw[pos] = w; }
            }
            __syncthreads();
            if (tid == 0) { s_count += tot; s_left += totl; }
            __syncthreads();
        }
        // degree: left fold in ascending column order, one thread (a hub's fold is as sequential as the reference's)
        if (tid == 0) {
            const uint32_t u = s_count, nl = s_left;
            double s = 0.0;
            if (a.final) { for (uint32_t t = 0; t < u; ++t) s = __dadd_rn(s, -gw[t + (t >= nl ? 1u : 0u)]); gcol[nl] = (uint32_t)i; gw[nl] = s; }
            else for (uint32_t t = 0; t < u; ++t) s = __dadd_rn(s, gw[t]);
            a.ulen[li] = u; a.deg[li] = s;
            if (a.row_nnz) a.row_nnz[li] = u + 1;
        }
        __syncthreads();
    }
}

// Pass B (normalised form only): CSR row lengths after the |v| <= 1e-9 drop rule.
__global__ void csr_row_nnz_kernel(const uint32_t* __restrict__ ulen, const double* __restrict__ deg,
                                   const uint64_t* __restrict__ rev_off, uint32_t k, const uint32_t* __restrict__ tmp_col,
                                   const double* __restrict__ tmp_w, uint64_t m, double thr, uint32_t* __restrict__ row_nnz) {
    const int lane = threadIdx.x & 31;
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= m) return;
    const uint32_t u = ulen[i];
    const double di = deg[i];
    uint32_t n = 0;
    if (di > thr) {
        const uint64_t to = i * (k + 1) + rev_off[i];
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

// Pass C: emit CSR rows.  A block takes a run of consecutive rows, assembles their entries (diagonal inserted in column
// order) in shared memory and writes the run's contiguous slice of `indices` / `data` with 16-byte stores where the
// slice is aligned, scalar stores at its ragged ends; runs whose slice does not fit are written row by row.
// r_begin: global index of local row 0 (the diagonal's column).
constexpr uint32_t EMIT_ROWS = 64, EMIT_CAP = 4032;
__global__ void __launch_bounds__(256) csr_emit_kernel(const uint32_t* __restrict__ ulen, const double* __restrict__ deg,
                                                       const uint64_t* __restrict__ rev_off, uint32_t k, const uint32_t* __restrict__ tmp_col,
                                                       const double* __restrict__ tmp_w, uint64_t m, uint64_t r_begin, int normalised, double thr,
                                                       const uint64_t* __restrict__ indptr, uint32_t* __restrict__ indices, double* __restrict__ data) {
    __shared__ __align__(16) double s_val[EMIT_CAP];
    __shared__ __align__(16) uint32_t s_col[EMIT_CAP];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint64_t row0 = (uint64_t)blockIdx.x * EMIT_ROWS;
    if (row0 >= m) return;
    const uint64_t row1 = row0 + EMIT_ROWS < m ? row0 + EMIT_ROWS : m;
    const uint64_t o0 = indptr[row0], o1 = indptr[row1];
    const bool staged = o1 - o0 <= EMIT_CAP;
    for (uint64_t i = row0 + wid; i < row1; i += nw) {
        const uint32_t u = ulen[i];
        const double di = deg[i];
        if (normalised && !(di > thr)) continue;  // isolated node: empty row (surfface-core laplacian.rs:349-353)
        const uint64_t to = i * (k + 1) + rev_off[i];
        const uint64_t o = indptr[i];
        const uint32_t gi = (uint32_t)(r_begin + i);
        const double diag_val = normalised ? 1.0 : di;
        uint32_t* oc = staged ? s_col + (o - o0) : indices + o;
        double* ov = staged ? s_val + (o - o0) : data + o;
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
                left = ok && c < gi;
            }
            n_left += __popc(__ballot_sync(FULL, left));
        }
        if (lane == 0) { oc[n_left] = gi; ov[n_left] = diag_val; }
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
                uint32_t dst = done + __popc(okb & ((1u << lane) - 1u)) + (c > gi ? 1u : 0u);
                oc[dst] = c; ov[dst] = v;
            }
            done += __popc(okb);
        }
    }
    if (!staged) return;
    __syncthreads();
    // the run's slice [o0, o1): scalar head up to the next 16-byte boundary, vector body, scalar tail
    const uint32_t n = (uint32_t)(o1 - o0);
    {   // data (f64): 2 per vector
        const uint32_t head = (uint32_t)((2 - (o0 & 1)) & 1) < n ? (uint32_t)((2 - (o0 & 1)) & 1) : n;
        const uint32_t body = (n - head) / 2;
        if (threadIdx.x < head) data[o0 + threadIdx.x] = s_val[threadIdx.x];
        double2* dst = reinterpret_cast<double2*>(data + o0 + head);
        for (uint32_t v = threadIdx.x; v < body; v += blockDim.x) dst[v] = make_double2(s_val[head + 2 * v], s_val[head + 2 * v + 1]);
        for (uint32_t t = head + 2 * body + threadIdx.x; t < n; t += blockDim.x) data[o0 + t] = s_val[t];
    }
    {   // indices (u32): 4 per vector
        const uint32_t mis = (uint32_t)(o0 & 3), h0 = (4 - mis) & 3;
        const uint32_t head = h0 < n ? h0 : n;
        const uint32_t body = (n - head) / 4;
        if (threadIdx.x < head) indices[o0 + threadIdx.x] = s_col[threadIdx.x];
        uint4* dst = reinterpret_cast<uint4*>(indices + o0 + head);
        for (uint32_t v = threadIdx.x; v < body; v += blockDim.x)
            dst[v] = make_uint4(s_col[head + 4 * v], s_col[head + 4 * v + 1], s_col[head + 4 * v + 2], s_col[head + 4 * v + 3]);
        for (uint32_t t = head + 4 * body + threadIdx.x; t < n; t += blockDim.x) indices[o0 + t] = s_col[t];
    }
}

// Pass C, unnormalised form: the scratch segments already hold the finished rows (lap_merge_*: final != 0); a block
// takes a run of consecutive rows, whose output is one contiguous slice of `indices` / `data`, and copies it with one
// thread per PAIR of entries: 16-byte stores of the values, 8-byte stores of the columns (scalar at the ragged ends).
constexpr uint32_t COPY_ROWS = 128;
__global__ void __launch_bounds__(256) csr_copy_kernel(const uint64_t* __restrict__ rev_off, uint32_t k, const uint32_t* __restrict__ tmp_col,
                                                       const double* __restrict__ tmp_w, uint64_t m, const uint64_t* __restrict__ indptr,
                                                       uint32_t* __restrict__ indices, double* __restrict__ data) {
    __shared__ uint64_t s_ptr[COPY_ROWS + 1], s_src[COPY_ROWS];
    const uint64_t row0 = (uint64_t)blockIdx.x * COPY_ROWS;
    if (row0 >= m) return;
    const uint32_t nrows = (uint32_t)(row0 + COPY_ROWS < m ? COPY_ROWS : m - row0);
    for (uint32_t r = threadIdx.x; r <= nrows; r += blockDim.x) {
        s_ptr[r] = indptr[row0 + r];
        if (r < nrows) s_src[r] = (row0 + r) * (k + 1) + rev_off[row0 + r];
    }
    __syncthreads();
    const uint64_t o0 = s_ptr[0], o1 = s_ptr[nrows];
    auto locate = [&](uint64_t e) {   // source position of output entry e: its row by bisection over the run's indptr
        uint32_t lo = 0, hi = nrows;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_ptr[mid] <= e) lo = mid; else hi = mid; }
        return s_src[lo] + (e - s_ptr[lo]);
    };
    for (uint64_t e = (o0 & ~1ull) + 2ull * threadIdx.x; e < o1; e += 2ull * blockDim.x) {
        const bool v0 = e >= o0, v1 = e + 1 < o1;
        uint32_t c0 = 0, c1 = 0; double w0 = 0.0, w1 = 0.0;
        if (v0) { const uint64_t sp = locate(e); c0 = tmp_col[sp]; w0 = tmp_w[sp]; }
        if (v1) { const uint64_t sp = locate(e + 1); c1 = tmp_col[sp]; w1 = tmp_w[sp]; }
        if (v0 && v1) {
            *reinterpret_cast<double2*>(data + e) = make_double2(w0, w1);
            *reinterpret_cast<uint2*>(indices + e) = make_uint2(c0, c1);
        } else if (v0) { data[e] = w0; indices[e] = c0; }
        else if (v1) { data[e + 1] = w1; indices[e + 1] = c1; }
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

// rows [r_begin, r_end) of the Laplacian of the graph in `a` (all of its lists are needed: the reverse edges of a row
// live in other rows' lists).  One device pass, one host round trip (the exact nnz, to size the CSR arrays).
static int32_t laplacian_build_rows(sfb_ctx* ctx, const sfb_adj* a, const sfb_lap_params* prm, uint64_t r_begin, uint64_t r_end, sfb_csr** out) {
    *out = nullptr;
    const uint64_t m = a->rows, nr = r_end - r_begin; const uint32_t k = a->k;
    StageTimer timer(ctx, &ctx->times.ms_laplacian);
    constexpr uint32_t WARP_CAP = 256, BLOCK_SMEM_CAP = 8192;

    // buffers sized from the N * k bound: the build never waits for a count
    DevBuf rev_cnt, rev_off, rev_src, rev_w, tmp_col, tmp_w, ulen, deg, row_nnz, long_list, n_long;
    const uint64_t rev_cap = m * k;   // every directed edge lands in at most one owned bucket
    SFB_CUDA(ctx, rev_cnt.alloc(sizeof(uint32_t) * nr));
    SFB_CUDA(ctx, rev_off.alloc(sizeof(uint64_t) * (nr + 1)));
    SFB_CUDA(ctx, rev_src.alloc(sizeof(uint32_t) * rev_cap));
    SFB_CUDA(ctx, rev_w.alloc(sizeof(double) * rev_cap));
    SFB_CUDA(ctx, tmp_col.alloc(sizeof(uint32_t) * (nr * (k + 1) + rev_cap)));
    SFB_CUDA(ctx, tmp_w.alloc(sizeof(double) * (nr * (k + 1) + rev_cap)));
    SFB_CUDA(ctx, ulen.alloc(sizeof(uint32_t) * nr));
    SFB_CUDA(ctx, deg.alloc(sizeof(double) * (prm->normalised ? m : nr)));
    SFB_CUDA(ctx, row_nnz.alloc(sizeof(uint32_t) * nr));
    SFB_CUDA(ctx, long_list.alloc(sizeof(uint32_t) * nr));
    SFB_CUDA(ctx, n_long.alloc(sizeof(uint32_t)));
    SFB_CUDA(ctx, cudaMemsetAsync(rev_cnt.p, 0, sizeof(uint32_t) * nr, ctx->stream));
    SFB_CUDA(ctx, cudaMemsetAsync(n_long.p, 0, sizeof(uint32_t), ctx->stream));

    rev_count_kernel<<<div_up(m * k, 256), 256, 0, ctx->stream>>>(a->idx, a->cnt, m, k, r_begin, r_end, rev_cnt.as<uint32_t>());
    SFB_LAUNCH_CHECK(ctx);
    SFB_TRY(sfb_scan_exclusive_u64(ctx, rev_cnt.as<uint32_t>(), nr, rev_off.as<uint64_t>()));
    rev_scatter_kernel<<<div_up(m * k, 256), 256, 0, ctx->stream>>>(a->idx, a->w, a->cnt, m, k, r_begin, r_end, rev_off.as<uint64_t>(),
                                                                    rev_cnt.as<uint32_t>(), rev_src.as<uint32_t>(), rev_w.as<double>());
    SFB_LAUNCH_CHECK(ctx);

    MergeArgs ma{a->idx, a->w, a->cnt, k, rev_off.as<uint64_t>(), rev_src.as<uint32_t>(), rev_w.as<double>(), r_begin, nr,
                 tmp_col.as<uint32_t>(), tmp_w.as<double>(), ulen.as<uint32_t>(), deg.as<double>(),
                 prm->normalised ? nullptr : row_nnz.as<uint32_t>(), long_list.as<uint32_t>(), n_long.as<uint32_t>(), prm->normalised ? 0 : 1, m < (1ull << 26) ? 1 : 0};
    {
        const int groups = 8;
        const size_t smem = (size_t)groups * WARP_CAP * (sizeof(double) + sizeof(uint32_t));
        if (m < (1ull << 27)) {
            auto kern = lap_merge_rows_kernel<WARP_CAP, true>;
            SFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<div_up(nr, groups), groups * 32, smem, ctx->stream>>>(ma);
        } else {
            auto kern = lap_merge_rows_kernel<WARP_CAP, false>;
            SFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<div_up(nr, groups), groups * 32, smem, ctx->stream>>>(ma);
        }
        SFB_LAUNCH_CHECK(ctx);
    }
    {   // hub rows (in-degree is unbounded): the kernel reads the list length on the device and leaves at once when it is empty
        const size_t smem = (size_t)BLOCK_SMEM_CAP * (sizeof(double) + sizeof(uint32_t));
        auto kern = lap_merge_long_kernel<BLOCK_SMEM_CAP>;
        SFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<2 * ctx->sm_count, 1024, smem, ctx->stream>>>(ma);
        SFB_LAUNCH_CHECK(ctx);
    }
    if (prm->normalised) {   // needs the degree of every neighbour: full range only (checked by the caller)
        csr_row_nnz_kernel<<<div_up(nr * 32, 256), 256, 0, ctx->stream>>>(ulen.as<uint32_t>(), deg.as<double>(), rev_off.as<uint64_t>(), k,
                                                                         tmp_col.as<uint32_t>(), tmp_w.as<double>(), nr, prm->weight_threshold, row_nnz.as<uint32_t>());
        SFB_LAUNCH_CHECK(ctx);
    }

    sfb_csr* L = new (std::nothrow) sfb_csr();
    if (!L) return SFB_ENOMEM;
    L->ctx = ctx; L->rows = nr; L->row0 = r_begin; L->cols = m;
    L->symmetric = nr == m ? 1 : 0;   // union / max symmetrisation, L_ij and L_ji are the same expression of (w, d_i, d_j): symmetric bit for bit
    if (sfb_dev_alloc(ctx, (void**)&L->indptr, sizeof(uint64_t) * (nr + 1)) != cudaSuccess) { sfb_csr_free(L); return sfb_fail(ctx, SFB_ENOMEM, "indptr"); }
    int32_t st = sfb_scan_exclusive_u64(ctx, row_nnz.as<uint32_t>(), nr, L->indptr);
    if (st != SFB_OK) { sfb_csr_free(L); return st; }
    cudaError_t e = cudaMemcpyAsync(&L->nnz, L->indptr + nr, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // the one host round trip: the exact nnz sizes the CSR arrays
    if (e != cudaSuccess) { sfb_csr_free(L); return sfb_fail(ctx, SFB_ECUDA, "Laplacian assembly: %s", cudaGetErrorString(e)); }
    if (sfb_dev_alloc(ctx, (void**)&L->indices, sizeof(uint32_t) * (L->nnz ? L->nnz : 1)) != cudaSuccess ||
        sfb_dev_alloc(ctx, (void**)&L->data, sizeof(double) * (L->nnz ? L->nnz : 1)) != cudaSuccess) {
        sfb_csr_free(L);
        return sfb_fail(ctx, SFB_ENOMEM, "CSR arrays (%llu nnz)", (unsigned long long)L->nnz);
    }
    if (prm->normalised)
        csr_emit_kernel<<<div_up(nr, EMIT_ROWS), 256, 0, ctx->stream>>>(ulen.as<uint32_t>(), deg.as<double>(), rev_off.as<uint64_t>(), k,
                                                                         tmp_col.as<uint32_t>(), tmp_w.as<double>(), nr, r_begin, prm->normalised,
                                                                         prm->weight_threshold, L->indptr, L->indices, L->data);
    else
        csr_copy_kernel<<<div_up(nr, COPY_ROWS), 256, 0, ctx->stream>>>(rev_off.as<uint64_t>(), k, tmp_col.as<uint32_t>(), tmp_w.as<double>(), nr,
                                                                         L->indptr, L->indices, L->data);
    ctx->times.kernel_launches++;
    timer.stop();   // synchronises: the scratch may go back to the cache
    e = cudaGetLastError();
    if (e != cudaSuccess) { sfb_csr_free(L); return sfb_fail(ctx, SFB_ECUDA, "CSR emit: %s", cudaGetErrorString(e)); }
    *out = L;
    return SFB_OK;
}

extern "C" int32_t sfb_laplacian_build(sfb_ctx* ctx, const sfb_adj* a, const sfb_lap_params* prm, sfb_csr** out) {
    if (!ctx || !a || !prm || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    return laplacian_build_rows(ctx, a, prm, 0, a->rows, out);
}

extern "C" int32_t sfb_laplacian_build_rows(sfb_ctx* ctx, const sfb_adj* a, const sfb_lap_params* prm, uint64_t row_begin, uint64_t row_end,
                                            sfb_csr** out) {
    if (!ctx || !a || !prm || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (row_begin >= row_end || row_end > a->rows) return sfb_fail(ctx, SFB_EINVAL, "bad row range [%llu, %llu) of %llu", (unsigned long long)row_begin, (unsigned long long)row_end, (unsigned long long)a->rows);
    if (prm->normalised && (row_begin != 0 || row_end != a->rows))
        return sfb_fail(ctx, SFB_EUNSUPPORTED, "the normalised form needs the degree of every neighbour: build all rows");
    return laplacian_build_rows(ctx, a, prm, row_begin, row_end, out);
}

extern "C" int32_t sfb_spmv(sfb_ctx* ctx, const sfb_csr* L, const double* x, double* y) {
    if (!ctx || !L || !x || !y) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (L->cols && L->cols != L->rows) return sfb_fail(ctx, SFB_EINVAL, "a row shard of a Laplacian is not a square operator");
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
    if (L->cols && L->cols != L->rows) return sfb_fail(ctx, SFB_EINVAL, "a row shard of a Laplacian is not a square operator");
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
