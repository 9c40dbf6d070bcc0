// project.cu -- the steps on either side of the lambda kernel:
//
//   * JL projection of the items (ImplicitProjection::project / project_matrix, src_legacy/reduction.rs:175-242;
//     successor f32 form surfface-core/src/clustering.rs:84-109): Y = X S / sqrt(r) with every output a LEFT FOLD
//     over the original dimension, bit for bit the reference's loop.  FP64-pipe bound (3 instructions per term,
//     no FMA: the reference rounds the product, the scaling and the sum separately).
//   * SortedLambdas::build_from (src_legacy/sorted_index.rs:22-46): lambdas ascending, equal lambdas ordered by
//     the decimal string of the item index, plus the f32 standard deviation of laplacian.rs:421-448.
//     The string order of the indices is a closed-form permutation; a stable LSD radix sort of that index array by
//     the lambda key (digits computed on the fly from lambda[idx]) goes on top of it.
#include "common.cuh"

// ---- projection ------------------------------------------------------------------------------------------------
// Block tile: (TY*4) rows x (TX*4) reduced columns, 256 threads, 4 x 4 outputs per thread (16 independent chains
// hide the DADD latency).  The original dimension is walked in chunks of PK terms staged in shared memory.
constexpr int PK = 16;

template <typename T, int TX, bool CORE>
__global__ void __launch_bounds__(256) project_rows_kernel(const double* __restrict__ x, uint64_t n, uint32_t f, const T* __restrict__ s /* f x r */, uint32_t r,
                                                           T scale, double* __restrict__ out) {
    constexpr int TY = 256 / TX, BR = TY * 4, BC = TX * 4;
    __shared__ T xs[PK][BR + 4];
    __shared__ T ss[PK][BC];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const uint32_t ncb = (r + BC - 1) / BC;                 // column tiles of one row tile are neighbours in launch order (x tile shared in L2)
    const uint64_t row0 = (uint64_t)(blockIdx.x / ncb) * BR;
    const uint32_t col0 = (blockIdx.x % ncb) * BC;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = T(0);
    for (uint32_t i0 = 0; i0 < f; i0 += PK) {
        // x tile: BR rows x PK terms, PK consecutive doubles of a row per 16 threads (128-byte segments)
        for (int e = threadIdx.x; e < BR * PK; e += 256) {
            int rr = e / PK, kk = e % PK;
            uint64_t row = row0 + rr;
            uint32_t i = i0 + kk;
            xs[kk][rr] = (row < n && i < f) ? (T)x[row * f + i] : T(0);
        }
        for (int e = threadIdx.x; e < PK * BC; e += 256) {
            int kk = e / BC, cc = e % BC;
            uint32_t i = i0 + kk, c = col0 + cc;
            ss[kk][cc] = (i < f && c < r) ? s[(uint64_t)i * r + c] : T(0);
        }
        __syncthreads();
        const int kmax = (f - i0) < (uint32_t)PK ? (int)(f - i0) : PK;   // padded terms must not be folded in: x + 0*s changes -0
        for (int kk = 0; kk < kmax; kk++) {
            T xv[4], sv[4];
#pragma unroll
            for (int a = 0; a < 4; a++) xv[a] = xs[kk][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; b++) sv[b] = ss[kk][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if constexpr (CORE) acc[a][b] = acc[a][b] + xv[a] * sv[b];              // clustering.rs:101
                    else acc[a][b] = acc[a][b] + (xv[a] * sv[b]) * scale;                   // reduction.rs:236
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        uint64_t row = row0 + ty * 4 + a;
        if (row >= n) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint32_t c = col0 + tx * 4 + b;
            if (c < r) out[row * r + c] = (double)(CORE ? acc[a][b] * scale : acc[a][b]);   // clustering.rs:104
        }
    }
}

template <typename T>
__global__ void cast_transpose_kernel(const double* __restrict__ in, uint32_t rows, uint32_t cols, int transpose, T* __restrict__ out) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)rows * cols) return;
    uint32_t i = (uint32_t)(g / cols), j = (uint32_t)(g % cols);   // output element (i, j) of rows x cols
    out[g] = (T)(transpose ? in[(uint64_t)j * rows + i] : in[g]);
}

template <typename T, bool CORE>
static int32_t project_launch(sfb_ctx* ctx, const sfb_mat* x, const T* s, uint32_t r, double* out) {
    const bool narrow = ((r + 31) / 32) * 32 < ((r + 63) / 64) * 64;   // less padding with 32-column tiles
    const T scale = T(1) / (T)sqrt((T)r);
    if (narrow) {
        const unsigned grid = ((r + 31) / 32) * div_up(x->rows, 128);
        project_rows_kernel<T, 8, CORE><<<grid, 256, 0, ctx->stream>>>(x->d, x->rows, x->cols, s, r, scale, out);
    } else {
        const unsigned grid = ((r + 63) / 64) * div_up(x->rows, 64);
        project_rows_kernel<T, 16, CORE><<<grid, 256, 0, ctx->stream>>>(x->d, x->rows, x->cols, s, r, scale, out);
    }
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

extern "C" int32_t sfb_project_rows(sfb_ctx* ctx, const sfb_mat* x, const double* samples, uint32_t reduced_dim, int32_t order, sfb_mat** out) {
    if (!ctx || !x || !samples || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (reduced_dim == 0) return sfb_fail(ctx, SFB_EINVAL, "reduced_dim must be positive");
    if (order != SFB_PROJECT_LEGACY && order != SFB_PROJECT_CORE_F32) return sfb_fail(ctx, SFB_EINVAL, "unknown projection order %d", order);
    if ((uint64_t)((reduced_dim + 31) / 32) * ((x->rows + 63) / 64) > 0x7FFFFFFFull) return sfb_fail(ctx, SFB_EUNSUPPORTED, "too many tiles for one projection launch");
    const uint32_t f = x->cols, r = reduced_dim;
    StageTimer t(ctx, &ctx->times.ms_lambda);
    DevBuf raw, s;
    SFB_CUDA(ctx, raw.alloc(sizeof(double) * f * r));
    SFB_CUDA(ctx, cudaMemcpyAsync(raw.p, samples, sizeof(double) * f * r, cudaMemcpyHostToDevice, ctx->stream));
    SFB_TRY(sfb_mat_alloc(ctx, x->rows, r, out));
    int32_t st;
    if (order == SFB_PROJECT_LEGACY) {
        st = project_launch<double, false>(ctx, x, raw.as<double>(), r, (*out)->d);
    } else {
        // the successor draws reduced-major (r x f): bring it to f x r, in f32
        st = SFB_OK;
        cudaError_t e = s.alloc(sizeof(float) * f * r);
        if (e != cudaSuccess) st = sfb_fail(ctx, SFB_ENOMEM, "projection samples: %s", cudaGetErrorString(e));
        if (st == SFB_OK) {
            cast_transpose_kernel<float><<<div_up((uint64_t)f * r, 256), 256, 0, ctx->stream>>>(raw.as<double>(), f, r, 1, s.as<float>());
            ctx->times.kernel_launches++;
            st = project_launch<float, true>(ctx, x, s.as<float>(), r, (*out)->d);
        }
    }
    if (st == SFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = sfb_fail(ctx, SFB_ECUDA, "projection: %s", cudaGetErrorString(cudaGetLastError()));
    if (st != SFB_OK) { sfb_mat_free(*out); *out = nullptr; }
    return st;
}

// compute_jl_dimension: src_legacy/reduction.rs:117-171 (core = 0) and surfface-core/src/clustering.rs:113-123 (core != 0).
// Host scalar arithmetic; Rust's float -> usize cast saturates.
static uint64_t sat_usize(double v) { return (v != v || v <= 0.0) ? 0 : (v >= 1.8446744073709552e19 ? UINT64_MAX : (uint64_t)v); }
extern "C" int32_t sfb_compute_jl_dimension(uint64_t n_points, uint64_t original_dim, double epsilon, int32_t core, uint64_t* out) {
    if (!out) return SFB_EINVAL;
    if (original_dim < 32) { *out = original_dim; return SFB_OK; }
    uint64_t v;
    if (core) {
        const float e = (float)epsilon;
        v = sat_usize((double)ceilf(8.0f * logf((float)n_points) / (e * e)));
    } else {
        const uint64_t bound = sat_usize(ceil(8.0 * log((double)n_points) / pow(epsilon, 2.0)));
        v = bound;
        if (original_dim > 2048) {
            const double ratio = (double)original_dim / (double)bound;
            v = sat_usize(ceil((double)bound * (ratio < 10.0 ? 1.2 : (ratio < 100.0 ? 1.5 : 2.0))));
        }
    }
    *out = v < 32 ? 32 : (v > original_dim ? original_dim : v);
    return SFB_OK;
}

// ---- SortedLambdas ----------------------------------------------------------------------------------------------
// OrderedFloat's total order as an unsigned key: -0 folded onto +0, every NaN onto one key above +inf.
__device__ __forceinline__ uint64_t ordered_key(double v) {
    if (v != v) return 0xFFFFFFFFFFFFFFFFull;
    if (v == 0.0) v = 0.0;
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// Position of idx among 0..n-1 ordered by idx.to_string() (the bucket order zadd keeps, sorted_index.rs:23-30), in
// closed form: "0" first, then for every prefix length L the numbers below n that branch off to a smaller digit
// at position L (sibling prefixes [first_L, P_L), each with all its 10^t extensions), plus the proper prefixes.
__device__ __forceinline__ uint32_t decimal_string_rank(uint32_t idx, uint64_t n) {
    if (idx == 0) return 0;
    uint32_t nd = 1;
    uint64_t pw = 10;
    while (nd < 10 && idx >= pw) { pw *= 10; nd++; }
    uint64_t rank = nd;   // "0" and the nd - 1 proper prefixes
    uint64_t div = pw / 10, prev = 0;
    for (uint32_t L = 1; L <= nd; L++, div /= 10) {
        const uint64_t P = idx / div;
        uint64_t a = L == 1 ? 1 : prev * 10, b = P;
        while (a < n) { rank += (b < n ? b : n) - a; a *= 10; b *= 10; }
        prev = P;
    }
    return (uint32_t)rank;
}
constexpr int SL_PASSES = 8;   // the 64-bit lambda key, 8 bits at a time, on top of the string order
__device__ __forceinline__ uint32_t sl_digit(const double* __restrict__ lam, uint32_t idx, int pass) {
    return (uint32_t)(ordered_key(lam[idx]) >> (8 * pass)) & 255u;
}
__global__ void sl_string_order_kernel(uint64_t n, uint32_t* __restrict__ perm) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) perm[decimal_string_rank((uint32_t)g, n)] = (uint32_t)g;
}

constexpr int SL_ITEMS = 8, SL_TILE = 256 * SL_ITEMS;

// hist[d * nblocks + b] = number of elements of tile b with digit d
__global__ void __launch_bounds__(256) sl_hist_kernel(const double* __restrict__ lam, const uint32_t* __restrict__ idx, uint64_t n, int pass,
                                                      uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * SL_TILE;
    for (int it = 0; it < SL_ITEMS; it++) {
        uint64_t e = base + it * 256 + threadIdx.x;
        if (e < n) atomicAdd(&h[sl_digit(lam, idx[e], pass)], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of the digit-major histogram, in place, one block, four entries per thread
__global__ void __launch_bounds__(1024) sl_scan_kernel(uint32_t* __restrict__ hist, uint64_t len) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint64_t base = 0; base < len; base += 4096) {
        const uint64_t e = base + (uint64_t)threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = e + q < len ? hist[e + q] : 0;
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t incl = mine;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(~0u, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint32_t s = warp_sums[lane], si = s;
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(~0u, si, o); if (lane >= o) si += t; }
            warp_sums[lane] = si - s;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        uint32_t run = carry + warp_sums[w] + incl - mine;
#pragma unroll
        for (int q = 0; q < 4; q++) { if (e + q < len) hist[e + q] = run; run += v[q]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = run;
        __syncthreads();
    }
}

// stable scatter: warp w owns elements [w*256, (w+1)*256) of the tile, taken 32 at a time in order
__global__ void __launch_bounds__(256) sl_scatter_kernel(const double* __restrict__ lam, const uint32_t* __restrict__ idx_in, uint64_t n, int pass,
                                                         const uint32_t* __restrict__ offs, uint32_t* __restrict__ idx_out) {
    __shared__ uint32_t cnt[8][256];
    for (int e = threadIdx.x; e < 8 * 256; e += 256) (&cnt[0][0])[e] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * SL_TILE + (uint64_t)w * (32 * SL_ITEMS);
    uint32_t id[SL_ITEMS], dg[SL_ITEMS], rk[SL_ITEMS];
#pragma unroll
    for (int it = 0; it < SL_ITEMS; it++) {
        uint64_t e = base + it * 32 + lane;
        const bool live = e < n;
        id[it] = live ? idx_in[e] : 0;
        dg[it] = live ? sl_digit(lam, id[it], pass) : 256u;   // dead lanes match only one another
        const uint32_t peers = __match_any_sync(~0u, dg[it]);
        if (live) {
            rk[it] = cnt[w][dg[it]] + __popc(peers & ((1u << lane) - 1));
        }
        __syncwarp();
        if (live && lane == (31 - __clz(peers))) cnt[w][dg[it]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // per digit: global offset of this tile, then exclusive over the eight warps
        const int d = threadIdx.x;
        uint32_t run = offs[(uint64_t)d * gridDim.x + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < 8; ww++) { uint32_t c = cnt[ww][d]; cnt[ww][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SL_ITEMS; it++)
        if (dg[it] < 256u) idx_out[cnt[w][dg[it]] + rk[it]] = id[it];
}

// The reference's std_dev is two strictly sequential folds (an f64 sum, then an f32 sum of squares, laplacian.rs:
// 421-448); sums do not reassociate, so ONE thread carries each chain.  Warps 1-3 stage the next tile in shared memory
// (and square the deviations for the second fold) while thread 0 folds the current one: the chain runs at the
// DADD / FADD latency, nothing else.  Launched beside the sort on its own stream.
constexpr int SD_TILE = 2048;
__global__ void __launch_bounds__(128) sl_stddev_kernel(const double* __restrict__ lam, uint64_t n, double* __restrict__ out) {
    __shared__ double buf[2][SD_TILE];
    __shared__ float mean_s;
    const int tid = threadIdx.x;
    const uint64_t ntiles = (n + SD_TILE - 1) / SD_TILE;
    for (int e = tid; e < SD_TILE; e += 128) buf[0][e] = (uint64_t)e < n ? lam[e] : 0.0;
    __syncthreads();
    double sum = 0.0;
    for (uint64_t t = 0; t < ntiles; t++) {
        if (tid >= 32) {
            const uint64_t base = (t + 1) * SD_TILE;
            if (base < n) for (int e = tid - 32; e < SD_TILE; e += 96) buf[(t + 1) & 1][e] = base + e < n ? lam[base + e] : 0.0;
        } else if (tid == 0) {
            const double* b = buf[t & 1];
            const int live = n - t * SD_TILE < (uint64_t)SD_TILE ? (int)(n - t * SD_TILE) : SD_TILE;
#pragma unroll 16
            for (int e = 0; e < live; e++) sum = sum + b[e];
        }
        __syncthreads();
    }
    if (tid == 0) mean_s = (float)sum / (float)n;   // laplacian.rs:422-426
    __syncthreads();
    const float mean = mean_s;
    float* sq = reinterpret_cast<float*>(&buf[0][0]);   // two tiles of SD_TILE floats
    for (int e = tid; e < SD_TILE; e += 128) { float d = mean - ((uint64_t)e < n ? (float)lam[e] : 0.f); sq[e] = d * d; }   // :437-439
    __syncthreads();
    float var = 0.f;
    for (uint64_t t = 0; t < ntiles; t++) {
        if (tid >= 32) {
            const uint64_t base = (t + 1) * SD_TILE;
            float* nx = sq + ((t + 1) & 1) * SD_TILE;
            if (base < n) for (int e = tid - 32; e < SD_TILE; e += 96) { float d = mean - (base + e < n ? (float)lam[base + e] : 0.f); nx[e] = d * d; }
        } else if (tid == 0) {
            const float* b = sq + (t & 1) * SD_TILE;
            const int live = n - t * SD_TILE < (uint64_t)SD_TILE ? (int)(n - t * SD_TILE) : SD_TILE;
#pragma unroll 16
            for (int e = 0; e < live; e++) var = var + b[e];
        }
        __syncthreads();
    }
    if (tid == 0) out[0] = (double)sqrtf(var / (float)n);
}

// the key a bucket reports is the one inserted first: only +-0 and NaN buckets can hold different bit patterns
__global__ void sl_first_special_kernel(const double* __restrict__ lam, uint64_t n, uint32_t* __restrict__ first /* [zero, nan] */) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const double v = lam[g];
    if (v == 0.0) atomicMin(&first[0], (uint32_t)g);
    else if (v != v) atomicMin(&first[1], (uint32_t)g);
}
__global__ void sl_gather_kernel(const double* __restrict__ lam, const uint32_t* __restrict__ idx, uint64_t n, const uint32_t* __restrict__ first,
                                 double* __restrict__ out) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double v = lam[idx[g]];
    if (v == 0.0) v = lam[first[0]];
    else if (v != v) v = lam[first[1]];
    out[g] = v;
}

extern "C" int32_t sfb_sorted_lambdas_build(sfb_ctx* ctx, const double* lambdas, uint64_t n, double* out_lambda, uint32_t* out_idx, double* out_std_dev) {
    if (!ctx || !lambdas || !out_lambda || !out_idx || !out_std_dev) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (n == 0) return sfb_fail(ctx, SFB_EINVAL, "cannot compute the standard deviation of zero lambdas");   // sorted_index.rs:36-40 panics
    if (n > 0xFFFFFFFEull) return sfb_fail(ctx, SFB_EINVAL, "item count must fit u32");
    const unsigned nblocks = div_up(n, SL_TILE);
    const uint64_t hlen = (uint64_t)256 * nblocks;
    StageTimer t(ctx, &ctx->times.ms_lambda);
    DevBuf lam, a, b, hist, sd, first, sorted;
    SFB_CUDA(ctx, lam.alloc(sizeof(double) * n)); SFB_CUDA(ctx, a.alloc(sizeof(uint32_t) * n)); SFB_CUDA(ctx, b.alloc(sizeof(uint32_t) * n));
    SFB_CUDA(ctx, hist.alloc(sizeof(uint32_t) * hlen)); SFB_CUDA(ctx, sd.alloc(sizeof(double))); SFB_CUDA(ctx, first.alloc(2 * sizeof(uint32_t)));
    SFB_CUDA(ctx, sorted.alloc(sizeof(double) * n));
    SFB_CUDA(ctx, cudaMemcpyAsync(lam.p, lambdas, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemsetAsync(first.p, 0xFF, 2 * sizeof(uint32_t), ctx->stream));
    // the serial folds run beside the sort
    cudaStream_t aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    struct Aux { cudaStream_t& s; cudaEvent_t &a, &b; ~Aux() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } aux_guard{aux, fork, join};
    SFB_CUDA(ctx, cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
    SFB_CUDA(ctx, cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    SFB_CUDA(ctx, cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    SFB_CUDA(ctx, cudaEventRecord(fork, ctx->stream));
    SFB_CUDA(ctx, cudaStreamWaitEvent(aux, fork, 0));
    sl_stddev_kernel<<<1, 128, 0, aux>>>(lam.as<double>(), n, sd.as<double>());
    SFB_LAUNCH_CHECK(ctx);
    SFB_CUDA(ctx, cudaEventRecord(join, aux));
    sl_first_special_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(lam.as<double>(), n, first.as<uint32_t>());
    SFB_LAUNCH_CHECK(ctx);
    // ties first: the string order of the indices is a fixed permutation; then a stable sort by lambda on top of it
    sl_string_order_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(n, a.as<uint32_t>());
    SFB_LAUNCH_CHECK(ctx);
    const uint32_t* src = a.as<uint32_t>();
    uint32_t* dst = b.as<uint32_t>();
    for (int pass = 0; pass < SL_PASSES; pass++) {
        sl_hist_kernel<<<nblocks, 256, 0, ctx->stream>>>(lam.as<double>(), src, n, pass, hist.as<uint32_t>());
        SFB_LAUNCH_CHECK(ctx);
        sl_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist.as<uint32_t>(), hlen);
        SFB_LAUNCH_CHECK(ctx);
        sl_scatter_kernel<<<nblocks, 256, 0, ctx->stream>>>(lam.as<double>(), src, n, pass, hist.as<uint32_t>(), dst);
        SFB_LAUNCH_CHECK(ctx);
        const uint32_t* done = dst;
        dst = const_cast<uint32_t*>(src);
        src = done;
    }
    SFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, join, 0));
    sl_gather_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(lam.as<double>(), src, n, first.as<uint32_t>(), sorted.as<double>());
    SFB_LAUNCH_CHECK(ctx);
    SFB_CUDA(ctx, cudaMemcpyAsync(out_idx, src, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(out_lambda, sorted.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(out_std_dev, sd.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}
