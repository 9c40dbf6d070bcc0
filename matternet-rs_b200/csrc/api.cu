// api.cu -- context, dense matrices, handle plumbing, synthetic rows, exclusive scan.
#include <stdarg.h>

#include <mutex>
#include <new>
#include <set>

#include "common.cuh"
#include "synth.cuh"

int32_t sfb_fail(sfb_ctx* ctx, int32_t code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return code;
}

thread_local sfb_ctx* sfb_tls_ctx = nullptr;

static void cache_release_all(sfb_ctx* ctx) {
    for (auto& kv : ctx->free_blocks) { cudaFree(kv.second); ctx->block_size.erase(kv.second); }
    ctx->free_blocks.clear();
    ctx->cached_bytes = 0;
}

cudaError_t sfb_dev_alloc(sfb_ctx* ctx, void** p, size_t bytes) {
    if (!bytes) bytes = 16;
    if (!ctx) return cudaMalloc(p, bytes);
    sfb_tls_ctx = ctx;
    const size_t want = (bytes + 511) & ~(size_t)511;
    auto it = ctx->free_blocks.lower_bound(want);
    if (it != ctx->free_blocks.end() && it->first <= want + want / 4 + (1u << 20)) {   // best fit, bounded waste
        *p = it->second;
        ctx->cached_bytes -= it->first;
        ctx->free_blocks.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e == cudaErrorMemoryAllocation && !ctx->free_blocks.empty()) {
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        cache_release_all(ctx);
        e = cudaMalloc(p, want);
    }
    if (e == cudaSuccess) ctx->block_size[*p] = want;
    return e;
}
// handles may outlive their context (garbage-collected host languages): only live contexts are dereferenced
static std::mutex g_live_mu;
static std::set<const sfb_ctx*> g_live;
static bool ctx_live(const sfb_ctx* c) { std::lock_guard<std::mutex> l(g_live_mu); return g_live.count(c) != 0; }
void sfb_dev_free(sfb_ctx* ctx, void* p) {
    if (!p) return;
    if (ctx && ctx_live(ctx)) {
        auto it = ctx->block_size.find(p);
        if (it != ctx->block_size.end()) { ctx->free_blocks.emplace(it->second, p); ctx->cached_bytes += it->second; return; }
    }
    cudaFree(p);
}

extern "C" int32_t sfb_abi_version(void) { return SFB_ABI_VERSION; }

extern "C" int32_t sfb_ctx_create(int32_t device_id, sfb_ctx** out) {
    if (!out) return SFB_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return SFB_ECUDA;  // no CPU fallback
    if (device_id < 0 || device_id >= n) return SFB_EINVAL;
    sfb_ctx* ctx = new (std::nothrow) sfb_ctx();
    if (!ctx) return SFB_ENOMEM;
    ctx->device = device_id;
    cudaDeviceProp prop;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SFB_ECUDA;
    }
    if (prop.major != 10) {  // sm_100a only: no multi-arch dispatch
        fprintf(stderr, "surfface_b200: device %d is sm_%d%d, this library is built for sm_100a only\n",
                device_id, prop.major, prop.minor);
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return SFB_EUNSUPPORTED;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    { std::lock_guard<std::mutex> l(g_live_mu); g_live.insert(ctx); }
    *out = ctx;
    return SFB_OK;
}

extern "C" void sfb_ctx_destroy(sfb_ctx* ctx) {
    if (!ctx) return;
    { std::lock_guard<std::mutex> l(g_live_mu); g_live.erase(ctx); }
    if (sfb_tls_ctx == ctx) sfb_tls_ctx = nullptr;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    sfb_comm_destroy(ctx);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaEventDestroy(ctx->side_fork); cudaEventDestroy(ctx->side_done); cudaStreamDestroy(ctx->side); }
    cache_release_all(ctx);
    if (ctx->timer0) { cudaEventDestroy(ctx->timer0); cudaEventDestroy(ctx->timer1); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* sfb_last_error(const sfb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" int32_t sfb_device_info(const sfb_ctx* ctx, char* name, int32_t* sm_count, uint64_t* hbm_bytes) {
    if (!ctx) return SFB_EINVAL;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) return SFB_ECUDA;
    if (name) { strncpy(name, prop.name, 255); name[255] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
    return SFB_OK;
}

extern "C" int32_t sfb_synchronize(sfb_ctx* ctx) {
    if (!ctx) return SFB_EINVAL;
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}

extern "C" int32_t sfb_timer_start(sfb_ctx* ctx) {
    if (!ctx) return SFB_EINVAL;
    if (!ctx->timer0) { SFB_CUDA(ctx, cudaEventCreate(&ctx->timer0)); SFB_CUDA(ctx, cudaEventCreate(&ctx->timer1)); }
    SFB_CUDA(ctx, cudaEventRecord(ctx->timer0, ctx->stream));
    return SFB_OK;
}
extern "C" int32_t sfb_timer_stop(sfb_ctx* ctx, double* ms) {
    if (!ctx || !ms || !ctx->timer0) return sfb_fail(ctx, SFB_EINVAL, "timer not started");
    SFB_CUDA(ctx, cudaEventRecord(ctx->timer1, ctx->stream));
    SFB_CUDA(ctx, cudaEventSynchronize(ctx->timer1));
    float f = 0.f;
    SFB_CUDA(ctx, cudaEventElapsedTime(&f, ctx->timer0, ctx->timer1));
    *ms = f;
    return SFB_OK;
}

extern "C" int32_t sfb_pinned_alloc(sfb_ctx* ctx, uint64_t bytes, void** out) {
    if (!ctx || !out) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault);
    if (e != cudaSuccess) return sfb_fail(ctx, SFB_ENOMEM, "cudaHostAlloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    return SFB_OK;
}
extern "C" void sfb_pinned_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int32_t sfb_timings(const sfb_ctx* ctx, sfb_stage_times* out) {
    if (!ctx || !out) return SFB_EINVAL;
    *out = ctx->times;
    return SFB_OK;
}
extern "C" int32_t sfb_timings_reset(sfb_ctx* ctx) {
    if (!ctx) return SFB_EINVAL;
    uint64_t l = ctx->times.kernel_launches;
    ctx->times = sfb_stage_times{};
    ctx->times.kernel_launches = l;
    return SFB_OK;
}

// ---- matrices ---------------------------------------------------------------------------------
int32_t sfb_mat_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t cols, sfb_mat** out) {
    if (!ctx || !out || rows == 0 || cols == 0) return sfb_fail(ctx, SFB_EINVAL, "matrix must be non-empty");
    if (rows > 0xFFFFFFFEull) return sfb_fail(ctx, SFB_EINVAL, "rows must fit u32 node indices");
    sfb_mat* m = new (std::nothrow) sfb_mat();
    if (!m) return SFB_ENOMEM;
    m->ctx = ctx; m->rows = rows; m->cols = cols;
    cudaError_t e = sfb_dev_alloc(ctx, (void**)&m->d, sizeof(double) * rows * cols);
    if (e != cudaSuccess) { delete m; return sfb_fail(ctx, SFB_ENOMEM, "cudaMalloc %llu x %u f64: %s", (unsigned long long)rows, cols, cudaGetErrorString(e)); }
    *out = m;
    return SFB_OK;
}

extern "C" int32_t sfb_mat_from_host(sfb_ctx* ctx, const double* x, uint64_t rows, uint32_t cols, sfb_mat** out) {
    if (!x) return sfb_fail(ctx, SFB_EINVAL, "null host pointer");
    SFB_TRY(sfb_mat_alloc(ctx, rows, cols, out));
    StageTimer t(ctx, &ctx->times.ms_h2d);
    cudaError_t e = cudaMemcpyAsync((*out)->d, x, sizeof(double) * rows * cols, cudaMemcpyHostToDevice, ctx->stream);
    t.stop();
    if (e != cudaSuccess) { sfb_mat_free(*out); *out = nullptr; return sfb_fail(ctx, SFB_ECUDA, "H2D: %s", cudaGetErrorString(e)); }
    return SFB_OK;
}

__global__ void widen_f32_kernel(const float* __restrict__ in, uint64_t n, double* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

extern "C" int32_t sfb_mat_from_host_f32(sfb_ctx* ctx, const float* x, uint64_t rows, uint32_t cols, sfb_mat** out) {
    if (!x) return sfb_fail(ctx, SFB_EINVAL, "null host pointer");
    SFB_TRY(sfb_mat_alloc(ctx, rows, cols, out));
    const uint64_t n = rows * cols;
    DevBuf stage;
    cudaError_t e = (sfb_tls_ctx = ctx, stage.alloc(sizeof(float) * n));
    if (e == cudaSuccess) {
        StageTimer t(ctx, &ctx->times.ms_h2d);
        e = cudaMemcpyAsync(stage.p, x, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            widen_f32_kernel<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(stage.as<float>(), n, (*out)->d);
            ctx->times.kernel_launches++;
            e = cudaGetLastError();
        }
        t.stop();   // synchronises: the staging buffer may be released
    }
    if (e != cudaSuccess) { sfb_mat_free(*out); *out = nullptr; return sfb_fail(ctx, SFB_ECUDA, "f32 upload: %s", cudaGetErrorString(e)); }
    return SFB_OK;
}

extern "C" int32_t sfb_mat_clone(sfb_ctx* ctx, const sfb_mat* a, sfb_mat** out) {
    if (!a) return sfb_fail(ctx, SFB_EINVAL, "null matrix");
    SFB_TRY(sfb_mat_alloc(ctx, a->rows, a->cols, out));
    cudaError_t e = cudaMemcpyAsync((*out)->d, a->d, sizeof(double) * a->rows * a->cols, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { sfb_mat_free(*out); *out = nullptr; return sfb_fail(ctx, SFB_ECUDA, "device copy: %s", cudaGetErrorString(e)); }
    return SFB_OK;
}

__global__ void generate_rows_kernel(double* __restrict__ out, int kind, uint64_t seed, uint64_t rows, uint32_t cols,
                                     uint32_t n_centres, double noise) {
    const uint32_t quads = (cols + 3) / 4;
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= rows * quads) return;
    uint64_t i = gid / quads;
    uint32_t q = (uint32_t)(gid % quads);
    double v[4];
    synth_row_quad(kind, seed, i, q, n_centres, noise, v);
    double* o = out + i * cols + (uint64_t)q * 4;
    for (uint32_t t = 0; t < 4 && q * 4 + t < cols; ++t) o[t] = v[t];
}

extern "C" int32_t sfb_mat_generate(sfb_ctx* ctx, int32_t kind, uint64_t seed, uint64_t rows, uint32_t cols,
                                    uint32_t n_centres, double noise, sfb_mat** out) {
    if (kind < 0 || kind > 2) return sfb_fail(ctx, SFB_EINVAL, "unknown synthetic kind %d", kind);
    if (kind == 1 && n_centres == 0) return sfb_fail(ctx, SFB_EINVAL, "clustered rows need n_centres > 0");
    SFB_TRY(sfb_mat_alloc(ctx, rows, cols, out));
    uint64_t total = rows * ((cols + 3) / 4);
    generate_rows_kernel<<<div_up(total, 256), 256, 0, ctx->stream>>>((*out)->d, kind, seed, rows, cols, n_centres, noise);
    SFB_LAUNCH_CHECK(ctx);
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}

// 32x32 tile transpose through shared memory: coalesced on both sides.  One block per tile, tiles
// numbered along the LONGER side first so either shape (few rows x millions of columns, or the reverse) fits.
__global__ void transpose_kernel(const double* __restrict__ a, double* __restrict__ b, uint64_t rows, uint64_t cols, uint64_t tiles_r) {
    __shared__ double tile[32][33];
    const uint64_t t = blockIdx.x;
    const uint64_t r0 = (t % tiles_r) * 32, c0 = (t / tiles_r) * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        uint64_t r = r0 + dy, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[dy][threadIdx.x] = a[r * cols + c];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        uint64_t c = c0 + dy, r = r0 + threadIdx.x;
        if (r < rows && c < cols) b[c * rows + r] = tile[threadIdx.x][dy];
    }
}

int32_t sfb_transpose_device(sfb_ctx* ctx, const double* a, uint64_t rows, uint64_t cols, double* b) {
    const uint64_t tiles_r = (rows + 31) / 32, tiles_c = (cols + 31) / 32;
    if (tiles_r * tiles_c > 0x7FFFFFFFull) return sfb_fail(ctx, SFB_EUNSUPPORTED, "matrix too large to transpose in one launch");
    transpose_kernel<<<(unsigned)(tiles_r * tiles_c), dim3(32, 8), 0, ctx->stream>>>(a, b, rows, cols, tiles_r);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

extern "C" int32_t sfb_mat_transpose(sfb_ctx* ctx, const sfb_mat* a, sfb_mat** out) {
    if (!a) return sfb_fail(ctx, SFB_EINVAL, "null matrix");
    if (a->rows > 0xFFFFFFFFull) return sfb_fail(ctx, SFB_EINVAL, "too many rows to transpose");
    SFB_TRY(sfb_mat_alloc(ctx, a->cols, (uint32_t)a->rows, out));
    return sfb_transpose_device(ctx, a->d, a->rows, a->cols, (*out)->d);
}

extern "C" int32_t sfb_mat_shape(const sfb_mat* a, uint64_t* rows, uint32_t* cols) {
    if (!a) return SFB_EINVAL;
    if (rows) *rows = a->rows;
    if (cols) *cols = a->cols;
    return SFB_OK;
}

extern "C" int32_t sfb_mat_copy_rows(sfb_ctx* ctx, const sfb_mat* a, uint64_t row0, uint64_t nrows, double* out) {
    if (!a || !out || row0 + nrows > a->rows) return sfb_fail(ctx, SFB_EINVAL, "row range out of bounds");
    SFB_CUDA(ctx, cudaMemcpyAsync(out, a->d + row0 * a->cols, sizeof(double) * nrows * a->cols, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}

extern "C" int32_t sfb_mat_view_rows(sfb_ctx* ctx, const sfb_mat* a, uint64_t row0, uint64_t nrows, sfb_mat** out) {
    if (!a || !out || nrows == 0 || row0 + nrows > a->rows) return sfb_fail(ctx, SFB_EINVAL, "row range out of bounds");
    sfb_mat* v = new (std::nothrow) sfb_mat();
    if (!v) return SFB_ENOMEM;
    v->ctx = ctx; v->d = a->d + row0 * a->cols; v->rows = nrows; v->cols = a->cols; v->owns = false;
    *out = v;
    return SFB_OK;
}

extern "C" void sfb_mat_free(sfb_mat* a) {
    if (!a) return;
    if (a->owns) sfb_dev_free(a->ctx, a->d);
    delete a;
}

// ---- kNN / adjacency / CSR handles ------------------------------------------------------------
// Lists from the host index device arrays downstream (adjacency_kernel reads in_cnt[j], the reverse-edge kernels count into
// rev_cnt[j]): a bad index would corrupt neighbouring blocks silently where the reference panics on an out-of-bounds index.
static int32_t sfb_check_lists(sfb_ctx* ctx, const uint32_t* idx, const uint32_t* cnt, uint64_t rows, uint32_t k) {
    for (uint64_t i = 0; i < rows; ++i) {
        if (cnt[i] > k) return sfb_fail(ctx, SFB_EINVAL, "row %llu: count %u exceeds k = %u", (unsigned long long)i, cnt[i], k);
        for (uint32_t t = 0; t < cnt[i]; ++t)
            if (idx[i * k + t] >= rows)
                return sfb_fail(ctx, SFB_EINVAL, "row %llu: neighbour %u is index %u, not below %llu rows", (unsigned long long)i, t, idx[i * k + t], (unsigned long long)rows);
    }
    return SFB_OK;
}
extern "C" int32_t sfb_knn_shape(const sfb_knn* g, uint64_t* rows, uint32_t* k, uint64_t* q_begin) {
    if (!g) return SFB_EINVAL;
    if (rows) *rows = g->rows;
    if (k) *k = g->k;
    if (q_begin) *q_begin = g->q_begin;
    return SFB_OK;
}
extern "C" int32_t sfb_knn_copy(sfb_ctx* ctx, const sfb_knn* g, uint32_t* idx, double* dist, uint32_t* cnt) {
    if (!g) return sfb_fail(ctx, SFB_EINVAL, "null kNN handle");
    StageTimer t(ctx, &ctx->times.ms_d2h);
    if (idx) SFB_CUDA(ctx, cudaMemcpyAsync(idx, g->idx, sizeof(uint32_t) * g->rows * g->k, cudaMemcpyDeviceToHost, ctx->stream));
    if (dist) SFB_CUDA(ctx, cudaMemcpyAsync(dist, g->dist, sizeof(double) * g->rows * g->k, cudaMemcpyDeviceToHost, ctx->stream));
    if (cnt) SFB_CUDA(ctx, cudaMemcpyAsync(cnt, g->cnt, sizeof(uint32_t) * g->rows, cudaMemcpyDeviceToHost, ctx->stream));
    t.stop();
    return SFB_OK;
}
extern "C" int32_t sfb_knn_stats_get(const sfb_knn* g, sfb_knn_stats* out) {
    if (!g || !out) return SFB_EINVAL;
    *out = g->stats;
    return SFB_OK;
}
int32_t sfb_knn_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t k, sfb_knn** out) {
    sfb_knn* g = new (std::nothrow) sfb_knn();
    if (!g) return SFB_ENOMEM;
    g->ctx = ctx; g->rows = rows; g->k = k; g->total = rows;
    if (sfb_dev_alloc(ctx, (void**)&g->idx, sizeof(uint32_t) * rows * k) != cudaSuccess ||
        sfb_dev_alloc(ctx, (void**)&g->dist, sizeof(double) * rows * k) != cudaSuccess ||
        sfb_dev_alloc(ctx, (void**)&g->cnt, sizeof(uint32_t) * rows) != cudaSuccess) {
        sfb_knn_free(g);
        return sfb_fail(ctx, SFB_ENOMEM, "kNN list allocation failed (%llu x %u)", (unsigned long long)rows, k);
    }
    *out = g;
    return SFB_OK;
}
extern "C" int32_t sfb_knn_from_host(sfb_ctx* ctx, const uint32_t* idx, const double* dist, const uint32_t* cnt,
                                     uint64_t rows, uint32_t k, sfb_knn** out) {
    if (!ctx || !idx || !dist || !cnt || !out || rows == 0 || k == 0) return sfb_fail(ctx, SFB_EINVAL, "bad kNN arrays");
    SFB_TRY(sfb_check_lists(ctx, idx, cnt, rows, k));
    SFB_TRY(sfb_knn_alloc(ctx, rows, k, out));
    sfb_knn* g = *out;
    SFB_CUDA(ctx, cudaMemcpyAsync(g->idx, idx, sizeof(uint32_t) * rows * k, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(g->dist, dist, sizeof(double) * rows * k, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(g->cnt, cnt, sizeof(uint32_t) * rows, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}
extern "C" void sfb_knn_free(sfb_knn* g) {
    if (!g) return;
    sfb_dev_free(g->ctx, g->idx); sfb_dev_free(g->ctx, g->dist); sfb_dev_free(g->ctx, g->cnt);
    delete g;
}

extern "C" int32_t sfb_adj_shape(const sfb_adj* a, uint64_t* rows, uint32_t* k) {
    if (!a) return SFB_EINVAL;
    if (rows) *rows = a->rows;
    if (k) *k = a->k;
    return SFB_OK;
}
extern "C" int32_t sfb_adj_copy(sfb_ctx* ctx, const sfb_adj* a, uint32_t* idx, double* w, uint32_t* cnt) {
    if (!a) return sfb_fail(ctx, SFB_EINVAL, "null adjacency handle");
    if (idx) SFB_CUDA(ctx, cudaMemcpyAsync(idx, a->idx, sizeof(uint32_t) * a->rows * a->k, cudaMemcpyDeviceToHost, ctx->stream));
    if (w) SFB_CUDA(ctx, cudaMemcpyAsync(w, a->w, sizeof(double) * a->rows * a->k, cudaMemcpyDeviceToHost, ctx->stream));
    if (cnt) SFB_CUDA(ctx, cudaMemcpyAsync(cnt, a->cnt, sizeof(uint32_t) * a->rows, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}
int32_t sfb_adj_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t k, sfb_adj** out) {
    sfb_adj* a = new (std::nothrow) sfb_adj();
    if (!a) return SFB_ENOMEM;
    a->ctx = ctx; a->rows = rows; a->k = k;
    if (sfb_dev_alloc(ctx, (void**)&a->idx, sizeof(uint32_t) * rows * k) != cudaSuccess ||
        sfb_dev_alloc(ctx, (void**)&a->w, sizeof(double) * rows * k) != cudaSuccess ||
        sfb_dev_alloc(ctx, (void**)&a->cnt, sizeof(uint32_t) * rows) != cudaSuccess) {
        sfb_adj_free(a);
        return sfb_fail(ctx, SFB_ENOMEM, "adjacency allocation failed");
    }
    *out = a;
    return SFB_OK;
}
extern "C" int32_t sfb_adj_from_host(sfb_ctx* ctx, const uint32_t* idx, const double* w, const uint32_t* cnt,
                                     uint64_t rows, uint32_t k, sfb_adj** out) {
    if (!ctx || !idx || !w || !cnt || !out || rows == 0 || k == 0) return sfb_fail(ctx, SFB_EINVAL, "bad adjacency arrays");
    SFB_TRY(sfb_check_lists(ctx, idx, cnt, rows, k));
    SFB_TRY(sfb_adj_alloc(ctx, rows, k, out));
    sfb_adj* a = *out;
    SFB_CUDA(ctx, cudaMemcpyAsync(a->idx, idx, sizeof(uint32_t) * rows * k, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(a->w, w, sizeof(double) * rows * k, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(a->cnt, cnt, sizeof(uint32_t) * rows, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}
extern "C" void sfb_adj_free(sfb_adj* a) {
    if (!a) return;
    sfb_dev_free(a->ctx, a->idx); sfb_dev_free(a->ctx, a->w); sfb_dev_free(a->ctx, a->cnt);
    delete a;
}

extern "C" int32_t sfb_csr_shape(const sfb_csr* L, uint64_t* rows, uint64_t* nnz) {
    if (!L) return SFB_EINVAL;
    if (rows) *rows = L->rows;
    if (nnz) *nnz = L->nnz;
    return SFB_OK;
}
extern "C" int32_t sfb_csr_copy(sfb_ctx* ctx, const sfb_csr* L, uint64_t* indptr, uint32_t* indices, double* data) {
    if (!L) return sfb_fail(ctx, SFB_EINVAL, "null CSR handle");
    StageTimer t(ctx, &ctx->times.ms_d2h);
    if (indptr) SFB_CUDA(ctx, cudaMemcpyAsync(indptr, L->indptr, sizeof(uint64_t) * (L->rows + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (indices && L->nnz) SFB_CUDA(ctx, cudaMemcpyAsync(indices, L->indices, sizeof(uint32_t) * L->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    if (data && L->nnz) SFB_CUDA(ctx, cudaMemcpyAsync(data, L->data, sizeof(double) * L->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    t.stop();
    return SFB_OK;
}
extern "C" int32_t sfb_csr_from_host(sfb_ctx* ctx, uint64_t rows, const uint64_t* indptr, const uint32_t* indices,
                                     const double* data, sfb_csr** out) {
    if (!ctx || !indptr || !out || rows == 0) return sfb_fail(ctx, SFB_EINVAL, "bad CSR arrays");
    uint64_t nnz = indptr[rows];
    if (nnz && (!indices || !data)) return sfb_fail(ctx, SFB_EINVAL, "bad CSR arrays");
    for (uint64_t r = 0; r < rows; ++r) {
        if (indptr[r] > indptr[r + 1]) return sfb_fail(ctx, SFB_EINVAL, "indptr not monotone at row %llu", (unsigned long long)r);
        for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e)
            if (indices[e] >= rows || (e > indptr[r] && indices[e] <= indices[e - 1]))
                return sfb_fail(ctx, SFB_EINVAL, "row %llu: column indices must be ascending and < rows", (unsigned long long)r);
    }
    sfb_csr* L = new (std::nothrow) sfb_csr();
    if (!L) return SFB_ENOMEM;
    L->ctx = ctx; L->rows = rows; L->nnz = nnz;
    if (sfb_dev_alloc(ctx, (void**)&L->indptr, sizeof(uint64_t) * (rows + 1)) != cudaSuccess ||
        sfb_dev_alloc(ctx, (void**)&L->indices, sizeof(uint32_t) * (nnz ? nnz : 1)) != cudaSuccess ||
        sfb_dev_alloc(ctx, (void**)&L->data, sizeof(double) * (nnz ? nnz : 1)) != cudaSuccess) {
        sfb_csr_free(L);
        return sfb_fail(ctx, SFB_ENOMEM, "CSR allocation failed");
    }
    SFB_CUDA(ctx, cudaMemcpyAsync(L->indptr, indptr, sizeof(uint64_t) * (rows + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        SFB_CUDA(ctx, cudaMemcpyAsync(L->indices, indices, sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice, ctx->stream));
        SFB_CUDA(ctx, cudaMemcpyAsync(L->data, data, sizeof(double) * nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = L;
    return SFB_OK;
}
extern "C" void sfb_csr_free(sfb_csr* L) {
    if (!L) return;
    sfb_dev_free(L->ctx, L->indptr); sfb_dev_free(L->ctx, L->indices); sfb_dev_free(L->ctx, L->data);
    sfb_dev_free(L->ctx, L->lt_recs); sfb_dev_free(L->ctx, L->lt_defect); sfb_dev_free(L->ctx, L->lt_meta);
    delete L;
}

// ---- exclusive scan u32 -> u64: ONE pass, decoupled look-back ---------------------------------------
// Tiles are taken in ticket order (atomic counter), each publishes its aggregate, then its inclusive prefix, in one
// 64-bit word (flag in the top two bits); a tile resolves its exclusive prefix by walking back over its predecessors'
// words until it meets an inclusive one.  One read and one write of the data, no block-sums array, no host round trip.
// in_b (optional) is added element-wise and `add` to every element: the callers scan cnt[i] + rev_len[i] or ulen[i] + 1
// without materialising them.
static constexpr int SCAN_BLOCK = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;
static constexpr unsigned long long LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total) {
    __shared__ uint64_t warp_sums[SCAN_BLOCK / 32];
    __shared__ uint64_t block_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = lane < SCAN_BLOCK / 32 ? warp_sums[lane] : 0, si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        if (lane < SCAN_BLOCK / 32) warp_sums[lane] = si - s;
        if (lane == 31) block_total = si;
    }
    __syncthreads();
    *total = block_total;
    return inc - v + warp_sums[wid];
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_lookback_kernel(const uint32_t* __restrict__ in_a, const uint32_t* __restrict__ in_b, uint32_t add,
                                                                   uint64_t n, uint64_t* __restrict__ out, volatile unsigned long long* state /* [tiles] + ticket */,
                                                                   uint64_t n_tiles) {
    __shared__ uint64_t s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd((unsigned long long*)(state + n_tiles), 1ull);
    __syncthreads();
    const uint64_t tile = s_tile;
    const uint64_t base = tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) {
        v[t] = base + t < n ? in_a[base + t] + (in_b ? in_b[base + t] : 0u) + add : 0u;
        s += v[t];
    }
    uint64_t total;
    const uint64_t ex_in_block = block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) {
        uint64_t prefix = 0;
        if (tile == 0) state[0] = LB_PREFIX | total;
        else {
            state[tile] = LB_AGG | total;
            for (uint64_t j = tile; j-- > 0;) {
                unsigned long long w;
                do { w = state[j]; } while ((w >> 62) == 0);
                prefix += w & LB_MASK;
                if ((w >> 62) == 2) break;
            }
            state[tile] = LB_PREFIX | (prefix + total);
        }
        s_prefix = prefix;
        if (tile == n_tiles - 1) out[n] = prefix + total;
    }
    __syncthreads();
    uint64_t ex = s_prefix + ex_in_block;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) { if (base + t < n) out[base + t] = ex; ex += v[t]; }
}

// out: n + 1 entries (out[n] = the total).  Asynchronous on the context's stream.
int32_t sfb_scan_exclusive_u64_ex(sfb_ctx* ctx, const uint32_t* in_a, const uint32_t* in_b, uint32_t add, uint64_t n, uint64_t* out) {
    if (n == 0) { SFB_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(uint64_t), ctx->stream)); return SFB_OK; }
    const uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf state;
    SFB_CUDA(ctx, state.alloc(sizeof(unsigned long long) * (nb + 1)));
    SFB_CUDA(ctx, cudaMemsetAsync(state.p, 0, sizeof(unsigned long long) * (nb + 1), ctx->stream));
    scan_lookback_kernel<<<(unsigned)nb, SCAN_BLOCK, 0, ctx->stream>>>(in_a, in_b, add, n, out, state.as<unsigned long long>(), nb);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;   // `state` returns to the context's cache; stream order keeps it valid until the kernel has run
}
int32_t sfb_scan_exclusive_u64(sfb_ctx* ctx, const uint32_t* in, uint64_t n, uint64_t* out) {
    return sfb_scan_exclusive_u64_ex(ctx, in, nullptr, 0u, n, out);
}
