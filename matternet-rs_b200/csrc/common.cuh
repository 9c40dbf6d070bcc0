// common.cuh -- context, handles, error plumbing shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>

#include <map>
#include <string>
#include <unordered_map>

#include "../../include/surfface_b200.h"

struct sfb_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;  // sfb_timer_start / stop
    std::string last_error;
    sfb_stage_times times{};
    // device-memory cache (api.cu: sfb_dev_alloc / sfb_dev_free)
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> block_size;
    size_t cached_bytes = 0;
    // side stream: work that hides behind the screen kernel (knn.cu: sfb_knn_build_columns_begin / _end)
    cudaStream_t side = nullptr;
    cudaEvent_t side_fork = nullptr, side_done = nullptr;
    struct sfb_pending* side_job = nullptr;   // registered, not launched yet
    // NCCL (loaded lazily, comm.cu)
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    bool knn_collective = false;   // inside sfb_knn_build_sharded: operand preparation is split across the ranks (knn_screen.cu)
    bool lambda_sharded = false;   // inside sfb_lambda_allgather: per-rank totals of CORE_F32SEM are all-reduced (lambda.cu)
};

struct sfb_mat {
    sfb_ctx* ctx;
    double* d = nullptr;  // rows x cols, row-major
    uint64_t rows = 0;
    uint32_t cols = 0;
    bool owns = true;  // false: a row-range view into another matrix
};

struct sfb_knn {
    sfb_ctx* ctx;
    uint32_t* idx = nullptr;  // rows x k
    double* dist = nullptr;   // rows x k
    uint32_t* cnt = nullptr;  // rows
    uint64_t rows = 0;        // query rows held
    uint64_t q_begin = 0;     // global index of local row 0
    uint64_t total = 0;       // corpus rows (node count)
    uint32_t k = 0;
    sfb_knn_stats stats{};
};

struct sfb_adj {
    sfb_ctx* ctx;
    uint32_t* idx = nullptr;  // rows x k
    double* w = nullptr;      // rows x k
    uint32_t* cnt = nullptr;  // rows
    uint64_t rows = 0;
    uint32_t k = 0;
};

struct sfb_csr {
    sfb_ctx* ctx;
    uint64_t* indptr = nullptr;  // rows + 1
    uint32_t* indices = nullptr;
    double* data = nullptr;
    uint64_t rows = 0, nnz = 0;
    uint64_t row0 = 0, cols = 0;  // a row shard of a larger matrix (sfb_laplacian_build_rows): global index of row 0, column count (0: square)
    mutable int symmetric = -1;  // -1 unknown, 1: structure and values symmetric bit for bit (set by the builder, else checked once in lambda.cu), 0: not
    // packed strict upper triangle for the lambda tile kernel (lambda.cu: lt_pack), built on first use, released by sfb_csr_free
    mutable void* lt_recs = nullptr; mutable void* lt_defect = nullptr; mutable void* lt_meta = nullptr;
};

int32_t sfb_fail(sfb_ctx* ctx, int32_t code, const char* fmt, ...);

// Device memory comes from a per-context cache of cudaMalloc'ed blocks: a build allocates and frees
// GB-sized scratch at every stage, and plain cudaMalloc / cudaFree synchronise the device and cost
// hundreds of milliseconds per step.  (cudaMallocAsync was tried first: its reuse-or-grow decision
// depends on whether earlier frees have retired, which made one step in five stall for up to a second.)
// All work of a context runs on its one stream, so a freed block can be handed out again at once:
// stream order is program order.  Blocks are returned to the driver when an allocation fails and when
// the context is destroyed.
extern thread_local sfb_ctx* sfb_tls_ctx;  // the context of the call in flight (set by SFB_CUDA / sfb_dev_alloc)
cudaError_t sfb_dev_alloc(sfb_ctx* ctx, void** p, size_t bytes);
void sfb_dev_free(sfb_ctx* ctx, void* p);

#define SFB_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        sfb_tls_ctx = (ctx);                                                                     \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return sfb_fail((ctx), e__ == cudaErrorMemoryAllocation ? SFB_ENOMEM : SFB_ECUDA,    \
                            "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define SFB_TRY(call)                   \
    do {                                \
        int32_t s__ = (call);           \
        if (s__ != SFB_OK) return s__;  \
    } while (0)

#define SFB_LAUNCH_CHECK(ctx)                  \
    do {                                       \
        (ctx)->times.kernel_launches++;        \
        SFB_CUDA((ctx), cudaGetLastError());   \
    } while (0)

// device scratch that frees itself on every return path
struct DevBuf {
    void* p = nullptr;
    sfb_ctx* owner = nullptr;
    ~DevBuf() { if (p) sfb_dev_free(owner, p); }
    cudaError_t alloc(size_t bytes) { owner = sfb_tls_ctx; return sfb_dev_alloc(owner, &p, bytes); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    void* release() { void* q = p; p = nullptr; return q; }
};

// device-time stopwatch on the context's stream (own events: timers nest)
struct StageTimer {
    sfb_ctx* ctx;
    double* acc;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool done = false;
    StageTimer(sfb_ctx* c, double* a) : ctx(c), acc(a) {
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx->stream);
    }
    ~StageTimer() { if (!done) stop(); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    double stop() {
        if (done) return 0.0;
        done = true;
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (acc) *acc += ms;
        return ms;
    }
};

// SFB_TRACE=1: host wall-clock checkpoints to stderr (synchronises the stream at every checkpoint)
struct HostTrace {
    sfb_ctx* ctx; const char* what; bool on; double t0 = 0.0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    HostTrace(sfb_ctx* c, const char* w) : ctx(c), what(w) { static const bool e = getenv("SFB_TRACE") != nullptr; on = e; if (on) { cudaStreamSynchronize(ctx->stream); t0 = now(); } }
    void mark(const char* label) { if (!on) return; cudaStreamSynchronize(ctx->stream); double t = now(); fprintf(stderr, "[sfb] %-14s %-18s %9.3f ms\n", what, label, t - t0); t0 = t; }
};

static inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// ---- internal entry points between translation units -------------------------------------------
int32_t sfb_scan_exclusive_u64(sfb_ctx* ctx, const uint32_t* in, uint64_t n, uint64_t* out /* n+1 */);   // asynchronous
int32_t sfb_scan_exclusive_u64_ex(sfb_ctx* ctx, const uint32_t* in_a, const uint32_t* in_b /* or null */, uint32_t add, uint64_t n, uint64_t* out);
int32_t sfb_knn_exact(sfb_ctx* ctx, const sfb_mat* x, const double* norms, int metric, uint32_t k, double eps,
                      const uint32_t* query_rows /* device, or null */, uint64_t nq, uint64_t q_begin,
                      uint32_t* out_idx, double* out_dist, uint32_t* out_cnt);
void sfb_comm_destroy(sfb_ctx* ctx);
int32_t sfb_comm_allgather_bytes(sfb_ctx* ctx, void* base, size_t bytes_per_rank);   // comm.cu: in-place, slot r = base + r * bytes
int32_t sfb_comm_allreduce_max_u64(sfb_ctx* ctx, unsigned long long* buf, size_t n);
int32_t sfb_mat_alloc(sfb_ctx* ctx, uint64_t rows, uint32_t cols, sfb_mat** out);  // api.cu: uninitialised rows x cols f64
// launches the registered side job, if any (called right after the screen kernel is enqueued, so that the persistent
// screen CTAs are placed first and the side kernel's small CTAs fill in beside them)
void sfb_side_job_fire(sfb_ctx* ctx);
int32_t sfb_row_norms(sfb_ctx* ctx, const sfb_mat* x, double* norms);
int32_t sfb_knn_screened(sfb_ctx* ctx, const sfb_mat* x, const double* norms, const sfb_knn_params* p,
                         uint64_t q_begin, uint64_t q_end, sfb_knn* out);
