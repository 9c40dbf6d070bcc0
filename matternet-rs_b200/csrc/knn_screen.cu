// knn_screen.cu -- tensor-core distance screen (tcgen05 + TMEM + TMA) with fused per-row top-k'
// selection, exact f64 rescore, per-row certification and exact fallback.
//
// Replaces the two CosinePair sweeps of _build_adjacency (src_legacy/laplacian.rs:213-229,245-254)
// and the brute-force loops of surfface-core/src/mst.rs:312-363 / src_legacy/energymaps.rs:875-892.
// The RESULT is defined by src_legacy/tests/test_helpers.rs:77-125 (distance in f64, (distance,
// index) order) and is reproduced bit-for-bit:
//
//   1. prepare   rows -> fp16 (or bf16) operands Q (cosine: unit rows * 64; L2: rows * 2^e), with the
//                exact per-row rounding residual |delta_i| and operand norm |q_i| measured in f64.
//   2. screen    S = Q Q^T on the tensor cores: TMA (SWIZZLE_128B) -> shared memory ->
//                tcgen05.mma kind::f16 (M=128, N=256, K=16 per instruction) -> fp32 accumulators in
//                TMEM (2 x 256 columns, double buffered) -> tcgen05.ld in the epilogue warps.  One
//                epilogue thread owns one query row: a register threshold filters the 32 values of a
//                chunk; survivors are appended to the row's candidate buffer (HBM/L2); a full buffer
//                is pruned to its k' best by a warp-cooperative radix select, tightening the threshold.
//   3. rescore   every surviving candidate gets the exact f64 distance with the reference's
//                arithmetic (left-fold, no FMA) and the exact top-k is selected by (distance, index).
//   4. certify   every candidate the screen dropped had key <= thr_i.  With
//                  |S~_ij - q_i.q_j| <= gamma |q_i||q_j|            (fp32 accumulation in the tensor core)
//                  |q_i.q_j - s^2 x^_i.x^_j| <= |delta_i||q_j| + s|delta_j|   (operand rounding, Cauchy-Schwarz)
//                a lower bound LB_i on the true distance of any dropped candidate follows; row i is
//                certified iff its exact k-th distance is < LB_i.  Uncertified rows (ties, duplicates,
//                zero rows, margins too tight for fp16) are recomputed by the exact f64 brute force.
//   => neighbour sets and distances equal the brute-force reference whatever the screen precision.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"
#include "topk_list.cuh"

namespace {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2, B_STAGE_BYTES = BN * BK * 2;
constexpr int SCREEN_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr double COS_SCALE = 64.0;   // unit rows are scaled by 2^6 so fp16 stays in its normal range
constexpr uint32_t MAX_CAP = 256;

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done, addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants ----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of one CTA's share of a pair operand; the transaction bytes complete on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at this shared offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 in bits [0,14), SBO (8 rows * 128 B = 1024 B) >> 4 in [32,46), version 1 in [46,48),
// layout SWIZZLE_128B (2) in [61,64).  LBO is unused for a single 128-byte swizzle atom along K.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t f32_sortable(float v) {
    uint32_t u = __float_as_uint(v);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sortable_f32(uint32_t k) { return __uint_as_float((k >> 31) ? (k & 0x7FFFFFFFu) : ~k); }

// ---- prepare ----------------------------------------------------------------------------------
__global__ void absmax_kernel(const double* __restrict__ x, uint64_t n, unsigned long long* out) {
    double m = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = fmax(m, fabs(x[i]));
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));  // non-negative doubles order as integers
}

// One warp per row of [row0, row0 + nrows).  aux[3 i + {0, 1, 2}] = |q_i|, |delta_i|, |q_i|^2 (f64).  gmax[0] = max |q|, gmax[1] = max |delta|.
template <bool BF16>
__global__ void prepare_kernel(const double* __restrict__ x, const double* __restrict__ norms, uint64_t row0, uint64_t nrows, uint32_t kd,
                               uint32_t kpad, int cosine, double scale, uint16_t* __restrict__ q, float* __restrict__ nq32,
                               double* __restrict__ aux, unsigned long long* __restrict__ gmax) {
    const int lane = threadIdx.x & 31;
    const uint64_t li = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (li >= nrows) return;
    const uint64_t i = row0 + li;
    double mul = scale;
    if (cosine) { double nrm = norms[i]; mul = nrm > 0.0 ? scale / nrm : 0.0; }
    if (!isfinite(mul)) mul = 0.0;
    double sq = 0.0, sd = 0.0;
    for (uint32_t d = lane; d < kpad; d += 32) {
        double t = d < kd ? x[i * kd + d] * mul : 0.0;
        double qv;
        uint16_t bits;
        if (BF16) { __nv_bfloat16 h = __double2bfloat16(t); qv = (double)__bfloat162float(h); bits = __bfloat16_as_ushort(h); }
        else {
            __half h = __double2half(t);
            qv = (double)__half2float(h);
            if (fabs(qv) < 6.103515625e-05) { qv = 0.0; bits = 0; }  // flush fp16 subnormals ourselves: delta stays exact
            else bits = __half_as_ushort(h);
        }
        if (!isfinite(qv)) { qv = 0.0; bits = 0; }  // overflowed element: dropped from the operand, fully charged to delta
        q[i * kpad + d] = bits;
        double dl = qv - t;
        sq += qv * qv; sd += dl * dl;
    }
    for (int o = 16; o; o >>= 1) { sq += __shfl_xor_sync(FULL, sq, o); sd += __shfl_xor_sync(FULL, sd, o); }
    if (lane == 0) {
        double qn = sqrt(sq) * (1.0 + 1e-12), dn = sqrt(sd) * (1.0 + 1e-12);
        if (!isfinite(dn)) dn = INFINITY;
        aux[3 * i] = qn; aux[3 * i + 1] = dn; aux[3 * i + 2] = sq;
        nq32[i] = (float)sq;
        atomicMax(&gmax[0], (unsigned long long)__double_as_longlong(qn));
        atomicMax(&gmax[1], (unsigned long long)__double_as_longlong(dn));
    }
}

// ---- the screen -------------------------------------------------------------------------------
struct ScreenArgs {
    uint64_t n_rows;        // corpus rows M (columns of S beyond it are padding)
    uint64_t q_begin, nq;   // query rows [q_begin, q_begin + nq)
    uint32_t kblocks;       // kpad / 64
    uint32_t tiles_total, tiles_per_split, n_splits;
    uint32_t idesc;
    uint32_t kprime, cap;
    const float* nq32;      // |q_j|^2 as f32 (L2 keys)
    uint2* buf;  // candidates as (key bits, corpus row) pairs: [(row_local * n_splits + split) * cap]
    uint32_t* out_cnt; float* out_thr;  // [row_local * n_splits + split]
    float* dump;            // DUMP mode: 128 x 256 accumulators of the CTA's first tile
};

// Warp-cooperative prune of the candidate buffers of the lanes in `need`: keep (at least) the k' largest keys
// and raise the row's threshold to the largest key dropped.
//   fast path  the keys are quantised to 8 bits over the buffer's [min, max] and the cut is found by an
//              8-step binary search on ballot counts: keeps k' plus the few entries sharing the cut's bin.
//              Quantisation is monotone, so every kept key is strictly greater than every dropped key.
//   exact path MSB-first radix select of the k'-th largest key (32 steps), keeps exactly k'; taken when the fast cut
//              would drop nothing (all keys equal, or one outlier stretching the range) or would keep more than
//              k' + 16 (many equal keys in the cut's bin, e.g. exact duplicates: cap >= k' + 64 guarantees 32 free
//              slots after every prune only if the prune keeps at most k' + 32).
template <int E>
__device__ __forceinline__ void prune_rows(uint32_t need, uint2* my_buf, uint32_t& cnt, float& thr,
                                           uint32_t kprime, int lane) {
    while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        uint2* bp = reinterpret_cast<uint2*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(my_buf), L));
        const uint32_t n = __shfl_sync(FULL, cnt, L);
        float kf[E];
        uint32_t idx[E];
        float mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < E; ++u) {
            const uint32_t e = lane + 32 * u;
            const bool v = e < n;
            const uint2 ent = v ? __ldcg(bp + e) : make_uint2(0xFF800000u, 0u);   // -inf
            kf[u] = __uint_as_float(ent.x);
            idx[u] = ent.y;
            if (v) mn = fminf(mn, kf[u]);
            mx = fmaxf(mx, kf[u]);
        }
        mn = sortable_f32(__reduce_min_sync(FULL, f32_sortable(mn)));   // REDUX on the order-preserving integer image
        mx = sortable_f32(__reduce_max_sync(FULL, f32_sortable(mx)));
        bool keep[E];
        uint32_t kept = 0;
        float new_thr = -INFINITY;
        {
            const float scale = 255.0f / (mx - mn);   // inf / NaN when all keys are equal: the cut below then drops nothing
            uint32_t q[E];
#pragma unroll
            for (int u = 0; u < E; ++u) {
                float t = (kf[u] - mn) * scale;
                q[u] = (uint32_t)(lane + 32 * u) < n ? (uint32_t)fminf(fmaxf(t, 0.0f), 255.0f) : 0u;  // NaN -> 0
            }
            uint32_t cut = 0;
#pragma unroll
            for (int b = 7; b >= 0; --b) {
                const uint32_t trial = cut | (1u << b);
                uint32_t c = 0;
#pragma unroll
                for (int u = 0; u < E; ++u) c += __popc(__ballot_sync(FULL, q[u] >= trial));   // invalid entries hold q = 0 < trial
                if (c >= kprime) cut = trial;
            }
            float dropped_max = -INFINITY;
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const bool v = (uint32_t)(lane + 32 * u) < n;
                keep[u] = v && q[u] >= cut;
                kept += __popc(__ballot_sync(FULL, keep[u]));
                if (v && !keep[u]) dropped_max = fmaxf(dropped_max, kf[u]);
            }
            dropped_max = sortable_f32(__reduce_max_sync(FULL, f32_sortable(dropped_max)));
            new_thr = dropped_max;
        }
        if (kept == n || kept > kprime + 16) {
            // exact path: k'-th largest key by MSB-first radix select.  Also taken when many keys share the cut's bin
            // (a cluster of exact duplicates): the buffer must come out of a prune with room for a full chunk of appends.
            uint32_t key[E];
#pragma unroll
            for (int u = 0; u < E; ++u) key[u] = (uint32_t)(lane + 32 * u) < n ? f32_sortable(kf[u]) : 0u;  // 0 sorts below every real key
            uint32_t prefix = 0, want = kprime;
#pragma unroll 1
            for (int b = 31; b >= 0; --b) {
                const uint32_t bit = 1u << b, hi = b == 31 ? 0u : ~((bit << 1) - 1u);
                uint32_t c = 0;
#pragma unroll
                for (int u = 0; u < E; ++u) c += __popc(__ballot_sync(FULL, (key[u] & hi) == (prefix & hi) && (key[u] & bit)));
                if (c >= want) prefix |= bit; else want -= c;
            }
            // keep key > T and `want` of the entries equal to T
            uint32_t eq_seen = 0;
#pragma unroll
            for (int u = 0; u < E; ++u) {
                bool gt = key[u] > prefix, eq = key[u] == prefix && (uint32_t)(lane + 32 * u) < n;
                uint32_t beq = __ballot_sync(FULL, eq);
                uint32_t my_eq_rank = eq_seen + __popc(beq & ((1u << lane) - 1u));
                keep[u] = gt || (eq && my_eq_rank < want);
                eq_seen += __popc(beq);
            }
            new_thr = sortable_f32(prefix);
        }
        uint32_t base = 0;
        uint32_t pos[E];
#pragma unroll
        for (int u = 0; u < E; ++u) {
            uint32_t bk = __ballot_sync(FULL, keep[u]);
            pos[u] = base + __popc(bk & ((1u << lane) - 1u));
            base += __popc(bk);
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < E; ++u)
            if (keep[u]) __stcg(bp + pos[u], make_uint2(__float_as_uint(kf[u]), idx[u]));
        __syncwarp();
        if (lane == L) { cnt = base; thr = fmaxf(thr, new_thr); }
    }
}

// Cycle counters of the epilogue, compiled in only with -DSFB_SCREEN_PROFILE (they cost ~50 ms at C2 even when idle).
// SFB_SCREEN_DBG=4: [0] cycles in the slow (hit) path, [1] slow-path entries, [2] cycles in prunes, [3] prune calls,
// [4] rows pruned, [5] cycles waiting for a full accumulator, [6] tiles -- summed over warp 2 of every CTA
__device__ unsigned long long g_screen_dbg[8];
__device__ int g_prof_on = 0;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // FMNMX3
    return d;
}

// Filters one 32-column chunk of a thread's query row against its threshold.  Hits are rare once the
// threshold has tightened (a few percent of chunks), so the common case is a branch-free FMNMX3 tree
// (20 instructions) and one warp vote; per-element compares run only inside groups of 4 whose max passes.
template <bool L2>
__device__ __forceinline__ void filter_chunk(const uint32_t (&v)[32], const float* __restrict__ nqv, uint32_t col0, uint64_t n_rows,
                                             bool row_valid, uint2* my_buf, uint32_t& cnt, float& thr,
                                             uint32_t cap, uint32_t kprime, int lane) {
    float key[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        key[c] = __uint_as_float(v[c]);
        if (L2) key[c] = fmaf(2.0f, key[c], -nqv[c]);  // -(|q_j|^2 - 2 q_i.q_j): larger is nearer
    }
    float g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = fmaxf(fmax3(key[4 * q], key[4 * q + 1], key[4 * q + 2]), key[4 * q + 3]);
    const float m = fmax3(fmax3(g[0], g[1], g[2]), fmax3(g[3], g[4], g[5]), fmaxf(g[6], g[7]));
    const bool hit = m > thr;
    if (!__any_sync(FULL, hit)) return;
#ifdef SFB_SCREEN_PROFILE
    const bool prof = g_prof_on && lane == 0 && (threadIdx.x >> 5) == 2;
    long long t_in = 0;
    if (prof) t_in = clock64();
#endif
    // (Measured alternatives: eight static per-group tests with their own append code, 684 vs 656 ms for the kernel; a
    // warp-uniform variant -- one vote per group of 4, predicated appends -- 705 vs 658 ms.)
    if (hit && row_valid) {
        // The mask of groups that hold a hit is built branch-free; per set bit ONE dispatch picks that group's four
        // keys (lanes with different groups diverge only over four moves) and one shared body appends the hits.
        const uint32_t n_rows32 = (uint32_t)n_rows;   // node indices are 32-bit
        uint32_t gm = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) gm |= (g[q] > thr ? 1u : 0u) << q;
        do {
            const int q = __ffs(gm) - 1;
            gm &= gm - 1;
            const uint32_t jb = col0 + 4u * (uint32_t)q;
            float k0, k1, k2, k3;
#ifdef SFB_SELECT_BRANCHLESS
            k0 = key[28]; k1 = key[29]; k2 = key[30]; k3 = key[31];
#pragma unroll
            for (int qq = 6; qq >= 0; --qq) {
                const bool sel = q == qq;
                k0 = sel ? key[4 * qq] : k0; k1 = sel ? key[4 * qq + 1] : k1; k2 = sel ? key[4 * qq + 2] : k2; k3 = sel ? key[4 * qq + 3] : k3;
            }
            if (false)
#endif
            switch (q) {
                case 0: k0 = key[0]; k1 = key[1]; k2 = key[2]; k3 = key[3]; break;
                case 1: k0 = key[4]; k1 = key[5]; k2 = key[6]; k3 = key[7]; break;
                case 2: k0 = key[8]; k1 = key[9]; k2 = key[10]; k3 = key[11]; break;
                case 3: k0 = key[12]; k1 = key[13]; k2 = key[14]; k3 = key[15]; break;
                case 4: k0 = key[16]; k1 = key[17]; k2 = key[18]; k3 = key[19]; break;
                case 5: k0 = key[20]; k1 = key[21]; k2 = key[22]; k3 = key[23]; break;
                case 6: k0 = key[24]; k1 = key[25]; k2 = key[26]; k3 = key[27]; break;
                default: k0 = key[28]; k1 = key[29]; k2 = key[30]; k3 = key[31]; break;
            }
            if (k0 > thr && jb < n_rows32) { __stcg(my_buf + cnt, make_uint2(__float_as_uint(k0), jb)); ++cnt; }
            if (k1 > thr && jb + 1 < n_rows32) { __stcg(my_buf + cnt, make_uint2(__float_as_uint(k1), jb + 1)); ++cnt; }
            if (k2 > thr && jb + 2 < n_rows32) { __stcg(my_buf + cnt, make_uint2(__float_as_uint(k2), jb + 2)); ++cnt; }
            if (k3 > thr && jb + 3 < n_rows32) { __stcg(my_buf + cnt, make_uint2(__float_as_uint(k3), jb + 3)); ++cnt; }
        } while (gm);
    }
    // a chunk appends at most 32 entries: prune whenever fewer than 32 slots remain
    const uint32_t need = __ballot_sync(FULL, cnt + 32 > cap);
    if (need) {
        __syncwarp();
#ifdef SFB_SCREEN_PROFILE
        long long t_p = 0;
        if (prof) t_p = clock64();
#endif
        if (cap <= 128) prune_rows<4>(need, my_buf, cnt, thr, kprime, lane);
        else prune_rows<8>(need, my_buf, cnt, thr, kprime, lane);
#ifdef SFB_SCREEN_PROFILE
        if (prof) { atomicAdd(&g_screen_dbg[2], (unsigned long long)(clock64() - t_p)); atomicAdd(&g_screen_dbg[3], 1ull); atomicAdd(&g_screen_dbg[4], (unsigned long long)__popc(need)); }
#endif
    }
#ifdef SFB_SCREEN_PROFILE
    if (prof) { atomicAdd(&g_screen_dbg[0], (unsigned long long)(clock64() - t_in)); atomicAdd(&g_screen_dbg[1], 1ull); }
#endif
}

// ---- eight epilogue warps on one set of row buffers (experiment) ------------------------------------------------
// Warps w and w + 4 read the same 32 TMEM lanes (same query rows), four 32-column chunks each.  A row's count and
// threshold live in shared memory; slots are reserved with atomicAdd; after every chunk the pair meets at a named
// barrier that also ORs "some row is within 64 slots of its capacity": only then the rows to prune are agreed on
// (from counts that no longer move), split between the two warps, pruned, and written back between two more barriers.
__device__ __forceinline__ bool pair_bar_or(int id, bool q) {
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %2, 0;\n\tbarrier.cta.red.or.pred p, %1, 64, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"(id), "r"((uint32_t)q) : "memory");
    return r != 0;
}
__device__ __forceinline__ void pair_bar(int id) { asm volatile("barrier.cta.sync %0, 64;" ::"r"(id) : "memory"); }

template <bool L2>
__device__ __forceinline__ void filter_chunk_shared(const uint32_t (&v)[32], const float* __restrict__ nqv, uint32_t col0, uint64_t n_rows,
                                                    bool row_valid, uint2* my_buf, uint32_t* s_cnt_row, float* s_thr_row,
                                                    uint32_t cap, uint32_t kprime, int lane, int group, int bar_id) {
    float key[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        key[c] = __uint_as_float(v[c]);
        if (L2) key[c] = fmaf(2.0f, key[c], -nqv[c]);
    }
    float g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = fmaxf(fmax3(key[4 * q], key[4 * q + 1], key[4 * q + 2]), key[4 * q + 3]);
    const float m = fmax3(fmax3(g[0], g[1], g[2]), fmax3(g[3], g[4], g[5]), fmaxf(g[6], g[7]));
    const float thr = *s_thr_row;
    const bool hit = m > thr;
    if (__any_sync(FULL, hit)) {
        if (hit && row_valid) {
            const uint32_t n_rows32 = (uint32_t)n_rows;
            uint32_t gm = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) gm |= (g[q] > thr ? 1u : 0u) << q;
            do {
                const int q = __ffs(gm) - 1;
                gm &= gm - 1;
                const uint32_t jb = col0 + 4u * (uint32_t)q;
                float k0, k1, k2, k3;
                switch (q) {
                    case 0: k0 = key[0]; k1 = key[1]; k2 = key[2]; k3 = key[3]; break;
                    case 1: k0 = key[4]; k1 = key[5]; k2 = key[6]; k3 = key[7]; break;
                    case 2: k0 = key[8]; k1 = key[9]; k2 = key[10]; k3 = key[11]; break;
                    case 3: k0 = key[12]; k1 = key[13]; k2 = key[14]; k3 = key[15]; break;
                    case 4: k0 = key[16]; k1 = key[17]; k2 = key[18]; k3 = key[19]; break;
                    case 5: k0 = key[20]; k1 = key[21]; k2 = key[22]; k3 = key[23]; break;
                    case 6: k0 = key[24]; k1 = key[25]; k2 = key[26]; k3 = key[27]; break;
                    default: k0 = key[28]; k1 = key[29]; k2 = key[30]; k3 = key[31]; break;
                }
                const bool h0 = k0 > thr && jb < n_rows32, h1 = k1 > thr && jb + 1 < n_rows32;
                const bool h2 = k2 > thr && jb + 2 < n_rows32, h3 = k3 > thr && jb + 3 < n_rows32;
                const uint32_t nh = (uint32_t)h0 + (uint32_t)h1 + (uint32_t)h2 + (uint32_t)h3;
                if (nh) {
                    uint32_t pos = atomicAdd(s_cnt_row, nh);   // both warps of the pair append to this row
                    if (h0) __stcg(my_buf + pos++, make_uint2(__float_as_uint(k0), jb));
                    if (h1) __stcg(my_buf + pos++, make_uint2(__float_as_uint(k1), jb + 1));
                    if (h2) __stcg(my_buf + pos++, make_uint2(__float_as_uint(k2), jb + 2));
                    if (h3) __stcg(my_buf + pos++, make_uint2(__float_as_uint(k3), jb + 3));
                }
            } while (gm);
        }
    }
    __syncwarp();
    // every chunk: meet the partner warp; the later of the two reads of a row's count sees both warps' appends
    const bool near_full = *s_cnt_row + 64 > cap;
    if (pair_bar_or(bar_id, near_full)) {
        // counts are frozen now (nobody appends before the last barrier below)
        uint32_t cnt = *s_cnt_row;
        float thr2 = *s_thr_row;
        const uint32_t need = __ballot_sync(FULL, cnt + 64 > cap);
        uint32_t mine = 0, nm = need;
        int turn = 0;
        while (nm) { const int L = __ffs(nm) - 1; nm &= nm - 1; if ((turn++ & 1) == group) mine |= 1u << L; }
        if (mine) {
            if (cap <= 128) prune_rows<4>(mine, my_buf, cnt, thr2, kprime, lane);
            else prune_rows<8>(mine, my_buf, cnt, thr2, kprime, lane);
        }
        pair_bar(bar_id);   // the partner has read the old counts
        if ((mine >> lane) & 1u) { *s_cnt_row = cnt; *s_thr_row = thr2; }
        pair_bar(bar_id);   // new counts / thresholds visible to both
    }
}

template <bool L2, bool DUMP>
__global__ void __launch_bounds__(SCREEN_THREADS, 1)
knn_screen_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ScreenArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve: A stages | B stages | barriers | tmem ptr | nq tile (L2)
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
    uint64_t* tempty_bar = tfull_bar + 2;       // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_nq = reinterpret_cast<float*>(tmem_ptr + 4);  // [2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mb = blockIdx.x, split = blockIdx.y;
    const uint32_t t0 = split * a.tiles_per_split;
    uint32_t t1 = t0 + a.tiles_per_split;
    if (t1 > a.tiles_total) t1 = a.tiles_total;
    const uint32_t n_my_tiles = t1 > t0 ? t1 - t0 : 0;
    const int m_row0 = (int)(a.q_begin + (uint64_t)mb * BM);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = 0; t < n_my_tiles; ++t) {
                const int n0 = (int)((t0 + t) * BN);
                for (uint32_t kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
                    tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, (int)(kb * BK), m_row0, &full_bar[stage]);
                    tma_load_2d(sB + stage * B_STAGE_BYTES, &tmB, (int)(kb * BK), n0, &full_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = 0; t < n_my_tiles; ++t) {
                const uint32_t as = t & 1, aphase = (t >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (uint32_t kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES), b_addr = smem_u32(sB + stage * B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_f16(tmem_d, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), a.idesc, (kb | (uint32_t)k) != 0u);
                    umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
                    if (kb + 1 == a.kblocks) umma_commit(&tfull_bar[as]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: thread = query row =====
        const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31 are this warp's
        const int row_in_tile = quarter * 32 + lane;
        const uint64_t row_local = (uint64_t)mb * BM + row_in_tile;
        const bool row_valid = row_local < a.nq;
        const size_t slot = (size_t)(row_valid ? row_local : 0) * a.n_splits + split;
        uint2* my_buf = a.buf + slot * a.cap;
        uint32_t cnt = 0;
        float thr = -INFINITY;
        const int etid = threadIdx.x - 64;            // 0..127 among the epilogue threads
        for (uint32_t t = 0; t < n_my_tiles; ++t) {
            const uint32_t as = t & 1, aphase = (t >> 1) & 1;
            const uint32_t n0 = (t0 + t) * BN;
            if (L2) {
                // |q_j|^2 of this tile's 256 columns, double buffered with the accumulator stage
                float* dst = s_nq + as * BN;
                for (int c = etid; c < BN; c += 128) dst[c] = (uint64_t)n0 + c < a.n_rows ? __ldg(a.nq32 + n0 + c) : INFINITY;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN;
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                uint32_t v[32];
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
                if (DUMP) {
                    if (t == 0)
                        for (int c = 0; c < 32; ++c) a.dump[(size_t)row_in_tile * BN + ch * 32 + c] = __uint_as_float(v[c]);
                    continue;
                }
                filter_chunk<L2>(v, s_nq + as * BN + ch * 32, n0 + ch * 32, a.n_rows, row_valid, my_buf, cnt, thr, a.cap, a.kprime, lane);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        if (!DUMP && row_valid) { a.out_cnt[slot] = cnt; a.out_thr[slot] = thr; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- the screen, CTA-pair version -------------------------------------------------------------
// The single-CTA kernel above re-streams both operands for every 128 x 256 tile: 48 KB of L2 -> shared
// traffic per 512 tensor-core cycles and SM, more than twice what the L2 can deliver to 148 SMs, and the
// fp16 corpus (0.77 GB at 1M x 384) falls out of the 126 MB L2 between two sweeps.  This version
//   * pairs two CTAs (cluster of 2, tcgen05.mma.cta_group::2, M = 256): each CTA stages only HALF of
//     every corpus tile (128 of the 256 rows) and the pair's tensor cores read both halves;
//   * keeps the CTA's 128 query rows (all of K) resident in shared memory for a whole work item, so only
//     the corpus streams: 16 KB per k-block and CTA -- a third of the single-CTA traffic;
//   * walks the corpus in L2-sized chunks: every pair sweeps chunk c for all of its query blocks before
//     any pair touches chunk c + 1 (persistent CTAs, static schedule), so corpus tiles are read from HBM
//     once per chunk instead of once per query block.  The per-row state (threshold, candidate count)
//     lives in HBM between chunks; the candidate buffers already do.
struct Screen2Args {
    uint64_t n_rows;
    uint64_t q_begin, nq;   // q_begin: first row of the QUERY operand (tmA) this launch covers
    uint32_t kblocks, stages;
    uint32_t resident;      // k-blocks of the query rows that stay in shared memory for a whole work item (all of them when K <= 512)
    uint32_t stage_bytes;   // 16 KB (corpus half-tile) when every k-block is resident, else 32 KB (+ the streamed query slab)
    uint32_t tiles_total, tiles_per_split, n_splits, chunk_tiles, n_chunks;
    uint32_t n_mb2;         // 256-row query blocks
    uint32_t idesc;
    uint32_t kprime, cap;
    const float* nq32;
    uint2* buf;
    uint32_t* out_cnt; float* out_thr;
    uint32_t dbg;           // SFB_SCREEN_DBG (timing experiments, results invalid): 1 = tcgen05.ld only, 2 = no epilogue work
};

constexpr int P_BM = 128, P_BNH = 128, P_SLAB_BYTES = 128 * BK * 2;  // 16 KB: one k-block of 128 rows

constexpr int SCREEN_THREADS8 = 320;  // + warps 6-9: the second epilogue group
template <bool L2, bool EW8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EW8 ? SCREEN_THREADS8 : SCREEN_THREADS, 1)
knn_screen_pair_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmA, Screen2Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve: A slabs [kblocks] | B stages [stages] | barriers | tmem ptr | nq tiles (L2)
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)a.resident * P_SLAB_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + (size_t)a.stages * a.stage_bytes);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tfull_bar = empty_bar + 8;     // [2]
    uint64_t* tempty_bar = tfull_bar + 2;    // [2]
    uint64_t* afull_bar = tempty_bar + 2;    // [1]
    uint64_t* aempty_bar = afull_bar + 1;    // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aempty_bar + 1);
    float* s_nq = reinterpret_cast<float*>(tmem_ptr + 4);  // [2][BN]
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_nq + 2 * BN);   // [128] (EW8)
    float* s_thr = reinterpret_cast<float*>(s_cnt + 128);              // [128] (EW8)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const uint32_t n_slots = a.n_mb2 * a.n_splits;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        for (uint32_t s = 0; s < a.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EW8 ? 16 : 8); }  // epilogue warps x 2 CTAs
        mbar_init(afull_bar, 1); mbar_init(aempty_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // Work items in schedule order; every role of both CTAs walks the same sequence.
#define SFB_FOR_EACH_ITEM                                                                              \
    for (uint32_t chunk = 0; chunk < a.n_chunks; ++chunk)                                              \
        for (uint32_t s2 = pair; s2 < n_slots; s2 += n_pairs)
#define SFB_ITEM_RANGE                                                                                 \
    const uint32_t mb2 = s2 / a.n_splits, split = s2 - mb2 * a.n_splits;                               \
    const uint32_t t_lo = split * a.tiles_per_split + chunk * a.chunk_tiles;                           \
    uint32_t t_hi = t_lo + a.chunk_tiles;                                                              \
    { uint32_t e = (split + 1) * a.tiles_per_split; if (e > a.tiles_total) e = a.tiles_total; if (t_hi > e) t_hi = e; } \
    if (t_lo >= t_hi) continue;

    if (warp == 0) {
        // ===== TMA producer (one thread per CTA) =====
        if (lane == 0) {
            const uint32_t l_afull = mapa_u32(afull_bar, 0);
            uint32_t stage = 0, phase = 0, iphase = 0;
            SFB_FOR_EACH_ITEM {
                SFB_ITEM_RANGE
                const int row0 = (int)(a.q_begin + (uint64_t)mb2 * 256 + rank * P_BM);
                mbar_wait(aempty_bar, iphase ^ 1);   // the previous item's MMAs have retired
                if (leader) mbar_expect_tx(afull_bar, 2u * a.resident * P_SLAB_BYTES);
                for (uint32_t kb = 0; kb < a.resident; ++kb)
                    tma_load_2d_pair(sA + (size_t)kb * P_SLAB_BYTES, &tmA, (int)(kb * BK), row0, l_afull);
                iphase ^= 1;
                for (uint32_t t = t_lo; t < t_hi; ++t) {
                    const int n0 = (int)(t * BN + rank * P_BNH);
                    for (uint32_t kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        const bool stream_a = kb >= a.resident;   // this k-block of the query rows rides along with the corpus
                        if (leader) mbar_expect_tx(&full_bar[stage], stream_a ? 4u * P_SLAB_BYTES : 2u * P_SLAB_BYTES);
                        const uint32_t l_full = mapa_u32(&full_bar[stage], 0);
                        uint8_t* st = sB + (size_t)stage * a.stage_bytes;
                        tma_load_2d_pair(st, &tm, (int)(kb * BK), n0, l_full);
                        if (stream_a) tma_load_2d_pair(st + P_SLAB_BYTES, &tmA, (int)(kb * BK), row0, l_full);
                        if (++stage == a.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores =====
        if (leader && lane == 0) {
            uint32_t stage = 0, phase = 0, iphase = 0, tc = 0;
            SFB_FOR_EACH_ITEM {
                SFB_ITEM_RANGE
                mbar_wait(afull_bar, iphase);
                iphase ^= 1;
                tc_fence_after();
                const uint32_t a_base = smem_u32(sA);
                for (uint32_t t = t_lo; t < t_hi; ++t, ++tc) {
                    const uint32_t as = tc & 1, aphase = (tc >> 1) & 1;
                    mbar_wait(&tempty_bar[as], aphase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * BN;
                    for (uint32_t kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(sB + (size_t)stage * a.stage_bytes);
                        const uint32_t a_addr = kb < a.resident ? a_base + kb * P_SLAB_BYTES : b_addr + P_SLAB_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16_pair(tmem_d, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), a.idesc, (kb | (uint32_t)k) != 0u);
                        umma_commit_pair(&empty_bar[stage]);
                        if (kb + 1 == a.kblocks) umma_commit_pair(&tfull_bar[as]);
                        if (++stage == a.stages) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit_pair(aempty_bar);   // A may be overwritten once every MMA of this item has retired
            }
        }
    } else if (EW8) {
        // ===== epilogue, two warps per 32 query rows: group 0 takes columns 0-127 of a tile, group 1 columns 128-255 =====
        const int quarter = warp & 3, group = (warp - 2) >> 2, bar_id = 2 + quarter;
        const int row_in_tile = quarter * 32 + lane;
        const int etid = threadIdx.x - 64;
        const uint32_t l_tempty0 = mapa_u32(&tempty_bar[0], 0), l_tempty1 = mapa_u32(&tempty_bar[1], 0);
        uint32_t tc = 0;
        SFB_FOR_EACH_ITEM {
            SFB_ITEM_RANGE
            const uint64_t row_local = (uint64_t)mb2 * 256 + rank * P_BM + row_in_tile;
            const bool row_valid = row_local < a.nq;
            const size_t slot = (size_t)(row_valid ? row_local : 0) * a.n_splits + split;
            uint2* my_buf = a.buf + slot * a.cap;
            if (group == 0) {
                uint32_t c0 = 0; float t0 = -INFINITY;
                if (chunk != 0 && row_valid) { c0 = a.out_cnt[slot]; t0 = a.out_thr[slot]; }
                if (a.dbg == 3) t0 = INFINITY;
                s_cnt[row_in_tile] = c0; s_thr[row_in_tile] = t0;
            }
            pair_bar(bar_id);
            for (uint32_t t = t_lo; t < t_hi; ++t, ++tc) {
                const uint32_t as = tc & 1, aphase = (tc >> 1) & 1;
                const uint32_t n0 = t * BN;
                if (L2) {
                    float* dst = s_nq + as * BN;
                    for (int c = etid; c < BN; c += 256) dst[c] = (uint64_t)n0 + c < a.n_rows ? __ldg(a.nq32 + n0 + c) : INFINITY;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                mbar_wait(&tfull_bar[as], aphase);
                tc_fence_after();
                const uint32_t half = (uint32_t)group * (BN / 2);
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + half;
                const float* nqt = s_nq + as * BN + half;
                uint32_t va[32], vb[32];
                if (a.dbg == 0 || a.dbg >= 3) {
                    tmem_ld32(taddr, va);
#pragma unroll 1
                    for (int ch = 0; ch < BN / 64; ch += 2) {
                        tmem_ld_wait();
                        tmem_ld32(taddr + (ch + 1) * 32, vb);
                        filter_chunk_shared<L2>(va, nqt + ch * 32, n0 + half + ch * 32, a.n_rows, row_valid, my_buf, s_cnt + row_in_tile, s_thr + row_in_tile,
                                                a.cap, a.kprime, lane, group, bar_id);
                        tmem_ld_wait();
                        if (ch + 2 < BN / 64) tmem_ld32(taddr + (ch + 2) * 32, va);
                        filter_chunk_shared<L2>(vb, nqt + (ch + 1) * 32, n0 + half + (ch + 1) * 32, a.n_rows, row_valid, my_buf, s_cnt + row_in_tile,
                                                s_thr + row_in_tile, a.cap, a.kprime, lane, group, bar_id);
                    }
                } else if (a.dbg == 1) {
                    uint32_t acc = 0;
#pragma unroll 1
                    for (int ch = 0; ch < BN / 64; ++ch) {
                        tmem_ld32(taddr + ch * 32, va);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc ^= va[c];
                    }
                    if (acc == 0x12345678u) s_thr[row_in_tile] = 0.0f;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(as ? l_tempty1 : l_tempty0);
            }
            pair_bar(bar_id);   // both warps are done with the item: the row state can be parked
            if (group == 0 && row_valid) {
                uint32_t c1 = s_cnt[row_in_tile]; float t1 = s_thr[row_in_tile];
                if (a.dbg == 3) { c1 = 0; t1 = -INFINITY; }
                a.out_cnt[slot] = c1; a.out_thr[slot] = t1;
            }
            pair_bar(bar_id);   // ... before group 0 re-initialises it for the next item
        }
    } else {
        // ===== epilogue: thread = query row =====
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        const int etid = threadIdx.x - 64;
        const uint32_t l_tempty0 = mapa_u32(&tempty_bar[0], 0), l_tempty1 = mapa_u32(&tempty_bar[1], 0);
        uint32_t tc = 0;
        SFB_FOR_EACH_ITEM {
            SFB_ITEM_RANGE
            const uint64_t row_local = (uint64_t)mb2 * 256 + rank * P_BM + row_in_tile;
            const bool row_valid = row_local < a.nq;
            const size_t slot = (size_t)(row_valid ? row_local : 0) * a.n_splits + split;
            uint2* my_buf = a.buf + slot * a.cap;
            uint32_t cnt = 0;
            float thr = -INFINITY;
            if (chunk != 0 && row_valid) { cnt = a.out_cnt[slot]; thr = a.out_thr[slot]; }
            if (a.dbg == 3) thr = INFINITY;   // timing experiment: the fast path only, nothing ever passes
            for (uint32_t t = t_lo; t < t_hi; ++t, ++tc) {
                const uint32_t as = tc & 1, aphase = (tc >> 1) & 1;
                const uint32_t n0 = t * BN;
                if (L2) {
                    float* dst = s_nq + as * BN;
                    for (int c = etid; c < BN; c += 128) dst[c] = (uint64_t)n0 + c < a.n_rows ? __ldg(a.nq32 + n0 + c) : INFINITY;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
#ifdef SFB_SCREEN_PROFILE
                const bool prof = a.dbg == 4 && lane == 0 && warp == 2;
                long long t_w = 0;
                if (prof) t_w = clock64();
#endif
                mbar_wait(&tfull_bar[as], aphase);
#ifdef SFB_SCREEN_PROFILE
                if (prof) { atomicAdd(&g_screen_dbg[5], (unsigned long long)(clock64() - t_w)); atomicAdd(&g_screen_dbg[6], 1ull); }
#endif
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN;
                const float* nqt = s_nq + as * BN;
                uint32_t va[32], vb[32];
                if (a.dbg == 0 || a.dbg >= 3) {
                    tmem_ld32(taddr, va);
#pragma unroll 1
                    for (int ch = 0; ch < BN / 32; ch += 2) {
                        tmem_ld_wait();
                        tmem_ld32(taddr + (ch + 1) * 32, vb);   // in flight while chunk ch is filtered
                        filter_chunk<L2>(va, nqt + ch * 32, n0 + ch * 32, a.n_rows, row_valid, my_buf, cnt, thr, a.cap, a.kprime, lane);
                        tmem_ld_wait();
                        if (ch + 2 < BN / 32) tmem_ld32(taddr + (ch + 2) * 32, va);
                        filter_chunk<L2>(vb, nqt + (ch + 1) * 32, n0 + (ch + 1) * 32, a.n_rows, row_valid, my_buf, cnt, thr, a.cap, a.kprime, lane);
                    }
                } else if (a.dbg == 1) {
                    uint32_t acc = 0;
#pragma unroll 1
                    for (int ch = 0; ch < BN / 32; ++ch) {
                        tmem_ld32(taddr + ch * 32, va);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc ^= va[c];
                    }
                    if (acc == 0x12345678u) thr = 0.0f;   // keeps the loads alive
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(as ? l_tempty1 : l_tempty0);
            }
            if (a.dbg == 3) { cnt = 0; thr = -INFINITY; }
            if (row_valid) { a.out_cnt[slot] = cnt; a.out_thr[slot] = thr; }
        }
    }
#undef SFB_FOR_EACH_ITEM
#undef SFB_ITEM_RANGE
    __syncwarp();         // the single-thread roles rejoin their warps before the aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();   // no CTA may leave while its peer can still signal its barriers or read its shared memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- rescore + certify ------------------------------------------------------------------------
struct RescoreArgs {
    const double* x; const double* norms; uint64_t m; uint32_t kd; int metric; uint32_t k; double eps;
    uint64_t q_begin, nq; uint32_t n_splits, cap;
    const uint2* buf; const uint32_t* cnt; const float* thr;
    const double* aux;      // |q|, |delta|, |q|^2
    double nmax, dmax, gamma, scale;
    uint32_t* out_idx; double* out_dist; uint32_t* out_cnt;
    uint32_t* fb_rows; uint32_t* fb_count;  // uncertified rows (global indices)
    double* max_margin;
    unsigned long long* n_cand;   // corpus rows rescored (a statistic: the rescore's traffic is n_cand x dims x 8 bytes)
    const uint32_t* qlist;  // null: query row rl is global row q_begin + rl; else global row qlist[rl] (re-screen of uncertified rows)
    uint64_t out_base;      // results of global row g go to output row g - out_base
};

// One warp per query row.
//   filter   the screen hands over k' .. cap candidates per row, but only those whose screen key is within
//            the error margin of the k-th best key can be among the exact k nearest: with S_k the k-th
//            largest key (self excluded) every candidate below S_k - 2 * margin (cosine; the L2 form goes
//            through the distance bounds) is provably farther than k others.  That cuts the rows gathered
//            in f64 -- the cost of this kernel -- from ~100 to ~k + a few.  Excluded candidates join the
//            screen's dropped set: the certificate below is evaluated against the larger threshold.
//   rescore  candidate rows are staged 32 at a time through a padded shared tile so global loads are
//            coalesced and every lane runs its candidate's left fold in dimension order (reference bits).
//   certify  row i is certified iff its exact k-th distance is below the proven lower bound on the
//            distance of everything that was not rescored; otherwise it goes to the exact fallback.
constexpr uint32_t RS_MAXC = 256;  // candidates a row can bring to the filter; beyond that all are rescored
// per-warp tile of one ring stage: NB candidate rows x 32 dimensions at row stride LD doubles, + the query chunk [32].
//   cp.async path  32 rows, stride 33 (conflict-free for lane = row)
//   bulk path      30 rows, stride 34: a row chunk is ONE 256-byte cp.async.bulk (16-byte aligned destination: even stride;
//                  lane = row reads are then 2-way conflicted, which the fold does not notice), 30 rows so that three CTAs of
//                  four warps fit an SM
//   CH = 16        (cp.async path) half the chunk: half a warp per row and copy instruction, stride 17, 4.4 KB per stage -- twice
//                  the resident warps for the same shared memory
template <bool BULK, int CH> struct RsGeom { static constexpr int LD = BULK ? 34 : CH + 1, NB = BULK ? 30 : 32, TILE = NB * LD + CH; };
__host__ __device__ constexpr size_t rs_warp_bytes(bool bulk, int ch, int nst, uint32_t k) {
    return ((size_t)nst * (bulk ? RsGeom<true, 32>::TILE : (ch == 16 ? RsGeom<false, 16>::TILE : RsGeom<false, 32>::TILE)) * sizeof(double) + (bulk ? (size_t)nst * 8 : 0) +
            (size_t)k * (sizeof(double) + sizeof(uint32_t)) + RS_MAXC * sizeof(uint32_t) + 15) & ~(size_t)15;   // 16-byte aligned per warp
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    const int n = valid ? 8 : 0;   // src-size 0: zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}

// NST: stages of the per-warp tile ring.  The gather is latency-bound: a warp waits for a chunk it issued one iteration
// earlier, ~3 us under load against ~0.4 us of folds, so the HBM rate is (bytes in flight) / latency -- 2.4 TB/s at C2 with two
// stages (one chunk in flight per warp, eight warps per SM).  Three stages keep two chunks in flight; they fit twice per SM
// up to k = 64 (the screen keys of the filter phase share the ring's memory), else the kernel runs with two.
//
// BULK: the row chunks come by cp.async.bulk (the TMA engine's linear copy: one 256-byte request per candidate row and chunk,
// issued by the lane that owns the row, completion counted in bytes on a per-stage mbarrier) instead of 8-byte cp.async, which
// costs the LSU one 8-byte element per lane and instruction: at C2 the kernel moved ~8 bytes per clock and SM whatever the ring
// depth, and only more resident warps made it faster.  Needs 16-byte aligned row chunks (even D, aligned base).
template <bool COS, int NST, bool BULK, int CH>
__global__ void __launch_bounds__(128, BULK ? 3 : (NST == 2 ? 4 : 2)) knn_rescore_kernel(RescoreArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(!BULK || CH == 32, "the bulk path moves 32-dimension chunks");
    constexpr int RS_LD = RsGeom<BULK, CH>::LD, RS_NB = RsGeom<BULK, CH>::NB, RS_TILE = RsGeom<BULK, CH>::TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // per warp: NST tiles | NST mbarriers (BULK) | list distances [k] | list indices [k] | candidate indices [RS_MAXC]
    unsigned char* wbase = smem_raw + (size_t)w * rs_warp_bytes(BULK, CH, NST, a.k);
    double* tile = reinterpret_cast<double*>(wbase);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tile + (size_t)NST * RS_TILE);
    double* ld = reinterpret_cast<double*>(bars + (BULK ? NST : 0));
    uint32_t* li = reinterpret_cast<uint32_t*>(ld + a.k);
    uint32_t* sidx = li + a.k;
    float* skey = reinterpret_cast<float*>(tile);         // filter phase only: RS_MAXC floats at the head of the (still idle) ring
    uint32_t phase = 0;                                   // BULK: parity of each stage's barrier, one bit per stage
    if (BULK) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }

    const uint64_t rl = (uint64_t)blockIdx.x * (blockDim.x >> 5) + w;
    if (rl >= a.nq) return;
    const uint32_t gi = a.qlist ? a.qlist[rl] : (uint32_t)(a.q_begin + rl);
    const uint64_t orow = (uint64_t)gi - a.out_base;
    const double* xi = a.x + (uint64_t)gi * a.kd;
    const double ni = COS ? a.norms[gi] : 0.0;
    const double qn = a.aux[3 * (size_t)gi], dn = a.aux[3 * (size_t)gi + 1], q2 = a.aux[3 * (size_t)gi + 2];
    const double s2 = a.scale * a.scale;
    // error model of the screen key (see the header comment of this file)
    const double cos_margin = ((a.gamma * qn + dn) * a.nmax + a.scale * a.dmax) * (1.0 + 1e-6) + s2 * 1e-9;  // key units
    const double eta_base = (1.2e-7 * a.nmax * a.nmax + (2.0 * a.gamma + 1.2e-7) * qn * a.nmax) * (1.0 + 1e-6);
    const double rho = (dn + a.dmax) * (1.0 + 1e-9);

    uint32_t total = 0;
    float thr = -INFINITY;
    for (uint32_t s = 0; s < a.n_splits; ++s) {
        const size_t slot = (size_t)rl * a.n_splits + s;
        total += a.cnt[slot];
        thr = fmaxf(thr, a.thr[slot]);
    }

    // ---- filter -------------------------------------------------------------------------------
    const bool use_list = total <= RS_MAXC;
    uint32_t n_list = 0;
    if (use_list) {
        for (uint32_t s = 0; s < a.n_splits; ++s) {
            const size_t slot = (size_t)rl * a.n_splits + s;
            const uint32_t n = a.cnt[slot];
            for (uint32_t b = 0; b < n; b += 32) {
                const bool v = b + lane < n;
                const uint2 ent = v ? __ldcg(a.buf + slot * a.cap + b + lane) : make_uint2(0xFF800000u, SFB_IDX_NONE);
                const uint32_t j = ent.y;
                const float kf = __uint_as_float(ent.x);
                const bool keep = v && j != gi;
                const uint32_t bal = __ballot_sync(FULL, keep);
                if (keep) { const uint32_t pos = n_list + __popc(bal & ((1u << lane) - 1u)); skey[pos] = kf; sidx[pos] = j; }
                n_list += __popc(bal);
            }
        }
        __syncwarp();
        if (n_list > a.k) {
            uint32_t key[RS_MAXC / 32];
#pragma unroll
            for (int u = 0; u < (int)(RS_MAXC / 32); ++u) { const uint32_t e = lane + 32 * u; key[u] = e < n_list ? f32_sortable(skey[e]) : 0u; }
            uint32_t prefix = 0, want = a.k;
            const int nreg = (int)((n_list + 31) / 32);   // registers that hold a key (warp-uniform): ~3 of the 8 at C2
#pragma unroll 1
            for (int b = 31; b >= 0; --b) {
                const uint32_t bit = 1u << b, hi = b == 31 ? 0u : ~((bit << 1) - 1u);
                uint32_t c = 0;
#pragma unroll
                for (int u = 0; u < (int)(RS_MAXC / 32); ++u)
                    if (u < nreg) c += __popc(__ballot_sync(FULL, (key[u] & hi) == (prefix & hi) && (key[u] & bit)));
                if (c >= want) prefix |= bit; else want -= c;
            }
            const double sk = (double)sortable_f32(prefix);   // k-th largest screen key among the other rows
            double tf;
            // The k rows with key >= sk have exact keys >= sk - margin, so the exact k-th key is >= sk - margin; a candidate
            // below sk - 2 margin has an exact key below sk - margin: it is farther than those k, and -- joining the dropped
            // set with its key + margin <= sk - margin -- it can never be what keeps the row from certifying.
            if (COS) tf = sk - 2.0 * (1.0 + 1e-6) * cos_margin;
            else {
                // the k rows with key >= sk are within T of row i (scaled units); exclude what is provably farther
                const double eta_k = eta_base + 2.4e-7 * (fabs(sk) + q2);
                const double d2k = q2 - sk + eta_k;
                const double T = (d2k > 0.0 ? sqrt(d2k) : 0.0) * (1.0 + 1e-9) + 2.0 * rho * (1.0 + 1e-6) + 1e-300;
                const double t2 = T * T * (1.0 + 1e-9);
                tf = q2 - t2 - (eta_base + 2.4e-7 * (q2 + t2)) * (1.0 + 1e-6);
            }
            const float thr_filter = isfinite(tf) ? __double2float_rd(tf) : -INFINITY;
            // compact in place (output position <= input position), remember the largest key excluded
            uint32_t out = 0;
            float ex_max = -INFINITY;
            for (uint32_t b = 0; b < n_list; b += 32) {
                const bool v = b + lane < n_list;
                const float kf = v ? skey[b + lane] : -INFINITY;
                const uint32_t j = v ? sidx[b + lane] : SFB_IDX_NONE;
                const bool keep = v && kf >= thr_filter;
                if (v && !keep) ex_max = fmaxf(ex_max, kf);
                const uint32_t bal = __ballot_sync(FULL, keep);
                __syncwarp();
                if (keep) sidx[out + __popc(bal & ((1u << lane) - 1u))] = j;
                out += __popc(bal);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) ex_max = fmaxf(ex_max, __shfl_xor_sync(FULL, ex_max, o));
            n_list = out;
            thr = fmaxf(thr, ex_max);   // the excluded candidates join the dropped set
            __syncwarp();
        }
    }

    // ---- exact rescore ------------------------------------------------------------------------
    uint32_t c = 0;
    auto run_batch = [&](uint32_t mine, uint32_t nb) {
        double acc = 0.0;
        // cp.async (8 bytes per lane: one 256-byte row segment per instruction), NST - 1 chunks in flight:
        // the gathers of chunks c + 1 .. c + NST - 1 overlap the folds of chunk c
        const uint32_t n_chunks = (a.kd + CH - 1) / CH;
        auto issue = [&](uint32_t ci) {   // chunk ci into slot ci % NST; cp.async: one commit group per call, empty past the end
            if (ci < n_chunks) {
                const uint32_t slot = ci % NST;
                double* tb = tile + slot * RS_TILE;
                const uint32_t d0 = ci * CH;
                if (BULK) {
                    const uint32_t bytes = (a.kd - d0 < 32 ? a.kd - d0 : 32) * 8;
                    // the slot's previous contents were read (or, in the filter phase, written) through the generic proxy
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (lane == 0) mbar_expect_tx(&bars[slot], (nb + 1) * bytes);
                    __syncwarp();
                    if (lane < (int)nb) bulk_g2s(&tb[lane * RS_LD], a.x + (uint64_t)mine * a.kd + d0, bytes, &bars[slot]);
                    if (lane == 31) bulk_g2s(&tb[RS_LD * RS_NB], xi + d0, bytes, &bars[slot]);
                } else if (CH == 32) {
                    const bool dv = d0 + lane < a.kd;
                    const uint32_t dd = dv ? d0 + lane : 0u;
                    for (uint32_t r = 0; r < nb; ++r) {
                        const uint32_t j = __shfl_sync(FULL, mine, r);
                        cp_async8(&tb[r * RS_LD + lane], a.x + (uint64_t)j * a.kd + dd, dv);
                    }
                    cp_async8(&tb[RS_LD * RS_NB + lane], xi + dd, dv);
                } else {
                    // half a warp per row: lanes 0-15 row r, lanes 16-31 row r + 1
                    const uint32_t hl = lane & 15u, hr = (uint32_t)lane >> 4;
                    const bool dv = d0 + hl < a.kd;
                    const uint32_t dd = dv ? d0 + hl : 0u;
                    for (uint32_t r = 0; r < nb; r += 2) {
                        const uint32_t rr = r + hr;
                        const uint32_t j = __shfl_sync(FULL, mine, rr & 31u);
                        if (rr < nb) cp_async8(&tb[rr * RS_LD + hl], a.x + (uint64_t)j * a.kd + dd, dv);
                    }
                    if (hr == 0) cp_async8(&tb[RS_LD * RS_NB + hl], xi + dd, dv);
                }
            }
            if (!BULK) asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int s = 0; s < NST - 1; ++s) issue((uint32_t)s);
        for (uint32_t ci = 0; ci < n_chunks; ++ci) {
            issue(ci + NST - 1);   // its slot was consumed in the previous iteration (__syncwarp at its end)
            if (BULK) {
                const uint32_t slot = ci % NST;
                mbar_wait(&bars[slot], (phase >> slot) & 1u);
                phase ^= 1u << slot;
            } else {
                asm volatile("cp.async.wait_group %0;" ::"n"(NST - 1) : "memory");
                __syncwarp();
            }
            const double* tb = tile + (ci % NST) * RS_TILE;
            const double* qc = tb + RS_LD * RS_NB;
            const uint32_t d0 = ci * CH, lim = a.kd - d0 < (uint32_t)CH ? a.kd - d0 : CH;
            if (lane < (int)nb) {
                for (uint32_t d = 0; d < lim; ++d) {
                    if (COS) acc = __dadd_rn(acc, __dmul_rn(qc[d], tb[lane * RS_LD + d]));
                    else { double t = __dadd_rn(qc[d], -tb[lane * RS_LD + d]); acc = __dadd_rn(acc, __dmul_rn(t, t)); }
                }
            }
            __syncwarp();   // the slot is refilled by the next iteration's issue
        }
        if (!BULK) asm volatile("cp.async.wait_group 0;" ::: "memory");
        double key = INFINITY;
        if (mine != SFB_IDX_NONE) {
            if (COS) {
                double denom = __dmul_rn(ni, a.norms[mine]), cosv = 0.0;
                if (denom > 1e-12) { cosv = __ddiv_rn(acc, denom); if (cosv < -1.0) cosv = -1.0; else if (cosv > 1.0) cosv = 1.0; }
                double rect = cosv > 0.0 ? cosv : 0.0;
                key = __dadd_rn(1.0, -rect);
            } else key = a.metric == SFB_METRIC_L2 ? __dsqrt_rn(acc) : acc;
        }
        double td = c == a.k ? ld[a.k - 1] : INFINITY;
        uint32_t ti = c == a.k ? li[a.k - 1] : SFB_IDX_NONE;
        bool pass = mine != SFB_IDX_NONE && mine != gi && key <= a.eps && topk_key_less(key, mine, td, ti);
        uint32_t bal = __ballot_sync(FULL, pass);
        if (c == 0) {
            // first batch into an empty list (the only batch of most rows): one 15-step bitonic network over the lanes in the
            // (distance, index) order instead of ~20 serial list insertions
            double sk_ = pass ? key : INFINITY;
            uint32_t si_ = pass ? mine : SFB_IDX_NONE;
#pragma unroll
            for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    const double ok = __shfl_xor_sync(FULL, sk_, j);
                    const uint32_t oi = __shfl_xor_sync(FULL, si_, j);
                    const bool want_min = ((lane & j) == 0) == ((lane & k2) == 0);
                    const bool swap = want_min ? topk_key_less(ok, oi, sk_, si_) : topk_key_less(sk_, si_, ok, oi);
                    if (swap) { sk_ = ok; si_ = oi; }
                }
            const uint32_t nv = __popc(bal);
            c = nv < a.k ? nv : a.k;
            if ((uint32_t)lane < c) { ld[lane] = sk_; li[lane] = si_; }
            __syncwarp();
            bal = 0;
        }
        while (bal) {
            int src = __ffs(bal) - 1; bal &= bal - 1;
            double kd_ = __shfl_sync(FULL, key, src);
            uint32_t jj = __shfl_sync(FULL, mine, src);
            warp_list_insert(ld, li, c, a.k, kd_, jj, lane);
        }
    };
    if (lane == 0) atomicAdd(a.n_cand, (unsigned long long)(use_list ? n_list : total));
    if (use_list) {
        for (uint32_t b = 0; b < n_list; b += RS_NB)
            run_batch(lane < RS_NB && b + lane < n_list ? sidx[b + lane] : SFB_IDX_NONE, n_list - b < (uint32_t)RS_NB ? n_list - b : RS_NB);
    } else {
        for (uint32_t s = 0; s < a.n_splits; ++s) {
            const size_t slot = (size_t)rl * a.n_splits + s;
            const uint32_t n = a.cnt[slot];
            const uint2* cand = a.buf + slot * a.cap;
            for (uint32_t b = 0; b < n; b += RS_NB)
                run_batch(lane < RS_NB && b + lane < n ? cand[b + lane].y : SFB_IDX_NONE, n - b < (uint32_t)RS_NB ? n - b : RS_NB);
        }
    }

    // ---- certification ------------------------------------------------------------------------
    const double bound = c == a.k ? ld[a.k - 1] : a.eps;  // everything not rescored must be farther than this
    bool certified;
    double margin = 0.0;
    if (thr == -INFINITY) certified = true;               // nothing was dropped or excluded for this row
    else {
        if (COS) {
            margin = cos_margin;
            double cos_ub = ((double)thr + margin) / s2;
            double lb = 1.0 - (cos_ub > 0.0 ? (cos_ub < 1.0 ? cos_ub : 1.0) : 0.0);
            certified = bound < lb;
            margin /= s2;
        } else {
            // not rescored: 2 S~ - nq32_j <= thr  =>  |q_i - q_j|^2 >= |q_i|^2 - thr - eta
            const double eta = eta_base + 2.4e-7 * fabs((double)thr);
            double d2 = q2 - (double)thr - eta;
            double lbs = (d2 > 0.0 ? sqrt(d2) * (1.0 - 1e-12) : 0.0) - rho;  // scaled lower bound on |y_i - y_j|
            double lb = lbs > 0.0 ? lbs / a.scale : 0.0;
            double b = bound * (1.0 + 1e-12);
            certified = a.metric == SFB_METRIC_L2 ? b < lb : b < lb * lb;
            margin = rho / a.scale;
        }
    }
    if (!(margin == margin)) certified = false;
    if (certified) {
        for (uint32_t t = lane; t < a.k; t += 32) {
            a.out_idx[orow * a.k + t] = t < c ? li[t] : SFB_IDX_NONE;
            a.out_dist[orow * a.k + t] = t < c ? ld[t] : INFINITY;
        }
        if (lane == 0) a.out_cnt[orow] = c;
    } else if (lane == 0) {
        a.fb_rows[atomicAdd(a.fb_count, 1u)] = gi;
    }
    if (lane == 0 && margin > 0.0 && isfinite(margin)) atomicMax(reinterpret_cast<unsigned long long*>(a.max_margin), (unsigned long long)__double_as_longlong(margin));
}

// query operand of the re-screen: the 16-bit rows of the uncertified queries, packed
__global__ void gather_rows16_kernel(const uint16_t* __restrict__ q, uint32_t kpad, const uint32_t* __restrict__ rows, uint32_t n,
                                     uint16_t* __restrict__ out) {
    const uint32_t per = kpad / 8;   // uint4 = 8 halves; kpad is a multiple of 64
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)n * per) return;
    const uint32_t r = (uint32_t)(gid / per), c = (uint32_t)(gid % per);
    reinterpret_cast<uint4*>(out)[(uint64_t)r * per + c] = __ldg(reinterpret_cast<const uint4*>(q) + (uint64_t)rows[r] * per + c);
}

__global__ void scatter_rows_kernel(const uint32_t* __restrict__ rows, uint32_t n, uint64_t q_begin, uint32_t k,
                                    const uint32_t* __restrict__ t_idx, const double* __restrict__ t_dist, const uint32_t* __restrict__ t_cnt,
                                    uint32_t* __restrict__ out_idx, double* __restrict__ out_dist, uint32_t* __restrict__ out_cnt) {
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)n * k) return;
    uint32_t r = (uint32_t)(gid / k), t = (uint32_t)(gid % k);
    uint64_t dst = (uint64_t)rows[r] - q_begin;
    out_idx[dst * k + t] = t_idx[gid]; out_dist[dst * k + t] = t_dist[gid];
    if (t == 0) out_cnt[dst] = t_cnt[r];
}

// ---- host -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int32_t make_tmap(sfb_ctx* ctx, CUtensorMap* map, void* base, uint64_t rows, uint32_t kpad, uint32_t box_rows, bool bf16) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || !p) return sfb_fail(ctx, SFB_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    cuuint64_t dims[2] = {kpad, rows};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sfb_fail(ctx, SFB_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SFB_OK;
}

uint32_t make_idesc(bool bf16) {
    // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format (F16 0 / BF16 1) @7 / @10, K-major A and B,
    // N >> 3 @17, M >> 4 @24
    uint32_t fmt = bf16 ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

size_t screen_smem_bytes() { return (size_t)STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 16 * 8 + 16 + 2 * BN * sizeof(float) + 1024; }

struct Prepared {
    DevBuf q, nq32, aux, gmax;
    uint64_t mpad = 0; uint32_t kpad = 0;
    double scale = 1.0, nmax = 0.0, dmax = 0.0;
    bool bf16 = false;
};

// `ctx->knn_collective` (sfb_knn_build_sharded: every rank of the communicator is in this call with the same matrix): each
// rank converts the rows of its own ceil-split shard and the operands, their norms and residuals are all-gathered over NVLink
// instead of being recomputed on every GPU.
int32_t prepare_operands(sfb_ctx* ctx, const sfb_mat* x, const double* norms, int metric, bool bf16, Prepared* P) {
    const uint64_t m = x->rows;
    const bool shared_work = ctx->knn_collective && ctx->world > 1;
    const uint64_t world = shared_work ? (uint64_t)ctx->world : 1, S = (m + world - 1) / world;
    const uint64_t r0 = shared_work ? ((uint64_t)ctx->rank * S < m ? (uint64_t)ctx->rank * S : m) : 0;
    const uint64_t r1 = shared_work ? (r0 + S < m ? r0 + S : m) : m;
    P->bf16 = bf16;
    P->kpad = (x->cols + BK - 1) / BK * BK;
    P->mpad = (m + BN - 1) / BN * BN;
    const uint64_t rows_alloc = P->mpad > world * S ? P->mpad : world * S;   // equal-count all-gather slots
    SFB_CUDA(ctx, P->q.alloc((size_t)rows_alloc * P->kpad * 2));
    SFB_CUDA(ctx, P->nq32.alloc(sizeof(float) * rows_alloc));
    SFB_CUDA(ctx, P->aux.alloc(sizeof(double) * 3 * rows_alloc));
    SFB_CUDA(ctx, P->gmax.alloc(4 * sizeof(unsigned long long)));
    SFB_CUDA(ctx, cudaMemsetAsync(P->q.p, 0, (size_t)rows_alloc * P->kpad * 2, ctx->stream));
    SFB_CUDA(ctx, cudaMemsetAsync(P->gmax.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    if (shared_work) {
        SFB_CUDA(ctx, cudaMemsetAsync(P->nq32.p, 0, sizeof(float) * rows_alloc, ctx->stream));
        SFB_CUDA(ctx, cudaMemsetAsync(P->aux.p, 0, sizeof(double) * 3 * rows_alloc, ctx->stream));
    }
    const bool cosine = metric == SFB_METRIC_COSINE;
    P->scale = COS_SCALE;
    if (!cosine) {
        if (r1 > r0) {
            absmax_kernel<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(x->d + r0 * x->cols, (r1 - r0) * x->cols, P->gmax.as<unsigned long long>() + 2);
            SFB_LAUNCH_CHECK(ctx);
        }
        if (shared_work) SFB_TRY(sfb_comm_allreduce_max_u64(ctx, P->gmax.as<unsigned long long>() + 2, 1));   // non-negative doubles order as integers
        double amax = 0.0;
        SFB_CUDA(ctx, cudaMemcpyAsync(&amax, P->gmax.as<unsigned long long>() + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!isfinite(amax)) return sfb_fail(ctx, SFB_EINVAL, "matrix holds non-finite values");
        // power-of-two scale that puts the largest |x| in [128, 256): exact in f64, far from fp16 overflow
        int e = 0;
        if (amax > 0.0) { frexp(amax, &e); e = 8 - e; }
        P->scale = ldexp(1.0, e);
    }
    if (r1 > r0) {
        if (bf16)
            prepare_kernel<true><<<div_up((r1 - r0) * 32, 256), 256, 0, ctx->stream>>>(x->d, norms, r0, r1 - r0, x->cols, P->kpad, cosine, P->scale, P->q.as<uint16_t>(),
                                                                                      P->nq32.as<float>(), P->aux.as<double>(), P->gmax.as<unsigned long long>());
        else
            prepare_kernel<false><<<div_up((r1 - r0) * 32, 256), 256, 0, ctx->stream>>>(x->d, norms, r0, r1 - r0, x->cols, P->kpad, cosine, P->scale, P->q.as<uint16_t>(),
                                                                                       P->nq32.as<float>(), P->aux.as<double>(), P->gmax.as<unsigned long long>());
        SFB_LAUNCH_CHECK(ctx);
    }
    if (shared_work) {
        SFB_TRY(sfb_comm_allgather_bytes(ctx, P->q.p, (size_t)S * P->kpad * 2));
        SFB_TRY(sfb_comm_allgather_bytes(ctx, P->nq32.p, (size_t)S * sizeof(float)));
        SFB_TRY(sfb_comm_allgather_bytes(ctx, P->aux.p, (size_t)S * 3 * sizeof(double)));
        SFB_TRY(sfb_comm_allreduce_max_u64(ctx, P->gmax.as<unsigned long long>(), 2));
    }
    double g[2];
    SFB_CUDA(ctx, cudaMemcpyAsync(g, P->gmax.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    P->nmax = g[0]; P->dmax = g[1];
    return SFB_OK;
}

template <bool DUMP>
int32_t launch_screen(sfb_ctx* ctx, const Prepared& P, int metric, ScreenArgs& sa, uint32_t m_blocks, void* a_base = nullptr, uint64_t a_rows = 0) {
    CUtensorMap tmA, tmB;
    SFB_TRY(make_tmap(ctx, &tmA, a_base ? a_base : P.q.p, a_base ? a_rows : P.mpad, P.kpad, BM, P.bf16));
    SFB_TRY(make_tmap(ctx, &tmB, P.q.p, P.mpad, P.kpad, BN, P.bf16));
    sa.idesc = make_idesc(P.bf16);
    size_t smem = screen_smem_bytes();
    dim3 grid(m_blocks, sa.n_splits);
    if (metric == SFB_METRIC_COSINE) {
        SFB_CUDA(ctx, cudaFuncSetAttribute(knn_screen_kernel<false, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_screen_kernel<false, DUMP><<<grid, SCREEN_THREADS, smem, ctx->stream>>>(tmA, tmB, sa);
    } else {
        SFB_CUDA(ctx, cudaFuncSetAttribute(knn_screen_kernel<true, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_screen_kernel<true, DUMP><<<grid, SCREEN_THREADS, smem, ctx->stream>>>(tmA, tmB, sa);
    }
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

// SFB_SCREEN_EW8=1: the eight-warp epilogue on shared row buffers (experiment)
static bool screen_ew8() {
    static const bool on = [] { const char* e = getenv("SFB_SCREEN_EW8"); return e && e[0] == '1'; }();
    return on;
}

int32_t launch_screen_pair(sfb_ctx* ctx, const Prepared& P, int metric, Screen2Args& sa, void* a_base, uint64_t a_rows) {
    CUtensorMap tm, tmA;
    SFB_TRY(make_tmap(ctx, &tm, P.q.p, P.mpad, P.kpad, P_BM, P.bf16));
    SFB_TRY(make_tmap(ctx, &tmA, a_base, a_rows, P.kpad, P_BM, P.bf16));
    // M = 256 across the pair, N = 256
    const uint32_t fmt = P.bf16 ? 1u : 0u;
    sa.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const size_t fixed = 1024 + 22 * 8 + 16 + 2 * BN * sizeof(float) + 1024;   // + row counts / thresholds of the 8-warp epilogue
    const size_t budget = ctx->smem_optin ? ctx->smem_optin : 232448;
    const size_t room = (budget - fixed) / P_SLAB_BYTES;   // 16 KB slabs that fit
    if (sa.kblocks + 5 <= room) {
        // every k-block of the query rows is resident; the stages carry corpus half-tiles only
        sa.resident = sa.kblocks; sa.stage_bytes = P_SLAB_BYTES;
        size_t stages = room - sa.kblocks;
        sa.stages = (uint32_t)(stages > 8 ? 8 : stages);
        if (const char* e = getenv("SFB_SCREEN_STAGES")) { int v = atoi(e); if (v >= 2 && (uint32_t)v <= sa.stages) sa.stages = (uint32_t)v; }
    } else {
        // K > 512: keep what fits beside four 32 KB stages, stream the remaining query slabs with the corpus
        sa.stages = 4; sa.stage_bytes = 2 * P_SLAB_BYTES;
        sa.resident = (uint32_t)(room - 2 * sa.stages);
        if (sa.resident > sa.kblocks) sa.resident = sa.kblocks;
    }
    const size_t smem = fixed + (size_t)sa.resident * P_SLAB_BYTES + (size_t)sa.stages * sa.stage_bytes;
    const uint32_t n_slots = sa.n_mb2 * sa.n_splits;
    uint32_t pairs = (uint32_t)ctx->sm_count / 2;
    if (pairs > n_slots) pairs = n_slots;
    const bool ew8 = screen_ew8();
    auto launch = [&](auto kernel, int threads) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<2 * pairs, threads, smem, ctx->stream>>>(tm, tmA, sa);
        return cudaSuccess;
    };
    if (metric == SFB_METRIC_COSINE) SFB_CUDA(ctx, ew8 ? launch(knn_screen_pair_kernel<false, true>, SCREEN_THREADS8) : launch(knn_screen_pair_kernel<false, false>, SCREEN_THREADS));
    else SFB_CUDA(ctx, ew8 ? launch(knn_screen_pair_kernel<true, true>, SCREEN_THREADS8) : launch(knn_screen_pair_kernel<true, false>, SCREEN_THREADS));
    SFB_LAUNCH_CHECK(ctx);
    sfb_side_job_fire(ctx);   // a pending feature-graph Gram rides beside the resident screen CTAs
    return SFB_OK;
}

// SFB_SCREEN_V1=1 forces the single-CTA kernel (A/B comparisons)
bool pair_kernel_applies(const Prepared& P) {
    const char* v1 = getenv("SFB_SCREEN_V1");
    return !(v1 && v1[0] == '1');
}

}  // namespace

namespace {

// One screen + rescore + certify pass over `nq` query rows.  The query operand is `a_base` (a_rows rows of
// 16-bit operands; query row rl is its row a_row0 + rl); qlist == null: query rl is global row a_row0 + rl,
// else global row qlist[rl].  Certified rows are written to `out` (row g - out_base); the global indices of the
// rest are appended to fb_rows / fb_count.
int32_t screen_level(sfb_ctx* ctx, const sfb_mat* x, const double* norms, const sfb_knn_params* p, const Prepared& P,
                     uint32_t kprime, void* a_base, uint64_t a_rows, uint64_t a_row0, uint64_t nq, const uint32_t* qlist,
                     uint64_t out_base, sfb_knn* out, uint32_t* fb_rows, uint64_t* fb_count /* [0] count, [1] max margin */,
                     double* ms_screen, double* ms_rescore) {
    const uint64_t m = x->rows;
    const bool cosine = p->metric == SFB_METRIC_COSINE;
    uint32_t slack = 64;   // appends a row's buffer takes between two prunes (minus the 32 a chunk may add)
    if (const char* e = getenv("SFB_SCREEN_SLACK")) { int v = atoi(e); if (v >= 64) slack = (uint32_t)v; }
    uint32_t cap = (kprime + slack + (screen_ew8() ? 32u : 0u) + 31) / 32 * 32;   // the 8-warp epilogue reserves 64 slots per chunk round
    if (cap > MAX_CAP) cap = MAX_CAP;
    const uint32_t tiles_total = (uint32_t)(P.mpad / BN);
    const bool use_pair = pair_kernel_applies(P);
    ScreenArgs sa{};
    Screen2Args s2{};
    uint32_t n_splits_used = 1;
    const uint32_t m_blocks = (uint32_t)((nq + BM - 1) / BM);
    if (use_pair) {
        s2.n_rows = m; s2.q_begin = a_row0; s2.nq = nq; s2.kblocks = P.kpad / BK;
        s2.tiles_total = tiles_total;
        s2.n_mb2 = (uint32_t)((nq + 255) / 256);
        const uint32_t pairs = (uint32_t)ctx->sm_count / 2;
        // corpus splits give small launches parallelism, but every split keeps its own k' candidates per row and the
        // rescore gathers all of them: more than 8 made the re-screen of 9 rows take 17 ms (148 x 192 candidates each)
        uint32_t n_splits = s2.n_mb2 >= 4 * pairs ? 1u : (2 * pairs + s2.n_mb2 - 1) / s2.n_mb2;
        if (n_splits > 8) n_splits = 8;
        if (n_splits > tiles_total) n_splits = tiles_total;
        s2.tiles_per_split = (tiles_total + n_splits - 1) / n_splits;
        s2.n_splits = (tiles_total + s2.tiles_per_split - 1) / s2.tiles_per_split;
        // corpus chunk that stays L2-resident while every pair sweeps it
        double chunk_mb = 24.0;
        if (const char* e = getenv("SFB_SCREEN_CHUNK_MB")) { double v = atof(e); if (v > 0.0) chunk_mb = v; }
        uint64_t ct = (uint64_t)(chunk_mb * 1048576.0 / ((double)BN * P.kpad * 2.0));
        if (ct < 4) ct = 4;
        if (ct > s2.tiles_per_split) ct = s2.tiles_per_split;
        s2.chunk_tiles = (uint32_t)ct;
        s2.n_chunks = (s2.tiles_per_split + s2.chunk_tiles - 1) / s2.chunk_tiles;
        s2.kprime = kprime; s2.cap = cap; s2.nq32 = P.nq32.as<float>();
        if (const char* e = getenv("SFB_SCREEN_DBG")) s2.dbg = (uint32_t)atoi(e);
        n_splits_used = s2.n_splits;
    } else {
        sa.n_rows = m; sa.q_begin = a_row0; sa.nq = nq; sa.kblocks = P.kpad / BK;
        sa.tiles_total = tiles_total;
        uint32_t want_ctas = 2u * (uint32_t)ctx->sm_count;
        uint32_t n_splits = m_blocks >= want_ctas ? 1u : (want_ctas + m_blocks - 1) / m_blocks;
        if (n_splits > 8) n_splits = 8;
        if (n_splits > sa.tiles_total) n_splits = sa.tiles_total;
        sa.tiles_per_split = (sa.tiles_total + n_splits - 1) / n_splits;
        sa.n_splits = (sa.tiles_total + sa.tiles_per_split - 1) / sa.tiles_per_split;
        sa.kprime = kprime; sa.cap = cap; sa.nq32 = P.nq32.as<float>();
        n_splits_used = sa.n_splits;
    }

    const size_t slots = (size_t)nq * n_splits_used;
    DevBuf buf, cnt, thr;
    SFB_CUDA(ctx, buf.alloc(slots * cap * sizeof(uint2)));
    SFB_CUDA(ctx, cnt.alloc(slots * sizeof(uint32_t)));
    SFB_CUDA(ctx, thr.alloc(slots * sizeof(float)));
    SFB_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, slots * sizeof(uint32_t), ctx->stream));
    sa.buf = s2.buf = buf.as<uint2>();
    sa.out_cnt = s2.out_cnt = cnt.as<uint32_t>(); sa.out_thr = s2.out_thr = thr.as<float>();
    sa.n_splits = n_splits_used;
    {
        StageTimer t(ctx, nullptr);
        if (use_pair && s2.dbg == 4) {
            unsigned long long z[8] = {0}; int on = 1;
            cudaMemcpyToSymbol(g_screen_dbg, z, sizeof z); cudaMemcpyToSymbol(g_prof_on, &on, sizeof on);
        }
        if (use_pair) SFB_TRY(launch_screen_pair(ctx, P, p->metric, s2, a_base, a_rows));
        if (use_pair && s2.dbg == 4) {
            unsigned long long z[8]; int off = 0;
            cudaStreamSynchronize(ctx->stream);
            cudaMemcpyFromSymbol(z, g_screen_dbg, sizeof z); cudaMemcpyToSymbol(g_prof_on, &off, sizeof off);
            fprintf(stderr, "[sfb] screen dbg (warp 2 of %d CTAs): tiles %llu, wait-for-accumulator %.0f cyc/tile, slow path %.2f entries/tile at %.0f cyc, "
                    "prunes %.3f calls/tile (%.2f rows each) at %.0f cyc\n", 2 * (ctx->sm_count / 2), z[6], z[6] ? (double)z[5] / z[6] : 0.0,
                    z[6] ? (double)z[1] / z[6] : 0.0, z[1] ? (double)z[0] / z[1] : 0.0, z[6] ? (double)z[3] / z[6] : 0.0, z[3] ? (double)z[4] / z[3] : 0.0,
                    z[3] ? (double)z[2] / z[3] : 0.0);
        }
        else SFB_TRY(launch_screen<false>(ctx, P, p->metric, sa, m_blocks, a_base, a_rows));
        *ms_screen += t.stop();
        SFB_CUDA(ctx, cudaGetLastError());
    }
    // tensor-core fp32 accumulation: K products, each partial sum off by at most 2 ulp of the running bound
    const double gamma = ((double)P.kpad + 64.0) * ldexp(1.0, -23);
    RescoreArgs ra{x->d, norms, m, x->cols, p->metric, p->k, p->eps, a_row0, nq, sa.n_splits, cap,
                   sa.buf, sa.out_cnt, sa.out_thr, P.aux.as<double>(), P.nmax, P.dmax, gamma, P.scale,
                   out->idx, out->dist, out->cnt, fb_rows, reinterpret_cast<uint32_t*>(fb_count),
                   reinterpret_cast<double*>(fb_count + 1), reinterpret_cast<unsigned long long*>(fb_count + 4), qlist, out_base};
    {
        StageTimer t(ctx, nullptr);
        const int wpb = 4;
        // Measured (C2 / C4, ms): cp.async two stages 27.3 / 136, three 33.4 / 133; bulk two stages 29.1 / 142, three 38.3 / 142 --
        // neither the gather mechanism nor the ring depth is the limit, the number of resident warps is (three CTAs per SM with two
        // stages; ncu: issue slots 42 % busy, stalls on fixed-latency dependencies and shared-memory loads, DRAM 3.7 TB/s at full
        // clock).  Default: cp.async, two stages.  SFB_RESCORE_BULK=1 selects the bulk gathers (16-byte aligned row chunks: even D),
        // SFB_RESCORE_NST=3 the deeper ring, SFB_RESCORE_CH=16 / 32 the chunk.
        const bool bulk = (x->cols & 1u) == 0 && (reinterpret_cast<uintptr_t>(x->d) & 15u) == 0 && getenv("SFB_RESCORE_BULK") != nullptr;
        int nst = 2, ch = 32;
        if (const char* e = getenv("SFB_RESCORE_NST")) { if (atoi(e) == 3) nst = 3; }
        if (const char* e = getenv("SFB_RESCORE_CH")) { if (atoi(e) == 16) ch = 16; }
        if (bulk) ch = 32;
        const size_t smem = (size_t)wpb * rs_warp_bytes(bulk, ch, nst, p->k);
#define SFB_RESCORE(C_, N_, B_, H_)                                                                                                      \
    do {                                                                                                                                 \
        SFB_CUDA(ctx, cudaFuncSetAttribute(knn_rescore_kernel<C_, N_, B_, H_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        knn_rescore_kernel<C_, N_, B_, H_><<<div_up(nq, wpb), wpb * 32, smem, ctx->stream>>>(ra);                                        \
    } while (0)
#define SFB_RESCORE_C(C_)                                                                            \
    do {                                                                                             \
        if (bulk) { if (nst == 3) SFB_RESCORE(C_, 3, true, 32); else SFB_RESCORE(C_, 2, true, 32); } \
        else if (ch == 16) { if (nst == 3) SFB_RESCORE(C_, 3, false, 16); else SFB_RESCORE(C_, 2, false, 16); } \
        else { if (nst == 3) SFB_RESCORE(C_, 3, false, 32); else SFB_RESCORE(C_, 2, false, 32); }    \
    } while (0)
        if (cosine) SFB_RESCORE_C(true); else SFB_RESCORE_C(false);
#undef SFB_RESCORE_C
#undef SFB_RESCORE
        SFB_LAUNCH_CHECK(ctx);
        *ms_rescore += t.stop();   // synchronises: the candidate buffers may be released on return
    }
    return SFB_OK;
}

}  // namespace

// Screen in up to three levels, each exact for the rows it certifies:
//   1. every query row with a small k' (cheap epilogue: the cost of the screen grows with k');
//   2. the rows level 1 could not certify -- near-ties at the margin, tight clusters -- once more through the tensor
//      cores with k' = 192, their 16-bit rows packed into a small query operand;
//   3. what is left (exact ties, zero rows, duplicates) by f64 brute force.
int32_t sfb_knn_screened(sfb_ctx* ctx, const sfb_mat* x, const double* norms, const sfb_knn_params* p, uint64_t q_begin,
                         uint64_t q_end, sfb_knn* out) {
    const uint64_t m = x->rows, nq = q_end - q_begin;
    const bool bf16 = p->screen == SFB_SCREEN_BF16;
    if (m > 0xFFFFFF00ull) return sfb_fail(ctx, SFB_EUNSUPPORTED, "too many rows for the screen");
    sfb_knn_stats& st = out->stats;
    st.screen_used = p->screen;

    // k' candidates survive per row and corpus split; the buffer has 64 slots of slack between prunes.
    // Measured at C2 (k = 16): k' = 96 / 64 / 48 / 32 / 24 take 814 / 717 / 689 / 662 / 645 ms of screen and leave
    // 0 / 0 / 0 / 9 / 3366 rows to the next level.
    const uint32_t kp_max = MAX_CAP - (screen_ew8() ? 96 : 64);
    uint32_t kprime = p->k_prime ? p->k_prime : (3 * p->k / 2 + 7) / 8 * 8;   // 1.5k (C5-like, k = 64: k' = 128 / 96 -> 980 / 895 ms), at least 32
    bool kp_forced = p->k_prime != 0;
    if (const char* e = getenv("SFB_SCREEN_KPRIME")) { int v = atoi(e); if (v > 0 && !p->k_prime) { kprime = (uint32_t)v; kp_forced = true; } }  // tuning aid
    if (kprime < 32 && !kp_forced) kprime = 32;
    if (kprime < p->k + 1) kprime = p->k + 1;
    if (kprime > kp_max) kprime = kp_max;
    if (kprime < p->k + 1) return sfb_fail(ctx, SFB_EUNSUPPORTED, "k too large for the screen buffers");
    st.k_prime = kprime;

    HostTrace tr(ctx, "knn_screened");
    Prepared P;
    {
        StageTimer t(ctx, nullptr);
        SFB_TRY(prepare_operands(ctx, x, norms, p->metric, bf16, &P));
        st.ms_prepare = t.stop();
    }
    tr.mark("prepare");
    DevBuf fb_rows, fb_rows2, fb_count;
    SFB_CUDA(ctx, fb_rows.alloc(nq * sizeof(uint32_t)));
    SFB_CUDA(ctx, fb_count.alloc(8 * sizeof(uint64_t)));   // level 1: [0] uncertified rows, [1] max margin, [4] rows rescored; level 2: [2], [3], [6]
    SFB_CUDA(ctx, cudaMemsetAsync(fb_count.p, 0, 8 * sizeof(uint64_t), ctx->stream));
    uint64_t* fbc = fb_count.as<uint64_t>();

    // level 1
    SFB_TRY(screen_level(ctx, x, norms, p, P, kprime, P.q.p, P.mpad, q_begin, nq, nullptr, q_begin, out, fb_rows.as<uint32_t>(), fbc,
                         &st.ms_screen, &st.ms_rescore));
    uint64_t h[8];
    SFB_CUDA(ctx, cudaMemcpyAsync(h, fb_count.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t n_fb = (uint32_t)(h[0] & 0xFFFFFFFFu);
    memcpy(&st.max_margin, &h[1], 8);
    st.candidates_rescored = h[4];
    const uint32_t* fb_final = fb_rows.as<uint32_t>();
    tr.mark("level 1");

    // level 2: re-screen the uncertified rows with the widest k'
    const bool rescreen_off = getenv("SFB_SCREEN_NO_RESCREEN") != nullptr;
    // a re-screen launch has ~4 ms of fixed cost; below a few dozen rows the f64 brute force (0.1 ms / row at 1M x 384) is cheaper
    if (n_fb >= 32 && kprime < kp_max && !rescreen_off && kp_max >= p->k + 1) {
        StageTimer t(ctx, nullptr);
        st.rows_rescreened = n_fb;
        DevBuf qa;
        const uint64_t a_rows = ((uint64_t)n_fb + 255) / 256 * 256;
        SFB_CUDA(ctx, qa.alloc(a_rows * P.kpad * 2));
        SFB_CUDA(ctx, cudaMemsetAsync(qa.p, 0, a_rows * P.kpad * 2, ctx->stream));
        gather_rows16_kernel<<<div_up((uint64_t)n_fb * (P.kpad / 8), 256), 256, 0, ctx->stream>>>(P.q.as<uint16_t>(), P.kpad, fb_rows.as<uint32_t>(), n_fb,
                                                                                                 qa.as<uint16_t>());
        SFB_LAUNCH_CHECK(ctx);
        SFB_CUDA(ctx, fb_rows2.alloc((size_t)n_fb * sizeof(uint32_t)));
        SFB_CUDA(ctx, cudaMemsetAsync(fbc + 2, 0, 2 * sizeof(uint64_t), ctx->stream));
        double ms_s = 0.0, ms_r = 0.0;
        SFB_TRY(screen_level(ctx, x, norms, p, P, kp_max, qa.p, a_rows, 0, n_fb, fb_rows.as<uint32_t>(), q_begin, out, fb_rows2.as<uint32_t>(),
                             fbc + 2, &ms_s, &ms_r));
        SFB_CUDA(ctx, cudaMemcpyAsync(h, fb_count.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        n_fb = (uint32_t)(h[2] & 0xFFFFFFFFu);
        st.candidates_rescored += h[6];
        fb_final = fb_rows2.as<uint32_t>();
        st.ms_rescreen = t.stop();
        tr.mark("level 2");
    }
    st.rows_fallback = n_fb;
    st.rows_certified = nq - n_fb;
    if (n_fb) {
        if (!p->allow_fallback) return sfb_fail(ctx, SFB_EUNCERTIFIED, "%u of %llu rows not certified by the screen", n_fb, (unsigned long long)nq);
        StageTimer t(ctx, nullptr);
        DevBuf t_idx, t_dist, t_cnt;
        SFB_CUDA(ctx, t_idx.alloc((size_t)n_fb * p->k * sizeof(uint32_t)));
        SFB_CUDA(ctx, t_dist.alloc((size_t)n_fb * p->k * sizeof(double)));
        SFB_CUDA(ctx, t_cnt.alloc((size_t)n_fb * sizeof(uint32_t)));
        SFB_TRY(sfb_knn_exact(ctx, x, norms, p->metric, p->k, p->eps, fb_final, n_fb, 0, t_idx.as<uint32_t>(),
                              t_dist.as<double>(), t_cnt.as<uint32_t>()));
        scatter_rows_kernel<<<div_up((uint64_t)n_fb * p->k, 256), 256, 0, ctx->stream>>>(fb_final, n_fb, q_begin, p->k,
                                                                                         t_idx.as<uint32_t>(), t_dist.as<double>(), t_cnt.as<uint32_t>(),
                                                                                         out->idx, out->dist, out->cnt);
        SFB_LAUNCH_CHECK(ctx);
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        st.ms_fallback = t.stop();
    }
    return SFB_OK;
}

// Diagnostic: the raw tensor-core accumulators of one 128 x 256 tile (rows row0.., columns col0.. rounded
// down to a multiple of 256) plus the operands the screen used, for validating the MMA path and the
// accumulation-error model in tests.  q_out: mpad x kpad operand values as f32 (may be NULL).
extern "C" int32_t sfb_debug_screen_tile(sfb_ctx* ctx, const sfb_mat* x, int32_t metric, int32_t screen, uint64_t row0,
                                         uint64_t col0, float* out_tile /* 128*256 */, float* q_rows /* 128*kpad */,
                                         float* q_cols /* 256*kpad */, uint32_t* kpad_out, double* scale_out) {
    if (!ctx || !x || !out_tile) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    DevBuf norms, dump;
    SFB_CUDA(ctx, norms.alloc(sizeof(double) * x->rows));
    if (metric == SFB_METRIC_COSINE) SFB_TRY(sfb_row_norms(ctx, x, norms.as<double>()));
    Prepared P;
    SFB_TRY(prepare_operands(ctx, x, norms.as<double>(), metric, screen == SFB_SCREEN_BF16, &P));
    SFB_CUDA(ctx, dump.alloc(sizeof(float) * BM * BN));
    ScreenArgs sa{};
    sa.n_rows = x->rows; sa.q_begin = row0; sa.nq = BM; sa.kblocks = P.kpad / BK;
    sa.tiles_total = (uint32_t)(col0 / BN) + 1; sa.tiles_per_split = sa.tiles_total; sa.n_splits = 1;
    // run only the requested tile: start the split at it
    sa.tiles_per_split = 1; sa.n_splits = sa.tiles_total;
    sa.kprime = 64; sa.cap = 128; sa.nq32 = P.nq32.as<float>(); sa.dump = dump.as<float>();
    // grid.y = n_splits CTAs would all dump; launch a single CTA positioned on the tile instead
    {
        CUtensorMap tmA, tmB;
        SFB_TRY(make_tmap(ctx, &tmA, P.q.p, P.mpad, P.kpad, BM, P.bf16));
        SFB_TRY(make_tmap(ctx, &tmB, (uint8_t*)P.q.p + (col0 / BN) * BN * (size_t)P.kpad * 2, P.mpad - (col0 / BN) * BN, P.kpad, BN, P.bf16));
        sa.idesc = make_idesc(P.bf16);
        sa.tiles_total = 1; sa.n_splits = 1;
        size_t smem = screen_smem_bytes();
        SFB_CUDA(ctx, cudaFuncSetAttribute(knn_screen_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_screen_kernel<false, true><<<dim3(1, 1), SCREEN_THREADS, smem, ctx->stream>>>(tmA, tmB, sa);
        SFB_LAUNCH_CHECK(ctx);
    }
    SFB_CUDA(ctx, cudaMemcpyAsync(out_tile, dump.p, sizeof(float) * BM * BN, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (kpad_out) *kpad_out = P.kpad;
    if (scale_out) *scale_out = P.scale;
    // operands back as f32 (host converts): copy raw 16-bit rows
    auto copy_rows = [&](uint64_t r0, uint32_t nr, float* dst) -> int32_t {
        if (!dst) return SFB_OK;
        std::string tmp;
        tmp.resize((size_t)nr * P.kpad * 2);
        uint64_t avail = P.mpad > r0 ? P.mpad - r0 : 0;
        uint32_t take = avail < nr ? (uint32_t)avail : nr;
        memset(&tmp[0], 0, tmp.size());
        if (take) SFB_CUDA(ctx, cudaMemcpy(&tmp[0], (uint8_t*)P.q.p + r0 * (size_t)P.kpad * 2, (size_t)take * P.kpad * 2, cudaMemcpyDeviceToHost));
        const uint16_t* h = reinterpret_cast<const uint16_t*>(tmp.data());
        for (size_t e = 0; e < (size_t)nr * P.kpad; ++e) {
            if (P.bf16) { uint32_t u = (uint32_t)h[e] << 16; memcpy(&dst[e], &u, 4); }
            else dst[e] = __half2float(__ushort_as_half(h[e]));
        }
        return SFB_OK;
    };
    SFB_TRY(copy_rows(row0, BM, q_rows));
    SFB_TRY(copy_rows(col0 / BN * BN, BN, q_cols));
    return SFB_OK;
}
