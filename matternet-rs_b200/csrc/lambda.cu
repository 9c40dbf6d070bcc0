// lambda.cu -- per-item taumode lambda, min-max normalisation, diffusion.
//
// Reference: TauMode::compute_taumode_lambdas_parallel / compute_synthetic_lambda /
// compute_rayleigh_quotient_from_matrix / compute_item_dispersion / select_tau
// (src_legacy/taumode.rs:29-70,117-318,326-408), node_energy_and_dispersion
// (src_legacy/energymaps.rs:923-1045), compute_lambdas_gpu (surfface-core/src/spectral/mod.rs:69-181),
// normalise_lambdas (src_legacy/core.rs:1341-1355), diffusion (src_legacy/energymaps.rs:520-546).
//
// L is F x F (feature graph, a few thousand non-zeros): it stays in L2/L1.  X (R x F, f64) is
// streamed once: one warp per item row, the row staged in shared memory with coalesced loads, the
// CSR rows of L dealt round-robin to the lanes.  The reference's dispersion probes all F^2 pairs
// (taumode.rs:371-383); only the nnz(L) stored pairs are non-zero, which is what is visited here, in
// the same (r asc, c asc) order per lane.  Cross-lane sums use a fixed butterfly, so results are
// deterministic and within 1e-12 relative of the reference's left fold.
#include <math.h>

#include "common.cuh"

int32_t sfb_comm_allreduce_sum_f32(sfb_ctx* ctx, float* buf, size_t n);

namespace {

constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_sum_u(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ uint64_t sort_key(double v) {
    uint64_t u = (uint64_t)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(uint64_t k) {
    uint64_t u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)u);
}

// Warp radix select over the finite entries of xs[0..f): returns the key of rank `rank` (0-based,
// ascending).  hist: 256 u32 of shared memory owned by this warp.  Also returns how many finite
// entries are strictly below / equal to the selected key.
__device__ uint64_t warp_select(const double* xs, uint32_t f, uint32_t rank, uint32_t* hist, int lane, uint32_t* n_below,
                                uint32_t* n_equal) {
    uint64_t prefix = 0, mask = 0;
    uint32_t below = 0, remaining = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = lane; b < 256; b += 32) hist[b] = 0;
        __syncwarp();
        for (uint32_t t = lane; t < f; t += 32) {
            double v = xs[t];
            if (!isfinite(v)) continue;
            uint64_t key = sort_key(v);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
        }
        __syncwarp();
        // each lane owns 8 consecutive bins
        uint32_t c[8], s = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; s += c[b]; }
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        uint32_t ex = inc - s;  // finite entries in lower bins (within the current prefix)
        uint32_t want = rank - below;
        int found_bin = -1; uint32_t found_below = 0, found_cnt = 0;
        if (want >= ex && want < ex + s) {
            uint32_t run = ex;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (found_bin < 0 && want < run + c[b]) { found_bin = lane * 8 + b; found_below = run; found_cnt = c[b]; }
                run += c[b];
            }
        }
        uint32_t who = __ballot_sync(FULL, found_bin >= 0);
        int src = __ffs(who) - 1;
        found_bin = __shfl_sync(FULL, found_bin, src);
        found_below = __shfl_sync(FULL, found_below, src);
        found_cnt = __shfl_sync(FULL, found_cnt, src);
        below += found_below;
        remaining = found_cnt;
        prefix |= (uint64_t)found_bin << shift;
        mask |= 0xFFull << shift;
        __syncwarp();
    }
    *n_below = below; *n_equal = remaining;
    return prefix;
}

// tau of one item row held in shared memory: TauMode::select_tau (taumode.rs:29-70)
__device__ double warp_select_tau(const double* xs, uint32_t f, int mode, double value, uint32_t* hist, int lane) {
    const double FLOOR = 1e-10;
    if (mode == SFB_TAU_FIXED) return (isfinite(value) && value > 0.0) ? value : FLOOR;
    uint32_t n = 0; double s = 0.0;
    for (uint32_t t = lane; t < f; t += 32) { double v = xs[t]; if (isfinite(v)) { ++n; s += v; } }
    n = warp_sum_u(n);
    if (n == 0) return FLOOR;
    if (mode == SFB_TAU_MEAN) { double mean = warp_sum(s) / (double)n; return mean > FLOOR ? mean : FLOOR; }
    uint32_t nb, ne;
    double r;
    if (mode == SFB_TAU_PERCENTILE) {
        double pp = value < 0.0 ? 0.0 : (value > 1.0 ? 1.0 : value);
        uint32_t idx = (uint32_t)round((double)(n - 1) * pp);
        r = key_value(warp_select(xs, f, idx, hist, lane, &nb, &ne));
    } else if (n & 1u) {
        r = key_value(warp_select(xs, f, n / 2, hist, lane, &nb, &ne));
    } else {
        uint64_t ka = warp_select(xs, f, n / 2 - 1, hist, lane, &nb, &ne);
        double a = key_value(ka), b;
        if (nb + ne > n / 2) b = a;  // duplicates of a reach rank n/2
        else {
            uint64_t kb = 0xFFFFFFFFFFFFFFFFull;
            for (uint32_t t = lane; t < f; t += 32) {
                double v = xs[t];
                if (!isfinite(v)) continue;
                uint64_t kv = sort_key(v);
                if (kv > ka && kv < kb) kb = kv;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) { uint64_t t = __shfl_xor_sync(FULL, kb, o); kb = t < kb ? t : kb; }
            b = key_value(kb);
        }
        r = 0.5 * (a + b);
    }
    return r > FLOOR ? r : FLOOR;
}

struct LambdaArgs {
    const uint64_t* indptr; const uint32_t* indices; const double* data; uint32_t f;
    const double* x; uint64_t n;
    int variant, tau_mode; double tau_value;
    double* out_lambda; double* out_disp; float* out_energy_f32;
    const double* tau_in;   // per item: tau chosen upstream (< 0: the item is a zero vector); null: select it from the row
};

// shared memory: per warp f doubles (the item row) + 256 u32 (select histogram)
template <int VARIANT>
__global__ void lambda_kernel(LambdaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* xs = reinterpret_cast<double*>(smem_raw) + (size_t)w * a.f;
    uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)wpb * a.f) + w * 256;
    const uint32_t f = a.f;
    for (uint64_t i = (uint64_t)blockIdx.x * wpb + w; i < a.n; i += (uint64_t)gridDim.x * wpb) {
        const double* xr = a.x + i * f;
        __syncwarp();
        bool zero = true;
        double den = 0.0;
        for (uint32_t t = lane; t < f; t += 32) {
            double v = xr[t];
            xs[t] = v;
            zero = zero && (fabs(v) <= 1e-10);
            den += v * v;
        }
        __syncwarp();
        zero = a.tau_in ? a.tau_in[i] < 0.0 : __all_sync(FULL, zero);   // projected items: the test was made on the unprojected vector
        if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE && zero) {  // taumode.rs:268-274
            if (lane == 0) { a.out_lambda[i] = 0.0; if (a.out_disp) a.out_disp[i] = 0.0; }
            continue;
        }
        if (VARIANT == SFB_LAMBDA_CORE_F32SEM) {
            float num = 0.f, den32 = 0.f, es = 0.f;
            for (uint32_t r = lane; r < f; r += 32) {
                float xv = (float)xs[r], lx = 0.f, wx = 0.f, wx2 = 0.f, dg = 0.f;
                for (uint64_t e = a.indptr[r]; e < a.indptr[r + 1]; ++e) {
                    float lv = (float)a.data[e], xc = (float)xs[a.indices[e]];
                    float wv = -lv > 0.f ? -lv : 0.f;
                    lx = __fadd_rn(lx, __fmul_rn(lv, xc));
                    wx = __fadd_rn(wx, __fmul_rn(wv, xc));
                    wx2 = __fadd_rn(wx2, __fmul_rn(wv, __fmul_rn(xc, xc)));
                    dg = __fadd_rn(dg, wv);
                }
                num += xv * lx; den32 += xv * xv;
                float ee = __fadd_rn(__fadd_rn(__fmul_rn(dg, __fmul_rn(xv, xv)), -__fmul_rn(__fmul_rn(xv, wx), 2.0f)), wx2);
                es += ee > 0.f ? ee : 0.f;
            }
            num = warp_sum_f(num); den32 = warp_sum_f(den32); es = warp_sum_f(es);
            float rq = num / (den32 + 1e-9f);
            rq = rq < -1e6f ? -1e6f : (rq > 1e6f ? 1e6f : rq);
            if (lane == 0) { a.out_lambda[i] = (double)rq; a.out_energy_f32[i] = es; }
            continue;
        }
        den = warp_sum(den);
        double num = 0.0, ssum = 0.0, qsum = 0.0;
        for (uint32_t r = lane; r < f; r += 32) {
            const double xv = xs[r];
            double rs = 0.0;
            const uint64_t e0 = a.indptr[r], e1 = a.indptr[r + 1];
            for (uint64_t e = e0; e < e1; ++e) {
                const uint32_t c = a.indices[e];
                const double lv = a.data[e], xc = xs[c];
                if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) rs = __dadd_rn(rs, __dmul_rn(__dmul_rn(xv, lv), xc));  // taumode.rs:348-350
                else rs = __dadd_rn(rs, __dmul_rn(lv, xc));                                                     // graph.rs:489-491
                const bool use = VARIANT == SFB_LAMBDA_LEGACY_TAUMODE ? (c != r) : (c > r);
                const double wv = -lv;
                if (use && wv > 0.0) {
                    double d = xv - xc;
                    double contrib = (wv * d) * d;
                    ssum += contrib;
                    qsum = fma(contrib, contrib, qsum);
                }
            }
            num += VARIANT == SFB_LAMBDA_LEGACY_TAUMODE ? rs : xv * rs;
        }
        num = warp_sum(num); ssum = warp_sum(ssum); qsum = warp_sum(qsum);
        double e_raw = 0.0;
        if (den > 1e-12) { e_raw = num / den; if (!(e_raw > 0.0)) e_raw = 0.0; }
        double g = 0.0;
        // a NaN sum: the taumode form tests `<= 1e-12 -> 0` and lets it propagate through f64::clamp (taumode.rs:385-407), the energy form tests `> 1e-12` (energymaps.rs:1006)
        if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE ? !(ssum <= 1e-12) : (ssum > 1e-12)) { g = qsum / (ssum * ssum); g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g); }
        double lam;
        if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) {
            const double tau = a.tau_in ? a.tau_in[i] : warp_select_tau(xs, f, a.tau_mode, a.tau_value, hist, lane);
            lam = tau * (e_raw / (e_raw + tau)) + (1.0 - tau) * g;  // taumode.rs:306-310
        } else {
            lam = e_raw;
        }
        if (lane == 0) { a.out_lambda[i] = lam; if (a.out_disp) a.out_disp[i] = g; }
    }
}

// ---- symmetric fast path -----------------------------------------------------------------------------------
// When L is symmetric bit for bit (every Laplacian this library builds; checked once for matrices from the host) the
// off-diagonal half below the diagonal repeats the half above it: x^T L x = sum_r L_rr x_r^2 + 2 sum_{r<c} L_rc x_r x_c
// and the dispersion terms of (r, c) and (c, r) are equal.  Moreover x^T L x = sum_r (L_rr + sum_{c != r} L_rc) x_r^2
// - sum_{r<c} L_rc (x_r - x_c)^2: with the row-sum defect precomputed (exactly 0 for L = D - W) the Rayleigh numerator is
// the same edge sum as the dispersion's S -- no cancellation (a constant vector gives exactly 0, as the reference's
// clamped value does) and no second product per edge.  The strict upper triangle is packed once per call
// ((r << 16 | c) u32 + value f64, CSR order: deterministic) and kept in shared memory with the diagonal; a warp owns
// an item, its lanes stride over the packed edges (coalesced shared-memory reads, two gathers of the item row per
// edge) -- half the edges of the row-wise kernel above and no per-row loop overhead.  tau's median / percentile is a
// quantised selection: min / max of the candidates, a 256-bin shared-memory histogram and a scan for the bin that
// holds the rank, recurse into that bin (usually <= 2 rounds for a few hundred entries), exact extraction at the end.
// Sums are reordered with respect to the reference's left folds (as in the kernel above): within 1e-12 relative.
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// STRIDE: doubles between consecutive entries (1: a row; ITEMS + 1: one item of a transposed tile)
// value of 0-based rank `rank` among the finite entries of xs[0..f); *n_le = number of finite entries <= that value
template <int STRIDE = 1>
__device__ double warp_kth_finite(const double* xs, uint32_t f, uint32_t rank, int lane, uint32_t* n_le, uint32_t* hist /* 256 u32 of this warp */) {
    double lo = -INFINITY, hi = INFINITY;   // candidates: finite v with lo <= v <= hi
    uint32_t below = 0;                     // finite entries < lo
    double mn = INFINITY, mx = -INFINITY;
    uint32_t cnt = 0;
    for (uint32_t t = lane; t < f; t += 32) { const double v = xs[(size_t)t * STRIDE]; if (isfinite(v)) { mn = fmin(mn, v); mx = fmax(mx, v); ++cnt; } }
    mn = warp_min_d(mn); mx = warp_max_d(mx); cnt = warp_sum_u(cnt);
    for (int iter = 0; iter < 40; ++iter) {
        lo = mn; hi = mx;
        const uint32_t target = rank - below;
        if (mn == mx) { *n_le = below + cnt; return mn; }
        const double scale = 256.0 / (mx - mn);
        if (cnt <= 8 || !(scale < 1e300) || iter == 39) break;
        // 256-bin histogram of the candidates in this warp's shared scratch, then a scan for the bin that holds `target`
        for (int b = lane; b < 256; b += 32) hist[b] = 0;
        __syncwarp();
        for (uint32_t t = lane; t < f; t += 32) {
            const double v = xs[(size_t)t * STRIDE];
            if (isfinite(v) && v >= lo && v <= hi) { int q = (int)((v - lo) * scale); q = q > 255 ? 255 : q; atomicAdd(&hist[q], 1u); }
        }
        __syncwarp();
        uint32_t c8[8], sum8 = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { c8[b] = hist[lane * 8 + b]; sum8 += c8[b]; }
        uint32_t inc = sum8;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        uint32_t run = inc - sum8;   // candidates in the bins of lower lanes
        int my_cut = -1; uint32_t my_below = 0;
        if (target >= run && target < run + sum8) {
#pragma unroll
            for (int b = 0; b < 8; ++b) { if (my_cut < 0 && target < run + c8[b]) { my_cut = lane * 8 + b; my_below = run; } run += c8[b]; }
        }
        const int src = __ffs(__ballot_sync(FULL, my_cut >= 0)) - 1;
        const uint32_t cut = (uint32_t)__shfl_sync(FULL, my_cut, src), below_cut = __shfl_sync(FULL, my_below, src);
        __syncwarp();
        // recurse into bin `cut`: its min, max and population
        double nmn = INFINITY, nmx = -INFINITY; uint32_t ncnt = 0;
        for (uint32_t t = lane; t < f; t += 32) {
            const double v = xs[(size_t)t * STRIDE];
            if (isfinite(v) && v >= lo && v <= hi) {
                int q = (int)((v - lo) * scale); q = q > 255 ? 255 : q;
                if ((uint32_t)q == cut) { nmn = fmin(nmn, v); nmx = fmax(nmx, v); ++ncnt; }
            }
        }
        mn = warp_min_d(nmn); mx = warp_max_d(nmx); cnt = warp_sum_u(ncnt);
        below += below_cut;
    }
    // exact tail: walk the distinct candidate values upwards
    const uint32_t target = rank - below;
    uint32_t seen = 0;
    double last = -INFINITY;
    bool first = true;
    for (;;) {
        double cand = INFINITY;
        for (uint32_t t = lane; t < f; t += 32) { const double v = xs[(size_t)t * STRIDE]; if (isfinite(v) && v >= lo && v <= hi && (first || v > last)) cand = fmin(cand, v); }
        cand = warp_min_d(cand);
        uint32_t mult = 0;
        for (uint32_t t = lane; t < f; t += 32) mult += xs[(size_t)t * STRIDE] == cand ? 1u : 0u;
        mult = warp_sum_u(mult);
        if (target < seen + mult || !(cand < INFINITY)) { *n_le = below + seen + mult; return cand; }
        seen += mult; last = cand; first = false;
    }
}

// TauMode::select_tau (taumode.rs:29-70) with the quantised selection
template <int STRIDE = 1>
__device__ double warp_select_tau_fast(const double* xs, uint32_t f, int mode, double value, int lane, uint32_t* hist) {
    const double FLOOR = 1e-10;
    if (mode == SFB_TAU_FIXED) return (isfinite(value) && value > 0.0) ? value : FLOOR;
    uint32_t n = 0; double s = 0.0;
    for (uint32_t t = lane; t < f; t += 32) { double v = xs[(size_t)t * STRIDE]; if (isfinite(v)) { ++n; s += v; } }
    n = warp_sum_u(n);
    if (n == 0) return FLOOR;
    if (mode == SFB_TAU_MEAN) { double mean = warp_sum(s) / (double)n; return mean > FLOOR ? mean : FLOOR; }
    uint32_t n_le;
    double r;
    if (mode == SFB_TAU_PERCENTILE) {
        double pp = value < 0.0 ? 0.0 : (value > 1.0 ? 1.0 : value);
        r = warp_kth_finite<STRIDE>(xs, f, (uint32_t)round((double)(n - 1) * pp), lane, &n_le, hist);
    } else if (n & 1u) {
        r = warp_kth_finite<STRIDE>(xs, f, n / 2, lane, &n_le, hist);
    } else {
        const double a = warp_kth_finite<STRIDE>(xs, f, n / 2 - 1, lane, &n_le, hist);
        double b = a;
        if (n_le <= n / 2) {   // rank n/2 is the next larger value
            b = INFINITY;
            for (uint32_t t = lane; t < f; t += 32) { double v = xs[(size_t)t * STRIDE]; if (isfinite(v) && v > a) b = fmin(b, v); }
            b = warp_min_d(b);
        }
        r = 0.5 * (a + b);
    }
    return r > FLOOR ? r : FLOOR;
}

// strict upper triangle of a symmetric CSR, packed in CSR order (deterministic), + the diagonal
__global__ void upper_count_kernel(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, uint32_t f, uint32_t* __restrict__ cnt) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= f) return;
    uint32_t c = 0;
    for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) c += indices[e] > r ? 1u : 0u;
    cnt[r] = c;
}
__global__ void upper_pack_kernel(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, const double* __restrict__ data,
                                  uint32_t f, const uint64_t* __restrict__ offs, uint32_t* __restrict__ rc, double* __restrict__ val,
                                  double* __restrict__ diag) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= f) return;
    uint64_t o = offs[r];
    double d = 0.0, fold = 0.0;
    for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
        const uint32_t c = indices[e];
        if (c > r) { rc[o] = (r << 16) | c; val[o] = data[e]; ++o; }
        if (c == r) d = data[e];
        else fold = __dadd_rn(fold, -data[e]);   // the fold that produced L_rr in the builder (ascending column, from 0.0)
    }
    diag[r] = __dadd_rn(d, -fold);   // row-sum defect: exactly 0 for a Laplacian assembled as D - W
}
// symmetric bit for bit?  every stored (r, c) must have a stored (c, r) with the same bits
__global__ void csr_symmetric_kernel(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, const double* __restrict__ data,
                                     uint64_t rows, int* __restrict__ bad) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
        const uint32_t c = indices[e];
        if (c == r) continue;
        if (c >= rows) { *bad = 1; return; }
        uint64_t a = indptr[c], b = indptr[c + 1];
        while (a < b) { uint64_t mid = (a + b) >> 1; if (indices[mid] < r) a = mid + 1; else b = mid; }
        if (a >= indptr[c + 1] || indices[a] != r || __double_as_longlong(data[a]) != __double_as_longlong(data[e])) { *bad = 1; return; }
    }
}

struct LambdaSymArgs {
    const uint32_t* rc; const double* val; const double* diag; uint32_t ne, f;
    const double* x; uint64_t n;
    int tau_mode; double tau_value;
    double* out_lambda; double* out_disp;
    const double* tau_in;   // as in LambdaArgs
};

template <int VARIANT>
__global__ void __launch_bounds__(256) lambda_sym_kernel(LambdaSymArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* s_val = reinterpret_cast<double*>(smem_raw);                 // [ne]
    double* s_diag = s_val + a.ne;                                       // [f]
    double* xs = s_diag + a.f + (size_t)w * a.f;                         // [wpb][f]
    uint32_t* s_rc = reinterpret_cast<uint32_t*>(s_diag + a.f + (size_t)wpb * a.f);   // [ne]
    uint32_t* hist = s_rc + a.ne + (size_t)w * 256;                                   // [wpb][256]
    for (uint32_t e = threadIdx.x; e < a.ne; e += blockDim.x) { s_val[e] = a.val[e]; s_rc[e] = a.rc[e]; }
    for (uint32_t r = threadIdx.x; r < a.f; r += blockDim.x) s_diag[r] = a.diag[r];
    __syncthreads();
    const uint32_t f = a.f;
    for (uint64_t i = (uint64_t)blockIdx.x * wpb + w; i < a.n; i += (uint64_t)gridDim.x * wpb) {
        const double* xr = a.x + i * f;
        __syncwarp();
        bool zero = true;
        double den = 0.0, num = 0.0;
        for (uint32_t t = lane; t < f; t += 32) {
            const double v = xr[t];
            xs[t] = v;
            zero = zero && (fabs(v) <= 1e-10);
            den += v * v;
            num += (v * s_diag[t]) * v;   // row-sum defect term
        }
        __syncwarp();
        zero = a.tau_in ? a.tau_in[i] < 0.0 : __all_sync(FULL, zero);   // projected items: the test was made on the unprojected vector
        if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE && zero) {  // taumode.rs:268-274
            if (lane == 0) { a.out_lambda[i] = 0.0; if (a.out_disp) a.out_disp[i] = 0.0; }
            continue;
        }
        double sall = 0.0, ssum = 0.0, qsum = 0.0;
        for (uint32_t e = lane; e < a.ne; e += 32) {
            const uint32_t rc = s_rc[e];
            const double xa = xs[rc >> 16], xb = xs[rc & 0xFFFFu], wv = -s_val[e];
            const double d = xa - xb;
            const double contrib = (wv * d) * d;
            sall += contrib;
            if (wv > 0.0) {
                ssum += contrib;
                qsum = fma(contrib, contrib, qsum);
            }
        }
        den = warp_sum(den); num = warp_sum(num) + warp_sum(sall);
        ssum = warp_sum(ssum); qsum = warp_sum(qsum);
        if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) { ssum *= 2.0; qsum *= 2.0; }   // both triangles (taumode.rs:371-383)
        double e_raw = 0.0;
        if (den > 1e-12) { e_raw = num / den; if (!(e_raw > 0.0)) e_raw = 0.0; }
        double g = 0.0;
        // a NaN sum: the taumode form tests `<= 1e-12 -> 0` and lets it propagate through f64::clamp (taumode.rs:385-407), the energy form tests `> 1e-12` (energymaps.rs:1006)
        if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE ? !(ssum <= 1e-12) : (ssum > 1e-12)) { g = qsum / (ssum * ssum); g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g); }
        double lam;
        if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) {
            const double tau = a.tau_in ? a.tau_in[i] : warp_select_tau_fast(xs, f, a.tau_mode, a.tau_value, lane, hist);
            lam = tau * (e_raw / (e_raw + tau)) + (1.0 - tau) * g;  // taumode.rs:306-310
        } else lam = e_raw;
        if (lane == 0) { a.out_lambda[i] = lam; if (a.out_disp) a.out_disp[i] = g; }
    }
}

// ---- tile kernel: lane = item -------------------------------------------------------------------------------
// The warp-per-item kernels above gather x[r], x[c] from shared memory with one lane per EDGE: 28 bytes of shared-memory
// traffic per item-edge and multi-way bank conflicts on the gathers (ncu, C2: 2.3e8 conflicts, issue slots 71 %, DRAM 5 %
// of peak).  Here a CTA stages a TILE of 32 items transposed (xs[feature][item], row stride 33 doubles) and the edge loop
// runs with lane = ITEM: the edge record (weight, offsets) is the same for the whole warp (one broadcast load), x[c] of
// 32 items is one conflict-free 256-byte read, and x[r] stays in a register across the run of edges that share row r:
// 9 bytes of shared-memory traffic per item-edge and 5 FP64 instructions (d, w*d, (w*d)*d, two accumulations).  The
// eight warps of a CTA split the packed upper-triangle edge list evenly; partial sums meet in shared memory in warp
// order (deterministic).  At C2 (F = 384, ~3k upper edges) the kernel is bound by the FP64 pipe: 1.5e10 DP instructions
// = 0.8 ms at full rate against 0.47 ms of HBM time for the 3.07 GB it reads.
//
// Phase 1 (one warp per item, coalesced 256-byte loads, the row in REGISTERS, E = ceil(F / 32) values per lane): zero
// test, |x|^2, row-sum-defect term, tau (median / percentile by one 256-bin quantised histogram + an exact ranking of
// the two bins that hold the wanted ranks), then the transposed store.  Phase 2: the edge loop.  Phase 3: warp 0
// combines, blends and writes 32 consecutive lambdas; the running min / max feed the normalisation without a second pass.
constexpr int LT_TS = 33;          // tile row stride in doubles (odd: the transposing store is conflict-free)

// w = -L_rc; coff / roff = BYTE offsets of x_c / x_r in the tile (index * LT_TS * 8).  First edge of a row: roff bit 31 set, and --
// when every weight of the matrix is positive (meta.any_nonpos == 0: the Laplacians built here) -- w stored NEGATED, so the edge
// loop tests one sign bit and multiplies by |w| (an operand modifier) instead of masking a flag out of an offset.
struct __align__(16) EdgeRec { double w; uint32_t coff; uint32_t roff; };

struct PackMeta { uint32_t ne; uint32_t any_defect; uint32_t any_nonpos; uint32_t pad; };

// one CTA, thread = a run of ceil(F / 1024) consecutive rows: strict upper triangle in CSR order -> EdgeRec list, row-sum
// defects, flags.  stride_bytes: bytes between consecutive features of the staged tile (LT_TS * 8, or (ITEMS + 1) * 8 for
// the narrow tiles of lambda_tile_narrow_kernel)
__global__ void __launch_bounds__(1024) lambda_pack_kernel(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                                                           const double* __restrict__ data, uint32_t f, uint32_t stride_bytes, EdgeRec* __restrict__ recs,
                                                           double* __restrict__ defect, PackMeta* __restrict__ meta) {
    __shared__ uint32_t s_scan[1024];
    __shared__ uint32_t s_flags[2];
    const uint32_t t = threadIdx.x, rpt = (f + 1023) / 1024;
    const uint32_t r_lo = t * rpt < f ? t * rpt : f, r_hi = r_lo + rpt < f ? r_lo + rpt : f;
    if (t < 2) s_flags[t] = 0;
    uint32_t c_up = 0;
    for (uint32_t r = r_lo; r < r_hi; ++r)
        for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) c_up += indices[e] > r ? 1u : 0u;
    s_scan[t] = c_up;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan
        uint32_t v = t >= o ? s_scan[t - o] : 0u;
        __syncthreads();
        s_scan[t] += v;
        __syncthreads();
    }
    uint32_t o = s_scan[t] - c_up;
    for (uint32_t r = r_lo; r < r_hi; ++r) {
        double d = 0.0, fold = 0.0;
        bool first = true, nonpos = false;
        for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
            const uint32_t c = indices[e];
            if (c > r) {
                EdgeRec rec; rec.w = -data[e]; rec.coff = c * stride_bytes; rec.roff = r * stride_bytes | (first ? 0x80000000u : 0u);
                recs[o++] = rec; first = false;
                nonpos = nonpos || !(rec.w > 0.0);
            }
            if (c == r) d = data[e];
            else fold = __dadd_rn(fold, -data[e]);   // the fold that produced L_rr in the builder (ascending column, from 0.0)
        }
        const double df = __dadd_rn(d, -fold);       // row-sum defect: exactly 0 for a Laplacian assembled as D - W
        defect[r] = df;
        if (df != 0.0) s_flags[0] = 1;
        if (nonpos) s_flags[1] = 1;
    }
    __syncthreads();
    if (t == 0) { meta->ne = s_scan[1023]; meta->any_defect = s_flags[0]; meta->any_nonpos = s_flags[1]; meta->pad = 0; }
    if (!s_flags[1]) {   // every weight positive: the row flag becomes the sign of w
        uint32_t q = s_scan[t] - c_up;
        for (uint32_t r = r_lo; r < r_hi; ++r) {
            uint32_t cu = 0;
            for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) cu += indices[e] > r ? 1u : 0u;
            if (cu) { EdgeRec* first = recs + q; first->w = -first->w; first->roff &= 0x7FFFFFFFu; }
            q += cu;
        }
    }
}

__device__ __forceinline__ uint32_t f32_key(float v) { uint32_t u = __float_as_uint(v); return (u >> 31) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key_f32(uint32_t k) { return __uint_as_float((k >> 31) ? (k & 0x7FFFFFFFu) : ~k); }

// Values of 0-based ranks k1 <= k2 <= k1 + 1 among the finite entries of v[] (E per lane, entry u of lane l is element
// l + 32u, valid while < f).  scratch: 128 u32 of this warp.  Falls back (returns false) when the quantised histogram cannot
// separate the candidates (all equal in f32, overflowing range, more than 64 values in the target bins).
template <int E>
__device__ __forceinline__ bool regs_select2(const double (&v)[E], uint32_t f, int lane, uint32_t k1, uint32_t k2, uint32_t* scratch,
                                             double* out1, double* out2) {
    float a[E];
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
    for (int u = 0; u < E; ++u) {
        const bool ok = (uint32_t)(lane + 32 * u) < f && isfinite(v[u]);
        a[u] = fminf(fmaxf(__double2float_rn(v[u]), -3.4028234e38f), 3.4028234e38f);
        if (ok) { const uint32_t kk = f32_key(a[u]); kmin = min(kmin, kk); kmax = max(kmax, kk); }
    }
    kmin = __reduce_min_sync(FULL, kmin); kmax = __reduce_max_sync(FULL, kmax);
    const float lo = key_f32(kmin), hi = key_f32(kmax);
    const float scale = 255.0f / (hi - lo);
    if (!(scale > 0.0f) || !(scale < 3.0e38f)) return false;
    // 256 bins, 16-bit counters packed two to a word
#pragma unroll
    for (int b = 0; b < 4; ++b) scratch[lane * 4 + b] = 0u;
    __syncwarp();
    uint32_t q[E];
#pragma unroll
    for (int u = 0; u < E; ++u) {
        const bool ok = (uint32_t)(lane + 32 * u) < f && isfinite(v[u]);
        q[u] = 0xFFFFu;
        if (ok) {
            int qq = (int)((a[u] - lo) * scale); qq = qq < 0 ? 0 : (qq > 255 ? 255 : qq);
            q[u] = (uint32_t)qq;
            atomicAdd(&scratch[qq >> 1], 1u << ((qq & 1) * 16));
        }
    }
    __syncwarp();
    const uint4 wv = *reinterpret_cast<const uint4*>(&scratch[lane * 4]);   // bins 8*lane .. 8*lane + 7
    const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t c8[8], s8 = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) { c8[b] = (ww[b >> 1] >> ((b & 1) * 16)) & 0xFFFFu; s8 += c8[b]; }
    uint32_t inc = s8;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    const uint32_t ex = inc - s8;
    // the bin of rank k1 (and of rank k2), with the number of entries in lower bins
    int bin1 = -1, bin2 = -1; uint32_t below1 = 0, cnt1 = 0, cnt2 = 0;
    {
        uint32_t run = ex;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (bin1 < 0 && k1 >= run && k1 < run + c8[b]) { bin1 = lane * 8 + b; below1 = run; cnt1 = c8[b]; }
            if (bin2 < 0 && k2 >= run && k2 < run + c8[b]) { bin2 = lane * 8 + b; cnt2 = c8[b]; }
            run += c8[b];
        }
    }
    const uint32_t m1 = __ballot_sync(FULL, bin1 >= 0), m2 = __ballot_sync(FULL, bin2 >= 0);
    if (!m1 || !m2) return false;   // rank beyond the population (cannot happen for k < n)
    const int l1 = __ffs(m1) - 1, l2 = __ffs(m2) - 1;
    bin1 = __shfl_sync(FULL, bin1, l1); below1 = __shfl_sync(FULL, below1, l1); cnt1 = __shfl_sync(FULL, cnt1, l1);
    bin2 = __shfl_sync(FULL, bin2, l2); cnt2 = __shfl_sync(FULL, cnt2, l2);
    const uint32_t nc = cnt1 + (bin2 != bin1 ? cnt2 : 0u);
    if (nc > 62) return false;
    __syncwarp();
    // candidates of the two bins -> scratch as doubles [0, 62), slot counter in word 126
    double* cand = reinterpret_cast<double*>(scratch);
    if (lane == 0) scratch[126] = 0u;
    __syncwarp();
#pragma unroll
    for (int u = 0; u < E; ++u)
        if (q[u] == (uint32_t)bin1 || q[u] == (uint32_t)bin2) cand[atomicAdd(&scratch[126], 1u)] = v[u];
    __syncwarp();
    // exact rank of every candidate (ties by slot: equal values are interchangeable)
    const uint32_t t1 = k1 - below1, t2 = k2 - below1;
    double r1 = 0.0, r2 = 0.0; bool h1 = false, h2 = false;
    for (uint32_t j = lane; j < nc; j += 32) {
        const double cj = cand[j];
        uint32_t rk = 0;
        for (uint32_t i = 0; i < nc; ++i) { const double ci = cand[i]; rk += (ci < cj || (ci == cj && i < j)) ? 1u : 0u; }
        if (rk == t1) { r1 = cj; h1 = true; }
        if (rk == t2) { r2 = cj; h2 = true; }
    }
    const uint32_t b1 = __ballot_sync(FULL, h1), b2 = __ballot_sync(FULL, h2);
    if (!b1 || !b2) return false;
    *out1 = __shfl_sync(FULL, r1, __ffs(b1) - 1);
    *out2 = __shfl_sync(FULL, r2, __ffs(b2) - 1);
    __syncwarp();
    return true;
}

// The common case of the selection above, stripped of everything it does not need: every lane slot holds a finite value
// (f == 32 E, no NaN / inf in the row).  One pass: f32 images -> min / max -> 256 linear bins counted with shared-memory
// atomics (one word per bin: the warp's 1 KB scratch) -> the bin that holds rank k1 -> exact ranking of that bin's
// members; rank k2 = k1 + 1 is the next member of the bin, or, when k1 is the bin's last, the smallest value above it.
// Returns false (nothing written) when the histogram cannot separate the candidates.
template <int E>
__device__ __forceinline__ bool regs_select2_fast(const double (&v)[E], int lane, uint32_t k1, uint32_t k2, uint32_t* scratch, double* out1, double* out2) {
    float a[E];
    float fmn = INFINITY, fmx = -INFINITY;
#pragma unroll
    for (int u = 0; u < E; ++u) { a[u] = __double2float_rn(v[u]); fmn = fminf(fmn, a[u]); fmx = fmaxf(fmx, a[u]); }
    const float lo = key_f32(__reduce_min_sync(FULL, f32_key(fmn))), hi = key_f32(__reduce_max_sync(FULL, f32_key(fmx)));
    const float scale = 255.0f / (hi - lo);
    if (!(scale > 0.0f) || !(scale < 3.0e38f)) return false;   // all equal in f32, or a range that overflows f32
    uint4* z = reinterpret_cast<uint4*>(scratch);
    z[lane * 2] = make_uint4(0u, 0u, 0u, 0u); z[lane * 2 + 1] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    int q[E];
#pragma unroll
    for (int u = 0; u < E; ++u) { q[u] = min((int)((a[u] - lo) * scale), 255); atomicAdd(&scratch[q[u]], 1u); }
    __syncwarp();
    const uint4 w0 = z[lane * 2], w1 = z[lane * 2 + 1];   // bins 8 * lane .. 8 * lane + 7
    const uint32_t c8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    uint32_t s8 = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) s8 += c8[b];
    uint32_t inc = s8;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    const uint32_t ex = inc - s8;
    const bool mine = k1 >= ex && k1 < inc;
    int bin1 = 0; uint32_t below1 = 0, cnt1 = 0;
    if (mine) {
        uint32_t run = ex;
#pragma unroll
        for (int b = 0; b < 8; ++b) { if (k1 >= run && k1 < run + c8[b]) { bin1 = lane * 8 + b; below1 = run; cnt1 = c8[b]; } run += c8[b]; }
    }
    const int src = __ffs(__ballot_sync(FULL, mine)) - 1;
    bin1 = __shfl_sync(FULL, bin1, src); below1 = __shfl_sync(FULL, below1, src); cnt1 = __shfl_sync(FULL, cnt1, src);
    if (cnt1 > 64) return false;
    __syncwarp();
    double* cand = reinterpret_cast<double*>(scratch);   // [0, 64): words 0 .. 127; slot counter in word 128
    if (lane == 0) scratch[128] = 0u;
    __syncwarp();
#pragma unroll
    for (int u = 0; u < E; ++u) if (q[u] == bin1) cand[atomicAdd(&scratch[128], 1u)] = v[u];
    __syncwarp();
    const uint32_t t1 = k1 - below1;
    const bool need2 = k2 != k1, same_bin = t1 + 1 < cnt1;
    double r1 = 0.0, r2 = 0.0; bool h1 = false, h2 = false;
    for (uint32_t j = lane; j < cnt1; j += 32) {
        const double cj = cand[j];
        uint32_t rk = 0;
        for (uint32_t i = 0; i < cnt1; ++i) { const double ci = cand[i]; rk += (ci < cj || (ci == cj && i < j)) ? 1u : 0u; }
        if (rk == t1) { r1 = cj; h1 = true; }
        if (rk == t1 + 1) { r2 = cj; h2 = true; }
    }
    const double va = __shfl_sync(FULL, r1, __ffs(__ballot_sync(FULL, h1)) - 1);
    double vb = va;
    if (need2) {
        if (same_bin) vb = __shfl_sync(FULL, r2, __ffs(__ballot_sync(FULL, h2)) - 1);
        else {   // the next larger value lives in a later bin
            double nx = INFINITY;
#pragma unroll
            for (int u = 0; u < E; ++u) if (q[u] > bin1) nx = fmin(nx, v[u]);
            vb = warp_min_d(nx);
        }
    }
    __syncwarp();
    *out1 = va; *out2 = vb;
    return true;
}

// exact selection of rank k among the finite register entries by bisection on the order-preserving 64-bit image (rare path)
template <int E>
__device__ double regs_select_exact(const double (&v)[E], uint32_t f, int lane, uint32_t k) {
    uint64_t prefix = 0;
    for (int b = 63; b >= 0; --b) {
        const uint64_t trial = prefix | (1ull << b);
        uint32_t c = 0;   // entries with key < trial
#pragma unroll
        for (int u = 0; u < E; ++u) {
            const bool ok = (uint32_t)(lane + 32 * u) < f && isfinite(v[u]);
            c += (ok && sort_key(v[u]) < trial) ? 1u : 0u;
        }
        c = __reduce_add_sync(FULL, c);
        if (c <= k) prefix = trial;   // at most k entries lie below trial: the rank-k key is >= trial
    }
    return key_value(prefix);
}

// TauMode::select_tau (taumode.rs:29-70) on a row held in registers
template <int E>
__device__ __forceinline__ double regs_select_tau(const double (&v)[E], uint32_t f, int lane, int mode, double value, uint32_t* scratch) {
    const double FLOOR = 1e-10;
    if (mode == SFB_TAU_FIXED) return (isfinite(value) && value > 0.0) ? value : FLOOR;
    if (mode != SFB_TAU_MEAN && f == 32u * E) {
        // full lanes: is every entry finite?  (exponent field of the high word below 0x7FF)
        bool fin = true;
#pragma unroll
        for (int u = 0; u < E; ++u) fin = fin && ((uint32_t)__double2hiint(v[u]) & 0x7FF00000u) != 0x7FF00000u;
        if (__all_sync(FULL, fin)) {
            uint32_t k1, k2;
            if (mode == SFB_TAU_PERCENTILE) {
                const double pp = value < 0.0 ? 0.0 : (value > 1.0 ? 1.0 : value);
                k1 = k2 = (uint32_t)round((double)(f - 1) * pp);
            } else { k1 = f / 2 - 1; k2 = f / 2; }   // f = 32 E is even
            double a, b;
            if (regs_select2_fast<E>(v, lane, k1, k2, scratch, &a, &b)) {
                const double r = k1 == k2 ? a : 0.5 * (a + b);
                return r > FLOOR ? r : FLOOR;
            }
        }
    }
    uint32_t n = 0; double s = 0.0;
#pragma unroll
    for (int u = 0; u < E; ++u) if ((uint32_t)(lane + 32 * u) < f && isfinite(v[u])) { ++n; s += v[u]; }
    n = __reduce_add_sync(FULL, n);
    if (n == 0) return FLOOR;
    if (mode == SFB_TAU_MEAN) { const double mean = warp_sum(s) / (double)n; return mean > FLOOR ? mean : FLOOR; }
    uint32_t k1, k2;
    if (mode == SFB_TAU_PERCENTILE) {
        const double pp = value < 0.0 ? 0.0 : (value > 1.0 ? 1.0 : value);
        k1 = k2 = (uint32_t)round((double)(n - 1) * pp);
    } else if (n & 1u) k1 = k2 = n / 2;
    else { k1 = n / 2 - 1; k2 = n / 2; }
    double a, b;
    if (!regs_select2<E>(v, f, lane, k1, k2, scratch, &a, &b)) {
        a = regs_select_exact<E>(v, f, lane, k1);
        b = k2 == k1 ? a : regs_select_exact<E>(v, f, lane, k2);
    }
    const double r = k1 == k2 ? a : 0.5 * (a + b);
    return r > FLOOR ? r : FLOOR;
}

struct LambdaTileArgs {
    const EdgeRec* recs; const double* defect; const PackMeta* meta; uint32_t f;
    const double* x; uint64_t n;
    int tau_mode; double tau_value;
    double* out_lambda; double* out_disp;
    const double* tau_in;
    unsigned long long* minmax;   // [0]: min of lambda, [1]: max(0, lambda), as order-preserving keys (atomicMin / atomicMax); may be null
};

__device__ __forceinline__ double lt_ld(const double* lane_base, uint32_t byte_off) {
    return *reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(lane_base) + byte_off);
}
__device__ __forceinline__ void lt_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// shared memory: xs [f][LT_TS] | den, defect term, tau [3][32] | one 1 KB region per warp: the tau selection's scratch in
// phase 1, the ring of edge records in phase 2 (2 x 32 records, filled by cp.async one chunk ahead), the warp's partial sums
// at the end of phase 2
// FULL: f == 32 E (every lane slot holds an entry): the bounds tests of phase 1 compile away
template <int VARIANT, int E, int LT_WARPS, bool FULL>
__global__ void __launch_bounds__(LT_WARPS * 32, LT_WARPS == 8 ? (E <= 4 ? 4 : 2) : 1) lambda_tile_kernel(LambdaTileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t f = FULL ? 32u * E : a.f;
    double* xs = reinterpret_cast<double*>(smem_raw);                         // [f][LT_TS]
    double* s_den = xs + (((size_t)f * LT_TS + 1) & ~(size_t)1);              // [32]   (16-byte aligned)
    double* s_dfc = s_den + 32;                                               // [32]
    double* s_tau = s_dfc + 32;                                               // [32]  < 0: zero vector
    unsigned char* region0 = reinterpret_cast<unsigned char*>(s_tau + 32);
    unsigned char* region = region0 + (size_t)w * 1024;
    uint32_t* scratch = reinterpret_cast<uint32_t*>(region);                  // phase 1
    uint4* ring = reinterpret_cast<uint4*>(region);                           // phase 2: [2][32] records
    double* part = reinterpret_cast<double*>(region);                         // end of phase 2: [3][32]
    const uint32_t ne = a.meta->ne;
    const bool has_defect = a.meta->any_defect != 0, nonpos = a.meta->any_nonpos != 0;
    const uint32_t e_lo = (uint32_t)((uint64_t)ne * w / LT_WARPS), e_hi = (uint32_t)((uint64_t)ne * (w + 1) / LT_WARPS);
    const uint32_t n_chunks = (e_hi - e_lo + 31) / 32;
    const double* xs_lane = xs + lane;
    double run_mn = INFINITY, run_mx = 0.0;
    const uint64_t n_tiles = (a.n + 31) / 32;
    constexpr int IPW = 32 / LT_WARPS;   // items per warp in phase 1
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t i0 = tile * 32;
        // ---- phase 1: this warp's items, the next item's loads in flight while the current one is processed
        constexpr bool PREFETCH = E <= 16;   // E = 24 runs 512 threads per CTA: 128 registers leave no room for a second row
        double vn[PREFETCH ? E : 1];
        if (PREFETCH) {
            const uint64_t i = i0 + w * IPW;
#pragma unroll
            for (int u = 0; u < E; ++u) { const uint32_t t = lane + 32 * u; vn[PREFETCH ? u : 0] = (i < a.n && (FULL || t < f)) ? __ldcs(a.x + i * f + t) : 0.0; }
        }
#pragma unroll 1
        for (int qi = 0; qi < IPW; ++qi) {
            const int it = w * IPW + qi;
            const uint64_t i = i0 + it;
            double v[E];
            if (PREFETCH) {
#pragma unroll
                for (int u = 0; u < E; ++u) v[u] = vn[PREFETCH ? u : 0];
                if (qi + 1 < IPW) {
                    const uint64_t i2 = i + 1;
#pragma unroll
                    for (int u = 0; u < E; ++u) { const uint32_t t = lane + 32 * u; vn[PREFETCH ? u : 0] = (i2 < a.n && (FULL || t < f)) ? __ldcs(a.x + i2 * f + t) : 0.0; }
                }
            } else {
#pragma unroll
                for (int u = 0; u < E; ++u) { const uint32_t t = lane + 32 * u; v[u] = (i < a.n && (FULL || t < f)) ? __ldcs(a.x + i * f + t) : 0.0; }
            }
            // zero test |v| <= 1e-10 for every entry (taumode.rs:268-274) as one running maximum (a NaN entry fails it, as in
            // the reference: max keeps the NaN out, so test it apart)
            double amax = 0.0, den = 0.0, dfc = 0.0;
            bool any_nan = false;
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const uint32_t t = lane + 32 * u;
                if (FULL || t < f) {
                    xs[(size_t)t * LT_TS + it] = v[u];
                    amax = fmax(amax, fabs(v[u]));
                    any_nan = any_nan || v[u] != v[u];
                    den = fma(v[u], v[u], den);
                }
            }
            const bool zero = amax <= 1e-10 && !any_nan;
            if (has_defect) {
#pragma unroll
                for (int u = 0; u < E; ++u) { const uint32_t t = lane + 32 * u; if (FULL || t < f) dfc += (v[u] * a.defect[t]) * v[u]; }
                dfc = warp_sum(dfc);
            }
            den = warp_sum(den);
            double tau = 0.0;
            if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) {
                if (a.tau_in) tau = i < a.n ? a.tau_in[i] : -1.0;           // projected items: tau and the zero test come from the unprojected rows
                else if (__all_sync(FULL, zero)) tau = -1.0;                // taumode.rs:268-274
                else tau = regs_select_tau<E>(v, f, lane, a.tau_mode, a.tau_value, scratch);
            }
            if (lane == 0) { s_den[it] = den; s_dfc[it] = dfc; s_tau[it] = tau; }
        }
        __syncthreads();
        // ---- phase 2: lane = item, this warp's slice of the edge list, records one chunk ahead in the ring
        double s0 = 0.0, q0 = 0.0, s1 = 0.0, q1 = 0.0, sa = 0.0;
        if (n_chunks) {
            { const uint32_t e = e_lo + lane; if (e < e_hi) lt_cp_async16(&ring[lane], a.recs + e); asm volatile("cp.async.commit_group;" ::: "memory"); }
            { const uint32_t e = e_lo + 32 + lane; if (e < e_hi) lt_cp_async16(&ring[32 + lane], a.recs + e); asm volatile("cp.async.commit_group;" ::: "memory"); }
            double xa = 0.0;
            for (uint32_t c = 0; c < n_chunks; ++c) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncwarp();
                const uint4* rr = ring + (c & 1) * 32;
                const uint32_t cnt = min(32u, e_hi - e_lo - c * 32);
                if (c == 0) xa = lt_ld(xs_lane, rr[0].w & 0x7FFFFFFFu);
                if (!nonpos) {
                    // every weight positive: a negative w marks the first edge of a row
                    uint32_t j = 0;
#pragma unroll 2
                    for (; j + 2 <= cnt; j += 2) {
                        const uint4 ra = rr[j], rb = rr[j + 1];
                        if ((int)ra.y < 0) xa = lt_ld(xs_lane, ra.w);
                        const double d0 = xa - lt_ld(xs_lane, ra.z);
                        if ((int)rb.y < 0) xa = lt_ld(xs_lane, rb.w);
                        const double d1 = xa - lt_ld(xs_lane, rb.z);
                        const double c0 = (fabs(__hiloint2double((int)ra.y, (int)ra.x)) * d0) * d0;
                        const double c1 = (fabs(__hiloint2double((int)rb.y, (int)rb.x)) * d1) * d1;
                        s0 += c0; q0 = fma(c0, c0, q0);
                        s1 += c1; q1 = fma(c1, c1, q1);
                    }
                    if (j < cnt) {
                        const uint4 ra = rr[j];
                        if ((int)ra.y < 0) xa = lt_ld(xs_lane, ra.w);
                        const double d0 = xa - lt_ld(xs_lane, ra.z);
                        const double c0 = (fabs(__hiloint2double((int)ra.y, (int)ra.x)) * d0) * d0;
                        s0 += c0; q0 = fma(c0, c0, q0);
                    }
                } else {
                    for (uint32_t j = 0; j < cnt; ++j) {
                        const uint4 ra = rr[j];
                        if (ra.w & 0x80000000u) xa = lt_ld(xs_lane, ra.w & 0x7FFFFFFFu);
                        const double wv = __hiloint2double((int)ra.y, (int)ra.x);
                        const double d0 = xa - lt_ld(xs_lane, ra.z);
                        const double c0 = (wv * d0) * d0;
                        sa += c0;
                        if (wv > 0.0) { s0 += c0; q0 = fma(c0, c0, q0); }
                    }
                }
                __syncwarp();
                { const uint32_t e = e_lo + (c + 2) * 32 + lane; if (e < e_hi) lt_cp_async16(&ring[(c & 1) * 32 + lane], a.recs + e); asm volatile("cp.async.commit_group;" ::: "memory"); }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        s0 += s1; q0 += q1;
        if (!nonpos) sa = s0;
        part[lane] = s0; part[32 + lane] = q0; part[64 + lane] = sa;
        __syncthreads();
        // ---- phase 3: combine, blend, write
        if (w == 0) {
            const uint64_t i = i0 + lane;
            double ssum = 0.0, qsum = 0.0, sall = 0.0;
#pragma unroll
            for (int ww = 0; ww < LT_WARPS; ++ww) {
                const double* pw = reinterpret_cast<const double*>(region0 + (size_t)ww * 1024);
                ssum += pw[lane]; qsum += pw[32 + lane]; sall += pw[64 + lane];
            }
            const double den = s_den[lane], num = s_dfc[lane] + sall, tau = s_tau[lane];
            if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) { ssum *= 2.0; qsum *= 2.0; }   // both triangles (taumode.rs:371-383)
            double e_raw = 0.0;
            if (den > 1e-12) { e_raw = num / den; if (!(e_raw > 0.0)) e_raw = 0.0; }
            double g = 0.0;
            // a NaN sum: the taumode form tests `<= 1e-12 -> 0` and lets it propagate through f64::clamp (taumode.rs:385-407), the energy form tests `> 1e-12` (energymaps.rs:1006)
        if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE ? !(ssum <= 1e-12) : (ssum > 1e-12)) { g = qsum / (ssum * ssum); g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g); }
            double lam;
            if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) {
                if (tau < 0.0) { lam = 0.0; g = 0.0; }
                else lam = tau * (e_raw / (e_raw + tau)) + (1.0 - tau) * g;       // taumode.rs:306-310
            } else lam = e_raw;
            if (i < a.n) {
                a.out_lambda[i] = lam;
                if (a.out_disp) a.out_disp[i] = g;
                run_mn = fmin(run_mn, lam); run_mx = fmax(run_mx, lam);
            }
        }
        __syncthreads();
    }
    if (w == 0 && a.minmax) {
        run_mn = warp_min_d(run_mn); run_mx = warp_max_d(run_mx);
        if (lane == 0) {
            if (run_mn == run_mn) atomicMin(&a.minmax[0], (unsigned long long)sort_key(run_mn));
            if (run_mx == run_mx) atomicMax(&a.minmax[1], (unsigned long long)sort_key(run_mx));
        }
    }
}

// ---- narrow tiles: F beyond the 32-item tile (768 < F <= 3111) ---------------------------------------------------
// A 32-item tile of F = 3072 would need 811 KB.  Here a tile holds ITEMS = 16 or 8 items (row stride ITEMS + 1 doubles:
// 1646 / 3111 features fit the 227 KB) and a warp walks SUB = 32 / ITEMS edges at a time: lane = (edge slot, item).  The
// record is no longer warp-uniform (SUB distinct 16-byte reads, broadcast inside a slot's lanes), x_r is fetched per edge
// (no run of equal rows to ride on) and the lanes of an item meet in a shuffle reduction at the end of the slice; the sums
// are reordered exactly as much as in the 32-item kernel (eight warp slices, then SUB slots), inside the 1e-9 of the
// contract.  Phase 1 streams each item row once (coalesced, eight loads in flight per lane) into the transposed tile and
// selects tau from the tile column with the strided quantised selection.
template <int VARIANT, int ITEMS>
__global__ void __launch_bounds__(256, 1) lambda_tile_narrow_kernel(LambdaTileArgs a) {
    constexpr int NW = 8, TS = ITEMS + 1, SUB = 32 / ITEMS, IPW = ITEMS / NW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t f = a.f;
    double* xs = reinterpret_cast<double*>(smem_raw);                         // [f][TS]
    double* s_den = xs + (((size_t)f * TS + 1) & ~(size_t)1);                 // [ITEMS]   (16-byte aligned)
    double* s_dfc = s_den + ITEMS;
    double* s_tau = s_dfc + ITEMS;                                            // < 0: zero vector
    unsigned char* region0 = reinterpret_cast<unsigned char*>(s_tau + ITEMS);
    unsigned char* region = region0 + (size_t)w * 1024;
    uint32_t* scratch = reinterpret_cast<uint32_t*>(region);                  // phase 1: 256-bin histogram
    uint4* ring = reinterpret_cast<uint4*>(region);                           // phase 2: [2][32] records
    double* part = reinterpret_cast<double*>(region);                         // end of phase 2: [3][ITEMS]
    const uint32_t ne = a.meta->ne;
    const bool has_defect = a.meta->any_defect != 0, nonpos = a.meta->any_nonpos != 0;
    const uint32_t e_lo = (uint32_t)((uint64_t)ne * w / NW), e_hi = (uint32_t)((uint64_t)ne * (w + 1) / NW);
    const uint32_t n_chunks = (e_hi - e_lo + 31) / 32;
    const int item = lane & (ITEMS - 1), sub = lane / ITEMS;
    const double* xs_item = xs + item;
    double run_mn = INFINITY, run_mx = 0.0;
    const uint64_t n_tiles = (a.n + ITEMS - 1) / ITEMS;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t i0 = tile * ITEMS;
        // ---- phase 1: warp w stages items w * IPW .. + IPW - 1
#pragma unroll 1
        for (int qi = 0; qi < IPW; ++qi) {
            const int it = w * IPW + qi;
            const uint64_t i = i0 + it;
            const bool live = i < a.n;
            const double* xr = a.x + (live ? i : 0) * f;
            double amax = 0.0, den = 0.0, dfc = 0.0;
            bool any_nan = false;
#pragma unroll 8
            for (uint32_t t = lane; t < f; t += 32) {
                const double v = live ? __ldcs(xr + t) : 0.0;
                xs[(size_t)t * TS + it] = v;
                amax = fmax(amax, fabs(v));
                any_nan = any_nan || v != v;
                den = fma(v, v, den);
                if (has_defect) dfc += (v * a.defect[t]) * v;
            }
            const bool zero = amax <= 1e-10 && !any_nan;
            den = warp_sum(den);
            if (has_defect) dfc = warp_sum(dfc);
            __syncwarp();
            double tau = 0.0;
            if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) {
                if (a.tau_in) tau = live ? a.tau_in[i] : -1.0;
                else if (__all_sync(FULL, zero)) tau = -1.0;                // taumode.rs:268-274
                else tau = warp_select_tau_fast<TS>(xs + it, f, a.tau_mode, a.tau_value, lane, scratch);
            }
            if (lane == 0) { s_den[it] = den; s_dfc[it] = dfc; s_tau[it] = tau; }
        }
        __syncthreads();
        // ---- phase 2: this warp's slice of the edge list, SUB edges per step, records one chunk ahead in the ring
        double s0 = 0.0, q0 = 0.0, s1 = 0.0, q1 = 0.0, sa = 0.0;
        if (n_chunks) {
            { const uint32_t e = e_lo + lane; if (e < e_hi) lt_cp_async16(&ring[lane], a.recs + e); asm volatile("cp.async.commit_group;" ::: "memory"); }
            { const uint32_t e = e_lo + 32 + lane; if (e < e_hi) lt_cp_async16(&ring[32 + lane], a.recs + e); asm volatile("cp.async.commit_group;" ::: "memory"); }
            for (uint32_t c = 0; c < n_chunks; ++c) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncwarp();
                const uint4* rr = ring + (c & 1) * 32;
                const uint32_t cnt = min(32u, e_hi - e_lo - c * 32);
                if (!nonpos) {
                    uint32_t j = sub;
                    for (; j + SUB < cnt; j += 2 * SUB) {
                        const uint4 ra = rr[j], rb = rr[j + SUB];
                        const double d0 = lt_ld(xs_item, ra.w) - lt_ld(xs_item, ra.z);
                        const double d1 = lt_ld(xs_item, rb.w) - lt_ld(xs_item, rb.z);
                        const double c0 = (fabs(__hiloint2double((int)ra.y, (int)ra.x)) * d0) * d0;
                        const double c1 = (fabs(__hiloint2double((int)rb.y, (int)rb.x)) * d1) * d1;
                        s0 += c0; q0 = fma(c0, c0, q0);
                        s1 += c1; q1 = fma(c1, c1, q1);
                    }
                    if (j < cnt) {
                        const uint4 ra = rr[j];
                        const double d0 = lt_ld(xs_item, ra.w) - lt_ld(xs_item, ra.z);
                        const double c0 = (fabs(__hiloint2double((int)ra.y, (int)ra.x)) * d0) * d0;
                        s0 += c0; q0 = fma(c0, c0, q0);
                    }
                } else {
                    for (uint32_t j = sub; j < cnt; j += SUB) {
                        const uint4 ra = rr[j];
                        const double wv = __hiloint2double((int)ra.y, (int)ra.x);
                        const double d0 = lt_ld(xs_item, ra.w & 0x7FFFFFFFu) - lt_ld(xs_item, ra.z);
                        const double c0 = (wv * d0) * d0;
                        sa += c0;
                        if (wv > 0.0) { s0 += c0; q0 = fma(c0, c0, q0); }
                    }
                }
                __syncwarp();
                { const uint32_t e = e_lo + (c + 2) * 32 + lane; if (e < e_hi) lt_cp_async16(&ring[(c & 1) * 32 + lane], a.recs + e); asm volatile("cp.async.commit_group;" ::: "memory"); }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        s0 += s1; q0 += q1;
        if (!nonpos) sa = s0;
#pragma unroll
        for (int o = ITEMS; o < 32; o <<= 1) {   // the SUB edge slots of an item
            s0 += __shfl_xor_sync(FULL, s0, o); q0 += __shfl_xor_sync(FULL, q0, o); sa += __shfl_xor_sync(FULL, sa, o);
        }
        if (sub == 0) { part[item] = s0; part[ITEMS + item] = q0; part[2 * ITEMS + item] = sa; }
        __syncthreads();
        // ---- phase 3: combine, blend, write
        if (w == 0 && lane < ITEMS) {
            const uint64_t i = i0 + lane;
            double ssum = 0.0, qsum = 0.0, sall = 0.0;
#pragma unroll
            for (int ww = 0; ww < NW; ++ww) {
                const double* pw = reinterpret_cast<const double*>(region0 + (size_t)ww * 1024);
                ssum += pw[lane]; qsum += pw[ITEMS + lane]; sall += pw[2 * ITEMS + lane];
            }
            const double den = s_den[lane], num = s_dfc[lane] + sall, tau = s_tau[lane];
            if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) { ssum *= 2.0; qsum *= 2.0; }   // both triangles (taumode.rs:371-383)
            double e_raw = 0.0;
            if (den > 1e-12) { e_raw = num / den; if (!(e_raw > 0.0)) e_raw = 0.0; }
            double g = 0.0;
            // NaN sums: see lambda_tile_kernel
            if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE ? !(ssum <= 1e-12) : (ssum > 1e-12)) { g = qsum / (ssum * ssum); g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g); }
            double lam;
            if (VARIANT == SFB_LAMBDA_LEGACY_TAUMODE) {
                if (tau < 0.0) { lam = 0.0; g = 0.0; }
                else lam = tau * (e_raw / (e_raw + tau)) + (1.0 - tau) * g;       // taumode.rs:306-310
            } else lam = e_raw;
            if (i < a.n) {
                a.out_lambda[i] = lam;
                if (a.out_disp) a.out_disp[i] = g;
                run_mn = fmin(run_mn, lam); run_mx = fmax(run_mx, lam);
            }
        }
        __syncthreads();
    }
    if (w == 0 && a.minmax) {
        run_mn = warp_min_d(run_mn); run_mx = warp_max_d(run_mx);
        if (lane == 0) {
            if (run_mn == run_mn) atomicMin(&a.minmax[0], (unsigned long long)sort_key(run_mn));
            if (run_mx == run_mx) atomicMax(&a.minmax[1], (unsigned long long)sort_key(run_mx));
        }
    }
}

// minmax keys -> {min, max(0, .)} as doubles (the input of the NCCL min / max exchange and of the normalisation)
__global__ void minmax_keys_kernel(const unsigned long long* __restrict__ keys, double* __restrict__ out) {
    out[0] = key_value(keys[0]); out[1] = key_value(keys[1]);
}
// (lambda - min) / max(max - min, 1e-9) with min / max read from device memory (core.rs:1341-1355)
__global__ void normalise_dev_kernel(double* __restrict__ v, uint64_t n, const double* __restrict__ mm) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double mn = mm[0];
    double rng = __dadd_rn(mm[1], -mn);
    if (!(rng > 1e-9)) rng = 1e-9;
    if (i < n) v[i] = __ddiv_rn(__dadd_rn(v[i], -mn), rng);
}

// CORE_F32SEM second pass: G_i = clamp(e_i / (sum e + 1e-12), 0, 1); lambda = R + G (f32)
__global__ void core_sum_energy_kernel(const float* __restrict__ e, uint64_t n, float* __restrict__ total) {
    // single block, ascending chunks: deterministic
    __shared__ float s[256];
    float acc = 0.f;
    for (uint64_t i = threadIdx.x; i < n; i += 256) acc += e[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) *total = s[0];
}
__global__ void core_finish_kernel(const float* __restrict__ e, const float* __restrict__ total, uint64_t n,
                                   double* __restrict__ lam, double* __restrict__ disp) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = e[i] / (*total + 1e-12f);
    g = g < 0.f ? 0.f : (g > 1.f ? 1.f : g);
    lam[i] = (double)((float)lam[i] + g);
    if (disp) disp[i] = (double)g;
}

// min / max(0, .) of lambda (core.rs:1345-1346), then (lambda - min) / max(max - min, 1e-9)
__global__ void minmax_kernel(const double* __restrict__ v, uint64_t n, double* __restrict__ out /* [2*gridDim.x] */) {
    __shared__ double smin[256], smax[256];
    double mn = INFINITY, mx = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double x = v[i];
        mn = fmin(mn, x); mx = fmax(mx, x);
    }
    smin[threadIdx.x] = mn; smax[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) { smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + o]); smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = smin[0]; out[2 * blockIdx.x + 1] = smax[0]; }
}

// diffusion: x_r' = x_r - eta * (L x)_r, `steps` times, row resident in shared memory (two slots)
__global__ void diffuse_kernel(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                               const double* __restrict__ data, uint32_t f, double* __restrict__ x, uint64_t n, double eta,
                               uint32_t steps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* cur = reinterpret_cast<double*>(smem_raw) + (size_t)w * 2 * f;
    double* nxt = cur + f;
    for (uint64_t i = (uint64_t)blockIdx.x * wpb + w; i < n; i += (uint64_t)gridDim.x * wpb) {
        double* xr = x + i * f;
        __syncwarp();
        for (uint32_t t = lane; t < f; t += 32) cur[t] = xr[t];
        __syncwarp();
        double* a = cur; double* b = nxt;
        for (uint32_t s = 0; s < steps; ++s) {
            for (uint32_t r = lane; r < f; r += 32) {
                double sum = 0.0;
                for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) sum = __dadd_rn(sum, __dmul_rn(data[e], a[indices[e]]));
                b[r] = __dadd_rn(a[r], -__dmul_rn(eta, sum));
            }
            __syncwarp();
            double* t = a; a = b; b = t;
        }
        for (uint32_t t = lane; t < f; t += 32) xr[t] = a[t];
    }
}

// diffusion, lane = item: a CTA stages a tile of 32 item rows transposed (as the lambda tile kernel does), keeps the whole CSR
// of L in shared memory as (value, byte offset of the column) records, and runs the steps between two tiles in shared memory:
// the row's record is one broadcast read for the warp, x[c] of 32 items one conflict-free read.  Every (L x)_r is the row's
// left fold with separately rounded multiply and add (graph.rs:486-493), so the result has the reference's bits.  One HBM
// read and one write of the rows, whatever the number of steps.
struct __align__(16) DiffRec { double v; uint32_t coff; uint32_t pad; };
__global__ void diffuse_pack_kernel(const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ indices, const double* __restrict__ data,
                                    uint32_t f, DiffRec* __restrict__ recs, uint32_t* __restrict__ ptr) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > f) return;
    ptr[r] = (uint32_t)indptr[r];
    if (r == f) return;
    for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) { DiffRec d; d.v = data[e]; d.coff = indices[e] * (LT_TS * 8); d.pad = 0; recs[e] = d; }
}
template <int NW>
__global__ void __launch_bounds__(NW * 32) diffuse_tile_kernel(const DiffRec* __restrict__ recs, const uint32_t* __restrict__ ptr, uint32_t nnz, uint32_t f,
                                                               double* __restrict__ x, uint64_t n, double eta, uint32_t steps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t tile = ((size_t)f * LT_TS + 1) & ~(size_t)1;
    double* t0 = reinterpret_cast<double*>(smem_raw);
    double* t1 = t0 + tile;
    DiffRec* s_rec = reinterpret_cast<DiffRec*>(t1 + tile);
    uint32_t* s_ptr = reinterpret_cast<uint32_t*>(s_rec + nnz);
    for (uint32_t e = threadIdx.x; e < nnz; e += NW * 32) s_rec[e] = recs[e];
    for (uint32_t r = threadIdx.x; r <= f; r += NW * 32) s_ptr[r] = ptr[r];
    __syncthreads();
    const uint64_t n_tiles = (n + 31) / 32;
    constexpr int IPW = 32 / NW;
    for (uint64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const uint64_t i0 = tl * 32;
#pragma unroll
        for (int qi = 0; qi < IPW; ++qi) {
            const int it = w * IPW + qi;
            const uint64_t i = i0 + it;
            for (uint32_t t = lane; t < f; t += 32) t0[(size_t)t * LT_TS + it] = i < n ? __ldcs(x + i * f + t) : 0.0;
        }
        __syncthreads();
        double* cur = t0; double* nxt = t1;
        for (uint32_t s = 0; s < steps; ++s) {
            const unsigned char* cur_lane = reinterpret_cast<const unsigned char*>(cur + lane);
            for (uint32_t r = w; r < f; r += NW) {
                double sum = 0.0;
                const uint32_t e1 = s_ptr[r + 1];
                for (uint32_t e = s_ptr[r]; e < e1; ++e) {
                    const DiffRec d = s_rec[e];
                    sum = __dadd_rn(sum, __dmul_rn(d.v, *reinterpret_cast<const double*>(cur_lane + d.coff)));
                }
                nxt[(size_t)r * LT_TS + lane] = __dadd_rn(cur[(size_t)r * LT_TS + lane], -__dmul_rn(eta, sum));
            }
            __syncthreads();
            double* tmp = cur; cur = nxt; nxt = tmp;
        }
#pragma unroll
        for (int qi = 0; qi < IPW; ++qi) {
            const int it = w * IPW + qi;
            const uint64_t i = i0 + it;
            if (i < n) for (uint32_t t = lane; t < f; t += 32) x[i * f + t] = cur[(size_t)t * LT_TS + it];
        }
        __syncthreads();
    }
}

// tau and the zero-vector test of every row (select_tau on the item, taumode.rs:174-175, and :268-274), for items that are
// projected before the Rayleigh quotient: both are taken from the UNPROJECTED vector.  One warp per row; -1 marks a zero vector.
__global__ void tau_rows_kernel(const double* __restrict__ x, uint64_t n, uint32_t f, int tau_mode, double tau_value, double* __restrict__ tau_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* xs = reinterpret_cast<double*>(smem_raw) + (size_t)w * f;
    uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)wpb * f) + w * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * wpb + w; i < n; i += (uint64_t)gridDim.x * wpb) {
        const double* xr = x + i * f;
        __syncwarp();
        bool zero = true;
        for (uint32_t t = lane; t < f; t += 32) { const double v = xr[t]; xs[t] = v; zero = zero && (fabs(v) <= 1e-10); }
        __syncwarp();
        zero = __all_sync(FULL, zero);
        const double tau = zero ? -1.0 : warp_select_tau_fast(xs, f, tau_mode, tau_value, lane, hist);
        if (lane == 0) tau_out[i] = tau;
    }
}

}  // namespace

namespace {

size_t lt_smem_bytes(uint32_t f, int nw) { return ((((size_t)f * LT_TS + 1) & ~(size_t)1) + 96) * sizeof(double) + (size_t)nw * 1024; }

template <int VARIANT, int E, int NW, bool FULL>
int32_t lt_launch_full(sfb_ctx* ctx, const LambdaTileArgs& a) {
    const size_t smem = lt_smem_bytes(a.f, NW);
    auto kern = lambda_tile_kernel<VARIANT, E, NW, FULL>;
    SFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SFB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const uint64_t tiles = (a.n + 31) / 32, cap = (uint64_t)ctx->sm_count * per_sm;
    kern<<<(unsigned)(tiles < cap ? tiles : cap), NW * 32, smem, ctx->stream>>>(a);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}
template <int VARIANT, int E, int NW>
int32_t lt_launch(sfb_ctx* ctx, const LambdaTileArgs& a) {
    return a.f == 32u * E ? lt_launch_full<VARIANT, E, NW, true>(ctx, a) : lt_launch_full<VARIANT, E, NW, false>(ctx, a);
}
constexpr uint32_t LT_MAX_F = 768, LTN_MAX_F16 = 1646, LTN_MAX_F8 = 3111;   // 32-, 16-, 8-item tiles in 227 KB
uint32_t lt_items(uint32_t f) { return f <= LT_MAX_F ? 32u : (f <= LTN_MAX_F16 ? 16u : 8u); }
size_t ltn_smem_bytes(uint32_t f, uint32_t items) { return ((((size_t)f * (items + 1) + 1) & ~(size_t)1) + 3 * items) * sizeof(double) + 8 * 1024; }

template <int VARIANT, int ITEMS>
int32_t ltn_launch(sfb_ctx* ctx, const LambdaTileArgs& a) {
    const size_t smem = ltn_smem_bytes(a.f, ITEMS);
    auto kern = lambda_tile_narrow_kernel<VARIANT, ITEMS>;
    SFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t tiles = (a.n + ITEMS - 1) / ITEMS, cap = (uint64_t)ctx->sm_count;
    kern<<<(unsigned)(tiles < cap ? tiles : cap), 256, smem, ctx->stream>>>(a);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

template <int VARIANT>
int32_t lt_dispatch(sfb_ctx* ctx, const LambdaTileArgs& a) {
    if (a.f > LT_MAX_F) return a.f <= LTN_MAX_F16 ? ltn_launch<VARIANT, 16>(ctx, a) : ltn_launch<VARIANT, 8>(ctx, a);
    const uint32_t e = (a.f + 31) / 32;
    if (e <= 4) return lt_launch<VARIANT, 4, 8>(ctx, a);
    if (e <= 8) return lt_launch<VARIANT, 8, 8>(ctx, a);
    if (e <= 12) return lt_launch<VARIANT, 12, 8>(ctx, a);
    if (e <= 16) return lt_launch<VARIANT, 16, 16>(ctx, a);
    return lt_launch<VARIANT, 24, 16>(ctx, a);
}

__global__ void minmax_final_kernel(const double* __restrict__ part, int nb, double* __restrict__ out) {
    double mn = INFINITY, mx = 0.0;
    for (int i = threadIdx.x; i < nb; i += 32) { mn = fmin(mn, part[2 * i]); mx = fmax(mx, part[2 * i + 1]); }
    mn = warp_min_d(mn); mx = warp_max_d(mx);
    if (threadIdx.x == 0) { out[0] = mn; out[1] = mx; }
}
__global__ void init_minmax_keys_kernel(unsigned long long* keys) { keys[0] = sort_key(INFINITY); keys[1] = sort_key(0.0); }

}  // namespace

// The packed upper triangle of a symmetric L for the tile kernel, built once per handle (sfb_csr_free releases it).
static int32_t lt_pack(sfb_ctx* ctx, const sfb_csr* L) {
    if (L->lt_recs) return SFB_OK;
    const uint32_t f = (uint32_t)L->rows;
    void *recs = nullptr, *defect = nullptr, *meta = nullptr;
    if (sfb_dev_alloc(ctx, &recs, sizeof(EdgeRec) * (L->nnz / 2 + 1)) != cudaSuccess || sfb_dev_alloc(ctx, &defect, sizeof(double) * f) != cudaSuccess ||
        sfb_dev_alloc(ctx, &meta, sizeof(PackMeta)) != cudaSuccess) {
        sfb_dev_free(ctx, recs); sfb_dev_free(ctx, defect); sfb_dev_free(ctx, meta);
        return sfb_fail(ctx, SFB_ENOMEM, "packed Laplacian");
    }
    lambda_pack_kernel<<<1, 1024, 0, ctx->stream>>>(L->indptr, L->indices, L->data, f, (lt_items(f) == 32 ? (uint32_t)LT_TS : lt_items(f) + 1) * 8u,
                                                    (EdgeRec*)recs, (double*)defect, (PackMeta*)meta);
    ctx->times.kernel_launches++;
    if (cudaGetLastError() != cudaSuccess) { sfb_dev_free(ctx, recs); sfb_dev_free(ctx, defect); sfb_dev_free(ctx, meta); return sfb_fail(ctx, SFB_ECUDA, "lambda_pack_kernel launch failed"); }
    L->lt_recs = recs; L->lt_defect = defect; L->lt_meta = meta;
    return SFB_OK;
}

// lambdas of rows [0, n) of a device matrix into a device array; shared by the single- and multi-GPU paths.
// d_mm (2 doubles on the device, may be null): receives {min lambda, max(0, max lambda)} of these rows (core.rs:1345-1346).
int32_t sfb_lambda_device(sfb_ctx* ctx, const sfb_csr* L, const double* x_dev, uint64_t n, uint32_t f,
                          const sfb_lambda_params* prm, double* d_lambda, double* d_disp, const double* tau_in, double* d_mm) {
    if (L->cols && L->cols != L->rows) return sfb_fail(ctx, SFB_EINVAL, "a row shard of a Laplacian is not a square operator");
    if (L->rows != f) return sfb_fail(ctx, SFB_EINVAL, "Matrix rows %llu must match vector length %u", (unsigned long long)L->rows, f);  // taumode.rs:330-337
    if (prm->variant < 0 || prm->variant > 2) return sfb_fail(ctx, SFB_EINVAL, "unknown lambda variant %d", prm->variant);
    if (prm->tau_mode < 0 || prm->tau_mode > 3) return sfb_fail(ctx, SFB_EINVAL, "unknown tau mode %d", prm->tau_mode);
    auto minmax_generic = [&]() -> int32_t {
        if (!d_mm) return SFB_OK;
        const int nb = 128;
        DevBuf part;
        SFB_CUDA(ctx, part.alloc(sizeof(double) * 2 * nb));
        minmax_kernel<<<nb, 256, 0, ctx->stream>>>(d_lambda, n, part.as<double>());
        SFB_LAUNCH_CHECK(ctx);
        minmax_final_kernel<<<1, 32, 0, ctx->stream>>>(part.as<double>(), nb, d_mm);
        SFB_LAUNCH_CHECK(ctx);
        return SFB_OK;   // part goes back to the context's cache; stream order keeps it valid
    };
    const bool sym_variant = prm->variant != SFB_LAMBDA_CORE_F32SEM;
    if (sym_variant && f <= 65535 && L->symmetric < 0 && !getenv("SFB_LAMBDA_ROWWISE")) {
        DevBuf bad;
        SFB_CUDA(ctx, bad.alloc(sizeof(int)));
        SFB_CUDA(ctx, cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
        csr_symmetric_kernel<<<div_up(L->rows, 128), 128, 0, ctx->stream>>>(L->indptr, L->indices, L->data, L->rows, bad.as<int>());
        SFB_LAUNCH_CHECK(ctx);
        int hb = 0;
        SFB_CUDA(ctx, cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        L->symmetric = hb ? 0 : 1;
    }
    // tile kernels (lane = item): LEGACY_TAUMODE / ENERGY_NODE, symmetric L, F <= 768 (32-item tiles) or <= 3111 (16 / 8 items)
    if (sym_variant && L->symmetric == 1 && f <= (getenv("SFB_LAMBDA_NO_NARROW") ? LT_MAX_F : LTN_MAX_F8) && L->nnz / 2 < 0x7FFFFFFFull / LT_TS && !getenv("SFB_LAMBDA_ROWWISE") && !getenv("SFB_LAMBDA_SYM")) {
        SFB_TRY(lt_pack(ctx, L));
        DevBuf keys;
        if (d_mm) {
            SFB_CUDA(ctx, keys.alloc(2 * sizeof(unsigned long long)));
            init_minmax_keys_kernel<<<1, 1, 0, ctx->stream>>>(keys.as<unsigned long long>());
            SFB_LAUNCH_CHECK(ctx);
        }
        LambdaTileArgs ta{(const EdgeRec*)L->lt_recs, (const double*)L->lt_defect, (const PackMeta*)L->lt_meta, f, x_dev, n, prm->tau_mode, prm->tau_value,
                          d_lambda, d_disp, tau_in, d_mm ? keys.as<unsigned long long>() : nullptr};
        {
            StageTimer tk(ctx, &ctx->times.ms_lambda_kernel);
            if (prm->variant == SFB_LAMBDA_LEGACY_TAUMODE) SFB_TRY(lt_dispatch<0>(ctx, ta));
            else SFB_TRY(lt_dispatch<1>(ctx, ta));
        }
        if (d_mm) {
            minmax_keys_kernel<<<1, 1, 0, ctx->stream>>>(keys.as<unsigned long long>(), d_mm);
            SFB_LAUNCH_CHECK(ctx);
        }
        return SFB_OK;
    }
    // symmetric fast path (LEGACY_TAUMODE / ENERGY_NODE): L symmetric bit for bit, packed upper triangle in shared memory
    if (sym_variant && f <= 65535 && !getenv("SFB_LAMBDA_ROWWISE")) {
        if (L->symmetric == 1) {
            DevBuf cnt, offs, rc, val, diag;
            SFB_CUDA(ctx, cnt.alloc(sizeof(uint32_t) * f));
            SFB_CUDA(ctx, offs.alloc(sizeof(uint64_t) * ((size_t)f + 1)));
            upper_count_kernel<<<div_up(f, 128), 128, 0, ctx->stream>>>(L->indptr, L->indices, f, cnt.as<uint32_t>());
            SFB_LAUNCH_CHECK(ctx);
            SFB_TRY(sfb_scan_exclusive_u64(ctx, cnt.as<uint32_t>(), f, offs.as<uint64_t>()));
            uint64_t ne64 = 0;
            SFB_CUDA(ctx, cudaMemcpyAsync(&ne64, offs.as<uint64_t>() + f, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
            SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const int wpb_s = 8;
            const size_t smem_s = (size_t)ne64 * (sizeof(double) + sizeof(uint32_t)) + (size_t)f * sizeof(double) * (1 + wpb_s) + (size_t)wpb_s * 256 * sizeof(uint32_t) + 16;
            if (smem_s <= 112 * 1024) {   // two CTAs per SM
                const uint32_t ne = (uint32_t)ne64;
                SFB_CUDA(ctx, rc.alloc(sizeof(uint32_t) * (ne ? ne : 1)));
                SFB_CUDA(ctx, val.alloc(sizeof(double) * (ne ? ne : 1)));
                SFB_CUDA(ctx, diag.alloc(sizeof(double) * f));
                upper_pack_kernel<<<div_up(f, 128), 128, 0, ctx->stream>>>(L->indptr, L->indices, L->data, f, offs.as<uint64_t>(), rc.as<uint32_t>(),
                                                                         val.as<double>(), diag.as<double>());
                SFB_LAUNCH_CHECK(ctx);
                LambdaSymArgs sa{rc.as<uint32_t>(), val.as<double>(), diag.as<double>(), ne, f, x_dev, n, prm->tau_mode, prm->tau_value, d_lambda, d_disp, tau_in};
                const int per_sm = (int)((ctx->smem_optin ? ctx->smem_optin : 232448) / (smem_s + 1024));
                uint64_t want = (n + wpb_s - 1) / wpb_s;
                const uint64_t cap_grid = (uint64_t)ctx->sm_count * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
                const unsigned grid = (unsigned)(want < cap_grid ? want : cap_grid);
                StageTimer tk(ctx, &ctx->times.ms_lambda_kernel);
                if (prm->variant == SFB_LAMBDA_LEGACY_TAUMODE) {
                    SFB_CUDA(ctx, cudaFuncSetAttribute(lambda_sym_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
                    lambda_sym_kernel<0><<<grid, wpb_s * 32, smem_s, ctx->stream>>>(sa);
                } else {
                    SFB_CUDA(ctx, cudaFuncSetAttribute(lambda_sym_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
                    lambda_sym_kernel<1><<<grid, wpb_s * 32, smem_s, ctx->stream>>>(sa);
                }
                SFB_LAUNCH_CHECK(ctx);
                tk.stop();   // synchronises: the packed arrays may be released on return
                return minmax_generic();
            }
        }
    }
    size_t per_warp = (size_t)f * sizeof(double) + 256 * sizeof(uint32_t);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
    if (per_warp * wpb > ctx->smem_optin) return sfb_fail(ctx, SFB_EUNSUPPORTED, "feature count %u too large for the shared-memory row staging", f);
    size_t smem = per_warp * wpb;
    DevBuf e32, tot;
    LambdaArgs a{L->indptr, L->indices, L->data, f, x_dev, n, prm->variant, prm->tau_mode, prm->tau_value, d_lambda, d_disp, nullptr, tau_in};
    // enough resident warps to cover HBM latency; grid = multiple of the SM count
    int blocks_per_sm = (int)((ctx->smem_optin) / (smem + 1024));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (blocks_per_sm > 8) blocks_per_sm = 8;
    uint64_t want = (n + wpb - 1) / wpb;
    unsigned grid = (unsigned)(want < (uint64_t)ctx->sm_count * blocks_per_sm ? want : (uint64_t)ctx->sm_count * blocks_per_sm);
    StageTimer tk(ctx, &ctx->times.ms_lambda_kernel);
    if (prm->variant == SFB_LAMBDA_LEGACY_TAUMODE) {
        SFB_CUDA(ctx, cudaFuncSetAttribute(lambda_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lambda_kernel<0><<<grid, wpb * 32, smem, ctx->stream>>>(a);
    } else if (prm->variant == SFB_LAMBDA_ENERGY_NODE) {
        SFB_CUDA(ctx, cudaFuncSetAttribute(lambda_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lambda_kernel<1><<<grid, wpb * 32, smem, ctx->stream>>>(a);
    } else {
        SFB_CUDA(ctx, e32.alloc(sizeof(float) * n));
        SFB_CUDA(ctx, tot.alloc(sizeof(float)));
        a.out_energy_f32 = e32.as<float>();
        SFB_CUDA(ctx, cudaFuncSetAttribute(lambda_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lambda_kernel<2><<<grid, wpb * 32, smem, ctx->stream>>>(a);
        SFB_LAUNCH_CHECK(ctx);
        core_sum_energy_kernel<<<1, 256, 0, ctx->stream>>>(e32.as<float>(), n, tot.as<float>());
        SFB_LAUNCH_CHECK(ctx);
        // sharded items (sfb_lambda_allgather): the dispersion is normalised by the energy of ALL items
        // (dirichlet_dispersion_gpu "normalize by global total", spectral/mod.rs:140-145): sum the per-rank totals
        if (ctx->world > 1 && ctx->lambda_sharded) SFB_TRY(sfb_comm_allreduce_sum_f32(ctx, tot.as<float>(), 1));
        core_finish_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(e32.as<float>(), tot.as<float>(), n, d_lambda, d_disp);
    }
    SFB_LAUNCH_CHECK(ctx);
    tk.stop();   // synchronises: e32 / tot may be released on return
    return minmax_generic();
}

// (lambda - min) / max(max - min, 1e-9) with {min, max} in device memory (core.rs:1341-1355)
int32_t sfb_normalise_device(sfb_ctx* ctx, double* d_lambda, uint64_t n, const double* d_mm) {
    normalise_dev_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(d_lambda, n, d_mm);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

// tau and the zero-vector test of every row of x_tau (the unprojected items) into a device array
int32_t sfb_tau_rows_device(sfb_ctx* ctx, const sfb_mat* x_tau, uint64_t rows, const sfb_lambda_params* prm, double* d_tau) {
    const uint32_t ft = x_tau->cols;
    const size_t per_warp = (size_t)ft * sizeof(double) + 256 * sizeof(uint32_t);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 100 * 1024) wpb >>= 1;
    if (per_warp * wpb > ctx->smem_optin) return sfb_fail(ctx, SFB_EUNSUPPORTED, "feature count %u too large for the shared-memory row staging", ft);
    SFB_CUDA(ctx, cudaFuncSetAttribute(tau_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)));
    const uint64_t want = (rows + wpb - 1) / wpb;
    const unsigned grid = (unsigned)(want < (uint64_t)ctx->sm_count * 8 ? want : (uint64_t)ctx->sm_count * 8);
    tau_rows_kernel<<<grid, wpb * 32, per_warp * wpb, ctx->stream>>>(x_tau->d, rows, ft, prm->tau_mode, prm->tau_value, d_tau);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

// x_tau != null: tau and the zero-vector test come from its rows (the unprojected items), energy and dispersion from x
static int32_t lambda_to_host(sfb_ctx* ctx, const sfb_csr* L, const sfb_mat* x, const sfb_mat* x_tau, const sfb_lambda_params* prm,
                              double* out_lambda, double* out_disp, double* stats) {
    double h_mm[2] = {0.0, 0.0};
    {
        StageTimer t(ctx, &ctx->times.ms_lambda);
        DevBuf lam, disp, tau, mm;
        SFB_CUDA(ctx, lam.alloc(sizeof(double) * x->rows));
        SFB_CUDA(ctx, mm.alloc(sizeof(double) * 2));
        if (out_disp) SFB_CUDA(ctx, disp.alloc(sizeof(double) * x->rows));
        if (x_tau) {
            SFB_CUDA(ctx, tau.alloc(sizeof(double) * x->rows));
            SFB_TRY(sfb_tau_rows_device(ctx, x_tau, x->rows, prm, tau.as<double>()));
        }
        const bool want_mm = prm->normalise_minmax || stats;
        SFB_TRY(sfb_lambda_device(ctx, L, x->d, x->rows, x->cols, prm, lam.as<double>(), out_disp ? disp.as<double>() : nullptr,
                                  x_tau ? tau.as<double>() : nullptr, want_mm ? mm.as<double>() : nullptr));
        if (prm->normalise_minmax) SFB_TRY(sfb_normalise_device(ctx, lam.as<double>(), x->rows, mm.as<double>()));
        t.stop();
        StageTimer t2(ctx, &ctx->times.ms_d2h);
        SFB_CUDA(ctx, cudaMemcpyAsync(out_lambda, lam.p, sizeof(double) * x->rows, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_disp) SFB_CUDA(ctx, cudaMemcpyAsync(out_disp, disp.p, sizeof(double) * x->rows, cudaMemcpyDeviceToHost, ctx->stream));
        if (want_mm) SFB_CUDA(ctx, cudaMemcpyAsync(h_mm, mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, ctx->stream));
        t2.stop();   // synchronises: the results are on the host, the scratch may go back to the cache
    }
    if (stats) { stats[0] = h_mm[0]; stats[1] = h_mm[1]; stats[2] = (h_mm[1] - h_mm[0]) > 1e-9 ? (h_mm[1] - h_mm[0]) : 1e-9; }
    return SFB_OK;
}

extern "C" int32_t sfb_lambda(sfb_ctx* ctx, const sfb_csr* L, const sfb_mat* x, const sfb_lambda_params* prm,
                              double* out_lambda, double* out_disp, double* stats) {
    if (!ctx || !L || !x || !prm || !out_lambda) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    return lambda_to_host(ctx, L, x, nullptr, prm, out_lambda, out_disp, stats);
}

extern "C" int32_t sfb_lambda_projected(sfb_ctx* ctx, const sfb_csr* L, const sfb_mat* x_original, const sfb_mat* x_projected,
                                        const sfb_lambda_params* prm, double* out_lambda, double* out_disp, double* stats) {
    if (!ctx || !L || !x_original || !x_projected || !prm || !out_lambda) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (prm->variant != SFB_LAMBDA_LEGACY_TAUMODE) return sfb_fail(ctx, SFB_EINVAL, "only the taumode lambda projects its items (taumode.rs:261-318)");
    if (x_original->rows != x_projected->rows) return sfb_fail(ctx, SFB_EINVAL, "original and projected items differ in count");
    // taumode.rs:287-297: the projected length must be the Laplacian's ("item seems neither projected nor unprojected" otherwise)
    if (L->rows != x_projected->cols) return sfb_fail(ctx, SFB_EINVAL, "projected items have %u dimensions, the Laplacian %llu rows", x_projected->cols, (unsigned long long)L->rows);
    return lambda_to_host(ctx, L, x_projected, x_original, prm, out_lambda, out_disp, stats);
}

// ---- the successor's Stage D seams ------------------------------------------------------------------------------
// compute_tau_mode_gpu (surfface-core/src/spectral/bridge.rs:27-32): f32 items in, f64 lambdas out, CORE_F32SEM.
extern "C" int32_t sfb_compute_tau_mode_lambdas(sfb_ctx* ctx, const sfb_csr* L, const float* data, uint64_t n_items, uint32_t n_features,
                                                double* out_lambdas) {
    if (!ctx || !L || !data || !out_lambdas) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    sfb_mat* x = nullptr;
    SFB_TRY(sfb_mat_from_host_f32(ctx, data, n_items, n_features, &x));
    sfb_lambda_params lp{SFB_LAMBDA_CORE_F32SEM, SFB_TAU_MEDIAN, 0.0, 0};
    int32_t st = lambda_to_host(ctx, L, x, nullptr, &lp, out_lambdas, nullptr, nullptr);
    sfb_mat_free(x);
    return st;
}

// compute_tau (surfface-core/src/taumode.rs:37-65): one f32 tau from the lambda distribution.
namespace {
// pass over the finite entries whose key matches `prefix` under `mask`: 256-bin histogram of the next 8 bits
__global__ void __launch_bounds__(256) tau_hist_kernel(const float* __restrict__ v, uint64_t n, uint32_t prefix, uint32_t mask, int shift,
                                                       uint32_t* __restrict__ hist /* 256 + 1 (finite count) */) {
    __shared__ uint32_t sh[257];
    for (int b = threadIdx.x; b < 257; b += 256) sh[b] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = v[i];
        if (!isfinite(x)) continue;
        const uint32_t key = f32_key(x);
        if ((key & mask) == prefix) { atomicAdd(&sh[(key >> shift) & 0xFFu], 1u); if (mask == 0u) atomicAdd(&sh[256], 1u); }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 257; b += 256) if (sh[b]) atomicAdd(&hist[b], sh[b]);
}
// Rust's iter().sum::<f32>() over the finite entries: one strictly sequential f32 chain (thread 0), tiles staged by warps 1-3
constexpr int TM_TILE = 4096;
__global__ void __launch_bounds__(128) tau_mean_kernel(const float* __restrict__ v, uint64_t n, float* __restrict__ out /* sum, count */) {
    __shared__ float buf[2][TM_TILE];
    const int tid = threadIdx.x;
    const uint64_t ntiles = (n + TM_TILE - 1) / TM_TILE;
    for (int e = tid; e < TM_TILE; e += 128) buf[0][e] = (uint64_t)e < n ? v[e] : INFINITY;   // non-finite entries are skipped by the fold
    __syncthreads();
    float sum = 0.0f; uint32_t cnt = 0;
    for (uint64_t t = 0; t < ntiles; t++) {
        if (tid >= 32) {
            const uint64_t base = (t + 1) * TM_TILE;
            if (base < n) for (int e = tid - 32; e < TM_TILE; e += 96) buf[(t + 1) & 1][e] = base + e < n ? v[base + e] : INFINITY;
        } else if (tid == 0) {
            const float* b = buf[t & 1];
#pragma unroll 16
            for (int e = 0; e < TM_TILE; e++) { const float x = b[e]; if (isfinite(x)) { sum = __fadd_rn(sum, x); ++cnt; } }
        }
        __syncthreads();
    }
    if (tid == 0) { out[0] = sum; out[1] = __uint_as_float(cnt); }
}
}  // namespace

extern "C" int32_t sfb_compute_tau(sfb_ctx* ctx, const float* lambdas, uint64_t n, int32_t tau_mode, float tau_value, float* out_tau) {
    if (!ctx || !out_tau || (n && !lambdas)) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (tau_mode < 0 || tau_mode > 3) return sfb_fail(ctx, SFB_EINVAL, "unknown tau mode %d", tau_mode);
    const float FLOOR = 1e-9f;
    *out_tau = FLOOR;
    if (n == 0) return SFB_OK;
    DevBuf dv, hist;
    SFB_CUDA(ctx, dv.alloc(sizeof(float) * n));
    SFB_CUDA(ctx, hist.alloc(sizeof(uint32_t) * 257));
    SFB_CUDA(ctx, cudaMemcpyAsync(dv.p, lambdas, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = (unsigned)(n / 256 + 1 < (uint64_t)ctx->sm_count * 8 ? n / 256 + 1 : (uint64_t)ctx->sm_count * 8);
    uint32_t h[257];
    auto pass = [&](uint32_t prefix, uint32_t mask, int shift) -> int32_t {
        SFB_CUDA(ctx, cudaMemsetAsync(hist.p, 0, sizeof(uint32_t) * 257, ctx->stream));
        tau_hist_kernel<<<grid, 256, 0, ctx->stream>>>(dv.as<float>(), n, prefix, mask, shift, hist.as<uint32_t>());
        SFB_LAUNCH_CHECK(ctx);
        SFB_CUDA(ctx, cudaMemcpyAsync(h, hist.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SFB_OK;
    };
    SFB_TRY(pass(0u, 0u, 24));
    const uint64_t c = h[256];                      // finite entries (n < 2^32: lambdas of u32-indexed items)
    if (c == 0) return SFB_OK;                      // finite.is_empty() -> TAU_FLOOR
    float r;
    if (tau_mode == SFB_TAU_FIXED) r = isfinite(tau_value) ? tau_value : FLOOR;
    else if (tau_mode == SFB_TAU_MEAN) {
        DevBuf o;
        SFB_CUDA(ctx, o.alloc(2 * sizeof(float)));
        tau_mean_kernel<<<1, 128, 0, ctx->stream>>>(dv.as<float>(), n, o.as<float>());
        SFB_LAUNCH_CHECK(ctx);
        float ho[2];
        SFB_CUDA(ctx, cudaMemcpyAsync(ho, o.p, sizeof(ho), cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        r = ho[0] / (float)c;
    } else {
        uint64_t rank;
        if (tau_mode == SFB_TAU_MEDIAN) rank = c / 2;
        else {
            float pp = tau_value;
            if (pp < 0.0f) pp = 0.0f; else if (pp > 1.0f) pp = 1.0f;
            const float fi = roundf(((float)c - 1.0f) * pp);
            rank = fi != fi ? 0 : (uint64_t)fi;
            if (rank >= c) rank = c - 1;
        }
        uint32_t prefix = 0, mask = 0;
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (shift != 24) SFB_TRY(pass(prefix, mask, shift));
            uint64_t run = 0; uint32_t b = 0;
            for (; b < 256; ++b) { if (rank < run + h[b]) break; run += h[b]; }
            if (b == 256) return sfb_fail(ctx, SFB_ECUDA, "radix select lost its rank");
            rank -= run;
            prefix |= b << shift; mask |= 0xFFu << shift;
        }
        const uint32_t u = (prefix >> 31) ? (prefix & 0x7FFFFFFFu) : ~prefix;
        memcpy(&r, &u, 4);
    }
    *out_tau = r > FLOOR ? r : FLOOR;   // f32::max(TAU_FLOOR)
    return SFB_OK;
}

// ---- energy pipeline: item -> sub-centroid mapping (src_legacy/energymaps.rs:1246-1342) ----------------------
// One warp per item.  The |delta lambda| scan over the S sub-centroids is a lexicographic (distance, index) warp
// minimum, i.e. the reference's "first strictly smaller"; ties within epsilon are rare, and only then are cosines
// computed: lane = candidate, each lane running the reference's left folds (dot, |centroid|^2) so the strict
// comparison between candidates sees the reference's bits.
namespace {
__global__ void map_items_kernel(const double* __restrict__ x, uint64_t n, uint32_t f, const double* __restrict__ item_lam,
                                 const double* __restrict__ subc, uint32_t s, const double* __restrict__ sub_lam, double eps,
                                 uint32_t* __restrict__ out_idx, double* __restrict__ out_lam, double* __restrict__ out_norm) {
    const int lane = threadIdx.x & 31;
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const double* xi = x + i * f;
    const double li = item_lam[i];
    double bd = INFINITY; uint32_t bi = 0xFFFFFFFFu;
    for (uint32_t c = lane; c < s; c += 32) { double d = fabs(li - sub_lam[c]); if (d < bd) { bd = d; bi = c; } }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        double od = __shfl_xor_sync(FULL, bd, o); uint32_t oi = __shfl_xor_sync(FULL, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    // |item| = sqrt of the left-fold sum of squares: lane 0 folds (F steps), the others count candidates meanwhile
    double nsq = 0.0;
    if (lane == 0) for (uint32_t t = 0; t < f; ++t) nsq = __dadd_rn(nsq, __dmul_rn(xi[t], xi[t]));
    nsq = __shfl_sync(FULL, nsq, 0);
    const double norm = __dsqrt_rn(nsq);
    uint32_t ncand = 0;
    for (uint32_t c = lane; c < s; c += 32) ncand += fabs(fabs(li - sub_lam[c]) - bd) < eps ? 1u : 0u;
    ncand = warp_sum_u(ncand);
    uint32_t best = bi;
    if (ncand > 1) {
        double best_cos = -INFINITY; uint32_t best_sc = 0xFFFFFFFFu;
        for (uint32_t c0 = 0; c0 < s; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool cand = c < s && fabs(fabs(li - sub_lam[c]) - bd) < eps;
            if (!__any_sync(FULL, cand)) continue;
            double cosv = -INFINITY;
            if (cand) {
                const double* y = subc + (uint64_t)c * f;
                double dot = 0.0, cn = 0.0;
                for (uint32_t t = 0; t < f; ++t) { dot = __dadd_rn(dot, __dmul_rn(xi[t], y[t])); cn = __dadd_rn(cn, __dmul_rn(y[t], y[t])); }
                const double cnorm = __dsqrt_rn(cn);
                cosv = (norm > 0.0 && cnorm > 0.0) ? __ddiv_rn(dot, __dmul_rn(norm, cnorm)) : 0.0;
            }
            // strictly-greater in ascending candidate order == lexicographic max of (cosine, -index)
            double mc = cosv; uint32_t mi = cand ? c : 0xFFFFFFFFu;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                double oc = __shfl_xor_sync(FULL, mc, o); uint32_t oi = __shfl_xor_sync(FULL, mi, o);
                if (oc > mc || (oc == mc && oi < mi)) { mc = oc; mi = oi; }
            }
            if (mi != 0xFFFFFFFFu && mc > best_cos) { best_cos = mc; best_sc = mi; }
        }
        if (best_sc != 0xFFFFFFFFu) best = best_sc;
    }
    if (lane == 0) { out_idx[i] = best; out_lam[i] = sub_lam[best]; out_norm[i] = norm; }
}
}  // namespace

extern "C" int32_t sfb_map_items_to_subcentroids(sfb_ctx* ctx, const sfb_mat* items, const double* item_lambdas, const sfb_mat* sub_centroids,
                                                 const double* sub_lambdas, double epsilon, uint32_t* out_idx, double* out_lambda,
                                                 double* out_norm) {
    if (!ctx || !items || !item_lambdas || !sub_centroids || !sub_lambdas || !out_idx) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (items->cols != sub_centroids->cols) return sfb_fail(ctx, SFB_EINVAL, "items have %u features, sub-centroids %u", items->cols, sub_centroids->cols);
    if (sub_centroids->rows == 0 || sub_centroids->rows > 0xFFFFFFFEull) return sfb_fail(ctx, SFB_EINVAL, "bad sub-centroid count");
    const uint64_t n = items->rows; const uint32_t s = (uint32_t)sub_centroids->rows;
    StageTimer t(ctx, &ctx->times.ms_lambda);
    DevBuf il, sl, oi, ol, on;
    SFB_CUDA(ctx, il.alloc(sizeof(double) * n)); SFB_CUDA(ctx, sl.alloc(sizeof(double) * s));
    SFB_CUDA(ctx, oi.alloc(sizeof(uint32_t) * n)); SFB_CUDA(ctx, ol.alloc(sizeof(double) * n)); SFB_CUDA(ctx, on.alloc(sizeof(double) * n));
    SFB_CUDA(ctx, cudaMemcpyAsync(il.p, item_lambdas, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    SFB_CUDA(ctx, cudaMemcpyAsync(sl.p, sub_lambdas, sizeof(double) * s, cudaMemcpyHostToDevice, ctx->stream));
    map_items_kernel<<<div_up(n * 32, 256), 256, 0, ctx->stream>>>(items->d, n, items->cols, il.as<double>(), sub_centroids->d, s, sl.as<double>(), epsilon,
                                                                   oi.as<uint32_t>(), ol.as<double>(), on.as<double>());
    SFB_LAUNCH_CHECK(ctx);
    SFB_CUDA(ctx, cudaMemcpyAsync(out_idx, oi.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_lambda) SFB_CUDA(ctx, cudaMemcpyAsync(out_lambda, ol.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_norm) SFB_CUDA(ctx, cudaMemcpyAsync(out_norm, on.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}

extern "C" int32_t sfb_diffuse(sfb_ctx* ctx, const sfb_csr* L, sfb_mat* x, double eta, uint32_t steps) {
    if (!ctx || !L || !x) return sfb_fail(ctx, SFB_EINVAL, "null argument");
    if (L->cols && L->cols != L->rows) return sfb_fail(ctx, SFB_EINVAL, "a row shard of a Laplacian is not a square operator");
    if (L->rows != x->cols) return sfb_fail(ctx, SFB_EINVAL, "Laplacian rows %llu must match feature count %u", (unsigned long long)L->rows, x->cols);  // energymaps.rs:507-512
    {   // tile kernel when two transposed tiles and the CSR records fit shared memory
        const uint32_t f = x->cols;
        const size_t tile = (((size_t)f * LT_TS + 1) & ~(size_t)1) * sizeof(double);
        const size_t smem_t = 2 * tile + (size_t)L->nnz * sizeof(DiffRec) + ((size_t)f + 1) * sizeof(uint32_t) + 16;
        if (smem_t <= (ctx->smem_optin ? ctx->smem_optin : 232448) && L->nnz < 0xFFFFFFFFull && f <= 0x7FFFFFFFu / (LT_TS * 8) && !getenv("SFB_DIFFUSE_ROWWISE")) {
            DevBuf recs, ptr;
            SFB_CUDA(ctx, recs.alloc(sizeof(DiffRec) * (L->nnz ? L->nnz : 1)));
            SFB_CUDA(ctx, ptr.alloc(sizeof(uint32_t) * ((size_t)f + 1)));
            StageTimer t(ctx, &ctx->times.ms_diffuse);
            diffuse_pack_kernel<<<div_up((uint64_t)f + 1, 128), 128, 0, ctx->stream>>>(L->indptr, L->indices, L->data, f, recs.as<DiffRec>(), ptr.as<uint32_t>());
            SFB_LAUNCH_CHECK(ctx);
            auto kern = diffuse_tile_kernel<8>;
            SFB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
            int per_sm = 1;
            SFB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem_t));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 4) per_sm = 4;
            const uint64_t tiles = (x->rows + 31) / 32, cap = (uint64_t)ctx->sm_count * per_sm;
            kern<<<(unsigned)(tiles < cap ? tiles : cap), 256, smem_t, ctx->stream>>>(recs.as<DiffRec>(), ptr.as<uint32_t>(), (uint32_t)L->nnz, f, x->d, x->rows, eta, steps);
            SFB_LAUNCH_CHECK(ctx);
            t.stop();   // synchronises: the records may be released
            return SFB_OK;
        }
    }
    size_t per_warp = (size_t)x->cols * 2 * sizeof(double);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
    if (per_warp * wpb > ctx->smem_optin) return sfb_fail(ctx, SFB_EUNSUPPORTED, "feature count %u too large", x->cols);
    size_t smem = per_warp * wpb;
    SFB_CUDA(ctx, cudaFuncSetAttribute(diffuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t want = (x->rows + wpb - 1) / wpb;
    unsigned grid = (unsigned)(want < (uint64_t)ctx->sm_count * 4 ? want : (uint64_t)ctx->sm_count * 4);
    StageTimer t(ctx, &ctx->times.ms_diffuse);
    diffuse_kernel<<<grid, wpb * 32, smem, ctx->stream>>>(L->indptr, L->indices, L->data, x->cols, x->d, x->rows, eta, steps);
    SFB_LAUNCH_CHECK(ctx);
    SFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}
