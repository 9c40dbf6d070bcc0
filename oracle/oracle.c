/*
 * oracle.c -- CPU restatement of the reference's graph-wiring arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (matternet-rs_b200/,
 * include/) may include, link or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs load liboracle.so, and there
 * only as the checker / the CPU baseline.
 *
 * Parity pin: the reference (tuned-org-uk/matternet-rs, a Rust workspace) cannot be
 * compiled here (no cargo/rustc; src_legacy is in no crate and needs the un-vendored
 * `smartcore`).  This restatement is pinned against every known-answer test the
 * reference holds for the path (tests/test_oracle_kat.py):
 *   select_tau table            src_legacy/tests/test_taumode.rs:14-160
 *   L = D - A on 3 points       src_legacy/tests/test_laplacian.rs:655-786
 *   cosine ordering 45/90/180   src_legacy/tests/test_laplacian.rs:156-213
 *   chain-graph lambda / R = 0  surfface-core/src/tests/test_spectral.rs:102-144,187-251
 *   scale invariance            src_legacy/tests/test_taumode.rs:643-682
 * Neighbour identity at the smartcore CosinePair boundary is "parity unpinned" in the
 * reference itself (no test pins which neighbours are chosen); the oracle follows the
 * repo's own brute-force restatement src_legacy/tests/test_helpers.rs:73-133, which
 * defines distance, eps filter and the (distance, index) order.
 *
 * All sums are left folds in ascending index order from 0.0, no FMA contraction
 * (build with -ffp-contract=off), matching Rust's iter().map().sum().
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_METRIC_COSINE 0 /* rectified cosine distance                    */
#define ORC_METRIC_L2 1     /* Euclidean: order and report sqrt(sum (a-b)^2) */
#define ORC_METRIC_L2SQ 2   /* squared Euclidean                             */

#define ORC_TAU_FIXED 0
#define ORC_TAU_MEDIAN 1
#define ORC_TAU_MEAN 2
#define ORC_TAU_PERCENTILE 3

#define ORC_LAMBDA_LEGACY_TAUMODE 0
#define ORC_LAMBDA_ENERGY_NODE 1
#define ORC_LAMBDA_CORE_F32SEM 2

#define ORC_IDX_NONE 0xFFFFFFFFu

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline must still use all host threads */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Synthetic rows: counter-based Philox4x32-10 + Box-Muller with transcendental-free,
 * bit-reproducible log/sin/cos (only + - * / sqrt in a fixed order), so the CUDA generator
 * (matternet-rs_b200/csrc/synth.cuh, written independently from the same spec in DESIGN.md)
 * produces the same bits.  Not part of the reference; the reference's fixtures
 * (src_legacy/tests/test_data.rs:68-238) use Rust RNG streams that cannot be replayed here,
 * so the distributions are re-created, not the bit streams.
 * ------------------------------------------------------------------------------------------ */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                 uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ln(u) for a normal positive double, via u = m*2^e, m in [sqrt(1/2), sqrt(2)),
 * ln m = 2*s*(1 + s2/3 + s2^2/5 + ...), s = (m-1)/(m+1). */
static inline double det_log(double u) {
    uint64_t bits; memcpy(&bits, &u, 8);
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m; memcpy(&m, &bits, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double p = 1.0 / 29.0;
    p = p * s2 + 1.0 / 27.0; p = p * s2 + 1.0 / 25.0; p = p * s2 + 1.0 / 23.0;
    p = p * s2 + 1.0 / 21.0; p = p * s2 + 1.0 / 19.0; p = p * s2 + 1.0 / 17.0;
    p = p * s2 + 1.0 / 15.0; p = p * s2 + 1.0 / 13.0; p = p * s2 + 1.0 / 11.0;
    p = p * s2 + 1.0 / 9.0;  p = p * s2 + 1.0 / 7.0;  p = p * s2 + 1.0 / 5.0;
    p = p * s2 + 1.0 / 3.0;  p = p * s2 + 1.0;
    return 2.0 * s * p + (double)e * 0.6931471805599453;
}

/* sin and cos of 2*pi*v for v = r * 2^-32 (exact), by octant reduction + Taylor. */
static inline void det_sincos_2pi(uint32_t r, double* sn, double* cs) {
    uint32_t q = r >> 29;                              /* octant 0..7        */
    double f = (double)(r & 0x1FFFFFFFu) * (1.0 / 536870912.0); /* frac in [0,1) exact */
    double a = f * 0.7853981633974483;                 /* phi in [0, pi/4)    */
    double a2 = a * a;
    double ps = -1.0 / 121645100408832000.0;           /* -1/19! */
    ps = ps * a2 + 1.0 / 355687428096000.0;            /* 1/17! */
    ps = ps * a2 - 1.0 / 1307674368000.0;              /* 1/15! */
    ps = ps * a2 + 1.0 / 6227020800.0;                 /* 1/13! */
    ps = ps * a2 - 1.0 / 39916800.0;                   /* 1/11! */
    ps = ps * a2 + 1.0 / 362880.0;                     /* 1/9!  */
    ps = ps * a2 - 1.0 / 5040.0;                       /* 1/7!  */
    ps = ps * a2 + 1.0 / 120.0;                        /* 1/5!  */
    ps = ps * a2 - 1.0 / 6.0;                          /* 1/3!  */
    ps = ps * a2 + 1.0;
    double s0 = a * ps;
    double pc = 1.0 / 6402373705728000.0;              /* 1/18! */
    pc = pc * a2 - 1.0 / 20922789888000.0;             /* 1/16! */
    pc = pc * a2 + 1.0 / 87178291200.0;                /* 1/14! */
    pc = pc * a2 - 1.0 / 479001600.0;                  /* 1/12! */
    pc = pc * a2 + 1.0 / 3628800.0;                    /* 1/10! */
    pc = pc * a2 - 1.0 / 40320.0;                      /* 1/8!  */
    pc = pc * a2 + 1.0 / 720.0;                        /* 1/6!  */
    pc = pc * a2 - 1.0 / 24.0;                         /* 1/4!  */
    pc = pc * a2 + 0.5;
    double c0 = 1.0 - a2 * pc;
    const double H = 0.7071067811865476;
    static const double SQ[8] = {0.0, 1.0, 1.0, 1.0, 0.0, -1.0, -1.0, -1.0};  /* sin(q pi/4)/{1,H} */
    static const double CQ[8] = {1.0, 1.0, 0.0, -1.0, -1.0, -1.0, 0.0, 1.0};  /* cos(q pi/4)/{1,H} */
    double sq = SQ[q], cq = CQ[q];
    if (q & 1u) { sq = sq * H; cq = cq * H; }
    *sn = sq * c0 + cq * s0;
    *cs = cq * c0 - sq * s0;
}

/* Four standard normals for (stream, row, quad): counter = (row_lo, row_hi, quad, stream). */
static inline void synth_normal4(uint64_t seed, uint32_t stream, uint64_t row, uint32_t quad,
                                 double z[4]) {
    uint32_t r[4];
    philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), quad, stream, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    for (int h = 0; h < 2; ++h) {
        double u1 = ((double)r[2 * h] + 0.5) * (1.0 / 4294967296.0);
        double rad = sqrt(-2.0 * det_log(u1));
        double sn, cs;
        det_sincos_2pi(r[2 * h + 1], &sn, &cs);
        z[2 * h] = rad * cs;
        z[2 * h + 1] = rad * sn;
    }
}

#define SYNTH_STREAM_NOISE 0u
#define SYNTH_STREAM_CENTRE 1u
#define SYNTH_STREAM_ASSIGN 2u
#define SYNTH_STREAM_SHIFT 3u

/* kind 0: x ~ N(0,1) iid.
 * kind 1: clustered: x_i = c[h(i)] + noise * N(0,I), c_j ~ N(0,I), h(i) = philox(i) mod n_centres
 *         (the law of make_gaussian_hd / make_energy_test_dataset, test_data.rs:118-238).
 * kind 2: anisotropic: x_i = N(0,I) + noise * s_i * 1, s_i ~ N(0,1) one shift per row.     */
void orc_generate_rows(int kind, uint64_t seed, uint64_t row0, uint64_t nrows, uint32_t kdim,
                       uint32_t n_centres, double noise, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < (int64_t)nrows; ++ii) {
        uint64_t i = row0 + (uint64_t)ii;
        double* o = out + (size_t)ii * kdim;
        uint64_t centre = 0; double shift = 0.0;
        if (kind == 1) {
            uint32_t r[4];
            philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0u, SYNTH_STREAM_ASSIGN, (uint32_t)seed,
                          (uint32_t)(seed >> 32), r);
            centre = r[0] % (n_centres ? n_centres : 1u);
        } else if (kind == 2) {
            double z[4]; synth_normal4(seed, SYNTH_STREAM_SHIFT, i, 0u, z); shift = noise * z[0];
        }
        for (uint32_t q = 0; q * 4 < kdim; ++q) {
            double z[4], c[4];
            synth_normal4(seed, SYNTH_STREAM_NOISE, i, q, z);
            if (kind == 1) synth_normal4(seed, SYNTH_STREAM_CENTRE, centre, q, c);
            for (uint32_t t = 0; t < 4 && q * 4 + t < kdim; ++t) {
                double v;
                if (kind == 1) v = c[t] + noise * z[t];
                else if (kind == 2) v = z[t] + shift;
                else v = z[t];
                o[q * 4 + t] = v;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * kNN.  Cosine: src_legacy/tests/test_helpers.rs:77-133 (norms :77-80, denom guard :94-104,
 * rectification :106, eps filter :107, (distance,index) order :116-120, truncate :122-125) and
 * src_legacy/laplacian.rs:245-257 (i != j, dist <= eps).  L2: src_legacy/energymaps.rs:875-892
 * (squared L2, stable sort => ties by index) and surfface-core/src/mst.rs:312-403 (Euclidean
 * orders by the sqrt value).
 * ------------------------------------------------------------------------------------------ */
void orc_row_norms(const double* x, uint64_t m, uint32_t kd, double* norms) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)m; ++i) {
        const double* r = x + (size_t)i * kd;
        double s = 0.0;
        for (uint32_t d = 0; d < kd; ++d) s += r[d] * r[d];
        norms[i] = sqrt(s);
    }
}

static inline double pair_key(const double* a, const double* b, uint32_t kd, int metric, double na,
                              double nb) {
    if (metric == ORC_METRIC_COSINE) {
        double denom = na * nb, cosv = 0.0;
        if (denom > 1e-12) {
            double dot = 0.0;
            for (uint32_t d = 0; d < kd; ++d) dot += a[d] * b[d];
            cosv = dot / denom;
            if (cosv < -1.0) cosv = -1.0; else if (cosv > 1.0) cosv = 1.0;  /* f64::clamp */
        }
        double rect = (cosv > 0.0) ? cosv : 0.0;  /* f64::max(0.0): NaN -> 0.0 */
        return 1.0 - rect;
    }
    double s = 0.0;
    for (uint32_t d = 0; d < kd; ++d) { double t = a[d] - b[d]; s += t * t; }
    return metric == ORC_METRIC_L2 ? sqrt(s) : s;
}

/* 8 corpus rows at a time: eight independent left-fold chains (each pair's own order is
 * untouched), so the CPU baseline is not artificially latency-bound. */
static inline void pair_keys8(const double* a, const double* const b[8], uint32_t kd, int metric,
                              double na, const double nb[8], double out[8]) {
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (metric == ORC_METRIC_COSINE) {
        for (uint32_t d = 0; d < kd; ++d) {
            double av = a[d];
            for (int t = 0; t < 8; ++t) s[t] += av * b[t][d];
        }
        for (int t = 0; t < 8; ++t) {
            double denom = na * nb[t], cosv = 0.0;
            if (denom > 1e-12) {
                cosv = s[t] / denom;
                if (cosv < -1.0) cosv = -1.0; else if (cosv > 1.0) cosv = 1.0;
            }
            double rect = (cosv > 0.0) ? cosv : 0.0;
            out[t] = 1.0 - rect;
        }
    } else {
        for (uint32_t d = 0; d < kd; ++d) {
            double av = a[d];
            for (int t = 0; t < 8; ++t) { double df = av - b[t][d]; s[t] += df * df; }
        }
        for (int t = 0; t < 8; ++t) out[t] = metric == ORC_METRIC_L2 ? sqrt(s[t]) : s[t];
    }
}

static inline int key_less(double da, uint32_t ia, double db, uint32_t ib) {
    return (da < db) || (da == db && ia < ib);
}

/* Insert (d, j) into the sorted top-k list (dist asc, idx asc). */
static inline void topk_insert(double* bd, uint32_t* bi, uint32_t* cnt, uint32_t k, double d,
                               uint32_t j) {
    uint32_t c = *cnt;
    if (c == k) {
        if (!key_less(d, j, bd[k - 1], bi[k - 1])) return;
        c = k - 1;
    }
    uint32_t p = c;
    while (p > 0 && key_less(d, j, bd[p - 1], bi[p - 1])) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
    bd[p] = d; bi[p] = j;
    *cnt = c + 1;
}

/* Brute-force kNN of `nq` query rows (query_rows == NULL: rows 0..nq-1) against all m rows.
 * out_idx/out_dist: nq x k, padded with ORC_IDX_NONE / +inf; out_cnt: nq. */
void orc_knn(const double* x, uint64_t m, uint32_t kd, int metric, uint32_t k, double eps,
             const uint64_t* query_rows, uint64_t nq, uint32_t* out_idx, double* out_dist,
             uint32_t* out_cnt) {
    double* norms = (double*)malloc(sizeof(double) * (size_t)m);
    if (metric == ORC_METRIC_COSINE) orc_row_norms(x, m, kd, norms);
    else memset(norms, 0, sizeof(double) * (size_t)m);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t qq = 0; qq < (int64_t)nq; ++qq) {
        uint64_t i = query_rows ? query_rows[qq] : (uint64_t)qq;
        const double* a = x + (size_t)i * kd;
        double* bd = out_dist + (size_t)qq * k;
        uint32_t* bi = out_idx + (size_t)qq * k;
        uint32_t cnt = 0;
        uint64_t j = 0;
        for (; j + 8 <= m; j += 8) {
            const double* b[8]; double nb[8], key[8];
            for (int t = 0; t < 8; ++t) { b[t] = x + (size_t)(j + t) * kd; nb[t] = norms[j + t]; }
            pair_keys8(a, b, kd, metric, norms[i], nb, key);
            for (int t = 0; t < 8; ++t) {
                if (j + t == i) continue;
                if (key[t] <= eps) topk_insert(bd, bi, &cnt, k, key[t], (uint32_t)(j + t));
            }
        }
        for (; j < m; ++j) {
            if (j == i) continue;
            double key = pair_key(a, x + (size_t)j * kd, kd, metric, norms[i], norms[j]);
            if (key <= eps) topk_insert(bd, bi, &cnt, k, key, (uint32_t)j);
        }
        for (uint32_t t = cnt; t < k; ++t) { bd[t] = INFINITY; bi[t] = ORC_IDX_NONE; }
        out_cnt[qq] = cnt;
    }
    free(norms);
}

/* ------------------------------------------------------------------------------------------
 * Kernel weights + inline sparsification: src_legacy/laplacian.rs:219-290.
 *   degrees = #kNN entries with i != j && dist <= eps (:219-229); sparsify iff mean > 10 (:231-232)
 *   w = 1/(1+(d/sigma)^p), kept iff w > 1e-12 (:255-257)
 *   score = w*sqrt((deg_i*deg_j) as f64) (:260-261); rows with len > 2 keep max(len/2,1) best (:276-282)
 * The reference's sort_unstable_by has no tie-break; the contract fixes (score desc, j asc).
 * adj_idx/adj_w: m x k padded (ORC_IDX_NONE / 0), adj_cnt: m.  force_sparsify: -1 = reference
 * rule, 0 = never, 1 = always.  Returns 1 when sparsification was applied.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t j; double w; double score; } orc_edge;

static int cmp_score_desc(const void* pa, const void* pb) {
    const orc_edge* a = (const orc_edge*)pa; const orc_edge* b = (const orc_edge*)pb;
    if (a->score > b->score) return -1;
    if (a->score < b->score) return 1;
    return (a->j > b->j) - (a->j < b->j);
}

int orc_build_adjacency(const uint32_t* knn_idx, const double* knn_dist, const uint32_t* knn_cnt,
                        uint64_t m, uint32_t k, double p, double sigma, int force_sparsify,
                        uint32_t* adj_idx, double* adj_w, uint32_t* adj_cnt) {
    uint64_t total = 0;
    for (uint64_t i = 0; i < m; ++i) total += knn_cnt[i];
    double avg_degree = (double)total / (double)m;
    int sparsify = force_sparsify < 0 ? (avg_degree > 10.0) : force_sparsify;
#pragma omp parallel
    {
        orc_edge* row = (orc_edge*)malloc(sizeof(orc_edge) * (k ? k : 1));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)m; ++i) {
            uint32_t len = 0;
            for (uint32_t t = 0; t < knn_cnt[i]; ++t) {
                uint32_t j = knn_idx[(size_t)i * k + t];
                double d = knn_dist[(size_t)i * k + t];
                double w = 1.0 / (1.0 + pow(d / sigma, p));
                if (w > 1e-12) {
                    double score = w;
                    if (sparsify)
                        score = w * sqrt((double)((uint64_t)knn_cnt[i] * (uint64_t)knn_cnt[j]));
                    row[len].j = j; row[len].w = w; row[len].score = score; ++len;
                }
            }
            if (sparsify && len > 2) {
                qsort(row, len, sizeof(orc_edge), cmp_score_desc);
                uint32_t keep = len / 2; if (keep < 1) keep = 1;
                len = keep;
            }
            for (uint32_t t = 0; t < k; ++t) {
                adj_idx[(size_t)i * k + t] = t < len ? row[t].j : ORC_IDX_NONE;
                adj_w[(size_t)i * k + t] = t < len ? row[t].w : 0.0;
            }
            adj_cnt[i] = len;
        }
        free(row);
    }
    return sparsify;
}

/* SF-GRASS standalone sparsifier: src_legacy/sparsification.rs:32-113.
 *   skip iff mean degree < 10 (:42-52); degrees = row lengths (:60)
 *   score = w*sqrt((deg_i*deg_j) as f64) (:79-80); keep = clamp(ceil(len*ratio),1,len) (:92-94)
 * ratio is clamped to [0.1, 1.0] (:26-29).  In place.  Returns 1 when applied. */
int orc_sfgrass(uint32_t* adj_idx, double* adj_w, uint32_t* adj_cnt, uint64_t m, uint32_t k,
                double ratio) {
    if (ratio < 0.1) ratio = 0.1;
    if (ratio > 1.0) ratio = 1.0;
    uint64_t total = 0;
    for (uint64_t i = 0; i < m; ++i) total += adj_cnt[i];
    double avg_degree = (double)total / (double)m;
    if (avg_degree < 10.0) return 0;
    uint32_t* deg = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)m);
    memcpy(deg, adj_cnt, sizeof(uint32_t) * (size_t)m);
#pragma omp parallel
    {
        orc_edge* row = (orc_edge*)malloc(sizeof(orc_edge) * (k ? k : 1));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)m; ++i) {
            uint32_t len = deg[i];
            if (len == 0) continue;
            for (uint32_t t = 0; t < len; ++t) {
                uint32_t j = adj_idx[(size_t)i * k + t];
                double w = adj_w[(size_t)i * k + t];
                row[t].j = j; row[t].w = w;
                row[t].score = w * sqrt((double)((uint64_t)deg[i] * (uint64_t)deg[j]));
            }
            qsort(row, len, sizeof(orc_edge), cmp_score_desc);
            uint32_t keep = (uint32_t)ceil((double)len * ratio);
            if (keep < 1) keep = 1;
            if (keep > len) keep = len;
            for (uint32_t t = 0; t < k; ++t) {
                adj_idx[(size_t)i * k + t] = t < keep ? row[t].j : ORC_IDX_NONE;
                adj_w[(size_t)i * k + t] = t < keep ? row[t].w : 0.0;
            }
            adj_cnt[i] = keep;
        }
        free(row);
    }
    free(deg);
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Symmetrise + Laplacian CSR.
 *   symmetrise: src_legacy/laplacian.rs:297-348 -- edge set {(i,j,w),(j,i,w)}, src != dst,
 *     rows sorted by column (:342).  Duplicate (i,j) carry identical w for a symmetric metric;
 *     the contract takes max (surfface-core/src/laplacian.rs:321-331).
 *   unnormalised L: src_legacy/laplacian.rs:364-385 -- L_ii = sum_j w_ij in ascending j
 *     (always stored, :372), L_ij = -w_ij; CSR rows sorted by column (:399, to_csr :161).
 *   normalised L_sym: surfface-core/src/laplacian.rs:333-372,209-219 -- L_ii = 1 if d_i > thr,
 *     L_ij = -w/sqrt(d_i d_j) if d_i, d_j > thr; entries |v| <= 1e-9 dropped.
 * Two-call pattern: orc_laplacian_build returns an opaque handle, nnz read, then copy.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t col; double w; } orc_nb;
typedef struct {
    uint64_t m; uint64_t nnz; uint64_t* indptr; uint32_t* indices; double* data;
} orc_csr;

static int cmp_nb(const void* pa, const void* pb) {
    const orc_nb* a = (const orc_nb*)pa; const orc_nb* b = (const orc_nb*)pb;
    if (a->col != b->col) return (a->col > b->col) - (a->col < b->col);
    return (a->w < b->w) - (a->w > b->w); /* larger w first => max survives dedupe */
}

orc_csr* orc_laplacian_build(const uint32_t* adj_idx, const double* adj_w, const uint32_t* adj_cnt,
                             uint64_t m, uint32_t k, int normalised, double weight_threshold) {
    /* reverse lists by counting sort */
    uint64_t* cnt = (uint64_t*)calloc((size_t)m + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < m; ++i)
        for (uint32_t t = 0; t < adj_cnt[i]; ++t) {
            uint32_t j = adj_idx[(size_t)i * k + t];
            if (j == i) continue;
            cnt[i + 1]++; cnt[(uint64_t)j + 1]++;
        }
    for (uint64_t i = 0; i < m; ++i) cnt[i + 1] += cnt[i];
    uint64_t tot = cnt[m];
    orc_nb* nb = (orc_nb*)malloc(sizeof(orc_nb) * (size_t)(tot ? tot : 1));
    uint64_t* fill = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)m);
    memcpy(fill, cnt, sizeof(uint64_t) * (size_t)m);
    for (uint64_t i = 0; i < m; ++i)
        for (uint32_t t = 0; t < adj_cnt[i]; ++t) {
            uint32_t j = adj_idx[(size_t)i * k + t];
            if (j == i) continue;
            double w = adj_w[(size_t)i * k + t];
            nb[fill[i]].col = j; nb[fill[i]].w = w; fill[i]++;
            nb[fill[j]].col = (uint32_t)i; nb[fill[j]].w = w; fill[j]++;
        }
    /* sort + dedupe each row in place; ulen[i] = unique neighbours */
    uint32_t* ulen = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)m);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)m; ++i) {
        orc_nb* r = nb + cnt[i];
        uint64_t len = cnt[i + 1] - cnt[i];
        qsort(r, (size_t)len, sizeof(orc_nb), cmp_nb);
        uint32_t u = 0;
        for (uint64_t t = 0; t < len; ++t)
            if (u == 0 || r[u - 1].col != r[t].col) r[u++] = r[t];
        ulen[i] = u;
    }
    double* deg = (double*)malloc(sizeof(double) * (size_t)m);
    for (uint64_t i = 0; i < m; ++i) {
        const orc_nb* r = nb + cnt[i];
        double s = 0.0;
        for (uint32_t t = 0; t < ulen[i]; ++t) s += r[t].w;
        deg[i] = s;
    }
    orc_csr* L = (orc_csr*)calloc(1, sizeof(orc_csr));
    L->m = m;
    L->indptr = (uint64_t*)calloc((size_t)m + 1, sizeof(uint64_t));
    /* pass 1: row lengths */
    for (uint64_t i = 0; i < m; ++i) {
        const orc_nb* r = nb + cnt[i];
        uint64_t len = 0;
        if (!normalised) {
            len = (uint64_t)ulen[i] + 1;
        } else {
            if (deg[i] > weight_threshold) len++;
            for (uint32_t t = 0; t < ulen[i]; ++t) {
                double dj = deg[r[t].col];
                if (deg[i] <= weight_threshold || dj <= weight_threshold) continue;
                double v = -r[t].w / sqrt(deg[i] * dj);
                if (fabs(v) > 1e-9) len++;
            }
        }
        L->indptr[i + 1] = L->indptr[i] + len;
    }
    L->nnz = L->indptr[m];
    L->indices = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(L->nnz ? L->nnz : 1));
    L->data = (double*)malloc(sizeof(double) * (size_t)(L->nnz ? L->nnz : 1));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)m; ++i) {
        const orc_nb* r = nb + cnt[i];
        uint64_t o = L->indptr[i];
        int diag_done = 0;
        double dv = normalised ? 1.0 : deg[i];
        int diag_keep = normalised ? (deg[i] > weight_threshold) : 1;
        for (uint32_t t = 0; t < ulen[i]; ++t) {
            if (!diag_done && r[t].col > (uint32_t)i) {
                if (diag_keep) { L->indices[o] = (uint32_t)i; L->data[o] = dv; ++o; }
                diag_done = 1;
            }
            if (!normalised) {
                L->indices[o] = r[t].col; L->data[o] = -r[t].w; ++o;
            } else {
                double dj = deg[r[t].col];
                if (deg[i] <= weight_threshold || dj <= weight_threshold) continue;
                double v = -r[t].w / sqrt(deg[i] * dj);
                if (fabs(v) > 1e-9) { L->indices[o] = r[t].col; L->data[o] = v; ++o; }
            }
        }
        if (!diag_done && diag_keep) { L->indices[o] = (uint32_t)i; L->data[o] = dv; ++o; }
    }
    free(cnt); free(nb); free(fill); free(ulen); free(deg);
    return L;
}
uint64_t orc_csr_nnz(const orc_csr* L) { return L->nnz; }
void orc_csr_copy(const orc_csr* L, uint64_t* indptr, uint32_t* indices, double* data) {
    memcpy(indptr, L->indptr, sizeof(uint64_t) * (size_t)(L->m + 1));
    memcpy(indices, L->indices, sizeof(uint32_t) * (size_t)L->nnz);
    memcpy(data, L->data, sizeof(double) * (size_t)L->nnz);
}
void orc_csr_free(orc_csr* L) {
    if (!L) return;
    free(L->indptr); free(L->indices); free(L->data); free(L);
}

/* ------------------------------------------------------------------------------------------
 * SpMV / Rayleigh: src_legacy/graph.rs:464-501 (row-sequential sum in CSR order), :422-461.
 * ------------------------------------------------------------------------------------------ */
void orc_spmv(const uint64_t* indptr, const uint32_t* indices, const double* data, uint64_t m,
              const double* x, double* y) {
    for (uint64_t r = 0; r < m; ++r) {
        double s = 0.0;
        for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) s += data[e] * x[indices[e]];
        y[r] = s;
    }
}

double orc_rayleigh(const uint64_t* indptr, const uint32_t* indices, const double* data, uint64_t m,
                    const double* x) {
    double* lx = (double*)malloc(sizeof(double) * (size_t)m);
    orc_spmv(indptr, indices, data, m, x, lx);
    double num = 0.0, den = 0.0;
    for (uint64_t r = 0; r < m; ++r) num += x[r] * lx[r];
    for (uint64_t r = 0; r < m; ++r) den += x[r] * x[r];
    free(lx);
    return den > 1e-12 ? num / den : 0.0;
}

/* ------------------------------------------------------------------------------------------
 * tau selection: src_legacy/taumode.rs:29-70 (TAU_FLOOR = 1e-10, :25).
 * ------------------------------------------------------------------------------------------ */
static int cmp_f64(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

double orc_select_tau(const double* e, uint64_t n, int mode, double value) {
    const double FLOOR = 1e-10;
    if (mode == ORC_TAU_FIXED) return (isfinite(value) && value > 0.0) ? value : FLOOR;
    if (mode == ORC_TAU_MEAN) {
        double s = 0.0; uint64_t c = 0;
        for (uint64_t i = 0; i < n; ++i) if (isfinite(e[i])) { s += e[i]; ++c; }
        if (c == 0) return FLOOR;
        double mean = s / (double)c;
        return mean > FLOOR ? mean : FLOOR;
    }
    double* v = (double*)malloc(sizeof(double) * (size_t)(n ? n : 1));
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i) if (isfinite(e[i])) v[c++] = e[i];
    if (c == 0) { free(v); return FLOOR; }
    qsort(v, (size_t)c, sizeof(double), cmp_f64);
    double r;
    if (mode == ORC_TAU_PERCENTILE) {
        double pp = value;
        if (pp < 0.0) pp = 0.0; else if (pp > 1.0) pp = 1.0;
        /* Rust f64::round == C round(): half away from zero */
        uint64_t idx = (uint64_t)round((double)(c - 1) * pp);
        r = v[idx];
    } else {
        r = (c % 2 == 1) ? v[c / 2] : 0.5 * (v[c / 2 - 1] + v[c / 2]);
    }
    free(v);
    return r > FLOOR ? r : FLOOR;
}

/* ------------------------------------------------------------------------------------------
 * Per-item lambda.
 *  LEGACY_TAUMODE  src_legacy/taumode.rs:261-408: zero-vector guard (:268-274, |v| <= 1e-10),
 *     E = max(num/den, 0) with num = sum_r sum_c (x_r*L_rc)*x_c (:340-352), den > 1e-12 (:356),
 *     G over both triangles (:366-408), lambda = tau*(E/(E+tau)) + (1-tau)*clamp(G,0,1) (:306-310).
 *  ENERGY_NODE     src_legacy/energymaps.rs:961-1036: lx = L x (graph.rs:486-493),
 *     lambda = max(sum x_r*lx_r / sum x_r^2, 0); G over the upper triangle only (not blended);
 *     out_g receives G when non-NULL.
 *  CORE_F32SEM     surfface-core/src/spectral/mod.rs:69-181 evaluated in f32:
 *     R = clamp(num/(den+1e-9), +-1e6); e_i = sum_f max(deg_f x_f^2 - 2 x_f (Wx)_f + (W x^2)_f, 0);
 *     G_i = clamp(e_i/(sum_i e_i + 1e-12), 0, 1); lambda = R + G.
 * ------------------------------------------------------------------------------------------ */
/* xt / ft: the vector the zero test and tau are taken from -- the item itself, or its UNPROJECTED form when the item
 * is JL-projected before the Rayleigh quotient (taumode.rs:174-175 select_tau(&item.item), :268-297). */
static double lambda_legacy_row(const uint64_t* indptr, const uint32_t* indices, const double* data,
                                uint64_t f, const double* x, const double* xt, uint64_t ft, int tau_mode, double tau_value,
                                double* out_e, double* out_g) {
    int all_zero = 1;
    for (uint64_t r = 0; r < ft; ++r) if (!(fabs(xt[r]) <= 1e-10)) { all_zero = 0; break; }
    if (all_zero) { if (out_e) *out_e = 0.0; if (out_g) *out_g = 0.0; return 0.0; }
    double tau = orc_select_tau(xt, ft, tau_mode, tau_value);
    double num = 0.0, den = 0.0;
    for (uint64_t r = 0; r < f; ++r) {
        double xi = x[r], rs = 0.0;
        for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) rs += xi * data[e] * x[indices[e]];
        num += rs;
    }
    for (uint64_t r = 0; r < f; ++r) den += x[r] * x[r];
    double e_raw = 0.0;
    if (den > 1e-12) { e_raw = num / den; if (!(e_raw > 0.0)) e_raw = 0.0; }
    double ssum = 0.0;
    for (uint64_t r = 0; r < f; ++r)
        for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
            uint32_t c = indices[e];
            if (c == r) continue;
            double w = -data[e]; if (!(w > 0.0)) continue;
            double d = x[r] - x[c];
            ssum += w * d * d;
        }
    double g = 0.0;
    if (!(ssum <= 1e-12)) {
        double acc = 0.0;
        for (uint64_t r = 0; r < f; ++r)
            for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
                uint32_t c = indices[e];
                if (c == r) continue;
                double w = -data[e]; if (!(w > 0.0)) continue;
                double d = x[r] - x[c];
                double share = (w * d * d) / ssum;
                acc += share * share;
            }
        g = acc < 0.0 ? 0.0 : (acc > 1.0 ? 1.0 : acc);
    }
    if (out_e) *out_e = e_raw;
    if (out_g) *out_g = g;
    double e_bounded = e_raw / (e_raw + tau);
    return tau * e_bounded + (1.0 - tau) * g;
}

static double lambda_energy_row(const uint64_t* indptr, const uint32_t* indices, const double* data,
                                uint64_t f, const double* x, double* lx, double* out_g) {
    orc_spmv(indptr, indices, data, f, x, lx);
    double num = 0.0, den = 0.0;
    for (uint64_t r = 0; r < f; ++r) num += x[r] * lx[r];
    for (uint64_t r = 0; r < f; ++r) den += x[r] * x[r];
    double lam = 0.0;
    if (den > 1e-12) { lam = num / den; if (!(lam > 0.0)) lam = 0.0; }
    if (out_g) {
        /* per-row local sums (energymaps.rs:990-1004), then a left fold over rows */
        double ssum = 0.0;
        for (uint64_t r = 0; r < f; ++r) {
            double local = 0.0;
            for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
                uint32_t c = indices[e];
                if (c <= r) continue;
                double w = -data[e]; if (!(w > 0.0)) continue;
                double d = x[r] - x[c];
                local += w * d * d;
            }
            ssum += local;
        }
        double g = 0.0;
        if (ssum > 1e-12) {
            double acc = 0.0;
            for (uint64_t r = 0; r < f; ++r) {
                double local = 0.0;
                for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
                    uint32_t c = indices[e];
                    if (c <= r) continue;
                    double w = -data[e]; if (!(w > 0.0)) continue;
                    double d = x[r] - x[c];
                    double share = (w * d * d) / ssum;
                    local += share * share;
                }
                acc += local;
            }
            g = acc < 0.0 ? 0.0 : (acc > 1.0 ? 1.0 : acc);
        }
        *out_g = g;
    }
    return lam;
}

/* x: n x f row-major; out_lambda: n; out_e / out_g optional (n each, may be NULL). */
void orc_lambda(const uint64_t* indptr, const uint32_t* indices, const double* data, uint64_t f,
                const double* x, uint64_t n, int variant, int tau_mode, double tau_value,
                double* out_lambda, double* out_e, double* out_g) {
    if (variant == ORC_LAMBDA_CORE_F32SEM) {
        /* dense f32 semantics; row sums in ascending feature order */
        float* W = (float*)calloc((size_t)(f * f), sizeof(float));
        float* Ld = (float*)calloc((size_t)(f * f), sizeof(float));
        for (uint64_t r = 0; r < f; ++r)
            for (uint64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
                float v = (float)data[e];
                Ld[r * f + indices[e]] = v;
                float w = -v; W[r * f + indices[e]] = w > 0.0f ? w : 0.0f;
            }
        float* deg = (float*)calloc((size_t)f, sizeof(float));
        for (uint64_t r = 0; r < f; ++r) { float s = 0.0f; for (uint64_t c = 0; c < f; ++c) s += W[r * f + c]; deg[r] = s; }
        float* rq = (float*)malloc(sizeof(float) * (size_t)n);
        float* en = (float*)malloc(sizeof(float) * (size_t)n);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const double* xd = x + (size_t)i * f;
            float num = 0.0f, den = 0.0f, es = 0.0f;
            for (uint64_t r = 0; r < f; ++r) {
                float xr = (float)xd[r], lx = 0.0f, wx = 0.0f, wx2 = 0.0f;
                for (uint64_t c = 0; c < f; ++c) {
                    float xc = (float)xd[c];
                    lx += Ld[r * f + c] * xc; wx += W[r * f + c] * xc; wx2 += W[r * f + c] * (xc * xc);
                }
                num += xr * lx; den += xr * xr;
                float ee = deg[r] * (xr * xr) - xr * wx * 2.0f + wx2;
                es += ee > 0.0f ? ee : 0.0f;
            }
            float r = num / (den + 1e-9f);
            if (r < -1e6f) r = -1e6f; else if (r > 1e6f) r = 1e6f;
            rq[i] = r; en[i] = es;
        }
        float total = 0.0f;
        for (uint64_t i = 0; i < n; ++i) total += en[i];
        for (uint64_t i = 0; i < n; ++i) {
            float g = en[i] / (total + 1e-12f);
            if (g < 0.0f) g = 0.0f; else if (g > 1.0f) g = 1.0f;
            out_lambda[i] = (double)(rq[i] + g);
            if (out_e) out_e[i] = (double)rq[i];
            if (out_g) out_g[i] = (double)g;
        }
        free(W); free(Ld); free(deg); free(rq); free(en);
        return;
    }
#pragma omp parallel
    {
        double* lx = (double*)malloc(sizeof(double) * (size_t)(f ? f : 1));
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const double* xi = x + (size_t)i * f;
            double e = 0.0, g = 0.0, lam;
            if (variant == ORC_LAMBDA_LEGACY_TAUMODE) {
                lam = lambda_legacy_row(indptr, indices, data, f, xi, xi, f, tau_mode, tau_value, &e, &g);
            } else {
                lam = lambda_energy_row(indptr, indices, data, f, xi, lx, out_g ? &g : NULL);
                e = lam;
            }
            out_lambda[i] = lam;
            if (out_e) out_e[i] = e;
            if (out_g) out_g[i] = g;
        }
        free(lx);
    }
}

/* taumode lambda of JL-projected items: x_proj n x f (f = rows of L), x_orig n x f_orig (tau and zero test). */
void orc_lambda_projected(const uint64_t* indptr, const uint32_t* indices, const double* data, uint64_t f,
                          const double* x_proj, uint64_t n, const double* x_orig, uint64_t f_orig, int tau_mode, double tau_value,
                          double* out_lambda) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        out_lambda[i] = lambda_legacy_row(indptr, indices, data, f, x_proj + (size_t)i * f, x_orig + (size_t)i * f_orig, f_orig,
                                          tau_mode, tau_value, NULL, NULL);
}

/* min-max normalisation: src_legacy/core.rs:1341-1355 (max fold starts at 0.0). stats = {min,max,range}. */
void orc_normalise_lambdas(double* lam, uint64_t n, double* stats) {
    double mn = INFINITY, mx = 0.0;
    for (uint64_t i = 0; i < n; ++i) { if (lam[i] < mn) mn = lam[i]; if (lam[i] > mx) mx = lam[i]; }
    double rng = mx - mn; if (!(rng > 1e-9)) rng = 1e-9;
    for (uint64_t i = 0; i < n; ++i) lam[i] = (lam[i] - mn) / rng;
    if (stats) { stats[0] = mn; stats[1] = mx; stats[2] = rng; }
}

/* diffusion: src_legacy/energymaps.rs:520-546: x_r' = x_r - eta*(L x)_r, `steps` times, in place. */
void orc_diffuse(const uint64_t* indptr, const uint32_t* indices, const double* data, uint64_t f,
                 double* x, uint64_t n, double eta, uint32_t steps) {
    for (uint32_t s = 0; s < steps; ++s) {
#pragma omp parallel
        {
            double* lx = (double*)malloc(sizeof(double) * (size_t)(f ? f : 1));
#pragma omp for schedule(static)
            for (int64_t i = 0; i < (int64_t)n; ++i) {
                double* xi = x + (size_t)i * f;
                orc_spmv(indptr, indices, data, f, xi, lx);
                for (uint64_t r = 0; r < f; ++r) xi[r] = xi[r] - eta * lx[r];
            }
            free(lx);
        }
    }
}

/* row-major transpose (the reference transposes centroids before the graph build, graph.rs:214-216) */
void orc_transpose(const double* x, uint64_t rows, uint64_t cols, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)cols; ++c)
        for (uint64_t r = 0; r < rows; ++r) out[(size_t)c * rows + r] = x[(size_t)r * cols + c];
}

/* ------------------------------------------------------------------------------------------
 * Successor Stage C ("next" row of the scope table): Bhattacharyya-coefficient kNN over the feature
 * nodes, f32 arithmetic as in the reference.
 *   bhattacharyya_coefficient      surfface-core/src/distance.rs:260-290
 *   compute_bhattacharyya_weights  surfface-core/src/laplacian.rs:254-298
 * means / vars are the centroid state [C, F] row-major (feature i at centroid c = a[c*F + i]); the
 * reference transposes first (laplacian.rs:165-166), the arithmetic per pair is the same left fold over c.
 * The reference's sort is unstable with no tie-break; the contract is (BC desc, j asc).
 * libm: logf / expf here are glibc's, the reference's are the platform's -- its own tests compare at 1e-5.
 * ------------------------------------------------------------------------------------------ */
/* Portable f32 log / exp: evaluated in f64 with + - * / only, in a fixed order (no libm, no FMA), then rounded to
 * f32 -- the f64 result is accurate to ~1e-16, so the f32 value is the correctly rounded one except when the exact
 * result lies within ~1e-9 ulp of a rounding boundary.  csrc/bc.cu evaluates the SAME sequence on the device, so the
 * device and this restatement agree bit for bit and a transcendental never decides a neighbour ORDER differently on
 * the two sides.  glibc's logf / expf (what the reference's f32::ln / f32::exp call on Linux)
 * differ from these in < 1 ulp; tests/test_oracle_kat.py measures how often. */
float orc_det_logf(float a) {
    if (a != a || a < 0.0f) return NAN;
    if (a == 0.0f) return -INFINITY;
    if (isinf(a)) return INFINITY;
    return (float)det_log((double)a);   /* every positive f32, subnormals included, is a normal f64 */
}
float orc_det_expf(float xf) {
    if (xf != xf) return NAN;
    double x = (double)xf;
    if (x > 100.0) return INFINITY;     /* expf overflows above 88.73 */
    if (x < -120.0) return 0.0f;        /* below the smallest f32 subnormal (ln = -103.3) */
    double kf = rint(x * 1.4426950408889634);
    double r = (x - kf * 6.93147180369123816490e-01) - kf * 1.90821492927058770002e-10;
    double p = 1.0 / 87178291200.0;                 /* 1/14! */
    p = p * r + 1.0 / 6227020800.0;  p = p * r + 1.0 / 479001600.0; p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;     p = p * r + 1.0 / 362880.0;    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;        p = p * r + 1.0 / 720.0;       p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;          p = p * r + 1.0 / 6.0;         p = p * r + 0.5;
    p = p * r + 1.0;                 p = p * r + 1.0;
    int64_t k = (int64_t)kf;
    uint64_t bits = (uint64_t)(k + 1023) << 52;     /* |k| <= 174: a normal f64 */
    double sc; memcpy(&sc, &bits, 8);
    return (float)(p * sc);
}

static inline float bc_pair(const float* means, const float* vars, uint32_t c, uint32_t f, uint32_t i, uint32_t j, float reg, int det) {
    float db = 0.0f;
    for (uint32_t cc = 0; cc < c; ++cc) {
        float vi = vars[(size_t)cc * f + i], vj = vars[(size_t)cc * f + j];
        vi = vi > reg ? vi : reg;                       /* f32::max ignores a NaN operand: NaN > reg is false */
        vj = vj > reg ? vj : reg;
        float v_sum = vi + vj;
        float dm = means[(size_t)cc * f + i] - means[(size_t)cc * f + j];
        float mean_term = (dm * dm) / (4.0f * v_sum);
        float arg = v_sum / (2.0f * sqrtf(vi * vj));
        float log_term = 0.5f * (det ? orc_det_logf(arg) : logf(arg));
        db += mean_term + log_term;
    }
    float bc = det ? orc_det_expf(-db) : expf(-db);
    if (bc < 0.0f) bc = 0.0f;
    if (bc > 1.0f) bc = 1.0f;
    return bc;
}
/* libm form (glibc logf / expf: the reference on Linux) and portable form (what the device evaluates) */
float orc_bc(const float* means, const float* vars, uint32_t c, uint32_t f, uint32_t i, uint32_t j, float reg) {
    return bc_pair(means, vars, c, f, i, j, reg, 0);
}
float orc_bc_det(const float* means, const float* vars, uint32_t c, uint32_t f, uint32_t i, uint32_t j, float reg) {
    return bc_pair(means, vars, c, f, i, j, reg, 1);
}

void orc_bc_matrix2(const float* means, const float* vars, uint32_t c, uint32_t f, float reg, int det, float* out /* f*f */) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < (int64_t)f; ++i)
        for (uint32_t j = 0; j < f; ++j) out[(size_t)i * f + j] = bc_pair(means, vars, c, f, (uint32_t)i, j, reg, det);
}
void orc_bc_matrix(const float* means, const float* vars, uint32_t c, uint32_t f, float reg, float* out /* f*f */) {
    orc_bc_matrix2(means, vars, c, f, reg, 0, out);
}

typedef struct { float w; uint32_t j; } orc_bc_pair;
static int orc_bc_cmp(const void* a, const void* b) {
    const orc_bc_pair* x = (const orc_bc_pair*)a; const orc_bc_pair* y = (const orc_bc_pair*)b;
    if (x->w > y->w) return -1;
    if (x->w < y->w) return 1;
    return x->j < y->j ? -1 : (x->j > y->j ? 1 : 0);
}

/* out_idx / out_w: f x k (padded with IDX_NONE / 0), out_cnt: f */
void orc_bc_knn2(const float* means, const float* vars, uint32_t c, uint32_t f, uint32_t k, float reg, float thr, int det,
                 uint32_t* out_idx, float* out_w, uint32_t* out_cnt) {
    uint32_t kk = k < (f ? f - 1 : 0) ? k : (f ? f - 1 : 0);
#pragma omp parallel
    {
        orc_bc_pair* sc = (orc_bc_pair*)malloc(sizeof(orc_bc_pair) * (size_t)(f ? f : 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < (int64_t)f; ++i) {
            uint32_t n = 0;
            for (uint32_t j = 0; j < f; ++j) {
                if (j == (uint32_t)i) continue;
                float w = bc_pair(means, vars, c, f, (uint32_t)i, j, reg, det);
                if (w > thr) { sc[n].w = w; sc[n].j = j; ++n; }
            }
            qsort(sc, n, sizeof(orc_bc_pair), orc_bc_cmp);
            uint32_t keep = n < kk ? n : kk;
            for (uint32_t t = 0; t < k; ++t) {
                out_idx[(size_t)i * k + t] = t < keep ? sc[t].j : ORC_IDX_NONE;
                out_w[(size_t)i * k + t] = t < keep ? sc[t].w : 0.0f;
            }
            out_cnt[i] = keep;
        }
        free(sc);
    }
}

void orc_bc_knn(const float* means, const float* vars, uint32_t c, uint32_t f, uint32_t k, float reg, float thr,
                uint32_t* out_idx, float* out_w, uint32_t* out_cnt) {
    orc_bc_knn2(means, vars, c, f, k, reg, thr, 0, out_idx, out_w, out_cnt);
}

/* ------------------------------------------------------------------------------------------
 * Energy pipeline, item -> sub-centroid mapping (src_legacy/energymaps.rs:1246-1342):
 *   best = first s minimising |lambda_item - lambda_s| (strict <, ascending s);
 *   candidates = { s : | |lambda_item - lambda_s| - best_dist | < epsilon } (epsilon = 1e-11 in the reference);
 *   more than one: the candidate with the strictly largest cosine to the item (ascending s; cosine = 0 when either
 *   norm is 0; dot and |centroid|^2 are left folds over the features);
 *   returns (index, lambda of the chosen sub-centroid, |item| as sqrt of the left-fold sum of squares).
 * ------------------------------------------------------------------------------------------ */
void orc_map_items(const double* items, uint64_t n, uint32_t f, const double* item_lambdas, const double* subc, uint32_t s,
                   const double* sub_lambdas, double epsilon, uint32_t* out_idx, double* out_lambda, double* out_norm) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const double* x = items + (size_t)i * f;
        const double li = item_lambdas[i];
        uint32_t best = 0; double best_d = INFINITY;
        for (uint32_t c = 0; c < s; ++c) { double d = fabs(li - sub_lambdas[c]); if (d < best_d) { best_d = d; best = c; } }
        double nsq = 0.0;
        for (uint32_t t = 0; t < f; ++t) nsq += x[t] * x[t];
        const double norm = sqrt(nsq);
        uint32_t ncand = 0;
        for (uint32_t c = 0; c < s; ++c) if (fabs(fabs(li - sub_lambdas[c]) - best_d) < epsilon) ++ncand;
        if (ncand > 1) {
            double best_cos = -INFINITY; uint32_t best_sc = best;
            for (uint32_t c = 0; c < s; ++c) {
                if (!(fabs(fabs(li - sub_lambdas[c]) - best_d) < epsilon)) continue;
                const double* y = subc + (size_t)c * f;
                double dot = 0.0, cn = 0.0;
                for (uint32_t t = 0; t < f; ++t) { dot += x[t] * y[t]; cn += y[t] * y[t]; }
                double cnorm = sqrt(cn);
                double cosv = (norm > 0.0 && cnorm > 0.0) ? dot / (norm * cnorm) : 0.0;
                if (cosv > best_cos) { best_cos = cosv; best_sc = c; }
            }
            best = best_sc;
        }
        out_idx[i] = best; out_lambda[i] = sub_lambdas[best]; out_norm[i] = norm;
    }
}

/* ---- JL projection of items ahead of lambda ("next" row 3) --------------------------------------
 * compute_jl_dimension: src_legacy/reduction.rs:117-171 (f64) and surfface-core/src/clustering.rs:113-123 (f32).
 * Pinned by the exact values in src_legacy/tests/test_reduction.rs:193-345 (tests/test_oracle_kat.py).
 * Rust's `as usize` saturates (negative and NaN -> 0). */
static uint64_t sat_usize(double v) { return (v != v || v <= 0.0) ? 0 : (v >= 1.8446744073709552e19 ? UINT64_MAX : (uint64_t)v); }
static uint64_t clamp_u64(uint64_t v, uint64_t lo, uint64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

uint64_t orc_jl_dimension(uint64_t n_points, uint64_t original_dim, double epsilon) {
    if (original_dim < 32) return original_dim;
    double log_n = log((double)n_points);
    double eps_sq = pow(epsilon, 2.0);
    uint64_t jl_bound = sat_usize(ceil(8.0 * log_n / eps_sq));
    if (original_dim > 2048) {
        double ratio = (double)original_dim / (double)jl_bound;
        double buffer = ratio < 10.0 ? 1.2 : (ratio < 100.0 ? 1.5 : 2.0);
        return clamp_u64(sat_usize(ceil((double)jl_bound * buffer)), 32, original_dim);
    }
    return clamp_u64(jl_bound, 32, original_dim);
}

uint64_t orc_jl_dimension_core(uint64_t n_points, uint64_t original_dim, float epsilon) {
    if (original_dim < 32) return original_dim;
    float log_n = logf((float)n_points);
    float eps_sq = epsilon * epsilon; /* powi(2) */
    return clamp_u64(sat_usize((double)ceilf(8.0f * log_n / eps_sq)), 32, original_dim);
}

/* ImplicitProjection::project (src_legacy/reduction.rs:225-242) with the StandardNormal samples handed in as a
 * matrix: samples[i*r + j] is the draw made for (original i, reduced j), the order the reference's nested loop
 * consumes its ChaCha8 stream in.  out[j] = fold over i of  acc + (x_i * s_ij) * scale,  scale = 1/sqrt(r).
 * project_matrix (reduction.rs:175-200) applies it to every row. */
void orc_project_rows(const double* x, uint64_t n, uint32_t f, const double* samples, uint32_t r, double* out) {
    const double scale = 1.0 / sqrt((double)r);
#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < (int64_t)n; row++) {
        double* o = out + (uint64_t)row * r;
        for (uint32_t j = 0; j < r; j++) o[j] = 0.0;
        for (uint32_t i = 0; i < f; i++) {
            double xi = x[(uint64_t)row * f + i];
            for (uint32_t j = 0; j < r; j++) o[j] = o[j] + (xi * samples[(uint64_t)i * r + j]) * scale;
        }
    }
}

/* successor: surfface-core/src/clustering.rs:84-109, f32; samples[j*f + i] (the draw order there is j-major):
 * out[j] = (fold over i of acc + x_i * s_ji) * scale. */
void orc_project_rows_core(const float* x, uint64_t n, uint32_t f, const float* samples, uint32_t r, float* out) {
    const float scale = 1.0f / sqrtf((float)r);
#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < (int64_t)n; row++)
        for (uint32_t j = 0; j < r; j++) {
            float sum = 0.0f;
            for (uint32_t i = 0; i < f; i++) sum = sum + x[(uint64_t)row * f + i] * samples[(uint64_t)j * f + i];
            out[(uint64_t)row * r + j] = sum * scale;
        }
}

/* ---- SortedLambdas::build_from ("next" row 4; src_legacy/sorted_index.rs:22-46, laplacian.rs:421-448) ----------
 * BTreeMap<OrderedFloat<f64>, Vec<(idx, idx.to_string())>>: keys ascending in OrderedFloat's total order (-0 == +0,
 * NaN == NaN and greater than everything); inside a bucket the items are ordered by their DECIMAL STRING id
 * ("10" < "2"); the key a bucket reports is the one first inserted (lowest idx).  to_vec() flattens that.
 * std_dev: f32 arithmetic on a sequential f64 sum, as laplacian.rs:421-448 writes it. */
static const double* g_sl_lam;
static int of_cmp(double a, double b) {
    int an = a != a, bn = b != b;
    if (an || bn) return an - bn; /* NaN greatest, NaN == NaN */
    return (a > b) - (a < b);     /* -0 == +0 */
}
static int sl_cmp(const void* pa, const void* pb) {
    uint32_t a = *(const uint32_t*)pa, b = *(const uint32_t*)pb;
    int c = of_cmp(g_sl_lam[a], g_sl_lam[b]);
    if (c) return c;
    char sa[24], sb[24];
    snprintf(sa, sizeof sa, "%u", a);
    snprintf(sb, sizeof sb, "%u", b);
    return strcmp(sa, sb);
}
int orc_sorted_lambdas(const double* lam, uint64_t n, double* out_lambda, uint32_t* out_idx, double* out_std_dev) {
    if (n == 0) return -1; /* mean() is None -> build_from panics */
    double sum = 0.0;
    for (uint64_t i = 0; i < n; i++) sum = sum + lam[i];
    float mean = (float)sum / (float)n;
    float var = 0.0f;
    for (uint64_t i = 0; i < n; i++) {
        float diff = mean - (float)lam[i];
        var = var + diff * diff;
    }
    var = var / (float)n;
    *out_std_dev = (double)sqrtf(var);
    for (uint64_t i = 0; i < n; i++) out_idx[i] = (uint32_t)i;
    g_sl_lam = lam;
    qsort(out_idx, n, sizeof(uint32_t), sl_cmp);
    for (uint64_t a = 0; a < n;) {
        uint64_t b = a;
        uint32_t first = out_idx[a];
        while (b < n && of_cmp(lam[out_idx[b]], lam[out_idx[a]]) == 0) { if (out_idx[b] < first) first = out_idx[b]; b++; }
        for (uint64_t t = a; t < b; t++) out_lambda[t] = lam[first];
        a = b;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Streaming forms for corpora that never exist on the host in full (C3: 10M x 768 f64 = 61 GB).
 * Same arithmetic and order as orc_knn; the caller feeds the corpus in ascending row blocks.
 *   q: nq x kd query vectors, q_norms: their norms (orc_row_norms; unused for L2), q_global: their row indices;
 *   io_idx / io_dist / io_cnt: running top-k lists (io_cnt zeroed before the first block).
 * orc_knn_block_finish pads the lists like orc_knn.
 * ------------------------------------------------------------------------------------------ */
void orc_knn_block(const double* q, const double* q_norms, const uint64_t* q_global, uint64_t nq, const double* block,
                   uint64_t block_row0, uint64_t block_rows, uint32_t kd, int metric, uint32_t k, double eps,
                   uint32_t* io_idx, double* io_dist, uint32_t* io_cnt) {
    double* norms = (double*)malloc(sizeof(double) * (size_t)(block_rows ? block_rows : 1));
    if (metric == ORC_METRIC_COSINE) orc_row_norms(block, block_rows, kd, norms);
    else memset(norms, 0, sizeof(double) * (size_t)block_rows);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qq = 0; qq < (int64_t)nq; ++qq) {
        const double* a = q + (size_t)qq * kd;
        const double na = metric == ORC_METRIC_COSINE ? q_norms[qq] : 0.0;
        double* bd = io_dist + (size_t)qq * k;
        uint32_t* bi = io_idx + (size_t)qq * k;
        uint32_t cnt = io_cnt[qq];
        uint64_t j = 0;
        for (; j + 8 <= block_rows; j += 8) {
            const double* b[8]; double nb[8], key[8];
            for (int t = 0; t < 8; ++t) { b[t] = block + (size_t)(j + t) * kd; nb[t] = norms[j + t]; }
            pair_keys8(a, b, kd, metric, na, nb, key);
            for (int t = 0; t < 8; ++t) {
                if (block_row0 + j + t == q_global[qq]) continue;
                if (key[t] <= eps) topk_insert(bd, bi, &cnt, k, key[t], (uint32_t)(block_row0 + j + t));
            }
        }
        for (; j < block_rows; ++j) {
            if (block_row0 + j == q_global[qq]) continue;
            double key = pair_key(a, block + (size_t)j * kd, kd, metric, na, norms[j]);
            if (key <= eps) topk_insert(bd, bi, &cnt, k, key, (uint32_t)(block_row0 + j));
        }
        io_cnt[qq] = cnt;
    }
    free(norms);
}
void orc_knn_block_finish(uint64_t nq, uint32_t k, uint32_t* io_idx, double* io_dist, const uint32_t* io_cnt) {
    for (uint64_t qq = 0; qq < nq; ++qq)
        for (uint32_t t = io_cnt[qq]; t < k; ++t) { io_dist[qq * k + t] = INFINITY; io_idx[qq * k + t] = ORC_IDX_NONE; }
}

/* Feature graph (nodes = columns) of a row-streamed item matrix, for a SAMPLE of the feature nodes: the dot
 * products of the sampled columns with every column and every column's sum of squares, each a left fold over the
 * item rows carried across blocks -- the sums orc_knn(orc_transpose(x)) forms (pair_key, orc_row_norms).
 *   acc: ns x f, nrm2: f, both zeroed before the first block. */
void orc_cols_gram_accumulate(const double* block, uint64_t rows, uint32_t f, const uint32_t* cols, uint32_t ns,
                              double* acc, double* nrm2) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)f; ++c) {
        double n2 = nrm2[c];
        for (uint64_t r = 0; r < rows; ++r) { double v = block[(size_t)r * f + c]; n2 += v * v; }
        nrm2[c] = n2;
        for (uint32_t s = 0; s < ns; ++s) {
            double a = acc[(size_t)s * f + c];
            const uint32_t cs = cols[s];
            for (uint64_t r = 0; r < rows; ++r) a += block[(size_t)r * f + cs] * block[(size_t)r * f + c];
            acc[(size_t)s * f + c] = a;
        }
    }
}
/* rectified-cosine kNN lists of the sampled feature nodes from those sums (pair_key's cosine branch) */
void orc_cols_gram_finish(const double* acc, const double* nrm2, uint32_t f, const uint32_t* cols, uint32_t ns, uint32_t k,
                          double eps, uint32_t* out_idx, double* out_dist, uint32_t* out_cnt) {
    for (uint32_t s = 0; s < ns; ++s) {
        const uint32_t i = cols[s];
        const double na = sqrt(nrm2[i]);
        double* bd = out_dist + (size_t)s * k; uint32_t* bi = out_idx + (size_t)s * k; uint32_t cnt = 0;
        for (uint32_t j = 0; j < f; ++j) {
            if (j == i) continue;
            double denom = na * sqrt(nrm2[j]), cosv = 0.0;
            if (denom > 1e-12) {
                cosv = acc[(size_t)s * f + j] / denom;
                if (cosv < -1.0) cosv = -1.0; else if (cosv > 1.0) cosv = 1.0;
            }
            double key = 1.0 - ((cosv > 0.0) ? cosv : 0.0);
            if (key <= eps) topk_insert(bd, bi, &cnt, k, key, j);
        }
        for (uint32_t t = cnt; t < k; ++t) { bd[t] = INFINITY; bi[t] = ORC_IDX_NONE; }
        out_cnt[s] = cnt;
    }
}

/* ------------------------------------------------------------------------------------------
 * The successor's tau: compute_tau (surfface-core/src/taumode.rs:37-65) -- ONE tau resolved from the lambda
 * DISTRIBUTION in f32 (TAU_FLOOR = 1e-9, :9): finite entries only; Fixed(t): t.max(floor) if finite else floor;
 * Mean: left-fold f32 sum / len; Median: sorted[len / 2] (upper median, no averaging); Percentile(p):
 * sorted[round((len - 1) as f32 * clamp(p, 0, 1))]; every result .max(floor).
 * ------------------------------------------------------------------------------------------ */
static int cmp_f32(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}
float orc_compute_tau_core(const float* lambdas, uint64_t n, int mode, float value) {
    const float FLOOR = 1e-9f;
    float* v = (float*)malloc(sizeof(float) * (size_t)(n ? n : 1));
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i) if (isfinite(lambdas[i])) v[c++] = lambdas[i];
    float r = FLOOR;
    if (c != 0) {
        if (mode == ORC_TAU_FIXED) r = isfinite(value) ? (value > FLOOR ? value : FLOOR) : FLOOR;
        else if (mode == ORC_TAU_MEAN) {
            float s = 0.0f;
            for (uint64_t i = 0; i < c; ++i) s += v[i];
            r = s / (float)c;
        } else {
            qsort(v, (size_t)c, sizeof(float), cmp_f32);
            if (mode == ORC_TAU_MEDIAN) r = v[c / 2];
            else {
                float pp = value;
                if (pp < 0.0f) pp = 0.0f; else if (pp > 1.0f) pp = 1.0f;
                float fi = roundf(((float)c - 1.0f) * pp);                 /* f32::round: half away from zero */
                uint64_t idx = fi != fi ? 0 : (uint64_t)fi;                /* `as usize` maps NaN to 0 */
                if (idx >= c) idx = c - 1;
                r = v[idx];
            }
        }
        if (!(r > FLOOR)) r = FLOOR;   /* f32::max(FLOOR): a NaN mean (inf - inf cannot occur: finite inputs; inf sum / c = inf) */
    }
    free(v);
    return r;
}
