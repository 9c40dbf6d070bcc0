/*
 * surfface_b200.h -- C ABI of libsurfface_b200.so: the B200 (sm_100a) graph-wiring build path
 * of surfface (tuned-org-uk/matternet-rs): kNN graph -> sparse Laplacian -> taumode lambda.
 *
 * The reference has NO FFI today: the seams are Rust functions.  Every
 * entry point below names the reference function it stands behind (paths relative to the
 * reference repo).  A Rust `surfface-b200-sys` crate binds exactly these symbols (see
 * INTEGRATION.md and matternet-rs_b200/rust/).
 *
 * Conventions
 *   ownership  caller owns every host buffer; the library owns device memory behind opaque
 *              handles (sfb_mat / sfb_knn / sfb_adj / sfb_csr) released with the matching
 *              *_free.  Reference: inputs are borrowed slices, outputs owned Vec / CsMat.
 *   errors     every call returns an sfb_status; sfb_last_error(ctx) holds the message.  The
 *              reference panics (src_legacy/laplacian.rs:130-135, taumode.rs:329-337); a Rust
 *              wrapper panics on non-zero to keep those signatures.
 *   threading  calls are blocking; one sfb_ctx per host thread (not internally locked).
 *   fallback   there is none: without a CUDA device sfb_ctx_create returns SFB_ECUDA.
 *   layout     matrices are row-major f64; kNN / adjacency lists are M x k, padded with
 *              SFB_IDX_NONE (+inf distance / 0 weight) past the row's count; CSR uses u64
 *              indptr, u32 indices (ascending per row), f64 data.
 */
#ifndef SURFFACE_B200_H
#define SURFFACE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_ABI_VERSION 4
#define SFB_IDX_NONE 0xFFFFFFFFu

typedef enum {
    SFB_OK = 0,
    SFB_EINVAL = 1,       /* bad argument (the reference would assert!/panic!)            */
    SFB_ECUDA = 2,        /* CUDA runtime / driver error, or no device                    */
    SFB_ENOMEM = 3,       /* device or host allocation failed                             */
    SFB_ENCCL = 4,        /* NCCL error / communicator not initialised                    */
    SFB_EUNCERTIFIED = 5, /* screen could not certify rows and exact fallback was disabled */
    SFB_EUNSUPPORTED = 6  /* valid request outside what this build implements             */
} sfb_status;

typedef struct sfb_ctx sfb_ctx;
typedef struct sfb_mat sfb_mat; /* dense row-major f64 matrix resident in HBM        */
typedef struct sfb_knn sfb_knn; /* per-row k nearest neighbours (idx, dist, count)   */
typedef struct sfb_adj sfb_adj; /* weighted directed adjacency lists (idx, w, count) */
typedef struct sfb_csr sfb_csr; /* CSR matrix (graph Laplacian) resident in HBM      */

/* ---- context -------------------------------------------------------------------------------
 * Replaces surfface-core/src/backend.rs:18-70 (SurffaceDevice::Cuda / get_device / dispatch). */
int32_t sfb_abi_version(void);
int32_t sfb_ctx_create(int32_t device_id, sfb_ctx** out);
void sfb_ctx_destroy(sfb_ctx* ctx);
const char* sfb_last_error(const sfb_ctx* ctx);
/* name: >= 256 bytes; sm_count, hbm_bytes may be NULL */
int32_t sfb_device_info(const sfb_ctx* ctx, char* name, int32_t* sm_count, uint64_t* hbm_bytes);
int32_t sfb_synchronize(sfb_ctx* ctx);
/* page-locked host memory for the host<->device copies of the one-shot entry points */
int32_t sfb_pinned_alloc(sfb_ctx* ctx, uint64_t bytes, void** out);
void sfb_pinned_free(void* p);

/* ---- dense matrices ------------------------------------------------------------------------
 * The reference moves whole flat Vec copies host<->device (surfface-core/src/laplacian.rs:157-158,
 * spectral/mod.rs:39-51,170-171).  sfb_mat_from_host is that upload. */
int32_t sfb_mat_from_host(sfb_ctx* ctx, const double* x, uint64_t rows, uint32_t cols, sfb_mat** out);
/* The successor's item matrices are f32 (compute_tau_mode_gpu(&LaplacianOutput, data: &[f32], ..),
 * surfface-core/src/spectral/bridge.rs:27-32): 4 bytes per value cross PCIe, widened to f64 on the device (exact). */
int32_t sfb_mat_from_host_f32(sfb_ctx* ctx, const float* x, uint64_t rows, uint32_t cols, sfb_mat** out);
/* device copy (sfb_diffuse works in place; the reference's diffusion clones its matrix first, energymaps.rs:520-524) */
int32_t sfb_mat_clone(sfb_ctx* ctx, const sfb_mat* a, sfb_mat** out);
/* Synthetic rows generated on the device: counter-based Philox4x32-10 +
 * Box-Muller with reproducible arithmetic, so any row can be regenerated on the CPU.
 * kind 0: N(0,1) iid; 1: clustered (n_centres centres ~ N(0,I), x = c + noise*N(0,I));
 * 2: N(0,I) + noise * (one N(0,1) shift per row). */
int32_t sfb_mat_generate(sfb_ctx* ctx, int32_t kind, uint64_t seed, uint64_t rows, uint32_t cols,
                         uint32_t n_centres, double noise, sfb_mat** out);
/* GraphFactory::build_laplacian_matrix_from_k_cluster transposes before the graph build
 * (src_legacy/graph.rs:214-216): nodes become the columns. */
int32_t sfb_mat_transpose(sfb_ctx* ctx, const sfb_mat* a, sfb_mat** out);
/* non-owning view of rows [row0, row0 + nrows) (the parent must outlive it); free with sfb_mat_free */
int32_t sfb_mat_view_rows(sfb_ctx* ctx, const sfb_mat* a, uint64_t row0, uint64_t nrows, sfb_mat** out);
int32_t sfb_mat_shape(const sfb_mat* a, uint64_t* rows, uint32_t* cols);
int32_t sfb_mat_copy_rows(sfb_ctx* ctx, const sfb_mat* a, uint64_t row0, uint64_t nrows, double* out);
void sfb_mat_free(sfb_mat* a);

/* ---- kNN graph -----------------------------------------------------------------------------
 * Replaces the two CosinePair sweeps of _build_adjacency (src_legacy/laplacian.rs:213-229,245-254),
 * with the semantics of the repo's brute-force restatement (src_legacy/tests/test_helpers.rs:77-125):
 *   cosine: d = 1 - max(0, clamp(x_i.x_j / (|x_i||x_j|), -1, 1)), cos := 0 when |x_i||x_j| <= 1e-12
 *   L2 / L2SQ: surfface-core/src/mst.rs:312-403, src_legacy/energymaps.rs:875-892
 *   keep j != i with d <= eps, order (d asc, j asc), first k.
 * Neighbour indices are bit-exact with that definition whichever screen is used: the tensor-core
 * screen only proposes candidates, every returned distance is the exact f64 value, and rows the
 * screen cannot certify are recomputed by brute force in f64. */
typedef enum { SFB_METRIC_COSINE = 0, SFB_METRIC_L2 = 1, SFB_METRIC_L2SQ = 2 } sfb_metric;
typedef enum {
    SFB_SCREEN_AUTO = 0,      /* tensor-core screen when the shape allows, else exact       */
    SFB_SCREEN_EXACT_F64 = 1, /* brute force in f64 (reference arithmetic, no screen)       */
    SFB_SCREEN_F16 = 2,       /* tcgen05 kind::f16, fp16 operands, fp32 accumulate in TMEM  */
    SFB_SCREEN_BF16 = 3       /* tcgen05 kind::f16, bf16 operands (8x looser margin)        */
} sfb_screen;

typedef struct {
    int32_t metric;    /* sfb_metric                                                         */
    uint32_t k;        /* neighbours kept per row, self excluded (GraphParams.topk)          */
    double eps;        /* keep d <= eps; +inf disables (GraphParams.eps)                     */
    int32_t screen;    /* sfb_screen                                                         */
    uint32_t k_prime;  /* screen candidates per row at the first level (0 = default max(1.5k, 32)) */
    uint64_t q_begin;  /* query-row shard [q_begin, q_end); q_end = 0 means all rows         */
    uint64_t q_end;
    int32_t allow_fallback; /* 0: return SFB_EUNCERTIFIED instead of recomputing rows exactly */
} sfb_knn_params;

typedef struct {
    uint64_t rows;            /* query rows computed                                         */
    uint64_t rows_certified;  /* rows whose screen margin proved the candidate set complete  */
    uint64_t rows_fallback;   /* rows recomputed by exact f64 brute force                    */
    uint32_t k_prime;         /* candidates rescored per row                                 */
    int32_t screen_used;      /* sfb_screen actually run                                     */
    double ms_prepare;        /* norms + operand conversion                                  */
    double ms_screen;         /* tensor-core distance screen + top-k' selection              */
    double ms_rescore;        /* exact f64 rescore + certification                           */
    double ms_fallback;       /* exact brute force for uncertified rows                      */
    double max_margin;        /* largest per-row screen margin used                          */
    uint64_t rows_rescreened; /* rows the first screen level left to the k' = 192 re-screen   */
    double ms_rescreen;       /* gather + re-screen + rescore of those rows                   */
    uint64_t candidates_rescored; /* corpus rows gathered and rescored in f64: rescore traffic = this x dims x 8 bytes */
} sfb_knn_stats;

int32_t sfb_knn_build(sfb_ctx* ctx, const sfb_mat* rows, const sfb_knn_params* params, sfb_knn** out);
/* The same graph over the COLUMNS of x (x is dims x nodes): the feature graph of
 * GraphFactory::build_laplacian_matrix_from_k_cluster (src_legacy/graph.rs:193-216) straight from the
 * item matrix, without materialising the transposed copy the reference makes. */
int32_t sfb_knn_build_columns(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* params, sfb_knn** out);
int32_t sfb_knn_shape(const sfb_knn* g, uint64_t* rows, uint32_t* k, uint64_t* q_begin);
/* idx, dist: rows x k; cnt: rows.  Any pointer may be NULL. */
int32_t sfb_knn_copy(sfb_ctx* ctx, const sfb_knn* g, uint32_t* idx, double* dist, uint32_t* cnt);
int32_t sfb_knn_stats_get(const sfb_knn* g, sfb_knn_stats* out);
int32_t sfb_knn_from_host(sfb_ctx* ctx, const uint32_t* idx, const double* dist, const uint32_t* cnt,
                          uint64_t rows, uint32_t k, sfb_knn** out);
void sfb_knn_free(sfb_knn* g);

/* The feature graph hidden behind the item graph's screen: _begin registers the exact f64 pair sums of the
 * columns of x (feature-graph shape only; anything else is computed in _end by the plain path); the next
 * sfb_knn_build on this context launches them on a side stream beside its tensor-core kernel; _end joins
 * (all-reduces when sharded != 0: collective, like sfb_knn_build_columns_sharded) and returns the graph.
 * x must stay alive and unchanged until _end; one pending build per context. */
typedef struct sfb_pending sfb_pending;
int32_t sfb_knn_build_columns_begin(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* params, int32_t sharded, sfb_pending** out);
int32_t sfb_knn_build_columns_end(sfb_ctx* ctx, sfb_pending* pending, sfb_knn** out);

/* ---- kernel weights + sparsification -------------------------------------------------------
 * sfb_adjacency_build: src_legacy/laplacian.rs:231-290 -- w = 1/(1+(d/sigma)^p), keep w > 1e-12;
 *   inline sparsification when mean degree > 10: score = w*sqrt(deg_i*deg_j), rows with more than
 *   2 edges keep their max(len/2,1) best.  sparsify: -1 reference rule, 0 never, 1 always.
 *   The reference's sort is unstable with no tie-break; the contract is (score desc, j asc).
 *   sigma is explicit: the reference has three different defaults.
 * sfb_sparsify_sfgrass: SfGrassSparsifier::sparsify_graph (src_legacy/sparsification.rs:32-113),
 *   in place; ratio clamped to [0.1, 1]; skipped when mean degree < 10. */
typedef struct {
    double p;
    double sigma;
    int32_t sparsify;
} sfb_adj_params;

int32_t sfb_adjacency_build(sfb_ctx* ctx, const sfb_knn* g, const sfb_adj_params* params,
                            sfb_adj** out, int32_t* sparsified);
int32_t sfb_sparsify_sfgrass(sfb_ctx* ctx, sfb_adj* adj, double ratio, int32_t* applied);
int32_t sfb_adj_shape(const sfb_adj* adj, uint64_t* rows, uint32_t* k);
int32_t sfb_adj_copy(sfb_ctx* ctx, const sfb_adj* adj, uint32_t* idx, double* w, uint32_t* cnt);
int32_t sfb_adj_from_host(sfb_ctx* ctx, const uint32_t* idx, const double* w, const uint32_t* cnt,
                          uint64_t rows, uint32_t k, sfb_adj** out);
void sfb_adj_free(sfb_adj* adj);

/* ---- symmetrise + Laplacian ----------------------------------------------------------------
 * _symmetrise_adjancency + _build_sparse_laplacian + to_csr (src_legacy/laplacian.rs:297-419,161):
 *   undirected edge set {(i,j,w),(j,i,w)}, duplicates -> max, rows sorted by column,
 *   L_ii = sum_j w_ij (ascending j; stored even when 0), L_ij = -w_ij.
 * normalised != 0: L_sym = I - D^-1/2 W D^-1/2 (surfface-core/src/laplacian.rs:333-372,209-219):
 *   L_ii = 1 if d_i > weight_threshold, L_ij = -w/sqrt(d_i d_j), entries |v| <= 1e-9 dropped. */
typedef struct {
    int32_t normalised;
    double weight_threshold;
} sfb_lap_params;

int32_t sfb_laplacian_build(sfb_ctx* ctx, const sfb_adj* adj, const sfb_lap_params* params, sfb_csr** out);
/* Row-owned assembly of a sharded build ("each GPU builds the CSR rows it owns"): rows
 * [row_begin, row_end) of the same Laplacian, from the all-gathered lists.  The handle holds row_end - row_begin rows,
 * indptr starting at 0, GLOBAL column indices; sfb_csr_shape / sfb_csr_copy work on it, the square-matrix consumers
 * (lambda, SpMV) refuse it.  Unnormalised form only (L_sym needs the degree of every neighbour). */
int32_t sfb_laplacian_build_rows(sfb_ctx* ctx, const sfb_adj* adj, const sfb_lap_params* params, uint64_t row_begin, uint64_t row_end,
                                 sfb_csr** out);
int32_t sfb_csr_shape(const sfb_csr* L, uint64_t* rows, uint64_t* nnz);
int32_t sfb_csr_copy(sfb_ctx* ctx, const sfb_csr* L, uint64_t* indptr, uint32_t* indices, double* data);
int32_t sfb_csr_from_host(sfb_ctx* ctx, uint64_t rows, const uint64_t* indptr, const uint32_t* indices,
                          const double* data, sfb_csr** out);
void sfb_csr_free(sfb_csr* L);
/* GraphLaplacian::multiply_vector / rayleigh_quotient (src_legacy/graph.rs:464-501,422-461). */
int32_t sfb_spmv(sfb_ctx* ctx, const sfb_csr* L, const double* x, double* y);
int32_t sfb_rayleigh_quotient(sfb_ctx* ctx, const sfb_csr* L, const double* x, double* out);

/* ---- per-item lambda -----------------------------------------------------------------------
 * L is M x M (M = feature nodes), X is R x M: one lambda per row of X, one HBM pass over X.
 *   LEGACY_TAUMODE  TauMode::compute_taumode_lambdas_parallel (src_legacy/taumode.rs:117-318,326-408)
 *   ENERGY_NODE     node_energy_and_dispersion (src_legacy/energymaps.rs:923-1045): lambda = E,
 *                   dispersion over the upper triangle returned separately
 *   CORE_F32SEM     compute_lambdas_gpu (surfface-core/src/spectral/mod.rs:69-181), f32 semantics
 * tau: TauMode::select_tau on the ITEM vector (src_legacy/taumode.rs:29-70).
 * normalise_minmax != 0: ArrowSpace::normalise_lambdas (src_legacy/core.rs:1341-1355);
 * stats = {min, max, range} of the raw lambdas. */
typedef enum { SFB_LAMBDA_LEGACY_TAUMODE = 0, SFB_LAMBDA_ENERGY_NODE = 1, SFB_LAMBDA_CORE_F32SEM = 2 } sfb_lambda_variant;
typedef enum { SFB_TAU_FIXED = 0, SFB_TAU_MEDIAN = 1, SFB_TAU_MEAN = 2, SFB_TAU_PERCENTILE = 3 } sfb_tau_mode;

typedef struct {
    int32_t variant;
    int32_t tau_mode;
    double tau_value; /* Fixed(t) or Percentile(p) */
    int32_t normalise_minmax;
} sfb_lambda_params;

/* out_lambda: R (host); out_dispersion: R or NULL; stats: 3 or NULL */
int32_t sfb_lambda(sfb_ctx* ctx, const sfb_csr* L, const sfb_mat* x, const sfb_lambda_params* params,
                   double* out_lambda, double* out_dispersion, double* stats);
/* Items that are JL-projected before the Rayleigh quotient (compute_synthetic_lambda with a projection,
 * src_legacy/taumode.rs:261-318): tau (select_tau on the item, :174-175) and the zero-vector test (:268-274) are taken
 * from the UNPROJECTED row of x_original, energy and dispersion from the row of x_projected (= sfb_project_rows of it);
 * L is reduced_dim x reduced_dim.  LEGACY_TAUMODE only. */
int32_t sfb_lambda_projected(sfb_ctx* ctx, const sfb_csr* L, const sfb_mat* x_original, const sfb_mat* x_projected,
                             const sfb_lambda_params* params, double* out_lambda, double* out_dispersion, double* stats);
/* compute_tau_mode_gpu (surfface-core/src/spectral/bridge.rs:27-32) = compute_lambdas_gpu (spectral/mod.rs:158-181) widened
 * to f64: data is N x F f32 on the host, L the F x F Laplacian of Stage C; lambda = Rayleigh + Dirichlet in f32 semantics
 * (SFB_LAMBDA_CORE_F32SEM), not normalised.  out_lambdas: n_items doubles. */
int32_t sfb_compute_tau_mode_lambdas(sfb_ctx* ctx, const sfb_csr* L, const float* data, uint64_t n_items, uint32_t n_features,
                                     double* out_lambdas);
/* compute_tau (surfface-core/src/taumode.rs:37-65): ONE f32 tau resolved from the lambda DISTRIBUTION (host array): finite
 * entries only, TAU_FLOOR = 1e-9; Fixed(t) / Mean (f32 left fold) / Median = sorted[len / 2] / Percentile(p) =
 * sorted[round((len - 1) * clamp(p, 0, 1))].  The selection runs on the device (radix select on the f32 keys). */
int32_t sfb_compute_tau(sfb_ctx* ctx, const float* lambdas, uint64_t n, int32_t tau_mode, float tau_value, float* out_tau);
/* diffusion step of diffuse_and_split_subcentroids (src_legacy/energymaps.rs:520-546):
 * X <- X - eta * X L^T, `steps` times, in place on the device matrix. */
int32_t sfb_diffuse(sfb_ctx* ctx, const sfb_csr* L, sfb_mat* x, double eta, uint32_t steps);

/* Energy pipeline, item -> sub-centroid mapping (src_legacy/energymaps.rs:1246-1342): nearest sub-centroid in
 * lambda space (first strict minimum of |lambda_item - lambda_s|), ties within epsilon (1e-11 in the reference)
 * broken by the strictly largest cosine in ascending s.  item_lambdas: N, sub_lambdas: S (host).  out_idx: N;
 * out_lambda (the chosen sub-centroid's lambda) and out_norm (|item|): N or NULL. */
int32_t sfb_map_items_to_subcentroids(sfb_ctx* ctx, const sfb_mat* items, const double* item_lambdas,
                                      const sfb_mat* sub_centroids, const double* sub_lambdas, double epsilon,
                                      uint32_t* out_idx, double* out_lambda, double* out_norm);

/* ---- JL projection of the items ahead of lambda ---------------------------------------------
 * sfb_project_rows = project_matrix / ImplicitProjection::project (src_legacy/reduction.rs:175-242), the step
 *   compute_synthetic_lambda applies to an unprojected item before the Rayleigh quotient (taumode.rs:277-297).
 *   The reference re-draws its ChaCha8 StandardNormal stream for every item; here the caller draws it ONCE
 *   (the Rust wrapper does, with the same rand crates, matternet-rs_b200/rust/) and hands the samples in:
 *     SFB_PROJECT_LEGACY    samples[i * r + j] (original_dim x reduced_dim, the reference's draw order);
 *                           out_j = fold over i of  acc + (x_i * s_ij) * scale,  scale = 1/sqrt(r), f64
 *     SFB_PROJECT_CORE_F32  samples[j * f + i] (reduced x original: surfface-core/src/clustering.rs:84-109 draws
 *                           reduced-major); out_j = (fold over i of acc + x_i * s_ji) * scale in f32, carried in f64
 *   Every output is a left fold in the reference's order: bit-exact.  out: x.rows x reduced_dim, on the device.
 * sfb_compute_jl_dimension = compute_jl_dimension (reduction.rs:117-171; core != 0: clustering.rs:113-123). */
typedef enum { SFB_PROJECT_LEGACY = 0, SFB_PROJECT_CORE_F32 = 1 } sfb_project_order;
int32_t sfb_project_rows(sfb_ctx* ctx, const sfb_mat* x, const double* samples, uint32_t reduced_dim, int32_t order,
                         sfb_mat** out);
int32_t sfb_compute_jl_dimension(uint64_t n_points, uint64_t original_dim, double epsilon, int32_t core, uint64_t* out);

/* ---- SortedLambdas ---------------------------------------------------------------------------
 * sfb_sorted_lambdas_build = SortedLambdas::build_from + to_vec (src_legacy/sorted_index.rs:22-57), the index
 *   ArrowSpace::build_lambdas_sorted fills after the lambdas (core.rs:937-940): lambdas ascending in OrderedFloat's
 *   order (-0 == +0, NaN last), equal lambdas ordered by the DECIMAL STRING of the item index (zadd sorts buckets
 *   by idx.to_string()), each bucket reporting the key first inserted.  out_std_dev = std_deviation(lambdas)
 *   (laplacian.rs:421-448: sequential f64 sum, then f32).  lambdas, out_lambda: n doubles, out_idx: n (host). */
int32_t sfb_sorted_lambdas_build(sfb_ctx* ctx, const double* lambdas, uint64_t n, double* out_lambda, uint32_t* out_idx,
                                 double* out_std_dev);

/* ---- reference-shaped one-shot entry points (host buffers in, host buffers out) ------------
 * sfb_build_laplacian_matrix = build_laplacian_matrix(transposed, &GraphParams, ..)
 *   (src_legacy/laplacian.rs:122-180): `items` is the already-transposed matrix, nodes = rows.
 *   normalise must be 0 (the reference's StandardScaler pre-scaling is out of scope).
 * sfb_compute_taumode_lambdas = TauMode::compute_taumode_lambdas_parallel + update_lambdas
 *   (src_legacy/taumode.rs:117-214, core.rs:1427-1443): items R x M, lambdas min-max normalised. */
typedef struct {
    double eps;
    uint32_t k;    /* carried for parity with GraphParams; unused, as in the reference */
    uint32_t topk; /* neighbours kept per node */
    double p;
    double sigma;  /* explicit (reference default sigma.unwrap_or(1.0), laplacian.rs:256) */
    int32_t normalise;
    int32_t sparsity_check; /* != 0: SFB_EINVAL if sparsity > 0.95 (graph.rs:232-240) */
} sfb_graph_params;

int32_t sfb_build_laplacian_matrix(sfb_ctx* ctx, const double* items, uint64_t nodes, uint32_t dims,
                                   const sfb_graph_params* params, int32_t screen, sfb_csr** out);
int32_t sfb_compute_taumode_lambdas(sfb_ctx* ctx, const sfb_csr* L, const double* items, uint64_t n_items,
                                    uint32_t n_features, int32_t tau_mode, double tau_value,
                                    double* out_lambdas);

/* ---- successor Stage C: Bhattacharyya feature graph (f32 semantics) ---------------------------
 * sfb_bc_adjacency_build = compute_bhattacharyya_weights (surfface-core/src/laplacian.rs:254-298) with
 *   bhattacharyya_coefficient (surfface-core/src/distance.rs:260-290): means / variances are the centroid
 *   state [C, F] row-major on the HOST (the reference pulls them to the CPU, laplacian.rs:157-158); nodes are
 *   the F features; per node the k.min(F-1) largest BC > weight_threshold, order (BC desc, j asc).
 *   Follow with sfb_laplacian_build {normalised, weight_threshold}: max-symmetrisation and L_sym / L as in
 *   build_laplacian_flat (laplacian.rs:312-394).  Values are f32-exact numbers carried in f64.
 * sfb_laplacian_stage_execute = LaplacianStage::execute (laplacian.rs:135-219) in one call. */
typedef struct {
    uint32_t k_neighbors;       /* default 15   (LaplacianConfig, laplacian.rs:49-77) */
    float variance_regularizer; /* default 1e-6 */
    int32_t normalize;          /* default 1    */
    float weight_threshold;     /* default 1e-9 */
} sfb_laplacian_config;

int32_t sfb_bc_adjacency_build(sfb_ctx* ctx, const float* means, const float* variances, uint32_t n_centroids,
                               uint32_t n_features, uint32_t k, float variance_regularizer, float weight_threshold,
                               sfb_adj** out);
/* degrees: n_features floats or NULL */
int32_t sfb_laplacian_stage_execute(sfb_ctx* ctx, const float* means, const float* variances, uint32_t n_centroids,
                                    uint32_t n_features, const sfb_laplacian_config* cfg, sfb_csr** out, float* degrees);

/* ---- diagnostics ----------------------------------------------------------------------------
 * Raw tensor-core accumulators of one 128 x 256 tile of the screen (query rows row0.., corpus rows
 * from col0 rounded down to a multiple of 256) and the 16-bit operands used, as f32.  Lets the tests
 * check the tcgen05 path and the accumulation-error model of the certification in isolation.
 * out_tile: 128*256; q_rows: 128*kpad or NULL; q_cols: 256*kpad or NULL (kpad = cols rounded up to 64). */
int32_t sfb_debug_screen_tile(sfb_ctx* ctx, const sfb_mat* x, int32_t metric, int32_t screen, uint64_t row0,
                              uint64_t col0, float* out_tile, float* q_rows, float* q_cols, uint32_t* kpad_out,
                              double* scale_out);

/* ---- stage timings of the last calls (device time, ms) ------------------------------------- */
typedef struct {
    double ms_h2d, ms_knn, ms_adjacency, ms_laplacian, ms_lambda, ms_d2h;
    uint64_t kernel_launches; /* kernels of this library launched since ctx creation */
    double ms_lambda_kernel;  /* the per-item lambda kernel alone (inside ms_lambda)         */
    double ms_diffuse;        /* sfb_diffuse                                                  */
    double ms_comm;           /* NCCL exchanges (all-gathers, all-reduces), incl. waiting for the slowest rank */
} sfb_stage_times;
int32_t sfb_timings(const sfb_ctx* ctx, sfb_stage_times* out);
/* device stopwatch on the context's stream (CUDA events): start, run any calls, stop -> ms */
int32_t sfb_timer_start(sfb_ctx* ctx);
int32_t sfb_timer_stop(sfb_ctx* ctx, double* ms);
int32_t sfb_timings_reset(sfb_ctx* ctx);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) -------------------------------------
 * Query rows are sharded across ranks (sfb_knn_params.q_begin/q_end); the exchange steps are the
 * all-gather of the kNN lists before symmetrisation and the min/max + all-gather of lambda.
 * id: 128 bytes from sfb_comm_unique_id on rank 0, distributed by the host. */
int32_t sfb_comm_unique_id(uint8_t id[128]);
int32_t sfb_comm_init(sfb_ctx* ctx, const uint8_t id[128], int32_t rank, int32_t world);
/* every rank uploads only its own row shard (same ceil split as below); the full matrix is assembled on every GPU
 * by one all-gather over NVLink instead of eight uploads of the whole matrix over PCIe */
int32_t sfb_mat_allgather_rows(sfb_ctx* ctx, const sfb_mat* shard, uint64_t total_rows, sfb_mat** out);
/* sfb_knn_build_columns as a COLLECTIVE: every rank passes the same matrix; when it has the feature-graph shape
 * (few columns, many rows) the exact f64 pair sums are split across the ranks and all-reduced (each sum is
 * produced by exactly one rank, so the result has the same bits as the single-GPU build). */
int32_t sfb_knn_build_columns_sharded(sfb_ctx* ctx, const sfb_mat* x, const sfb_knn_params* params, sfb_knn** out);
/* sfb_knn_build as a COLLECTIVE: every rank passes the same (replicated) matrix and its own ceil-split query shard in
 * params->q_begin / q_end.  Same result as sfb_knn_build; the operand preparation of the tensor-core screen (row norms,
 * 16-bit conversion, rounding residuals: one pass over the f64 matrix) is done by each rank for its own rows only and
 * all-gathered over NVLink. */
int32_t sfb_knn_build_sharded(sfb_ctx* ctx, const sfb_mat* rows, const sfb_knn_params* params, sfb_knn** out);
/* gathers equal row shards of every rank into a full-M kNN handle (all ranks get a copy) */
int32_t sfb_knn_allgather(sfb_ctx* ctx, const sfb_knn* shard, uint64_t total_rows, sfb_knn** out);
/* lambda over this rank's rows of X, min/max all-reduced, normalised, all-gathered into out (total_rows) */
int32_t sfb_lambda_allgather(sfb_ctx* ctx, const sfb_csr* L, const sfb_mat* x_shard, uint64_t row0,
                             uint64_t total_rows, const sfb_lambda_params* params, double* out_lambda,
                             double* stats);
int32_t sfb_comm_barrier(sfb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SURFFACE_B200_H */
