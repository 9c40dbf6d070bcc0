// surfface_b200.hpp -- C++17 host-side mirror of the reference's operator interface for the graph-wiring path, over
// the C ABI of surfface_b200.h (header-only; link libsurfface_b200.so).
//
// The reference is Rust and its seams are functions; this header keeps their names, argument
// meaning and error behaviour so that host code and tests read like the reference's own:
//   GraphParams, GraphLaplacian, build_laplacian_matrix          src_legacy/graph.rs:94-136, src_legacy/laplacian.rs:122-180
//   GraphFactory::build_laplacian_matrix_from_k_cluster           src_legacy/graph.rs:193-255
//   GraphLaplacian::{multiply_vector, rayleigh_quotient, degrees} src_legacy/graph.rs:353-373,422-501
//   TauMode, TauMode::compute_taumode_lambdas_parallel            src_legacy/taumode.rs:16-23,117-214 (+ core.rs:1427-1443)
//   SfGrassSparsifier::sparsify_graph                             src_legacy/sparsification.rs:14-113
//   LaplacianConfig, LaplacianOutput, LaplacianStage::execute     surfface-core/src/laplacian.rs:49-219
//   compute_tau_mode_gpu                                          surfface-core/src/spectral/bridge.rs:27-32
//   CoreTauMode, compute_tau                                      surfface-core/src/taumode.rs:12-65
//   compute_jl_dimension, ImplicitProjection, project_matrix      src_legacy/reduction.rs:117-248
//   SortedLambdas::{build_from, to_vec, range_bylambda}           src_legacy/sorted_index.rs:8-79
// The reference panics (assert! / panic!) on bad input; the mirror throws surfface_b200::Error carrying the status
// and the library's message.  Everything computes on the GPU; there is no CPU fallback.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "surfface_b200.h"

namespace surfface_b200 {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& msg) : std::runtime_error("surfface_b200: status " + std::to_string(st) + ": " + msg), status(st) {}
};

// backend::get_device / SurffaceDevice::Cuda (surfface-core/src/backend.rs:18-70): one context per host thread
class Context {
public:
    explicit Context(int device = 0) {
        int st = sfb_ctx_create(device, &h_);
        if (st != SFB_OK) throw Error(st, "sfb_ctx_create failed (no CUDA device, or not sm_100)");
    }
    ~Context() { sfb_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    sfb_ctx* get() const { return h_; }
    void check(int st) const { if (st != SFB_OK) throw Error(st, sfb_last_error(h_)); }
    static Context& thread_default() { thread_local Context c(0); return c; }
private:
    sfb_ctx* h_ = nullptr;
};

// sprs::CsMat<f64> as three vectors (indices ascending per row, diagonal always stored)
struct CsrMatrix {
    size_t rows = 0;
    std::vector<uint64_t> indptr;
    std::vector<uint32_t> indices;
    std::vector<double> data;
    size_t nnz() const { return indices.size(); }
    std::pair<size_t, size_t> shape() const { return {rows, rows}; }
    const double* get(size_t i, size_t j) const {
        for (uint64_t e = indptr[i]; e < indptr[i + 1]; ++e) if (indices[e] == j) return &data[e];
        return nullptr;
    }
};

// src_legacy/graph.rs:94-102.  `k` is carried but unused, as in the reference (only topk drives the neighbour count).
struct GraphParams {
    double eps = 1e-3;
    size_t k = 6;
    size_t topk = 3;
    double p = 2.0;
    std::optional<double> sigma;   // None -> 1.0 (`params.sigma.unwrap_or(1.0)`, laplacian.rs:256)
    bool normalise = false;
    bool sparsity_check = false;
    bool operator==(const GraphParams& o) const {
        return eps == o.eps && k == o.k && topk == o.topk && p == o.p && sigma == o.sigma && normalise == o.normalise && sparsity_check == o.sparsity_check;
    }
};

// The lambda-graph half of the reference's builder (surfface-pipeline/src/builder.rs): defaults :105-111,
// with_lambda_graph :629-657, define_result_k :785-793 (run first thing by every build, :839,1096).
struct LambdaGraphBuilder {
    double lambda_eps = 1e-3;
    size_t lambda_k = 6, lambda_topk = 3;
    double lambda_p = 2.0;
    std::optional<double> lambda_sigma;
    bool normalise = false, sparsity_check = false;
    LambdaGraphBuilder& with_lambda_graph(double eps, size_t k, size_t topk, double p, std::optional<double> sigma_override = std::nullopt) {
        lambda_eps = eps; lambda_k = k; lambda_topk = topk; lambda_p = p; lambda_sigma = sigma_override;
        return *this;
    }
    void define_result_k() { if (lambda_k <= 5) lambda_topk = 3; else if (lambda_k < 10) lambda_topk = 4; }
    GraphParams graph_params() { define_result_k(); return GraphParams{lambda_eps, lambda_k, lambda_topk, lambda_p, lambda_sigma, normalise, sparsity_check}; }
};

// src_legacy/graph.rs:127-136
struct GraphLaplacian {
    std::vector<double> init_data;   // the matrix the graph was built from (row-major, one row per graph node)
    size_t init_rows = 0, init_cols = 0;
    CsrMatrix matrix;
    size_t nnodes = 0;
    GraphParams graph_params;
    bool energy = false;
    std::shared_ptr<sfb_csr> device;   // the same Laplacian resident in HBM

    // graph.rs:464-501
    std::vector<double> multiply_vector(const std::vector<double>& x) const {
        if (x.size() != matrix.rows) throw Error(SFB_EINVAL, "Vector length must match number of nodes");
        std::vector<double> y(x.size());
        Context& c = Context::thread_default();
        c.check(sfb_spmv(c.get(), device.get(), x.data(), y.data()));
        return y;
    }
    // graph.rs:422-461
    double rayleigh_quotient(const std::vector<double>& x) const {
        if (x.size() != matrix.rows) throw Error(SFB_EINVAL, "Vector length must match number of nodes");
        double r = 0.0;
        Context& c = Context::thread_default();
        c.check(sfb_rayleigh_quotient(c.get(), device.get(), x.data(), &r));
        return r;
    }
    // graph.rs:353-373: the diagonal
    std::vector<double> degrees() const {
        std::vector<double> d(matrix.rows, 0.0);
        for (size_t r = 0; r < matrix.rows; ++r) if (const double* v = matrix.get(r, r)) d[r] = *v;
        return d;
    }
    // graph.rs:626-632
    static double sparsity(const CsrMatrix& m) { return 1.0 - (double)m.nnz() / ((double)m.rows * (double)m.rows); }
};

namespace detail {
inline CsrMatrix fetch(Context& c, const sfb_csr* l) {
    CsrMatrix m;
    uint64_t rows = 0, nnz = 0;
    c.check(sfb_csr_shape(l, &rows, &nnz));
    m.rows = rows; m.indptr.resize(rows + 1); m.indices.resize(nnz ? nnz : 1); m.data.resize(nnz ? nnz : 1);
    c.check(sfb_csr_copy(c.get(), l, m.indptr.data(), m.indices.data(), m.data.data()));
    c.check(sfb_synchronize(c.get()));
    m.indices.resize(nnz); m.data.resize(nnz);
    return m;
}
}  // namespace detail

// src_legacy/laplacian.rs:122-180: `transposed` is row-major rows x cols, one ROW per graph node
// (`let (d, n) = transposed.shape()`: nnodes defaults to the column count n)
inline GraphLaplacian build_laplacian_matrix(const std::vector<double>& transposed, size_t rows, size_t cols, const GraphParams& params,
                                             std::optional<size_t> n_items = std::nullopt, bool energy = false) {
    if (transposed.size() != rows * cols) throw Error(SFB_EINVAL, "matrix size does not match its shape");
    Context& c = Context::thread_default();
    sfb_graph_params gp{params.eps, (uint32_t)params.k, (uint32_t)params.topk, params.p, params.sigma.value_or(1.0),
                        params.normalise ? 1 : 0, params.sparsity_check ? 1 : 0};
    sfb_csr* l = nullptr;
    c.check(sfb_build_laplacian_matrix(c.get(), transposed.data(), rows, (uint32_t)cols, &gp, SFB_SCREEN_AUTO, &l));
    GraphLaplacian gl;
    gl.device = std::shared_ptr<sfb_csr>(l, [](sfb_csr* p) { sfb_csr_free(p); });
    gl.matrix = detail::fetch(c, l);
    gl.init_data = transposed; gl.init_rows = rows; gl.init_cols = cols;
    gl.nnodes = n_items.value_or(cols);
    gl.graph_params = params; gl.energy = energy;
    return gl;
}

struct GraphFactory {
    // src_legacy/graph.rs:193-255: `clustered` is items x features; the graph is over the FEATURES (its columns)
    static GraphLaplacian build_laplacian_matrix_from_k_cluster(const std::vector<double>& clustered, size_t rows, size_t cols, double eps, size_t k,
                                                                size_t topk, double p, std::optional<double> sigma_override, bool normalise,
                                                                bool sparsity_check, size_t n_items) {
        if (rows > n_items) throw Error(SFB_EINVAL, "clustered.shape().0 <= n_items");   // graph.rs:212
        std::vector<double> t(clustered.size());
        for (size_t r = 0; r < rows; ++r) for (size_t cc = 0; cc < cols; ++cc) t[cc * rows + r] = clustered[r * cols + cc];
        GraphParams params{eps, k, topk, p, sigma_override, normalise, sparsity_check};
        return build_laplacian_matrix(t, cols, rows, params, n_items, false);
    }
};

// src_legacy/taumode.rs:16-23
struct TauMode {
    enum Kind { KFixed = SFB_TAU_FIXED, KMedian = SFB_TAU_MEDIAN, KMean = SFB_TAU_MEAN, KPercentile = SFB_TAU_PERCENTILE } kind = KMedian;
    double value = 0.0;
    static TauMode Fixed(double t) { return {KFixed, t}; }
    static TauMode Median() { return {KMedian, 0.0}; }
    static TauMode Mean() { return {KMean, 0.0}; }
    static TauMode Percentile(double p) { return {KPercentile, p}; }
    // taumode.rs:117-214 + ArrowSpace::update_lambdas (core.rs:1427-1443): per-item synthetic lambda against the F x F
    // Laplacian, min-max normalised.  items: n_items x n_features row-major.
    static std::vector<double> compute_taumode_lambdas_parallel(const std::vector<double>& items, size_t n_items, size_t n_features,
                                                                const GraphLaplacian& gl, TauMode mode) {
        if (items.size() != n_items * n_features) throw Error(SFB_EINVAL, "items size does not match its shape");
        std::vector<double> out(n_items);
        Context& c = Context::thread_default();
        c.check(sfb_compute_taumode_lambdas(c.get(), gl.device.get(), items.data(), n_items, (uint32_t)n_features, (int)mode.kind, mode.value, out.data()));
        return out;
    }
};

// src_legacy/sparsification.rs:14-113
struct SfGrassSparsifier {
    double target_ratio = 0.5;
    SfGrassSparsifier& with_target_ratio(double r) { target_ratio = r < 0.1 ? 0.1 : (r > 1.0 ? 1.0 : r); return *this; }
    using Row = std::vector<std::pair<size_t, double>>;
    std::vector<Row> sparsify_graph(const std::vector<Row>& adj_rows, size_t n_nodes) const {
        size_t k = 1;
        for (const Row& r : adj_rows) if (r.size() > k) k = r.size();
        std::vector<uint32_t> idx(n_nodes * k, SFB_IDX_NONE), cnt(n_nodes, 0);
        std::vector<double> w(n_nodes * k, 0.0);
        for (size_t i = 0; i < adj_rows.size() && i < n_nodes; ++i) {
            cnt[i] = (uint32_t)adj_rows[i].size();
            for (size_t t = 0; t < adj_rows[i].size(); ++t) { idx[i * k + t] = (uint32_t)adj_rows[i][t].first; w[i * k + t] = adj_rows[i][t].second; }
        }
        Context& c = Context::thread_default();
        sfb_adj* a = nullptr;
        c.check(sfb_adj_from_host(c.get(), idx.data(), w.data(), cnt.data(), n_nodes, (uint32_t)k, &a));
        std::unique_ptr<sfb_adj, void (*)(sfb_adj*)> guard(a, sfb_adj_free);
        int32_t applied = 0;
        c.check(sfb_sparsify_sfgrass(c.get(), a, target_ratio, &applied));
        c.check(sfb_adj_copy(c.get(), a, idx.data(), w.data(), cnt.data()));
        c.check(sfb_synchronize(c.get()));
        std::vector<Row> out(n_nodes);
        for (size_t i = 0; i < n_nodes; ++i) for (uint32_t t = 0; t < cnt[i]; ++t) out[i].push_back({idx[i * k + t], w[i * k + t]});
        return out;
    }
};

// surfface-core/src/laplacian.rs:49-99
struct LaplacianConfig {
    size_t k_neighbors = 15;
    float variance_regularizer = 1e-6f;
    bool normalize = true;
    float weight_threshold = 1e-9f;
};
struct LaplacianOutput {
    CsrMatrix matrix;   // values are f32-exact numbers
    size_t n_features = 0, nnz = 0;
    std::vector<float> degrees;
    float sparsity = 0.f;
    std::shared_ptr<sfb_csr> device;
};
// surfface-core/src/laplacian.rs:101-219: means / variances are the centroid state [C, F] row-major
class LaplacianStage {
public:
    explicit LaplacianStage(LaplacianConfig cfg = {}) : config(cfg) {}
    static LaplacianStage with_defaults() { return LaplacianStage(LaplacianConfig{}); }
    LaplacianOutput execute(const std::vector<float>& means, const std::vector<float>& variances, size_t c, size_t f) const {
        if (means.size() != c * f || variances.size() != c * f) throw Error(SFB_EINVAL, "means / variances must be [C, F]");
        Context& ctx = Context::thread_default();
        sfb_laplacian_config lc{(uint32_t)config.k_neighbors, config.variance_regularizer, config.normalize ? 1 : 0, config.weight_threshold};
        LaplacianOutput out;
        out.degrees.resize(f);
        sfb_csr* l = nullptr;
        ctx.check(sfb_laplacian_stage_execute(ctx.get(), means.data(), variances.data(), (uint32_t)c, (uint32_t)f, &lc, &l, out.degrees.data()));
        out.device = std::shared_ptr<sfb_csr>(l, [](sfb_csr* p) { sfb_csr_free(p); });
        out.matrix = detail::fetch(ctx, l);
        out.n_features = f; out.nnz = out.matrix.nnz();
        out.sparsity = 1.0f - (float)out.nnz / (float)(f * f);
        return out;
    }
    LaplacianConfig config;
};

// Stage D seam of the successor: compute_tau_mode_gpu(&LaplacianOutput, data: &[f32], n_items, n_features) -> Vec<f64>
// (surfface-core/src/spectral/bridge.rs:27-32): Rayleigh + Dirichlet per item in f32 semantics, widened to f64, not
// normalised.  `data` crosses PCIe as f32 and is widened on the device.
inline std::vector<double> compute_tau_mode_gpu(const LaplacianOutput& laplacian, const std::vector<float>& data, size_t n_items, size_t n_features) {
    if (data.size() != n_items * n_features) throw Error(SFB_EINVAL, "data must be n_items x n_features");
    Context& ctx = Context::thread_default();
    std::vector<double> out(n_items);
    ctx.check(sfb_compute_tau_mode_lambdas(ctx.get(), laplacian.device.get(), data.data(), n_items, (uint32_t)n_features, out.data()));
    return out;
}

// TauMode of the successor (surfface-core/src/taumode.rs:12-23): tau is resolved from the lambda DISTRIBUTION.
struct CoreTauMode {
    int kind = SFB_TAU_MEDIAN;
    float value = 0.f;
    static CoreTauMode Median() { return {SFB_TAU_MEDIAN, 0.f}; }
    static CoreTauMode Mean() { return {SFB_TAU_MEAN, 0.f}; }
    static CoreTauMode Fixed(float t) { return {SFB_TAU_FIXED, t}; }
    static CoreTauMode Percentile(float p) { return {SFB_TAU_PERCENTILE, p}; }
};
// compute_tau(lambdas: &[f32], mode: &TauMode) -> f32 (surfface-core/src/taumode.rs:37-65)
inline float compute_tau(const std::vector<float>& lambdas, const CoreTauMode& mode = CoreTauMode::Median()) {
    Context& ctx = Context::thread_default();
    float out = 0.f;
    ctx.check(sfb_compute_tau(ctx.get(), lambdas.empty() ? nullptr : lambdas.data(), lambdas.size(), mode.kind, mode.value, &out));
    return out;
}

// ---- JL projection ahead of lambda (src_legacy/reduction.rs) -----------------------------------------------------
inline size_t compute_jl_dimension(size_t n_points, size_t original_dim, double epsilon) {
    uint64_t out = 0;
    int st = sfb_compute_jl_dimension(n_points, original_dim, epsilon, 0, &out);
    if (st != SFB_OK) throw Error(st, "sfb_compute_jl_dimension");
    return (size_t)out;
}

// reduction.rs:202-248.  The reference keeps the seed and re-draws its ChaCha8 StandardNormal stream for every item;
// the device wants the draws once: `samples[i * reduced_dim + j]` in the reference's draw order (the Rust wrapper
// fills them with the reference's own rand crates; a C++ host supplies its own N(0,1) matrix).
struct ImplicitProjection {
    size_t original_dim = 0, reduced_dim = 0;
    std::vector<double> samples;
    ImplicitProjection(size_t original, size_t reduced, std::vector<double> draws) : original_dim(original), reduced_dim(reduced), samples(std::move(draws)) {
        if (samples.size() != original_dim * reduced_dim) throw Error(SFB_EINVAL, "samples must be original_dim x reduced_dim");
    }
    size_t get_reduced_dim() const { return reduced_dim; }
    std::vector<double> project(const std::vector<double>& query) const;
};

// reduction.rs:175-200: row-major n_rows x original_dim in, n_rows x reduced_dim out
inline std::vector<double> project_matrix(const std::vector<double>& data, size_t n_rows, const ImplicitProjection& projection) {
    if (data.size() != n_rows * projection.original_dim) throw Error(SFB_EINVAL, "data must be n_rows x original_dim");
    Context& c = Context::thread_default();
    sfb_mat *x = nullptr, *y = nullptr;
    c.check(sfb_mat_from_host(c.get(), data.data(), n_rows, (uint32_t)projection.original_dim, &x));
    int st = sfb_project_rows(c.get(), x, projection.samples.data(), (uint32_t)projection.reduced_dim, SFB_PROJECT_LEGACY, &y);
    sfb_mat_free(x);
    c.check(st);
    std::vector<double> out(n_rows * projection.reduced_dim);
    st = sfb_mat_copy_rows(c.get(), y, 0, n_rows, out.data());
    sfb_mat_free(y);
    c.check(st);
    return out;
}
inline std::vector<double> ImplicitProjection::project(const std::vector<double>& query) const {
    return project_matrix(std::vector<double>(query.begin(), query.begin() + (std::ptrdiff_t)original_dim), 1, *this);   // `.take(original_dim)`, :233
}

// TauMode::compute_taumode_lambdas_parallel for an ArrowSpace that carries a projection (taumode.rs:117-214 with
// aspace.projection_matrix = Some(..)): items n_items x original_dim, gl over the reduced_dim projected features; tau and
// the zero-vector test come from the unprojected item (:174-175,268-274), energy and dispersion from the projected one.
inline std::vector<double> compute_taumode_lambdas_projected(const std::vector<double>& items, size_t n_items, const ImplicitProjection& projection,
                                                             const GraphLaplacian& gl, TauMode mode) {
    if (items.size() != n_items * projection.original_dim) throw Error(SFB_EINVAL, "items must be n_items x original_dim");
    Context& c = Context::thread_default();
    sfb_mat *x = nullptr, *y = nullptr;
    c.check(sfb_mat_from_host(c.get(), items.data(), n_items, (uint32_t)projection.original_dim, &x));
    int st = sfb_project_rows(c.get(), x, projection.samples.data(), (uint32_t)projection.reduced_dim, SFB_PROJECT_LEGACY, &y);
    if (st != SFB_OK) { sfb_mat_free(x); c.check(st); }
    std::vector<double> out(n_items);
    sfb_lambda_params prm{SFB_LAMBDA_LEGACY_TAUMODE, (int)mode.kind, mode.value, 1};   // update_lambdas normalises (core.rs:1427-1443)
    st = sfb_lambda_projected(c.get(), gl.device.get(), x, y, &prm, out.data(), nullptr, nullptr);
    if (st == SFB_OK) st = sfb_synchronize(c.get());
    sfb_mat_free(x);
    sfb_mat_free(y);
    c.check(st);
    return out;
}

// ---- SortedLambdas (src_legacy/sorted_index.rs:8-79) ----------------------------------------------------------------
class SortedLambdas {
public:
    // build_from (:32-46): throws where the reference panics (no lambdas)
    void build_from(const std::vector<double>& lambdas) {
        Context& c = Context::thread_default();
        lambdas_.assign(lambdas.size(), 0.0);
        indices_.assign(lambdas.size(), 0u);
        c.check(sfb_sorted_lambdas_build(c.get(), lambdas.data(), lambdas.size(), lambdas_.data(), indices_.data(), &std_dev_));
    }
    // to_vec (:48-57): (lambda, idx) in map order
    std::vector<std::pair<double, size_t>> to_vec() const {
        std::vector<std::pair<double, size_t>> out(lambdas_.size());
        for (size_t i = 0; i < out.size(); ++i) out[i] = {lambdas_[i], indices_[i]};
        return out;
    }
    // range_bylambda (:60-79): the first k items with lambda in [q - band, q + band], band = std_dev / 2^p
    std::vector<std::pair<size_t, double>> range_bylambda(double lambda_q, size_t k, double p) const {
        const double band = std_dev_ / std::pow(2.0, p), lo = lambda_q - band, hi = lambda_q + band;
        std::vector<std::pair<size_t, double>> out;
        size_t a = 0, b = lambdas_.size();
        while (a < b) { size_t m = (a + b) / 2; if (lambdas_[m] < lo) a = m + 1; else b = m; }
        for (size_t i = a; i < lambdas_.size() && lambdas_[i] <= hi && out.size() < k; ++i) out.emplace_back(indices_[i], lambdas_[i]);
        return out;
    }
    double std_dev() const { return std_dev_; }
private:
    std::vector<double> lambdas_;
    std::vector<uint32_t> indices_;
    double std_dev_ = 0.0;
};

}  // namespace surfface_b200
