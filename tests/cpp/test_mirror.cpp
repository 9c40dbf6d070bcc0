// test_mirror.cpp -- the reference's own tests for the graph-wiring path, restated against the C++ mirror
// (include/surfface_b200.hpp) over the C ABI, plus parity with the CPU oracle on seeded inputs.
//   test_basic_laplacian_construction / test_laplacian_mathematical_properties   src_legacy/tests/test_laplacian.rs:34-154
//   test_with_adjacency_output (L = D - A on three points)                        src_legacy/tests/test_laplacian.rs:655-786
//   chain-graph lambda, constant vector => 0, scale invariance                    surfface-core/src/tests/test_spectral.rs:187-251,
//                                                                                 src_legacy/tests/test_taumode.rs:643-682
// Built and run by tests/test_cpp_mirror.py (needs a B200); links libsurfface_b200.so and liboracle.so (the checker).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <algorithm>
#include <tuple>
#include <vector>

#include "surfface_b200.hpp"

using namespace surfface_b200;

extern "C" {   // oracle/oracle.c (test infrastructure)
void orc_knn(const double* x, uint64_t m, uint32_t kd, int metric, uint32_t k, double eps, const uint64_t* query_rows, uint64_t nq,
             uint32_t* idx, double* dist, uint32_t* cnt);
int orc_build_adjacency(const uint32_t* idx, const double* dist, const uint32_t* cnt, uint64_t m, uint32_t k, double p, double sigma,
                        int force, uint32_t* a_idx, double* a_w, uint32_t* a_cnt);
void* orc_laplacian_build(const uint32_t* a_idx, const double* a_w, const uint32_t* a_cnt, uint64_t m, uint32_t k, int normalised, double thr);
uint64_t orc_csr_nnz(void* h);
void orc_csr_copy(void* h, uint64_t* indptr, uint32_t* indices, double* data);
void orc_csr_free(void* h);
void orc_lambda(const uint64_t* indptr, const uint32_t* indices, const double* data, uint64_t f, const double* x, uint64_t n, int variant,
                int tau_mode, double tau_value, double* lam, double* e, double* g);
void orc_normalise_lambdas(double* lam, uint64_t n, double* stats);
uint64_t orc_jl_dimension(uint64_t n_points, uint64_t original_dim, double epsilon);
void orc_project_rows(const double* x, uint64_t n, uint32_t f, const double* samples, uint32_t r, double* out);
int orc_sorted_lambdas(const double* lam, uint64_t n, double* out_lambda, uint32_t* out_idx, double* out_std_dev);
float orc_compute_tau_core(const float* lambdas, uint64_t n, int mode, float value);
}

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++failures; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)

static std::vector<double> transpose(const std::vector<double>& a, size_t rows, size_t cols) {
    std::vector<double> t(a.size());
    for (size_t r = 0; r < rows; ++r) for (size_t c = 0; c < cols; ++c) t[c * rows + r] = a[r * cols + c];
    return t;
}

// create_test_vectors / default_params (test_laplacian.rs:11-32)
static const std::vector<double> kItems = {1.0, 0.0, 0.0, 0.8, 0.6, 0.0, 0.0, 1.0, 0.0, 0.0, 0.8, 0.6, 0.0, 0.0, 1.0};
static GraphParams default_params() { return GraphParams{0.5, 3, 2, 2.0, 0.1, false, true}; }

static void test_basic_laplacian_construction() {
    GraphParams params = default_params();
    GraphLaplacian laplacian = build_laplacian_matrix(transpose(kItems, 5, 3), 3, 5, params, std::nullopt, false);
    CHECK(laplacian.nnodes == 5, "nnodes %zu", laplacian.nnodes);
    CHECK(laplacian.matrix.shape() == std::make_pair((size_t)3, (size_t)3), "shape");
    CHECK(laplacian.graph_params == params, "graph_params");
}

static void check_properties(const GraphLaplacian& laplacian, size_t max_nnz) {
    const CsrMatrix& m = laplacian.matrix;
    for (size_t i = 0; i < m.rows; ++i) {   // row sums zero, symmetric, diag >= 0, off-diag <= 0
        double row_sum = 0.0;
        for (uint64_t e = m.indptr[i]; e < m.indptr[i + 1]; ++e) {
            row_sum += m.data[e];
            size_t j = m.indices[e];
            const double* lji = m.get(j, i);
            CHECK(lji && std::fabs(m.data[e] - *lji) <= 1e-12, "symmetric L[%zu,%zu]", i, j);
            if (i != j) CHECK(m.data[e] <= 1e-12, "off-diagonal L[%zu,%zu] = %g", i, j, m.data[e]);
            if (e > m.indptr[i]) CHECK(m.indices[e] > m.indices[e - 1], "sorted columns in row %zu", i);
        }
        CHECK(std::fabs(row_sum) <= 1e-12, "row %zu sum %g", i, row_sum);
        const double* d = m.get(i, i);
        CHECK(d && *d >= -1e-12, "diagonal %zu", i);
    }
    CHECK(m.nnz() <= max_nnz, "nnz %zu > %zu", m.nnz(), max_nnz);
    CHECK(m.indices.size() == m.nnz() && m.data.size() == m.nnz(), "CSR integrity");
}

static void test_laplacian_mathematical_properties() {
    GraphParams params = default_params();
    GraphLaplacian laplacian = build_laplacian_matrix(transpose(kItems, 5, 3), 3, 5, params);
    check_properties(laplacian, laplacian.nnodes * (params.k + 1));
}

// three points [[1,0],[.9,.1],[0,1]], eps .5, topk 1, p 1, sigma .2: values derivable by hand
static void test_with_adjacency_output() {
    const std::vector<double> items = {1.0, 0.0, 0.9, 0.1, 0.0, 1.0};
    GraphParams params{0.5, 2, 1, 1.0, 0.2, false, false};
    GraphLaplacian gl = build_laplacian_matrix(items, 3, 2, params);   // nodes = the three rows
    const double cos01 = 0.9 / std::sqrt(0.82), d01 = 1.0 - cos01, w01 = 1.0 / (1.0 + d01 / 0.2);
    const CsrMatrix& m = gl.matrix;
    CHECK(m.get(0, 1) && std::fabs(*m.get(0, 1) + w01) <= 1e-12, "L[0,1] = %g, want %g", m.get(0, 1) ? *m.get(0, 1) : NAN, -w01);
    CHECK(m.get(0, 0) && std::fabs(*m.get(0, 0) - w01) <= 1e-12, "L[0,0]");
    CHECK(m.get(2, 2) && *m.get(2, 2) == 0.0 && m.indptr[3] - m.indptr[2] == 1, "node 2 is isolated (distance 1 - 0 > eps... to 0; 0.89 to 1): diagonal stored as 0");
    check_properties(gl, 9);
    // L = D - A: the diagonal is the sum of the row's off-diagonal magnitudes
    for (size_t i = 0; i < 3; ++i) {
        double deg = 0.0;
        for (uint64_t e = m.indptr[i]; e < m.indptr[i + 1]; ++e) if (m.indices[e] != i) deg += -m.data[e];
        CHECK(std::fabs(*m.get(i, i) - deg) <= 1e-10, "diag %zu", i);
    }
}

// chain graph L = [[1,-1,0],[-1,2,-1],[0,-1,1]] arises from three collinear... built directly through the C ABI
static void test_lambda_kats() {
    Context& c = Context::thread_default();
    const uint64_t indptr[4] = {0, 2, 5, 7};
    const uint32_t indices[7] = {0, 1, 0, 1, 2, 1, 2};
    const double data[7] = {1, -1, -1, 2, -1, -1, 1};
    sfb_csr* l = nullptr;
    c.check(sfb_csr_from_host(c.get(), 3, indptr, indices, data, &l));
    GraphLaplacian gl; gl.device = std::shared_ptr<sfb_csr>(l, [](sfb_csr* p) { sfb_csr_free(p); }); gl.matrix.rows = 3;
    CHECK(std::fabs(gl.rayleigh_quotient({1.0, 1.0, 1.0})) <= 1e-15, "R(constant) = 0");
    CHECK(gl.rayleigh_quotient({1.0, 0.0, -1.0}) > 0.5, "R(alternating) > 0");
    std::vector<double> y = gl.multiply_vector({1.0, 2.0, 4.0});
    CHECK(y[0] == -1.0 && y[1] == -1.0 && y[2] == 2.0, "L x");
    // per-item lambdas, raw (sfb_lambda) : constant vector => 0; lambda(x) == lambda(2x) to 1e-10 with tau Fixed
    const std::vector<double> items = {1.0, 1.0, 1.0, 1.0, 0.0, -1.0, 2.0, 0.0, -2.0};
    sfb_mat* xm = nullptr;
    c.check(sfb_mat_from_host(c.get(), items.data(), 3, 3, &xm));
    sfb_lambda_params lp{SFB_LAMBDA_LEGACY_TAUMODE, SFB_TAU_FIXED, 0.5, 0};
    double lam[3];
    c.check(sfb_lambda(c.get(), l, xm, &lp, lam, nullptr, nullptr));
    c.check(sfb_synchronize(c.get()));
    sfb_mat_free(xm);
    CHECK(lam[0] >= 0.0 && lam[0] <= 1e-12, "lambda(constant) = %g", lam[0]);
    CHECK(lam[1] > lam[0] && std::fabs(lam[1] - lam[2]) <= 1e-10, "scale invariance %g vs %g", lam[1], lam[2]);
}

// seeded Gaussian items: feature graph + lambdas against the oracle (indices exact, values 1e-9)
static void test_oracle_parity() {
    const size_t n = 700, f = 24, topk = 3;
    std::mt19937_64 rng(5);
    std::normal_distribution<double> nd;
    std::vector<double> items(n * f);
    for (double& v : items) v = nd(rng);
    GraphLaplacian gl = GraphFactory::build_laplacian_matrix_from_k_cluster(items, n, f, INFINITY, 6, topk, 2.0, 1.0, false, false, n);
    CHECK(gl.matrix.rows == f && gl.nnodes == n, "shape");
    std::vector<double> xt = transpose(items, n, f);
    std::vector<uint32_t> idx(f * topk), cnt(f), a_idx(f * topk), a_cnt(f);
    std::vector<double> dist(f * topk), a_w(f * topk);
    orc_knn(xt.data(), f, (uint32_t)n, 0, topk, INFINITY, nullptr, f, idx.data(), dist.data(), cnt.data());
    orc_build_adjacency(idx.data(), dist.data(), cnt.data(), f, topk, 2.0, 1.0, -1, a_idx.data(), a_w.data(), a_cnt.data());
    void* h = orc_laplacian_build(a_idx.data(), a_w.data(), a_cnt.data(), f, topk, 0, 0.0);
    uint64_t nnz = orc_csr_nnz(h);
    std::vector<uint64_t> ip(f + 1); std::vector<uint32_t> ix(nnz); std::vector<double> dv(nnz);
    orc_csr_copy(h, ip.data(), ix.data(), dv.data()); orc_csr_free(h);
    CHECK(gl.matrix.indptr == ip && gl.matrix.indices == ix, "CSR structure differs from the oracle");
    for (size_t e = 0; e < nnz && e < gl.matrix.data.size(); ++e)
        CHECK(std::fabs(gl.matrix.data[e] - dv[e]) <= 1e-9 * std::fabs(dv[e]) + 1e-300, "L value %zu", e);
    std::vector<double> lam = TauMode::compute_taumode_lambdas_parallel(items, n, f, gl, TauMode::Median());
    std::vector<double> o(n), stats(3);
    orc_lambda(ip.data(), ix.data(), dv.data(), f, items.data(), n, 0, 1 /* median */, 0.0, o.data(), nullptr, nullptr);
    orc_normalise_lambdas(o.data(), n, stats.data());
    for (size_t i = 0; i < n; ++i) CHECK(std::fabs(lam[i] - o[i]) <= 1e-9 * std::fabs(o[i]) + 1e-12, "lambda %zu: %g vs %g", i, lam[i], o[i]);
}

static void test_errors_are_exceptions() {
    bool threw = false;
    try { build_laplacian_matrix({1.0, 2.0, 3.0}, 1, 3, GraphParams{}); } catch (const Error& e) { threw = e.status == SFB_EINVAL; }   // assert!(n >= 2 && d >= 2)
    CHECK(threw, "a 1-row matrix must be refused like the reference's assert!");
    threw = false;
    GraphParams sparse = default_params(); sparse.eps = 1e-9;   // nothing within eps: 3 stored diagonals of 9 -> not > 0.95 sparse... use a bigger graph
    std::vector<double> big(40 * 6);
    for (size_t i = 0; i < big.size(); ++i) big[i] = std::sin(0.37 * (double)i) + 2.0;
    try { build_laplacian_matrix(big, 40, 6, sparse); } catch (const Error& e) { threw = e.status == SFB_EINVAL; }               // "too sparse", graph.rs:232-240
    CHECK(threw, "sparsity_check must refuse an empty graph");
}

static void test_sfgrass_and_stage() {
    using Row = SfGrassSparsifier::Row;
    std::vector<Row> adj(12);
    for (size_t i = 0; i < 12; ++i) for (size_t t = 1; t <= 11; ++t) adj[i].push_back({(i + t) % 12, 1.0 / (double)t});
    std::vector<Row> out = SfGrassSparsifier().with_target_ratio(0.5).sparsify_graph(adj, 12);
    for (const Row& r : out) CHECK(r.size() == 6, "ceil(11 * 0.5) edges kept, got %zu", r.size());
    std::vector<float> means(8 * 20), vars(8 * 20);
    for (size_t i = 0; i < means.size(); ++i) { means[i] = std::sin(0.7f * (float)i); vars[i] = 0.2f + 0.1f * std::cos(0.3f * (float)i); }
    LaplacianOutput lo = LaplacianStage::with_defaults().execute(means, vars, 8, 20);
    CHECK(lo.n_features == 20 && lo.matrix.rows == 20 && lo.degrees.size() == 20, "stage C shape");
    for (size_t i = 0; i < 20; ++i) if (lo.degrees[i] > 1e-9f) CHECK(lo.matrix.get(i, i) && std::fabs(*lo.matrix.get(i, i) - 1.0) <= 1e-6, "L_sym diagonal");
}

// the reference's own reduction tests (src_legacy/tests/test_reduction.rs) + SortedLambdas, against the oracle
static void test_projection_and_sorted_lambdas() {
    CHECK(compute_jl_dimension(100, 16, 0.3) == 16 && compute_jl_dimension(10, 100, 0.3) == 100, "jl: low dims / never expands");   // :193-210
    CHECK(compute_jl_dimension(2, 1000, 0.9) == 32 && compute_jl_dimension(1000, 512, 0.1) == 512, "jl: clamp");                      // :219-243
    CHECK(compute_jl_dimension(1, 100, 0.1) == 32 && compute_jl_dimension(1, 10, 0.1) == 10, "jl: single point");                      // :468-475
    for (auto [n, d, e] : {std::tuple<size_t, size_t, double>{10000, 5000, 0.3}, {100, 2049, 0.3}, {100, 100000, 0.3}})
        CHECK(compute_jl_dimension(n, d, e) == orc_jl_dimension(n, d, e), "jl(%zu, %zu, %g)", n, d, e);
    std::mt19937_64 rng(42);
    std::normal_distribution<double> nd;
    const size_t n = 333, f = 40, r = 10;
    std::vector<double> draws(f * r), data(n * f);
    for (double& v : draws) v = nd(rng);
    for (double& v : data) v = nd(rng);
    ImplicitProjection proj(f, r, draws);
    std::vector<double> zero = proj.project(std::vector<double>(f, 0.0));                                                              // :60-69
    for (double v : zero) CHECK(v == 0.0, "zero vector must project to zeros");
    std::vector<double> ones(f, 1.0), twos(f, 2.0);
    std::vector<double> p1 = proj.project(ones), p2 = proj.project(twos);                                                              // :72-93
    for (size_t j = 0; j < r; ++j) CHECK(p2[j] == 2.0 * p1[j], "linearity at %zu", j);
    std::vector<double> got = project_matrix(data, n, proj), want(n * r);                                                              // :128-148
    orc_project_rows(data.data(), n, (uint32_t)f, draws.data(), (uint32_t)r, want.data());
    CHECK(got == want, "projected matrix differs from the oracle's left folds");
    // lambdas of unprojected items against the Laplacian of the projected features: normalised, the zero item at the minimum
    {
        std::vector<double> pos(data);
        for (double& v : pos) v = std::fabs(v);
        for (size_t j = 0; j < f; ++j) pos[7 * f + j] = 0.0;
        std::vector<double> y = project_matrix(pos, n, proj);
        GraphLaplacian glp = GraphFactory::build_laplacian_matrix_from_k_cluster(y, n, r, INFINITY, 6, 3, 2.0, 1.0, false, false, n);
        std::vector<double> lp = compute_taumode_lambdas_projected(pos, n, proj, glp, TauMode::Median());
        CHECK(lp.size() == n && lp[7] == 0.0, "zero unprojected item must sit at the normalised minimum, got %g", lp[7]);
        double mx = 0.0; for (double v : lp) { CHECK(v >= 0.0 && v <= 1.0, "normalised lambda out of [0,1]"); mx = std::max(mx, v); }
        CHECK(mx == 1.0, "normalised maximum");
    }
    std::vector<double> lam(5000);
    for (size_t i = 0; i < lam.size(); ++i) lam[i] = std::floor(100.0 * std::fabs(std::sin(0.1 * (double)i))) / 100.0;              // heavy ties
    SortedLambdas sl;
    sl.build_from(lam);
    std::vector<double> wl(lam.size()); std::vector<uint32_t> wi(lam.size()); double wsd = 0.0;
    orc_sorted_lambdas(lam.data(), lam.size(), wl.data(), wi.data(), &wsd);
    auto v = sl.to_vec();
    bool same = v.size() == lam.size() && sl.std_dev() == wsd;
    for (size_t i = 0; same && i < v.size(); ++i) same = v[i].first == wl[i] && v[i].second == wi[i];
    CHECK(same, "SortedLambdas order / std_dev differ from the oracle");
    auto hits = sl.range_bylambda(0.5, 7, 2.0);
    CHECK(hits.size() == 7, "range_bylambda returns the first k in the band");
    for (auto& h : hits) CHECK(std::fabs(h.second - 0.5) <= sl.std_dev() / 4.0, "range_bylambda band");
    bool threw = false;
    try { SortedLambdas().build_from({}); } catch (const Error& e) { threw = e.status == SFB_EINVAL; }                                 // sorted_index.rs:36-40 panics
    CHECK(threw, "empty lambdas must be refused");
}

// Stage D seam (surfface-core/src/tests/test_spectral.rs:30-80,165-185: finite lambdas, zero-vector safety) and compute_tau
static void test_stage_d_seam() {
    // Stage C on 3 features / 4 centroids, then compute_tau_mode_gpu on two items (test_spectral.rs:38-78)
    const std::vector<float> means = {1.0f, 0.9f, 0.1f, 0.8f, 1.0f, 0.2f, 0.2f, 0.1f, 1.0f, 0.5f, 0.5f, 0.5f};
    const std::vector<float> vars(12, 0.1f);
    LaplacianConfig cfg; cfg.k_neighbors = 2;
    LaplacianOutput lap = LaplacianStage(cfg).execute(means, vars, 4, 3);
    CHECK(lap.n_features == 3 && lap.nnz > 0, "Stage C output");
    const double* d00 = lap.matrix.get(0, 0);
    CHECK(d00 && std::fabs(*d00 - 1.0) < 1e-5, "L_sym diagonal is 1 for connected nodes");
    const std::vector<float> data = {1.0f, 0.0f, 0.0f, 0.5f, 0.5f, 0.5f};
    std::vector<double> lambdas = compute_tau_mode_gpu(lap, data, 2, 3);
    CHECK(lambdas.size() == 2 && std::isfinite(lambdas[0]) && std::isfinite(lambdas[1]), "lambdas finite");
    // zero vector: finite (test_zero_vector_safety)
    std::vector<double> z = compute_tau_mode_gpu(lap, {0.f, 0.f, 0.f, 1.f, 1.f, 1.f}, 2, 3);
    CHECK(std::isfinite(z[0]) && std::isfinite(z[1]), "zero vector is safe");
    // compute_tau against the oracle's restatement of taumode.rs:37-65
    std::mt19937 rng(5);
    std::vector<float> lam(1001);
    for (auto& v : lam) v = std::generate_canonical<float, 24>(rng);
    lam[17] = NAN; lam[400] = INFINITY;
    const CoreTauMode modes[] = {CoreTauMode::Median(), CoreTauMode::Mean(), CoreTauMode::Fixed(0.25f), CoreTauMode::Fixed(-1.f), CoreTauMode::Percentile(0.9f)};
    for (const auto& m : modes) CHECK(compute_tau(lam, m) == orc_compute_tau_core(lam.data(), lam.size(), m.kind, m.value), "compute_tau mode %d", m.kind);
    CHECK(compute_tau({}, CoreTauMode::Median()) == 1e-9f, "empty distribution -> TAU_FLOOR");
}

int main() {
    test_stage_d_seam();
    test_basic_laplacian_construction();
    test_laplacian_mathematical_properties();
    test_with_adjacency_output();
    test_lambda_kats();
    test_oracle_parity();
    test_errors_are_exceptions();
    test_sfgrass_and_stage();
    test_projection_and_sorted_lambdas();
    std::printf(failures ? "%d FAILED\n" : "all C++ mirror tests passed\n", failures);
    return failures ? 1 : 0;
}
