//! Safe wrappers with the reference's signatures (shipped as source; not compiled in this image: no Rust toolchain).
//!
//! | here                                                        | reference                                                                     |
//! |-------------------------------------------------------------|-------------------------------------------------------------------------------|
//! | `build_laplacian_matrix(transposed, &params, n_items, energy) -> GraphLaplacian` | `src_legacy/laplacian.rs:122-128`                          |
//! | `TauMode::compute_taumode_lambdas_parallel(&mut aspace, &gl, taumode)`           | `src_legacy/taumode.rs:117-121` (+ `update_lambdas`, `core.rs:1427-1443`) |
//! | `SfGrassSparsifier::sparsify_graph(&adj_rows, n_nodes)`                         | `src_legacy/sparsification.rs:32-36`                        |
//! | `diffuse(&gl, &mut x, rows, eta, steps)`, `map_items_to_subcentroids(..)`       | `src_legacy/energymaps.rs:520-546,1246-1342`                |
//! | `LaplacianStage::execute(means, variances, c, f) -> LaplacianOutput`            | `surfface-core/src/laplacian.rs:135` (state pulled to the host as the reference does, `:157-158`) |
//! | `compute_tau_mode_gpu(&LaplacianOutput, data: &[f32], n_items, n_features) -> Vec<f64>` | `surfface-core/src/spectral/bridge.rs:27-32`        |
//! | `compute_tau(lambdas: &[f32], mode: &CoreTauMode) -> f32`                       | `surfface-core/src/taumode.rs:37-65`                        |
//! | `Comm`, `build_laplacian_rows_sharded`, `compute_taumode_lambdas_sharded`       | none (the reference is single-process): one process per GPU |
//!
//! Two small traits stand where the reference's own types plug in, so that this crate does not depend on the
//! reference crate (the dependency points the other way): `DenseLike` (implemented by smartcore's `DenseMatrix<f64>`
//! behind the `smartcore` feature, and by `RowMajor`) and `LambdaSpace` (implemented by `ArrowSpace` in five lines,
//! see INTEGRATION.md).
//!
//! The reference panics on bad input (`assert!`, `panic!`): so do these, with the library's message.
use std::cell::RefCell;
use std::ffi::CStr;

use sprs::CsMat;
use surfface_b200_sys as sys;

// ---- the two seams where the reference's own types plug in --------------------------------------------------------
/// Anything that can hand over a dense f64 matrix row-major.  `smartcore::linalg::basic::matrix::DenseMatrix<f64>` is
/// column-major inside (`DenseMatrix::from_iterator(.., rows, cols, 1)`): its impl walks `get((r, c))`.
pub trait DenseLike {
    fn shape(&self) -> (usize, usize);
    fn to_row_major(&self) -> Vec<f64>;
}

/// A plain row-major matrix (tests, callers without smartcore).
#[derive(Clone, Debug, PartialEq)]
pub struct RowMajor {
    pub data: Vec<f64>,
    pub rows: usize,
    pub cols: usize,
}
impl DenseLike for RowMajor {
    fn shape(&self) -> (usize, usize) {
        (self.rows, self.cols)
    }
    fn to_row_major(&self) -> Vec<f64> {
        self.data.clone()
    }
}

#[cfg(feature = "smartcore")]
impl DenseLike for smartcore::linalg::basic::matrix::DenseMatrix<f64> {
    fn shape(&self) -> (usize, usize) {
        smartcore::linalg::basic::arrays::Array::shape(self)
    }
    fn to_row_major(&self) -> Vec<f64> {
        use smartcore::linalg::basic::arrays::Array;
        let (r, c) = Array::shape(self);
        let mut out = Vec::with_capacity(r * c);
        for i in 0..r {
            for j in 0..c {
                out.push(*self.get((i, j)));
            }
        }
        out
    }
}

/// What `compute_taumode_lambdas_parallel` needs from an `ArrowSpace` (`src_legacy/core.rs`): the items, the optional
/// projection, and `update_lambdas` (which min-max normalises, `core.rs:1427-1443` -- here the device already did).
pub trait LambdaSpace {
    /// `(items row-major, n_items, n_features)`; `aspace.data` as the reference reads it (`taumode.rs:129-136`)
    fn items_row_major(&self) -> (Vec<f64>, usize, usize);
    /// `aspace.projection_matrix` (`taumode.rs:277-297`)
    fn projection(&self) -> Option<&ImplicitProjection> {
        None
    }
    /// store NORMALISED lambdas (the wrapper passes the min-max normalised vector: `store_lambdas_normalised`)
    fn store_lambdas_normalised(&mut self, lambdas: Vec<f64>);
}

// ---- configuration types (field for field the reference's) -----------------------------------------------------------
/// `GraphParams` (`src_legacy/graph.rs:94-102`)
#[derive(Clone, Debug, PartialEq)]
pub struct GraphParams {
    pub eps: f64,
    pub k: usize,
    pub topk: usize,
    pub p: f64,
    pub sigma: Option<f64>,
    pub normalise: bool,
    pub sparsity_check: bool,
}

/// `TauMode` (`src_legacy/taumode.rs:16-23`)
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum TauMode {
    Fixed(f64),
    Median,
    Mean,
    Percentile(f64),
}
impl TauMode {
    fn code(self) -> (i32, f64) {
        match self {
            TauMode::Fixed(t) => (0, t),
            TauMode::Median => (1, 0.0),
            TauMode::Mean => (2, 0.0),
            TauMode::Percentile(p) => (3, p),
        }
    }
}

// ---- context ---------------------------------------------------------------------------------------------------------
struct Ctx(*mut sys::sfb_ctx);
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::sfb_ctx_destroy(self.0) }
    }
}

thread_local! {
    // one context per host thread (the library's threading convention); SURFFACE_B200_DEVICE picks the GPU (one process per GPU)
    static CTX: RefCell<Option<Ctx>> = RefCell::new(None);
}

fn with_ctx<R>(f: impl FnOnce(*mut sys::sfb_ctx) -> R) -> R {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            let dev = std::env::var("SURFFACE_B200_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
            let mut h = std::ptr::null_mut();
            let st = unsafe { sys::sfb_ctx_create(dev, &mut h) };
            assert!(st == sys::SFB_OK, "surfface_b200: no sm_100 device {dev} (status {st}); there is no CPU fallback");
            *c = Some(Ctx(h));
        }
        f(c.as_ref().unwrap().0)
    })
}

fn check(ctx: *mut sys::sfb_ctx, st: i32) {
    if st != sys::SFB_OK {
        let msg = unsafe { CStr::from_ptr(sys::sfb_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("surfface_b200: status {st}: {msg}");
    }
}

fn fetch(ctx: *mut sys::sfb_ctx, l: *mut sys::sfb_csr, cols: usize) -> CsMat<f64> {
    let (mut rows, mut nnz) = (0u64, 0u64);
    check(ctx, unsafe { sys::sfb_csr_shape(l, &mut rows, &mut nnz) });
    let n = rows as usize;
    let mut indptr = vec![0u64; n + 1];
    let mut indices = vec![0u32; (nnz as usize).max(1)];
    let mut data = vec![0f64; (nnz as usize).max(1)];
    check(ctx, unsafe { sys::sfb_csr_copy(ctx, l, indptr.as_mut_ptr(), indices.as_mut_ptr(), data.as_mut_ptr()) });
    check(ctx, unsafe { sys::sfb_synchronize(ctx) });
    indices.truncate(nnz as usize);
    data.truncate(nnz as usize);
    CsMat::new((n, cols), indptr.iter().map(|&p| p as usize).collect(), indices.iter().map(|&j| j as usize).collect(), data)
}

// ---- GraphLaplacian + build_laplacian_matrix ------------------------------------------------------------------------
/// `GraphLaplacian` (`src_legacy/graph.rs:127-136`): same public fields; the device handle rides along so that the lambda
/// pass does not upload the matrix again.
pub struct GraphLaplacian {
    pub init_data: RowMajor,
    pub matrix: CsMat<f64>,
    pub nnodes: usize,
    pub graph_params: GraphParams,
    pub energy: bool,
    handle: *mut sys::sfb_csr,
}
impl Drop for GraphLaplacian {
    fn drop(&mut self) {
        unsafe { sys::sfb_csr_free(self.handle) }
    }
}

/// `build_laplacian_matrix(transposed, &params, n_items, energy)` (`src_legacy/laplacian.rs:122-180`): one row of
/// `transposed` per graph node.  `nnodes = n_items.unwrap_or(cols)` (`:129,166-169`).
pub fn build_laplacian_matrix<M: DenseLike>(transposed: M, params: &GraphParams, n_items: Option<usize>, energy: bool) -> GraphLaplacian {
    let (rows, cols) = transposed.shape();
    assert!(rows >= 2 && cols >= 2, "items should be at least of shape (2,2): ({cols},{rows})"); // laplacian.rs:130-135
    let data = transposed.to_row_major();
    with_ctx(|ctx| {
        let gp = sys::sfb_graph_params {
            eps: params.eps,
            k: params.k as u32,
            topk: params.topk as u32,
            p: params.p,
            sigma: params.sigma.unwrap_or(1.0), // laplacian.rs:256
            normalise: params.normalise as i32,
            sparsity_check: params.sparsity_check as i32,
        };
        let mut l = std::ptr::null_mut();
        check(ctx, unsafe { sys::sfb_build_laplacian_matrix(ctx, data.as_ptr(), rows as u64, cols as u32, &gp, 0, &mut l) });
        GraphLaplacian {
            matrix: fetch(ctx, l, rows),
            init_data: RowMajor { data, rows, cols },
            nnodes: n_items.unwrap_or(cols),
            graph_params: params.clone(),
            energy,
            handle: l,
        }
    })
}

impl TauMode {
    /// `TauMode::compute_taumode_lambdas_parallel(&mut aspace, &gl, taumode)` (`src_legacy/taumode.rs:117-214`): per-item
    /// synthetic lambda against `gl`, written back through `update_lambdas` (min-max normalised, `core.rs:1427-1443`).
    /// With a projection on the space, tau and the zero-vector test come from the unprojected item (`:174-175,268-297`).
    pub fn compute_taumode_lambdas_parallel<S: LambdaSpace>(aspace: &mut S, gl: &GraphLaplacian, taumode: TauMode) {
        let (items, n_items, n_features) = aspace.items_row_major();
        assert_eq!(items.len(), n_items * n_features);
        let (mode, value) = taumode.code();
        let mut out = vec![0f64; n_items];
        match aspace.projection().cloned() {
            None => with_ctx(|ctx| {
                check(ctx, unsafe {
                    sys::sfb_compute_taumode_lambdas(ctx, gl.handle, items.as_ptr(), n_items as u64, n_features as u32, mode, value, out.as_mut_ptr())
                })
            }),
            Some(projection) => {
                assert_eq!(n_features, projection.original_dim);
                let samples = projection.materialise();
                with_ctx(|ctx| unsafe {
                    let (mut x, mut y) = (std::ptr::null_mut(), std::ptr::null_mut());
                    check(ctx, sys::sfb_mat_from_host(ctx, items.as_ptr(), n_items as u64, n_features as u32, &mut x));
                    let st = sys::sfb_project_rows(ctx, x, samples.as_ptr(), projection.reduced_dim as u32, 0, &mut y);
                    if st != sys::SFB_OK {
                        sys::sfb_mat_free(x);
                    }
                    check(ctx, st);
                    let prm = sys::sfb_lambda_params { variant: 0, tau_mode: mode, tau_value: value, normalise_minmax: 1 };
                    let st = sys::sfb_lambda_projected(ctx, gl.handle, x, y, &prm, out.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut());
                    sys::sfb_mat_free(x);
                    sys::sfb_mat_free(y);
                    check(ctx, st);
                    check(ctx, sys::sfb_synchronize(ctx));
                })
            }
        }
        aspace.store_lambdas_normalised(out);
    }
    /// `TauMode::compute_taumode_lambdas` (`taumode.rs:411-413`)
    pub fn compute_taumode_lambdas<S: LambdaSpace>(aspace: &mut S, gl: &GraphLaplacian, taumode: TauMode) {
        Self::compute_taumode_lambdas_parallel(aspace, gl, taumode)
    }
}

// ---- SfGrassSparsifier (src_legacy/sparsification.rs:14-113) -----------------------------------------------------------
pub struct SfGrassSparsifier {
    target_ratio: f64,
}
impl Default for SfGrassSparsifier {
    fn default() -> Self {
        Self::new()
    }
}
impl SfGrassSparsifier {
    pub fn new() -> Self {
        Self { target_ratio: 0.5 }
    }
    pub fn with_target_ratio(mut self, ratio: f64) -> Self {
        self.target_ratio = ratio.clamp(0.1, 1.0); // sparsification.rs:26-29
        self
    }
    /// `sparsify_graph(&self, adj_rows: &[Vec<(usize, f64)>], n_nodes) -> Vec<Vec<(usize, f64)>>` (`:32-36`)
    pub fn sparsify_graph(&self, adj_rows: &[Vec<(usize, f64)>], n_nodes: usize) -> Vec<Vec<(usize, f64)>> {
        assert_eq!(adj_rows.len(), n_nodes);
        let k = adj_rows.iter().map(|r| r.len()).max().unwrap_or(0).max(1);
        let mut idx = vec![sys::SFB_IDX_NONE; n_nodes * k];
        let mut w = vec![0f64; n_nodes * k];
        let mut cnt = vec![0u32; n_nodes];
        for (i, row) in adj_rows.iter().enumerate() {
            cnt[i] = row.len() as u32;
            for (t, &(j, wv)) in row.iter().enumerate() {
                idx[i * k + t] = j as u32;
                w[i * k + t] = wv;
            }
        }
        with_ctx(|ctx| unsafe {
            let mut a = std::ptr::null_mut();
            check(ctx, sys::sfb_adj_from_host(ctx, idx.as_ptr(), w.as_ptr(), cnt.as_ptr(), n_nodes as u64, k as u32, &mut a));
            let mut applied = 0i32;
            let st = sys::sfb_sparsify_sfgrass(ctx, a, self.target_ratio, &mut applied);
            if st == sys::SFB_OK {
                check(ctx, sys::sfb_adj_copy(ctx, a, idx.as_mut_ptr(), w.as_mut_ptr(), cnt.as_mut_ptr()));
            }
            sys::sfb_adj_free(a);
            check(ctx, st);
        });
        (0..n_nodes).map(|i| (0..cnt[i] as usize).map(|t| (idx[i * k + t] as usize, w[i * k + t])).collect()).collect()
    }
}

// ---- energy pipeline steps (src_legacy/energymaps.rs) ------------------------------------------------------------------
/// The diffusion of `diffuse_and_split_subcentroids` (`energymaps.rs:520-546`): `x` (`rows x F` row-major) `<- x - eta * x L^T`,
/// `steps` times, in place.
pub fn diffuse(gl: &GraphLaplacian, x: &mut [f64], rows: usize, eta: f64, steps: usize) {
    let f = gl.matrix.rows();
    assert_eq!(x.len(), rows * f);
    with_ctx(|ctx| unsafe {
        let mut m = std::ptr::null_mut();
        check(ctx, sys::sfb_mat_from_host(ctx, x.as_ptr(), rows as u64, f as u32, &mut m));
        let st = sys::sfb_diffuse(ctx, gl.handle, m, eta, steps as u32);
        if st == sys::SFB_OK {
            check(ctx, sys::sfb_mat_copy_rows(ctx, m, 0, rows as u64, x.as_mut_ptr()));
        }
        sys::sfb_mat_free(m);
        check(ctx, st);
    })
}

/// Item -> sub-centroid mapping (`energymaps.rs:1246-1342`): `(index, lambda of the chosen sub-centroid, |item|)` per item.
pub fn map_items_to_subcentroids(
    items: &[f64], n_items: usize, item_lambdas: &[f64], sub_centroids: &[f64], n_sub: usize, sub_lambdas: &[f64], n_features: usize, epsilon: f64,
) -> (Vec<usize>, Vec<f64>, Vec<f64>) {
    assert!(items.len() == n_items * n_features && sub_centroids.len() == n_sub * n_features);
    assert!(item_lambdas.len() == n_items && sub_lambdas.len() == n_sub);
    let (mut idx, mut lam, mut norm) = (vec![0u32; n_items], vec![0f64; n_items], vec![0f64; n_items]);
    with_ctx(|ctx| unsafe {
        let (mut xi, mut xs) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(ctx, sys::sfb_mat_from_host(ctx, items.as_ptr(), n_items as u64, n_features as u32, &mut xi));
        let st = sys::sfb_mat_from_host(ctx, sub_centroids.as_ptr(), n_sub as u64, n_features as u32, &mut xs);
        if st != sys::SFB_OK {
            sys::sfb_mat_free(xi);
        }
        check(ctx, st);
        let st = sys::sfb_map_items_to_subcentroids(ctx, xi, item_lambdas.as_ptr(), xs, sub_lambdas.as_ptr(), epsilon, idx.as_mut_ptr(), lam.as_mut_ptr(), norm.as_mut_ptr());
        sys::sfb_mat_free(xi);
        sys::sfb_mat_free(xs);
        check(ctx, st);
    });
    (idx.into_iter().map(|i| i as usize).collect(), lam, norm)
}

// ---- successor crate: Stage C and Stage D ----------------------------------------------------------------------------
/// `LaplacianConfig` / `LaplacianStage` / `LaplacianOutput` (`surfface-core/src/laplacian.rs:49-219`).
#[derive(Clone, Debug)]
pub struct LaplacianConfig {
    pub k_neighbors: usize,
    pub variance_regularizer: f32,
    pub normalize: bool,
    pub weight_threshold: f32,
}
impl Default for LaplacianConfig {
    fn default() -> Self {
        Self { k_neighbors: 15, variance_regularizer: 1e-6, normalize: true, weight_threshold: 1e-9 }
    }
}
pub struct LaplacianOutput {
    pub matrix: CsMat<f32>,
    pub n_features: usize,
    pub nnz: usize,
    pub degrees: Vec<f32>,
    pub sparsity: f32,
}
pub struct LaplacianStage {
    pub config: LaplacianConfig,
}
impl LaplacianStage {
    pub fn new(config: LaplacianConfig) -> Self {
        Self { config }
    }
    pub fn with_defaults() -> Self {
        Self::new(LaplacianConfig::default())
    }
    /// `execute(&self, state: &CentroidState<B>) -> LaplacianOutput` (`laplacian.rs:135`): the reference pulls the state to
    /// the host first (`:157-158`: `state.means.to_data().to_vec()`), which is what the caller passes here: `[C, F]` row-major.
    pub fn execute(&self, means: &[f32], variances: &[f32], c: usize, f: usize) -> LaplacianOutput {
        assert!(means.len() == c * f && variances.len() == c * f);
        with_ctx(|ctx| {
            let cfg = sys::sfb_laplacian_config {
                k_neighbors: self.config.k_neighbors as u32,
                variance_regularizer: self.config.variance_regularizer,
                normalize: self.config.normalize as i32,
                weight_threshold: self.config.weight_threshold,
            };
            let mut degrees = vec![0f32; f];
            let mut l = std::ptr::null_mut();
            check(ctx, unsafe { sys::sfb_laplacian_stage_execute(ctx, means.as_ptr(), variances.as_ptr(), c as u32, f as u32, &cfg, &mut l, degrees.as_mut_ptr()) });
            let m64 = fetch(ctx, l, f);
            unsafe { sys::sfb_csr_free(l) };
            let nnz = m64.nnz();
            let matrix = m64.map(|&v| v as f32); // values are f32-exact
            LaplacianOutput { matrix, n_features: f, nnz, degrees, sparsity: 1.0 - nnz as f32 / (f * f) as f32 }
        })
    }
}

/// `compute_tau_mode_gpu(&LaplacianOutput, data: &[f32], n_items, n_features) -> Vec<f64>` (`spectral/bridge.rs:27-32`):
/// Rayleigh + Dirichlet per item in f32 semantics, widened to f64, not normalised.  `data` crosses PCIe as f32.
pub fn compute_tau_mode_gpu(laplacian: &LaplacianOutput, data: &[f32], n_items: usize, n_features: usize) -> Vec<f64> {
    assert_eq!(data.len(), n_items * n_features);
    assert_eq!(laplacian.n_features, n_features);
    let m = &laplacian.matrix;
    let indptr: Vec<u64> = m.indptr().raw_storage().iter().map(|&p| p as u64).collect();
    let indices: Vec<u32> = m.indices().iter().map(|&j| j as u32).collect();
    let values: Vec<f64> = m.data().iter().map(|&v| v as f64).collect();
    let mut out = vec![0f64; n_items];
    with_ctx(|ctx| unsafe {
        let mut l = std::ptr::null_mut();
        check(ctx, sys::sfb_csr_from_host(ctx, n_features as u64, indptr.as_ptr(), indices.as_ptr(), values.as_ptr(), &mut l));
        let st = sys::sfb_compute_tau_mode_lambdas(ctx, l, data.as_ptr(), n_items as u64, n_features as u32, out.as_mut_ptr());
        sys::sfb_csr_free(l);
        check(ctx, st);
    });
    out
}

/// `TauMode` of the successor (`surfface-core/src/taumode.rs:12-23`): tau is resolved from the lambda DISTRIBUTION.
#[derive(Clone, Debug, Default, PartialEq)]
pub enum CoreTauMode {
    #[default]
    Median,
    Mean,
    Fixed(f32),
    Percentile(f32),
}
/// `compute_tau(lambdas: &[f32], mode: &TauMode) -> f32` (`surfface-core/src/taumode.rs:37-65`)
pub fn compute_tau(lambdas: &[f32], mode: &CoreTauMode) -> f32 {
    let (m, v) = match mode {
        CoreTauMode::Fixed(t) => (0, *t),
        CoreTauMode::Median => (1, 0.0),
        CoreTauMode::Mean => (2, 0.0),
        CoreTauMode::Percentile(p) => (3, *p),
    };
    let mut out = 0f32;
    with_ctx(|ctx| check(ctx, unsafe { sys::sfb_compute_tau(ctx, lambdas.as_ptr(), lambdas.len() as u64, m, v, &mut out) }));
    out
}

// ---- JL projection ahead of lambda (src_legacy/reduction.rs:175-248) --------------------------------------------
use rand::SeedableRng;
use rand_chacha::ChaCha8Rng;
use rand_distr::{Distribution, StandardNormal};

/// `compute_jl_dimension` (`reduction.rs:117-171`)
pub fn compute_jl_dimension(n_points: usize, original_dim: usize, epsilon: f64) -> usize {
    let mut out = 0u64;
    let st = unsafe { sys::sfb_compute_jl_dimension(n_points as u64, original_dim as u64, epsilon, 0, &mut out) };
    assert!(st == sys::SFB_OK);
    out as usize
}

/// `ImplicitProjection` (`reduction.rs:202-248`): still seed-only.
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct ImplicitProjection {
    pub original_dim: usize,
    pub reduced_dim: usize,
    pub seed: u64,
}
impl ImplicitProjection {
    pub fn new(original_dim: usize, reduced_dim: usize, seed: Option<u64>) -> Self {
        Self { original_dim, reduced_dim, seed: seed.unwrap_or_else(rand::random) }
    }
    /// The draws `project` makes, once: `samples[i * r + j]` is the StandardNormal drawn for (original i, reduced j)
    /// -- the reference restarts `ChaCha8Rng::seed_from_u64(seed)` for every item and consumes it in exactly this
    /// order (`reduction.rs:228-239`), so every item sees the same F x r matrix.
    pub fn materialise(&self) -> Vec<f64> {
        let mut rng = ChaCha8Rng::seed_from_u64(self.seed);
        (0..self.original_dim * self.reduced_dim).map(|_| StandardNormal.sample(&mut rng)).collect()
    }
    pub fn get_reduced_dim(&self) -> usize {
        self.reduced_dim
    }
}

/// `project_matrix(data, projection)` (`reduction.rs:175-200`)
pub fn project_matrix<M: DenseLike>(data: &M, projection: &ImplicitProjection) -> RowMajor {
    let (n_rows, f) = data.shape();
    assert_eq!(f, projection.original_dim);
    let flat = data.to_row_major();
    let samples = projection.materialise();
    let r = projection.reduced_dim;
    let mut out = vec![0f64; n_rows * r];
    with_ctx(|ctx| unsafe {
        let (mut x, mut y) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(ctx, sys::sfb_mat_from_host(ctx, flat.as_ptr(), n_rows as u64, f as u32, &mut x));
        let st = sys::sfb_project_rows(ctx, x, samples.as_ptr(), r as u32, 0, &mut y);
        sys::sfb_mat_free(x);
        check(ctx, st);
        let st = sys::sfb_mat_copy_rows(ctx, y, 0, n_rows as u64, out.as_mut_ptr());
        sys::sfb_mat_free(y);
        check(ctx, st);
    });
    RowMajor { data: out, rows: n_rows, cols: r }
}

// ---- SortedLambdas (src_legacy/sorted_index.rs) ------------------------------------------------------------------
use ordered_float::OrderedFloat;
use std::collections::BTreeMap;

/// `SortedLambdas::build_from` (`sorted_index.rs:32-46`): the device returns the items already in map order, so the
/// BTreeMap is bulk-built from a sorted run (O(N)) instead of N `zadd` calls that each re-sort a bucket of strings.
pub fn build_sorted_lambdas(lambdas: &[f64]) -> (BTreeMap<OrderedFloat<f64>, Vec<(usize, String)>>, f64) {
    let n = lambdas.len();
    let (mut sorted, mut idx, mut std_dev) = (vec![0f64; n], vec![0u32; n], 0f64);
    with_ctx(|ctx| check(ctx, unsafe { sys::sfb_sorted_lambdas_build(ctx, lambdas.as_ptr(), n as u64, sorted.as_mut_ptr(), idx.as_mut_ptr(), &mut std_dev) }));
    let mut runs: Vec<(OrderedFloat<f64>, Vec<(usize, String)>)> = Vec::new();
    for (lam, i) in sorted.into_iter().zip(idx) {
        match runs.last_mut() {
            Some((k, bucket)) if *k == OrderedFloat(lam) => bucket.push((i as usize, i.to_string())),
            _ => runs.push((OrderedFloat(lam), vec![(i as usize, i.to_string())])),
        }
    }
    (runs.into_iter().collect(), std_dev)
}

// ---- multi-GPU: one process per GPU (no counterpart in the single-process reference) ---------------------------------
/// The NCCL communicator of this process' context.  Rank 0 calls `Comm::unique_id()` and ships the 128 bytes to the other
/// ranks by whatever the host has (MPI, a file, a socket); every rank then calls `Comm::init`.
pub struct Comm {
    pub rank: usize,
    pub world: usize,
}
impl Comm {
    pub fn unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        assert!(unsafe { sys::sfb_comm_unique_id(id.as_mut_ptr()) } == sys::SFB_OK, "libnccl could not be loaded");
        id
    }
    pub fn init(id: &[u8; 128], rank: usize, world: usize) -> Comm {
        with_ctx(|ctx| check(ctx, unsafe { sys::sfb_comm_init(ctx, id.as_ptr(), rank as i32, world as i32) }));
        Comm { rank, world }
    }
    /// rows `[lo, hi)` of `n` owned by this rank (ceil split, the convention of every sharded entry point)
    pub fn shard(&self, n: usize) -> (usize, usize) {
        let s = (n + self.world - 1) / self.world;
        let lo = (self.rank * s).min(n);
        (lo, (lo + s).min(n))
    }
}

/// The item graph of a sharded build: this rank uploads its rows of `items` (`n x d` row-major, `my_rows` = rows
/// `comm.shard(n)`), the corpus is replicated over NVLink, the kNN of the shard runs on the tensor cores, the lists are
/// all-gathered, and the rank assembles the CSR rows it owns (`sfb_laplacian_build_rows`).  Returns those rows
/// (`hi - lo` x `n`, global column indices).
pub fn build_laplacian_rows_sharded(comm: &Comm, my_rows: &[f64], n: usize, d: usize, metric: i32, k: usize, p: f64, sigma: f64) -> CsMat<f64> {
    let (lo, hi) = comm.shard(n);
    assert_eq!(my_rows.len(), (hi - lo) * d);
    with_ctx(|ctx| unsafe {
        let (mut xs, mut x, mut g, mut ga, mut a, mut l) =
            (std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
        check(ctx, sys::sfb_mat_from_host(ctx, my_rows.as_ptr(), (hi - lo) as u64, d as u32, &mut xs));
        let st = sys::sfb_mat_allgather_rows(ctx, xs, n as u64, &mut x);
        sys::sfb_mat_free(xs);
        check(ctx, st);
        let kp = sys::sfb_knn_params { metric, k: k as u32, eps: f64::INFINITY, screen: 0, k_prime: 0, q_begin: lo as u64, q_end: hi as u64, allow_fallback: 1 };
        let st = sys::sfb_knn_build(ctx, x, &kp, &mut g);
        sys::sfb_mat_free(x);
        check(ctx, st);
        let st = sys::sfb_knn_allgather(ctx, g, n as u64, &mut ga);
        sys::sfb_knn_free(g);
        check(ctx, st);
        let ap = sys::sfb_adj_params { p, sigma, sparsify: -1 };
        let st = sys::sfb_adjacency_build(ctx, ga, &ap, &mut a, std::ptr::null_mut());
        sys::sfb_knn_free(ga);
        check(ctx, st);
        let lp = sys::sfb_lap_params { normalised: 0, weight_threshold: 0.0 };
        let st = sys::sfb_laplacian_build_rows(ctx, a, &lp, lo as u64, hi as u64, &mut l);
        sys::sfb_adj_free(a);
        check(ctx, st);
        let m = fetch(ctx, l, n);
        sys::sfb_csr_free(l);
        m
    })
}

/// Lambdas of a sharded item set: this rank's rows against the (replicated) feature Laplacian, global min / max, all ranks
/// receive the whole normalised vector.
pub fn compute_taumode_lambdas_sharded(comm: &Comm, my_rows: &[f64], n: usize, f: usize, gl: &GraphLaplacian, taumode: TauMode) -> Vec<f64> {
    let (lo, hi) = comm.shard(n);
    assert_eq!(my_rows.len(), (hi - lo) * f);
    let (mode, value) = taumode.code();
    let mut out = vec![0f64; n];
    with_ctx(|ctx| unsafe {
        let mut xs = std::ptr::null_mut();
        check(ctx, sys::sfb_mat_from_host(ctx, my_rows.as_ptr(), (hi - lo) as u64, f as u32, &mut xs));
        let prm = sys::sfb_lambda_params { variant: 0, tau_mode: mode, tau_value: value, normalise_minmax: 1 };
        let st = sys::sfb_lambda_allgather(ctx, gl.handle, xs, lo as u64, n as u64, &prm, out.as_mut_ptr(), std::ptr::null_mut());
        sys::sfb_mat_free(xs);
        check(ctx, st);
    });
    out
}
