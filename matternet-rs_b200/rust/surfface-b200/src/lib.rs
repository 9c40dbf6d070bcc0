//! Safe wrappers with the reference's signatures (shipped as source; not compiled in this image).
//!
//! * `build_laplacian_matrix`            replaces `src_legacy/laplacian.rs:122-180`
//! * `compute_taumode_lambdas`           replaces `TauMode::compute_taumode_lambdas_parallel` (`src_legacy/taumode.rs:117-214`)
//!                                       + `ArrowSpace::update_lambdas` (`src_legacy/core.rs:1427-1443`)
//! * `LaplacianStage::execute`           replaces `surfface-core/src/laplacian.rs:135-219`
//!
//! The reference panics on bad input (`assert!`, `panic!`): so do these, with the library's message.
use std::cell::RefCell;
use std::ffi::CStr;

use sprs::CsMat;
use surfface_b200_sys as sys;

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

struct Ctx(*mut sys::sfb_ctx);
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::sfb_ctx_destroy(self.0) }
    }
}

thread_local! {
    // one context per host thread (the library's threading convention)
    static CTX: RefCell<Option<Ctx>> = RefCell::new(None);
}

fn with_ctx<R>(f: impl FnOnce(*mut sys::sfb_ctx) -> R) -> R {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            let mut h = std::ptr::null_mut();
            let st = unsafe { sys::sfb_ctx_create(0, &mut h) };
            assert!(st == sys::SFB_OK, "surfface_b200: no sm_100 device (status {st}); there is no CPU fallback");
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

/// A Laplacian resident in HBM plus its host copy (`GraphLaplacian.matrix`, `src_legacy/graph.rs:127-136`).
pub struct DeviceLaplacian {
    handle: *mut sys::sfb_csr,
    pub matrix: CsMat<f64>,
}
impl Drop for DeviceLaplacian {
    fn drop(&mut self) {
        unsafe { sys::sfb_csr_free(self.handle) }
    }
}

fn fetch(ctx: *mut sys::sfb_ctx, l: *mut sys::sfb_csr) -> CsMat<f64> {
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
    CsMat::new((n, n), indptr.iter().map(|&p| p as usize).collect(), indices.iter().map(|&j| j as usize).collect(), data)
}

/// `build_laplacian_matrix(transposed, &params, n_items, energy)`: `transposed` is row-major `rows x cols`, one row per
/// graph node.  Returns the device-resident Laplacian (host copy in `.matrix`) and `nnodes`
/// (`n_items.unwrap_or(cols)`, `laplacian.rs:129,166-169`).
pub fn build_laplacian_matrix(transposed: &[f64], rows: usize, cols: usize, params: &GraphParams, n_items: Option<usize>) -> (DeviceLaplacian, usize) {
    assert_eq!(transposed.len(), rows * cols);
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
        check(ctx, unsafe { sys::sfb_build_laplacian_matrix(ctx, transposed.as_ptr(), rows as u64, cols as u32, &gp, 0, &mut l) });
        (DeviceLaplacian { handle: l, matrix: fetch(ctx, l) }, n_items.unwrap_or(cols))
    })
}

/// Per-item taumode lambdas against the F x F Laplacian, min-max normalised (items: `n_items x n_features` row-major).
pub fn compute_taumode_lambdas(items: &[f64], n_items: usize, n_features: usize, gl: &DeviceLaplacian, taumode: TauMode) -> Vec<f64> {
    assert_eq!(items.len(), n_items * n_features);
    let (mode, value) = match taumode {
        TauMode::Fixed(t) => (0, t),
        TauMode::Median => (1, 0.0),
        TauMode::Mean => (2, 0.0),
        TauMode::Percentile(p) => (3, p),
    };
    let mut out = vec![0f64; n_items];
    with_ctx(|ctx| check(ctx, unsafe { sys::sfb_compute_taumode_lambdas(ctx, gl.handle, items.as_ptr(), n_items as u64, n_features as u32, mode, value, out.as_mut_ptr()) }));
    out
}

/// `LaplacianConfig` / `LaplacianStage` / `LaplacianOutput` of the successor crate (`surfface-core/src/laplacian.rs:49-219`).
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
    /// `means` / `variances`: the centroid state `[C, F]` row-major (what `state.means.to_data().to_vec()` yields).
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
            let m64 = fetch(ctx, l);
            unsafe { sys::sfb_csr_free(l) };
            let nnz = m64.nnz();
            let matrix = m64.map(|&v| v as f32); // values are f32-exact
            LaplacianOutput { matrix, n_features: f, nnz, degrees, sparsity: 1.0 - nnz as f32 / (f * f) as f32 }
        })
    }
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

/// `project_matrix(data, projection)` (`reduction.rs:175-200`): row-major `n x original_dim` in, `n x reduced_dim` out.
pub fn project_matrix(data: &[f64], n_rows: usize, projection: &ImplicitProjection) -> Vec<f64> {
    assert_eq!(data.len(), n_rows * projection.original_dim);
    let samples = projection.materialise();
    let r = projection.reduced_dim;
    let mut out = vec![0f64; n_rows * r];
    with_ctx(|ctx| unsafe {
        let (mut x, mut y) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(ctx, sys::sfb_mat_from_host(ctx, data.as_ptr(), n_rows as u64, projection.original_dim as u32, &mut x));
        let st = sys::sfb_project_rows(ctx, x, samples.as_ptr(), r as u32, 0, &mut y);
        sys::sfb_mat_free(x);
        check(ctx, st);
        let st = sys::sfb_mat_copy_rows(ctx, y, 0, n_rows as u64, out.as_mut_ptr());
        sys::sfb_mat_free(y);
        check(ctx, st);
    });
    out
}

/// `TauMode::compute_taumode_lambdas_parallel` for an ArrowSpace that carries a projection (`taumode.rs:117-214` with
/// `aspace.projection_matrix = Some(..)`): items are `n_items x original_dim`, `gl` is `reduced_dim x reduced_dim`;
/// tau and the zero-vector test come from the unprojected item, energy and dispersion from the projected one.
pub fn compute_taumode_lambdas_projected(items: &[f64], n_items: usize, projection: &ImplicitProjection, gl: &DeviceLaplacian, taumode: TauMode) -> Vec<f64> {
    assert_eq!(items.len(), n_items * projection.original_dim);
    let (mode, value) = match taumode {
        TauMode::Fixed(t) => (0, t),
        TauMode::Median => (1, 0.0),
        TauMode::Mean => (2, 0.0),
        TauMode::Percentile(p) => (3, p),
    };
    let samples = projection.materialise();
    let mut out = vec![0f64; n_items];
    with_ctx(|ctx| unsafe {
        let (mut x, mut y) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(ctx, sys::sfb_mat_from_host(ctx, items.as_ptr(), n_items as u64, projection.original_dim as u32, &mut x));
        let st = sys::sfb_project_rows(ctx, x, samples.as_ptr(), projection.reduced_dim as u32, 0, &mut y);
        if st != sys::SFB_OK { sys::sfb_mat_free(x); }
        check(ctx, st);
        let prm = sys::sfb_lambda_params { variant: 0, tau_mode: mode, tau_value: value, normalise_minmax: 1 };   // update_lambdas normalises
        let st = sys::sfb_lambda_projected(ctx, gl.handle, x, y, &prm, out.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut());
        sys::sfb_mat_free(x);
        sys::sfb_mat_free(y);
        check(ctx, st);
        check(ctx, sys::sfb_synchronize(ctx));
    });
    out
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
