//! Raw FFI to `libsurfface_b200` -- one declaration per symbol of `include/surfface_b200.h` (generated from the header;
//! `tests/test_abi.py` checks that none is missing).  Shipped as source: this image has no Rust toolchain.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub const SFB_OK: i32 = 0;
pub const SFB_IDX_NONE: u32 = 0xFFFF_FFFF;

#[repr(C)] pub struct sfb_ctx { _p: [u8; 0] }
#[repr(C)] pub struct sfb_mat { _p: [u8; 0] }
#[repr(C)] pub struct sfb_knn { _p: [u8; 0] }
#[repr(C)] pub struct sfb_adj { _p: [u8; 0] }
#[repr(C)] pub struct sfb_csr { _p: [u8; 0] }
#[repr(C)] pub struct sfb_pending { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy)]
pub struct sfb_knn_params { pub metric: i32, pub k: u32, pub eps: f64, pub screen: i32, pub k_prime: u32, pub q_begin: u64, pub q_end: u64, pub allow_fallback: i32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct sfb_knn_stats { pub rows: u64, pub rows_certified: u64, pub rows_fallback: u64, pub k_prime: u32, pub screen_used: i32, pub ms_prepare: f64,
                           pub ms_screen: f64, pub ms_rescore: f64, pub ms_fallback: f64, pub max_margin: f64, pub rows_rescreened: u64, pub ms_rescreen: f64,
                           pub candidates_rescored: u64 }
#[repr(C)] #[derive(Clone, Copy)] pub struct sfb_adj_params { pub p: f64, pub sigma: f64, pub sparsify: i32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct sfb_lap_params { pub normalised: i32, pub weight_threshold: f64 }
#[repr(C)] #[derive(Clone, Copy)] pub struct sfb_lambda_params { pub variant: i32, pub tau_mode: i32, pub tau_value: f64, pub normalise_minmax: i32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sfb_graph_params { pub eps: f64, pub k: u32, pub topk: u32, pub p: f64, pub sigma: f64, pub normalise: i32, pub sparsity_check: i32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sfb_laplacian_config { pub k_neighbors: u32, pub variance_regularizer: f32, pub normalize: i32, pub weight_threshold: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct sfb_stage_times { pub ms_h2d: f64, pub ms_knn: f64, pub ms_adjacency: f64, pub ms_laplacian: f64, pub ms_lambda: f64, pub ms_d2h: f64, pub kernel_launches: u64, pub ms_lambda_kernel: f64, pub ms_diffuse: f64, pub ms_comm: f64 }

extern "C" {
    pub fn sfb_abi_version() -> i32;
    pub fn sfb_ctx_create(device_id: i32, out: *mut *mut sfb_ctx) -> i32;
    pub fn sfb_ctx_destroy(ctx: *mut sfb_ctx);
    pub fn sfb_last_error(ctx: *const sfb_ctx) -> *const c_char;
    pub fn sfb_device_info(ctx: *const sfb_ctx, name: *mut c_char, sm_count: *mut i32, hbm_bytes: *mut u64) -> i32;
    pub fn sfb_synchronize(ctx: *mut sfb_ctx) -> i32;
    pub fn sfb_pinned_alloc(ctx: *mut sfb_ctx, bytes: u64, out: *mut *mut c_void) -> i32;
    pub fn sfb_pinned_free(p: *mut c_void);
    pub fn sfb_mat_from_host(ctx: *mut sfb_ctx, x: *const f64, rows: u64, cols: u32, out: *mut *mut sfb_mat) -> i32;
    pub fn sfb_mat_from_host_f32(ctx: *mut sfb_ctx, x: *const f32, rows: u64, cols: u32, out: *mut *mut sfb_mat) -> i32;
    pub fn sfb_mat_clone(ctx: *mut sfb_ctx, a: *const sfb_mat, out: *mut *mut sfb_mat) -> i32;
    pub fn sfb_mat_generate(ctx: *mut sfb_ctx, kind: i32, seed: u64, rows: u64, cols: u32, n_centres: u32, noise: f64, out: *mut *mut sfb_mat) -> i32;
    pub fn sfb_mat_transpose(ctx: *mut sfb_ctx, a: *const sfb_mat, out: *mut *mut sfb_mat) -> i32;
    pub fn sfb_mat_view_rows(ctx: *mut sfb_ctx, a: *const sfb_mat, row0: u64, nrows: u64, out: *mut *mut sfb_mat) -> i32;
    pub fn sfb_mat_shape(a: *const sfb_mat, rows: *mut u64, cols: *mut u32) -> i32;
    pub fn sfb_mat_copy_rows(ctx: *mut sfb_ctx, a: *const sfb_mat, row0: u64, nrows: u64, out: *mut f64) -> i32;
    pub fn sfb_mat_free(a: *mut sfb_mat);
    pub fn sfb_knn_build(ctx: *mut sfb_ctx, rows: *const sfb_mat, params: *const sfb_knn_params, out: *mut *mut sfb_knn) -> i32;
    pub fn sfb_knn_build_sharded(ctx: *mut sfb_ctx, rows: *const sfb_mat, params: *const sfb_knn_params, out: *mut *mut sfb_knn) -> i32;
    pub fn sfb_knn_build_columns(ctx: *mut sfb_ctx, x: *const sfb_mat, params: *const sfb_knn_params, out: *mut *mut sfb_knn) -> i32;
    pub fn sfb_knn_shape(g: *const sfb_knn, rows: *mut u64, k: *mut u32, q_begin: *mut u64) -> i32;
    pub fn sfb_knn_copy(ctx: *mut sfb_ctx, g: *const sfb_knn, idx: *mut u32, dist: *mut f64, cnt: *mut u32) -> i32;
    pub fn sfb_knn_stats_get(g: *const sfb_knn, out: *mut sfb_knn_stats) -> i32;
    pub fn sfb_knn_from_host(ctx: *mut sfb_ctx, idx: *const u32, dist: *const f64, cnt: *const u32, rows: u64, k: u32, out: *mut *mut sfb_knn) -> i32;
    pub fn sfb_knn_free(g: *mut sfb_knn);
    pub fn sfb_knn_build_columns_begin(ctx: *mut sfb_ctx, x: *const sfb_mat, params: *const sfb_knn_params, sharded: i32, out: *mut *mut sfb_pending) -> i32;
    pub fn sfb_knn_build_columns_end(ctx: *mut sfb_ctx, pending: *mut sfb_pending, out: *mut *mut sfb_knn) -> i32;
    pub fn sfb_adjacency_build(ctx: *mut sfb_ctx, g: *const sfb_knn, params: *const sfb_adj_params, out: *mut *mut sfb_adj, sparsified: *mut i32) -> i32;
    pub fn sfb_sparsify_sfgrass(ctx: *mut sfb_ctx, adj: *mut sfb_adj, ratio: f64, applied: *mut i32) -> i32;
    pub fn sfb_adj_shape(adj: *const sfb_adj, rows: *mut u64, k: *mut u32) -> i32;
    pub fn sfb_adj_copy(ctx: *mut sfb_ctx, adj: *const sfb_adj, idx: *mut u32, w: *mut f64, cnt: *mut u32) -> i32;
    pub fn sfb_adj_from_host(ctx: *mut sfb_ctx, idx: *const u32, w: *const f64, cnt: *const u32, rows: u64, k: u32, out: *mut *mut sfb_adj) -> i32;
    pub fn sfb_adj_free(adj: *mut sfb_adj);
    pub fn sfb_laplacian_build(ctx: *mut sfb_ctx, adj: *const sfb_adj, params: *const sfb_lap_params, out: *mut *mut sfb_csr) -> i32;
    pub fn sfb_laplacian_build_rows(ctx: *mut sfb_ctx, adj: *const sfb_adj, params: *const sfb_lap_params, row_begin: u64, row_end: u64, out: *mut *mut sfb_csr) -> i32;
    pub fn sfb_csr_shape(L: *const sfb_csr, rows: *mut u64, nnz: *mut u64) -> i32;
    pub fn sfb_csr_copy(ctx: *mut sfb_ctx, L: *const sfb_csr, indptr: *mut u64, indices: *mut u32, data: *mut f64) -> i32;
    pub fn sfb_csr_from_host(ctx: *mut sfb_ctx, rows: u64, indptr: *const u64, indices: *const u32, data: *const f64, out: *mut *mut sfb_csr) -> i32;
    pub fn sfb_csr_free(L: *mut sfb_csr);
    pub fn sfb_spmv(ctx: *mut sfb_ctx, L: *const sfb_csr, x: *const f64, y: *mut f64) -> i32;
    pub fn sfb_rayleigh_quotient(ctx: *mut sfb_ctx, L: *const sfb_csr, x: *const f64, out: *mut f64) -> i32;
    pub fn sfb_lambda(ctx: *mut sfb_ctx, L: *const sfb_csr, x: *const sfb_mat, params: *const sfb_lambda_params, out_lambda: *mut f64, out_dispersion: *mut f64, stats: *mut f64) -> i32;
    pub fn sfb_lambda_projected(ctx: *mut sfb_ctx, L: *const sfb_csr, x_original: *const sfb_mat, x_projected: *const sfb_mat, params: *const sfb_lambda_params, out_lambda: *mut f64, out_dispersion: *mut f64, stats: *mut f64) -> i32;
    pub fn sfb_compute_tau_mode_lambdas(ctx: *mut sfb_ctx, L: *const sfb_csr, data: *const f32, n_items: u64, n_features: u32, out_lambdas: *mut f64) -> i32;
    pub fn sfb_compute_tau(ctx: *mut sfb_ctx, lambdas: *const f32, n: u64, tau_mode: i32, tau_value: f32, out_tau: *mut f32) -> i32;
    pub fn sfb_diffuse(ctx: *mut sfb_ctx, L: *const sfb_csr, x: *mut sfb_mat, eta: f64, steps: u32) -> i32;
    pub fn sfb_map_items_to_subcentroids(ctx: *mut sfb_ctx, items: *const sfb_mat, item_lambdas: *const f64, sub_centroids: *const sfb_mat, sub_lambdas: *const f64, epsilon: f64, out_idx: *mut u32, out_lambda: *mut f64, out_norm: *mut f64) -> i32;
    pub fn sfb_project_rows(ctx: *mut sfb_ctx, x: *const sfb_mat, samples: *const f64, reduced_dim: u32, order: i32, out: *mut *mut sfb_mat) -> i32;
    pub fn sfb_compute_jl_dimension(n_points: u64, original_dim: u64, epsilon: f64, core: i32, out: *mut u64) -> i32;
    pub fn sfb_sorted_lambdas_build(ctx: *mut sfb_ctx, lambdas: *const f64, n: u64, out_lambda: *mut f64, out_idx: *mut u32, out_std_dev: *mut f64) -> i32;
    pub fn sfb_build_laplacian_matrix(ctx: *mut sfb_ctx, items: *const f64, nodes: u64, dims: u32, params: *const sfb_graph_params, screen: i32, out: *mut *mut sfb_csr) -> i32;
    pub fn sfb_compute_taumode_lambdas(ctx: *mut sfb_ctx, L: *const sfb_csr, items: *const f64, n_items: u64, n_features: u32, tau_mode: i32, tau_value: f64, out_lambdas: *mut f64) -> i32;
    pub fn sfb_bc_adjacency_build(ctx: *mut sfb_ctx, means: *const f32, variances: *const f32, n_centroids: u32, n_features: u32, k: u32, variance_regularizer: f32, weight_threshold: f32, out: *mut *mut sfb_adj) -> i32;
    pub fn sfb_laplacian_stage_execute(ctx: *mut sfb_ctx, means: *const f32, variances: *const f32, n_centroids: u32, n_features: u32, cfg: *const sfb_laplacian_config, out: *mut *mut sfb_csr, degrees: *mut f32) -> i32;
    pub fn sfb_debug_screen_tile(ctx: *mut sfb_ctx, x: *const sfb_mat, metric: i32, screen: i32, row0: u64, col0: u64, out_tile: *mut f32, q_rows: *mut f32, q_cols: *mut f32, kpad_out: *mut u32, scale_out: *mut f64) -> i32;
    pub fn sfb_timings(ctx: *const sfb_ctx, out: *mut sfb_stage_times) -> i32;
    pub fn sfb_timer_start(ctx: *mut sfb_ctx) -> i32;
    pub fn sfb_timer_stop(ctx: *mut sfb_ctx, ms: *mut f64) -> i32;
    pub fn sfb_timings_reset(ctx: *mut sfb_ctx) -> i32;
    pub fn sfb_comm_unique_id(id: *mut u8) -> i32;
    pub fn sfb_comm_init(ctx: *mut sfb_ctx, id: *const u8, rank: i32, world: i32) -> i32;
    pub fn sfb_mat_allgather_rows(ctx: *mut sfb_ctx, shard: *const sfb_mat, total_rows: u64, out: *mut *mut sfb_mat) -> i32;
    pub fn sfb_knn_build_columns_sharded(ctx: *mut sfb_ctx, x: *const sfb_mat, params: *const sfb_knn_params, out: *mut *mut sfb_knn) -> i32;
    pub fn sfb_knn_allgather(ctx: *mut sfb_ctx, shard: *const sfb_knn, total_rows: u64, out: *mut *mut sfb_knn) -> i32;
    pub fn sfb_lambda_allgather(ctx: *mut sfb_ctx, L: *const sfb_csr, x_shard: *const sfb_mat, row0: u64, total_rows: u64, params: *const sfb_lambda_params, out_lambda: *mut f64, stats: *mut f64) -> i32;
    pub fn sfb_comm_barrier(ctx: *mut sfb_ctx) -> i32;
}
