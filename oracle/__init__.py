"""ctypes bindings for the CPU oracle (oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

METRIC_COSINE, METRIC_L2, METRIC_L2SQ = 0, 1, 2
TAU_FIXED, TAU_MEDIAN, TAU_MEAN, TAU_PERCENTILE = 0, 1, 2, 3
LAMBDA_LEGACY_TAUMODE, LAMBDA_ENERGY_NODE, LAMBDA_CORE_F32SEM = 0, 1, 2
IDX_NONE = 0xFFFFFFFF

_c = ctypes
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile liboracle.so (gcc, -ffp-contract=off).  Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = ctypes.CDLL(_SO)
    L.orc_num_threads.restype = _c.c_int
    L.orc_set_num_threads.argtypes = [_c.c_int]
    L.orc_generate_rows.argtypes = [_c.c_int, _c.c_uint64, _c.c_uint64, _c.c_uint64, _c.c_uint32,
                                    _c.c_uint32, _c.c_double, _dp]
    L.orc_row_norms.argtypes = [_dp, _c.c_uint64, _c.c_uint32, _dp]
    L.orc_knn.argtypes = [_dp, _c.c_uint64, _c.c_uint32, _c.c_int, _c.c_uint32, _c.c_double,
                          _c.c_void_p, _c.c_uint64, _u32p, _dp, _u32p]
    L.orc_build_adjacency.argtypes = [_u32p, _dp, _u32p, _c.c_uint64, _c.c_uint32, _c.c_double,
                                      _c.c_double, _c.c_int, _u32p, _dp, _u32p]
    L.orc_build_adjacency.restype = _c.c_int
    L.orc_sfgrass.argtypes = [_u32p, _dp, _u32p, _c.c_uint64, _c.c_uint32, _c.c_double]
    L.orc_sfgrass.restype = _c.c_int
    L.orc_laplacian_build.argtypes = [_u32p, _dp, _u32p, _c.c_uint64, _c.c_uint32, _c.c_int,
                                      _c.c_double]
    L.orc_laplacian_build.restype = _c.c_void_p
    L.orc_csr_nnz.argtypes = [_c.c_void_p]
    L.orc_csr_nnz.restype = _c.c_uint64
    L.orc_csr_copy.argtypes = [_c.c_void_p, _u64p, _u32p, _dp]
    L.orc_csr_free.argtypes = [_c.c_void_p]
    L.orc_spmv.argtypes = [_u64p, _u32p, _dp, _c.c_uint64, _dp, _dp]
    L.orc_rayleigh.argtypes = [_u64p, _u32p, _dp, _c.c_uint64, _dp]
    L.orc_rayleigh.restype = _c.c_double
    L.orc_select_tau.argtypes = [_dp, _c.c_uint64, _c.c_int, _c.c_double]
    L.orc_select_tau.restype = _c.c_double
    L.orc_lambda.argtypes = [_u64p, _u32p, _dp, _c.c_uint64, _dp, _c.c_uint64, _c.c_int, _c.c_int,
                             _c.c_double, _dp, _c.c_void_p, _c.c_void_p]
    L.orc_normalise_lambdas.argtypes = [_dp, _c.c_uint64, _dp]
    L.orc_diffuse.argtypes = [_u64p, _u32p, _dp, _c.c_uint64, _dp, _c.c_uint64, _c.c_double,
                              _c.c_uint32]
    L.orc_transpose.argtypes = [_dp, _c.c_uint64, _c.c_uint64, _dp]
    _fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    L.orc_map_items.argtypes = [_dp, _c.c_uint64, _c.c_uint32, _dp, _dp, _c.c_uint32, _dp, _c.c_double, _u32p, _dp, _dp]
    L.orc_bc.argtypes = [_fp, _fp, _c.c_uint32, _c.c_uint32, _c.c_uint32, _c.c_uint32, _c.c_float]
    L.orc_bc.restype = _c.c_float
    L.orc_bc_matrix.argtypes = [_fp, _fp, _c.c_uint32, _c.c_uint32, _c.c_float, _fp]
    L.orc_bc_knn.argtypes = [_fp, _fp, _c.c_uint32, _c.c_uint32, _c.c_uint32, _c.c_float, _c.c_float, _u32p, _fp, _u32p]
    L.orc_jl_dimension.argtypes = [_c.c_uint64, _c.c_uint64, _c.c_double]
    L.orc_jl_dimension.restype = _c.c_uint64
    L.orc_jl_dimension_core.argtypes = [_c.c_uint64, _c.c_uint64, _c.c_float]
    L.orc_jl_dimension_core.restype = _c.c_uint64
    L.orc_project_rows.argtypes = [_dp, _c.c_uint64, _c.c_uint32, _dp, _c.c_uint32, _dp]
    L.orc_project_rows_core.argtypes = [_fp, _c.c_uint64, _c.c_uint32, _fp, _c.c_uint32, _fp]
    L.orc_sorted_lambdas.argtypes = [_dp, _c.c_uint64, _dp, _u32p, _c.POINTER(_c.c_double)]
    L.orc_sorted_lambdas.restype = _c.c_int
    L.orc_lambda_projected.argtypes = [_u64p, _u32p, _dp, _c.c_uint64, _dp, _c.c_uint64, _dp, _c.c_uint64, _c.c_int, _c.c_double, _dp]
    L.orc_bc_det.argtypes = L.orc_bc.argtypes
    L.orc_bc_det.restype = _c.c_float
    L.orc_bc_matrix2.argtypes = [_fp, _fp, _c.c_uint32, _c.c_uint32, _c.c_float, _c.c_int, _fp]
    L.orc_bc_knn2.argtypes = [_fp, _fp, _c.c_uint32, _c.c_uint32, _c.c_uint32, _c.c_float, _c.c_float, _c.c_int, _u32p, _fp, _u32p]
    L.orc_det_logf.argtypes = [_c.c_float]
    L.orc_det_logf.restype = _c.c_float
    L.orc_det_expf.argtypes = [_c.c_float]
    L.orc_det_expf.restype = _c.c_float
    L.orc_knn_block.argtypes = [_dp, _dp, _u64p, _c.c_uint64, _dp, _c.c_uint64, _c.c_uint64, _c.c_uint32, _c.c_int, _c.c_uint32,
                                _c.c_double, _u32p, _dp, _u32p]
    L.orc_knn_block_finish.argtypes = [_c.c_uint64, _c.c_uint32, _u32p, _dp, _u32p]
    L.orc_cols_gram_accumulate.argtypes = [_dp, _c.c_uint64, _c.c_uint32, _u32p, _c.c_uint32, _dp, _dp]
    L.orc_cols_gram_finish.argtypes = [_dp, _dp, _c.c_uint32, _u32p, _c.c_uint32, _c.c_uint32, _c.c_double, _u32p, _dp, _u32p]
    L.orc_compute_tau_core.argtypes = [_fp, _c.c_uint64, _c.c_int, _c.c_float]
    L.orc_compute_tau_core.restype = _c.c_float
    _lib = L
    return L


def num_threads():
    return int(lib().orc_num_threads())


def use_all_threads():
    """OpenMP threads = host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


def generate_rows(kind, seed, row0, nrows, kdim, n_centres=0, noise=0.0):
    out = np.empty((nrows, kdim), dtype=np.float64)
    lib().orc_generate_rows(kind, seed, row0, nrows, kdim, n_centres, noise, out)
    return out


def row_norms(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(x.shape[0], dtype=np.float64)
    lib().orc_row_norms(x, x.shape[0], x.shape[1], out)
    return out


def knn(x, k, metric=METRIC_COSINE, eps=np.inf, query_rows=None):
    """Brute-force kNN (test_helpers.rs:73-133): returns idx (nq,k) u32, dist (nq,k) f64, cnt (nq,)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    m, kd = x.shape
    if query_rows is None:
        nq, qp = m, None
    else:
        q = np.ascontiguousarray(query_rows, dtype=np.uint64)
        nq, qp = q.shape[0], q.ctypes.data_as(_c.c_void_p)
    idx = np.empty((nq, k), dtype=np.uint32)
    dist = np.empty((nq, k), dtype=np.float64)
    cnt = np.empty(nq, dtype=np.uint32)
    lib().orc_knn(x, m, kd, metric, k, float(eps), qp, nq, idx, dist, cnt)
    return idx, dist, cnt


def build_adjacency(idx, dist, cnt, p, sigma, force_sparsify=-1):
    """laplacian.rs:219-290: kernel weights + inline sparsification.  Returns (adj_idx, adj_w, adj_cnt, applied)."""
    m, k = idx.shape
    a_idx = np.empty((m, k), dtype=np.uint32)
    a_w = np.empty((m, k), dtype=np.float64)
    a_cnt = np.empty(m, dtype=np.uint32)
    applied = lib().orc_build_adjacency(np.ascontiguousarray(idx), np.ascontiguousarray(dist),
                                        np.ascontiguousarray(cnt), m, k, float(p), float(sigma),
                                        int(force_sparsify), a_idx, a_w, a_cnt)
    return a_idx, a_w, a_cnt, bool(applied)


def sfgrass(a_idx, a_w, a_cnt, ratio=0.5):
    """sparsification.rs:32-113 on copies.  Returns (adj_idx, adj_w, adj_cnt, applied)."""
    a_idx, a_w, a_cnt = a_idx.copy(), a_w.copy(), a_cnt.copy()
    m, k = a_idx.shape
    applied = lib().orc_sfgrass(a_idx, a_w, a_cnt, m, k, float(ratio))
    return a_idx, a_w, a_cnt, bool(applied)


def laplacian(a_idx, a_w, a_cnt, normalised=False, weight_threshold=1e-9):
    """laplacian.rs:297-419: symmetrise + L = D - W as CSR (indptr u64, indices u32, data f64)."""
    m, k = a_idx.shape
    h = lib().orc_laplacian_build(np.ascontiguousarray(a_idx), np.ascontiguousarray(a_w),
                                  np.ascontiguousarray(a_cnt), m, k, int(normalised),
                                  float(weight_threshold))
    nnz = lib().orc_csr_nnz(h)
    indptr = np.empty(m + 1, dtype=np.uint64)
    indices = np.empty(max(nnz, 1), dtype=np.uint32)
    data = np.empty(max(nnz, 1), dtype=np.float64)
    lib().orc_csr_copy(h, indptr, indices, data)
    lib().orc_csr_free(h)
    return indptr, indices[:nnz], data[:nnz]


def _csr_args(indptr, indices, data):
    return (np.ascontiguousarray(indptr, dtype=np.uint64),
            np.ascontiguousarray(indices, dtype=np.uint32) if len(indices) else np.zeros(1, np.uint32),
            np.ascontiguousarray(data, dtype=np.float64) if len(data) else np.zeros(1, np.float64))


def spmv(indptr, indices, data, x):
    ip, ix, dv = _csr_args(indptr, indices, data)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().orc_spmv(ip, ix, dv, len(ip) - 1, x, y)
    return y


def rayleigh(indptr, indices, data, x):
    ip, ix, dv = _csr_args(indptr, indices, data)
    return float(lib().orc_rayleigh(ip, ix, dv, len(ip) - 1, np.ascontiguousarray(x, dtype=np.float64)))


def select_tau(energies, mode, value=0.0):
    e = np.ascontiguousarray(energies, dtype=np.float64)
    if e.size == 0:
        e = np.zeros(1, dtype=np.float64)
        n = 0
    else:
        n = e.size
    return float(lib().orc_select_tau(e, n, mode, float(value)))


def lambdas(indptr, indices, data, x, variant=LAMBDA_LEGACY_TAUMODE, tau_mode=TAU_MEDIAN,
            tau_value=0.0, with_parts=False):
    ip, ix, dv = _csr_args(indptr, indices, data)
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, f = x.shape
    assert f == len(ip) - 1
    lam = np.empty(n, dtype=np.float64)
    e = np.empty(n, dtype=np.float64)
    g = np.empty(n, dtype=np.float64)
    lib().orc_lambda(ip, ix, dv, f, x, n, variant, tau_mode, float(tau_value), lam,
                     e.ctypes.data_as(_c.c_void_p), g.ctypes.data_as(_c.c_void_p))
    return (lam, e, g) if with_parts else lam


def normalise_lambdas(lam):
    lam = np.array(lam, dtype=np.float64, copy=True)
    stats = np.empty(3, dtype=np.float64)
    lib().orc_normalise_lambdas(lam, lam.size, stats)
    return lam, stats


def diffuse(indptr, indices, data, x, eta, steps):
    ip, ix, dv = _csr_args(indptr, indices, data)
    x = np.array(x, dtype=np.float64, copy=True, order="C")
    lib().orc_diffuse(ip, ix, dv, len(ip) - 1, x, x.shape[0], float(eta), int(steps))
    return x


def transpose(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty((x.shape[1], x.shape[0]), dtype=np.float64)
    lib().orc_transpose(x, x.shape[0], x.shape[1], out)
    return out


# ---- successor Stage C: Bhattacharyya-coefficient feature graph (f32) --------------------------
def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def bc(means, variances, i, j, reg=1e-6, det=False):
    """bhattacharyya_coefficient (surfface-core/src/distance.rs:260-290) of features i, j of a [C, F] state.
    det=False: glibc logf / expf (the reference on Linux); det=True: the portable log / exp the device evaluates."""
    m, v = _f32(means), _f32(variances)
    fn = lib().orc_bc_det if det else lib().orc_bc
    return float(fn(m, v, m.shape[0], m.shape[1], i, j, reg))


def bc_matrix(means, variances, reg=1e-6, det=False):
    m, v = _f32(means), _f32(variances)
    out = np.empty((m.shape[1], m.shape[1]), np.float32)
    lib().orc_bc_matrix2(m, v, m.shape[0], m.shape[1], reg, int(det), out)
    return out


def bc_knn(means, variances, k, reg=1e-6, thr=1e-9, det=False):
    """compute_bhattacharyya_weights (surfface-core/src/laplacian.rs:254-298): (idx [F,k], w [F,k] f32, cnt [F])."""
    m, v = _f32(means), _f32(variances)
    f = m.shape[1]
    idx = np.empty((f, k), np.uint32)
    w = np.empty((f, k), np.float32)
    cnt = np.empty(f, np.uint32)
    lib().orc_bc_knn2(m, v, m.shape[0], f, k, reg, thr, int(det), idx, w, cnt)
    return idx, w, cnt


def det_logf(a):
    return float(lib().orc_det_logf(float(a)))


def det_expf(a):
    return float(lib().orc_det_expf(float(a)))


def compute_tau_core(lambdas, mode, value=0.0):
    """compute_tau (surfface-core/src/taumode.rs:37-65): one f32 tau from the lambda distribution."""
    v = _f32(lambdas)
    if v.size == 0:
        return float(lib().orc_compute_tau_core(np.zeros(1, np.float32), 0, mode, float(value)))
    return float(lib().orc_compute_tau_core(v, v.size, mode, float(value)))


class KnnStream:
    """orc_knn over a corpus fed in ascending row blocks (for matrices that never exist on the host in full)."""

    def __init__(self, queries, query_rows, k, metric=METRIC_COSINE, eps=np.inf):
        self.q = np.ascontiguousarray(queries, dtype=np.float64)
        self.rows = np.ascontiguousarray(query_rows, dtype=np.uint64)
        self.k, self.metric, self.eps = int(k), int(metric), float(eps)
        nq = self.q.shape[0]
        self.qn = row_norms(self.q) if metric == METRIC_COSINE else np.zeros(nq)
        self.idx = np.zeros((nq, self.k), np.uint32)
        self.dist = np.zeros((nq, self.k), np.float64)
        self.cnt = np.zeros(nq, np.uint32)

    def feed(self, block, row0):
        b = np.ascontiguousarray(block, dtype=np.float64)
        lib().orc_knn_block(self.q, self.qn, self.rows, self.q.shape[0], b, int(row0), b.shape[0], b.shape[1], self.metric,
                            self.k, self.eps, self.idx, self.dist, self.cnt)

    def finish(self):
        lib().orc_knn_block_finish(self.q.shape[0], self.k, self.idx, self.dist, self.cnt)
        return self.idx, self.dist, self.cnt


class ColsGramStream:
    """Rectified-cosine kNN lists of a SAMPLE of the feature nodes (columns) of a row-streamed item matrix: the sums
    orc_knn(orc_transpose(x)) would form, carried across row blocks."""

    def __init__(self, f, cols, k, eps=np.inf):
        self.f, self.k, self.eps = int(f), int(k), float(eps)
        self.cols = np.ascontiguousarray(cols, dtype=np.uint32)
        self.acc = np.zeros((len(self.cols), self.f), np.float64)
        self.nrm2 = np.zeros(self.f, np.float64)

    def feed(self, block):
        b = np.ascontiguousarray(block, dtype=np.float64)
        lib().orc_cols_gram_accumulate(b, b.shape[0], self.f, self.cols, len(self.cols), self.acc, self.nrm2)

    def finish(self):
        ns = len(self.cols)
        idx = np.empty((ns, self.k), np.uint32); dist = np.empty((ns, self.k), np.float64); cnt = np.empty(ns, np.uint32)
        lib().orc_cols_gram_finish(self.acc, self.nrm2, self.f, self.cols, ns, self.k, self.eps, idx, dist, cnt)
        return idx, dist, cnt


def map_items(items, item_lambdas, sub_centroids, sub_lambdas, epsilon=1e-11):
    """Item -> sub-centroid mapping by |delta lambda| with cosine tie-break (energymaps.rs:1246-1342)."""
    x = np.ascontiguousarray(items, dtype=np.float64)
    sc = np.ascontiguousarray(sub_centroids, dtype=np.float64)
    il = np.ascontiguousarray(item_lambdas, dtype=np.float64)
    sl = np.ascontiguousarray(sub_lambdas, dtype=np.float64)
    n, f = x.shape
    idx = np.empty(n, np.uint32); lam = np.empty(n, np.float64); norm = np.empty(n, np.float64)
    lib().orc_map_items(x, n, f, il, sc, sc.shape[0], sl, float(epsilon), idx, lam, norm)
    return idx, lam, norm


def jl_dimension(n_points, original_dim, epsilon, core=False):
    """compute_jl_dimension: reduction.rs:117-171 (f64) / surfface-core clustering.rs:113-123 (f32)."""
    if core:
        return int(lib().orc_jl_dimension_core(n_points, original_dim, float(epsilon)))
    return int(lib().orc_jl_dimension(n_points, original_dim, float(epsilon)))


def project_rows(x, samples):
    """project_matrix (reduction.rs:175-242): samples is original_dim x reduced_dim (draw order of the reference)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    s = np.ascontiguousarray(samples, dtype=np.float64)
    assert s.shape[0] == x.shape[1]
    out = np.empty((x.shape[0], s.shape[1]), np.float64)
    lib().orc_project_rows(x, x.shape[0], x.shape[1], s, s.shape[1], out)
    return out


def project_rows_core(x, samples):
    """surfface-core clustering.rs:84-109 (f32): samples is reduced_dim x original_dim (its draw order)."""
    x = _f32(x)
    s = _f32(samples)
    assert s.shape[1] == x.shape[1]
    out = np.empty((x.shape[0], s.shape[0]), np.float32)
    lib().orc_project_rows_core(x, x.shape[0], x.shape[1], s, s.shape[0], out)
    return out


def sorted_lambdas(lam):
    """SortedLambdas::build_from + to_vec (sorted_index.rs:22-57): (lambda_sorted, idx, std_dev)."""
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    out = np.empty(lam.shape[0], np.float64)
    idx = np.empty(lam.shape[0], np.uint32)
    sd = _c.c_double()
    if lib().orc_sorted_lambdas(lam, lam.shape[0], out, idx, _c.byref(sd)) != 0:
        raise ValueError("empty lambdas: the reference panics")
    return out, idx, sd.value


def lambdas_projected(indptr, indices, data, x_projected, x_original, tau_mode=TAU_MEDIAN, tau_value=0.0):
    """compute_synthetic_lambda with a projection (taumode.rs:261-318): tau and the zero test from the unprojected
    item, energy and dispersion from the projected one."""
    ip, ix, dv = _csr_args(indptr, indices, data)
    xp = np.ascontiguousarray(x_projected, dtype=np.float64)
    xo = np.ascontiguousarray(x_original, dtype=np.float64)
    assert xp.shape[0] == xo.shape[0] and xp.shape[1] == len(ip) - 1
    lam = np.empty(xp.shape[0], dtype=np.float64)
    lib().orc_lambda_projected(ip, ix, dv, xp.shape[1], xp, xp.shape[0], xo, xo.shape[1], tau_mode, float(tau_value), lam)
    return lam
