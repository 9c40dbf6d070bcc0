"""surfface_b200 -- host-side mirror of the reference's graph-wiring API over the sm_100a C ABI.

Two layers:
  * handle layer (Context, Matrix, KnnGraph, Adjacency, Csr): one Python object per C-ABI handle;
  * reference layer: the names and argument meaning of the reference's operator interface for this
    path -- GraphParams / GraphLaplacian / GraphFactory / build_laplacian_matrix
    (src_legacy/graph.rs:94-136,193-255, src_legacy/laplacian.rs:122-180), TauMode
    (src_legacy/taumode.rs:16-23,117-121), SfGrassSparsifier (src_legacy/sparsification.rs:14-36).

The reference is Rust and there is no Rust toolchain here, so the tests drive
the C ABI through this module; the Rust `-sys` crate a maintainer would add is in rust/ and
INTEGRATION.md.  Everything computes on the GPU; nothing here falls back to the CPU.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _ffi
from ._ffi import IDX_NONE, SfbError, lib

METRIC_COSINE, METRIC_L2, METRIC_L2SQ = 0, 1, 2
SCREEN_AUTO, SCREEN_EXACT_F64, SCREEN_F16, SCREEN_BF16 = 0, 1, 2, 3
LAMBDA_LEGACY_TAUMODE, LAMBDA_ENERGY_NODE, LAMBDA_CORE_F32SEM = 0, 1, 2
TAU_FIXED, TAU_MEDIAN, TAU_MEAN, TAU_PERCENTILE = 0, 1, 2, 3
SYNTH_GAUSSIAN, SYNTH_CLUSTERED, SYNTH_ANISOTROPIC = 0, 1, 2
PROJECT_LEGACY, PROJECT_CORE_F32 = 0, 1


_PINNED = {}  # page-locked buffers stay alive for the life of the process


# ------------------------------------------------------------------------------------------------
# handle layer
# ------------------------------------------------------------------------------------------------
class Context:
    """sfb_ctx: one per host thread; owns the device, its stream and (optionally) an NCCL communicator."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        st = lib().sfb_ctx_create(int(device), C.byref(self._h))
        if st != _ffi.OK:
            self._h = C.c_void_p()
            raise SfbError(st, f"sfb_ctx_create(device={device}) failed (no CUDA device, or not sm_100)")
        self.device = device

    def check(self, st):
        if st != _ffi.OK:
            raise SfbError(st, lib().sfb_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().sfb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self.check(lib().sfb_synchronize(self._h))

    def device_info(self):
        name = C.create_string_buffer(256)
        sm = C.c_int32()
        mem = C.c_uint64()
        self.check(lib().sfb_device_info(self._h, name, C.byref(sm), C.byref(mem)))
        return name.value.decode(), sm.value, mem.value

    def timings(self):
        t = _ffi.StageTimes()
        self.check(lib().sfb_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in t._fields_}

    def timer_start(self):
        self.check(lib().sfb_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self.check(lib().sfb_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def timings_reset(self):
        self.check(lib().sfb_timings_reset(self._h))

    def pinned_empty(self, shape, dtype=np.float64):
        """numpy array over page-locked host memory (kept alive by the returned array's base)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(lib().sfb_pinned_alloc(self._h, n, C.byref(p)))
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        _PINNED[p.value] = buf
        return arr

    # -- matrices
    def matrix(self, x):
        x = _ffi.f64(x)
        if x.ndim != 2:
            raise ValueError("matrix must be 2-D")
        h = C.c_void_p()
        self.check(lib().sfb_mat_from_host(self._h, _ffi.ptr(x), x.shape[0], x.shape[1], C.byref(h)))
        self.synchronize()
        return Matrix(self, h)

    def matrix_f32(self, x):
        """Upload an f32 host matrix (4 bytes per value over PCIe), widened to f64 on the device."""
        x = _ffi.f32(x)
        if x.ndim != 2:
            raise ValueError("matrix must be 2-D")
        h = C.c_void_p()
        self.check(lib().sfb_mat_from_host_f32(self._h, _ffi.ptr(x), x.shape[0], x.shape[1], C.byref(h)))
        return Matrix(self, h)

    def matrix_copy(self, m):
        """Device copy of a Matrix (or of a row view)."""
        h = C.c_void_p()
        self.check(lib().sfb_mat_clone(self._h, m._h, C.byref(h)))
        return Matrix(self, h)

    def compute_tau(self, lambdas, mode=TAU_MEDIAN, value=0.0):
        """compute_tau (surfface-core/src/taumode.rs:37-65): one f32 tau from the lambda distribution."""
        v = _ffi.f32(lambdas).ravel()
        out = C.c_float()
        self.check(lib().sfb_compute_tau(self._h, _ffi.ptr(v) if v.size else None, v.size, int(mode), float(value), C.byref(out)))
        return out.value

    def generate(self, kind, seed, rows, cols, n_centres=0, noise=0.0):
        h = C.c_void_p()
        self.check(lib().sfb_mat_generate(self._h, kind, seed, rows, cols, n_centres, float(noise), C.byref(h)))
        return Matrix(self, h)

    # -- multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self.check(lib().sfb_comm_init(self._h, buf, rank, world))

    def barrier(self):
        self.check(lib().sfb_comm_barrier(self._h))


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    st = lib().sfb_comm_unique_id(buf)
    if st != _ffi.OK:
        raise SfbError(st, "sfb_comm_unique_id failed (libnccl not loadable?)")
    return bytes(buf)


class _Handle:
    _free = None

    def __init__(self, ctx, h):
        self.ctx, self._h = ctx, h

    def free(self):
        if self._h:
            getattr(lib(), self._free)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Matrix(_Handle):
    _free = "sfb_mat_free"

    @property
    def shape(self):
        r, c = C.c_uint64(), C.c_uint32()
        lib().sfb_mat_shape(self._h, C.byref(r), C.byref(c))
        return r.value, c.value

    def transpose(self):
        h = C.c_void_p()
        self.ctx.check(lib().sfb_mat_transpose(self.ctx._h, self._h, C.byref(h)))
        return Matrix(self.ctx, h)

    def allgather_rows(self, total_rows):
        """Collective: this rank's row shard -> the full matrix on every GPU (all-gather over NVLink)."""
        h = C.c_void_p()
        self.ctx.check(lib().sfb_mat_allgather_rows(self.ctx._h, self._h, total_rows, C.byref(h)))
        return Matrix(self.ctx, h)

    def view_rows(self, row0, nrows):
        h = C.c_void_p()
        self.ctx.check(lib().sfb_mat_view_rows(self.ctx._h, self._h, row0, nrows, C.byref(h)))
        v = Matrix(self.ctx, h)
        v._parent = self  # keep the parent alive
        return v

    def rows(self, row0=0, nrows=None):
        r, c = self.shape
        nrows = r - row0 if nrows is None else nrows
        out = np.empty((nrows, c), dtype=np.float64)
        self.ctx.check(lib().sfb_mat_copy_rows(self.ctx._h, self._h, row0, nrows, _ffi.ptr(out)))
        return out

    def knn(self, k, metric=METRIC_COSINE, eps=math.inf, screen=SCREEN_AUTO, k_prime=0, q_begin=0, q_end=0,
            allow_fallback=True, sharded=False):
        """sharded=True: the collective form (every rank of the communicator calls it with the same matrix and its own
        ceil-split [q_begin, q_end)): the operand preparation is split across the ranks."""
        p = _ffi.KnnParams(metric, k, float(eps), screen, k_prime, q_begin, q_end, int(allow_fallback))
        h = C.c_void_p()
        fn = lib().sfb_knn_build_sharded if sharded else lib().sfb_knn_build
        self.ctx.check(fn(self.ctx._h, self._h, C.byref(p), C.byref(h)))
        return KnnGraph(self.ctx, h)

    def knn_columns(self, k, metric=METRIC_COSINE, eps=math.inf, screen=SCREEN_AUTO, sharded=False):
        """kNN graph over the COLUMNS of this matrix (the feature graph, graph.rs:193-216), no transposed copy.
        sharded=True: collective over the context's communicator (every rank holds the same matrix)."""
        p = _ffi.KnnParams(metric, k, float(eps), screen, 0, 0, 0, 1)
        h = C.c_void_p()
        fn = lib().sfb_knn_build_columns_sharded if sharded else lib().sfb_knn_build_columns
        self.ctx.check(fn(self.ctx._h, self._h, C.byref(p), C.byref(h)))
        return KnnGraph(self.ctx, h)

    def knn_columns_begin(self, k, metric=METRIC_COSINE, eps=math.inf, sharded=False):
        """Registers the feature-graph build so that it runs beside the next knn() of this context (hidden behind the
        tensor-core screen); call .end() on the result to get the graph.  This matrix must stay alive until then."""
        p = _ffi.KnnParams(metric, k, float(eps), SCREEN_AUTO, 0, 0, 0, 1)
        h = C.c_void_p()
        self.ctx.check(lib().sfb_knn_build_columns_begin(self.ctx._h, self._h, C.byref(p), int(sharded), C.byref(h)))
        return PendingKnn(self.ctx, h, self)

    def debug_screen_tile(self, metric=METRIC_COSINE, screen=SCREEN_F16, row0=0, col0=0):
        """Diagnostic: (tile 128x256 f32 accumulators, operands of the 128 rows, operands of the 256 columns, scale)."""
        r, c = self.shape
        kpad = (c + 63) // 64 * 64
        tile = np.empty((128, 256), np.float32)
        qr = np.empty((128, kpad), np.float32)
        qc = np.empty((256, kpad), np.float32)
        kp, sc = C.c_uint32(), C.c_double()
        self.ctx.check(lib().sfb_debug_screen_tile(self.ctx._h, self._h, metric, screen, row0, col0, _ffi.ptr(tile),
                                                   _ffi.ptr(qr), _ffi.ptr(qc), C.byref(kp), C.byref(sc)))
        assert kp.value == kpad
        return tile, qr, qc, sc.value

    def map_to_subcentroids(self, item_lambdas, sub_centroids, sub_lambdas, epsilon=1e-11):
        """Energy pipeline (energymaps.rs:1246-1342): (index of the chosen sub-centroid, its lambda, |item|) per row."""
        n = self.shape[0]
        il, sl = _ffi.f64(item_lambdas), _ffi.f64(sub_lambdas)
        idx = np.empty(n, np.uint32); lam = np.empty(n, np.float64); norm = np.empty(n, np.float64)
        self.ctx.check(lib().sfb_map_items_to_subcentroids(self.ctx._h, self._h, _ffi.ptr(il), sub_centroids._h, _ffi.ptr(sl),
                                                           float(epsilon), _ffi.ptr(idx), _ffi.ptr(lam), _ffi.ptr(norm)))
        return idx, lam, norm

    def diffuse(self, L, eta, steps):
        self.ctx.check(lib().sfb_diffuse(self.ctx._h, L._h, self._h, float(eta), int(steps)))
        return self

    def project(self, samples, order=PROJECT_LEGACY):
        """project_matrix (reduction.rs:175-242): `samples` are the StandardNormal draws, original_dim x reduced_dim
        (PROJECT_LEGACY) or reduced_dim x original_dim (PROJECT_CORE_F32, clustering.rs:84-109).  Stays on the device."""
        s = _ffi.f64(samples)
        f = self.shape[1]
        if s.ndim != 2 or s.shape[0 if order == PROJECT_LEGACY else 1] != f:
            raise ValueError(f"samples {s.shape} do not match {f} original dimensions")
        r = s.shape[1 if order == PROJECT_LEGACY else 0]
        out = C.c_void_p()
        self.ctx.check(lib().sfb_project_rows(self.ctx._h, self._h, _ffi.ptr(s), r, int(order), C.byref(out)))
        return Matrix(self.ctx, out)


class PendingKnn:
    """sfb_pending: a feature-graph build in flight (Matrix.knn_columns_begin)."""

    def __init__(self, ctx, h, keep):
        self.ctx, self._h, self._keep = ctx, h, keep

    def end(self):
        if not self._h:
            raise ValueError("end() was already called")
        out = C.c_void_p()
        h, self._h = self._h, None
        self.ctx.check(lib().sfb_knn_build_columns_end(self.ctx._h, h, C.byref(out)))
        return KnnGraph(self.ctx, out)


class KnnGraph(_Handle):
    _free = "sfb_knn_free"

    @property
    def shape(self):
        r, k, q = C.c_uint64(), C.c_uint32(), C.c_uint64()
        lib().sfb_knn_shape(self._h, C.byref(r), C.byref(k), C.byref(q))
        return r.value, k.value

    @property
    def q_begin(self):
        q = C.c_uint64()
        lib().sfb_knn_shape(self._h, None, None, C.byref(q))
        return q.value

    def to_host(self, out=None):
        r, k = self.shape
        if out is None:
            out = (np.empty((r, k), np.uint32), np.empty((r, k), np.float64), np.empty(r, np.uint32))
        idx, dist, cnt = out
        self.ctx.check(lib().sfb_knn_copy(self.ctx._h, self._h, _ffi.ptr(idx), _ffi.ptr(dist), _ffi.ptr(cnt)))
        self.ctx.synchronize()
        return idx, dist, cnt

    def stats(self):
        s = _ffi.KnnStats()
        lib().sfb_knn_stats_get(self._h, C.byref(s))
        return {n: getattr(s, n) for n, _ in s._fields_}

    @staticmethod
    def from_host(ctx, idx, dist, cnt):
        idx, dist, cnt = _ffi.u32(idx), _ffi.f64(dist), _ffi.u32(cnt)
        h = C.c_void_p()
        ctx.check(lib().sfb_knn_from_host(ctx._h, _ffi.ptr(idx), _ffi.ptr(dist), _ffi.ptr(cnt), idx.shape[0],
                                           idx.shape[1], C.byref(h)))
        return KnnGraph(ctx, h)

    def allgather(self, total_rows):
        h = C.c_void_p()
        self.ctx.check(lib().sfb_knn_allgather(self.ctx._h, self._h, total_rows, C.byref(h)))
        return KnnGraph(self.ctx, h)

    def adjacency(self, p=2.0, sigma=1.0, sparsify=-1):
        prm = _ffi.AdjParams(float(p), float(sigma), int(sparsify))
        h, applied = C.c_void_p(), C.c_int32()
        self.ctx.check(lib().sfb_adjacency_build(self.ctx._h, self._h, C.byref(prm), C.byref(h), C.byref(applied)))
        a = Adjacency(self.ctx, h)
        a.sparsified = bool(applied.value)
        return a


class Adjacency(_Handle):
    _free = "sfb_adj_free"
    sparsified = False

    @property
    def shape(self):
        r, k = C.c_uint64(), C.c_uint32()
        lib().sfb_adj_shape(self._h, C.byref(r), C.byref(k))
        return r.value, k.value

    def to_host(self):
        r, k = self.shape
        idx = np.empty((r, k), np.uint32)
        w = np.empty((r, k), np.float64)
        cnt = np.empty(r, np.uint32)
        self.ctx.check(lib().sfb_adj_copy(self.ctx._h, self._h, _ffi.ptr(idx), _ffi.ptr(w), _ffi.ptr(cnt)))
        return idx, w, cnt

    @staticmethod
    def from_host(ctx, idx, w, cnt):
        idx, w, cnt = _ffi.u32(idx), _ffi.f64(w), _ffi.u32(cnt)
        h = C.c_void_p()
        ctx.check(lib().sfb_adj_from_host(ctx._h, _ffi.ptr(idx), _ffi.ptr(w), _ffi.ptr(cnt), idx.shape[0],
                                           idx.shape[1], C.byref(h)))
        return Adjacency(ctx, h)

    def sfgrass(self, ratio=0.5):
        applied = C.c_int32()
        self.ctx.check(lib().sfb_sparsify_sfgrass(self.ctx._h, self._h, float(ratio), C.byref(applied)))
        return bool(applied.value)

    def laplacian(self, normalised=False, weight_threshold=1e-9, rows=None):
        """rows=(begin, end): only those rows of the Laplacian (row-owned assembly of a sharded build): a Csr of
        end - begin rows with global column indices."""
        prm = _ffi.LapParams(int(normalised), float(weight_threshold))
        h = C.c_void_p()
        if rows is None:
            self.ctx.check(lib().sfb_laplacian_build(self.ctx._h, self._h, C.byref(prm), C.byref(h)))
        else:
            self.ctx.check(lib().sfb_laplacian_build_rows(self.ctx._h, self._h, C.byref(prm), int(rows[0]), int(rows[1]), C.byref(h)))
        return Csr(self.ctx, h)


class Csr(_Handle):
    _free = "sfb_csr_free"

    @property
    def shape(self):
        r, nnz = C.c_uint64(), C.c_uint64()
        lib().sfb_csr_shape(self._h, C.byref(r), C.byref(nnz))
        return r.value, nnz.value

    def to_host(self, out=None):
        r, nnz = self.shape
        if out is None:
            out = (np.empty(r + 1, np.uint64), np.empty(max(nnz, 1), np.uint32), np.empty(max(nnz, 1), np.float64))
        indptr, indices, data = out
        assert len(indices) >= nnz and len(data) >= nnz
        self.ctx.check(lib().sfb_csr_copy(self.ctx._h, self._h, _ffi.ptr(indptr), _ffi.ptr(indices), _ffi.ptr(data)))
        self.ctx.synchronize()
        return indptr, indices[:nnz], data[:nnz]

    @staticmethod
    def from_host(ctx, indptr, indices, data):
        indptr = _ffi.u64(indptr)
        indices = _ffi.u32(indices) if len(indices) else np.zeros(1, np.uint32)
        data = _ffi.f64(data) if len(data) else np.zeros(1, np.float64)
        h = C.c_void_p()
        ctx.check(lib().sfb_csr_from_host(ctx._h, len(indptr) - 1, _ffi.ptr(indptr), _ffi.ptr(indices),
                                           _ffi.ptr(data), C.byref(h)))
        return Csr(ctx, h)

    def spmv(self, x):
        x = _ffi.f64(x)
        y = np.empty_like(x)
        self.ctx.check(lib().sfb_spmv(self.ctx._h, self._h, _ffi.ptr(x), _ffi.ptr(y)))
        return y

    def rayleigh_quotient(self, x):
        x = _ffi.f64(x)
        out = C.c_double()
        self.ctx.check(lib().sfb_rayleigh_quotient(self.ctx._h, self._h, _ffi.ptr(x), C.byref(out)))
        return out.value

    def lambdas(self, x: Matrix, variant=LAMBDA_LEGACY_TAUMODE, tau_mode=TAU_MEDIAN, tau_value=0.0,
                normalise=False, with_dispersion=False):
        n = x.shape[0]
        prm = _ffi.LambdaParams(variant, tau_mode, float(tau_value), int(normalise))
        lam = np.empty(n, np.float64)
        disp = np.empty(n, np.float64) if with_dispersion else None
        stats = np.empty(3, np.float64)
        self.ctx.check(lib().sfb_lambda(self.ctx._h, self._h, x._h, C.byref(prm), _ffi.ptr(lam), _ffi.ptr(disp),
                                        _ffi.ptr(stats)))
        self.ctx.synchronize()
        return (lam, disp, stats) if with_dispersion else (lam, stats)

    def lambdas_projected(self, x_original: Matrix, x_projected: Matrix, tau_mode=TAU_MEDIAN, tau_value=0.0, normalise=False):
        """Taumode lambda of JL-projected items (taumode.rs:261-318): tau and the zero test from the unprojected rows."""
        n = x_projected.shape[0]
        prm = _ffi.LambdaParams(LAMBDA_LEGACY_TAUMODE, tau_mode, float(tau_value), int(normalise))
        lam = np.empty(n, np.float64)
        stats = np.empty(3, np.float64)
        self.ctx.check(lib().sfb_lambda_projected(self.ctx._h, self._h, x_original._h, x_projected._h, C.byref(prm), _ffi.ptr(lam), None,
                                                  _ffi.ptr(stats)))
        self.ctx.synchronize()
        return lam, stats

    def lambdas_allgather(self, x_shard: Matrix, row0, total_rows, variant=LAMBDA_LEGACY_TAUMODE,
                          tau_mode=TAU_MEDIAN, tau_value=0.0, normalise=True):
        prm = _ffi.LambdaParams(variant, tau_mode, float(tau_value), int(normalise))
        lam = np.empty(total_rows, np.float64)
        stats = np.empty(3, np.float64)
        self.ctx.check(lib().sfb_lambda_allgather(self.ctx._h, self._h, x_shard._h, row0, total_rows, C.byref(prm),
                                                  _ffi.ptr(lam), _ffi.ptr(stats)))
        return lam, stats


# ------------------------------------------------------------------------------------------------
# reference layer
# ------------------------------------------------------------------------------------------------
_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


@dataclass
class GraphParams:
    """src_legacy/graph.rs:94-102.  `k` is carried but unused, as in the reference (only `topk`
    drives the neighbour count, laplacian.rs:213,223,248).  sigma=None means 1.0
    (`params.sigma.unwrap_or(1.0)`, laplacian.rs:256)."""
    eps: float = 1e-3
    k: int = 6
    topk: int = 3
    p: float = 2.0
    sigma: Optional[float] = None
    normalise: bool = False
    sparsity_check: bool = False


@dataclass
class LambdaGraphBuilder:
    """The lambda-graph half of the reference's builder (surfface-pipeline/src/builder.rs): defaults :105-111,
    `with_lambda_graph` :629-657, `define_result_k` :785-793 (run first thing by every build, :839,1096), and the
    GraphParams the build hands to the Laplacian stage (`GraphFactory::build_laplacian_matrix_from_k_cluster`)."""
    lambda_eps: float = 1e-3
    lambda_k: int = 6
    lambda_topk: int = 3
    lambda_p: float = 2.0
    lambda_sigma: Optional[float] = None
    normalise: bool = False
    sparsity_check: bool = False

    def with_lambda_graph(self, eps, k, topk, p, sigma_override=None):
        self.lambda_eps, self.lambda_k, self.lambda_topk, self.lambda_p, self.lambda_sigma = float(eps), int(k), int(topk), float(p), sigma_override
        return self

    def define_result_k(self):
        if self.lambda_k <= 5:
            self.lambda_topk = 3
        elif self.lambda_k < 10:
            self.lambda_topk = 4
        return self

    def graph_params(self) -> "GraphParams":
        self.define_result_k()
        return GraphParams(self.lambda_eps, self.lambda_k, self.lambda_topk, self.lambda_p, self.lambda_sigma, self.normalise, self.sparsity_check)


@dataclass
class GraphLaplacian:
    """src_legacy/graph.rs:127-136: `matrix` is the CSR Laplacian (here: a device handle plus lazily
    fetched host arrays), `nnodes` the item count of the original data, `init_data` the matrix the
    graph was built from."""
    matrix: Csr
    nnodes: int
    graph_params: GraphParams
    init_data: Optional[np.ndarray] = None
    energy: bool = False
    _host: Optional[tuple] = field(default=None, repr=False)

    def csr(self):
        if self._host is None:
            self._host = self.matrix.to_host()
        return self._host

    def shape(self):
        r, _ = self.matrix.shape
        return (r, r)

    def nnz(self):
        return self.matrix.shape[1]

    @staticmethod
    def sparsity(matrix: Csr):
        """graph.rs:626-632"""
        r, nnz = matrix.shape
        return 1.0 - nnz / float(r * r)

    def degrees(self):
        """graph.rs:353-373: the diagonal"""
        indptr, indices, data = self.csr()
        out = np.zeros(len(indptr) - 1)
        for r in range(len(out)):
            s, e = int(indptr[r]), int(indptr[r + 1])
            hit = np.nonzero(indices[s:e] == r)[0]
            if len(hit):
                out[r] = data[s + hit[0]]
        return out

    def multiply_vector(self, x):
        """graph.rs:464-501"""
        if len(x) != self.matrix.shape[0]:
            raise ValueError(f"Vector length {len(x)} must match number of nodes {self.matrix.shape[0]}")
        return self.matrix.spmv(x)

    def rayleigh_quotient(self, vector):
        """graph.rs:422-461"""
        if len(vector) != self.matrix.shape[0]:
            raise ValueError(f"Vector length {len(vector)} must match number of nodes {self.matrix.shape[0]}")
        return self.matrix.rayleigh_quotient(vector)

    def neighbors_of(self, i):
        """graph.rs:510-525: (j, w) with w = -L_ij > 0"""
        indptr, indices, data = self.csr()
        s, e = int(indptr[i]), int(indptr[i + 1])
        return [(int(j), -float(v)) for j, v in zip(indices[s:e], data[s:e]) if j != i and -v > 0.0]


def build_laplacian_matrix(transposed, params: GraphParams, n_items=None, energy=False, screen=SCREEN_AUTO,
                           ctx=None) -> GraphLaplacian:
    """src_legacy/laplacian.rs:122-180.  `transposed` has one row per graph node.  Runs through the
    one-shot C entry point (host buffers in; the CSR stays on the device)."""
    ctx = ctx or default_context()
    x = _ffi.f64(transposed)
    if x.ndim != 2:
        raise ValueError("items must be a 2-D matrix")
    n, d = x.shape
    gp = _ffi.GraphParamsC(float(params.eps), int(params.k), int(params.topk), float(params.p),
                           float(1.0 if params.sigma is None else params.sigma), int(params.normalise),
                           int(params.sparsity_check))
    h = C.c_void_p()
    ctx.check(lib().sfb_build_laplacian_matrix(ctx._h, _ffi.ptr(x), n, d, C.byref(gp), screen, C.byref(h)))
    # `let (d, n) = transposed.shape(); ... nnodes: n_items.unwrap_or(n)` (laplacian.rs:129,166-169): the COLUMN count
    return GraphLaplacian(matrix=Csr(ctx, h), nnodes=d if n_items is None else n_items, graph_params=params,
                          init_data=x, energy=energy)


class GraphFactory:
    @staticmethod
    def build_laplacian_matrix_from_k_cluster(clustered, eps, k, topk, p, sigma_override, normalise, sparsity_check,
                                              n_items, ctx=None) -> GraphLaplacian:
        """src_legacy/graph.rs:193-255: transposes (features become the nodes) and builds."""
        clustered = np.asarray(clustered, dtype=np.float64)
        if clustered.shape[0] > n_items:
            raise ValueError("clustered.shape().0 <= n_items")  # graph.rs:212
        params = GraphParams(eps, k, topk, p, sigma_override, normalise, sparsity_check)
        return build_laplacian_matrix(np.ascontiguousarray(clustered.T), params, n_items, False, ctx=ctx)


class TauMode:
    """src_legacy/taumode.rs:16-23."""

    def __init__(self, kind, value=0.0):
        self.kind, self.value = kind, value

    @staticmethod
    def Fixed(t):
        return TauMode(TAU_FIXED, t)

    @staticmethod
    def Percentile(p):
        return TauMode(TAU_PERCENTILE, p)

    def __repr__(self):
        return {TAU_FIXED: f"Fixed({self.value})", TAU_MEDIAN: "Median", TAU_MEAN: "Mean",
                TAU_PERCENTILE: f"Percentile({self.value})"}[self.kind]

    @staticmethod
    def compute_taumode_lambdas_parallel(items, gl: GraphLaplacian, taumode, ctx=None):
        """taumode.rs:117-214 + ArrowSpace::update_lambdas (core.rs:1427-1443): per-item synthetic
        lambda against the F x F Laplacian, then min-max normalised.  Returns the lambda vector."""
        ctx = ctx or gl.matrix.ctx
        x = _ffi.f64(items)
        n, f = x.shape
        out = np.empty(n, np.float64)
        ctx.check(lib().sfb_compute_taumode_lambdas(ctx._h, gl.matrix._h, _ffi.ptr(x), n, f, taumode.kind,
                                                    float(taumode.value), _ffi.ptr(out)))
        return out

    compute_taumode_lambdas = compute_taumode_lambdas_parallel  # taumode.rs:411-413


TauMode.Median = TauMode(TAU_MEDIAN)
TauMode.Mean = TauMode(TAU_MEAN)


class SfGrassSparsifier:
    """src_legacy/sparsification.rs:14-36."""

    def __init__(self):
        self.target_ratio = 0.5

    def with_target_ratio(self, ratio):
        self.target_ratio = min(max(ratio, 0.1), 1.0)
        return self

    def sparsify_graph(self, adj_rows, n_nodes, ctx=None):
        """adj_rows: list of lists of (j, w).  Returns the same shape."""
        ctx = ctx or default_context()
        k = max(1, max((len(r) for r in adj_rows), default=1))
        idx = np.full((n_nodes, k), IDX_NONE, np.uint32)
        w = np.zeros((n_nodes, k), np.float64)
        cnt = np.zeros(n_nodes, np.uint32)
        for i, r in enumerate(adj_rows):
            cnt[i] = len(r)
            for t, (j, wv) in enumerate(r):
                idx[i, t], w[i, t] = j, wv
        a = Adjacency.from_host(ctx, idx, w, cnt)
        a.sfgrass(self.target_ratio)
        idx, w, cnt = a.to_host()
        return [[(int(idx[i, t]), float(w[i, t])) for t in range(int(cnt[i]))] for i in range(n_nodes)]


# ---- successor Stage C (surfface-core/src/laplacian.rs) ---------------------------------------------------------
@dataclass
class LaplacianConfig:
    """surfface-core/src/laplacian.rs:49-77."""
    k_neighbors: int = 15
    variance_regularizer: float = 1e-6
    normalize: bool = True
    weight_threshold: float = 1e-9


@dataclass
class LaplacianOutput:
    """surfface-core/src/laplacian.rs:84-99 (`matrix` is a device CSR handle; values are f32-exact)."""
    matrix: Csr
    n_features: int
    nnz: int
    degrees: np.ndarray
    sparsity: float


def bc_adjacency(means, variances, k, reg=1e-6, thr=1e-9, ctx=None) -> Adjacency:
    """compute_bhattacharyya_weights (laplacian.rs:254-298): directed top-k Bhattacharyya affinities per feature."""
    ctx = ctx or default_context()
    m = np.ascontiguousarray(means, dtype=np.float32)
    v = np.ascontiguousarray(variances, dtype=np.float32)
    if m.ndim != 2 or m.shape != v.shape:
        raise ValueError("means and variances must be [C, F] arrays of the same shape")
    h = C.c_void_p()
    ctx.check(lib().sfb_bc_adjacency_build(ctx._h, _ffi.ptr(m), _ffi.ptr(v), m.shape[0], m.shape[1], int(k), float(reg),
                                           float(thr), C.byref(h)))
    return Adjacency(ctx, h)


class LaplacianStage:
    """LaplacianStage (surfface-core/src/laplacian.rs:101-219)."""

    def __init__(self, config: Optional[LaplacianConfig] = None):
        self.config = config or LaplacianConfig()

    @staticmethod
    def with_defaults():
        return LaplacianStage(LaplacianConfig())

    def execute(self, means, variances, ctx=None) -> LaplacianOutput:
        ctx = ctx or default_context()
        m = np.ascontiguousarray(means, dtype=np.float32)
        v = np.ascontiguousarray(variances, dtype=np.float32)
        if m.ndim != 2 or m.shape != v.shape:
            raise ValueError("means and variances must be [C, F] arrays of the same shape")
        c, f = m.shape
        cfg = _ffi.LaplacianConfigC(int(self.config.k_neighbors), float(self.config.variance_regularizer),
                                    int(self.config.normalize), float(self.config.weight_threshold))
        deg = np.empty(f, np.float32)
        h = C.c_void_p()
        ctx.check(lib().sfb_laplacian_stage_execute(ctx._h, _ffi.ptr(m), _ffi.ptr(v), c, f, C.byref(cfg), C.byref(h), _ffi.ptr(deg)))
        L = Csr(ctx, h)
        nnz = L.shape[1]
        return LaplacianOutput(matrix=L, n_features=f, nnz=nnz, degrees=deg, sparsity=1.0 - nnz / float(f * f))


def compute_tau_mode_gpu(laplacian: LaplacianOutput, data, n_items, n_features, ctx=None):
    """Stage D seam of the successor, compute_tau_mode_gpu(&LaplacianOutput, data: &[f32], n_items, n_features) -> Vec<f64>
    (surfface-core/src/spectral/bridge.rs:27-32): Rayleigh + Dirichlet per item in f32 semantics, widened to f64,
    not normalised.  `data` is the flat row-major f32 item matrix; it crosses PCIe as f32."""
    ctx = ctx or laplacian.matrix.ctx
    x = _ffi.f32(data).reshape(-1)
    if x.size != n_items * n_features:
        raise ValueError(f"data has {x.size} values, expected {n_items} x {n_features}")
    out = np.empty(n_items, np.float64)
    ctx.check(lib().sfb_compute_tau_mode_lambdas(ctx._h, laplacian.matrix._h, _ffi.ptr(x), n_items, n_features, _ffi.ptr(out)))
    return out


class CoreTauMode:
    """TauMode of the successor (surfface-core/src/taumode.rs:12-23): tau is resolved from the lambda DISTRIBUTION."""
    Median, Mean = (TAU_MEDIAN, 0.0), (TAU_MEAN, 0.0)

    @staticmethod
    def Fixed(t):
        return (TAU_FIXED, float(t))

    @staticmethod
    def Percentile(p):
        return (TAU_PERCENTILE, float(p))


def compute_tau(lambdas, mode=CoreTauMode.Median, ctx=None):
    """compute_tau(lambdas: &[f32], mode) -> f32 (surfface-core/src/taumode.rs:37-65)."""
    ctx = ctx or default_context()
    return ctx.compute_tau(lambdas, mode[0], mode[1])


def compute_jl_dimension(n_points, original_dim, epsilon, core=False):
    """reduction.rs:117-171 (core=True: surfface-core/src/clustering.rs:113-123)."""
    out = C.c_uint64()
    if lib().sfb_compute_jl_dimension(int(n_points), int(original_dim), float(epsilon), int(bool(core)), C.byref(out)) != 0:
        raise SfbError("sfb_compute_jl_dimension")
    return int(out.value)


class ImplicitProjection:
    """reduction.rs:202-248.  The reference keeps only the seed and re-draws the ChaCha8 StandardNormal stream per
    item; the device path wants the draws once, so the mirror carries them (`samples`: original_dim x reduced_dim in
    the reference's draw order -- the Rust wrapper fills them with the reference's own rand crates)."""

    def __init__(self, original_dim, reduced_dim, samples):
        self.original_dim, self.reduced_dim = int(original_dim), int(reduced_dim)
        self.samples = _ffi.f64(samples)
        if self.samples.shape != (self.original_dim, self.reduced_dim):
            raise ValueError("samples must be original_dim x reduced_dim")

    def get_reduced_dim(self):
        return self.reduced_dim

    def project(self, query, ctx=None):
        q = _ffi.f64(query).reshape(1, -1)
        ctx = ctx or default_context()
        return ctx.matrix(q[:, :self.original_dim]).project(self.samples).rows()[0]


def project_matrix(data, projection: ImplicitProjection, ctx=None) -> Matrix:
    """reduction.rs:175-200; `data` is a host array or a device Matrix; the result stays on the device."""
    ctx = ctx or default_context()
    m = data if isinstance(data, Matrix) else ctx.matrix(data)
    return m.project(projection.samples)


class SortedLambdas:
    """sorted_index.rs:8-57,59-79: items ordered by lambda (ties by the decimal string of the index); built by a
    device radix sort instead of N BTreeMap insertions."""

    def __init__(self):
        self.lambdas = np.empty(0)
        self.indices = np.empty(0, np.uint32)
        self.std_dev = 0.0

    def build_from(self, lambdas, ctx=None):
        ctx = ctx or default_context()
        lam = _ffi.f64(lambdas)
        n = lam.shape[0]
        self.lambdas, self.indices = np.empty(n, np.float64), np.empty(n, np.uint32)
        sd = C.c_double()
        ctx.check(lib().sfb_sorted_lambdas_build(ctx._h, _ffi.ptr(lam), n, _ffi.ptr(self.lambdas), _ffi.ptr(self.indices), C.byref(sd)))
        self.std_dev = sd.value
        return self

    def to_vec(self):
        return list(zip(self.lambdas.tolist(), self.indices.tolist()))

    def range_bylambda(self, lambda_q, k, p):
        """(idx, lambda) of the first k items with lambda in [q - band, q + band], band = std_dev / 2^p (:60-79)."""
        band = self.std_dev / 2.0 ** p
        lo = int(np.searchsorted(self.lambdas, lambda_q - band, "left"))
        hi = int(np.searchsorted(self.lambdas, lambda_q + band, "right"))
        hi = min(hi, lo + k)
        return list(zip(self.indices[lo:hi].tolist(), self.lambdas[lo:hi].tolist()))
