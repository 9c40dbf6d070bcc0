"""ctypes binding of libsurfface_b200.so (the C ABI declared in include/surfface_b200.h).

There is no CPU fallback: a missing library raises ImportError, a missing GPU raises SfbError
from Context()."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsurfface_b200.so")

IDX_NONE = 0xFFFFFFFF
OK, EINVAL, ECUDA, ENOMEM, ENCCL, EUNCERTIFIED, EUNSUPPORTED = range(7)
STATUS_NAMES = ["SFB_OK", "SFB_EINVAL", "SFB_ECUDA", "SFB_ENOMEM", "SFB_ENCCL", "SFB_EUNCERTIFIED", "SFB_EUNSUPPORTED"]


class SfbError(RuntimeError):
    def __init__(self, status, message):
        self.status = status
        name = STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else str(status)
        super().__init__(f"{name}: {message}")


class KnnParams(C.Structure):
    _fields_ = [("metric", C.c_int32), ("k", C.c_uint32), ("eps", C.c_double), ("screen", C.c_int32),
                ("k_prime", C.c_uint32), ("q_begin", C.c_uint64), ("q_end", C.c_uint64),
                ("allow_fallback", C.c_int32)]


class KnnStats(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("rows_certified", C.c_uint64), ("rows_fallback", C.c_uint64),
                ("k_prime", C.c_uint32), ("screen_used", C.c_int32), ("ms_prepare", C.c_double),
                ("ms_screen", C.c_double), ("ms_rescore", C.c_double), ("ms_fallback", C.c_double),
                ("max_margin", C.c_double), ("rows_rescreened", C.c_uint64), ("ms_rescreen", C.c_double),
                ("candidates_rescored", C.c_uint64)]


class AdjParams(C.Structure):
    _fields_ = [("p", C.c_double), ("sigma", C.c_double), ("sparsify", C.c_int32)]


class LapParams(C.Structure):
    _fields_ = [("normalised", C.c_int32), ("weight_threshold", C.c_double)]


class LambdaParams(C.Structure):
    _fields_ = [("variant", C.c_int32), ("tau_mode", C.c_int32), ("tau_value", C.c_double),
                ("normalise_minmax", C.c_int32)]


class GraphParamsC(C.Structure):
    _fields_ = [("eps", C.c_double), ("k", C.c_uint32), ("topk", C.c_uint32), ("p", C.c_double),
                ("sigma", C.c_double), ("normalise", C.c_int32), ("sparsity_check", C.c_int32)]


class LaplacianConfigC(C.Structure):
    _fields_ = [("k_neighbors", C.c_uint32), ("variance_regularizer", C.c_float), ("normalize", C.c_int32),
                ("weight_threshold", C.c_float)]


class StageTimes(C.Structure):
    _fields_ = [("ms_h2d", C.c_double), ("ms_knn", C.c_double), ("ms_adjacency", C.c_double),
                ("ms_laplacian", C.c_double), ("ms_lambda", C.c_double), ("ms_d2h", C.c_double),
                ("kernel_launches", C.c_uint64), ("ms_lambda_kernel", C.c_double), ("ms_diffuse", C.c_double), ("ms_comm", C.c_double)]


# every symbol include/surfface_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
SYMBOLS = {
    "sfb_abi_version": (C.c_int32, []),
    "sfb_ctx_create": (C.c_int32, [C.c_int32, _PP]),
    "sfb_ctx_destroy": (None, [_P]),
    "sfb_last_error": (C.c_char_p, [_P]),
    "sfb_device_info": (C.c_int32, [_P, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "sfb_synchronize": (C.c_int32, [_P]),
    "sfb_pinned_alloc": (C.c_int32, [_P, C.c_uint64, _PP]),
    "sfb_pinned_free": (None, [_P]),
    "sfb_mat_from_host": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint32, _PP]),
    "sfb_mat_from_host_f32": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint32, _PP]),
    "sfb_mat_clone": (C.c_int32, [_P, _P, _PP]),
    "sfb_mat_generate": (C.c_int32, [_P, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, _PP]),
    "sfb_mat_transpose": (C.c_int32, [_P, _P, _PP]),
    "sfb_mat_view_rows": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint64, _PP]),
    "sfb_mat_allgather_rows": (C.c_int32, [_P, _P, C.c_uint64, _PP]),
    "sfb_mat_shape": (C.c_int32, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "sfb_mat_copy_rows": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "sfb_mat_free": (None, [_P]),
    "sfb_knn_build": (C.c_int32, [_P, _P, C.POINTER(KnnParams), _PP]),
    "sfb_knn_build_sharded": (C.c_int32, [_P, _P, C.POINTER(KnnParams), _PP]),
    "sfb_knn_build_columns": (C.c_int32, [_P, _P, C.POINTER(KnnParams), _PP]),
    "sfb_knn_build_columns_sharded": (C.c_int32, [_P, _P, C.POINTER(KnnParams), _PP]),
    "sfb_knn_build_columns_begin": (C.c_int32, [_P, _P, C.POINTER(KnnParams), C.c_int32, _PP]),
    "sfb_knn_build_columns_end": (C.c_int32, [_P, _P, _PP]),
    "sfb_knn_shape": (C.c_int32, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "sfb_knn_copy": (C.c_int32, [_P, _P, _P, _P, _P]),
    "sfb_knn_stats_get": (C.c_int32, [_P, C.POINTER(KnnStats)]),
    "sfb_knn_from_host": (C.c_int32, [_P, _P, _P, _P, C.c_uint64, C.c_uint32, _PP]),
    "sfb_knn_free": (None, [_P]),
    "sfb_adjacency_build": (C.c_int32, [_P, _P, C.POINTER(AdjParams), _PP, C.POINTER(C.c_int32)]),
    "sfb_sparsify_sfgrass": (C.c_int32, [_P, _P, C.c_double, C.POINTER(C.c_int32)]),
    "sfb_adj_shape": (C.c_int32, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "sfb_adj_copy": (C.c_int32, [_P, _P, _P, _P, _P]),
    "sfb_adj_from_host": (C.c_int32, [_P, _P, _P, _P, C.c_uint64, C.c_uint32, _PP]),
    "sfb_adj_free": (None, [_P]),
    "sfb_laplacian_build": (C.c_int32, [_P, _P, C.POINTER(LapParams), _PP]),
    "sfb_laplacian_build_rows": (C.c_int32, [_P, _P, C.POINTER(LapParams), C.c_uint64, C.c_uint64, _PP]),
    "sfb_csr_shape": (C.c_int32, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "sfb_csr_copy": (C.c_int32, [_P, _P, _P, _P, _P]),
    "sfb_csr_from_host": (C.c_int32, [_P, C.c_uint64, _P, _P, _P, _PP]),
    "sfb_csr_free": (None, [_P]),
    "sfb_spmv": (C.c_int32, [_P, _P, _P, _P]),
    "sfb_rayleigh_quotient": (C.c_int32, [_P, _P, _P, C.POINTER(C.c_double)]),
    "sfb_lambda": (C.c_int32, [_P, _P, _P, C.POINTER(LambdaParams), _P, _P, _P]),
    "sfb_lambda_projected": (C.c_int32, [_P, _P, _P, _P, C.POINTER(LambdaParams), _P, _P, _P]),
    "sfb_compute_tau_mode_lambdas": (C.c_int32, [_P, _P, _P, C.c_uint64, C.c_uint32, _P]),
    "sfb_compute_tau": (C.c_int32, [_P, _P, C.c_uint64, C.c_int32, C.c_float, C.POINTER(C.c_float)]),
    "sfb_diffuse": (C.c_int32, [_P, _P, _P, C.c_double, C.c_uint32]),
    "sfb_build_laplacian_matrix": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint32, C.POINTER(GraphParamsC), C.c_int32, _PP]),
    "sfb_compute_taumode_lambdas": (C.c_int32, [_P, _P, _P, C.c_uint64, C.c_uint32, C.c_int32, C.c_double, _P]),
    "sfb_debug_screen_tile": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, _P, _P, _P,
                                          C.POINTER(C.c_uint32), C.POINTER(C.c_double)]),
    "sfb_bc_adjacency_build": (C.c_int32, [_P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float, _PP]),
    "sfb_laplacian_stage_execute": (C.c_int32, [_P, _P, _P, C.c_uint32, C.c_uint32, C.POINTER(LaplacianConfigC), _PP, _P]),
    "sfb_map_items_to_subcentroids": (C.c_int32, [_P, _P, _P, _P, _P, C.c_double, _P, _P, _P]),
    "sfb_project_rows": (C.c_int32, [_P, _P, _P, C.c_uint32, C.c_int32, _PP]),
    "sfb_compute_jl_dimension": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_double, C.c_int32, C.POINTER(C.c_uint64)]),
    "sfb_sorted_lambdas_build": (C.c_int32, [_P, _P, C.c_uint64, _P, _P, C.POINTER(C.c_double)]),
    "sfb_timings": (C.c_int32, [_P, C.POINTER(StageTimes)]),
    "sfb_timings_reset": (C.c_int32, [_P]),
    "sfb_timer_start": (C.c_int32, [_P]),
    "sfb_timer_stop": (C.c_int32, [_P, C.POINTER(C.c_double)]),
    "sfb_comm_unique_id": (C.c_int32, [_P]),
    "sfb_comm_init": (C.c_int32, [_P, _P, C.c_int32, C.c_int32]),
    "sfb_knn_allgather": (C.c_int32, [_P, _P, C.c_uint64, _PP]),
    "sfb_lambda_allgather": (C.c_int32, [_P, _P, _P, C.c_uint64, C.c_uint64, C.POINTER(LambdaParams), _P, _P]),
    "sfb_comm_barrier": (C.c_int32, [_P]),
}

_lib = None


def lib():
    """Load the CUDA library.  Raises ImportError when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). surfface_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the header and the library drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)
