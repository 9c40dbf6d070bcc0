"""Oracle pieces added for full-size verification and the successor's seams (CPU only):
  * the block-streamed kNN / sampled column-Gram forms equal the whole-matrix oracle bit for bit;
  * compute_tau of surfface-core (taumode.rs:37-65) against values derived by hand from its code;
  * the portable f32 log / exp the Bhattacharyya kernel evaluates, against the host libm (what the reference calls)."""
import ctypes

import numpy as np
import pytest


@pytest.mark.parametrize("metric", [0, 1, 2])
def test_knn_stream_equals_whole_matrix_oracle(oracle, metric):
    rng = np.random.default_rng(metric)
    x = rng.normal(size=(2500, 24))
    x[100] = x[7]            # an exact duplicate: ties by index across block boundaries
    x[1300] = 0.0            # a zero row (cosine guard)
    q = np.array([7, 100, 1300, 0, 2499, 1023, 1024], dtype=np.uint64)
    want = oracle.knn(x, 9, metric, query_rows=q)
    for step in (2500, 1024, 333):
        s = oracle.KnnStream(x[q], q, 9, metric)
        for r0 in range(0, 2500, step):
            s.feed(x[r0:r0 + step], r0)
        got = s.finish()
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), step


def test_knn_stream_eps_and_short_lists(oracle):
    x = np.random.default_rng(3).normal(size=(40, 6))
    q = np.arange(40, dtype=np.uint64)
    want = oracle.knn(x, 5, 0, eps=0.3)
    s = oracle.KnnStream(x, q, 5, 0, eps=0.3)
    s.feed(x[:17], 0); s.feed(x[17:], 17)
    got = s.finish()
    assert all(np.array_equal(a, b) for a, b in zip(got, want)) and want[2].min() < 5


def test_cols_gram_stream_equals_feature_graph_oracle(oracle):
    x = np.random.default_rng(4).normal(size=(3001, 30)) + 0.2
    x[:, 11] = x[:, 4]       # duplicate feature
    x[:, 29] = 0.0           # zero feature
    want = oracle.knn(oracle.transpose(x), 6, oracle.METRIC_COSINE)
    cols = np.array([0, 4, 11, 29, 17], dtype=np.uint32)
    for step in (3001, 512, 7):
        g = oracle.ColsGramStream(30, cols, 6)
        for r0 in range(0, 3001, step):
            g.feed(x[r0:r0 + step])
        got = g.finish()
        assert np.array_equal(got[0], want[0][cols]) and np.array_equal(got[1], want[1][cols]) and np.array_equal(got[2], want[2][cols])


def test_compute_tau_core_follows_the_reference_code(oracle):
    f32 = np.float32
    lam = [0.5, 0.1, 0.4, 0.2, 0.3]
    assert oracle.compute_tau_core(lam, oracle.TAU_MEDIAN) == f32(0.3)            # sorted[5 / 2]
    assert oracle.compute_tau_core(lam[:4], oracle.TAU_MEDIAN) == f32(0.4)        # even: sorted[4 / 2], the UPPER median, no averaging
    assert oracle.compute_tau_core(lam, oracle.TAU_MEAN) == (f32(0.5) + f32(0.1) + f32(0.4) + f32(0.2) + f32(0.3)) / f32(5)
    assert oracle.compute_tau_core(lam, oracle.TAU_FIXED, 0.25) == f32(0.25)
    assert oracle.compute_tau_core(lam, oracle.TAU_FIXED, -3.0) == f32(1e-9)      # (-3).max(TAU_FLOOR)
    assert oracle.compute_tau_core(lam, oracle.TAU_FIXED, float("nan")) == f32(1e-9)
    assert oracle.compute_tau_core([], oracle.TAU_FIXED, 0.25) == f32(1e-9)       # finite.is_empty() comes first
    assert oracle.compute_tau_core([float("nan"), float("inf")], oracle.TAU_MEAN) == f32(1e-9)
    assert oracle.compute_tau_core(lam, oracle.TAU_PERCENTILE, 0.0) == f32(0.1)
    assert oracle.compute_tau_core(lam, oracle.TAU_PERCENTILE, 1.0) == f32(0.5)
    assert oracle.compute_tau_core(lam, oracle.TAU_PERCENTILE, 0.5) == f32(0.3)   # round(4 * 0.5) = 2
    assert oracle.compute_tau_core(lam, oracle.TAU_PERCENTILE, 0.625) == f32(0.4)  # round(2.5) = 3: half away from zero
    assert oracle.compute_tau_core(lam, oracle.TAU_PERCENTILE, 7.0) == f32(0.5)   # clamp(p, 0, 1)
    assert oracle.compute_tau_core([1e-12, 3e-12, 2e-12], oracle.TAU_MEDIAN) == f32(1e-9)   # floored
    assert oracle.compute_tau_core([0.2, float("nan"), 0.1, float("-inf"), 0.3], oracle.TAU_MEDIAN) == f32(0.2)


def test_portable_logf_expf_against_libm(oracle):
    """The reference's f32::ln / f32::exp are the platform libm's (glibc here); the portable forms the device evaluates
    are the correctly rounded values up to ~1e-9 ulp, so they may differ from glibc by at most 1 ulp, and rarely."""
    libm = ctypes.CDLL("libm.so.6")
    for fn in (libm.logf, libm.expf):
        fn.restype = ctypes.c_float
        fn.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(0)
    args = np.concatenate([np.exp(rng.uniform(-30, 30, 20000)), 1.0 + rng.uniform(-1e-3, 1e-3, 5000), [1.0, 2.0, 0.5, 1e-45, 3e38]]).astype(np.float32)
    bad = 0
    for a in args:
        got, want = np.float32(oracle.det_logf(a)), np.float32(libm.logf(float(a)))
        if got != want:
            bad += 1
            assert abs(int(got.view(np.int32)) - int(want.view(np.int32))) <= 1 or abs(float(got) - float(want)) <= 1e-45
    assert bad <= 0.002 * len(args), bad
    args = np.concatenate([rng.uniform(-104, 88, 20000), rng.uniform(-1, 0, 5000), [0.0, -0.0, -103.9, 88.7]]).astype(np.float32)
    bad = 0
    for a in args:
        got, want = np.float32(oracle.det_expf(a)), np.float32(libm.expf(float(a)))
        if got != want:
            bad += 1
            assert abs(int(got.view(np.int32)) - int(want.view(np.int32))) <= 1
    assert bad <= 0.002 * len(args), bad
    # special values (f32::ln / f32::exp conventions)
    assert oracle.det_logf(1.0) == 0.0 and oracle.det_logf(0.0) == -np.inf and np.isnan(oracle.det_logf(-1.0)) and oracle.det_logf(np.inf) == np.inf
    assert oracle.det_expf(0.0) == 1.0 and oracle.det_expf(-np.inf) == 0.0 and oracle.det_expf(np.inf) == np.inf and np.isnan(oracle.det_expf(np.nan))


def test_bc_portable_form_within_one_ulp_of_libm_form(oracle):
    rng = np.random.default_rng(1)
    means = rng.normal(size=(40, 60)).astype(np.float32)
    variances = rng.uniform(0.05, 2.0, size=(40, 60)).astype(np.float32)
    a, b = oracle.bc_matrix(means, variances, det=True), oracle.bc_matrix(means, variances, det=False)
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-30)
    ia, wa, ca = oracle.bc_knn(means, variances, 10, det=True)
    ib, wb, cb = oracle.bc_knn(means, variances, 10, det=False)
    assert np.array_equal(ca, cb) and np.mean(ia != ib) < 0.01
