"""Size-independent properties of the CPU oracle (hypothesis; CPU only).  The same properties are what bench.py
--verify and the GPU tests lean on at sizes the oracle cannot brute-force: if they did not hold for the oracle they
would mean nothing for the CUDA path."""
import numpy as np
from hypothesis import given, settings, strategies as st

SET = settings(max_examples=25, deadline=None)


def _x(seed, m, kd):
    return np.random.default_rng(seed).normal(size=(m, kd))


@SET
@given(st.integers(0, 10_000), st.integers(8, 60), st.integers(2, 12), st.integers(1, 6), st.sampled_from([0, 1, 2]))
def test_knn_is_equivariant_under_row_permutation(oracle, seed, m, kd, k, metric):
    """Relabelling the rows relabels the neighbours: distances are identical bit for bit, and wherever a row's
    distances are all distinct (no tie to break by index) so are the neighbours."""
    x = _x(seed, m, kd)
    k = min(k, m - 1)
    perm = np.random.default_rng(seed + 1).permutation(m)
    idx, dist, cnt = oracle.knn(x, k, metric)
    pidx, pdist, pcnt = oracle.knn(x[perm], k, metric)
    inv = np.argsort(perm)
    assert np.array_equal(pcnt[inv], cnt) and np.array_equal(pdist[inv], dist)
    # a tie between the k-th neighbour and the first one left out is broken by index too: look one neighbour further
    k1 = min(k + 1, m - 1)
    _, d1, c1 = oracle.knn(x, k1, metric)
    distinct = np.array([len(set(d1[i, :c1[i]])) == c1[i] and (k1 > k or cnt[i] < k) for i in range(m)])
    mapped = perm[pidx[inv].astype(np.int64) % m]
    for i in np.nonzero(distinct)[0]:
        assert list(mapped[i, :cnt[i]]) == list(idx[i, :cnt[i]])


@SET
@given(st.integers(0, 10_000), st.integers(10, 80), st.integers(2, 10), st.integers(1, 8))
def test_laplacian_invariants_hold_for_any_input(oracle, seed, m, kd, k):
    """L = D - W: symmetric structure and values, zero row sums (to rounding), non-positive off-diagonals, a stored
    diagonal in every row, column indices ascending (laplacian.rs:351-419; test_laplacian.rs:52-154)."""
    x = _x(seed, m, kd)
    k = min(k, m - 1)
    a = oracle.build_adjacency(*oracle.knn(x, k, 0), 2.0, 1.0)
    ptr, ind, dat = oracle.laplacian(a[0], a[1], a[2])
    ptr = ptr.astype(np.int64)
    dense = np.zeros((m, m))
    for r in range(m):
        cols = ind[ptr[r]:ptr[r + 1]]
        assert np.all(np.diff(cols.astype(np.int64)) > 0) and r in cols
        dense[r, cols] = dat[ptr[r]:ptr[r + 1]]
    assert np.array_equal(dense, dense.T)
    assert np.all(np.abs(dense.sum(axis=1)) <= 1e-12 * np.maximum(1.0, np.abs(np.diag(dense))))
    off = dense - np.diag(np.diag(dense))
    assert np.all(off <= 0.0)
    assert len(ind) <= m * (2 * k + 1)


@SET
@given(st.integers(0, 10_000), st.integers(3, 40), st.integers(4, 24), st.floats(0.1, 50.0))
def test_lambda_is_scale_invariant_with_fixed_tau(oracle, seed, n, f, scale):
    """E and G are ratios: lambda(c x) = lambda(x) for a Fixed tau (test_taumode.rs:643-682, 1e-10)."""
    x = _x(seed, n, f)
    a = oracle.build_adjacency(*oracle.knn(oracle.transpose(x), min(3, f - 1), 0), 2.0, 1.0)
    ptr, ind, dat = oracle.laplacian(a[0], a[1], a[2])
    l1 = oracle.lambdas(ptr, ind, dat, x, tau_mode=oracle.TAU_FIXED, tau_value=0.5)
    l2 = oracle.lambdas(ptr, ind, dat, scale * x, tau_mode=oracle.TAU_FIXED, tau_value=0.5)
    np.testing.assert_allclose(l1, l2, rtol=1e-10, atol=1e-12)


@SET
@given(st.integers(0, 10_000), st.integers(1, 30), st.integers(2, 40), st.integers(1, 16))
def test_projection_is_linear_in_the_items(oracle, seed, n, f, r):
    """project(2^k x) = 2^k project(x) exactly; project(x + y) = project(x) + project(y) to rounding."""
    rng = np.random.default_rng(seed)
    x, y, s = rng.normal(size=(n, f)), rng.normal(size=(n, f)), rng.normal(size=(f, r))
    px = oracle.project_rows(x, s)
    assert np.array_equal(oracle.project_rows(8.0 * x, s), 8.0 * px)
    np.testing.assert_allclose(oracle.project_rows(x + y, s), px + oracle.project_rows(y, s), rtol=1e-9, atol=1e-12)


@SET
@given(st.integers(0, 10_000), st.integers(1, 400), st.integers(1, 4))
def test_sorted_lambdas_is_a_sorted_permutation(oracle, seed, n, decimals):
    lam = np.round(np.random.default_rng(seed).random(n), decimals)
    srt, idx, sd = oracle.sorted_lambdas(lam)
    assert sorted(idx.tolist()) == list(range(n)) and np.array_equal(srt, lam[idx]) and np.all(np.diff(srt) >= 0)
    for a, b in zip(range(n - 1), range(1, n)):   # ties by the decimal string of the index
        if srt[a] == srt[b]:
            assert str(int(idx[a])) < str(int(idx[b]))
    assert sd >= 0.0
