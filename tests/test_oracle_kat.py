"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md section 8c).  CPU only."""
import math

import numpy as np
import pytest

FLOOR = 1e-10  # src_legacy/taumode.rs:25


# ---- select_tau table: src_legacy/tests/test_taumode.rs:14-160 --------------------------------
def test_select_tau_fixed(oracle):
    e = [0.1, 0.5, 1.0]
    assert oracle.select_tau(e, oracle.TAU_FIXED, 0.3) == 0.3
    for bad in (-0.1, 0.0, float("nan"), float("inf")):
        assert oracle.select_tau(e, oracle.TAU_FIXED, bad) == FLOOR


def test_select_tau_mean(oracle):
    assert abs(oracle.select_tau([1.0, 2.0, 3.0], oracle.TAU_MEAN) - 2.0) < 1e-12
    assert abs(oracle.select_tau([1.0, math.nan, 3.0, math.inf, 2.0], oracle.TAU_MEAN) - 2.0) < 1e-12
    assert oracle.select_tau([math.nan, math.inf, -math.inf], oracle.TAU_MEAN) == FLOOR
    assert oracle.select_tau([], oracle.TAU_MEAN) == FLOOR


def test_select_tau_median(oracle):
    assert oracle.select_tau([3.0, 1.0, 2.0], oracle.TAU_MEDIAN) == 2.0
    assert abs(oracle.select_tau([1.0, 2.0, 3.0, 4.0], oracle.TAU_MEDIAN) - 2.5) < 1e-12
    assert oracle.select_tau([5.0], oracle.TAU_MEDIAN) == 5.0
    assert oracle.select_tau([math.nan, 1.0, 3.0, math.inf, 2.0], oracle.TAU_MEDIAN) == 2.0
    assert oracle.select_tau([math.nan, math.inf], oracle.TAU_MEDIAN) == FLOOR
    assert oracle.select_tau([], oracle.TAU_MEDIAN) == FLOOR


def test_select_tau_percentile(oracle):
    e = [1.0, 2.0, 3.0, 4.0, 5.0]
    P = oracle.TAU_PERCENTILE
    assert oracle.select_tau(e, P, 0.0) == 1.0
    assert oracle.select_tau(e, P, 1.0) == 5.0
    assert oracle.select_tau(e, P, 0.5) == 3.0
    assert oracle.select_tau(e, P, -0.1) == 1.0
    assert oracle.select_tau(e, P, 1.5) == 5.0
    assert oracle.select_tau([], P, 0.5) == FLOOR


def test_select_tau_floor(oracle):
    assert oracle.select_tau([FLOOR * 2.0], oracle.TAU_MEAN) == FLOOR * 2.0
    assert oracle.select_tau([FLOOR / 2.0], oracle.TAU_MEAN) == FLOOR
    assert oracle.select_tau([0.0], oracle.TAU_MEAN) == FLOOR


# ---- pure-Python restatement of test_helpers.rs:73-170 for tiny cases -------------------------
def helper_adjacency(items, eps, topk, p, sigma):
    n = len(items)
    norms = [math.sqrt(sum(v * v for v in it)) for it in items]
    adj = [dict() for _ in range(n)]
    for i in range(n):
        cand = []
        for j in range(n):
            if i == j:
                continue
            denom = norms[i] * norms[j]
            if denom > 1e-12:
                dot = 0.0
                for a, b in zip(items[i], items[j]):
                    dot += a * b
                cos = min(max(dot / denom, -1.0), 1.0)
            else:
                cos = 0.0
            dist = 1.0 - max(cos, 0.0)
            if dist <= eps:
                w = 1.0 / (1.0 + (dist / sigma) ** p)
                if w > 1e-12:
                    cand.append((dist, j, w))
        cand.sort()
        for dist, j, w in cand[:topk]:
            adj[i][j] = w
    for i in range(n):
        for j in list(adj[i].keys()):
            w = adj[i][j]
            if w > 1e-12 and adj[j].get(i, 0.0) < 1e-12:
                adj[j][i] = w
    return adj


def oracle_graph(oracle, items, eps, topk, p, sigma):
    x = np.array(items, dtype=np.float64)
    idx, dist, cnt = oracle.knn(x, topk, oracle.METRIC_COSINE, eps)
    a_idx, a_w, a_cnt, _ = oracle.build_adjacency(idx, dist, cnt, p, sigma)
    indptr, indices, data = oracle.laplacian(a_idx, a_w, a_cnt)
    n = len(items)
    dense = np.zeros((n, n))
    for r in range(n):
        for e in range(int(indptr[r]), int(indptr[r + 1])):
            dense[r, indices[e]] = data[e]
    return dense, (indptr, indices, data)


# ---- L = D - A on 3 points: src_legacy/tests/test_laplacian.rs:655-786 ------------------------
def test_with_adjacency_output(oracle):
    items = [[1.0, 0.0], [0.9, 0.1], [0.0, 1.0]]
    eps, topk, p, sigma = 0.5, 1, 1.0, 0.2
    adj = helper_adjacency(items, eps, topk, p, sigma)
    L, (indptr, indices, data) = oracle_graph(oracle, items, eps, topk, p, sigma)
    # hand-derived: only edge 0-1, d = 1 - 0.9/sqrt(0.82)
    d01 = 1.0 - 0.9 / math.sqrt(0.9 * 0.9 + 0.1 * 0.1)
    w01 = 1.0 / (1.0 + d01 / 0.2)
    assert adj[0] == {1: pytest.approx(w01, rel=1e-15)} and adj[2] == {}
    for i in range(3):
        deg = sum(adj[i].values())
        assert abs(L[i, i] - deg) < 1e-10
        for j in range(3):
            if i != j:
                assert abs(L[i, j] + adj[i].get(j, 0.0)) < 1e-10
    assert np.allclose(L, L.T, atol=1e-10)
    # diagonal always stored, even 0 (laplacian.rs:372): 3 diagonals + 2 off-diagonals
    assert len(data) == 5 and int(indptr[3]) - int(indptr[2]) == 1 and data[-1] == 0.0
    assert L[0, 1] == -w01


# ---- cosine ordering: src_legacy/tests/test_laplacian.rs:156-213 ------------------------------
def test_cosine_similarity_based_construction(oracle):
    items = [[1.0, 0.0], [0.707, 0.707], [0.0, 1.0], [-1.0, 0.0]]
    eps, topk, p, sigma = 2.0, 2, 1.0, 0.5
    adj = helper_adjacency(items, eps, topk, p, sigma)
    L, _ = oracle_graph(oracle, items, eps, topk, p, sigma)
    for i in range(4):
        for j in range(4):
            if i != j:
                assert L[i, j] == -adj[i].get(j, 0.0)
    a01, a02, a03 = -L[0, 1], -L[0, 2], -L[0, 3]
    assert a01 > a02
    assert a02 >= a03
    assert a01 > a02 + 0.05


# ---- Laplacian invariants: src_legacy/tests/test_laplacian.rs:52-154 --------------------------
def test_laplacian_invariants(oracle):
    rng = np.random.default_rng(3)
    x = rng.normal(size=(60, 12))
    k = 5
    idx, dist, cnt = oracle.knn(x, k, oracle.METRIC_COSINE, np.inf)
    a_idx, a_w, a_cnt, applied = oracle.build_adjacency(idx, dist, cnt, 2.0, 1.0)
    assert not applied  # mean degree 5 <= 10
    indptr, indices, data = oracle.laplacian(a_idx, a_w, a_cnt)
    n = x.shape[0]
    L = np.zeros((n, n))
    for r in range(n):
        cols = indices[int(indptr[r]):int(indptr[r + 1])]
        assert np.all(np.diff(cols.astype(np.int64)) > 0)  # sorted, unique
        L[r, cols] = data[int(indptr[r]):int(indptr[r + 1])]
    assert np.allclose(L.sum(axis=1), 0.0, atol=1e-12)
    assert np.array_equal(L, L.T)
    assert np.all(np.diag(L) >= 0.0)
    assert np.all(L[~np.eye(n, dtype=bool)] <= 0.0)
    assert len(data) <= n * (2 * k + 1)


# ---- Rayleigh / chain graph: surfface-core/src/tests/test_spectral.rs:102-144,187-251 ----------
def csr_from_dense(d):
    d = np.asarray(d, dtype=np.float64)
    indptr, indices, data = [0], [], []
    for r in range(d.shape[0]):
        for c in range(d.shape[1]):
            if d[r, c] != 0.0:
                indices.append(c)
                data.append(d[r, c])
        indptr.append(len(indices))
    return (np.array(indptr, np.uint64), np.array(indices, np.uint32), np.array(data, np.float64))


def test_rayleigh_eigenvector(oracle):
    L = csr_from_dense([[1.0, -1.0], [-1.0, 1.0]])
    assert oracle.rayleigh(*L, np.array([1.0, 1.0])) == 0.0
    lam = oracle.lambdas(*L, np.array([[1.0, 1.0]]), oracle.LAMBDA_CORE_F32SEM)
    assert abs(lam[0]) < 1e-5


def test_dispersion_uniform(oracle):
    L = csr_from_dense([[1.0, -0.5], [-0.5, 1.0]])
    x = np.array([[1.0, 1.0], [1.0, 1.0]])
    for variant in (oracle.LAMBDA_LEGACY_TAUMODE, oracle.LAMBDA_ENERGY_NODE, oracle.LAMBDA_CORE_F32SEM):
        _, _, g = oracle.lambdas(*L, x, variant, with_parts=True)
        assert np.all(np.abs(g) < 1e-5)


def test_chain_graph_lambda(oracle):
    L = csr_from_dense([[1, -1, 0], [-1, 2, -1], [0, -1, 1]])
    x = np.array([[1.0, 1.0, 1.0], [1.0, 0.0, -1.0]])
    for variant in (oracle.LAMBDA_LEGACY_TAUMODE, oracle.LAMBDA_ENERGY_NODE, oracle.LAMBDA_CORE_F32SEM):
        lam = oracle.lambdas(*L, x, variant, oracle.TAU_FIXED, 0.5)
        assert abs(lam[0]) < 1e-5 and lam[1] > lam[0] and np.all(np.isfinite(lam))
    # hand value: L x = [1, 0, -1], x.Lx = 2, x.x = 2 => E = 1
    _, e, g = oracle.lambdas(*L, x, oracle.LAMBDA_ENERGY_NODE, with_parts=True)
    assert e[1] == 1.0
    # two unit edges with (x_r - x_c)^2 = 1 each => shares 1/2, 1/2 => G = 1/2 (upper triangle)
    assert g[1] == 0.5
    # taumode counts both triangles: four shares of 1/4 => 1/4 (exactly half of the energy-node value)
    _, e2, g2 = oracle.lambdas(*L, x, oracle.LAMBDA_LEGACY_TAUMODE, oracle.TAU_FIXED, 0.5, with_parts=True)
    assert g2[1] == 0.25 and e2[1] == 1.0
    lam = oracle.lambdas(*L, x, oracle.LAMBDA_LEGACY_TAUMODE, oracle.TAU_FIXED, 0.5)
    assert lam[1] == 0.5 * (1.0 / 1.5) + 0.5 * 0.25


# ---- scale invariance: src_legacy/tests/test_taumode.rs:643-682 -------------------------------
def test_scale_invariance(oracle):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(40, 4))
    idx, dist, cnt = oracle.knn(oracle.transpose(x), 2, oracle.METRIC_COSINE, np.inf)
    a = oracle.build_adjacency(idx, dist, cnt, 2.0, 1.0)
    L = oracle.laplacian(*a[:3])
    v = np.array([[1.0, 2.0, 3.0, 1.0], [2.0, 4.0, 6.0, 2.0]])
    lam = oracle.lambdas(*L, v, oracle.LAMBDA_LEGACY_TAUMODE, oracle.TAU_FIXED, 0.5)
    assert abs(lam[0] - lam[1]) <= 1e-10 * max(abs(lam[0]), 1.0)


def test_zero_vector_lambda(oracle):
    L = csr_from_dense([[1, -1, 0], [-1, 2, -1], [0, -1, 1]])
    lam = oracle.lambdas(*L, np.array([[0.0, 1e-11, -1e-11]]), oracle.LAMBDA_LEGACY_TAUMODE)
    assert lam[0] == 0.0  # taumode.rs:268-274


# ---- normalisation: src_legacy/core.rs:1341-1355 ----------------------------------------------
def test_normalise_lambdas(oracle):
    lam, stats = oracle.normalise_lambdas([0.2, 0.5, 0.3])
    assert stats[0] == 0.2 and stats[1] == 0.5 and lam[0] == 0.0 and lam[1] == 1.0
    lam, stats = oracle.normalise_lambdas([-2.0, -1.0])  # max fold starts at 0.0
    assert stats[1] == 0.0 and stats[2] == 2.0 and lam[1] == 0.5
    lam, stats = oracle.normalise_lambdas([0.3, 0.3])
    assert stats[2] == 1e-9 and lam[0] == 0.0  # range floor


# ---- distances: surfface-core/src/tests/test_distance.rs:254,266,284 --------------------------
def test_distance_kats(oracle):
    x = np.array([[0.0, 0.0], [3.0, 4.0]])
    _, d, _ = oracle.knn(x, 1, oracle.METRIC_L2)
    assert d[0, 0] == 5.0
    _, d, _ = oracle.knn(x, 1, oracle.METRIC_L2SQ)
    assert d[0, 0] == 25.0
    x = np.array([[1.0, 0.0], [2.0, 0.0], [0.0, 3.0]])
    idx, d, _ = oracle.knn(x, 2, oracle.METRIC_COSINE)
    assert idx[0, 0] == 1 and d[0, 0] == 0.0 and d[0, 1] == 1.0  # parallel -> cos 1, orthogonal -> cos 0


# ---- ties by index, eps filter, ragged rows ----------------------------------------------------
def test_knn_ties_and_eps(oracle):
    x = np.array([[1.0, 0.0], [1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])
    idx, d, cnt = oracle.knn(x, 3, oracle.METRIC_COSINE)
    assert list(idx[0]) == [1, 2, 3] and list(d[0]) == [0.0, 0.0, 1.0]
    assert list(idx[4]) == [0, 1, 2] and list(d[4]) == [1.0, 1.0, 1.0]  # zero row: denom guard -> cos 0
    idx, d, cnt = oracle.knn(x, 3, oracle.METRIC_COSINE, eps=0.5)
    assert list(cnt) == [2, 2, 2, 0, 0]
    assert idx[3, 0] == oracle.IDX_NONE and math.isinf(d[3, 0])


# ---- sparsification: laplacian.rs:258-282, sparsification.rs:32-113 ---------------------------
def test_sparsify_rules(oracle):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(80, 6))
    idx, dist, cnt = oracle.knn(x, 12, oracle.METRIC_COSINE)
    a_idx, a_w, a_cnt, applied = oracle.build_adjacency(idx, dist, cnt, 2.0, 1.0)
    assert applied and np.all(a_cnt == 6)  # mean degree 12 > 10: keep len/2
    full = oracle.build_adjacency(idx, dist, cnt, 2.0, 1.0, force_sparsify=0)
    assert np.all(full[2] == 12)
    s_idx, s_w, s_cnt, applied = oracle.sfgrass(*full[:3], ratio=0.3)
    assert applied and np.all(s_cnt == math.ceil(12 * 0.3))
    # kept edges are the top scores with (score desc, j asc)
    deg = full[2]
    for i in (0, 17, 79):
        sc = [(-full[1][i, t] * math.sqrt(float(int(deg[i]) * int(deg[full[0][i, t]]))), int(full[0][i, t])) for t in range(12)]
        sc.sort()
        assert [j for _, j in sc[:4]] == list(s_idx[i, :4])
    sparse = oracle.sfgrass(*oracle.build_adjacency(*oracle.knn(x, 5), 2.0, 1.0)[:3])
    assert not sparse[3]  # mean degree 5 < 10: skipped


def _ragged(rows, k):
    m = len(rows)
    idx = np.full((m, k), 0xFFFFFFFF, np.uint32); w = np.zeros((m, k)); cnt = np.zeros(m, np.uint32)
    for i, r in enumerate(rows):
        cnt[i] = len(r)
        for t, (j, v) in enumerate(r):
            idx[i, t], w[i, t] = j, v
    return idx, w, cnt


def test_sfgrass_reference_cases(oracle):
    """src_legacy/tests/test_sparsification.rs:3-46 on ragged rows: the 3-node graph is below the mean-degree
    gate (every row survives, non-empty); the 50-node (i + j) % 3 == 0 graph loses edges, each row keeping
    clamp(ceil(len * 0.5), 1, len) of them."""
    basic = [[(1, 1.0), (2, 0.5)], [(0, 1.0), (2, 0.8)], [(0, 0.5), (1, 0.8)]]
    idx, w, cnt, applied = oracle.sfgrass(*_ragged(basic, 2))
    assert not applied and list(cnt) == [2, 2, 2] and idx[0].tolist() == [1, 2]
    n = 50
    rows = [[(j, 1.0 / (1.0 + abs(i - j))) for j in range(n) if i != j and (i + j) % 3 == 0] for i in range(n)]
    k = max(len(r) for r in rows)
    idx, w, cnt, applied = oracle.sfgrass(*_ragged(rows, k))
    assert applied and int(cnt.sum()) < sum(len(r) for r in rows)
    assert list(cnt) == [min(max(math.ceil(len(r) * 0.5), 1), len(r)) for r in rows]
    # the survivors of row 4 are its best (w * sqrt(len_i * len_j)) scores, ties by index
    i = 4
    sc = sorted((-v * math.sqrt(float(len(rows[i]) * len(rows[j]))), j) for j, v in rows[i])
    assert [j for _, j in sc[:cnt[i]]] == idx[i, :cnt[i]].tolist()


# ---- diffusion: energymaps.rs:520-546 ---------------------------------------------------------
def test_diffusion(oracle):
    L = csr_from_dense([[1, -1, 0], [-1, 2, -1], [0, -1, 1]])
    x = np.array([[1.0, 0.0, -1.0], [2.0, 2.0, 2.0]])
    y = oracle.diffuse(*L, x, 0.1, 1)
    assert np.array_equal(y[0], np.array([1.0 - 0.1 * 1.0, 0.0, -1.0 + 0.1 * 1.0]))
    assert np.array_equal(y[1], x[1])
    y4 = oracle.diffuse(*L, x, 0.1, 4)
    z = x.copy()
    for _ in range(4):
        z = oracle.diffuse(*L, z, 0.1, 1)
    assert np.array_equal(y4, z)


# ---- synthetic generator sanity ---------------------------------------------------------------
def test_generator_statistics(oracle):
    x = oracle.generate_rows(0, 42, 0, 4000, 64)
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    from scipy import stats
    assert stats.kstest(x.ravel()[:50000], "norm").pvalue > 1e-3
    # counter-based: any row range is reproducible
    y = oracle.generate_rows(0, 42, 1000, 10, 64)
    assert np.array_equal(x[1000:1010], y)
    c = oracle.generate_rows(1, 7, 0, 2000, 32, n_centres=8, noise=0.3)
    assert c.shape == (2000, 32) and np.all(np.isfinite(c))


# ---- successor Stage C: Bhattacharyya coefficient (surfface-core/src/distance.rs:260-290) --------------------
def test_bhattacharyya_kats(oracle):
    """Identical distributions => BC = 1 (tests/test_distance.rs:10-25); unit variances, mean gap d => exp(-d^2/8);
    equal means, variances (1, 4) => sqrt(2*1*2/(1+4)) per centroid; symmetric; decays with separation
    (tests/test_distance.rs:412-470); variances below the floor are clamped to it."""
    m = np.array([[0.0, 0.0, 1.0, 3.0]], np.float32)
    v = np.array([[1.0, 1.0, 1.0, 1.0]], np.float32)
    assert oracle.bc(m, v, 0, 1) == 1.0
    assert oracle.bc(m, v, 0, 2) == pytest.approx(math.exp(-1.0 / 8.0), rel=1e-6)
    assert oracle.bc(m, v, 0, 3) == pytest.approx(math.exp(-9.0 / 8.0), rel=1e-6)
    assert oracle.bc(m, v, 0, 2) == oracle.bc(m, v, 2, 0)
    assert oracle.bc(m, v, 0, 1) > oracle.bc(m, v, 0, 2) > oracle.bc(m, v, 0, 3)
    m2 = np.zeros((3, 2), np.float32)
    v2 = np.array([[1.0, 4.0]] * 3, np.float32)
    assert oracle.bc(m2, v2, 0, 1) == pytest.approx((2.0 * 1.0 * 2.0 / 5.0) ** 1.5, rel=1e-6)   # (2 s_i s_j / (v_i + v_j))^(C/2)
    v3 = np.array([[0.0, 1e-9]], np.float32)
    assert oracle.bc(np.zeros((1, 2), np.float32), v3, 0, 1, reg=1e-6) == 1.0


def test_bhattacharyya_knn_order_and_threshold(oracle):
    """Top-k by (BC desc, j asc), k.min(F-1), BC <= threshold dropped (laplacian.rs:254-298)."""
    m = np.array([[0.0, 0.5, 0.5, 40.0, 0.1]], np.float32)
    v = np.ones((1, 5), np.float32)
    idx, w, cnt = oracle.bc_knn(m, v, 3, thr=1e-9)
    assert list(idx[0]) == [4, 1, 2] and cnt[0] == 3          # 1 and 2 tie: index order
    assert cnt[3] == 0 and np.all(idx[3] == oracle.IDX_NONE)  # BC(3, .) ~ exp(-200) underflows below the threshold
    idx, w, cnt = oracle.bc_knn(m, v, 10)
    assert cnt[0] == 3 and idx.shape == (5, 10)               # feature 3 is out of reach of everyone


# ---- JL projection ahead of lambda: src_legacy/tests/test_reduction.rs -------------------------------------------
def test_jl_dimension_table(oracle):
    """Every exact value the reference asserts (test_reduction.rs:193-229,233-243,317-357,453-475,525-539,556-562)."""
    jl = oracle.jl_dimension
    assert [jl(100, 16, 0.3), jl(1000, 8, 0.1), jl(50, 31, 0.2), jl(10, 1, 0.5)] == [16, 8, 31, 1]
    assert jl(10, 100, 0.3) == 100 and jl(10, 50, 0.3) == 50
    assert 148 <= jl(100, 200, 0.5) <= 149
    assert jl(2, 1000, 0.9) == 32 and jl(2, 20, 0.9) == 20
    assert jl(1000, 512, 0.1) == 512
    assert 921 <= jl(100, 2000, 0.2) <= 923
    bound = lambda n, e: math.ceil(8.0 * math.log(n) / (e * e))
    assert jl(10_000, 5_000, 0.3) == math.ceil(bound(10_000, 0.3) * 1.2)
    assert jl(5_000, 3_000, 0.3) == min(math.ceil(bound(5_000, 0.3) * 1.2), 3_000)
    assert jl(1000, 1500, 0.15) == min(max(bound(1000, 0.15), 32), 1500)
    assert jl(1, 100, 0.1) == 32 and jl(1, 10, 0.1) == 10
    assert jl(100, 2048, 0.3) == min(max(bound(100, 0.3), 32), 2048)
    assert jl(100, 2049, 0.3) == min(max(math.ceil(bound(100, 0.3) * 1.2), 32), 2049)
    assert 800 <= jl(100, 100_000, 0.3) <= 850                      # 2.0x buffer tier
    for n, d, e in ((500, 3000, 0.3), (100, 100_000, 0.3), (10, 50_000, 0.3)):   # :255-315, within +-2 of the tier
        tier = 1.2 if d / bound(n, e) < 10 else (1.5 if d / bound(n, e) < 100 else 2.0)
        assert abs(jl(n, d, e) - math.ceil(bound(n, e) * tier)) <= 2
    assert jl(10, 10_000, 0.3) < jl(100, 10_000, 0.3)
    # successor (surfface-core/src/clustering.rs:113-123): plain clamp, f32
    assert oracle.jl_dimension(100, 16, 0.3, core=True) == 16
    assert oracle.jl_dimension(1000, 4096, 0.3, core=True) == math.ceil(np.float32(8.0) * np.log(np.float32(1000)) / (np.float32(0.3) * np.float32(0.3)))


def test_projection_kats(oracle):
    """Zero vector -> zeros (test_reduction.rs:60-69), project(2x) = 2 project(x) (:72-93), norm roughly preserved
    (:96-109), rows of a matrix project independently (:139-148); and the fold order itself on a hand case."""
    rng = np.random.default_rng(42)
    s = rng.standard_normal((40, 10))
    assert np.all(oracle.project_rows(np.zeros((1, 40)), s) == 0.0)
    s = rng.standard_normal((25, 6))
    p1 = oracle.project_rows(np.ones((1, 25)), s)
    p2 = oracle.project_rows(2.0 * np.ones((1, 25)), s)
    assert np.array_equal(p2, 2.0 * p1)               # doubling is exact in binary floating point
    s = rng.standard_normal((50, 15))
    ratio = np.linalg.norm(oracle.project_rows(np.ones((1, 50)), s)) / math.sqrt(50.0)
    assert 0.5 < ratio < 2.0
    x = rng.standard_normal((7, 50))
    whole = oracle.project_rows(x, s)
    for i in range(7):
        assert np.array_equal(whole[i:i + 1], oracle.project_rows(x[i:i + 1], s))
    # ((x_i * s_ij) * scale) added left to right: a case where any other association differs in the last bit
    x = np.array([[0.1, 0.2, 0.3]])
    s = np.array([[0.7], [1.1], [-0.3]])
    want = 0.0
    for i in range(3):
        want = want + (x[0, i] * s[i, 0]) * (1.0 / math.sqrt(1.0))
    assert oracle.project_rows(x, s)[0, 0] == want
    s3 = rng.standard_normal((3, 3))
    scale = 1.0 / math.sqrt(3.0)
    want = [0.0] * 3
    for i in range(3):
        for j in range(3):
            want[j] = want[j] + (x[0, i] * s3[i, j]) * scale
    assert list(oracle.project_rows(x, s3)[0]) == want
    # successor, f32: (sum_i x_i s_ji) * scale
    xs = rng.standard_normal((2, 9)).astype(np.float32)
    sc = rng.standard_normal((4, 9)).astype(np.float32)
    got = oracle.project_rows_core(xs, sc)
    for r in range(2):
        for j in range(4):
            acc = np.float32(0)
            for i in range(9):
                acc = np.float32(acc + np.float32(xs[r, i] * sc[j, i]))
            assert got[r, j] == np.float32(acc * np.float32(np.float32(1.0) / np.sqrt(np.float32(4.0))))


# ---- SortedLambdas (src_legacy/sorted_index.rs:22-57) ------------------------------------------------------------
def test_sorted_lambdas_semantics(oracle):
    """Keys ascending, equal keys in one bucket ordered by the DECIMAL STRING of the index (zadd sorts by id), the
    key of a +-0 bucket is the first one inserted; std_dev in f32 as laplacian.rs:421-448 computes it."""
    lam = np.array([0.5, 0.1, 0.5, 0.0, -0.0, 0.1, 1.0, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5])
    srt, idx, sd = oracle.sorted_lambdas(lam)
    # the reference's structure, restated with Python's own containers
    buckets = {}
    for i, v in enumerate(lam):
        buckets.setdefault(float(v) + 0.0 if v != 0 else 0.0, []).append(i)
    want_idx = []
    for key in sorted(buckets):
        want_idx += sorted(buckets[key], key=str)
    assert list(idx) == want_idx == [3, 4, 1, 5, 0, 10, 11, 12, 2, 7, 8, 9, 6]
    assert np.array_equal(srt, np.sort(lam)) and not np.signbit(srt[0]) and not np.signbit(srt[1])
    lam2 = lam.copy(); lam2[3], lam2[4] = -0.0, 0.0
    srt2, idx2, _ = oracle.sorted_lambdas(lam2)
    assert list(idx2) == want_idx and np.signbit(srt2[0]) and np.signbit(srt2[1])
    mean = np.float32(np.float32(np.sum(lam)) / np.float32(len(lam)))
    var = np.float32(0)
    for v in lam:
        d = np.float32(mean - np.float32(v))
        var = np.float32(var + np.float32(d * d))
    assert sd == float(np.sqrt(np.float32(var / np.float32(len(lam)))))
    with pytest.raises(ValueError):
        oracle.sorted_lambdas(np.zeros(0))
    # NaN keys sort last and share one bucket (OrderedFloat)
    srt3, idx3, _ = oracle.sorted_lambdas(np.array([math.nan, 0.3, math.nan, math.inf]))
    assert list(idx3) == [1, 3, 0, 2] and np.isnan(srt3[2]) and np.isnan(srt3[3])


def test_projected_lambda_uses_the_unprojected_item_for_tau(oracle):
    """taumode.rs:174-175,261-318: tau = select_tau(&item.item) and the zero test see the UNPROJECTED item; E and G the
    projected one.  Same vector on both sides == the plain lambda; otherwise the blend of (tau(original), E, G(projected))."""
    rng = np.random.default_rng(8)
    ip = np.array([0, 2, 5, 7], np.uint64)
    ind = np.array([0, 1, 0, 1, 2, 1, 2], np.uint32)
    dat = np.array([1.0, -1.0, -1.0, 2.0, -1.0, -1.0, 1.0])       # chain graph (test_spectral.rs:187-251)
    xp = rng.standard_normal((6, 3))
    xo = np.abs(rng.standard_normal((6, 5)))
    xo[2] = 0.0
    assert np.array_equal(oracle.lambdas_projected(ip, ind, dat, xp, xp), oracle.lambdas(ip, ind, dat, xp))
    got = oracle.lambdas_projected(ip, ind, dat, xp, xo, tau_mode=oracle.TAU_MEAN)
    _, e, g = oracle.lambdas(ip, ind, dat, xp, tau_mode=oracle.TAU_MEAN, with_parts=True)
    for i in range(6):
        if i == 2:
            assert got[i] == 0.0                                    # zero ORIGINAL vector, whatever the projected one holds
            continue
        tau = oracle.select_tau(xo[i], oracle.TAU_MEAN)
        assert got[i] == tau * (e[i] / (e[i] + tau)) + (1.0 - tau) * min(max(g[i], 0.0), 1.0)
