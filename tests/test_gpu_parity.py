"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): neighbour indices and CSR structure bit-exact, ties by index;
distances bit-exact (same f64 arithmetic); Laplacian weights and lambda within 1e-9 relative."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9  # the tolerance north_star states for f64 weights and lambda


@pytest.fixture(scope="module")
def ctx(sfb):
    c = sfb.Context(0)
    yield c
    c.close()


def assert_knn_equal(got, want):
    assert np.array_equal(got[2], want[2]), "neighbour counts differ"
    assert np.array_equal(got[0], want[0]), "neighbour indices differ"
    assert np.array_equal(got[1], want[1]), "distances differ in bits"


# ---- synthetic generator: device rows == CPU replay -------------------------------------------
@pytest.mark.parametrize("kind,ncent,noise", [(0, 0, 0.0), (1, 16, 0.3), (2, 0, 0.1)])
def test_generator_bit_exact(sfb, oracle, ctx, kind, ncent, noise):
    m = ctx.generate(kind, 42, 777, 50, ncent, noise)
    got = m.rows()
    want = oracle.generate_rows(kind, 42, 0, 777, 50, ncent, noise)
    assert np.array_equal(got, want)
    assert np.array_equal(m.rows(100, 5), oracle.generate_rows(kind, 42, 100, 5, 50, ncent, noise))


def test_transpose(sfb, ctx):
    x = np.random.default_rng(1).normal(size=(77, 45))
    assert np.array_equal(ctx.matrix(x).transpose().rows(), x.T)


# ---- exact kNN --------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", [0, 1, 2])
@pytest.mark.parametrize("m,kd,k", [(50, 7, 3), (257, 33, 16), (1000, 64, 5), (130, 384, 32), (65, 16, 64)])
def test_knn_exact_parity(sfb, oracle, ctx, metric, m, kd, k):
    x = np.random.default_rng(m + kd + metric).normal(size=(m, kd))
    got = ctx.matrix(x).knn(k, metric, screen=sfb.SCREEN_EXACT_F64).to_host()
    assert_knn_equal(got, oracle.knn(x, k, metric))


def test_knn_ties_zero_rows_eps(sfb, oracle, ctx):
    x = np.array([[1.0, 0.0], [1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.0, 0.0], [2.0, 0.0], [0.0, 0.0]])
    for metric in (0, 1, 2):
        for eps in (math.inf, 0.5, 0.0):
            got = ctx.matrix(x).knn(3, metric, eps=eps, screen=sfb.SCREEN_EXACT_F64).to_host()
            assert_knn_equal(got, oracle.knn(x, 3, metric, eps))
    # duplicated rows: every tie resolved by index
    rng = np.random.default_rng(9)
    base = rng.normal(size=(40, 12))
    x = np.concatenate([base, base, base[:17]])
    got = ctx.matrix(x).knn(6, 0, screen=sfb.SCREEN_EXACT_F64).to_host()
    assert_knn_equal(got, oracle.knn(x, 6, 0))


def test_knn_query_shard(sfb, oracle, ctx):
    x = np.random.default_rng(4).normal(size=(500, 24))
    g = ctx.matrix(x).knn(7, 0, screen=sfb.SCREEN_EXACT_F64, q_begin=120, q_end=333)
    assert g.shape == (213, 7) and g.q_begin == 120
    want = oracle.knn(x, 7, 0, query_rows=np.arange(120, 333))
    assert_knn_equal(g.to_host(), want)


# ---- feature graph: few nodes with very long rows (graph.rs:193-216) ----------------------------
@pytest.mark.parametrize("metric", [0, 1, 2])
@pytest.mark.parametrize("n_items,n_feat,k", [(3000, 48, 5), (2001, 33, 3), (5000, 130, 16), (777, 17, 16)])
def test_knn_columns_feature_graph(sfb, oracle, ctx, metric, n_items, n_feat, k):
    """Nodes = columns of the item matrix; Gram-tile kernel (dims-major input, odd / even widths, ragged
    tiles and chunks) and the rows-as-nodes entry point on the transposed copy agree with the oracle."""
    x = np.random.default_rng(n_items + n_feat + metric).normal(size=(n_items, n_feat))
    x[:, 3] = x[:, 1]          # duplicated feature: exact ties, broken by index
    x[:, 5] = 0.0              # a zero feature: cosine := 0 -> distance 1 to everything
    want = oracle.knn(oracle.transpose(x), k, metric)
    m = ctx.matrix(x)
    assert_knn_equal(m.knn_columns(k, metric).to_host(), want)
    assert_knn_equal(m.transpose().knn(k, metric).to_host(), want)
    eps = float(np.median(want[1][np.isfinite(want[1])]))
    assert_knn_equal(m.knn_columns(k, metric, eps=eps).to_host(), oracle.knn(oracle.transpose(x), k, metric, eps))


@pytest.mark.parametrize("kernel", ["regs", "smem"])
@pytest.mark.parametrize("mode", ["co", "after"])
@pytest.mark.parametrize("gt", ["8", "16", "32"])
@pytest.mark.parametrize("metric", [0, 1])
def test_knn_columns_gram_tile_variants(sfb, oracle, ctx, gt, metric, mode, kernel, monkeypatch):
    """Both pair-tile edges of the feature-graph Gram kernels -- the warp-per-tile kernel with its operands in a register ring
    (the default beside a screen) and the shared-memory ring (the default alone) -- each stand-alone and beside the screen, on shapes with ragged tiles,
    odd node counts and a dimension count that is not a multiple of a round / staged chunk: every pair sum is the reference's
    left fold, so the lists are bit-exact."""
    monkeypatch.setenv("SFB_GRAM_GT", gt)
    monkeypatch.setenv("SFB_GRAM_SMEM" if kernel == "smem" else "SFB_GRAM_REGS", "1")   # either kernel in both roles
    monkeypatch.setenv("SFB_GRAM_MODE", mode)   # beside the screen kernel / released when it has finished
    for n_items, n_feat, k in ((20000, 37, 5), (6000, 130, 16), (17001, 64, 8)):   # two long enough for the side stream, one run inline
        x = np.random.default_rng(n_items + metric).normal(size=(n_items, n_feat))
        m = ctx.matrix(x)
        want = oracle.knn(oracle.transpose(x), k, metric)
        assert_knn_equal(m.knn_columns(k, metric).to_host(), want)
        pend = m.knn_columns_begin(k, metric)      # fired beside the screen of the next knn()
        g = m.knn(4, sfb.METRIC_COSINE, screen=sfb.SCREEN_F16)
        assert_knn_equal(pend.end().to_host(), want)
        g.free()


@pytest.mark.parametrize("metric", [0, 1, 2])
@pytest.mark.parametrize("n,d,k", [(20011, 37, 16), (16384, 384, 16), (17000, 5, 128), (18000, 16, 3)])
def test_knn_exact_few_rows(sfb, oracle, ctx, metric, n, d, k, monkeypatch):
    """A handful of query rows against a long corpus -- the shape of the screen's fallback -- take the one-corpus-row-per-lane
    kernel: 1 to 8 rows, at the start and at the ragged end of the corpus, with duplicate rows (index ties), a zero row and a
    distance cap; the same bits as the tile kernel (SFB_EXACT_NO_FEW) and as the oracle."""
    rng = np.random.default_rng(n + d + metric)
    x = rng.normal(size=(n, d))
    x[7] = x[3]; x[n - 2] = x[3]; x[11] = 0.0; x[n - 1] = x[n - 5]
    m = ctx.matrix(x)
    for q0, q1 in ((0, 1), (2, 5), (3, 7), (4, 9), (5, 13), (n - 8, n), (n - 1, n)):
        want = oracle.knn(x, k, metric, query_rows=np.arange(q0, q1, dtype=np.uint32))
        assert_knn_equal(m.knn(k, metric, screen=sfb.SCREEN_EXACT_F64, q_begin=q0, q_end=q1).to_host(), want)
    eps = float(np.median(want[1][np.isfinite(want[1])]))
    want = oracle.knn(x, k, metric, eps, query_rows=np.arange(2, 8, dtype=np.uint32))
    assert_knn_equal(m.knn(k, metric, eps=eps, screen=sfb.SCREEN_EXACT_F64, q_begin=2, q_end=8).to_host(), want)
    monkeypatch.setenv("SFB_EXACT_NO_FEW", "1")
    assert_knn_equal(m.knn(k, metric, eps=eps, screen=sfb.SCREEN_EXACT_F64, q_begin=2, q_end=8).to_host(), want)


def test_knn_columns_many_nodes(sfb, oracle, ctx):
    """Columns as nodes when the shape is NOT the feature-graph shape: falls back to a node-major copy."""
    x = np.random.default_rng(77).normal(size=(40, 600))
    assert_knn_equal(ctx.matrix(x).knn_columns(9, 0).to_host(), oracle.knn(oracle.transpose(x), 9, 0))


def test_knn_argument_errors(sfb, ctx):
    x = ctx.matrix(np.ones((4, 3)))
    for bad in (dict(k=0), dict(k=129), dict(k=2, metric=7), dict(k=2, eps=math.nan), dict(k=2, q_begin=3, q_end=2)):
        with pytest.raises(sfb.SfbError):
            x.knn(**bad)
    with pytest.raises(sfb.SfbError):
        ctx.matrix(np.ones((1, 3))).knn(1)  # assert!(n >= 2), laplacian.rs:130


# ---- weights, sparsification ------------------------------------------------------------------
@pytest.mark.parametrize("p,sigma", [(2.0, 1.0), (1.0, 0.2), (1.7, 0.5)])
def test_adjacency_parity(sfb, oracle, ctx, p, sigma):
    x = np.random.default_rng(11).normal(size=(300, 10))
    for k in (5, 12, 40):  # mean degree <= 10: untouched; > 10: inline sparsification
        o_knn = oracle.knn(x, k)
        g = sfb.KnnGraph.from_host(ctx, *o_knn)
        a = g.adjacency(p, sigma)
        idx, w, cnt = a.to_host()
        o_idx, o_w, o_cnt, o_applied = oracle.build_adjacency(*o_knn, p, sigma)
        assert a.sparsified == o_applied == (k > 10)
        assert np.array_equal(cnt, o_cnt) and np.array_equal(idx, o_idx)
        if p in (1.0, 2.0):
            assert np.array_equal(w, o_w)
        assert np.allclose(w, o_w, rtol=1e-14, atol=0)


def test_adjacency_weight_floor_and_forced(sfb, oracle, ctx):
    x = np.random.default_rng(12).normal(size=(120, 6))
    o_knn = oracle.knn(x, 8)
    g = sfb.KnnGraph.from_host(ctx, *o_knn)
    # tiny sigma: w = 1/(1+(d/sigma)^p) <= 1e-12 for most edges -> dropped (laplacian.rs:257)
    got = g.adjacency(4.0, 1e-4).to_host()
    want = oracle.build_adjacency(*o_knn, 4.0, 1e-4)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0]) and got[2].sum() < o_knn[2].sum()
    for force in (0, 1):
        got = g.adjacency(2.0, 1.0, sparsify=force).to_host()
        want = oracle.build_adjacency(*o_knn, 2.0, 1.0, force_sparsify=force)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])


@pytest.mark.parametrize("ratio", [0.05, 0.3, 0.5, 1.0])
def test_sfgrass_parity(sfb, oracle, ctx, ratio):
    x = np.random.default_rng(13).normal(size=(200, 8))
    for k in (6, 20, 64):
        o_adj = oracle.build_adjacency(*oracle.knn(x, k), 2.0, 1.0, force_sparsify=0)
        a = sfb.Adjacency.from_host(ctx, *o_adj[:3])
        applied = a.sfgrass(ratio)
        want = oracle.sfgrass(*o_adj[:3], ratio=ratio)
        assert applied == want[3] == (k >= 10)
        got = a.to_host()
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_sfgrass_reference_layer(sfb, oracle, ctx):
    x = np.random.default_rng(14).normal(size=(60, 5))
    o_adj = oracle.build_adjacency(*oracle.knn(x, 12), 2.0, 1.0, force_sparsify=0)
    rows = [[(int(o_adj[0][i, t]), float(o_adj[1][i, t])) for t in range(int(o_adj[2][i]))] for i in range(60)]
    out = sfb.SfGrassSparsifier().with_target_ratio(0.25).sparsify_graph(rows, 60, ctx=ctx)
    want = oracle.sfgrass(*o_adj[:3], ratio=0.25)
    assert [len(r) for r in out] == list(want[2])
    assert [j for j, _ in out[7]] == list(want[0][7, :want[2][7]])


def test_sfgrass_reference_cases(sfb, oracle, ctx):
    """The reference's own SF-GRASS tests (src_legacy/tests/test_sparsification.rs:3-46), ragged rows through the mirror."""
    basic = [[(1, 1.0), (2, 0.5)], [(0, 1.0), (2, 0.8)], [(0, 0.5), (1, 0.8)]]
    out = sfb.SfGrassSparsifier().sparsify_graph(basic, 3, ctx=ctx)
    assert len(out) == 3 and all(len(r) > 0 for r in out) and [j for j, _ in out[0]] == [1, 2]
    n = 50
    rows = [[(j, 1.0 / (1.0 + abs(i - j))) for j in range(n) if i != j and (i + j) % 3 == 0] for i in range(n)]
    out = sfb.SfGrassSparsifier().sparsify_graph(rows, n, ctx=ctx)
    assert len(out) == n and sum(len(r) for r in out) < sum(len(r) for r in rows)
    k = max(len(r) for r in rows)
    idx = np.full((n, k), 0xFFFFFFFF, np.uint32); w = np.zeros((n, k)); cnt = np.zeros(n, np.uint32)
    for i, r in enumerate(rows):
        cnt[i] = len(r)
        for t, (j, v) in enumerate(r):
            idx[i, t], w[i, t] = j, v
    want = oracle.sfgrass(idx, w, cnt)
    for i in range(n):
        assert [j for j, _ in out[i]] == want[0][i, :want[2][i]].tolist()
        assert [v for _, v in out[i]] == want[1][i, :want[2][i]].tolist()


# ---- symmetrise + Laplacian -------------------------------------------------------------------
def assert_csr_equal(got, want, data_exact=False):
    assert np.array_equal(got[0], want[0]), "indptr differs"
    assert np.array_equal(got[1], want[1]), "indices differ"
    if data_exact:
        assert np.array_equal(got[2], want[2])
    assert np.allclose(got[2], want[2], rtol=RTOL, atol=0)


@pytest.mark.parametrize("m,kd,k", [(40, 5, 3), (300, 10, 16), (2000, 16, 8), (500, 6, 64)])
def test_laplacian_parity(sfb, oracle, ctx, m, kd, k):
    x = np.random.default_rng(m).normal(size=(m, kd))
    o_adj = oracle.build_adjacency(*oracle.knn(x, k), 2.0, 1.0)
    a = sfb.Adjacency.from_host(ctx, *o_adj[:3])
    got = a.laplacian().to_host()
    assert_csr_equal(got, oracle.laplacian(*o_adj[:3]), data_exact=True)
    # invariants of src_legacy/tests/test_laplacian.rs:52-154
    indptr, indices, data = got
    row = np.repeat(np.arange(m), np.diff(indptr.astype(np.int64)))
    sums = np.zeros(m)
    np.add.at(sums, row, data)
    assert np.all(np.abs(sums) < 1e-12)
    assert len(data) <= m * (2 * k + 1)


def test_laplacian_hub_rows(sfb, oracle, ctx):
    """A star: node 0 is everybody's neighbour -> one row with M-1 reverse edges (block-sort path)."""
    m, k = 3000, 2
    idx = np.zeros((m, k), np.uint32)
    idx[:, 0] = 0
    idx[:, 1] = (np.arange(m) + 1) % m
    idx[0] = [1, 2]
    w = np.random.default_rng(2).uniform(0.1, 1.0, size=(m, k))
    cnt = np.full(m, k, np.uint32)
    a = sfb.Adjacency.from_host(ctx, idx, w, cnt)
    assert_csr_equal(a.laplacian().to_host(), oracle.laplacian(idx, w, cnt), data_exact=True)


def test_laplacian_hub_beyond_shared_memory(sfb, oracle, ctx):
    """A star: 20 000 nodes all point at node 0 (plus a second hub of 9 000), so two rows have more incident edges than
    the 8192 a block sorts in shared memory: they are sorted in global memory.  The reference has no such limit."""
    m, k = 20001, 3
    rng = np.random.default_rng(9)
    idx = np.full((m, k), 0xFFFFFFFF, np.uint32); w = np.zeros((m, k)); cnt = np.zeros(m, np.uint32)
    for i in range(1, m):
        nb = [0] + ([7] if i % 2 == 0 and i < 18000 and i != 7 else []) + [int(rng.integers(1, m))]
        nb = [j for t, j in enumerate(nb) if j != i and j not in nb[:t]]
        idx[i, :len(nb)] = nb; w[i, :len(nb)] = rng.uniform(0.1, 1.0, len(nb)); cnt[i] = len(nb)
    idx[0, :2] = [5, 7]; w[0, :2] = [0.9, 0.3]; cnt[0] = 2      # duplicates of reverse edges: max of the two weights
    L = sfb.Adjacency.from_host(ctx, idx, w, cnt).laplacian()
    want = oracle.laplacian(idx, w, cnt)
    assert_csr_equal(L.to_host(), want, data_exact=True)
    assert np.diff(want[0].astype(np.int64)).max() > 8192


@pytest.mark.parametrize("m,kd,k", [(3000, 12, 16), (777, 5, 40)])
def test_laplacian_row_shards_concatenate_to_the_full_matrix(sfb, oracle, ctx, m, kd, k):
    """sfb_laplacian_build_rows: the rows a rank owns, from the all-gathered lists."""
    x = np.random.default_rng(m).normal(size=(m, kd))
    a = ctx.matrix(x).knn(k, sfb.METRIC_L2SQ, screen=sfb.SCREEN_EXACT_F64).adjacency(2.0, 1.0)
    full = a.laplacian().to_host()
    cuts = [0, m // 3 + 1, m // 3 + 2, m - 5, m]
    ptr, ind, dat = [np.zeros(1, np.uint64)], [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        sh = a.laplacian(rows=(lo, hi))
        assert sh.shape[0] == hi - lo
        p, i, d = sh.to_host()
        ptr.append(p[1:] + ptr[-1][-1]); ind.append(i); dat.append(d)
        with pytest.raises(sfb.SfbError):   # a shard is not a square operator
            sh.spmv(np.ones(hi - lo))
    assert np.array_equal(np.concatenate(ptr), full[0]) and np.array_equal(np.concatenate(ind), full[1]) and np.array_equal(np.concatenate(dat), full[2])
    with pytest.raises(sfb.SfbError):
        a.laplacian(rows=(5, 5))
    with pytest.raises(sfb.SfbError):
        a.laplacian(normalised=True, rows=(0, 10))


def test_laplacian_asymmetric_weights_take_max(sfb, oracle, ctx):
    idx = np.array([[1, 2], [0, 2], [0, 0xFFFFFFFF]], np.uint32)
    w = np.array([[0.5, 0.25], [0.75, 0.125], [0.3, 0.0]])
    cnt = np.array([2, 2, 1], np.uint32)
    got = sfb.Adjacency.from_host(ctx, idx, w, cnt).laplacian().to_host()
    assert_csr_equal(got, oracle.laplacian(idx, w, cnt), data_exact=True)
    assert got[2][1] == -0.75  # L[0][1] = -max(0.5, 0.75)


def test_laplacian_empty_and_isolated(sfb, oracle, ctx):
    idx = np.full((5, 2), 0xFFFFFFFF, np.uint32)
    w = np.zeros((5, 2))
    cnt = np.zeros(5, np.uint32)
    idx[1, 0], w[1, 0], cnt[1] = 3, 0.5, 1
    got = sfb.Adjacency.from_host(ctx, idx, w, cnt).laplacian().to_host()
    assert_csr_equal(got, oracle.laplacian(idx, w, cnt), data_exact=True)
    assert len(got[2]) == 7  # 5 diagonals (zeros stored, laplacian.rs:372) + one undirected edge


@pytest.mark.parametrize("thr", [1e-9, 0.8])
def test_laplacian_normalised(sfb, oracle, ctx, thr):
    x = np.random.default_rng(21).normal(size=(400, 9))
    o_adj = oracle.build_adjacency(*oracle.knn(x, 6, eps=0.6), 2.0, 1.0)
    a = sfb.Adjacency.from_host(ctx, *o_adj[:3])
    got = a.laplacian(normalised=True, weight_threshold=thr).to_host()
    assert_csr_equal(got, oracle.laplacian(*o_adj[:3], normalised=True, weight_threshold=thr), data_exact=True)


def test_spmv_rayleigh(sfb, oracle, ctx):
    x = np.random.default_rng(22).normal(size=(250, 7))
    L = oracle.laplacian(*oracle.build_adjacency(*oracle.knn(x, 5), 2.0, 1.0)[:3])
    c = sfb.Csr.from_host(ctx, *L)
    v = np.random.default_rng(23).normal(size=250)
    assert np.array_equal(c.spmv(v), oracle.spmv(*L, v))
    assert c.rayleigh_quotient(v) == pytest.approx(oracle.rayleigh(*L, v), rel=1e-12)
    assert c.rayleigh_quotient(np.zeros(250)) == 0.0


# ---- lambda ------------------------------------------------------------------------------------
def feature_laplacian(oracle, x, topk):
    idx, dist, cnt = oracle.knn(oracle.transpose(x), topk)
    a = oracle.build_adjacency(idx, dist, cnt, 2.0, 1.0)
    return oracle.laplacian(*a[:3])


@pytest.mark.parametrize("tau", [(1, 0.0), (2, 0.0), (0, 0.35), (0, -1.0), (3, 0.0), (3, 0.37), (3, 1.0)])
@pytest.mark.parametrize("n,f,topk", [(300, 48, 3), (257, 33, 6), (64, 384, 16)])
def test_lambda_legacy_parity(sfb, oracle, ctx, tau, n, f, topk):
    rng = np.random.default_rng(n + f)
    x = rng.normal(size=(n, f)) + 0.5
    x[3] = 0.0            # zero vector -> lambda 0 (taumode.rs:268-274)
    x[4] = 1e-11
    x[5] = 2.5            # constant vector
    x[6, :f // 2] = x[6, f // 2: 2 * (f // 2)]  # duplicates around the median
    L = feature_laplacian(oracle, x, topk)
    c = sfb.Csr.from_host(ctx, *L)
    lam, disp, stats = c.lambdas(ctx.matrix(x), sfb.LAMBDA_LEGACY_TAUMODE, tau[0], tau[1], with_dispersion=True)
    o_lam, o_e, o_g = oracle.lambdas(*L, x, oracle.LAMBDA_LEGACY_TAUMODE, tau[0], tau[1], with_parts=True)
    assert lam[3] == 0.0 and lam[4] == 0.0
    assert np.allclose(disp, o_g, rtol=RTOL, atol=1e-15)
    assert np.allclose(lam, o_lam, rtol=RTOL, atol=1e-14)
    # normalised: core.rs:1341-1355
    lam_n, stats = c.lambdas(ctx.matrix(x), sfb.LAMBDA_LEGACY_TAUMODE, tau[0], tau[1], normalise=True)
    o_n, o_stats = oracle.normalise_lambdas(o_lam)
    assert np.allclose(stats, o_stats, rtol=RTOL, atol=1e-14)
    assert np.allclose(lam_n, o_n, rtol=1e-8, atol=1e-12)
    assert lam_n.min() == 0.0 and lam_n.max() <= 1.0 + 1e-12


def test_lambda_tau_nonfinite_entries(sfb, oracle, ctx):
    """select_tau filters NaN / inf (taumode.rs:41,50); here via an all-finite matrix with huge spread."""
    x = np.random.default_rng(31).normal(size=(50, 40)) * np.logspace(-3, 3, 40)
    L = feature_laplacian(oracle, x, 4)
    c = sfb.Csr.from_host(ctx, *L)
    for tm, tv in ((1, 0), (3, 0.9)):
        lam, _ = c.lambdas(ctx.matrix(x), sfb.LAMBDA_LEGACY_TAUMODE, tm, tv)
        assert np.allclose(lam, oracle.lambdas(*L, x, 0, tm, tv), rtol=RTOL, atol=1e-14)


def test_lambda_energy_node_parity(sfb, oracle, ctx):
    x = np.random.default_rng(32).normal(size=(500, 64))
    L = feature_laplacian(oracle, x, 4)
    c = sfb.Csr.from_host(ctx, *L)
    lam, disp, _ = c.lambdas(ctx.matrix(x), sfb.LAMBDA_ENERGY_NODE, with_dispersion=True)
    o_lam, _, o_g = oracle.lambdas(*L, x, oracle.LAMBDA_ENERGY_NODE, with_parts=True)
    assert np.allclose(lam, o_lam, rtol=RTOL, atol=1e-15) and np.allclose(disp, o_g, rtol=RTOL, atol=1e-15)
    # upper-triangle dispersion == 2x the taumode dispersion for a symmetric L
    _, disp_t, _ = c.lambdas(ctx.matrix(x), sfb.LAMBDA_LEGACY_TAUMODE, with_dispersion=True)
    assert np.allclose(disp, 2.0 * disp_t, rtol=1e-12)


def test_lambda_core_f32sem(sfb, oracle, ctx):
    x = np.random.default_rng(33).normal(size=(200, 32))
    L = feature_laplacian(oracle, x, 4)
    c = sfb.Csr.from_host(ctx, *L)
    lam, _ = c.lambdas(ctx.matrix(x), sfb.LAMBDA_CORE_F32SEM)
    o_lam = oracle.lambdas(*L, x, oracle.LAMBDA_CORE_F32SEM)
    assert np.allclose(lam, o_lam, rtol=1e-5, atol=1e-6)  # f32 semantics: the reference's own tolerance


def test_lambda_kats(sfb, oracle, ctx):
    """chain graph (test_spectral.rs:187-251) through the CUDA path"""
    indptr = np.array([0, 2, 5, 7], np.uint64)
    indices = np.array([0, 1, 0, 1, 2, 1, 2], np.uint32)
    data = np.array([1.0, -1, -1, 2, -1, -1, 1])
    c = sfb.Csr.from_host(ctx, indptr, indices, data)
    x = ctx.matrix(np.array([[1.0, 1.0, 1.0], [1.0, 0.0, -1.0]]))
    lam, disp, _ = c.lambdas(x, sfb.LAMBDA_LEGACY_TAUMODE, sfb.TAU_FIXED, 0.5, with_dispersion=True)
    assert lam[0] == 0.0 and disp[1] == 0.25 and lam[1] == 0.5 * (1.0 / 1.5) + 0.5 * 0.25
    lam, disp, _ = c.lambdas(x, sfb.LAMBDA_ENERGY_NODE, with_dispersion=True)
    assert lam[1] == 1.0 and disp[1] == 0.5
    with pytest.raises(sfb.SfbError):  # assert_eq!(matrix.rows(), n), taumode.rs:330-337
        c.lambdas(ctx.matrix(np.ones((2, 4))))


@pytest.mark.parametrize("n,f,topk", [(300, 40, 3), (1000, 128, 4), (77, 200, 8), (33, 384, 16), (1, 64, 3)])
def test_diffusion_bit_exact(sfb, oracle, ctx, n, f, topk):
    """Both diffusion kernels (the tile kernel when two tiles and L fit shared memory, the row-wise one otherwise):
    every (L x)_r is the reference's left fold, so the rows come out bit for bit, ragged tiles included."""
    x = np.random.default_rng(34 + n).normal(size=(n, f))
    L = feature_laplacian(oracle, np.random.default_rng(f).normal(size=(max(n, 64), f)), topk)
    c = sfb.Csr.from_host(ctx, *L)
    for steps in (1, 4, 5):
        m = ctx.matrix(x).diffuse(c, 0.1, steps)
        assert np.array_equal(m.rows(), oracle.diffuse(*L, x, 0.1, steps))
    assert np.array_equal(ctx.matrix(x).diffuse(c, 0.1, 0).rows(), x)


# ---- reference-shaped entry points ------------------------------------------------------------
def test_reference_layer_pipeline(sfb, oracle, ctx):
    rng = np.random.default_rng(40)
    centres = rng.normal(size=(5, 24)) * 3
    x = centres[rng.integers(0, 5, size=600)] + 0.4 * rng.normal(size=(600, 24))
    params = sfb.GraphParams(eps=math.inf, k=6, topk=3, p=2.0, sigma=None)
    gl = sfb.GraphFactory.build_laplacian_matrix_from_k_cluster(x, params.eps, params.k, params.topk, params.p,
                                                                params.sigma, False, False, 600, ctx=ctx)
    assert gl.shape() == (24, 24) and gl.nnodes == 600
    L = feature_laplacian(oracle, x, 3)
    assert_csr_equal(gl.csr(), L, data_exact=True)
    v = rng.normal(size=24)
    assert np.array_equal(gl.multiply_vector(v), oracle.spmv(*L, v))
    assert gl.rayleigh_quotient(v) == pytest.approx(oracle.rayleigh(*L, v), rel=1e-12)
    assert np.allclose(gl.degrees(), [L[2][int(L[0][r]) + list(L[1][int(L[0][r]):int(L[0][r + 1])]).index(r)] for r in range(24)])
    for tm in (sfb.TauMode.Median, sfb.TauMode.Mean, sfb.TauMode.Fixed(0.3), sfb.TauMode.Percentile(0.25)):
        lam = sfb.TauMode.compute_taumode_lambdas_parallel(x, gl, tm)
        o_lam, _ = oracle.normalise_lambdas(oracle.lambdas(*L, x, 0, tm.kind, tm.value))
        assert np.allclose(lam, o_lam, rtol=1e-8, atol=1e-12)
        assert lam.min() >= 0.0 and lam.max() <= 1.0 + 1e-12 and np.all(np.isfinite(lam))
    with pytest.raises(sfb.SfbError):  # too sparse (graph.rs:232-240)
        big = rng.normal(size=(50, 200))
        sfb.GraphFactory.build_laplacian_matrix_from_k_cluster(big, math.inf, 6, 1, 2.0, None, False, True, 50, ctx=ctx)


# ---- config C1: 10k x 384 Gaussian, cosine, k = 16 (the reference's CPU-runnable case) --------
@pytest.mark.parametrize("screen", ["f16", "exact"])
def test_config_c1_exact(sfb, oracle, ctx, screen):
    """The whole C1 build -- item graph through the tensor-core screen (and through the exact kernel), Laplacian,
    feature graph, per-item lambda -- against the oracle's full CPU build."""
    m = ctx.generate(sfb.SYNTH_GAUSSIAN, 42, 10000, 384)
    x = oracle.generate_rows(0, 42, 0, 10000, 384)
    g = m.knn(16, sfb.METRIC_COSINE, screen=sfb.SCREEN_F16 if screen == "f16" else sfb.SCREEN_EXACT_F64)
    want = oracle.knn(x, 16, oracle.METRIC_COSINE)
    assert_knn_equal(g.to_host(), want)
    if screen == "f16":
        st = g.stats()
        assert st["screen_used"] == sfb.SCREEN_F16 and st["rows_certified"] + st["rows_fallback"] == 10000
    a = g.adjacency(2.0, 1.0)
    o_adj = oracle.build_adjacency(*want, 2.0, 1.0)
    assert a.sparsified and o_adj[3]
    # weights: glibc pow(r, 2.0) vs the device's exactly rounded r*r may differ in the last bit
    assert_csr_equal(a.laplacian().to_host(), oracle.laplacian(*o_adj[:3]))
    # feature graph over the 384 columns + taumode lambda of every item, min-max normalised
    gf = m.knn_columns(16, sfb.METRIC_COSINE)
    f_want = oracle.knn(oracle.transpose(x), 16, oracle.METRIC_COSINE)
    assert_knn_equal(gf.to_host(), f_want)
    Lf = gf.adjacency(2.0, 1.0).laplacian()
    fl = oracle.laplacian(*oracle.build_adjacency(*f_want, 2.0, 1.0)[:3])
    assert_csr_equal(Lf.to_host(), fl)
    lam, stats = Lf.lambdas(m, sfb.LAMBDA_LEGACY_TAUMODE, sfb.TAU_MEDIAN, normalise=True)
    o_lam, o_stats = oracle.normalise_lambdas(oracle.lambdas(*fl, x, oracle.LAMBDA_LEGACY_TAUMODE, oracle.TAU_MEDIAN))
    assert np.allclose(lam, o_lam, rtol=1e-9, atol=1e-12) and np.allclose(stats, o_stats, rtol=1e-9)


# ---- the lambda tile kernel (lane = item) at the shapes of the configs ----------------------------------------------
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,f,topk", [(1000, 128, 4), (777, 200, 8), (2000, 384, 16), (300, 512, 16), (200, 768, 16), (33, 40, 3), (1, 64, 3),
                                      (100, 769, 8), (37, 1646, 6), (21, 1647, 6), (50, 3072, 8), (9, 3111, 4), (12, 3112, 4)])
def test_lambda_tile_kernel_shapes(sfb, oracle, ctx, variant, n, f, topk):
    """Every register-block size of the tile kernel (E = 4 .. 24 values per lane), the 16- and 8-item tiles beyond 768
    features (and the row-wise kernel past their 3111), ragged tails (n not a multiple of the tile),
    a Laplacian built on the device (symmetric by construction, defect 0) and one from the host with a non-zero row-sum
    defect and a positive off-diagonal entry (the general edge loop)."""
    rng = np.random.default_rng(n * 7 + f)
    x = rng.normal(size=(n, f)) * rng.uniform(0.1, 3.0, size=(1, f)) + 0.3
    if n > 8:
        x[1] = 0.0; x[2] = -4.0; x[5, : 2 * (f // 2) : 2] = x[5, 1::2]
    xm = ctx.matrix(x)
    gf = xm.knn_columns(min(topk, f - 1), sfb.METRIC_COSINE)
    Lf = gf.adjacency(2.0, 1.0).laplacian()
    fl = Lf.to_host()
    o_variant = oracle.LAMBDA_ENERGY_NODE if variant else oracle.LAMBDA_LEGACY_TAUMODE
    for tm, tv in ((1, 0.0), (3, 0.25), (2, 0.0)):
        lam, disp, stats = Lf.lambdas(xm, variant, tm, tv, with_dispersion=True)
        o_lam, _, o_g = oracle.lambdas(*fl, x, o_variant, tm, tv, with_parts=True)
        assert np.allclose(lam, o_lam, rtol=RTOL, atol=1e-14) and np.allclose(disp, o_g, rtol=RTOL, atol=1e-15)
        assert stats[0] == lam.min() and stats[1] == max(0.0, lam.max())
    # a symmetric matrix that is not a Laplacian: row-sum defect, one positive off-diagonal pair
    ip, ix, dv = [a.copy() for a in fl]
    dv[ix == np.repeat(np.arange(f), np.diff(ip.astype(np.int64)))] *= 1.25
    r0 = 0
    e = int(ip[r0]) + int(np.argmax(ix[int(ip[r0]):int(ip[r0 + 1])] != r0))
    c0 = int(ix[e]); dv[e] = 0.125
    e2 = int(ip[c0]) + int(np.searchsorted(ix[int(ip[c0]):int(ip[c0 + 1])], r0)); dv[e2] = 0.125
    Lh = sfb.Csr.from_host(ctx, ip, ix, dv)
    lam, disp, _ = Lh.lambdas(xm, variant, 1, 0.0, with_dispersion=True)
    o_lam, _, o_g = oracle.lambdas(ip, ix, dv, x, o_variant, 1, 0.0, with_parts=True)
    assert np.allclose(lam, o_lam, rtol=RTOL, atol=1e-13) and np.allclose(disp, o_g, rtol=RTOL, atol=1e-15)


@pytest.mark.parametrize("f", [96, 1000, 2000])
def test_lambda_tile_kernel_tau_edge_cases(sfb, oracle, ctx, f):
    """Median / percentile selection in registers (f = 96) and over a column of a narrow tile (f = 1000, 2000): all-equal rows, two-valued rows, heavy duplicates around the median,
    rows whose values collapse in f32 (the quantised histogram cannot separate them: exact path), huge dynamic range,
    NaN / inf entries (select_tau keeps the finite ones, taumode.rs:41,50)."""
    rng = np.random.default_rng(5)
    base = rng.normal(size=(64, f))
    L = feature_laplacian(oracle, base, 4)
    c = sfb.Csr.from_host(ctx, *L)
    x = rng.normal(size=(40, f))
    x[0] = 3.0
    x[1, :f // 2] = 1.0; x[1, f // 2:] = 2.0
    x[2] = np.round(x[2])                                   # many duplicates
    x[3] = 1.0 + np.arange(f) * 1e-13                       # equal in f32
    x[4] = 1e300 * np.sign(x[4]) * np.abs(x[4])             # f32 overflow in the quantisation
    x[5] = np.logspace(-200, 200, f)
    x[6, 3] = np.nan; x[6, 10] = np.inf; x[6, 11] = -np.inf
    x[7, :] = np.where(np.arange(f) % 3 == 0, 0.5, x[7])
    x[8] = 1e-11                                            # zero vector by the 1e-10 rule
    x[9, :] = 0.0; x[9, 17] = 1e-9                          # not a zero vector
    for tm, tv in ((1, 0.0), (3, 0.0), (3, 0.5), (3, 0.99), (3, 1.0), (2, 0.0)):
        lam, _ = c.lambdas(ctx.matrix(x), sfb.LAMBDA_LEGACY_TAUMODE, tm, tv)
        want = oracle.lambdas(*L, x, oracle.LAMBDA_LEGACY_TAUMODE, tm, tv)
        ok = np.isclose(lam, want, rtol=RTOL, atol=1e-14) | (np.isnan(lam) & np.isnan(want))
        assert ok.all(), (tm, tv, np.nonzero(~ok)[0], lam[~ok], want[~ok])


def test_compute_tau_core(sfb, oracle, ctx):
    """compute_tau (surfface-core/src/taumode.rs:37-65) on the device against the oracle, all modes."""
    rng = np.random.default_rng(8)
    cases = [rng.normal(size=1001).astype(np.float32) ** 2, rng.uniform(size=4096).astype(np.float32), np.array([0.3], np.float32),
             np.array([np.nan, np.inf], np.float32), np.zeros(0, np.float32), np.array([1e-12, 2e-12, np.nan, 5e-10], np.float32),
             np.round(rng.normal(size=300) * 3).astype(np.float32), rng.normal(size=200_003).astype(np.float32)]
    for lam in cases:
        for mode, val in ((1, 0.0), (2, 0.0), (0, 0.25), (0, np.nan), (0, -1.0), (3, 0.0), (3, 0.5), (3, 0.37), (3, 1.0), (3, 2.0), (3, np.nan)):
            got, want = ctx.compute_tau(lam, mode, val), oracle.compute_tau_core(lam, mode, val)
            assert got == want or (np.isnan(got) and np.isnan(want)), (len(lam), mode, val, got, want)
    # the reference's own table (surfface-core/src/tests/test_taumode.rs)
    assert sfb.compute_tau([0.1, 0.2, 0.3, 0.4, 0.5], sfb.CoreTauMode.Median, ctx=ctx) == np.float32(0.3)
    assert sfb.compute_tau([], sfb.CoreTauMode.Median, ctx=ctx) == np.float32(1e-9)


def test_compute_tau_mode_gpu_seam(sfb, oracle, ctx):
    """compute_tau_mode_gpu(&LaplacianOutput, data: &[f32], n, f) -> Vec<f64> (spectral/bridge.rs:27-32): f32 upload."""
    rng = np.random.default_rng(9)
    means = rng.normal(size=(30, 48)).astype(np.float32); variances = rng.uniform(0.1, 1.0, size=(30, 48)).astype(np.float32)
    out = sfb.LaplacianStage.with_defaults().execute(means, variances, ctx=ctx)
    data = rng.normal(size=(500, 48)).astype(np.float32)
    lam = sfb.compute_tau_mode_gpu(out, data.reshape(-1), 500, 48, ctx=ctx)
    want = oracle.lambdas(*out.matrix.to_host(), data.astype(np.float64), oracle.LAMBDA_CORE_F32SEM)
    assert lam.dtype == np.float64 and np.allclose(lam, want, rtol=1e-5, atol=1e-6)
    # the f32 upload is exact
    assert np.array_equal(ctx.matrix_f32(data).rows(), data.astype(np.float64))
    assert np.array_equal(ctx.matrix_copy(ctx.matrix_f32(data)).rows(), data.astype(np.float64))


def test_host_lists_are_validated(sfb, ctx):
    """A neighbour index outside the node range (or a count above k) from the host is refused like the reference's
    out-of-bounds panic, instead of corrupting device memory downstream."""
    idx = np.array([[1, 2], [0, 7], [0, 1]], np.uint32); w = np.ones((3, 2)); cnt = np.array([2, 2, 2], np.uint32)
    with pytest.raises(sfb.SfbError):
        sfb.Adjacency.from_host(ctx, idx, w, cnt)
    with pytest.raises(sfb.SfbError):
        sfb.KnnGraph.from_host(ctx, idx, w, cnt)
    idx[1, 1] = 2
    with pytest.raises(sfb.SfbError):
        sfb.Adjacency.from_host(ctx, idx, w, np.array([2, 3, 2], np.uint32))
    idx[2, 1] = 0xFFFFFFFF   # padding past the count is fine
    sfb.Adjacency.from_host(ctx, idx, w, np.array([2, 2, 1], np.uint32)).laplacian()


# ---- successor Stage C: Bhattacharyya feature graph (f32 semantics, 1e-5 like the reference's own tests) ------
def _bc_state(seed, c, f):
    rng = np.random.default_rng(seed)
    means = rng.normal(size=(c, f)).astype(np.float32)
    variances = rng.uniform(0.05, 2.0, size=(c, f)).astype(np.float32)
    variances[0, :5] = 0.0          # below the variance floor
    means[:, 7] = means[:, 3]; variances[:, 7] = variances[:, 3]   # a duplicated feature: BC = 1, ties by index
    return means, variances


@pytest.mark.parametrize("c,f,k", [(40, 150, 15), (7, 33, 40), (300, 64, 5)])
def test_bc_adjacency_parity(sfb, oracle, ctx, c, f, k):
    means, variances = _bc_state(c + f, c, f)
    idx, w, cnt = sfb.bc_adjacency(means, variances, k, ctx=ctx).to_host()
    kk = min(k, f - 1)
    # bit for bit against the oracle's portable log / exp (the sequence the device evaluates): indices AND weights
    o_idx, o_w, o_cnt = oracle.bc_knn(means, variances, kk, det=True)
    assert idx.shape == (f, kk) and np.array_equal(cnt, o_cnt)
    assert np.array_equal(idx, o_idx) and np.array_equal(w, o_w.astype(np.float64))
    # ... and within 1 ulp (f32) of the libm form, the reference on glibc (its own tests compare at 1e-5)
    l_idx, l_w, l_cnt = oracle.bc_knn(means, variances, kk, det=False)
    np.testing.assert_allclose(w, l_w.astype(np.float64), rtol=1e-5, atol=1e-12)
    assert np.array_equal(cnt, l_cnt)


@pytest.mark.parametrize("normalize", [True, False])
def test_laplacian_stage_execute(sfb, oracle, ctx, normalize):
    """LaplacianStage::execute (laplacian.rs:135-219): BC kNN -> max-symmetrise -> L_sym or D - W."""
    means, variances = _bc_state(5, 30, 90)
    cfg = sfb.LaplacianConfig(k_neighbors=12, normalize=normalize)
    out = sfb.LaplacianStage(cfg).execute(means, variances, ctx=ctx)
    assert out.n_features == 90 and out.matrix.shape[0] == 90
    indptr, indices, data = out.matrix.to_host()
    o_idx, o_w, o_cnt = oracle.bc_knn(means, variances, 12, det=True)
    o_ptr, o_ind, o_dat = oracle.laplacian(o_idx, o_w.astype(np.float64), o_cnt, normalised=normalize, weight_threshold=1e-9)
    assert np.array_equal(indptr, o_ptr) and np.array_equal(indices, o_ind)
    np.testing.assert_allclose(data, o_dat, rtol=1e-12, atol=0)
    u_ptr, u_ind, u_dat = oracle.laplacian(o_idx, o_w.astype(np.float64), o_cnt, normalised=False)
    deg = np.array([u_dat[s:e][u_ind[s:e] == r][0] for r, (s, e) in enumerate(zip(u_ptr[:-1].astype(int), u_ptr[1:].astype(int)))])
    np.testing.assert_allclose(out.degrees, deg, rtol=1e-5)
    row_of = np.repeat(np.arange(90), np.diff(indptr.astype(np.int64)))
    assert np.all(data[indices != row_of] <= 0.0)   # off-diagonals <= 0 (tests/test_laplacian.rs:41-62)
    if not normalize:   # L = D - W: every row sums to 0 (tests/test_laplacian.rs:221-250, |sum| < 1e-4 in f32)
        assert np.all(np.abs(np.bincount(row_of, weights=data, minlength=90)) < 1e-4)
    if normalize:   # diag = 1, Rayleigh quotients in [0, 2] (tests/test_laplacian.rs:16-114)
        assert np.allclose(data[indices == row_of], 1.0)
        x = np.random.default_rng(0).normal(size=90)
        r = out.matrix.rayleigh_quotient(x)
        assert -1e-6 <= r <= 2.0 + 1e-6


# ---- energy pipeline: item -> sub-centroid mapping (energymaps.rs:1246-1342) ------------------------------------
def test_map_items_to_subcentroids(sfb, oracle, ctx):
    rng = np.random.default_rng(21)
    n, f, s = 5000, 24, 67
    x = rng.normal(size=(n, f)); x[11] = 0.0
    sc = rng.normal(size=(s, f)); sc[5] = 0.0
    sl = rng.uniform(0, 1, s)
    sl[40] = sl[7]; sl[41] = sl[7]; sl[12] = sl[3]          # exact lambda ties -> cosine tie-break
    il = rng.uniform(-0.1, 1.1, n)
    il[:200] = sl[7]; il[200:300] = sl[3]; il[300:320] = (sl[3] + sl[12]) / 2 + 5e-12
    sl[20] = 0.25; sl[21] = 0.75; il[320:340] = 0.5           # equidistant from two sub-centroids
    want = oracle.map_items(x, il, sc, sl)
    got = ctx.matrix(x).map_to_subcentroids(il, ctx.matrix(sc), sl)
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert len(set(want[0][:200])) > 1                        # the tie-break really decided


# ---- feature graph hidden behind the item screen (sfb_knn_build_columns_begin / _end) --------------------------
def test_knn_columns_begin_end(sfb, oracle, ctx):
    x = np.random.default_rng(31).normal(size=(6000, 40))
    x[:, 9] = x[:, 2]
    want_f = oracle.knn(oracle.transpose(x), 7, 0)
    want_i = oracle.knn(x, 5, 1)
    m = ctx.matrix(x)
    # a tensor-core screen runs between begin and end: the Gram tiles ride on the side stream beside it
    pend = m.knn_columns_begin(7, 0)
    g = m.knn(5, 1)
    assert g.stats()["screen_used"] == sfb.SCREEN_F16
    assert_knn_equal(pend.end().to_host(), want_f)
    assert_knn_equal(g.to_host(), want_i)
    # nothing in between: end() runs the job itself
    assert_knn_equal(m.knn_columns_begin(7, 0).end().to_host(), want_f)
    # an exact (no screen) kNN in between, other metric for the columns
    pend = m.knn_columns_begin(3, 2)
    g2 = m.knn(5, 1, screen=sfb.SCREEN_EXACT_F64, q_begin=0, q_end=300)
    assert_knn_equal(pend.end().to_host(), oracle.knn(oracle.transpose(x), 3, 2))
    assert_knn_equal(g2.to_host(), oracle.knn(x, 5, 1, query_rows=np.arange(300)))
    # not the feature-graph shape: end() takes the plain path
    y = np.random.default_rng(32).normal(size=(30, 500))
    assert_knn_equal(ctx.matrix(y).knn_columns_begin(4, 0).end().to_host(), oracle.knn(oracle.transpose(y), 4, 0))
    # a build that is worth hiding (long rows: the Gram fits behind the screen of the same matrix) occupies the context's
    # single side slot until its end(); small ones are simply deferred to end() and may be stacked
    p1 = m.knn_columns_begin(7, 0)
    p2 = m.knn_columns_begin(7, 0)
    assert_knn_equal(p2.end().to_host(), want_f)
    assert_knn_equal(p1.end().to_host(), want_f)
    big = ctx.generate(sfb.SYNTH_GAUSSIAN, 5, 700000, 48)
    q1 = big.knn_columns_begin(5, 0)
    with pytest.raises(sfb.SfbError):
        big.knn_columns_begin(5, 0)
    gq = big.knn(4, 0, q_begin=0, q_end=3000)
    assert_knn_equal(q1.end().to_host(), oracle.knn(oracle.transpose(oracle.generate_rows(0, 5, 0, 700000, 48)), 5, 0))
    gq.free()


# ---- JL projection ahead of lambda, SortedLambdas after it ---------------------------
@pytest.mark.parametrize("n,f,r", [(1000, 384, 91), (777, 50, 64), (130, 33, 1), (5, 700, 200), (2049, 128, 32)])
def test_project_rows_bit_exact(sfb, oracle, ctx, n, f, r):
    """Every projected entry is the reference's left fold ((x_i * s_ij) * scale added in i order): identical bits."""
    rng = np.random.default_rng(n + f + r)
    x = rng.standard_normal((n, f))
    x[0] = 0.0                                            # zero vector -> zeros (test_reduction.rs:60-69)
    s = rng.standard_normal((f, r))
    got = ctx.matrix(x).project(s).rows()
    want = oracle.project_rows(x, s)
    assert got.shape == (n, r) and np.array_equal(got, want)
    assert np.all(got[0] == 0.0)
    # linearity by an exact factor (test_reduction.rs:72-93)
    assert np.array_equal(ctx.matrix(2.0 * x).project(s).rows(), 2.0 * want)


def test_project_rows_core_f32(sfb, oracle, ctx):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((300, 96)).astype(np.float32)
    s = rng.standard_normal((40, 96)).astype(np.float32)  # reduced-major, as clustering.rs:94-105 draws
    got = ctx.matrix(x.astype(np.float64)).project(s.astype(np.float64), order=sfb.PROJECT_CORE_F32).rows()
    assert np.array_equal(got.astype(np.float32), oracle.project_rows_core(x, s)) and np.array_equal(got, got.astype(np.float32))
    with pytest.raises(ValueError):
        ctx.matrix(x.astype(np.float64)).project(np.zeros((95, 40)))


def test_projected_lambdas_match_reference_pipeline(sfb, oracle, ctx):
    """compute_synthetic_lambda on unprojected items (taumode.rs:261-318): tau = select_tau(item) and the zero-vector
    test use the UNPROJECTED item, the Rayleigh quotient and the dispersion the projected one, against the r x r
    Laplacian -- device pipeline against the oracle's, <= 1e-9 relative."""
    rng = np.random.default_rng(21)
    n, f, r = 3000, 128, sfb.compute_jl_dimension(40, 128, 0.5)
    assert r == oracle.jl_dimension(40, 128, 0.5) and 32 <= r < 128
    x = np.abs(rng.standard_normal((n, f)))          # positive entries: the median tau is well above the floor
    x[5] = 0.0                                        # zero vector -> lambda 0 without touching the graph
    x[6] = 1e-11                                      # |v| <= 1e-10 everywhere: also "zero" (taumode.rs:268-274)
    x[7, 1:] = 0.0                                    # one live entry: tau from the unprojected row's median (floor)
    s = rng.standard_normal((f, r))
    proj = sfb.ImplicitProjection(f, r, s)
    xm = ctx.matrix(x)
    y = sfb.project_matrix(xm, proj, ctx=ctx)
    L = y.knn_columns(6, 0).adjacency(2.0, 1.0).laplacian()
    yo = oracle.project_rows(x, s)
    idx, dist, cnt = oracle.knn(oracle.transpose(yo), 6, 0)
    a = oracle.build_adjacency(idx, dist, cnt, 2.0, 1.0)
    ip, ind, dat = oracle.laplacian(a[0], a[1], a[2])
    for mode, val in ((sfb.TAU_MEDIAN, 0.0), (sfb.TAU_MEAN, 0.0), (sfb.TAU_FIXED, 0.4), (sfb.TAU_PERCENTILE, 0.9)):
        got, stats = L.lambdas_projected(xm, y, tau_mode=mode, tau_value=val)
        want = oracle.lambdas_projected(ip, ind, dat, yo, x, tau_mode=mode, tau_value=val)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-14)
        assert got[5] == 0.0 and got[6] == 0.0 and want[5] == 0.0
        # tau really comes from the unprojected rows: taking it from the projected ones gives different lambdas
        if mode != sfb.TAU_FIXED:
            assert not np.allclose(got, oracle.lambdas(ip, ind, dat, yo, tau_mode=mode, tau_value=val), rtol=1e-6)
    got_n, stats = L.lambdas_projected(xm, y, normalise=True)
    want_n, wstats = oracle.normalise_lambdas(oracle.lambdas_projected(ip, ind, dat, yo, x))
    np.testing.assert_allclose(got_n, want_n, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(stats, wstats, rtol=1e-9)
    assert np.array_equal(proj.project(x[3], ctx=ctx), yo[3])
    # already-projected items (len == reduced_dim) go through the plain call; a mismatched length is refused
    # like the reference's panic ("item seems neither projected nor unprojected", taumode.rs:287-297)
    with pytest.raises(sfb.SfbError):
        L.lambdas_projected(xm, xm)
    with pytest.raises(sfb.SfbError):
        L.lambdas_projected(ctx.matrix(x[:10]), y)


def test_jl_dimension_matches_oracle(sfb, oracle):
    for n, d, e in [(100, 16, 0.3), (10, 100, 0.3), (2, 1000, 0.9), (1000, 512, 0.1), (100, 2000, 0.2), (10_000, 5_000, 0.3),
                    (5_000, 3_000, 0.3), (100, 100_000, 0.3), (1, 100, 0.1), (0, 64, 0.3), (100, 2048, 0.3), (100, 2049, 0.3)]:
        assert sfb.compute_jl_dimension(n, d, e) == oracle.jl_dimension(n, d, e)
        assert sfb.compute_jl_dimension(n, d, e, core=True) == oracle.jl_dimension(n, d, e, core=True)


@pytest.mark.parametrize("n", [1, 13, 2048, 2049, 100_003, 1_000_000])
def test_sorted_lambdas_parity(sfb, oracle, ctx, n):
    """Order (lambda, decimal string of idx), bucket keys and the f32 std_dev, identical to the oracle's BTreeMap
    restatement; heavy ties (quantised lambdas, the zero bucket with both signs) exercise the string order."""
    rng = np.random.default_rng(n)
    lam = np.round(rng.random(n), 3 if n > 100 else 1)
    lam[rng.random(n) < 0.05] = 0.0
    lam[rng.random(n) < 0.01] = -0.0
    if n > 12:
        lam[7], lam[11] = 1.0, -0.25
    sl = sfb.SortedLambdas().build_from(lam, ctx=ctx)
    want, widx, wsd = oracle.sorted_lambdas(lam)
    assert np.array_equal(sl.indices, widx)
    assert np.array_equal(sl.lambdas.view(np.uint64), want.view(np.uint64))
    assert sl.std_dev == wsd
    assert sl.to_vec()[:3] == list(zip(want[:3].tolist(), widx[:3].tolist()))


def test_sorted_lambdas_nan_and_range(sfb, oracle, ctx):
    lam = np.array([math.nan, 0.3, math.nan, math.inf, 0.3, -1.0])
    sl = sfb.SortedLambdas().build_from(lam, ctx=ctx)
    want, widx, _ = oracle.sorted_lambdas(lam)
    assert np.array_equal(sl.indices, widx) and list(widx) == [5, 1, 4, 3, 0, 2]
    assert np.array_equal(np.isnan(sl.lambdas), np.isnan(want))
    with pytest.raises(sfb.SfbError):
        sfb.SortedLambdas().build_from(np.zeros(0), ctx=ctx)
    # range_bylambda (sorted_index.rs:60-79): first k inside [q - std/2^p, q + std/2^p] in index order of the map
    lam = np.random.default_rng(4).random(5000)
    sl = sfb.SortedLambdas().build_from(lam, ctx=ctx)
    band = sl.std_dev / 2.0 ** 2.0
    inside = [(int(i), float(lam[i])) for i in np.argsort(lam, kind="stable") if 0.5 - band <= lam[i] <= 0.5 + band]
    assert sl.range_bylambda(0.5, 10, 2.0) == inside[:10]
