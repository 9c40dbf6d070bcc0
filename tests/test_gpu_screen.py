"""The tensor-core screen: tcgen05 accumulators against the operands it was fed, and end-to-end
neighbour parity (bit-exact) with the brute-force oracle for every screen precision."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(sfb):
    c = sfb.Context(0)
    yield c
    c.close()


def assert_knn_equal(got, want):
    assert np.array_equal(got[2], want[2]), "neighbour counts differ"
    assert np.array_equal(got[0], want[0]), "neighbour indices differ"
    assert np.array_equal(got[1], want[1]), "distances differ in bits"


@pytest.mark.parametrize("screen", [2, 3])
@pytest.mark.parametrize("m,kd,row0,col0", [(700, 64, 0, 0), (1500, 200, 256, 512), (900, 384, 640, 256), (300, 3072, 128, 0)])
def test_tile_accumulators_match_operands(sfb, ctx, screen, m, kd, row0, col0):
    """S~ = Q Q^T on the tensor cores vs the same 16-bit operands multiplied in f64, within the
    accumulation bound gamma*|q_i||q_j| the certification assumes (and far inside it in practice)."""
    x = np.random.default_rng(m + kd).normal(size=(m, kd))
    tile, qr, qc, scale = ctx.matrix(x).debug_screen_tile(sfb.METRIC_COSINE, screen, row0, col0)
    ref = qr.astype(np.float64) @ qc.astype(np.float64).T
    nr = np.linalg.norm(qr.astype(np.float64), axis=1)
    nc = np.linalg.norm(qc.astype(np.float64), axis=1)
    kpad = qr.shape[1]
    gamma = (kpad + 64) * 2.0 ** -23
    err = np.abs(tile.astype(np.float64) - ref)
    bound = gamma * np.outer(nr, nc) + 1e-30
    assert np.all(err <= bound), f"max err/bound {np.max(err / bound):.3f}"
    # operands are the scaled unit rows
    assert scale == 64.0
    rows = x[row0:row0 + 128]
    unit = rows / np.linalg.norm(rows, axis=1, keepdims=True) * scale
    got = qr[:len(rows), :kd]
    assert np.allclose(got, unit, rtol=2.0 ** -7 if screen == 3 else 2.0 ** -10, atol=1e-4)
    # padding rows / columns are zero
    assert np.all(qr[:, kd:] == 0)
    print(f"screen={screen} kd={kd}: max err/bound = {np.max(err / bound):.4f}")


@pytest.mark.parametrize("screen", [2, 3])
@pytest.mark.parametrize("metric", [0, 1, 2])
@pytest.mark.parametrize("m,kd,k", [(5000, 96, 16), (4097, 50, 8), (6000, 384, 32), (4500, 700, 8), (4200, 1030, 5)])
def test_knn_screen_parity_gaussian(sfb, oracle, ctx, screen, metric, m, kd, k):
    x = np.random.default_rng(m + kd + metric).normal(size=(m, kd))
    g = ctx.matrix(x).knn(k, metric, screen=screen)
    assert_knn_equal(g.to_host(), oracle.knn(x, k, metric))
    st = g.stats()
    assert st["screen_used"] == screen and st["rows_certified"] + st["rows_fallback"] == m
    print(st)


@pytest.mark.parametrize("gather", ["bulk2", "bulk3", "lsu2", "lsu3", "half2", "half3"])
@pytest.mark.parametrize("metric", [0, 1])
def test_knn_screen_parity_clustered(sfb, oracle, ctx, metric, gather, monkeypatch):
    """Tight clusters: neighbour gaps comparable to the fp16 margin -> many rows fall back; still exact.  Both gather paths
    of the rescore kernel (cp.async.bulk row chunks / 8-byte cp.async, the latter with 32- and 16-dimension chunks) at both
    ring depths."""
    if gather.startswith("bulk"):
        monkeypatch.setenv("SFB_RESCORE_BULK", "1")
    monkeypatch.setenv("SFB_RESCORE_CH", "16" if gather.startswith("half") else "32")
    monkeypatch.setenv("SFB_RESCORE_NST", gather[-1])
    m = ctx.generate(sfb.SYNTH_CLUSTERED, 7, 20000, 128, 32, 0.3)
    x = oracle.generate_rows(1, 7, 0, 20000, 128, 32, 0.3)
    g = m.knn(16, metric, screen=sfb.SCREEN_F16)
    assert_knn_equal(g.to_host(), oracle.knn(x, 16, metric))
    print(g.stats())


def test_knn_screen_duplicates_zero_rows_eps(sfb, oracle, ctx):
    rng = np.random.default_rng(3)
    base = rng.normal(size=(3000, 40))
    x = np.concatenate([base, base[:1500], np.zeros((7, 40))])  # exact ties and zero rows -> fallback rows
    for metric in (0, 1):
        for eps in (math.inf, 0.7):
            g = ctx.matrix(x).knn(10, metric, eps=eps, screen=sfb.SCREEN_F16)
            assert_knn_equal(g.to_host(), oracle.knn(x, 10, metric, eps))
            if metric == 0 and eps == math.inf:
                # a zero row ties with everything at distance 1: the margin cannot separate -> exact fallback
                assert g.stats()["rows_fallback"] >= 7


def test_knn_screen_query_shard_and_kprime(sfb, oracle, ctx):
    x = np.random.default_rng(5).normal(size=(7000, 72))
    mat = ctx.matrix(x)
    g = mat.knn(12, 0, screen=sfb.SCREEN_F16, q_begin=1000, q_end=3333)
    assert_knn_equal(g.to_host(), oracle.knn(x, 12, 0, query_rows=np.arange(1000, 3333)))
    for kp in (16, 32, 192):
        g = mat.knn(12, 0, screen=sfb.SCREEN_F16, k_prime=kp, q_begin=0, q_end=2000)
        assert_knn_equal(g.to_host(), oracle.knn(x, 12, 0, query_rows=np.arange(0, 2000)))
    # a zero row is at cosine distance exactly 1 from every row: no margin separates its k-th neighbour
    # from the dropped candidates, so without the exact fallback the call refuses
    with pytest.raises(sfb.SfbError):
        ctx.matrix(np.concatenate([x[:3000], np.zeros((3, 72))])).knn(4, 0, screen=sfb.SCREEN_F16, allow_fallback=False)
    # duplicated rows certify as long as the k-th gap is wide: both copies carry the same screen key
    g = ctx.matrix(np.concatenate([x[:3000], x[:3000]])).knn(4, 0, screen=sfb.SCREEN_F16, allow_fallback=False)
    assert_knn_equal(g.to_host(), oracle.knn(np.concatenate([x[:3000], x[:3000]]), 4, 0))


def test_knn_auto_uses_screen(sfb, oracle, ctx):
    x = np.random.default_rng(6).normal(size=(4500, 64))
    g = ctx.matrix(x).knn(16, 0)
    assert g.stats()["screen_used"] == sfb.SCREEN_F16
    assert_knn_equal(g.to_host(), oracle.knn(x, 16, 0))


def test_knn_screen_edge_shapes(sfb, oracle, ctx):
    """Ragged and extreme shapes through the screen: k at the ABI maximum, more neighbours asked than rows can supply
    under eps (counts < k, padded with IDX_NONE / +inf), a one-row query shard, a corpus that is not a multiple of the
    256-row tile, K = 1 and a K that is not a multiple of 64."""
    rng = np.random.default_rng(11)
    x = rng.normal(size=(4321, 70))
    mat = ctx.matrix(x)
    # k = 128 (ABI maximum): k' = 192 leaves k + 64
    assert_knn_equal(mat.knn(128, 1, screen=sfb.SCREEN_F16, q_begin=100, q_end=400).to_host(),
                     oracle.knn(x, 128, 1, query_rows=np.arange(100, 400)))
    # tight eps: most rows keep fewer than k neighbours
    d = oracle.knn(x, 8, 0)[1]
    eps = float(np.quantile(d[:, 3], 0.5))
    got = mat.knn(8, 0, eps=eps, screen=sfb.SCREEN_F16).to_host()
    want = oracle.knn(x, 8, 0, eps)
    assert_knn_equal(got, want)
    assert (got[2] < 8).any() and np.all(got[0][got[2] == 0] == sfb.IDX_NONE if (got[2] == 0).any() else True)
    # single query row, last row of the matrix
    assert_knn_equal(mat.knn(5, 2, screen=sfb.SCREEN_F16, q_begin=4320, q_end=4321).to_host(),
                     oracle.knn(x, 5, 2, query_rows=np.array([4320])))
    # one-dimensional rows: cosine is +-1 everywhere (ties by index), L2 is |a - b|
    y = rng.normal(size=(4200, 1))
    for metric in (0, 1):
        assert_knn_equal(ctx.matrix(y).knn(4, metric, screen=sfb.SCREEN_F16).to_host(), oracle.knn(y, 4, metric))


def test_knn_screen_non_finite_input(sfb, ctx):
    """The reference would propagate NaN through its sort; the screen refuses non-finite rows for L2 instead of
    returning an arbitrary order (cosine rows with non-finite norms become zero operands and fall back)."""
    x = np.random.default_rng(12).normal(size=(4200, 16))
    x[17, 3] = np.inf
    with pytest.raises(sfb.SfbError):
        ctx.matrix(x).knn(4, 1, screen=sfb.SCREEN_F16)


def test_knn_three_levels(sfb, oracle, ctx):
    """Tight clusters with a deliberately small k': level 1 leaves many rows uncertified, the k' = 192 re-screen
    certifies most of them, the rest go to f64 brute force -- and the result is still the oracle's, bit for bit."""
    m = ctx.generate(sfb.SYNTH_CLUSTERED, 3, 30000, 64, 16, 0.05)
    x = oracle.generate_rows(1, 3, 0, 30000, 64, 16, 0.05)
    g = m.knn(8, 0, screen=sfb.SCREEN_F16, k_prime=9)
    st = g.stats()
    assert st["rows_rescreened"] >= 32 and st["rows_fallback"] < st["rows_rescreened"]
    assert st["rows_certified"] + st["rows_fallback"] == 30000
    assert_knn_equal(g.to_host(), oracle.knn(x, 8, 0))
    print(st)


def test_knn_screen_many_exact_duplicates(sfb, oracle, ctx):
    """A cluster of exact duplicates puts more equal keys into a row's candidate buffer than the quantised prune can
    separate: the prune must still leave room for the next chunk's appends (it switches to the exact select), and the
    result is the oracle's -- ties by index -- for the duplicates, their neighbours and every other row."""
    rng = np.random.default_rng(15)
    base = rng.normal(size=(4500, 48))
    hot = base[7] + 0.01 * rng.normal(size=48)
    x = np.concatenate([base[:2000], np.tile(hot, (150, 1)), base[2000:], np.tile(base[11], (90, 1))])
    for metric in (0, 1):
        g = ctx.matrix(x).knn(16, metric, screen=sfb.SCREEN_F16)
        assert_knn_equal(g.to_host(), oracle.knn(x, 16, metric))
        st = g.stats()
        assert st["rows_certified"] + st["rows_fallback"] == x.shape[0]


# ---- the accumulation-error model of the certificate under adversarial operands -----------------------------------
def _adversarial_rows(case, m, kd, rng):
    """Rows that stress fp32 accumulation in the tensor core (cosine metric: the operands are the unit rows x 64).
      cancel   products +c for the first half of the dimensions, -c for the second: partial sums reach K*c/2, the dot is ~0
      range    every row spans 2^14 in magnitude (the widest spread fp16 operands keep after the x 64 scaling)
      drop     one dominant coordinate A and K-1 coordinates t with t*t just under one ulp of A*A: an accumulator that
               truncates each addend to the running sum's exponent loses every one of them (the worst case the model allows)
      drop16   the same with t*t = 0.9 ulp / 16: sixteen products (one MMA instruction's worth) together stay under an ulp"""
    if case == "cancel":
        x = np.ones((m, kd)) + rng.normal(size=(m, kd)) * 1e-3
        x[1::2, kd // 2:] *= -1.0
        return x
    if case == "range":
        return np.sign(rng.normal(size=(m, kd))) * 2.0 ** rng.uniform(-7, 7, size=(m, kd))
    t = math.sqrt(0.9 * 2.0 ** -23 / (16.0 if case == "drop16" else 1.0))
    x = np.full((m, kd), t) * (1.0 + rng.uniform(0, 0.05, size=(m, kd)))
    x[:, 0] = 1.0
    x[2::3, 0] = -1.0   # negative dots too
    return x


@pytest.mark.parametrize("screen", [2, 3])
@pytest.mark.parametrize("kd", [16, 384, 3072])
@pytest.mark.parametrize("case", ["cancel", "range", "drop", "drop16"])
def test_tile_accumulators_adversarial(sfb, ctx, screen, kd, case):
    """|S~ - q_i.q_j| <= gamma |q_i||q_j| with gamma = (kpad + 64) 2^-23 (knn_screen.cu: screen_level) must hold for operands
    built to maximise the accumulation error, not only for Gaussian rows; the worst err / bound is printed."""
    rng = np.random.default_rng(kd + len(case))
    x = _adversarial_rows(case, 512, kd, rng)
    worst = 0.0
    for row0, col0 in ((0, 0), (128, 256)):
        tile, qr, qc, scale = ctx.matrix(x).debug_screen_tile(sfb.METRIC_COSINE, screen, row0, col0)
        ref = qr.astype(np.float64) @ qc.astype(np.float64).T
        nr = np.linalg.norm(qr.astype(np.float64), axis=1)
        nc = np.linalg.norm(qc.astype(np.float64), axis=1)
        gamma = (qr.shape[1] + 64) * 2.0 ** -23
        bound = gamma * np.outer(nr, nc) + 1e-30
        worst = max(worst, float(np.max(np.abs(tile.astype(np.float64) - ref) / bound)))
    print(f"ACCUM case={case} screen={screen} kd={kd}: worst err/bound = {worst:.4f}")
    assert worst <= 0.5, f"accumulation error reaches {worst:.3f} of the certificate's bound: widen gamma"


@pytest.mark.parametrize("screen", [2, 3])
@pytest.mark.parametrize("metric", [0, 1, 2])
@pytest.mark.parametrize("case,kd", [("cancel", 384), ("range", 384), ("drop", 384), ("range", 16), ("drop", 3072), ("cancel", 1030)])
def test_knn_screen_parity_adversarial(sfb, oracle, ctx, screen, metric, case, kd):
    """End to end on the same operands: whatever the screen drops or mis-ranks, the certificate must notice (rows it
    cannot certify go to the exact path), so the lists equal the brute-force oracle bit for bit."""
    rng = np.random.default_rng(kd * 3 + metric)
    m = 4300 if kd <= 1030 else 4100
    x = _adversarial_rows(case, m, kd, rng)
    x[:64] = x[64:128]                      # exact duplicates: ties by index
    g = ctx.matrix(x).knn(8, metric, screen=screen)
    assert_knn_equal(g.to_host(), oracle.knn(x, 8, metric))
    st = g.stats()
    assert st["rows_certified"] + st["rows_fallback"] == m
    print(case, kd, metric, {k: st[k] for k in ("rows_certified", "rows_fallback", "rows_rescreened", "max_margin")})
