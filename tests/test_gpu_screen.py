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


@pytest.mark.parametrize("metric", [0, 1])
def test_knn_screen_parity_clustered(sfb, oracle, ctx, metric):
    """Tight clusters: neighbour gaps comparable to the fp16 margin -> many rows fall back; still exact."""
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
