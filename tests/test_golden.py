"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle at small sizes).

CPU: the oracle still reproduces them bit for bit (they freeze it against drift).
GPU: the CUDA path, through the C ABI, reproduces them: indices / counts / CSR structure and distances bit-exact,
weights, Laplacian values and lambda within 1e-9 relative (north_star)."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)
CASES = make_golden.CASES
RTOL = 1e-9


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle, name):
    want = load(name)
    got = make_golden.build(CASES[name])
    assert sorted(got) == sorted(want)
    for key in want:
        assert np.array_equal(np.asarray(got[key]), want[key]), key


@pytest.mark.gpu
@pytest.mark.parametrize("screen", [1, 2])   # exact f64 brute force, tensor-core screen
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_golden(sfb, name, screen):
    kind, seed, rows, cols, centres, noise, metric, k, p, sigma = CASES[name]
    g = load(name)
    ctx = sfb.Context(0)
    x = ctx.generate(kind, seed, rows, cols, centres, noise)
    assert np.array_equal(x.rows(), g["x"])
    knn = x.knn(k, metric, screen=screen)
    idx, dist, cnt = knn.to_host()
    assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(cnt, g["knn_cnt"]) and np.array_equal(dist, g["knn_dist"])
    adj = knn.adjacency(p, sigma)
    a_idx, a_w, a_cnt = adj.to_host()
    assert adj.sparsified == bool(g["adj_sparsified"][0])
    assert np.array_equal(a_idx, g["adj_idx"]) and np.array_equal(a_cnt, g["adj_cnt"])
    np.testing.assert_allclose(a_w, g["adj_w"], rtol=RTOL, atol=0)
    for normalised, pre in ((False, "lap"), (True, "nlap")):
        ptr, ind, dat = adj.laplacian(normalised=normalised).to_host()
        assert np.array_equal(ptr, g[pre + "_indptr"]) and np.array_equal(ind, g[pre + "_indices"])
        np.testing.assert_allclose(dat, g[pre + "_data"], rtol=RTOL, atol=1e-300)
    # feature graph over the columns + lambda
    gf = x.knn_columns(min(3, cols - 1), sfb.METRIC_COSINE)
    f_idx, f_dist, f_cnt = gf.to_host()
    assert np.array_equal(f_idx, g["f_idx"]) and np.array_equal(f_cnt, g["f_cnt"]) and np.array_equal(f_dist, g["f_dist"])
    Lf = gf.adjacency(p, sigma).laplacian()
    fptr, find, fdat = Lf.to_host()
    assert np.array_equal(fptr, g["flap_indptr"]) and np.array_equal(find, g["flap_indices"])
    np.testing.assert_allclose(fdat, g["flap_data"], rtol=RTOL, atol=1e-300)
    lam, disp, _ = Lf.lambdas(x, sfb.LAMBDA_LEGACY_TAUMODE, sfb.TAU_MEDIAN, with_dispersion=True)
    np.testing.assert_allclose(lam, g["lambda_legacy"], rtol=RTOL, atol=1e-15)
    lam_e, disp_e, _ = Lf.lambdas(x, sfb.LAMBDA_ENERGY_NODE, sfb.TAU_MEDIAN, with_dispersion=True)
    np.testing.assert_allclose(lam_e, g["lambda_energy"], rtol=RTOL, atol=1e-15)
    np.testing.assert_allclose(disp_e, g["disp_energy"], rtol=RTOL, atol=1e-15)
    lam_n, stats = Lf.lambdas(x, sfb.LAMBDA_LEGACY_TAUMODE, sfb.TAU_MEDIAN, normalise=True)
    np.testing.assert_allclose(lam_n, g["lambda_legacy_norm"], rtol=RTOL, atol=1e-15)
    np.testing.assert_allclose(stats, g["lambda_stats"], rtol=RTOL, atol=1e-15)
    x.diffuse(Lf, 0.1, 4)
    np.testing.assert_allclose(x.rows(), g["diffused"], rtol=RTOL, atol=1e-15)
    ctx.close()


def test_oracle_reproduces_post_steps_golden(oracle):
    want = load("post_steps")
    got = make_golden.build_post()
    assert sorted(got) == sorted(want)
    for key in want:
        assert np.array_equal(np.asarray(got[key]), want[key]), key


@pytest.mark.gpu
def test_cuda_path_reproduces_post_steps_golden(sfb):
    """JL projection (bit-exact), lambda of the projected items, SortedLambdas order / keys / std_dev against the fixture."""
    g = load("post_steps")
    ctx = sfb.Context(0)
    x = ctx.matrix(g["x"])
    y = x.project(g["samples"])
    assert np.array_equal(y.rows(), g["projected"])
    assert [sfb.compute_jl_dimension(a, b, c) for a, b, c in make_golden.POST_JL] == g["jl"].tolist()
    L = y.knn_columns(3, sfb.METRIC_COSINE).adjacency(2.0, 1.0).laplacian()
    ptr, ind, dat = L.to_host()
    assert np.array_equal(ptr, g["plap_indptr"]) and np.array_equal(ind, g["plap_indices"])
    np.testing.assert_allclose(dat, g["plap_data"], rtol=RTOL, atol=1e-300)
    lam, _ = L.lambdas_projected(x, y, tau_mode=sfb.TAU_MEDIAN)
    np.testing.assert_allclose(lam, g["lambda_projected"], rtol=RTOL, atol=1e-15)
    assert lam[11] == 0.0
    sl = sfb.SortedLambdas().build_from(g["lambda_quantised"], ctx=ctx)
    assert np.array_equal(sl.indices, g["sorted_idx"]) and np.array_equal(sl.lambdas, g["sorted_lambda"])
    assert sl.std_dev == float(g["std_dev"][0])
    ctx.close()
