"""Generates the golden fixtures of the graph-wiring path from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference is Rust and cannot run here, so these
vectors freeze the oracle -- itself pinned to the reference's known-answer tests (tests/test_oracle_kat.py) --
at small sizes; tests/test_golden.py checks both the oracle and the CUDA path against them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (kind, seed, rows, cols, centres, noise, metric, k, p, sigma)
    "gauss_cos": (0, 42, 96, 24, 0, 0.0, 0, 6, 2.0, 1.0),
    "clustered_l2": (1, 7, 120, 16, 5, 0.3, 1, 12, 2.0, 0.5),
    "aniso_l2sq": (2, 13, 80, 40, 0, 0.1, 2, 4, 1.0, 2.0),
}


def build(case):
    kind, seed, rows, cols, centres, noise, metric, k, p, sigma = case
    x = oracle.generate_rows(kind, seed, 0, rows, cols, centres, noise)
    out = {"x": x}
    idx, dist, cnt = oracle.knn(x, k, metric)
    out.update(knn_idx=idx, knn_dist=dist, knn_cnt=cnt)
    a_idx, a_w, a_cnt, applied = oracle.build_adjacency(idx, dist, cnt, p, sigma)
    out.update(adj_idx=a_idx, adj_w=a_w, adj_cnt=a_cnt, adj_sparsified=np.array([applied]))
    ptr, ind, dat = oracle.laplacian(a_idx, a_w, a_cnt)
    out.update(lap_indptr=ptr, lap_indices=ind, lap_data=dat)
    nptr, nind, ndat = oracle.laplacian(a_idx, a_w, a_cnt, normalised=True)
    out.update(nlap_indptr=nptr, nlap_indices=nind, nlap_data=ndat)
    # feature graph (nodes = columns, graph.rs:214-216) + per-item lambda in the three variants
    xt = oracle.transpose(x)
    kf = min(3, cols - 1)
    f_idx, f_dist, f_cnt = oracle.knn(xt, kf, oracle.METRIC_COSINE)
    fa = oracle.build_adjacency(f_idx, f_dist, f_cnt, p, sigma)
    fptr, find, fdat = oracle.laplacian(*fa[:3])
    out.update(f_idx=f_idx, f_dist=f_dist, f_cnt=f_cnt, flap_indptr=fptr, flap_indices=find, flap_data=fdat)
    for name, variant in (("legacy", oracle.LAMBDA_LEGACY_TAUMODE), ("energy", oracle.LAMBDA_ENERGY_NODE)):
        lam, e, g = oracle.lambdas(fptr, find, fdat, x, variant, oracle.TAU_MEDIAN, with_parts=True)
        out["lambda_" + name] = lam
        out["disp_" + name] = g
    out["lambda_legacy_norm"], out["lambda_stats"] = oracle.normalise_lambdas(out["lambda_legacy"])
    out["diffused"] = oracle.diffuse(fptr, find, fdat, x, 0.1, 4)
    return out


def build_post():
    """The steps either side of lambda: JL projection, taumode lambda of projected items (tau and
    the zero test from the unprojected rows), SortedLambdas.  The projection draws are Philox Gaussians here (the
    reference's ChaCha8 + ziggurat stream is made by the Rust wrapper)."""
    n, f, r = 150, 48, 12
    x = np.abs(oracle.generate_rows(1, 23, 0, n, f, 6, 0.4))
    x[11] = 0.0
    samples = oracle.generate_rows(0, 99, 0, f, r)
    y = oracle.project_rows(x, samples)
    out = {"x": x, "samples": samples, "projected": y,
           "jl": np.array([oracle.jl_dimension(a, b, c) for a, b, c in POST_JL], np.uint64)}
    idx, dist, cnt = oracle.knn(oracle.transpose(y), 3, oracle.METRIC_COSINE)
    a = oracle.build_adjacency(idx, dist, cnt, 2.0, 1.0)
    ptr, ind, dat = oracle.laplacian(a[0], a[1], a[2])
    out.update(plap_indptr=ptr, plap_indices=ind, plap_data=dat)
    lam = oracle.lambdas_projected(ptr, ind, dat, y, x, oracle.TAU_MEDIAN)
    out["lambda_projected"] = lam
    lam_n, _ = oracle.normalise_lambdas(lam)
    q = np.round(lam_n, 2)                      # quantised: many equal lambdas, ordered by the decimal string of the index
    srt, order, sd = oracle.sorted_lambdas(q)
    out.update(lambda_quantised=q, sorted_lambda=srt, sorted_idx=order, std_dev=np.array([sd]))
    return out


POST_JL = [(100, 16, 0.3), (10, 100, 0.3), (2, 1000, 0.9), (1000, 512, 0.1), (100, 2000, 0.2), (10000, 5000, 0.3), (100, 100000, 0.3)]


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "post_steps.npz"), **build_post())
    print("wrote post_steps")
    for name, case in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(case))
        print("wrote", name)
