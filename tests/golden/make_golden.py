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


if __name__ == "__main__":
    for name, case in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(case))
        print("wrote", name)
