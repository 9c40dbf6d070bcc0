"""Persistence wire format (src_legacy/storage/parquet.rs): schema, Snappy, CSR <-> COO round trip."""
import numpy as np
import pytest


def test_sparse_matrix_parquet_round_trip(sfb, oracle, tmp_path):
    pq = pytest.importorskip("pyarrow.parquet")
    from matternet_rs_b200 import storage
    x = oracle.generate_rows(0, 3, 0, 60, 9)
    idx, dist, cnt = oracle.knn(x, 4, 0)
    a = oracle.build_adjacency(idx, dist, cnt, 2.0, 1.0)
    indptr, indices, data = oracle.laplacian(*a[:3])
    f = storage.save_sparse_matrix(indptr, indices, data, str(tmp_path), "laplacian")
    t = pq.read_table(f)
    assert t.column_names == ["name_id", "n_rows", "n_cols", "nnz", "row", "col", "value"]          # parquet.rs:435-443
    assert [str(c.type) for c in t.columns] == ["string", "uint64", "uint64", "uint64", "uint64", "uint64", "double"]
    assert pq.ParquetFile(f).metadata.row_group(0).column(6).compression == "SNAPPY"
    assert t.num_rows == len(indices) and t["name_id"][0].as_py() == "laplacian" and t["nnz"][0].as_py() == len(indices)
    ip, ix, dv, shape = storage.load_sparse_matrix(f)
    assert shape == (60, 60) and np.array_equal(ip, indptr) and np.array_equal(ix, indices) and np.array_equal(dv, data)


def test_lambda_parquet_round_trip(sfb, tmp_path):
    pq = pytest.importorskip("pyarrow.parquet")
    from matternet_rs_b200 import storage
    lam = np.random.default_rng(0).uniform(size=1000)
    f = storage.save_lambda(lam, str(tmp_path), "lambdas")
    t = pq.read_table(f)
    assert t.column_names == ["name_id", "n_values", "row_index", "lambda"]                          # parquet.rs:743-748
    assert np.array_equal(storage.load_lambda(f), lam)
    with pytest.raises(ValueError):
        storage.save_lambda([], str(tmp_path), "empty")


def test_duplicate_triplets_are_summed_like_trimat_to_csr(sfb, tmp_path):
    pa = pytest.importorskip("pyarrow")
    pq = pytest.importorskip("pyarrow.parquet")
    from matternet_rs_b200 import storage
    rows = np.array([0, 1, 1, 0, 2, 1], np.uint64); cols = np.array([1, 0, 0, 0, 2, 2], np.uint64)
    vals = np.array([1.0, 2.0, 0.5, 3.0, 4.0, 5.0])
    t = pa.table({"name_id": ["m"] * 6, "n_rows": np.full(6, 3, np.uint64), "n_cols": np.full(6, 3, np.uint64),
                  "nnz": np.full(6, 6, np.uint64), "row": rows, "col": cols, "value": vals})
    f = str(tmp_path / "m.parquet")
    pq.write_table(t, f)
    ip, ix, dv, shape = storage.load_sparse_matrix(f)
    assert shape == (3, 3) and ip.tolist() == [0, 2, 4, 5] and ix.tolist() == [0, 1, 0, 2, 2] and dv.tolist() == [3.0, 1.0, 2.5, 5.0, 4.0]
    bad = t.set_column(3, "nnz", pa.array(np.full(6, 7, np.uint64)))
    pq.write_table(bad, f)
    with pytest.raises(ValueError):
        storage.load_sparse_matrix(f)


def test_metadata_sidecar_matches_the_reference_layout(sfb, tmp_path):
    """`<name_id>_metadata.json` (parquet.rs:32-56,131-145,486-503,788-806): serde's externally tagged ConfigValue."""
    pytest.importorskip("pyarrow")
    import json
    from matternet_rs_b200 import storage
    b = sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 8, 4, 2.0, None)
    cfg = {"lambda_eps": b.lambda_eps, "lambda_k": b.lambda_k, "lambda_topk": b.lambda_topk, "lambda_p": b.lambda_p,
           "lambda_sigma": {"OptionF64": b.lambda_sigma}, "normalise": b.normalise, "synthesis": {"TauMode": "Median"}}
    indptr = np.array([0, 1, 3], np.uint64); indices = np.array([0, 0, 1], np.uint32); data = np.array([1.0, -1.0, 1.0])
    storage.save_sparse_matrix(indptr, indices, data, str(tmp_path), "lap", builder_config=cfg)
    m = json.load(open(tmp_path / "lap_metadata.json"))
    assert set(m) == {"name_id", "timestamp", "n_rows", "n_cols", "builder_config", "files"}
    assert m["builder_config"]["lambda_eps"] == {"F64": 0.5} and m["builder_config"]["lambda_k"] == {"Usize": 8}
    assert m["builder_config"]["normalise"] == {"Bool": False} and m["builder_config"]["lambda_sigma"] == {"OptionF64": None}
    assert m["files"]["matrix"]["file_type"] == "sparse" and m["files"]["matrix"]["nnz"] == 3 and m["files"]["matrix"]["filename"] == "lap.parquet"
    storage.save_lambda([0.1, 0.2], str(tmp_path), "lam", builder_config=cfg)
    m = storage.load_metadata(str(tmp_path), "lam")
    assert m["n_rows"] == 2 and m["n_cols"] == 1 and m["files"]["lambda_vector"]["nnz"] is None
