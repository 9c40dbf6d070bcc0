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
