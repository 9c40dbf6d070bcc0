"""Wire format of the path's outputs: the reference's Parquet layout (src_legacy/storage/parquet.rs), so that a
Laplacian and a lambda vector built here can be read by the reference's loaders and vice versa.

    save_sparse_matrix / load_sparse_matrix   parquet.rs:412-520,520-590: COO rows
        name_id Utf8 | n_rows u64 | n_cols u64 | nnz u64 | row u64 | col u64 | value f64      (Snappy, one row per non-zero,
        in CSR order: row-major, ascending column)
    save_lambda / load_lambda                 parquet.rs:728-826:
        name_id Utf8 | n_values u64 | row_index u64 | lambda f64
Files are `<path>/<name_id>.parquet`.  Host-side I/O only (the reference's persistence is not on the compute path)."""
import os

import numpy as np


def _pa():
    import pyarrow as pa
    import pyarrow.parquet as pq
    return pa, pq


def save_sparse_matrix(indptr, indices, data, path, name_id, n_cols=None):
    """CSR (as returned by Csr.to_host()) -> `<path>/<name_id>.parquet` in the reference's COO schema."""
    pa, pq = _pa()
    indptr = np.asarray(indptr, dtype=np.uint64)
    n_rows = len(indptr) - 1
    nnz = int(indptr[-1])
    n_cols = n_rows if n_cols is None else n_cols
    rows = np.repeat(np.arange(n_rows, dtype=np.uint64), np.diff(indptr.astype(np.int64)))
    table = pa.table({
        "name_id": pa.array([name_id] * nnz, pa.string()),
        "n_rows": pa.array(np.full(nnz, n_rows, np.uint64)),
        "n_cols": pa.array(np.full(nnz, n_cols, np.uint64)),
        "nnz": pa.array(np.full(nnz, nnz, np.uint64)),
        "row": pa.array(rows),
        "col": pa.array(np.asarray(indices[:nnz], dtype=np.uint64)),
        "value": pa.array(np.asarray(data[:nnz], dtype=np.float64)),
    })
    schema = pa.schema([pa.field(n, t, nullable=False) for n, t in zip(table.column_names, [c.type for c in table.columns])])
    os.makedirs(path, exist_ok=True)
    out = os.path.join(path, f"{name_id}.parquet")
    pq.write_table(table.cast(schema), out, compression="snappy")
    return out


def load_sparse_matrix(file_path):
    """The reference's COO Parquet -> (indptr u64, indices u32, data f64, (n_rows, n_cols)); triplets are summed into CSR
    in (row, col) order like sprs' TriMat::to_csr."""
    pa, pq = _pa()
    t = pq.read_table(file_path)
    n_rows, n_cols = int(t["n_rows"][0].as_py()), int(t["n_cols"][0].as_py())
    row = t["row"].to_numpy().astype(np.int64)
    col = t["col"].to_numpy().astype(np.int64)
    val = t["value"].to_numpy().astype(np.float64)
    order = np.lexsort((col, row))
    row, col, val = row[order], col[order], val[order]
    indptr = np.zeros(n_rows + 1, np.uint64)
    np.add.at(indptr, row + 1, 1)
    return np.cumsum(indptr).astype(np.uint64), col.astype(np.uint32), val, (n_rows, n_cols)


def save_lambda(lambdas, path, name_id):
    pa, pq = _pa()
    lam = np.asarray(lambdas, dtype=np.float64)
    if lam.size == 0:
        raise ValueError("Cannot save empty lambda vector")   # parquet.rs:736-740
    n = lam.size
    table = pa.table({
        "name_id": pa.array([name_id] * n, pa.string()),
        "n_values": pa.array(np.full(n, n, np.uint64)),
        "row_index": pa.array(np.arange(n, dtype=np.uint64)),
        "lambda": pa.array(lam),
    })
    schema = pa.schema([pa.field(nm, c.type, nullable=False) for nm, c in zip(table.column_names, table.columns)])
    os.makedirs(path, exist_ok=True)
    out = os.path.join(path, f"{name_id}.parquet")
    pq.write_table(table.cast(schema), out, compression="snappy")
    return out


def load_lambda(file_path):
    pa, pq = _pa()
    t = pq.read_table(file_path)
    idx = t["row_index"].to_numpy().astype(np.int64)
    lam = np.empty(int(t["n_values"][0].as_py()), np.float64)
    lam[idx] = t["lambda"].to_numpy()
    return lam
