"""Wire format of the path's outputs: the reference's Parquet layout (src_legacy/storage/parquet.rs), so that a
Laplacian and a lambda vector built here can be read by the reference's loaders and vice versa.

    save_sparse_matrix / load_sparse_matrix   parquet.rs:412-520,520-590: COO rows
        name_id Utf8 | n_rows u64 | n_cols u64 | nnz u64 | row u64 | col u64 | value f64      (Snappy, one row per non-zero,
        in CSR order: row-major, ascending column)
    save_lambda / load_lambda                 parquet.rs:728-826:
        name_id Utf8 | n_values u64 | row_index u64 | lambda f64
    save_metadata / load_metadata             parquet.rs:131-170: `<name_id>_metadata.json`, the ArrowSpaceMetadata sidecar the
        reference writes next to a file when a builder configuration is passed (parquet.rs:486-503,788-806): name_id,
        timestamp (RFC 3339), n_rows, n_cols, builder_config (serde's externally tagged ConfigValue: {"F64": 0.001},
        {"Usize": 6}, {"OptionF64": null}, {"TauMode": "Median"}), files {key: FileInfo}
Files are `<path>/<name_id>.parquet`.  Host-side I/O only (the reference's persistence is not on the compute path; a Rust
host keeps using the reference's own storage module on the CsMat / Vec the wrapper returns)."""
import datetime
import json
import os

import numpy as np


def config_value(v):
    """A Python value as serde serialises the reference's ConfigValue (surfface-pipeline/src/builder.rs:1531-1544)."""
    if isinstance(v, dict):          # already tagged
        return v
    if isinstance(v, bool):
        return {"Bool": v}
    if isinstance(v, int):
        return {"Usize": v}
    if isinstance(v, float):
        return {"F64": v}
    if isinstance(v, str):
        return {"String": v}
    if v is None:
        return {"OptionF64": None}
    raise TypeError(f"no ConfigValue form for {type(v).__name__}")


def save_metadata(path, name_id, n_rows, n_cols, builder_config, files):
    """ArrowSpaceMetadata -> `<path>/<name_id>_metadata.json` (parquet.rs:131-145), pretty-printed like serde_json."""
    meta = {"name_id": name_id, "timestamp": datetime.datetime.now(datetime.timezone.utc).isoformat(),
            "n_rows": int(n_rows), "n_cols": int(n_cols),
            "builder_config": {k: config_value(v) for k, v in builder_config.items()}, "files": files}
    os.makedirs(path, exist_ok=True)
    out = os.path.join(path, f"{name_id}_metadata.json")
    with open(out, "w") as f:
        json.dump(meta, f, indent=2)
    return out


def load_metadata(path, name_id):
    with open(os.path.join(path, f"{name_id}_metadata.json")) as f:
        return json.load(f)


def _pa():
    import pyarrow as pa
    import pyarrow.parquet as pq
    return pa, pq


def save_sparse_matrix(indptr, indices, data, path, name_id, n_cols=None, builder_config=None):
    """CSR (as returned by Csr.to_host()) -> `<path>/<name_id>.parquet` in the reference's COO schema; with a
    builder_config also the `<name_id>_metadata.json` sidecar (parquet.rs:486-503)."""
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
    if builder_config is not None:
        save_metadata(path, name_id, n_rows, n_cols, builder_config,
                      {"matrix": {"filename": f"{name_id}.parquet", "file_type": "sparse", "rows": n_rows, "cols": int(n_cols),
                                  "nnz": nnz, "size_bytes": os.path.getsize(out)}})
    return out


def load_sparse_matrix(file_path):
    """The reference's COO Parquet -> (indptr u64, indices u32, data f64, (n_rows, n_cols)).  Triplets are ordered by
    (row, col) and duplicates of a position SUMMED, like sprs' TriMat::to_csr (parquet.rs:520-590)."""
    pa, pq = _pa()
    t = pq.read_table(file_path)
    if t.num_rows == 0:
        raise ValueError("empty sparse-matrix file: the dimensions live in its rows")
    n_rows, n_cols, nnz = int(t["n_rows"][0].as_py()), int(t["n_cols"][0].as_py()), int(t["nnz"][0].as_py())
    if nnz != t.num_rows:
        raise ValueError(f"file declares nnz = {nnz} but holds {t.num_rows} triplets")
    row = t["row"].to_numpy().astype(np.int64)
    col = t["col"].to_numpy().astype(np.int64)
    val = t["value"].to_numpy().astype(np.float64)
    if row.max() >= n_rows or col.max() >= n_cols:
        raise ValueError("triplet outside the declared shape")
    order = np.lexsort((col, row))
    row, col, val = row[order], col[order], val[order]
    first = np.ones(len(row), bool)
    first[1:] = (row[1:] != row[:-1]) | (col[1:] != col[:-1])
    starts = np.nonzero(first)[0]
    val = np.add.reduceat(val, starts)
    row, col = row[starts], col[starts]
    indptr = np.zeros(n_rows + 1, np.uint64)
    np.add.at(indptr, row + 1, 1)
    return np.cumsum(indptr).astype(np.uint64), col.astype(np.uint32), val, (n_rows, n_cols)


def save_lambda(lambdas, path, name_id, builder_config=None):
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
    if builder_config is not None:   # parquet.rs:788-806
        save_metadata(path, name_id, n, 1, builder_config,
                      {"lambda_vector": {"filename": f"{name_id}.parquet", "file_type": "lambda_vector", "rows": n, "cols": 1,
                                         "nnz": None, "size_bytes": os.path.getsize(out)}})
    return out


def load_lambda(file_path):
    pa, pq = _pa()
    t = pq.read_table(file_path)
    idx = t["row_index"].to_numpy().astype(np.int64)
    lam = np.empty(int(t["n_values"][0].as_py()), np.float64)
    lam[idx] = t["lambda"].to_numpy()
    return lam
