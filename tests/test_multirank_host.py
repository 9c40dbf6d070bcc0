"""N > 1 host logic on CPU (gloo, world_size 2): the row sharding convention of bench.py / comm.cu
(rank r holds rows [r*S, min((r+1)*S, n)), S = ceil(n / world); equal-count all-gather of padded shards) and the
global min/max normalisation of lambda, checked against the single-process oracle.  The GPU exchange itself
(NCCL, csrc/comm.cu) runs in bench.py --gpus N and in tests marked gpu."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = oracle.generate_rows(1, 7, 0, n, d, 8, 0.3)           # replicated corpus
    lo, hi = bench.shard(n, rank, world)
    S = (n + world - 1) // world
    idx, dist_, cnt = oracle.knn(x, k, 0, query_rows=np.arange(lo, hi))
    # padded equal-count shards, gathered in rank order (the layout of sfb_knn_allgather)
    pad = lambda a, fill: np.concatenate([a, np.full((S - a.shape[0],) + a.shape[1:], fill, a.dtype)])
    parts = []
    for arr, fill in ((idx.astype(np.int64), 0xFFFFFFFF), (dist_, 0.0), (cnt.astype(np.int64), 0)):
        t = torch.from_numpy(pad(arr, fill))
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t)
        parts.append(torch.cat(bufs)[:n].numpy())
    g_idx, g_dist, g_cnt = parts[0].astype(np.uint32), parts[1], parts[2].astype(np.uint32)
    # every rank assembles the same Laplacian from the gathered lists
    a = oracle.build_adjacency(g_idx, g_dist, g_cnt, 2.0, 1.0)
    L = oracle.laplacian(*a[:3])
    # feature graph is replicated; lambda rows are sharded, min/max all-reduced, then gathered
    f = oracle.knn(oracle.transpose(x), 3, 0)
    fa = oracle.build_adjacency(*f, 2.0, 1.0)
    fl = oracle.laplacian(*fa[:3])
    lam = oracle.lambdas(*fl, x[lo:hi])
    mn = torch.tensor([lam.min() if len(lam) else np.inf]); mx = torch.tensor([max(0.0, lam.max()) if len(lam) else 0.0])
    dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    rng = max(float(mx) - float(mn), 1e-9)
    lam_n = (lam - float(mn)) / rng
    t = torch.from_numpy(pad(lam_n, 0.0))
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t)
    lam_all = torch.cat(bufs)[:n].numpy()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=g_idx, dist=g_dist, cnt=g_cnt, indptr=L[0], indices=L[1], data=L[2],
             lam=lam_all)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [101, 64])   # ragged and even splits
def test_sharded_build_matches_single_process(oracle, tmp_path, n):
    import torch.multiprocessing as mp
    d, k, world = 12, 5, 2
    mp.spawn(_worker, args=(world, _free_port(), n, d, k, str(tmp_path)), nprocs=world, join=True)
    x = oracle.generate_rows(1, 7, 0, n, d, 8, 0.3)
    idx, dist_, cnt = oracle.knn(x, k, 0)
    a = oracle.build_adjacency(idx, dist_, cnt, 2.0, 1.0)
    L = oracle.laplacian(*a[:3])
    f = oracle.knn(oracle.transpose(x), 3, 0)
    fl = oracle.laplacian(*oracle.build_adjacency(*f, 2.0, 1.0)[:3])
    lam, _ = oracle.normalise_lambdas(oracle.lambdas(*fl, x))
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert np.array_equal(got["idx"], idx) and np.array_equal(got["dist"], dist_) and np.array_equal(got["cnt"], cnt)
        assert np.array_equal(got["indptr"], L[0]) and np.array_equal(got["indices"], L[1]) and np.array_equal(got["data"], L[2])
        np.testing.assert_allclose(got["lam"], lam, rtol=1e-12, atol=0)


def test_shard_convention():
    sys.path.insert(0, ROOT)
    import bench
    for n in (1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            spans = [bench.shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            S = (n + world - 1) // world
            assert all(hi - lo <= S for lo, hi in spans)
