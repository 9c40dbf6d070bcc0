"""The NCCL exchange calls of the row-sharded build on two GPUs (one process per GPU), against the single-GPU results
and the oracle: sfb_knn_allgather, sfb_lambda_allgather (taumode and the f32 core variant, whose dispersion is
normalised by the energy of ALL items), sfb_mat_allgather_rows, sfb_knn_build_columns_sharded and the sharded
begin / end form.  Skipped when fewer than two devices are visible."""
import os
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(n, rank, world):
    s = (n + world - 1) // world
    lo = min(rank * s, n)
    return lo, min(lo + s, n)


def _worker(rank, world, tmp, n, d, k):
    sys.path.insert(0, ROOT)
    from sfb_loader import load
    sfb = load()
    ctx = sfb.Context(rank)
    idf = os.path.join(tmp, "nccl_id")
    if rank == 0:
        with open(idf + ".tmp", "wb") as f:
            f.write(sfb.comm_unique_id())
        os.rename(idf + ".tmp", idf)
    t0 = time.time()
    while not os.path.exists(idf):
        if time.time() - t0 > 60:
            raise RuntimeError("rank 0 never published the NCCL id")
        time.sleep(0.05)
    ctx.comm_init(open(idf, "rb").read(), rank, world)
    lo, hi = _shard(n, rank, world)
    X = ctx.generate(sfb.SYNTH_CLUSTERED, 7, n, d, 16, 0.3)
    out = {}
    # (1) every rank uploads its shard, the corpus is assembled by one all-gather over NVLink
    full = X.rows()
    Xg = ctx.matrix(full[lo:hi]).allgather_rows(n)
    out["mat"] = Xg.rows()
    # (2) kNN of the shard through the tensor-core screen, lists all-gathered
    g = Xg.knn(k, sfb.METRIC_COSINE, screen=sfb.SCREEN_F16, q_begin=lo, q_end=hi)
    ga = g.allgather(n)
    out["idx"], out["dist"], out["cnt"] = ga.to_host()
    # (3) feature graph: pair sums split across the ranks and all-reduced (plain and begin / end forms)
    gf = Xg.knn_columns(min(k, d - 1), sfb.METRIC_COSINE, sharded=True)
    out["f_idx"], out["f_dist"], out["f_cnt"] = gf.to_host()
    pend = Xg.knn_columns_begin(min(k, d - 1), sfb.METRIC_COSINE, sharded=True)
    g2 = Xg.knn(k, sfb.METRIC_L2, screen=sfb.SCREEN_F16, q_begin=lo, q_end=hi)
    gf2 = pend.end()
    out["f2_idx"], out["f2_dist"], out["f2_cnt"] = gf2.to_host()
    out["l2_idx"], out["l2_dist"], out["l2_cnt"] = g2.allgather(n).to_host()
    # (4) lambda of the shard, global min / max, all-gathered: taumode, energy-node and the f32 core variant
    Lf = gf.adjacency(2.0, 1.0).laplacian()
    xs = Xg.view_rows(lo, hi - lo)
    for name, variant in (("lam_legacy", sfb.LAMBDA_LEGACY_TAUMODE), ("lam_energy", sfb.LAMBDA_ENERGY_NODE), ("lam_core", sfb.LAMBDA_CORE_F32SEM)):
        lam, stats = Lf.lambdas_allgather(xs, lo, n, variant=variant, normalise=True)
        out[name], out[name + "_stats"] = lam, stats
        raw, _ = Lf.lambdas_allgather(xs, lo, n, variant=variant, normalise=False)
        out[name + "_raw"] = raw
    ctx.barrier()
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), **out)
    ctx.close()


@pytest.mark.parametrize("n,d,k", [(6001, 96, 8), (4096, 64, 5)])   # ragged and even shards
def test_two_gpu_exchange_matches_single_gpu_and_oracle(sfb, oracle, tmp_path, n, d, k):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, str(tmp_path), n, d, k), nprocs=2, join=True)
    x = oracle.generate_rows(1, 7, 0, n, d, 16, 0.3)
    want = oracle.knn(x, k, oracle.METRIC_COSINE)
    want_l2 = oracle.knn(x, k, oracle.METRIC_L2)
    f_want = oracle.knn(oracle.transpose(x), min(k, d - 1), oracle.METRIC_COSINE)
    fl = oracle.laplacian(*oracle.build_adjacency(*f_want, 2.0, 1.0)[:3])
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert np.array_equal(got["mat"], x)
        for pre, w in (("", want), ("l2_", want_l2), ("f_", f_want), ("f2_", f_want)):
            assert np.array_equal(got[pre + "idx"], w[0]) and np.array_equal(got[pre + "dist"], w[1]) and np.array_equal(got[pre + "cnt"], w[2]), pre
        for name, variant, tol in (("lam_legacy", oracle.LAMBDA_LEGACY_TAUMODE, 1e-9), ("lam_energy", oracle.LAMBDA_ENERGY_NODE, 1e-9),
                                   ("lam_core", oracle.LAMBDA_CORE_F32SEM, 1e-5)):
            raw = oracle.lambdas(*fl, x, variant, oracle.TAU_MEDIAN)
            o_n, o_stats = oracle.normalise_lambdas(raw)
            assert np.allclose(got[name + "_raw"], raw, rtol=tol, atol=1e-6 if tol > 1e-6 else 1e-14), name
            assert np.allclose(got[name + "_stats"], o_stats, rtol=tol, atol=1e-6 if tol > 1e-6 else 1e-14), name
            assert np.allclose(got[name], o_n, rtol=tol * 10, atol=1e-6 if tol > 1e-6 else 1e-12), name
    # both ranks hold the same bits
    a, b = (np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(2))
    for key in a.files:
        assert np.array_equal(a[key], b[key], equal_nan=True), key
