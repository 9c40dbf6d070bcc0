#!/usr/bin/env python
"""bench.py -- graph-wiring build throughput (kNN graph -> Laplacian -> taumode lambda), vectors/s.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one full build over one synthetic embedding matrix (BASELINE.json configs[1] by default:
1M x 384 clustered, cosine k=16, + Laplacian + taumode lambda):
    item graph     kNN (tcgen05 screen + exact f64 rescore) -> kernel weights / sparsification
                   -> symmetrise + CSR Laplacian                       (src_legacy/laplacian.rs:122-419)
    feature graph  the reference's own call shape: kNN over the D feature nodes (the columns), Laplacian
                   (src_legacy/graph.rs:193-255); its exact f64 pair sums run on a side stream beside the screen
    lambda         per-item taumode lambda against the D x D feature Laplacian, min-max normalised
                   (src_legacy/taumode.rs:117-318, core.rs:1341-1355)
`value`  = rows / device time with the f64 matrix already resident in HBM.
`e2e`    = the same build through the public API with the matrix in pinned HOST memory: H2D upload,
           build, D2H of the item Laplacian (CSR) and the lambda vector inside the timed region.
N > 1: one process per GPU (torchrun); query rows are sharded, the f64 corpus is replicated, the kNN
lists and lambdas are all-gathered with NCCL (strong scaling: the matrix is the same at every N).
The reference arm times the CPU oracle (a C restatement of the reference; Rust cannot be built here)
on a bounded sample of the same workload and extrapolates per stage -- see `cpu_baseline.sample`.
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC_NAME = "kNN-graph+Laplacian+lambda build throughput"
UNIT = "vectors/s"

# BASELINE.json configs (SURVEY.md section 8d gives the synthetic laws)
WORKLOADS = {
    "c1": dict(name="C1 10k x 384 Gaussian, cosine k=16", rows=10_000, cols=384, kind=0, seed=42, centres=0, noise=0.0,
               metric=0, k=16),
    "c2": dict(name="C2 1M x 384 clustered, cosine k=16 + Laplacian + taumode lambda", rows=1_000_000, cols=384, kind=1,
               seed=7, centres=1024, noise=0.3, metric=0, k=16),
    "c3": dict(name="C3 10M x 768 clustered, L2 k=16 + lambda, query-row sharded", rows=10_000_000, cols=768, kind=1, seed=11,
               centres=4096, noise=0.3, metric=1, k=16),
    "c4": dict(name="C4 100k x 3072 anisotropic, cosine k=32", rows=100_000, cols=3072, kind=2, seed=13, centres=0,
               noise=0.1, metric=0, k=32),
    "c5": dict(name="C5 5M x 128 clustered, L2 k=64", rows=5_000_000, cols=128, kind=1, seed=17, centres=2048, noise=0.5,
               metric=1, k=64),
}
P_WEIGHT, SIGMA = 2.0, 1.0


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons sampled every 200 ms DURING the timed region, through NVML in a
    thread of this process.  (An `nvidia-smi -lms 200` child process was measured to slow the timed step
    by ~25 %: every query stalls the CUDA calls of the process being measured.)"""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, gpu_index, period=0.2):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[gpu_index]) if gpu_index < len(ids) else gpu_index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return

        def loop():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    for name, bit in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                self._stop.wait(period)

        self._t = threading.Thread(target=loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm),
                "sm_mhz_min": sm[0], "how": "NVML in-process thread, 200 ms period, timed region only"}


def shard(n, rank, world):
    s = (n + world - 1) // world
    lo = min(rank * s, n)
    return lo, min(lo + s, n)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
_TRACE = bool(os.environ.get("SFB_BENCH_TRACE"))


def _timed(ctx, name, f):
    """SFB_BENCH_TRACE=1: wall time of every public call, to stderr (development aid; adds synchronisation)."""
    if not _TRACE:
        return f()
    ctx.synchronize()
    t0 = time.perf_counter()
    r = f()
    ctx.synchronize()
    sys.stderr.write(f"  {name:14s} {(time.perf_counter() - t0) * 1e3:9.2f} ms\n")
    return r


def build_once(sfb, ctx, X, wl, rank, world, out_lambda=None, keep=False):
    """One full build on device-resident X.  Returns (handles or None, stats)."""
    n, d = X.shape
    lo, hi = shard(n, rank, world)
    # the feature graph (the reference's call shape: nodes = columns) is registered first and runs on a side stream
    # beside the item graph's tensor-core screen; .end() joins it
    pend = X.knn_columns_begin(min(wl["k"], d - 1), sfb.METRIC_COSINE, sharded=world > 1)
    g = _timed(ctx, "knn", lambda: X.knn(wl["k"], wl["metric"], q_begin=lo, q_end=hi))
    st = g.stats()
    if world > 1:
        g_all = _timed(ctx, "allgather", lambda: g.allgather(n))
        g.free()
        g = g_all
    adj = _timed(ctx, "adjacency", lambda: g.adjacency(P_WEIGHT, SIGMA))
    L = _timed(ctx, "laplacian", lambda: adj.laplacian())
    # feature Laplacian + per-item lambda
    gf = _timed(ctx, "knn_columns", lambda: pend.end())
    adjf = _timed(ctx, "adjacency_f", lambda: gf.adjacency(P_WEIGHT, SIGMA))
    Lf = _timed(ctx, "laplacian_f", lambda: adjf.laplacian())
    xs = X.view_rows(lo, hi - lo)
    lam, lstats = _timed(ctx, "lambda", lambda: Lf.lambdas_allgather(xs, lo, n, normalise=True))
    if out_lambda is not None:
        out_lambda[:] = lam
    if keep == "all":
        for h in (xs, adjf, gf, adj):
            h.free()
        return (L, Lf, lam, g), st
    _timed(ctx, "free", lambda: [h.free() for h in (xs, adjf, gf, adj, g)])
    if keep:
        return (L, Lf, lam), st
    L.free(); Lf.free()
    if _TRACE:
        sys.stderr.write(f"  knn stats: prepare {st['ms_prepare']:.2f} screen {st['ms_screen']:.2f} rescore {st['ms_rescore']:.2f} fallback {st['ms_fallback']:.2f}\n")
    return None, st


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from sfb_loader import load
    sfb = load()  # raises if libsurfface_b200.so is missing: there is no fallback

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one process per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sfb.Context(local)
    if world > 1:
        ids = [sfb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)

    wl = dict(WORKLOADS[args.config])
    if args.rows:
        wl["rows"] = args.rows
        wl["name"] += f" (rows overridden to {args.rows})"
    n, d, k = wl["rows"], wl["cols"], wl["k"]
    lo, hi = shard(n, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    X = ctx.generate(wl["kind"], wl["seed"], n, d, wl["centres"], wl["noise"])
    ctx.synchronize()

    # ---- device-resident timing --------------------------------------------------------------
    for _ in range(args.warmup):
        build_once(sfb, ctx, X, wl, rank, world)
    barrier()
    ctx.timings_reset()
    sampler = ClockSampler(local) if rank == 0 and not args.no_clocks else None
    stats = []
    t_wall = time.perf_counter()
    ctx.timer_start()
    step_wall = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        _, st = build_once(sfb, ctx, X, wl, rank, world)
        stats.append(st)
        step_wall.append((time.perf_counter() - t_s) * 1e3)  # every step ends with a synchronising D2H of lambda
    ms = ctx.timer_stop()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if sampler else None
    tm = ctx.timings()
    t = torch.tensor([ms, t_wall * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, wall_ms = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = n / (ms_step * 1e-3)

    # ---- parity at full size (--verify): sampled rows against the CPU oracle, structural invariants ---
    verify = None
    if args.verify:
        verify = verify_build(sfb, ctx, X, wl, rank, world)

    # ---- end to end: pinned host matrix in, CSR + lambda out ----------------------------------
    e2e = None
    if not args.no_e2e:
        xh = ctx.pinned_empty((n, d))
        step_rows = 1 << 16
        for r0 in range(0, n, step_rows):
            xh[r0:r0 + step_rows] = X.rows(r0, min(step_rows, n - r0))
        X.free()
        nnz_cap = n * (2 * k + 1)
        out_csr = (ctx.pinned_empty(n + 1, np.uint64), ctx.pinned_empty(nnz_cap, np.uint32), ctx.pinned_empty(nnz_cap, np.float64))
        out_lam = ctx.pinned_empty(n, np.float64)
        d2h = 0

        def e2e_step():
            nonlocal d2h
            if world > 1:   # upload this rank's rows only; the corpus is replicated over NVLink, not PCIe
                xs_ = ctx.matrix(xh[lo:hi])
                Xd = xs_.allgather_rows(n)
                xs_.free()
            else:
                Xd = ctx.matrix(xh)
            (L, Lf, lam), _ = build_once(sfb, ctx, Xd, wl, rank, world, out_lambda=out_lam, keep=True)
            indptr, indices, data = L.to_host(out_csr)
            d2h = indptr.nbytes + indices.nbytes + data.nbytes + out_lam.nbytes
            L.free(); Lf.free(); Xd.free()

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()  # warm-up
        barrier()
        t0 = time.perf_counter()
        ctx.timer_start()
        for _ in range(e2e_steps):
            e2e_step()
        ms_e = ctx.timer_stop()
        barrier()
        wall_e = (time.perf_counter() - t0) * 1e3
        te = torch.tensor([max(ms_e, wall_e)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ms_e_step = float(te[0]) / e2e_steps
        e2e = {"value": n / (ms_e_step * 1e-3), "unit": UNIT, "ms_per_step": ms_e_step, "steps": e2e_steps,
               "h2d_bytes_per_step": int(hi - lo) * d * 8 * world, "d2h_bytes_per_step": int(d2h) * world,
               "api": "Context.matrix(host f64 rows of this rank) [-> Matrix.allgather_rows] -> Matrix.knn -> adjacency -> laplacian -> Csr.to_host; "
                      "Csr.lambdas_allgather -> host (every rank fetches the full CSR and lambda)"}
    else:
        X.free()

    # ---- roofline of the dominant kernel (tcgen05 distance screen) ----------------------------
    pk = peaks()
    scr_ms = sum(s["ms_screen"] for s in stats) / len(stats)
    screened = stats[0]["screen_used"] in (sfb.SCREEN_F16, sfb.SCREEN_BF16)
    flops = 2.0 * (hi - lo) * n * d  # algorithmic: every ordered (query, corpus) pair, D multiply-adds
    roof = None
    if screened and scr_ms > 0:
        ach = flops / (scr_ms * 1e-3) / 1e12
        peak = pk["tc_sustained"] if scr_ms > 100 else pk["tc_burst"]
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r01e_screen_traffic.json")
        if args.config == "c2" and not args.rows and world == 1 and os.path.exists(tpath):
            tj = json.load(open(tpath))   # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
            traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
        roof = {"kernel": "knn_screen_pair_kernel", "bound": "tensor", "achieved": ach,
                "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "ms_per_launch": scr_ms, "flops_per_launch": flops,
                "peak_source": pk["source"] + (", sustained" if scr_ms > 100 else ", burst")}

    if rank == 0:
        line = {
            "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "dtype_note": "results are the exact f64 values of the reference arithmetic; the tensor-core screen that proposes candidates runs in fp16 with fp32 accumulation",
            "data": "synthetic (counter-based Philox + Box-Muller, generated on device)",
            "config": {"workload": wl["name"], "rows": n, "cols": d, "k": k, "metric": ["cosine", "l2", "l2sq"][wl["metric"]],
                       "p": P_WEIGHT, "sigma": SIGMA, "sharding": f"query rows / {world}, corpus replicated" if world > 1 else "single GPU",
                       "l2_policy": f"inputs larger than L2 ({n * d * 8 / 1e9:.2f} GB f64 matrix streamed every step)"},
            "impl": "surfface_b200", "wall_ms_per_step": wall_ms / args.steps, "wall_ms_steps": [round(v, 2) for v in step_wall],
            "knn_ms_steps": [[round(s_[k_], 2) for k_ in ("ms_prepare", "ms_screen", "ms_rescore")] for s_ in stats],
            "stages_ms_per_step": {kk: tm[kk] / args.steps for kk in ("ms_knn", "ms_adjacency", "ms_laplacian", "ms_lambda")},
            "knn": {kk: stats[-1][kk] for kk in ("rows", "rows_certified", "rows_fallback", "k_prime", "screen_used", "ms_prepare",
                                                 "ms_screen", "ms_rescore", "ms_fallback", "max_margin", "rows_rescreened", "ms_rescreen")},
            "gpu_launches": int(tm["kernel_launches"]),
            "clocks": clocks, "roofline": roof, "e2e": e2e,
        }
        if verify is not None:
            line["verify"] = verify
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_sample(wl, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def verify_build(sfb, ctx, X, wl, rank, world, n_sample=48):
    """Full-size parity (every rank runs the build; rank 0 checks): kNN lists of sampled rows bit-exact against
    the oracle's brute force over the FULL corpus; CSR invariants of the item Laplacian (sorted columns, stored
    diagonal, row sums 0, symmetry on sampled entries); feature Laplacian and sampled lambdas against the oracle."""
    import numpy as np
    import oracle
    oracle.build()
    oracle.use_all_threads()
    n, d, k = wl["rows"], wl["cols"], wl["k"]
    (L, Lf, lam, g), _ = build_once(sfb, ctx, X, wl, rank, world, keep="all")
    out = None
    if rank == 0:
        xh = np.empty((n, d))
        step = 1 << 16
        for r0 in range(0, n, step):
            xh[r0:r0 + step] = X.rows(r0, min(step, n - r0))
        rows = np.unique(np.random.default_rng(1).integers(0, n, n_sample))
        idx, dist_, cnt = g.to_host()
        o_idx, o_dist, o_cnt = oracle.knn(xh, k, wl["metric"], query_rows=rows)
        knn_ok = bool(np.array_equal(idx[rows], o_idx) and np.array_equal(dist_[rows], o_dist) and np.array_equal(cnt[rows], o_cnt))
        indptr, indices, data = L.to_host()
        ip = indptr.astype(np.int64)
        deg = np.diff(ip)
        row_of = np.repeat(np.arange(n), deg)
        sorted_ok = bool(np.all((np.diff(indices.astype(np.int64)) > 0) | (np.diff(row_of) > 0)))
        diag_ok = bool(np.count_nonzero(indices == row_of) == n)
        rs = np.add.reduceat(data, ip[:-1])
        scale = np.add.reduceat(np.abs(data), ip[:-1]) + 1e-300
        rowsum_ok = bool(np.max(np.abs(rs) / scale) < 1e-12)
        sym_ok = True
        for r in rows[:16]:
            for e in range(ip[r], ip[r + 1]):
                c = int(indices[e])
                pos = ip[c] + np.searchsorted(indices[ip[c]:ip[c + 1]], r)
                sym_ok = sym_ok and pos < ip[c + 1] and indices[pos] == r and data[pos] == data[e]
        # the item graph from the gathered lists == the oracle's assembly of the same lists
        a = oracle.build_adjacency(idx, dist_, cnt, P_WEIGHT, SIGMA)
        o_ptr, o_ind, o_dat = oracle.laplacian(*a[:3])
        lap_ok = bool(np.array_equal(indptr, o_ptr) and np.array_equal(indices, o_ind) and np.allclose(data, o_dat, rtol=1e-9, atol=0))
        # feature graph + lambda
        f = oracle.knn(oracle.transpose(xh), min(k, d - 1), oracle.METRIC_COSINE)
        fl = oracle.laplacian(*oracle.build_adjacency(*f, P_WEIGHT, SIGMA)[:3])
        fptr, find, fdat = Lf.to_host()
        flap_ok = bool(np.array_equal(fptr, fl[0]) and np.array_equal(find, fl[1]) and np.allclose(fdat, fl[2], rtol=1e-9, atol=0))
        o_lam, _ = oracle.normalise_lambdas(oracle.lambdas(*fl, xh))
        lam_ok = bool(np.allclose(lam, o_lam, rtol=1e-9, atol=1e-12))
        out = {"knn_rows_checked": int(len(rows)), "knn_bit_exact": knn_ok, "csr_sorted": sorted_ok, "csr_diagonal_stored": diag_ok,
               "csr_row_sums_zero": rowsum_ok, "csr_symmetric_sample": bool(sym_ok), "item_laplacian_matches_oracle": lap_ok,
               "feature_laplacian_matches_oracle": flap_ok, "lambda_all_rows_within_1e-9": lam_ok}
        out["ok"] = all(v for kk, v in out.items() if kk != "knn_rows_checked")
    for h in (L, Lf, g):
        h.free()
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference) on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def cpu_setup(wl):
    """Corpus on the host + the inputs of the non-kNN stages (untimed)."""
    import numpy as np
    import oracle
    oracle.build()
    oracle.use_all_threads()
    key = (wl["rows"], wl["cols"], wl["seed"])
    if key in _CPU_CACHE:
        return _CPU_CACHE[key]
    n, d = wl["rows"], wl["cols"]
    x = oracle.generate_rows(wl["kind"], wl["seed"], 0, n, d, wl["centres"], wl["noise"])
    nb = min(n, 8192)  # block on which the linear-cost stages are timed
    xb = np.ascontiguousarray(x[:nb])
    b_idx, b_dist, b_cnt = oracle.knn(xb, min(wl["k"], nb - 1), wl["metric"])
    _CPU_CACHE[key] = (x, xb, (b_idx, b_dist, b_cnt))
    return _CPU_CACHE[key]


def cpu_step(wl, nq):
    """One sampled pass.  Returns (vectors/s extrapolated, detail)."""
    import numpy as np
    import oracle
    x, xb, lists = cpu_setup(wl)
    n, d, k = wl["rows"], wl["cols"], wl["k"]
    nb = xb.shape[0]
    q = np.linspace(0, n - 1, nq).astype(np.uint64)
    t0 = time.perf_counter()
    oracle.knn(x, k, wl["metric"], query_rows=q)           # nq query rows against the FULL corpus
    t_knn = time.perf_counter() - t0
    t0 = time.perf_counter()
    a = oracle.build_adjacency(*lists, P_WEIGHT, SIGMA)
    L = oracle.laplacian(*a[:3])
    xt = oracle.transpose(xb)
    f = oracle.knn(xt, min(k, d - 1), oracle.METRIC_COSINE)
    fa = oracle.build_adjacency(*f, P_WEIGHT, SIGMA)
    fl = oracle.laplacian(*fa[:3])
    lam = oracle.lambdas(*fl, xb, oracle.LAMBDA_LEGACY_TAUMODE, oracle.TAU_MEDIAN)
    oracle.normalise_lambdas(lam)
    t_rest = time.perf_counter() - t0
    per_row = t_knn / nq + t_rest / nb
    return 1.0 / per_row, {"t_knn_s": t_knn, "nq": nq, "t_rest_s": t_rest, "rest_rows": nb}


def cpu_sample(wl, seconds):
    import oracle
    cpu_setup(wl)
    n, d = wl["rows"], wl["cols"]
    _, probe = cpu_step(wl, 16)
    nq = int(max(16, min(4096, seconds / max(probe["t_knn_s"] / 16, 1e-9))))
    v, det = cpu_step(wl, nq)
    return {"value": v, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": (f"oracle (C restatement of the reference, OpenMP): kNN of {det['nq']} query rows against the full {n} x {d} "
                       f"corpus in {det['t_knn_s']:.2f} s, scaled by rows; weights + Laplacian + feature graph + lambda on a "
                       f"{det['rest_rows']}-row block in {det['t_rest_s']:.2f} s, scaled by rows")}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import oracle
    wl = dict(WORKLOADS[args.config])
    if args.rows:
        wl["rows"] = args.rows
    cpu_setup(wl)
    n, d = wl["rows"], wl["cols"]
    _, probe = cpu_step(wl, 16)
    nq = int(max(16, min(4096, args.cpu_seconds / max(probe["t_knn_s"] / 16, 1e-9))))
    for _ in range(args.warmup):
        cpu_step(wl, max(16, nq // 8))
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        v, det = cpu_step(wl, nq)
        vals.append(1.0 / v)
    wall = time.perf_counter() - t0
    value = 1.0 / (sum(vals) / len(vals))
    sample = (f"per step: kNN of {nq} query rows against the full {n} x {d} corpus, scaled by rows; weights + Laplacian + "
              f"feature graph + lambda on a {det['rest_rows']}-row block, scaled by rows")
    line = {"impl": "reference", "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (same counter-based generator, on the host)",
            "config": {"workload": wl["name"], "rows": n, "cols": d, "k": wl["k"], "metric": ["cosine", "l2", "l2sq"][wl["metric"]],
                       "p": P_WEIGHT, "sigma": SIGMA},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference is a Rust workspace that cannot be compiled in this image (no cargo/rustc); the CPU arm is the "
                    "oracle port of its algorithm, all host threads"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the row count (development only)")
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="CPU time budget of one sampled kNN pass")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--verify", action="store_true", help="after timing, check the full-size build against the CPU oracle (adds ~1 min)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi during the timed region (diagnosis)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
