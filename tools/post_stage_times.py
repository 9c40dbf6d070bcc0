"""Device time of the steps either side of the lambda kernel at C2 scale:
JL projection 1M x 384 -> r, lambda on the projected items, SortedLambdas of 1M / 10M lambdas.  Development aid."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sfb_loader import load
sfb = load()
ctx = sfb.Context(0)
QUICK = os.environ.get("POST_QUICK") == "1"   # one pass per call, no 10M sort: the run ncu profiles
n, f = 1_000_000, 384
X = ctx.generate(1, 7, n, f, 1024, 0.3)
out = {}
def timed(name, fn, reps=3):
    best = None
    for _ in range(1 if QUICK else reps + 1):
        ctx.synchronize(); t = time.perf_counter(); r = fn(); ctx.synchronize(); dt = (time.perf_counter() - t) * 1e3
        best = dt if best is None else min(best, dt)
    out[name] = round(best, 3); print(f"{name:34s} {best:9.3f} ms", flush=True); return r
for r in (sfb.compute_jl_dimension(1024, f, 0.3), 64, 128):
    s = np.random.default_rng(r).standard_normal((f, r))
    y = timed(f"project 1M x 384 -> {r}", lambda: X.project(s))
    # algorithmic work: 3 FP64 instructions per term (two DMUL + one DADD, no FMA by specification)
    out[f"project_{r}_fp64_ginstr_per_s"] = round(3.0 * n * f * r / (out[f"project 1M x 384 -> {r}"] * 1e-3) / 1e9, 1)
    if r != 128:
        y.free()
L = y.knn_columns(16, 0).adjacency(2.0, 1.0).laplacian()
lam = timed("lambda on projected 1M x 128", lambda: L.lambdas(y, normalise=True))[0]
timed("lambda, tau from unprojected rows", lambda: L.lambdas_projected(X, y, normalise=True))
sl = sfb.SortedLambdas()
timed("sorted lambdas 1M (h2d+sort+d2h)", lambda: sl.build_from(lam, ctx=ctx))
assert np.all(np.diff(sl.lambdas) >= 0) and sorted(sl.indices.tolist()) == list(range(n))
if not QUICK:
    big = np.random.default_rng(1).random(10_000_000)
    timed("sorted lambdas 10M", lambda: sl.build_from(big, ctx=ctx), reps=1)
    assert np.all(np.diff(sl.lambdas) >= 0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/post_stage_times%s.json" % ("_quick" if QUICK else ""), "w"), indent=1)
