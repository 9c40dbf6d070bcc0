"""Feature-graph Gram chains at C4's shape (3072 nodes x 100k dimensions), stand-alone: 16- against 32-edge pair tiles."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GT = os.environ.get("GRAM_C4_GT")
if GT is None:
    for gt in ("16", "32"):
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, GRAM_C4_GT=gt, SFB_GRAM_GT=gt), capture_output=True, text=True, timeout=200)
        print(f"gt={gt}: {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
    sys.exit(0)
from sfb_loader import load
sfb = load()
ctx = sfb.Context(0)
X = ctx.generate(2, 11, 100000, 3072, 0, 0.1)
best = 1e30
for _ in range(3):
    ctx.synchronize(); t = time.perf_counter(); g = X.knn_columns(32, 0); ctx.synchronize()
    best = min(best, (time.perf_counter() - t) * 1e3); g.free()
ops = 3072 * 3073 / 2 * 100000 * 2
print(f"{best:7.2f} ms  {ops / best / 1e9:6.2f} T FP64 instr/s = {ops / (best * 1e-3) / 148 / 1.9e9:5.1f} lanes/clk/SM at 1.9 GHz")
