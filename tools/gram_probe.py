"""Feature-graph Gram chains: time per fold step of the two kernels (register ring / shared-memory ring), stand-alone and
beside a screen, at the tile counts one rank sees at 1 and at 8 GPUs.  Prints one line per case; run on one B200."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASE = os.environ.get("GRAM_PROBE_CASE")
if CASE is None:
    # every case in its own process: the kernels read their knobs from the environment
    for smem in ("", "1"):
        for gt in ("8", "16"):
            for mode in ("alone", "co", "after"):
                env = dict(os.environ, GRAM_PROBE_CASE=mode, SFB_GRAM_GT=gt)
                if smem:
                    env["SFB_GRAM_SMEM"] = "1"
                if mode != "alone":
                    env["SFB_GRAM_MODE"] = mode
                r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=600)
                print(f"kernel={'smem' if smem else 'regs'} gt={gt:>2} {mode:5}: {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
    sys.exit(0)

from sfb_loader import load
sfb = load()
ctx = sfb.Context(0)
N = 1000000


def timed(f, reps=3):
    best = 1e30
    for _ in range(reps):
        ctx.synchronize(); t = time.perf_counter(); r = f(); ctx.synchronize()
        best = min(best, (time.perf_counter() - t) * 1e3)
        r.free() if hasattr(r, "free") else [o.free() for o in r]
    return best


out = []
X = ctx.generate(1, 7, N, 384, 1024, 0.3)        # all 300 / 1176 tiles: one GPU's load
Y = ctx.generate(1, 9, N, 136, 1024, 0.3)        # 45 / 153 tiles: one rank's share at 8 GPUs
if CASE == "alone":
    for name, M in (("384", X), ("136", Y)):
        t = timed(lambda: M.knn_columns(16, 0))
        out.append(f"{name}: {t:7.2f} ms ({t * 1e6 / N:5.1f} ns/step)")
else:
    # the screen of one rank's rows at 8 GPUs (125k queries against the 1M corpus) with the chains fired beside / behind it
    t_scr = timed(lambda: X.knn(16, 0, q_begin=0, q_end=N // 8))
    def both(M):
        pend = M.knn_columns_begin(16, 0)
        g = X.knn(16, 0, q_begin=0, q_end=N // 8)
        return [g, pend.end()]
    for name, M in (("384", X), ("136", Y)):
        t = timed(lambda: both(M))
        out.append(f"{name}: {t:7.2f} ms vs knn alone {t_scr:6.2f} (+{t - t_scr:6.2f})")
print("  ".join(out))
