"""Step-time jitter hunt (development aid): same C2 build repeated, with optional torch import / NVML sampler."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if "torch" in mode:
    import torch
    torch.cuda.set_device(0)
from sfb_loader import load
sfb = load()
ctx = sfb.Context(0)
X = ctx.generate(1, 7, 1000000, 384, 1024, 0.3)
ctx.synchronize()
stop = threading.Event()
if "nvml" in mode:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    def loop():
        while not stop.is_set():
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            stop.wait(0.2)
    threading.Thread(target=loop, daemon=True).start()
times = []
for it in range(14):
    t0 = time.perf_counter()
    g = X.knn(16, 0); adj = g.adjacency(2.0, 1.0); L = adj.laplacian()
    gf = X.knn_columns(16, 0); adjf = gf.adjacency(2.0, 1.0); Lf = adjf.laplacian()
    lam = Lf.lambdas_allgather(X, 0, 1000000, normalise=True)
    for hh in (g, adj, L, gf, adjf, Lf):
        hh.free()
    times.append(round((time.perf_counter() - t0) * 1e3, 1))
stop.set()
print(mode, times)
