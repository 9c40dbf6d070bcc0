"""Wall-clock time of every public call of one C2 build (development aid; not part of the bench contract)."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfb_loader import load
sfb = load()
ctx = sfb.Context(0)
X = ctx.generate(1, 7, 1000000, 384, 1024, 0.3)
ctx.synchronize()
def T(name, f):
    ctx.synchronize(); t=time.perf_counter(); r=f(); ctx.synchronize(); print(f"{name:20s} {(time.perf_counter()-t)*1e3:9.2f} ms"); return r
for it in range(3):
    print("iter", it)
    g = T("knn", lambda: X.knn(16, 0))
    adj = T("adjacency", lambda: g.adjacency(2.0, 1.0))
    L = T("laplacian", lambda: adj.laplacian())
    gf = T("knn_columns", lambda: X.knn_columns(16, 0))
    adjf = T("adjf", lambda: gf.adjacency(2.0, 1.0))
    Lf = T("lapf", lambda: adjf.laplacian())
    lam = T("lambda", lambda: Lf.lambdas_allgather(X, 0, 1000000, normalise=True))
    T("free", lambda: [h.free() for h in (g, adj, L, gf, adjf, Lf)])
    print(g.stats() if False else "")
