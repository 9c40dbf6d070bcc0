"""One small pass through every kernel family of the library (a quick all-kernels smoke run; compute-sanitizer is
closed on this pool, so bad accesses are hunted with the parity tests instead)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sfb_loader import load
sfb = load()
ctx = sfb.Context(0)
rng = np.random.default_rng(0)
x = rng.normal(size=(4500, 70))
m = ctx.matrix(x)
for metric in (0, 1):
    pend = m.knn_columns_begin(5, 0)
    g = m.knn(9, metric)                       # tensor-core screen (pair kernel), rescore, certify
    gf = pend.end()
    adj = g.adjacency(2.0, 1.0)
    adj.sfgrass(0.5)
    L = adj.laplacian()
    Ln = adj.laplacian(normalised=True)
    Lf = gf.adjacency(2.0, 1.0).laplacian()
    lam, stats = Lf.lambdas(m, normalise=True)
    Lf.lambdas(m, sfb.LAMBDA_ENERGY_NODE)
    Lf.lambdas(m, sfb.LAMBDA_CORE_F32SEM)
    print(metric, g.stats()["screen_used"], L.shape, Ln.shape, float(lam.mean()))
g = m.knn(7, 0, screen=sfb.SCREEN_EXACT_F64, q_begin=10, q_end=700)
g = m.knn(7, 2, screen=sfb.SCREEN_BF16, k_prime=8)   # small k': re-screen + fallback levels
print(g.stats())
y = ctx.matrix(rng.normal(size=(4200, 600)))            # K > 512: partially resident query slabs
print(y.knn(4, 0).stats()["rows_certified"])
means = rng.normal(size=(20, 60)).astype(np.float32); var = rng.uniform(0.1, 1, size=(20, 60)).astype(np.float32)
out = sfb.LaplacianStage().execute(means, var, ctx=ctx)
print(out.nnz)
idx, l2, nrm = m.map_to_subcentroids(rng.uniform(size=4500), ctx.matrix(rng.normal(size=(30, 70))), rng.uniform(size=30))
m.diffuse(Lf, 0.1, 2)
print("done")
