"""Does the feature-graph Gram kernel overlap with the item screen when issued from a second context/thread?"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfb_loader import load
sfb = load()
c1, c2 = sfb.Context(0), sfb.Context(0)
X = c1.generate(1, 7, 1000000, 384, 1024, 0.3)
c1.synchronize()
import ctypes as C
# a view of X usable from the second context: same device memory, other stream
def T(f):
    c1.synchronize(); c2.synchronize(); t = time.perf_counter(); r = f(); c1.synchronize(); c2.synchronize(); return (time.perf_counter() - t) * 1e3, r
for it in range(3):
    t_knn, g = T(lambda: X.knn(16, 0)); g.free()
    Xv = sfb.Matrix(c2, C.c_void_p(X._h.value)); Xv.free = lambda: None
    t_col, gf = T(lambda: Xv.knn_columns(16, 0)); gf.free()
    def both():
        out = {}
        th = threading.Thread(target=lambda: out.__setitem__("g", X.knn(16, 0)))
        th.start()
        time.sleep(0.05)          # the persistent screen CTAs are resident by now
        out["gf"] = Xv.knn_columns(16, 0)
        th.join()
        return out
    t_both, o = T(both)
    o["g"].free(); o["gf"].free()
    print(f"knn {t_knn:.1f} ms, knn_columns {t_col:.1f} ms, concurrent {t_both:.1f} ms")
