import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, bench
import fastoverlap_b200 as fob
ctx = fob.Context(0)
wl = bench.Blj256(); wl.setup(ctx)
A, B, _ = wl.make(2, 0)
al = wl.al
from fastoverlap_b200.utils import oh_operations
R = oh_operations()[17]
pos2 = A[0] @ R.T + 0.3
for k in range(3):
    t = time.perf_counter(); r = al.align_oh(A[0], pos2); dt = time.perf_counter() - t
    print("align_oh dist %.6f op-match %s  %.1f ms" % (r[0], np.allclose(r[-1], np.linalg.inv(R)) or np.allclose(r[-1], R.T), dt * 1e3))
