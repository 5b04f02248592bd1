"""Full-alignment rate (GPU hot path + native host refinement pool), repeated, per workload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fastoverlap_b200 as fob
from fastoverlap_b200 import _lib

ctx = fob.Context(0)
for name in ("lj38", "blj256"):
    wl = bench.WORKLOADS[name]()
    wl.setup(ctx)
    A, B, _ = wl.make(4096, 0)
    nthr = os.cpu_count()
    for rep in range(4):
        t0 = time.perf_counter()
        if name == "lj38":
            X1 = A - A.mean(1, keepdims=True); X2 = B - B.mean(1, keepdims=True)
            Rs = wl.sa._grid_search(X1, X2, [np.arange(38)], True).reshape(len(X1), -1, 3)
            t1 = time.perf_counter()
            r = _lib.host_refine_spherical(X1, X2, Rs, [np.arange(38)], nthr)
        else:
            bi, bv, fr, _, st = ctx.per_align_pairs(wl.params, A, B)
            t1 = time.perf_counter()
            r = _lib.host_refine_periodic(wl.params, wl.perm, A, B, fr, 10, nthr)
        t2 = time.perf_counter()
        print(name, "rep", rep, "gpu+wrapper %.1f ms  host refine %.1f ms  -> %.0f pairs/s (threads %d)" % (
            (t1 - t0) * 1e3, (t2 - t1) * 1e3, 4096 / (t2 - t0), nthr), flush=True)
    # the public batched call, in one piece and with the chunk pipeline (batch.overlap_chunks)
    al = wl.sa if name == "lj38" else wl.al
    for chunk in (0, 512, 1024, 2048):
        best = 1e9
        for rep in range(4):
            t0 = time.perf_counter()
            d = al.align_batch(A, B, nthreads=nthr, chunk=chunk)[0]
            best = min(best, time.perf_counter() - t0)
        print(name, "align_batch chunk %4d: %.1f ms -> %.0f pairs/s (median distance %.4f)" % (
            chunk, best * 1e3, 4096 / best, float(np.median(d))), flush=True)
