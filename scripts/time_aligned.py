"""End-to-end ALIGNED pairs/s (fo_*_align_pairs_full: hot path + device screening + host pool), per workload,
against the two-step path (hot path, then the host pool on every pair).
    python scripts/time_aligned.py [pairs] [threads]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
import fastoverlap_b200 as fob
from fastoverlap_b200 import _lib

P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nthr = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
ctx = fob.Context(0)
for name in ("blj256", "lj38"):
    wl = bench.WORKLOADS[name]()
    wl.setup(ctx)
    A, B, _ = wl.make(P, 0)
    if name == "lj38":
        A -= A.mean(1, keepdims=True)
        B -= B.mean(1, keepdims=True)
    hA = torch.from_numpy(A).pin_memory()
    hB = torch.from_numpy(B).pin_memory()
    for pinned in (True, False):
        a, b = (hA.numpy(), hB.numpy()) if pinned else (A, B)
        for want_perm in (True, False):
            best, nh = 1e9, -1
            for rep in range(4):
                t0 = time.perf_counter()
                if name == "lj38":
                    r = ctx.sph_align_pairs_full(a, b, 15, 0.3, invert=True, nthreads=nthr, want_perm=want_perm)
                    d, nh = r[0], r[-1]
                else:
                    r = ctx.per_align_pairs_full(wl.params, a, b, niter=10, nthreads=nthr, want_perm=want_perm)
                    d, nh = r[0], r[-1]
                best = min(best, time.perf_counter() - t0)
            print("%s full: %d pairs, %s buffers, perm %d, %d threads: %.1f ms -> %.0f aligned pairs/s "
                  "(host LAP for %d, median distance %.4f)" % (name, P, "pinned" if pinned else "pageable",
                                                               want_perm, nthr, best * 1e3, P / best, nh,
                                                               float(np.median(d))), flush=True)
    # the two steps apart, on a slice
    n = min(P, 16384)
    t0 = time.perf_counter()
    if name == "lj38":
        from fastoverlap_b200.utils import indtoEuler
        fr = ctx.sph_align_pairs(A[:n], B[:n], 15, 0.3, invert=True)[2]
        t1 = time.perf_counter()
        _lib.host_refine_spherical(A[:n], B[:n], indtoEuler(fr.reshape(-1, 3), 32).reshape(fr.shape), None, nthr)
    else:
        fr = ctx.per_align_pairs(wl.params, A[:n], B[:n])[2]
        t1 = time.perf_counter()
        _lib.host_refine_periodic(wl.params, wl.perm, A[:n], B[:n], fr, 10, nthr)
    t2 = time.perf_counter()
    print("%s two-step on %d pairs: hot path %.1f ms, host pool on every pair %.1f ms (%.0f pairs/s of host pool)" % (
        name, n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, n / (t2 - t1)), flush=True)
    ctx.profile_begin()
    if name == "lj38":
        ctx.sph_align_pairs_full(hA.numpy(), hB.numpy(), 15, 0.3, invert=True, nthreads=nthr)
    else:
        ctx.per_align_pairs_full(wl.params, hA.numpy(), hB.numpy(), niter=10, nthreads=nthr)
    print(name, "kernel classes (ms, launches):", ctx.profile_end(), flush=True)
