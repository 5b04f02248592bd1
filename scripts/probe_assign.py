import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, bench
import fastoverlap_b200 as fob
ctx = fob.Context(0)
wl = bench.Blj256(); wl.setup(ctx)
A, B, _ = wl.make(3256, 0)
for k in range(3): r = ctx.per_align_pairs_full(wl.params, A, B, niter=10, nthreads=8)
print("nhost", r[-1])
