"""Where does the end-to-end (host buffer) step spend its time?  Wall clock of fo_per_align_pairs /
fo_sph_align_pairs vs the CUDA-event sum of its kernels (fo_profile_*), pinned vs pageable inputs."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import fastoverlap_b200 as fob  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "blj256"
wl = bench.WORKLOADS[name]()
ctx = fob.Context(0)
wl.setup(ctx)
P = wl.default_pairs
A, B, _ = wl.make(P, 0)
hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
for label, a, b in (("pinned", hA.numpy(), hB.numpy()), ("pageable", A, B)):
    for _ in range(3):
        wl.run_host(ctx, a, b)
    ctx.profile_begin()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        wl.run_host(ctx, a, b)
    wall = (time.perf_counter() - t0) / n * 1e3
    prof = ctx.profile_end()
    ksum = sum(v[0] for v in prof.values()) / n
    print("%s %s: wall %.3f ms/step, kernels %.3f ms/step, gap %.3f ms; %s" % (
        name, label, wall, ksum, wall - ksum, {k: round(v[0] / n, 3) for k, v in prof.items()}))
dA, dB = hA.cuda(), hB.cuda()
out = wl.out_tensors(torch, P)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    wl.run_dev(ctx, dA, dB, P, out)
torch.cuda.synchronize()
ctx.profile_begin()
t0 = time.perf_counter()
for _ in range(10):
    wl.run_dev(ctx, dA, dB, P, out)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10 * 1e3
prof = ctx.profile_end()
print("%s device-resident: wall %.3f ms/step, kernels %.3f ms/step" % (name, wall, sum(v[0] for v in prof.values()) / 10))
