"""End-to-end (host buffers, full alignment) step time of LJ38 against the chunk budget.
    python scripts/probe_e2e_chunks_sph.py [pairs] [chunk budget in MB -> option sph_chunk_mb] [option=value ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import fastoverlap_b200 as fob
OPTS = [a.split("=") for a in sys.argv[1:] if "=" in a]
sys.argv = [a for a in sys.argv if "=" not in a]
P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
wl = bench.Lj38()
ctx = fob.Context(0)
wl.setup(ctx)
MB = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # chunk budget in MB (0: the library's default)
if MB:
    ctx.set_option("sph_chunk_mb", MB)
for k, v in OPTS:
    ctx.set_option(k, int(v))
A, B, _ = wl.make(P, 0)
hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
for _ in range(2):
    wl.run_host_full(ctx, hA.numpy(), hB.numpy(), 16)
ctx.profile_begin()
t0 = time.perf_counter()
n = 4
for _ in range(n):
    wl.run_host_full(ctx, hA.numpy(), hB.numpy(), 16)
wall = (time.perf_counter() - t0) / n * 1e3
prof = ctx.profile_end()
print(str(OPTS) + " sph_chunk_mb=%s MB: %.2f ms per %d pairs (%.0f pairs/s); kernels %.2f ms %s" % (
    (MB or "default"), wall, P, P / wall * 1e3, sum(v[0] for v in prof.values()) / n,
    {k: round(v[0] / n, 2) for k, v in prof.items()}))
