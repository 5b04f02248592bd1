"""Under torchrun: end-to-end LJ38 step per rank -- full alignment (host Kearsley) vs hot path only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
import fastoverlap_b200 as fob
rank, local, world = bench.dist_env()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = 65536
wl = bench.Lj38()
ctx = fob.Context(local)
wl.setup(ctx)
A, B, _ = wl.make(P, rank)
hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
nt = max(1, bench.host_cores() // world)
def timed(fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    w = torch.tensor([(time.perf_counter() - t0) / n * 1e3], device="cuda")
    if world > 1:
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    return float(w.item())
out1 = [None]
def full():
    out1[0] = wl.run_host_full(ctx, hA.numpy(), hB.numpy(), nt, out=out1[0])
def hot():
    wl.run_host_hot(ctx, hA.numpy(), hB.numpy())
res = {"full": timed(full), "hot path": timed(hot)}
fob._lib.load_library().fo_host_refine_counters
if rank == 0:
    print("world %d, %d host threads per rank: " % (world, nt) +
          ", ".join("%s %.1f ms (%.2f M pairs/s)" % (k, v, world * P / v / 1e3) for k, v in res.items()), flush=True)
