"""Quick GPU probe: periodic hot path throughput (device-resident and host API) + FP64 peak."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastoverlap_b200 as fob

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = fob.Context(0)
print("device", ctx.device_info())
print("fp64 peak TFLOP/s", ctx.measure_fp64_peak())
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "periodic_blj256.npz"))
box = g["box"]; perm = [np.arange(204), np.arange(204, 256)]
rng = np.random.default_rng(256)
base = g["pos1"]
pos1 = np.broadcast_to(base, (P, 256, 3)).copy()
pos2 = pos1 + rng.uniform(0, 1, size=(P, 1, 3)) * box + rng.normal(scale=0.05, size=(P, 256, 3))
al = fob.PeriodicAlign(256, box, perm, ctx=ctx)
p = al._params()
# host API
for _ in range(2):
    t = time.perf_counter(); r = ctx.per_align_pairs(p, pos1, pos2); dt = time.perf_counter() - t
print("host API: %d pairs in %.4f s -> %.0f pairs/s" % (P, dt, P / dt))
# device API
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
dA = torch.from_numpy(pos1).cuda(); dB = torch.from_numpy(pos2).cuda()
bi = torch.empty((P, 3), dtype=torch.int64, device="cuda"); bv = torch.empty(P, dtype=torch.float64, device="cuda")
fr = torch.empty((P, 3), dtype=torch.float64, device="cuda")
for rep in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    ctx.per_align_pairs_dev(p, dA.data_ptr(), dB.data_ptr(), P, bi.data_ptr(), bv.data_ptr(), fr.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("dev API: %d pairs in %.3f ms -> %.0f pairs/s" % (P, ms, P / ms * 1e3))
assert np.array_equal(bi.cpu().numpy(), r[0])
