"""A/B probe on one box: device-resident hot-path rate and per-kernel-class CUDA-event times.
    python scripts/probe_ab.py [blj256|lj38] [pairs] [reps] [option=value ...]
Library under test: FASTOVERLAP_B200_LIB (default: the in-tree build)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastoverlap_b200 as fob

wl = sys.argv[1] if len(sys.argv) > 1 else "blj256"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
ctx = fob.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for kv in sys.argv[4:]:  # library options, e.g. per_xf_variant=4
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
gold = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
rng = np.random.default_rng(256)
if wl == "blj256":
    g = np.load(os.path.join(gold, "periodic_blj256.npz"))
    box = g["box"]; perm = [np.arange(204), np.arange(204, 256)]
    pos1 = np.broadcast_to(g["pos1"], (P, 256, 3)).copy()
    pos2 = pos1 + rng.uniform(0, 1, size=(P, 1, 3)) * box + rng.normal(scale=0.05, size=(P, 256, 3))
    al = fob.PeriodicAlign(256, box, perm, ctx=ctx)
    p = al._params()
    dA = torch.from_numpy(pos1).cuda(); dB = torch.from_numpy(pos2).cuda()
    bi = torch.empty((P, 3), dtype=torch.int64, device="cuda"); bv = torch.empty(P, dtype=torch.float64, device="cuda")
    fr = torch.empty((P, 3), dtype=torch.float64, device="cuda")
    run = lambda: ctx.per_align_pairs_dev(p, dA.data_ptr(), dB.data_ptr(), P, bi.data_ptr(), bv.data_ptr(), fr.data_ptr())
else:
    g = np.load(os.path.join(gold, "spherical_lj38.npz"))
    base = g["pos1"] - g["pos1"].mean(0)
    pos1 = base[None] + rng.normal(scale=0.05, size=(P, 38, 3)); pos1 -= pos1.mean(1, keepdims=True)
    pos2 = base[None] + rng.normal(scale=0.05, size=(P, 38, 3)); pos2 -= pos2.mean(1, keepdims=True)
    ctx.set_perm([np.arange(38)], 38)
    dA = torch.from_numpy(pos1).cuda(); dB = torch.from_numpy(pos2).cuda()
    bi = torch.empty((P, 2, 3), dtype=torch.int64, device="cuda"); bv = torch.empty((P, 2), dtype=torch.float64, device="cuda")
    fr = torch.empty((P, 2, 3), dtype=torch.float64, device="cuda")
    run = lambda: ctx.sph_align_pairs_dev(dA.data_ptr(), dB.data_ptr(), P, 38, 15, 0.3, True, bi.data_ptr(), bv.data_ptr(), fr.data_ptr())
for _ in range(3):
    run()
torch.cuda.synchronize()
import subprocess
def timed():
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "100"],
                       stdout=subprocess.PIPE, text=True)
blocks = [timed() for _ in range(4)]
smi.terminate()
lines = [l.strip() for l in smi.stdout.read().splitlines() if l.strip()]
print("blocks ms/step:", [round(b, 3) for b in blocks], "| smi:", lines[:: max(1, len(lines) // 6)][:8])
ms = min(blocks)
ctx.profile_begin()
for _ in range(4):
    run()
torch.cuda.synchronize()
prof = ctx.profile_end()
print("%s lib=%s P=%d: %.3f ms/step -> %.0f pairs/s; per class ms/step: %s; checksum %d %.6f" % (
    wl, os.environ.get("FASTOVERLAP_B200_LIB", "in-tree"), P, ms, P / ms * 1e3,
    {k: round(v[0] / 4, 3) for k, v in prof.items()}, int(bi.sum().item()), float(bv.sum().item())))
