"""GPU probe: BLJ256 with a fine k-grid (bench.py `blj256_fine`: n = 16, F = 72; optionally other n / F) -- kernel-class
times of the host-buffer call from the library's CUDA-event profile, and the full-alignment rate.
usage: python scripts/probe_fine.py [pairs] [n] [F] [option=value ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastoverlap_b200 as fob  # noqa: E402
from fastoverlap_b200 import _lib  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if "=" not in a]
    opts = [a.split("=") for a in sys.argv[1:] if "=" in a]
    P = int(args[0]) if len(args) > 0 else 4096
    n = int(args[1]) if len(args) > 1 else 16
    F = int(args[2]) if len(args) > 2 else 72
    ctx = fob.Context(0)
    for k, v in opts:
        ctx.set_option(k, int(v))
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "periodic_blj256.npz"))
    box = np.full(3, 5.975206329)
    perm = [np.arange(204), np.arange(204, 256)]
    rng = np.random.default_rng(256)
    A = np.broadcast_to(g["pos1"], (P, 256, 3)).copy()
    B = A + rng.uniform(0, 1, size=(P, 1, 3)) * box + rng.normal(scale=0.05, size=(P, 256, 3))
    ctx.set_perm(perm, 256)
    p = _lib.Context.per_params(256, box, n, F, 0.3)
    ctx.per_align_pairs(p, A[:64], B[:64])
    for rep in range(3):
        ctx.profile_begin()
        t = time.perf_counter()
        r = ctx.per_align_pairs(p, A, B)
        dt = time.perf_counter() - t
        prof = ctx.profile_end()
        print("n=%d F=%d P=%d hot path, host buffers: %.2f ms -> %.0f pairs/s; classes (ms, launches): %s"
              % (n, F, P, dt * 1e3, P / dt, {k: (round(v[0], 3), v[1]) for k, v in prof.items()}), flush=True)
    for rep in range(2):
        t = time.perf_counter()
        out = ctx.per_align_pairs_full(p, A, B, niter=10, nthreads=0)
        dt = time.perf_counter() - t
        print("full alignment: %.2f ms -> %.0f pairs/s, host LAP pairs %d, median dist %.4f"
              % (dt * 1e3, P / dt, out[5], np.median(out[0])), flush=True)


if __name__ == "__main__":
    main()
