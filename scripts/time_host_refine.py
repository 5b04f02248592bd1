"""Times the native host refinement pool (fo_host_refine_periodic; no GPU needed) on the bench.py BLJ256
workload: base + random translation + N(0, jitter^2), permuted within species.  The fractional index
handed over is the exact translation (what the device hot path recovers to ~1e-3 cells).
usage: python scripts/time_host_refine.py [pairs] [threads] [jitter]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastoverlap_b200 import _lib  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    jitter = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
    g = np.load(os.path.join(ROOT, "tests", "golden", "periodic_blj256.npz"))
    base, N, F = g["pos1"], 256, 40
    box = np.full(3, 5.975206329)
    groups = [np.arange(204), np.arange(204, 256)]
    rng = np.random.default_rng(256)
    A = np.repeat(base[None], P, 0)
    shift = rng.uniform(0, 1, size=(P, 1, 3)) * box
    B = A + shift + rng.normal(scale=jitter, size=A.shape)
    for i in range(P):
        B[i] = B[i][np.concatenate([gg[0] + rng.permutation(len(gg)) for gg in groups])]
    frac = shift[:, 0, :] / box * F + rng.normal(scale=2e-3, size=(P, 3))
    pp = _lib.Context.per_params(N, box, 9, F, 0.3)
    best = 1e9
    for _ in range(5):
        t = time.perf_counter()
        dist, pm, disp = _lib.host_refine_periodic(pp, groups, A, B, frac, nthreads=nt)
        best = min(best, time.perf_counter() - t)
    print("P=%d threads=%d jitter=%.3f: %.1f us per pair per thread, %.0f pairs/s; mean dist %.4f "
          "(noise level %.4f); checksum %.12f" % (P, nt, jitter, best / P * nt * 1e6, P / best, dist.mean(),
                                                  jitter * np.sqrt(3 * N), dist.sum()))


if __name__ == "__main__":
    main()
