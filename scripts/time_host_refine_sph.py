"""Times fo_host_refine_spherical (no GPU needed) on LJ38-like pairs: orientation 0 carries the right Euler
angles (jittered copy, permuted), orientation 1 (the inverted structure) random ones -- the usual case of
one easy and one hard assignment per pair.  usage: python scripts/time_host_refine_sph.py [pairs] [threads]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastoverlap_b200 import _lib  # noqa: E402
from fastoverlap_b200.utils import EulerM  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    g = np.load(os.path.join(ROOT, "tests", "golden", "spherical_lj38.npz"))
    X = g["pos1"] - g["pos1"].mean(0)
    N = len(X)
    rng = np.random.default_rng(3)
    A = np.repeat(X[None], P, 0) + rng.normal(scale=0.05, size=(P, N, 3))
    A -= A.mean(1, keepdims=True)
    eul = np.zeros((P, 2, 3))
    B = np.empty_like(A)
    for q in range(P):
        a, b, c = rng.uniform(0, 2 * np.pi), rng.uniform(0.2, np.pi - 0.2), rng.uniform(0, 2 * np.pi)
        B[q] = (X @ EulerM(a, b, c).T)[rng.permutation(N)]
        eul[q, 0] = (a, b, c)
        eul[q, 1] = rng.uniform(0, 3, 3)
    for no in (1, 2):
        best = 1e9
        for _ in range(5):
            t = time.perf_counter()
            d, o, pm, rm = _lib.host_refine_spherical(A, B, eul[:, :no], nthreads=nt)
            best = min(best, time.perf_counter() - t)
        print("orientations %d, threads %d: %.1f us per pair per thread; mean dist %.4f (noise level %.4f), "
              "checksum %.12f" % (no, nt, best / P * nt * 1e6, d.mean(), 0.05 * np.sqrt(3 * N), d.sum()))


if __name__ == "__main__":
    main()
