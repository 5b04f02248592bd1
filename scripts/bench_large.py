#!/usr/bin/env python
"""BASELINE.json configs[3]: large-cluster SphericalAlign (LJ1000-size synthetic clusters, SURVEY 8d
C4: 1000 lattice points with spacing 1.12 inside a sphere + N(0, 0.03^2) jitter; partner = rotated +
permuted + N(0, 0.05^2) copy; sigma = 0.37) at Jmax 31 / 63.  Prints per-kernel-class times."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(P, N=1000, seed=1000):
    rng = np.random.default_rng(seed)
    m = int(np.ceil((3 * N / (4 * np.pi)) ** (1 / 3))) + 2
    g = np.arange(-m, m + 1) * 1.12
    pts = np.array(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1).T
    pts = pts[np.argsort(np.linalg.norm(pts, axis=1), kind="stable")[:N]]
    A = np.empty((P, N, 3))
    B = np.empty((P, N, 3))
    for i in range(P):
        a = pts + rng.normal(scale=0.03, size=pts.shape)
        a -= a.mean(0)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[w*w+x*x-y*y-z*z, 2*(x*y-w*z), 2*(x*z+w*y)],
                      [2*(x*y+w*z), w*w-x*x+y*y-z*z, 2*(y*z-w*x)],
                      [2*(x*z-w*y), 2*(y*z+w*x), w*w-x*x-y*y+z*z]])
        b = (a + rng.normal(scale=0.05, size=a.shape)).dot(R.T)[rng.permutation(N)]
        A[i] = a
        B[i] = b - b.mean(0)
    return A, B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--natoms", type=int, default=1000)
    ap.add_argument("--jmax", type=int, nargs="+", default=[31, 63])
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--align", action="store_true", help="also run the full alignment of pair 0")
    args = ap.parse_args()
    import fastoverlap_b200 as fob
    ctx = fob.Context(0)
    A, B = make(args.pairs, args.natoms)
    ctx.set_perm([np.arange(args.natoms)], args.natoms)
    for J in args.jmax:
        ctx.sph_align_pairs(A[:1], B[:1], J, 0.37, invert=True)  # warm-up: Wigner table, scratch
        ctx.profile_begin()
        t = time.perf_counter()
        bi, bv, fr, _, st = ctx.sph_align_pairs(A, B, J, 0.37, invert=True)
        dt = time.perf_counter() - t
        prof = ctx.profile_end()
        L = J
        nnz = (L + 1) * (2 * L + 1) * (2 * L + 3) // 3
        N = args.natoms
        flops_coef = 2.0 * N * N * (L + 1) * (L + 2) + 8.0 * N * sum((l + 1) ** 2 for l in range(L + 1))
        out = {"natoms": N, "Jmax": J, "pairs": args.pairs, "pairs_per_s": args.pairs / dt,
               "ms_per_pair": dt / args.pairs * 1e3,
               "kernel_ms": {k: v[0] for k, v in prof.items()},
               "coef_gemm_tflops": flops_coef * args.pairs / (prof.get("sph_coef", (1e30, 0))[0] * 1e-3) / 1e12,
               "best_val": bv[0].tolist(), "status": int(st.max())}
        if args.align:
            sa = fob.SphericalAlign(0.37, J, ctx=ctx)
            t = time.perf_counter()
            d = sa(A[0], B[0])[0]
            out["align_dist"] = float(d)
            out["align_s"] = time.perf_counter() - t
            out["noise_norm"] = float(0.05 * np.sqrt(3 * N))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
