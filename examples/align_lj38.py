#!/usr/bin/env python
"""Align the two LJ38 minima of the reference's examples/LJ38 (the coordinates are stored in
tests/golden/spherical_lj38.npz) with the drop-in classes -- the counterpart of the reference's
examples/alignSpherical.py.  Expected distance: 1.4767670631638872 (sphericalAlignment.py:692).

    python examples/align_lj38.py            (needs a CUDA device)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fastoverlap_b200 import SphericalAlign, SphericalHarmonicAlign, SphericalAlignFortran  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "spherical_lj38.npz"))
pos1, pos2 = g["pos1"], g["pos2"]

soap = SphericalAlign(0.3, 15)                        # scale, Jmax: same constructor as the reference
dist, X1, X2 = soap(pos1, pos2)
print("SphericalAlign           distance %.13f   |X1 - X2| %.13f" % (dist, np.linalg.norm(X1 - X2)))

harm = SphericalHarmonicAlign(0.3, 1.0, 20, 15)       # scale, harmscale, nmax, Jmax
print("SphericalHarmonicAlign   distance %.13f" % harm(pos1, pos2)[0])

fort = SphericalAlignFortran(0.3, 15)                 # the f2py-wrapper class, bound to the C ABI
dist, X1, X2, rmat = fort(pos1, pos2)
print("SphericalAlignFortran    distance %.13f   det(R) %+.3f" % (dist, np.linalg.det(rmat)))

# the ten best rotations of the overlap grid (device top-k peak search) and their distances
I = soap.calcSO3Coeffs(*soap.COM_shift(pos1, pos2))
Rs, amplitude = soap.findRotations(I, nrot=10)[:2]
for R, a in zip(Rs, amplitude):
    print("  rotation (%.4f %.4f %.4f)  peak %.4f  ->  distance %.6f" % (
        R[0], R[1], R[2], a, soap.refine(*soap.COM_shift(pos1, pos2), R)[0]))

# a batch: 1000 perturbed, rotated, permuted copies in one GPU call + the native host refinement pool
rng = np.random.default_rng(0)
A = pos1[None] + rng.normal(scale=0.05, size=(1000, 38, 3))
B = np.array([a[rng.permutation(38)] for a in A + rng.normal(scale=0.05, size=A.shape)])
dists, eulers = soap.align_batch(A, B)
print("batch of 1000 pairs: median distance %.4f" % np.median(dists))
