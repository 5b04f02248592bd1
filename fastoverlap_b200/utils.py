"""Helper names of reference fastoverlap/utils.py that the drop-in classes and their callers use, as thin
calls into the native library (include/fastoverlap_b200.h) -- the numerics live in csrc/:

  find_best_permutation (utils.py:82-167)   fo_host_best_permutation  (Jonker-Volgenant LAP per group)
  findrotation          (utils.py:169-253)  fo_host_kearsley          (Kearsley quaternion fit)
  _next_fast_len        (utils.py:278-313)  fo_next_fast_len
  findMax               (utils.py:319-338)  host arrays only; on the alignment path the interpolated maximum
                                            comes fused out of the transform kernels
  findPeaks             (utils.py:366-396)  fo_grid_find_peaks (peaks.py)
  indtoEuler / EulerM / calcThetaPhiR       two-line conversions between grid indices, angles and matrices
"""
import numpy as np

from . import _lib


def find_best_permutation(X1, X2, permlist=None, user_cost_matrix=None, reshape=True, box=None):
    """Permutation of X2 that best matches X1, group by group.  Returns (-1.0, perm) like the reference's
    munkres fallback (its distance is not meaningful; callers recompute it).  box: minimum-image
    distance costs (the periodic classes); user_cost_matrix: any other cost, solved with scipy's LAP."""
    X1 = np.asarray(X1, float).reshape(-1, 3)
    X2 = np.asarray(X2, float).reshape(-1, 3)
    if user_cost_matrix is None:
        return -1.0, _lib.host_best_permutation(X1, X2, permlist, box).tolist()
    from scipy.optimize import linear_sum_assignment
    perm = np.arange(len(X1))
    for g in ([perm.copy()] if permlist is None else permlist):
        g = np.asarray(g, int)
        if len(g):
            r, c = linear_sum_assignment(np.asarray(user_cost_matrix(X1[g], X2[g])))
            perm[g[r]] = g[c]
    return -1.0, perm.tolist()


def findrotation(x1, x2, align_com=True):
    """(distance, rotation matrix) of the best rotation of x2 onto x1 about their centroids."""
    if not align_com:
        raise NotImplementedError("findrotation always removes the centroids (as every caller in the reference does)")
    return _lib.host_kearsley(x1, x2)


def _next_fast_len(target):
    """Smallest 5-smooth integer >= target."""
    return int(_lib.load_library().fo_next_fast_len(int(target)))


def findMax(a):
    """Interpolated maximum of a periodic N-d host array in fractional index units: arg-max (first in C order),
    then per axis the vertex of the parabola through |a| at the maximum and its two periodic neighbours."""
    a = np.abs(np.asanyarray(a))
    top = np.array(np.unravel_index(int(a.argmax()), a.shape))
    centre = a[tuple(top)]
    out = top.astype(float)
    for ax, n in enumerate(a.shape):
        step = np.zeros(a.ndim, int)
        step[ax] = 1
        up, down = a[tuple((top + step) % n)], a[tuple((top - step) % n)]
        out[ax] -= (down - up) / (2 * (2 * centre - up - down))
    return out


def indtoEuler(ind, n):
    """Fractional grid index (a, b, g) of a (n, n, n) Euler grid -> angles: a, g in steps of 2 pi / n from 0,
    b in steps of pi / n from pi / 2n."""
    f = np.pi / n
    return (np.atleast_2d(np.asarray(ind, float)) * (2 * f, f, 2 * f) + (0.0, 0.5 * f, 0.0)).squeeze()


def calcThetaPhiR(pos):
    """(polar angle, azimuth, radius) of every row."""
    x, y, z = np.atleast_2d(pos).T
    r = np.sqrt(x * x + y * y + z * z)
    return np.arccos(z / r), np.arctan2(y, x), r


def _rz(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def EulerM(a, b, y):
    """Rotation matrix of the Euler angles in the reference's convention: Rz(y) . Ry'(b) . Rz(a) with
    Ry'(b) = [[cos b, 0, -sin b], [0, 1, 0], [sin b, 0, cos b]]; row vectors rotate as X . EulerM."""
    cb, sb = np.cos(b), np.sin(b)
    return _rz(y) @ np.array([[cb, 0.0, -sb], [0.0, 1.0, 0.0], [sb, 0.0, cb]]) @ _rz(a)


def BruteOverlap(pos1, pos2, scale):
    """Exact O(N^2) Gaussian overlap at the identity; test helper."""
    d = np.atleast_2d(pos1)[:, None, :] - np.atleast_2d(pos2)[None, :, :]
    return np.exp(-(d * d).sum(2) / (4 * scale ** 2)).sum() * (np.pi * scale ** 2) ** 1.5


def oh_operations():
    """The 48 operations of the octahedral group O_h as (48, 3, 3) signed permutation matrices,
    identity first, the 24 proper rotations before the 24 improper ones.  The reference holds the
    same set as a data table (OHOPS, alignutils.f90:830-987; OHOPSMAT, fastbulk.f90:102-248); the
    order only decides which of several exactly equal distances is reported."""
    import itertools
    ops = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1.0, -1.0), repeat=3):
            R = np.zeros((3, 3))
            for r in range(3):
                R[r, perm[r]] = signs[r]
            ops.append(R)
    ops.sort(key=lambda R: (np.linalg.det(R) < 0, not np.array_equal(R, np.eye(3))))
    return np.array(ops)
