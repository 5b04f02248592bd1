"""Host-side helpers that stay on the CPU (north_star: "The Hungarian/munkres permutation step
and the Kabsch refinement stay on the host").

Mirrors the names of reference fastoverlap/utils.py so the drop-in classes read the same:
find_best_permutation (utils.py:82-167), findrotation / q2mx (utils.py:169-275),
_next_fast_len (:278-313), findMax (:319-338), indtoEuler (:340-345), calcThetaPhiR (:439-445),
EulerM (:447-460).  The linear assignment is solved with scipy's Jonker-Volgenant
implementation (the reference uses `munkres` or pele's JV; any exact LAP solver returns the same
optimum up to degenerate ties).
"""
import numpy as np
from numpy import cos, sin, pi
from scipy.optimize import linear_sum_assignment
from scipy.spatial.distance import cdist


def lap(cost):
    """Solve the linear assignment problem; returns the column chosen for each row
    (reference utils.py:34-41)."""
    r, c = linear_sum_assignment(np.asarray(cost))
    out = np.empty(len(r), dtype=int)
    out[r] = c
    return out.tolist()


def _make_cost_matrix(X1, X2):
    """Squared-distance cost matrix (reference utils.py:48-56)."""
    return cdist(X1, X2, 'sqeuclidean')


def find_best_permutation(X1, X2, permlist=None, user_cost_matrix=_make_cost_matrix, reshape=True):
    """Permutation of X2 that best matches X1, group by group (reference utils.py:82-167).
    Returns (dist, perm); as in the reference fallback, `dist` is not meaningful --
    callers recompute the distance."""
    if reshape:
        X1 = X1.reshape([-1, 3])
        X2 = X2.reshape([-1, 3])
    if permlist is None:
        permlist = [list(range(len(X1)))]
    newperm = list(range(len(X1)))
    for atomlist in permlist:
        if len(atomlist) == 0:
            continue
        atomlist = np.asarray(atomlist)
        perm = lap(user_cost_matrix(X1[atomlist], X2[atomlist]))
        for i, atom in enumerate(atomlist):
            newperm[atom] = atomlist[perm[i]]
    return -1.0, newperm


def q2mx(qin):
    """Quaternion -> rotation matrix (reference utils.py:255-275)."""
    Q = qin / np.linalg.norm(qin)
    q0, q1, q2, q3 = Q
    return np.array([
        [2. * (0.5 - q2 * q2 - q3 * q3), 2. * (q1 * q2 - q0 * q3), 2. * (q1 * q3 + q0 * q2)],
        [2. * (q1 * q2 + q0 * q3), 2. * (0.5 - q1 * q1 - q3 * q3), 2. * (q2 * q3 - q0 * q1)],
        [2. * (q1 * q3 - q0 * q2), 2. * (q2 * q3 + q0 * q1), 2. * (0.5 - q1 * q1 - q2 * q2)]])


def findrotation(x1, x2, align_com=True):
    """Kearsley quaternion fit: rotation that best maps x2 onto x1 and the resulting distance
    (reference utils.py:169-253; Kearsley, Acta Cryst. A 45, 208 (1989))."""
    x1 = np.array(x1, dtype=float).reshape(-1, 3)
    x2 = np.array(x2, dtype=float).reshape(-1, 3)
    if x1.shape != x2.shape:
        raise ValueError("dimension of arrays does not match")
    if align_com:
        x1 = x1 - x1.mean(axis=0)
        x2 = x2 - x2.mean(axis=0)
    m = x1 - x2
    p = x1 + x2
    xm, ym, zm = m[:, 0], m[:, 1], m[:, 2]
    xp, yp, zp = p[:, 0], p[:, 1], p[:, 2]
    Q = np.empty((4, 4))
    Q[0, 0] = np.sum(xm * xm + ym * ym + zm * zm)
    Q[0, 1] = Q[1, 0] = np.sum(ym * zp - yp * zm)
    Q[0, 2] = Q[2, 0] = np.sum(xp * zm - xm * zp)
    Q[0, 3] = Q[3, 0] = np.sum(xm * yp - xp * ym)
    Q[1, 1] = np.sum(yp * yp + zp * zp + xm * xm)
    Q[1, 2] = Q[2, 1] = np.sum(xm * ym - xp * yp)
    Q[1, 3] = Q[3, 1] = np.sum(xm * zm - xp * zp)
    Q[2, 2] = np.sum(xp * xp + zp * zp + ym * ym)
    Q[2, 3] = Q[3, 2] = np.sum(ym * zm - yp * zp)
    Q[3, 3] = np.sum(xp * xp + yp * yp + zm * zm)
    eigs, vecs = np.linalg.eigh(Q)
    eigmin = eigs[0]
    if eigmin < 0.:
        eigmin = 0. if abs(eigmin) < 1e-6 else -eigmin
    return np.sqrt(eigmin), q2mx(vecs[:, 0])


def _next_fast_len(target):
    """Smallest 5-smooth integer >= target (reference utils.py:278-313)."""
    target = int(target)
    if target <= 6:
        return target
    best = None
    p5 = 1
    while p5 < 2 * target:
        p35 = p5
        while p35 < 2 * target:
            v = p35
            while v < target:
                v *= 2
            if best is None or v < best:
                best = v
            p35 *= 3
        p5 *= 5
    return best


def findMax(a):
    """Interpolated maximum of a periodic N-d array in fractional index units
    (reference utils.py:319-338).  Host version for arrays that are already on the host."""
    a = np.asanyarray(a)
    shape = a.shape
    ind = np.unravel_index(a.argmax(), shape)
    d = np.empty(len(shape))
    for ax in range(len(shape)):
        ip = list(ind)
        im = list(ind)
        ip[ax] = (ind[ax] + 1) % shape[ax]
        im[ax] = ind[ax] - 1
        y1, y2, y3 = np.abs(a[tuple(ip)]), np.abs(a[tuple(ind)]), np.abs(a[tuple(im)])
        d[ax] = (y3 - y1) / (2 * (2 * y2 - y1 - y3))
    return np.array(ind) - d


def indtoEuler(ind, n):
    """Grid index -> Euler angles (reference utils.py:340-345, soft.py:127-130)."""
    ind = np.atleast_2d(ind)
    rot = np.array([2 * pi / n, pi / n, 2 * pi / n]) * ind
    rot[:, 1] += 0.5 * pi / n
    return rot.squeeze()


def calcThetaPhiR(pos):
    """(theta, phi, r) of each row (reference utils.py:439-445)."""
    pos = np.atleast_2d(pos)
    X, Y, Z = pos.T
    R = np.linalg.norm(pos, axis=1)
    return np.arccos(Z / R), np.arctan2(Y, X), R


def EulerM(a, b, y):
    """ZYZ Euler rotation matrix in the reference's convention (utils.py:447-460)."""
    sina, cosa = sin(a), cos(a)
    sinb, cosb = sin(b), cos(b)
    siny, cosy = sin(y), cos(y)
    Ma = np.array(((cosa, -sina, 0), (sina, cosa, 0), (0, 0, 1)))
    Mb = np.array(((cosb, 0, -sinb), (0, 1, 0), (sinb, 0, cosb)))
    My = np.array(((cosy, -siny, 0), (siny, cosy, 0), (0, 0, 1)))
    return My.dot(Mb).dot(Ma)


def BruteOverlap(pos1, pos2, scale):
    """Exact O(N^2) Gaussian overlap at the identity (reference utils.py:402-406); test helper."""
    pos1 = np.atleast_2d(pos1)
    pos2 = np.atleast_2d(pos2)
    rs2 = cdist(pos1, pos2, 'sqeuclidean')
    return np.exp(-rs2 / 4 / scale ** 2).sum() * (pi * scale ** 2) ** 1.5


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


def norm_harmonicBasis(n, l, r0):
    """Normalisation N_nl of the harmonic-oscillator radial basis function (reference utils.py:408-412)."""
    from scipy.special import factorial, gamma
    return np.sqrt(2 * factorial(n) * r0 ** (-2 * l - 3) / gamma(1.5 + n + l))


def coeffs_harmonicBasis(n, l, r0):
    """Coefficients g_s, s = 0 .. 2n+l, of  N_nl r^l L_n^{l+1/2}(r^2) = sum_s g_s r^s  (reference
    utils.py:414-427).  As in the reference the Laguerre argument is not scaled by r0 (exact for the
    default harmscale r0 = 1); the device kernel does not use this expansion.
    L_n^a(x) = sum_k (-1)^k binom(n + a, n - k) x^k / k!."""
    from scipy.special import binom, factorial
    k = np.arange(n + 1)
    out = np.zeros(2 * n + l + 1)
    out[l::2] = (-1.0) ** k * binom(n + l + 0.5, n - k) / factorial(k)
    return out * norm_harmonicBasis(n, l, r0)
