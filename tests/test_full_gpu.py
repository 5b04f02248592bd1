"""GPU: the full alignment entry points (fo_per_align_pairs_full / fo_sph_align_pairs_full: hot path + device
screening of the assignment + host pool) against the two-step path they replace (hot path through the C ABI,
then fo_host_refine_* on every pair), against scipy's LAP through the reference's loop, and against the
reference's golden distances.

Bars: permutations identical, distances / displacements bit-identical between the device-settled and the
host-settled path (same operations in the same order), 1e-8 against the reference (north_star)."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _blj_pairs(P, jitter, seed=256):
    g = golden("periodic_blj256.npz")
    box = np.ones(3) * 5.975206329
    rng = np.random.default_rng(seed)
    shift = rng.uniform(0, 1, size=(P, 1, 3)) * box
    B = g["pos1"][None] + shift + rng.normal(scale=jitter, size=(P, 256, 3))
    B -= np.round(B / box) * box
    for i in range(P):
        B[i] = B[i][np.concatenate([rng.permutation(204), 204 + rng.permutation(52)])]
    A = np.broadcast_to(g["pos1"], B.shape).copy()
    return A, B, box


@pytest.mark.parametrize("jitter,P", [(0.05, 300), (0.0, 64), (0.12, 128), (0.3, 96)])
def test_periodic_full_matches_two_step(ctx, jitter, P):
    """Bench configuration (256 atoms [204, 52], n = 9, F = 40) at the bench's jitter (every pair settled on
    the device), without noise, and at two larger jitters (a growing share goes to the host LAP pool)."""
    from fastoverlap_b200 import PeriodicAlign, _lib
    A, B, box = _blj_pairs(P, jitter)
    perm = [np.arange(204), np.arange(204, 256)]
    al = PeriodicAlign(256, box, perm, ctx=ctx)
    p = al._params()
    fr = ctx.per_align_pairs(p, A, B)[2]
    d_ref, pm_ref, s_ref = _lib.host_refine_periodic(p, perm, A, B, fr, 10, 4)
    dist, pm, disp, fr2, st, nhost = ctx.per_align_pairs_full(p, A, B, niter=10, nthreads=4)
    assert np.array_equal(fr, fr2)
    assert np.array_equal(pm, pm_ref)
    assert np.array_equal(dist, d_ref), np.abs(dist - d_ref).max()
    assert np.array_equal(disp, s_ref)
    if jitter <= 0.05:
        assert nhost == 0
        assert dist.max() < 3 * max(jitter, 1e-9) * np.sqrt(3 * 256)
    if jitter >= 0.3:
        assert nhost > 0
    # the permutation is optional
    d2 = ctx.per_align_pairs_full(p, A, B, niter=10, nthreads=4, want_perm=False)
    assert d2[1] is None and np.array_equal(d2[0], dist)
    # align_batch is the same call
    db, sb, pb = al.align_batch(A, B, nthreads=4)
    assert np.array_equal(db, dist) and np.array_equal(pb, pm) and al.last_nhost == nhost


def test_periodic_full_vs_scipy_loop(ctx):
    """Ragged groups, non-cubic box, atoms outside the cell, one atom in no group: the full path against
    scipy's linear_sum_assignment through the reference's loop (periodicAlignment.py:27-80)."""
    from scipy.optimize import linear_sum_assignment
    from fastoverlap_b200 import PeriodicAlign
    rng = np.random.default_rng(11)
    N, box = 62, np.array([4.0, 4.5, 5.0])
    groups = [np.arange(37), np.arange(37, 50), np.arange(50, 61)]   # atom 61 is in no group
    P = 40
    A = (np.stack(np.unravel_index(rng.permutation(64)[:N], (4, 4, 4)), 1)[None] + 0.5 +
         rng.uniform(-0.15, 0.15, size=(P, N, 3))) / 4 * box + rng.integers(-2, 3, size=(P, N, 3)) * box
    jit = np.where(np.arange(P) % 2 == 0, 0.03, 0.25)[:, None, None]
    B = A + rng.uniform(0, 1, size=(P, 1, 3)) * box + rng.normal(size=A.shape) * jit
    for i in range(P):
        B[i] = B[i][np.concatenate([g[0] + rng.permutation(len(g)) for g in groups] + [[61]])]
    al = PeriodicAlign(N, box, groups, ctx=ctx)
    p = al._params()
    dist, pm, disp, fr, st, nhost = ctx.per_align_pairs_full(p, A, B, niter=10, nthreads=2)
    assert 0 < nhost < P
    F = al.fshape[0]

    def mi(d):
        return d - np.rint(d / box) * box

    def bestperm(x, y):
        perm = np.arange(N)
        for g in groups:
            c = np.linalg.norm(mi(x[g][:, None, :] - y[g][None, :, :]), axis=2)
            r, cc = linear_sum_assignment(c)
            perm[g[r]] = g[cc]
        return perm

    for q in range(P):
        x, y, d = A[q], B[q], fr[q] * box / F
        save = bestperm(x, y - d)
        perm = save
        for _ in range(10):
            d = d - mi(x - (y[save] - d)).mean(0)
            perm = bestperm(x, y - d)
            if np.array_equal(perm, save):
                break
            save = perm
        d = d - mi(x - (y[perm] - d)).mean(0)
        ref = np.sqrt((mi(mi(x) - mi(y[perm] - d)) ** 2).sum())
        assert np.array_equal(perm, pm[q]), q
        assert abs(ref - dist[q]) < 1e-12 and np.allclose(d, disp[q], atol=1e-12)


def test_periodic_full_reference_pair(ctx):
    """examples/BLJ256 pair: 1.5590835031549872 and the reference's permutation through the full path."""
    from fastoverlap_b200 import PeriodicAlign
    g = golden("periodic_blj256.npz")
    al = PeriodicAlign(256, g["box"], [np.arange(204), np.arange(204, 256)], ctx=ctx)
    dists, disps, perms = al.align_batch(g["pos1"][None], g["pos2"][None])
    assert abs(dists[0] - 1.5590835031549872) < 1e-8
    assert np.array_equal(perms[0], g["perm"])
    assert np.allclose(disps[0], g["disp"], atol=1e-8)


def _lj38_pairs(P, jitter, seed=20171013):
    g = golden("spherical_lj38.npz")
    minima = [g["pos1"] - g["pos1"].mean(0), g["pos2"] - g["pos2"].mean(0)]
    rng = np.random.default_rng(seed)
    A, B = np.empty((P, 38, 3)), np.empty((P, 38, 3))
    for i in range(P):
        a = minima[i % 2] + rng.normal(scale=jitter, size=(38, 3))
        b = minima[i % 2] + rng.normal(scale=jitter, size=(38, 3))
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[w*w+x*x-y*y-z*z, 2*(x*y-w*z), 2*(x*z+w*y)],
                      [2*(x*y+w*z), w*w-x*x+y*y-z*z, 2*(y*z-w*x)],
                      [2*(x*z-w*y), 2*(y*z+w*x), w*w-x*x-y*y+z*z]])
        b = b.dot(R.T)[rng.permutation(38)]
        A[i], B[i] = a - a.mean(0), b - b.mean(0)
    return A, B


@pytest.mark.parametrize("jitter,P", [(0.05, 256), (0.0, 64), (0.25, 96)])
def test_spherical_full_matches_two_step(ctx, jitter, P):
    """Bench configuration (LJ38, Jmax = 15, sigma = 0.3, both orientations): the full path (screening AND
    Kearsley fit of the settled pairs on the device, host pool for the rest) against the hot path +
    fo_host_refine_spherical on every (pair, orientation): same orientation and permutation, distance and rotation
    matrix to rounding (the device sums the quaternion matrix in a different order and rotates with its own
    sincos; an exact copy has distance sqrt(rounding residue) ~ 1e-8, compared at that level)."""
    from fastoverlap_b200 import SphericalAlign, _lib
    from fastoverlap_b200.utils import indtoEuler
    A, B = _lj38_pairs(P, jitter)
    ctx.set_perm([np.arange(38)], 38)
    bi, bv, fr, _, st = ctx.sph_align_pairs(A, B, 15, 0.3, invert=True)
    eul = indtoEuler(fr.reshape(-1, 3), 32).reshape(fr.shape)
    d_ref, o_ref, pm_ref, r_ref = _lib.host_refine_spherical(A, B, eul, None, 4)
    dist, orient, pm, rmat, eu, st2, nhost = ctx.sph_align_pairs_full(A, B, 15, 0.3, invert=True, nthreads=4)
    assert np.array_equal(eu, eul)
    assert np.allclose(dist, d_ref, rtol=1e-11, atol=1e-12 if jitter > 0 else 1e-6) and np.array_equal(orient, o_ref)
    assert np.array_equal(pm, pm_ref) and np.allclose(rmat, r_ref, atol=1e-9 if jitter > 0 else 1e-6)
    if jitter <= 0.05:
        # the correct orientation of a perturbed copy is always settled by the screening
        assert nhost <= P
        assert np.median(dist) < 3 * max(jitter, 1e-9) * np.sqrt(3 * 38)
    sa = SphericalAlign(0.3, 15, ctx=ctx)
    db, Rb = sa.align_batch(A, B, nthreads=4)
    # align_batch re-centres the (already centred) structures: last-bit differences in the coordinates.  An exact
    # copy (jitter 0) has distance sqrt(rounding residue) ~ 1e-9: only the noise level can be compared there
    assert np.allclose(db, dist, atol=1e-12 if jitter > 0 else 1e-7)
    assert np.allclose(Rb, eul, atol=1e-9) and np.array_equal(sa.last_perms, pm)


def test_spherical_full_groups_and_reference_pair(ctx):
    """Two permutation groups; and the reference's LJ38 pair (1.4767670631638872, inverted orientation)."""
    from fastoverlap_b200 import SphericalAlign, _lib
    from fastoverlap_b200.utils import indtoEuler
    g = golden("spherical_lj38.npz")
    sa = SphericalAlign(0.3, 15, ctx=ctx)
    d, R = sa.align_batch(g["pos1"][None], g["pos2"][None])
    assert abs(d[0] - 1.4767670631638872) < 1e-8 and sa.last_orient[0] == 1
    rng = np.random.default_rng(2)
    N, P = 21, 50
    groups = [np.arange(9), np.arange(9, 21)]
    A = rng.normal(size=(P, N, 3))
    A -= A.mean(1, keepdims=True)
    B = np.empty_like(A)
    for i in range(P):
        th = rng.uniform(0, 6)
        Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
        order = np.concatenate([rng.permutation(9), 9 + rng.permutation(12)])
        B[i] = (A[i] + rng.normal(scale=0.02, size=(N, 3))).dot(Rz.T)[order]
    B -= B.mean(1, keepdims=True)
    ctx.set_perm(groups, N)
    fr = ctx.sph_align_pairs(A, B, 9, 0.5, invert=True)[2]
    eul = indtoEuler(fr.reshape(-1, 3), 20).reshape(fr.shape)
    d_ref, o_ref, pm_ref, r_ref = _lib.host_refine_spherical(A, B, eul, groups, 2)
    dist, orient, pm, rmat, eu, st, nhost = ctx.sph_align_pairs_full(A, B, 9, 0.5, invert=True, nthreads=2)
    assert np.allclose(dist, d_ref, rtol=1e-11, atol=1e-12) and np.array_equal(pm, pm_ref) and np.array_equal(orient, o_ref)
    assert np.median(dist) < 3 * 0.02 * np.sqrt(3 * N)
