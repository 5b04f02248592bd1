"""GPU parity of the spherical hot path: CUDA (through the C ABI) vs golden vectors from the
unmodified reference and vs the C oracle on seeded inputs.

Tolerances (BASELINE.json north_star): overlap grids within 1e-10 relative (to the grid max),
identical best grid index, final distance after the same host permutation step within 1e-8."""
import numpy as np
import pytest

import oracle
from conftest import golden, groups_from

pytestmark = pytest.mark.gpu

GRID_RTOL = 1e-10
DIST_ATOL = 1e-8


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max()


def test_wigner_table(ctx):
    s = golden("soft_tables.npz")
    for bw in (4, 8, 11, 16):
        assert np.abs(ctx.sph_wigner_table(bw - 1) - s["Ds_%d" % bw]).max() < 1e-12, bw


def test_soft_isoft_complex_and_roundtrip(ctx):
    from fastoverlap_b200 import SOFT
    s = golden("soft_tables.npz")
    for bw in (4, 8, 11, 16):
        soft = SOFT(bw, ctx=ctx)
        out = soft.iSOFT(s["flmm_%d" % bw])
        assert rel(out, s["isoft_%d" % bw]) < 1e-12
        back = soft.SOFT(out)      # SOFT(iSOFT(f)) == f  (SURVEY section 4 round-trip KAT)
        assert rel(back, s["flmm_%d" % bw]) < 1e-12
        assert np.abs(soft.weights - s["weights_%d" % bw]).max() < 1e-15


def test_lj38_direct_coeffs_grid_argmax(ctx):
    from fastoverlap_b200 import SphericalAlign
    g = golden("spherical_lj38.npz")
    X1 = g["pos1"] - g["pos1"].mean(0)
    X2 = g["pos2"] - g["pos2"].mean(0)
    for J in (14, 15):
        k = "J%d_" % J
        sa = SphericalAlign(0.3, J, ctx=ctx)
        I = sa.calcSO3Coeffs(X1, X2)
        assert rel(I, g[k + "Ilmm"]) < 1e-12
        assert rel(sa.calcSO3Coeffs(X1, -X2), g[k + "Ilmm_inv"]) < 1e-12
        bi, bv, fr, grid, st = ctx.sph_align_pairs(X1, X2, J, 0.3, invert=True, want_grid=True)
        assert st[0] == 0
        assert rel(grid[0, 0], g[k + "grid"]) < GRID_RTOL
        assert rel(grid[0, 1], g[k + "grid_inv"]) < GRID_RTOL
        assert tuple(bi[0, 0]) == tuple(g[k + "argmax"])
        assert tuple(bi[0, 1]) == tuple(g[k + "argmax_inv"])
        assert np.allclose(fr[0, 0], g[k + "findmax"].real, atol=1e-7)
        assert np.allclose(fr[0, 1], g[k + "findmax_inv"].real, atol=1e-7)
        assert abs(bv[0, 0] - g[k + "grid"].max()) < 1e-10 * g[k + "grid"].max()
        # iSOFT of the reference's own coefficients
        bi2, bv2, fr2, grid2 = ctx.sph_isoft_argmax(g[k + "Ilmm"], J, want_grid=True)
        assert rel(grid2[0, 0], g[k + "grid"]) < GRID_RTOL
        assert tuple(bi2[0, 0]) == tuple(g[k + "argmax"])


def test_lj38_known_answer(ctx):
    """sphericalAlignment.py:680-706: distance should = 1.4767, also for the inversion isomer."""
    from fastoverlap_b200 import SphericalAlign, SphericalHarmonicAlign
    g = golden("spherical_lj38.npz")
    for J in (14, 15):
        sa = SphericalAlign(0.3, J, ctx=ctx)
        d = sa(g["pos1"], g["pos2"])[0]
        assert abs(d - 1.4767670631638872) < DIST_ATOL
        assert abs(d - float(g["J%d_dist" % J])) < DIST_ATOL
        assert abs(sa(g["pos1"], -g["pos2"])[0] - float(g["J%d_dist_inv" % J])) < DIST_ATOL
        # per-orientation refined distances
        X1, X2 = sa.COM_shift(g["pos1"], g["pos2"])
        Rs = sa._grid_search(X1, X2, [np.arange(38)], True)
        assert abs(sa.refine(X1, X2, Rs[0])[0] - float(g["J%d_dist_normal_only" % J])) < DIST_ATOL
        assert abs(sa.refine(X1, -X2, Rs[1])[0] - float(g["J%d_dist_inverted_only" % J])) < DIST_ATOL
    # numpy orientation rule (L-BFGS-refined overlaps) gives the same answer here
    sa = SphericalAlign(0.3, 15, ctx=ctx, orientation="overlap")
    assert abs(sa(g["pos1"], g["pos2"])[0] - 1.4767670631638872) < DIST_ATOL
    sh = SphericalHarmonicAlign(0.3, 1.0, 20, 15, ctx=ctx)
    assert abs(sh(g["pos1"], g["pos2"])[0] - float(g["H_dist"])) < DIST_ATOL


def test_lj38_harmonic_path(ctx):
    from fastoverlap_b200 import SphericalHarmonicAlign
    g = golden("spherical_lj38.npz")
    X1 = g["pos1"] - g["pos1"].mean(0)
    X2 = g["pos2"] - g["pos2"].mean(0)
    sh = SphericalHarmonicAlign(0.3, 1.0, 20, 15, ctx=ctx)
    c1 = sh.calcHarmCoeff(X1)
    # vs the closed-form oracle (exact to rounding) and vs the reference's numpy values, whose own
    # cancellation noise at nmax=20 is ~1e-10 (DESIGN.md "harmonic radial integrals")
    assert rel(c1, oracle.sph_harm_coeffs(X1, 20, 15, 1.0, 0.3)) < 1e-12
    assert rel(c1, g["H_c1"]) < 5e-10
    bi, bv, fr, avg, grid = sh._bank_pair(X1, X2, [np.arange(38)], True, want_grid=True)
    assert rel(grid[0, 0], g["H_grid"]) < GRID_RTOL
    assert rel(grid[0, 1], g["H_grid_inv"]) < GRID_RTOL
    assert tuple(bi[0, 0]) == tuple(np.unravel_index(g["H_grid"].argmax(), g["H_grid"].shape))
    I = g["H_Ilmm"]
    assert abs(avg[0] - (np.abs(I) ** 2).sum()) < 1e-9 * (np.abs(I) ** 2).sum()


def test_synthetic_cases_vs_reference_golden(ctx):
    from fastoverlap_b200 import SphericalAlign, SphericalHarmonicAlign
    g = golden("spherical_synth.npz")
    for i in range(int(g["ncases"])):
        k = "c%d_" % i
        perm = groups_from(g[k + "groups"], g[k + "gsizes"])
        p1, p2 = g[k + "pos1"], g[k + "pos2"]
        J, sc = int(g[k + "Jmax"]), float(g[k + "scale"])
        sa = SphericalAlign(sc, J, perm=perm if len(perm) > 1 else None, ctx=ctx)
        X1, X2 = sa.COM_shift(p1, p2)
        assert rel(sa._coeffs(X1, X2, perm), g[k + "Ilmm"]) < 1e-12, i
        Rs = sa._grid_search(X1, X2, perm, False, want_grid=True)
        assert rel(sa._grid[0, 0], g[k + "grid"]) < GRID_RTOL, i
        assert np.allclose(sa._frac_idx[0, 0], g[k + "findmax"].real, atol=1e-7), i
        assert abs(sa(p1, p2)[0] - float(g[k + "dist"])) < DIST_ATOL, i
        sh = SphericalHarmonicAlign(sc, 1.0, 12, J, perm=perm if len(perm) > 1 else None, ctx=ctx)
        ctx.set_perm(perm, len(X1))
        C, _ = ctx.sph_harm_coeffs(np.stack([X1, X2]), 12, J, 1.0, sc)
        assert rel(C[0], g[k + "H_c1"]) < 1e-9, i
        assert rel(sh._coeffs(X1, X2, perm), g[k + "H_Ilmm"]) < 1e-9, i


@pytest.mark.parametrize("N,J,sigma,groups", [(7, 3, 0.6, None), (24, 9, 0.5, [14, 10]),
                                               (38, 15, 0.3, None), (60, 12, 0.45, [20, 25, 15]),
                                               (16, 21, 0.4, None)])
def test_batch_vs_oracle(ctx, N, J, sigma, groups):
    rng = np.random.default_rng(77 + N)
    P = 5
    A = rng.normal(size=(P, N, 3)) * 1.2
    B = rng.normal(size=(P, N, 3)) * 1.2
    A -= A.mean(1, keepdims=True)
    B -= B.mean(1, keepdims=True)
    perm = None
    if groups:
        o = np.cumsum([0] + groups)
        perm = [np.arange(o[i], o[i + 1]) for i in range(len(groups))]
    ctx.set_perm(perm if perm else [np.arange(N)], N)
    bi, bv, fr, grid, st = ctx.sph_align_pairs(A, B, J, sigma, invert=True, want_grid=True)
    obi, obv, ofr, ogrid, _ = oracle.sph_align_pairs(A, B, J, sigma, True, perm, want_grid=True)
    for p in range(P):
        for o_ in range(2):
            assert rel(grid[p, o_], ogrid[p, o_]) < GRID_RTOL
    assert np.array_equal(bi, obi)
    assert np.allclose(bv, obv, rtol=1e-11)
    assert np.allclose(fr, ofr, atol=1e-6)
    I, _ = ctx.sph_coeffs_direct(A, B, J, sigma)
    for p in range(P):
        assert rel(I[p], oracle.sph_coeffs_direct(A[p], B[p], J, sigma, perm)) < 1e-12
    bi2, bv2, fr2, _, _ = ctx.sph_align_pairs(A, B, J, sigma, invert=True)
    assert np.array_equal(bi, bi2) and np.array_equal(bv, bv2) and np.array_equal(fr, fr2)
    ctx.set_perm([np.arange(N)], N)


def test_bench_configuration_vs_oracle(ctx):
    """The bench workload itself (bench.py Lj38: perturbed LJ38 minima, random rotation + permutation, Jmax = 15,
    sigma = 0.3, both orientations): P = 64 pairs in one call -- arg-max, peak value and interpolated maximum of
    every (pair, orientation) against the oracle, the full 32^3 grids on a subsample."""
    import bench
    wl = bench.Lj38()
    A, B, _ = wl.make(64, 3)
    ctx.set_perm([np.arange(38)], 38)
    bi, bv, fr, _, st = ctx.sph_align_pairs(A, B, 15, 0.3, invert=True)
    obi, obv, ofr, _, _ = oracle.sph_align_pairs(A, B, 15, 0.3, True, None, nthreads=0)
    assert np.all(st == 0)
    assert np.array_equal(bi, obi)
    assert np.allclose(bv, obv, rtol=1e-11)
    assert np.allclose(fr, ofr, atol=1e-6)
    sub = [0, 21, 63]
    grid = ctx.sph_align_pairs(A[sub], B[sub], 15, 0.3, invert=True, want_grid=True)[3]
    ogrid = oracle.sph_align_pairs(A[sub], B[sub], 15, 0.3, True, None, want_grid=True)[3]
    for p in range(len(sub)):
        for o_ in range(2):
            assert rel(grid[p, o_], ogrid[p, o_]) < GRID_RTOL


def test_rotation_recovery(ctx):
    """sphericalAlignment.py:711-732: random cloud vs rotated + permuted copy, distance ~ 0,
    also for the inverted copy; batched API; BruteOverlap cross-check of the coefficients."""
    from fastoverlap_b200 import SphericalAlign
    from fastoverlap_b200.utils import BruteOverlap, EulerM
    rng = np.random.default_rng(11)
    N, P = 50, 6
    sa = SphericalAlign(0.5, 12, ctx=ctx)
    A = rng.normal(size=(P, N, 3)) * 2
    B = np.empty_like(A)
    for i in range(P):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        a, b, c, d = q
        R = np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                      [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                      [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])
        B[i] = (A[i].dot(R.T))[rng.permutation(N)] * (1 if i % 2 == 0 else -1)
    dists, Rs = sa.align_batch(A, B)
    assert np.all(dists < 1e-6)
    assert sa(A[0], B[0])[0] < 1e-6 and sa(A[1], B[1])[0] < 1e-6
    # sum_l I^l . D^l(identity) approximates the exact Gaussian overlap (utils.py:402-406)
    X = A[0] - A[0].mean(0)
    I = sa.calcSO3Coeffs(X, X)
    l = np.arange(13)
    tr = sum(np.trace(I[j]) for j in l).real
    assert abs(tr - BruteOverlap(X, X, 0.5)) / BruteOverlap(X, X, 0.5) < 0.05


def test_harmonic_all_vs_all(ctx):
    from fastoverlap_b200 import SphericalHarmonicAlign
    g = golden("spherical_lj38.npz")
    rng = np.random.default_rng(3)
    base = [g["pos1"], g["pos2"]]
    coords = np.array([base[i % 2] + rng.normal(scale=0.02, size=(38, 3)) for i in range(4)])
    sh = SphericalHarmonicAlign(0.3, 1.0, 20, 15, ctx=ctx)
    avg, mx, navg, nmax_ = sh.compareList(coords)
    assert np.allclose(np.diag(navg), 1) and np.allclose(np.diag(nmax_), 1)
    assert np.allclose(avg, avg.T)
    assert navg[0, 2] > navg[0, 1]  # same minimum is more similar than the other minimum
    d = sh.alignGroup(coords)
    assert d[0, 2] < 0.5 and abs(d[0, 1] - 1.4767) < 0.3


def test_edge_cases(ctx):
    from fastoverlap_b200 import FastOverlapError
    rng = np.random.default_rng(5)
    A = rng.normal(size=(2, 6, 3))
    B = rng.normal(size=(2, 6, 3))
    A[1, 3] = 0.0   # atom exactly at the origin: reference gives NaN (utils.py:444); we flag it
    ctx.set_perm([np.arange(6)], 6)
    bi, bv, fr, _, st = ctx.sph_align_pairs(A, B, 5, 0.5)
    assert st[0] == 0 and (st[1] & 2) and np.all(np.isfinite(bv))
    A[0, 0, 0] = np.nan
    bi, bv, fr, _, st = ctx.sph_align_pairs(A, B, 5, 0.5)
    assert st[0] & 1
    r = ctx.sph_align_pairs(np.zeros((0, 6, 3)), np.zeros((0, 6, 3)), 5, 0.5)
    assert r[0].shape == (0, 2, 3)
    with pytest.raises(FastOverlapError):
        ctx.sph_align_pairs(B, B, 200, 0.5)
    with pytest.raises(FastOverlapError):
        ctx.sph_align_pairs(B, B, 5, -1.0)


def test_fortran_wrapper_facade(ctx):
    """The f2py-module facade (fastoverlap_b200.f90) through the reference's wrapper classes:
    same call signatures / return tuples as sphericalAlignment.py:441-663, periodicAlignment.py:482-605."""
    import fastoverlap_b200 as fob
    g = golden("spherical_lj38.npz")
    al = fob.SphericalAlignFortran(0.3, 15)
    dist, X1, X2, rmat = al(g["pos1"], g["pos2"])
    assert abs(dist - 1.4767670631638872) < DIST_ATOL
    assert X1.shape == (38, 3) and rmat.shape == (3, 3)
    assert abs(abs(np.linalg.det(rmat)) - 1) < 1e-9
    assert abs(np.linalg.norm(X1 - X2) - dist) < 1e-9
    assert abs(al.align(g["pos1"], g["pos2"])[0] - 1.4767670631638872) < DIST_ATOL
    ah = fob.SphericalHarmonicAlignFortran(0.3, 15, 1.0, 20)
    assert abs(ah(g["pos1"], g["pos2"])[0] - 1.4767670631638872) < 1e-6
    avg, mx, navg, nmx = ah.compareList(np.array([g["pos1"], g["pos2"], g["pos1"]]))
    assert abs(navg[0, 2] - 1) < 1e-12 and navg[0, 1] < 1
    p = golden("periodic_blj256.npz")
    ap = fob.PeriodicAlignFortran(256, p["box"], perm=[np.arange(204), np.arange(204, 256)])
    dist, Y1, Y2, perm = ap.align(p["pos1"], p["pos2"], ndisps=1)
    assert abs(dist - 1.5590835031549872) < DIST_ATOL
    assert np.array_equal(perm - 1, p["perm"])
    d, aligned = ap.alignGroup(np.array([p["pos1"], p["pos2"]]))
    assert aligned.shape == (256, 3, 2, 2) and abs(d[0, 1] - 1.5590835031549872) < DIST_ATOL


@pytest.mark.parametrize("N,J", [(20, 7), (38, 15), (25, 9), (18, 21), (20, 14), (16, 13), (15, 8),
                                 (14, 6), (12, 4), (10, 2), (9, 1), (22, 11), (21, 12)])
def test_fast_and_generic_isoft_agree(ctx, N, J):
    """sph_isoft2_kernel (tensor-core, persistent) vs sph_isoft_kernel (any size)."""
    rng = np.random.default_rng(N + J)
    A = rng.normal(size=(4, N, 3))
    B = rng.normal(size=(4, N, 3))
    A -= A.mean(1, keepdims=True)
    B -= B.mean(1, keepdims=True)
    ctx.set_perm([np.arange(N)], N)
    fast = ctx.sph_align_pairs(A, B, J, 0.5, invert=True, want_grid=True)
    fast_ng = ctx.sph_align_pairs(A, B, J, 0.5, invert=True)
    ctx.set_option("force_generic", 1)
    try:
        gen = ctx.sph_align_pairs(A, B, J, 0.5, invert=True, want_grid=True)
    finally:
        ctx.set_option("force_generic", 0)
    assert np.array_equal(fast[0], gen[0]) and np.array_equal(fast[0], fast_ng[0])
    assert np.allclose(fast[1], gen[1], rtol=1e-12) and np.array_equal(fast[1], fast_ng[1])
    assert np.allclose(fast[2], gen[2], atol=1e-6)
    assert rel(fast[3], gen[3]) < 1e-12


@pytest.mark.parametrize("J,invert", [(15, True), (15, False), (13, True), (7, True), (11, True), (3, True), (14, True), (8, False)])
def test_isoft_kernel_variants_agree(ctx, J, invert):
    """sph_isoft4_kernel (stages A -> B chained in registers; four planes per CTA for odd Jmax, two for even) against
    sph_isoft3_kernel (stage A -> shared memory -> stage B) and, for odd Jmax, its own two-plane form: same arg-max
    and interpolated maximum, grids to rounding."""
    rng = np.random.default_rng(100 + J)
    N = 38
    A = rng.normal(size=(5, N, 3))
    B = rng.normal(size=(5, N, 3))
    A -= A.mean(1, keepdims=True)
    B -= B.mean(1, keepdims=True)
    ctx.set_perm([np.arange(N)], N)
    new = ctx.sph_align_pairs(A, B, J, 0.45, invert=invert, want_grid=True)
    res = []
    for variant in (3, 42, 4, 5):   # 4: sph_isoft4_kernel, 5: sph_isoft5_kernel (FFT form; Jmax = 15, else the default)
        ctx.set_option("sph_isoft_variant", variant)
        try:
            res.append(ctx.sph_align_pairs(A, B, J, 0.45, invert=invert, want_grid=True))
        finally:
            ctx.set_option("sph_isoft_variant", 0)
    for old in res:
        assert np.array_equal(new[0], old[0])
        assert np.allclose(new[1], old[1], rtol=1e-12)
        assert np.allclose(new[2], old[2], atol=1e-7)
        assert rel(new[3], old[3]) < 1e-12


@pytest.mark.parametrize("N,J,groups", [(38, 15, None), (13, 7, None), (55, 20, [30, 25]), (8, 3, None), (63, 15, None),
                                        (24, 31, [10, 14]), (3, 2, None)])
def test_direct_coefficient_kernels_agree(ctx, N, J, groups):
    """sph_prep2 / sph_bessel2 / sph_direct2 (operands in the DMMA fragment layouts: swizzled 8-column groups for
    every tile count mod 4, row-permuted second structure, mbarrier ring, chained second product) against
    sph_prep / sph_bessel / sph_direct_mma (option sph_direct_ring = -1) and against the oracle."""
    rng = np.random.default_rng(1000 + N + J)
    P = 6
    A = rng.normal(size=(P, N, 3)) * 1.1
    B = rng.normal(size=(P, N, 3)) * 1.1
    A -= A.mean(1, keepdims=True)
    B -= B.mean(1, keepdims=True)
    perm = None
    if groups:
        o = np.cumsum([0] + groups)
        perm = [np.arange(o[i], o[i + 1]) for i in range(len(groups))]
    ctx.set_perm(perm if perm else [np.arange(N)], N)
    new, st = ctx.sph_coeffs_direct(A, B, J, 0.45)
    ctx.set_option("sph_direct_ring", -1)
    try:
        old, _ = ctx.sph_coeffs_direct(A, B, J, 0.45)
    finally:
        ctx.set_option("sph_direct_ring", 0)
    assert (st == 0).all()
    for p in range(P):
        assert rel(new[p], old[p]) < 1e-13
    assert rel(new[0], oracle.sph_coeffs_direct(A[0], B[0], J, 0.45, perm)) < 1e-12
    # two ring slots (three CTAs per SM) give the same numbers bit for bit
    ctx.set_option("sph_direct_ring", 2)
    try:
        two, _ = ctx.sph_coeffs_direct(A, B, J, 0.45)
    finally:
        ctx.set_option("sph_direct_ring", 0)
    assert np.array_equal(new, two)
    ctx.set_perm([np.arange(N)], N)


@pytest.mark.parametrize("N,J", [(10, 33), (12, 40), (9, 63)])
def test_large_bandwidth_isoft(ctx, N, J):
    """Jmax > 32 (grids up to 128^3, BASELINE.json configs[3]): sph_isoft_big_kernel with the
    plane-major Wigner table vs the oracle (DSOFT.f90:265-329 restatement)."""
    rng = np.random.default_rng(N * 100 + J)
    P = 2
    A = rng.normal(size=(P, N, 3))
    B = rng.normal(size=(P, N, 3))
    A -= A.mean(1, keepdims=True)
    B -= B.mean(1, keepdims=True)
    ctx.set_perm([np.arange(N)], N)
    bi, bv, fr, grid, st = ctx.sph_align_pairs(A, B, J, 0.45, invert=True, want_grid=True)
    obi, obv, ofr, ogrid, _ = oracle.sph_align_pairs(A, B, J, 0.45, True, None, want_grid=True)
    for p in range(P):
        for o_ in range(2):
            assert rel(grid[p, o_], ogrid[p, o_]) < GRID_RTOL
    assert np.array_equal(bi, obi)
    assert np.allclose(bv, obv, rtol=1e-11)
    assert np.allclose(fr, ofr, atol=1e-6)
    bi2, bv2, fr2, _, _ = ctx.sph_align_pairs(A, B, J, 0.45, invert=True)
    assert np.array_equal(bi, bi2) and np.array_equal(bv, bv2) and np.array_equal(fr, fr2)
    if J == 40:  # Wigner table export from the plane-major layout
        assert rel(ctx.sph_wigner_table(J), oracle.wigner_table(J + 1)) < 1e-11


@pytest.mark.parametrize("N,J,groups,force", [(24, 9, [14, 10], True), (38, 15, None, True),
                                              (131, 12, [70, 61], False), (200, 21, None, False)])
def test_direct_coeffs_tensor_core_gemm(ctx, N, J, groups, force):
    """Large-cluster form of the direct coefficients (sph_direct_gemm_kernel, DMMA) vs the oracle
    (fastclusters.f90:868-916 restatement) and vs the small-cluster kernel."""
    rng = np.random.default_rng(N + 7 * J)
    P = 3
    A = rng.normal(size=(P, N, 3)) * 1.5
    B = rng.normal(size=(P, N, 3)) * 1.5
    A -= A.mean(1, keepdims=True)
    B -= B.mean(1, keepdims=True)
    perm = None
    if groups:
        o = np.cumsum([0] + groups)
        perm = [np.arange(o[i], o[i + 1]) for i in range(len(groups))]
    ctx.set_perm(perm if perm else [np.arange(N)], N)
    try:
        ctx.set_option("direct_gemm_min_atoms", 1 if force else 64)
        I, _ = ctx.sph_coeffs_direct(A, B, J, 0.5)
        res = ctx.sph_align_pairs(A, B, J, 0.5, invert=True)
        ctx.set_option("direct_gemm_min_atoms", 1 << 30)
        I0, _ = ctx.sph_coeffs_direct(A, B, J, 0.5)
        res0 = ctx.sph_align_pairs(A, B, J, 0.5, invert=True)
    finally:
        ctx.set_option("direct_gemm_min_atoms", 64)
        ctx.set_perm([np.arange(N)], N)
    for p in range(P):
        assert rel(I[p], oracle.sph_coeffs_direct(A[p], B[p], J, 0.5, perm)) < 1e-12
    assert rel(I, I0) < 1e-12
    assert np.array_equal(res[0], res0[0])
    assert np.allclose(res[1], res0[1], rtol=1e-11)


@pytest.mark.parametrize("N,J", [(400, 31), (1000, 63)])
def test_large_cluster_rotation_recovery(ctx, N, J):
    """BASELINE.json configs[3] (LJ1000-size clusters at high Jmax; also at reduced size): lattice blob,
    rotated + permuted + jittered partner; the alignment recovers the transformation (distance ~ noise
    level).  Size-independent property at the full configuration size."""
    from fastoverlap_b200 import SphericalAlign
    rng = np.random.default_rng(1000)
    g = np.arange(-8, 9) * 1.12
    pts = np.array(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1).T
    pts = pts[np.argsort(np.linalg.norm(pts, axis=1), kind="stable")[:N]]
    A = pts + rng.normal(scale=0.03, size=pts.shape)
    A -= A.mean(0)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    R = np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                  [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                  [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])
    noise = rng.normal(scale=0.02, size=A.shape)
    B = (A + noise).dot(R.T)[rng.permutation(N)]
    B -= B.mean(0)
    sa = SphericalAlign(0.37, J, ctx=ctx)
    dist = sa(A, B)[0]
    assert dist < 2.0 * np.linalg.norm(noise), dist


def test_sphharm_stage_method(ctx):
    """a1 as a stage method: sphHarm(theta, phi) (sphericalAlignment.py:57-65) on the device vs scipy, in the
    reference's [l, m wrapped, atom] layout; r from fo_sph_ylm; SOFT.calcWignerMatrices; _align alias."""
    from scipy.special import sph_harm_y
    from fastoverlap_b200 import SphericalAlign, SOFT
    rng = np.random.default_rng(57)
    for L in (4, 15, 31):
        sa = SphericalAlign(0.3, L, ctx=ctx)
        pos = rng.normal(size=(23, 3))
        r = np.linalg.norm(pos, axis=1)
        theta, phi = np.arccos(pos[:, 2] / r), np.arctan2(pos[:, 1], pos[:, 0])
        Y = sa.sphHarm(theta, phi)
        assert Y.shape == (L + 1, 2 * L + 1, 23)
        ref = np.zeros_like(Y)
        for l in range(L + 1):
            for m in range(-l, l + 1):
                ref[l, m] = sph_harm_y(l, m, theta, phi)
        assert np.abs(Y - ref).max() < 2e-13, L
        Y2, r2, st = ctx.sph_ylm(pos, L)
        assert np.abs(Y2[0] - ref).max() < 2e-13 and np.abs(r2[0] - r).max() < 1e-15 and st[0] == 0
    s = golden("soft_tables.npz")
    assert np.abs(SOFT(8, ctx=ctx).calcWignerMatrices() - s["Ds_8"]).max() < 1e-12
    g = golden("spherical_lj38.npz")
    sa = SphericalAlign(0.3, 15, ctx=ctx)
    assert abs(sa._align(g["pos1"], g["pos2"])[0] - float(g["J15_dist"])) < DIST_ATOL
