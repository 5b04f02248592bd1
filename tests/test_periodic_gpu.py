"""GPU parity of the periodic hot path: CUDA (through the C ABI) vs golden vectors from the
unmodified reference and vs the C oracle on seeded inputs.

Tolerances (BASELINE.json north_star): overlap grids within 1e-10 relative (to the grid max),
identical best grid index, final distance after the same host permutation step within 1e-8.
"""
import numpy as np
import pytest

import oracle
from conftest import golden, groups_from

pytestmark = pytest.mark.gpu

GRID_RTOL = 1e-10
DIST_ATOL = 1e-8


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max()


def test_blj256_structure_factors(ctx):
    from fastoverlap_b200 import PeriodicAlign
    g = golden("periodic_blj256.npz")
    al = PeriodicAlign(256, g["box"], [np.arange(204), np.arange(204, 256)], ctx=ctx)
    assert al.n == int(g["n"]) and al.fshape == (int(g["F"]),) * 3
    assert abs(al.scale - float(g["scale"])) < 1e-15
    C1 = al.calcFourierCoeff(g["pos1"])
    assert C1.shape == g["C1"].shape
    assert rel(C1, g["C1"]) < 1e-13


def test_blj256_grid_argmax_distance(ctx):
    from fastoverlap_b200 import PeriodicAlign
    g = golden("periodic_blj256.npz")
    al = PeriodicAlign(256, g["box"], [np.arange(204), np.arange(204, 256)], ctx=ctx)
    dist, X1, X2, perm, disp = al(g["pos1"], g["pos2"])
    assert rel(al.fabs, g["fabs"]) < GRID_RTOL
    assert tuple(np.unravel_index(al.fabs.argmax(), al.fabs.shape)) == (10, 38, 32)
    assert tuple(al._best_idx) == (10, 38, 32)
    assert abs(al._best_val - 55419.12238387397) < 1e-10 * 55419.0
    assert np.allclose(al._frac_idx, g["findmax"], atol=1e-8)
    assert abs(dist - 1.5590835031549872) < DIST_ATOL
    assert abs(dist - float(g["dist"])) < DIST_ATOL
    assert np.array_equal(np.asarray(perm), g["perm"])
    assert np.allclose(disp, g["disp"], atol=1e-8)
    # precomputed-coefficient hook (examples/alignPeriodic.py:35-42 "quickAlign")
    c1, c2 = al.calcFourierCoeff(g["pos1"]), al.calcFourierCoeff(g["pos2"])
    d2 = al.align(g["pos1"], g["pos2"], [c1, c2])[0]
    assert abs(d2 - dist) < 1e-12
    # and with the reference's own coefficients
    d3 = al.align(g["pos1"], g["pos2"], [g["C1"], g["C2"]])[0]
    assert abs(d3 - float(g["dist"])) < DIST_ATOL


def test_synthetic_cases_vs_reference_golden(ctx):
    from fastoverlap_b200 import PeriodicAlign
    g = golden("periodic_synth.npz")
    for i in range(int(g["ncases"])):
        k = "c%d_" % i
        perm = groups_from(g[k + "groups"], g[k + "gsizes"])
        N = len(g[k + "pos1"])
        al = PeriodicAlign(N, g[k + "box"], perm, scale=float(g[k + "scale"]), n=int(g[k + "n"]), ctx=ctx)
        assert al.fshape[0] == int(g[k + "F"])
        dist, X1, X2, p, disp = al(g[k + "pos1"], g[k + "pos2"])
        assert rel(al.fabs, g[k + "fabs"]) < GRID_RTOL, i
        assert tuple(al._best_idx) == tuple(g[k + "argmax"]), i
        assert np.allclose(al._frac_idx, g[k + "findmax"], atol=1e-8), i
        assert abs(dist - float(g[k + "dist"])) < DIST_ATOL, i
        assert rel(al.calcFourierCoeff(g[k + "pos1"]), g[k + "C1"]) < 1e-13, i


def _random_pairs(rng, P, N, box, jitter=0.05):
    pos1 = rng.uniform(-0.5, 0.5, size=(P, N, 3)) * box
    shift = rng.uniform(0, 1, size=(P, 1, 3)) * box
    pos2 = pos1 + shift + rng.normal(scale=jitter, size=(P, N, 3))
    return pos1, pos2, shift[:, 0, :]


@pytest.mark.parametrize("N,n,groups", [(17, 3, None), (40, 6, [23, 17]), (64, 9, [50, 14]),
                                         (33, 11, None), (5, 1, None),
                                         # per_sf3: several atom tiles per group, an empty and a tiny group
                                         (230, 9, [120, 0, 3, 107]),
                                         (36, 16, [20, 16]), (24, 32, None)])  # fine k-grids: F = 72, 135
def test_batch_vs_oracle(ctx, N, n, groups):
    """Seeded random batches: every pair's grid, arg-max and interpolated maximum vs the oracle."""
    from fastoverlap_b200 import PeriodicAlign
    rng = np.random.default_rng(1000 + N)
    box = np.array([4.0, 4.5, 5.1])
    perm = None
    if groups:
        o = np.cumsum([0] + groups)
        perm = [np.arange(o[i], o[i + 1]) for i in range(len(groups))]
    P = 7 if n <= 11 else 2
    pos1, pos2, _ = _random_pairs(rng, P, N, box)
    al = PeriodicAlign(N, box, perm, n=n, ctx=ctx)
    F = al.fshape[0]
    p = al._params()
    bi, bv, fr, grids, st = ctx.per_align_pairs(p, pos1, pos2, want_grid=True)
    obi, obv, ofr, ogrids, _ = oracle.per_align_pairs(pos1, pos2, box, n, F, al.scale, perm,
                                                      want_grid=True)
    assert np.all(st == 0)
    for i in range(P):
        assert rel(grids[i], ogrids[i]) < GRID_RTOL
    assert np.array_equal(bi, obi)
    assert np.allclose(bv, obv, rtol=1e-12)
    assert np.allclose(fr, ofr, atol=1e-7)
    # results without the grid output are identical (the grid is only a side output)
    bi2, bv2, fr2, _, _ = ctx.per_align_pairs(p, pos1, pos2)
    assert np.array_equal(bi, bi2) and np.array_equal(bv, bv2) and np.array_equal(fr, fr2)


def test_bench_configuration_vs_oracle(ctx):
    """The bench workload itself (bench.py Blj256: 256 atoms [204, 52], box 5.975206329, n = 9, F = 40, partner
    = translated + N(0, 0.05^2) + permuted within species): P = 64 pairs in one call -- arg-max, peak value and
    interpolated maximum of every pair against the oracle, the full 40^3 grid on a subsample."""
    from fastoverlap_b200 import PeriodicAlign
    import bench
    wl = bench.Blj256()
    A, B, shift = wl.make(64, 5)
    al = PeriodicAlign(256, wl.box, wl.perm, ctx=ctx)
    p = al._params()
    bi, bv, fr, _, st = ctx.per_align_pairs(p, A, B)
    obi, obv, ofr, _, _ = oracle.per_align_pairs(A, B, wl.box, 9, 40, al.scale, wl.perm, nthreads=0)
    assert np.all(st == 0)
    assert np.array_equal(bi, obi)
    assert np.allclose(bv, obv, rtol=1e-12)
    assert np.allclose(fr, ofr, atol=1e-7)
    sub = [0, 17, 42, 63]
    grids = ctx.per_align_pairs(p, A[sub], B[sub], want_grid=True)[3]
    ogrids = oracle.per_align_pairs(A[sub], B[sub], wl.box, 9, 40, al.scale, wl.perm, want_grid=True)[3]
    for i in range(len(sub)):
        assert rel(grids[i], ogrids[i]) < GRID_RTOL
    d = fr * wl.box / 40 - shift
    d -= np.round(d / wl.box) * wl.box
    assert np.abs(d).max() < wl.box[0] / 40


@pytest.mark.parametrize("ngroups", [1, 2, 3])
def test_fused_pairs_path_matches_bank_path(ctx, ngroups):
    """per_sfx_kernel (structure factors of both structures + cross-spectrum in one kernel, S_A stashed in tensor
    memory, groups accumulated into the image by red.add) against the bank path (per_sf3 x 2 + per_cross6) at the
    bench k-grid: same arg-max, overlap values and grids to rounding, for one, two and three permutation groups
    (an empty leading group included) and odd pair counts."""
    from fastoverlap_b200 import PeriodicAlign
    import bench
    wl = bench.Blj256()
    A, B, _ = wl.make(37, 11)
    perm = {1: [np.arange(256)], 2: wl.perm, 3: [np.arange(0), np.arange(120), np.arange(120, 256)]}[ngroups]
    al = PeriodicAlign(256, wl.box, perm, ctx=ctx)
    p = al._params()
    fused = ctx.per_align_pairs(p, A, B, want_grid=True)
    ctx.set_option("per_pairs_fused", 0)
    try:
        bank = ctx.per_align_pairs(p, A, B, want_grid=True)
    finally:
        ctx.set_option("per_pairs_fused", 1)
    assert np.array_equal(fused[0], bank[0])
    assert np.allclose(fused[1], bank[1], rtol=1e-13)
    assert np.allclose(fused[2], bank[2], atol=1e-9)
    assert rel(fused[3], bank[3]) < 1e-13


def test_translation_recovery_full_size(ctx):
    """Size-independent property at BASELINE size (N=256, n=9, F=40): a pure translation plus a
    permutation within species is recovered; the overlap peak sits at the translation."""
    from fastoverlap_b200 import PeriodicAlign
    g = golden("periodic_blj256.npz")
    rng = np.random.default_rng(256)
    box = g["box"]
    perm = [np.arange(204), np.arange(204, 256)]
    P = 300  # > one chunk of SMs, exercises the persistent loop
    base = g["pos1"]
    shift = rng.uniform(0, 1, size=(P, 3)) * box
    pos2 = np.empty((P, 256, 3))
    for i in range(P):
        order = np.concatenate([rng.permutation(204), 204 + rng.permutation(52)])
        pos2[i] = (base + shift[i])[order]
    pos1 = np.broadcast_to(base, pos2.shape).copy()
    al = PeriodicAlign(256, box, perm, ctx=ctx)
    disps, bi, bv = al.findDisps_batch(pos1, pos2)
    err = disps - shift
    err -= np.round(err / box) * box
    assert np.abs(err).max() < 0.02 * box[0] / 40 * 40 / 10  # well inside one grid cell
    # the peak height is that of the self-overlap up to grid sampling of the peak
    assert np.ptp(bv) / bv.mean() < 2e-2
    dists, out_disp, perms = al.align_batch(pos1[:5], pos2[:5])
    assert np.all(dists < 1e-6)


def test_all_vs_all_bank_matches_pairs(ctx):
    from fastoverlap_b200 import PeriodicAlign
    rng = np.random.default_rng(5)
    N, box = 24, np.array([3.3, 3.3, 3.3])
    coords = rng.uniform(-0.5, 0.5, size=(5, N, 3)) * box
    al = PeriodicAlign(N, box, ctx=ctx)
    p = al._params()
    bank = ctx.per_bank_create(p, coords)
    pairs = np.array([(i, j) for i in range(5) for j in range(5)])
    bi, bv, fr, _, _ = ctx.per_align_bank(p, bank, pairs)
    bi2, bv2, fr2, _, _ = ctx.per_align_pairs(p, coords[pairs[:, 0]], coords[pairs[:, 1]])
    assert np.array_equal(bi, bi2) and np.array_equal(bv, bv2) and np.array_equal(fr, fr2)
    dists = al.alignGroup(coords)
    assert np.allclose(np.diag(dists), 0, atol=1e-7)
    assert np.allclose(dists, dists.T, atol=1e-6)


def test_edge_cases(ctx):
    from fastoverlap_b200 import PeriodicAlign, FastOverlapError
    box = np.array([3.0, 3.0, 3.0])
    al = PeriodicAlign(4, box, ctx=ctx)
    p = al._params()
    # empty batch
    bi, bv, fr, _, st = ctx.per_align_pairs(p, np.zeros((0, 4, 3)), np.zeros((0, 4, 3)))
    assert bi.shape == (0, 3)
    # NaN coordinate is flagged per pair, the other pair is unaffected
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, size=(2, 4, 3))
    b = a + 0.3
    a[1, 2, 1] = np.nan
    bi, bv, fr, _, st = ctx.per_align_pairs(p, a, b)
    assert st[0] == 0 and st[1] != 0
    # identical structures: maximum at zero displacement
    bi, bv, fr, _, st = ctx.per_align_pairs(p, a[:1], a[:1])
    assert tuple(bi[0]) == (0, 0, 0)
    # invalid parameters raise, never abort
    bad = ctx.per_params(4, box, 0, 10, 0.3)
    with pytest.raises(FastOverlapError):
        ctx.per_align_pairs(bad, a[:1], a[:1])
    # a permutation group may be empty
    al2 = PeriodicAlign(4, box, [np.arange(4), np.array([], int)], ctx=ctx)
    r = ctx.per_align_pairs(al2._params(), a[:1], b[:1])
    r0 = ctx.per_align_pairs(p, a[:1], b[:1])
    ctx.set_perm([np.arange(4)], 4)
    assert np.array_equal(r[0], r0[0])


def test_multigpu_thread_sharding_identical(ctx):
    """One Context + one host thread per GPU (here: two contexts on the available device):
    per-pair results are bit-identical to the single-context run for any shard count."""
    import torch
    from fastoverlap_b200 import PeriodicAlign
    from fastoverlap_b200.batch import MultiGPU
    rng = np.random.default_rng(8)
    N, box = 20, np.array([3.0, 3.2, 3.4])
    pos1, pos2, _ = _random_pairs(rng, 11, N, box)
    al = PeriodicAlign(N, box, ctx=ctx)
    ref = ctx.per_align_pairs(al._params(), pos1, pos2)
    ndev = torch.cuda.device_count()
    mg = MultiGPU([i % ndev for i in range(3)])

    def fn(c, a, b):
        c.set_perm([np.arange(N)], N)
        return c.per_align_pairs(al._params(), a, b)[:3]
    out = mg.map_pairs(fn, pos1, pos2)
    assert all(np.array_equal(o, r) for o, r in zip(out, ref[:3]))


@pytest.mark.parametrize("N,n", [(24, 3), (30, 6), (64, 9), (20, 4)])
def test_fast_and_generic_kernels_agree(ctx, N, n):
    """The tensor-core fast paths (per_sf2 / per_xf4) and the any-size kernels (per_sf / per_xf) are
    two implementations of the same mathematics: same arg-max, values within rounding."""
    from fastoverlap_b200 import PeriodicAlign
    rng = np.random.default_rng(N * 100 + n)
    box = np.array([4.1, 4.4, 4.9])
    pos1, pos2, _ = _random_pairs(rng, 6, N, box)
    al = PeriodicAlign(N, box, n=n, ctx=ctx)
    p = al._params()
    fast = ctx.per_align_pairs(p, pos1, pos2, want_grid=True)
    ctx.set_option("force_generic", 1)
    try:
        gen = ctx.per_align_pairs(p, pos1, pos2, want_grid=True)
    finally:
        ctx.set_option("force_generic", 0)
    assert np.array_equal(fast[0], gen[0])
    assert np.allclose(fast[1], gen[1], rtol=1e-12)
    assert np.allclose(fast[2], gen[2], atol=1e-7)
    assert rel(fast[3], gen[3]) < 1e-12


def test_oh_cell_symmetries(ctx):
    """O_h branch (ALIGN1 with OHCELLT, fastbulk.f90:458-480; broken in the reference): a cubic
    cell, the partner is an octahedral image + translation + noise + permutation of the first structure.
    The plain alignment cannot match it, the 48-operation search recovers the noise-level distance and
    the operation."""
    from fastoverlap_b200 import PeriodicAlign, PeriodicAlignFortran
    from fastoverlap_b200.utils import oh_operations
    rng = np.random.default_rng(48)
    N, box = 40, np.array([5.0, 5.0, 5.0])
    groups = [np.arange(30), np.arange(30, 40)]
    ops = oh_operations()
    assert ops.shape == (48, 3, 3) and len({o.tobytes() for o in ops}) == 48
    al = PeriodicAlign(N, box, groups, ctx=ctx)
    pos1 = rng.uniform(-0.5, 0.5, size=(N, 3)) * box
    noise = 0.01
    for k in (7, 19, 41):  # two proper rotations / one improper operation (none the identity)
        R0 = ops[k]
        pos2 = pos1.dot(R0.T) + rng.uniform(0, 1, size=3) * box + rng.normal(scale=noise, size=(N, 3))
        order = np.arange(N)
        for g in groups:
            order[g] = rng.permutation(g)
        pos2 = pos2[order]
        plain = al(pos1, pos2)[0]
        dist, X1, X2, perm, disp, R = al.align_oh(pos1, pos2)
        assert dist < 3 * noise * np.sqrt(3 * N), (k, dist)
        assert plain > 10 * dist, (k, plain, dist)
        assert np.array_equal(R, R0.T), k
        d = X1 - X2
        d -= np.round(d / box) * box
        assert abs(np.linalg.norm(d) - dist) < 1e-9
        fw = PeriodicAlignFortran(N, box, perm=groups)
        assert abs(fw.align(pos1, pos2, ohcell=True)[0] - dist) < 1e-9
    with pytest.raises(ValueError):
        PeriodicAlign(N, [5.0, 5.0, 6.0], groups, ctx=ctx).align_oh(pos1, pos1)


@pytest.mark.parametrize("N,n,groups", [(40, None, [np.arange(30), np.arange(30, 40)]), (256, 9, None)])
def test_oh_index_permutation_matches_transformed_coordinates(ctx, N, n, groups):
    """fo_per_align_bank_ops (OHTRANSFORMCOEFFS, fastbulk.f90:863-1380): the structure factors of the 48 images
    R B as index permutations of B's bank entry, against the hot path on the explicitly transformed coordinates:
    same arg-max, same overlap value and interpolated maximum; a non-cubic box and a bad code are refused."""
    from fastoverlap_b200 import PeriodicAlign, _lib
    from fastoverlap_b200.utils import oh_operations
    rng = np.random.default_rng(4848 + N)
    box = np.array([5.3, 5.3, 5.3])
    if groups is None:
        groups = [np.arange(204), np.arange(204, 256)]
    al = PeriodicAlign(N, box, groups, ctx=ctx) if n is None else PeriodicAlign(N, box, groups, n=n, ctx=ctx)
    p = al._params()
    pos1 = rng.uniform(-0.5, 0.5, size=(N, 3)) * box
    pos2 = pos1.dot(oh_operations()[29].T) + 0.37 + rng.normal(scale=0.02, size=(N, 3))
    ops = oh_operations()
    X2s = np.einsum("oij,aj->oai", ops, pos2)
    ref = ctx.per_align_pairs(p, np.broadcast_to(pos1, X2s.shape).copy(), X2s)
    bank = ctx.per_bank_create(p, np.stack([pos1, pos2]))
    got = ctx.per_align_bank_ops(p, bank, np.tile([0, 1], (48, 1)), ops)
    assert np.array_equal(got[0], ref[0])
    assert np.allclose(got[1], ref[1], rtol=1e-11)
    assert np.allclose(got[2], ref[2], atol=1e-6)
    ident = ctx.per_align_bank(p, bank, [[0, 1]])
    assert np.array_equal(ident[0][0], got[0][0]) and ident[1][0] == got[1][0]
    bad = ops.copy()
    bad[3] = 0.5
    with pytest.raises(ValueError):
        ctx.per_align_bank_ops(p, bank, np.tile([0, 1], (48, 1)), bad)
    bank.close()
    al2 = PeriodicAlign(N, [5.3, 5.3, 6.0], groups, ctx=ctx)
    bank2 = ctx.per_bank_create(al2._params(), np.stack([pos1, pos2]))
    with pytest.raises(_lib.FastOverlapError):
        ctx.per_align_bank_ops(al2._params(), bank2, [[0, 1]], ops[:1])
    bank2.close()
