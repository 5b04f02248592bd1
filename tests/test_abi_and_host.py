"""CPU: the C-ABI library loads and exports every symbol include/fastoverlap_b200.h declares;
host-side logic (LAP, Kearsley, findMax, peaks, refine) against the reference's golden values.
No compute call is made on the library (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fastoverlap_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import fastoverlap_b200 as fob
    from fastoverlap_b200 import _lib
    path = fob.library_path()
    if not os.path.exists(path):
        from fastoverlap_b200 import build
        build.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    # the ctypes signature table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_no_gpu_fails_loudly():
    import fastoverlap_b200 as fob
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fob.FastOverlapError):
        fob.Context(0)
    with pytest.raises(fob.FastOverlapError):
        fob.PeriodicAlign(4, [1.0, 1.0, 1.0]).calcFourierCoeff(np.zeros((4, 3)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fastoverlap_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dp, f)
                assert "/root/reference" not in txt, os.path.join(dp, f)


def test_next_fast_len():
    from fastoverlap_b200.utils import _next_fast_len
    tab = golden("next_fast_len.npz")["table"]
    assert all(_next_fast_len(i) == tab[i] for i in range(len(tab)))
    lib = ctypes.CDLL(__import__("fastoverlap_b200").library_path())
    lib.fo_next_fast_len.restype = ctypes.c_int64
    lib.fo_next_fast_len.argtypes = [ctypes.c_int64]
    assert all(lib.fo_next_fast_len(i) == tab[i] for i in range(len(tab)))
    s, n, F = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    box = (ctypes.c_double * 3)(5.975206329, 5.975206329, 5.975206329)
    lib.fo_per_defaults.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]
    assert lib.fo_per_defaults(256, box, ctypes.byref(s), ctypes.byref(n), ctypes.byref(F)) == 0
    g = golden("periodic_blj256.npz")
    assert n.value == int(g["n"]) and F.value == int(g["F"]) and abs(s.value - float(g["scale"])) < 1e-15


def test_findmax_host():
    from fastoverlap_b200.utils import findMax
    g = golden("periodic_blj256.npz")
    assert np.allclose(findMax(g["fabs"]), g["findmax"], atol=1e-12)
    s = golden("spherical_lj38.npz")
    assert np.allclose(findMax(s["J15_grid"]), s["J15_findmax"].real, atol=1e-12)


def test_periodic_refine_host_matches_reference():
    """BasePeriodicAlignment.refine with the reference's displacement reproduces the documented
    1.559 (periodicAlignment.py:609) -- host LAP uses the pele cost-matrix orientation (Q9)."""
    from fastoverlap_b200.periodic import PeriodicAlign
    g = golden("periodic_blj256.npz")
    al = PeriodicAlign.__new__(PeriodicAlign)
    al.Natoms, al.boxvec, al.dim = 256, g["box"], 3
    al.perm = [np.arange(204), np.arange(204, 256)]
    dist, X1, X2, perm, disp = al.refine(g["pos1"].copy(), g["pos2"].copy(), g["disp0"][None, :])
    assert abs(dist - 1.5590835031549872) < 1e-10
    assert np.array_equal(perm, g["perm"])
    assert np.allclose(disp, g["disp"], atol=1e-10)


def test_spherical_refine_host_matches_reference():
    from fastoverlap_b200.spherical import SphericalAlign
    from fastoverlap_b200.utils import indtoEuler
    g = golden("spherical_lj38.npz")
    sa = SphericalAlign.__new__(SphericalAlign)
    sa.perm, sa.scale, sa.Jmax = None, 0.3, 15
    X1, X2 = sa.COM_shift(g["pos1"], g["pos2"])
    R = indtoEuler(g["J15_findmax_inv"].real, 32)
    assert abs(sa.refine(X1, -X2, R)[0] - 1.4767670631638872) < 1e-10
    R = indtoEuler(g["J15_findmax"].real, 32)
    assert abs(sa.refine(X1, X2, R)[0] - float(g["J15_dist_normal_only"])) < 1e-10


def test_kearsley_and_lap():
    from fastoverlap_b200.utils import findrotation, find_best_permutation, EulerM
    rng = np.random.default_rng(0)
    X = rng.normal(size=(30, 3))
    X -= X.mean(0)
    R = EulerM(0.3, 1.1, -2.0)
    perm = rng.permutation(30)
    Y = X.dot(R)[perm]
    _, p = find_best_permutation(X.dot(R), Y)
    assert np.array_equal(np.asarray(Y)[p], X.dot(R))
    d, M = findrotation(X, X.dot(R))
    assert d < 1e-7 and np.allclose(X.dot(R).dot(M.T), X, atol=1e-7)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 38, 63, 64, 65, 100, 204])
def test_lap_vector_kernels_all_sizes(n):
    """The per-ISA cost / row-scan kernels of the assignment (padded column arrays, in-register column pick) against
    scipy.optimize.linear_sum_assignment on unrelated point sets (many free rows after the column reduction: long
    augmenting paths), free and periodic costs, every size around the vector widths; and on integer coordinates
    (exact ties: any optimal assignment, equal cost)."""
    from scipy.optimize import linear_sum_assignment
    from fastoverlap_b200 import _lib
    rng = np.random.default_rng(4000 + n)
    box = np.array([3.0, 3.5, 4.0])
    for trial in range(3):
        # every instruction set this CPU has (2-wide, AVX2, AVX-512), the widest one last = the default again
        isa = _lib.host_lap_isa(trial + 1)
        if isa < 0:
            isa = _lib.host_lap_isa(0)
        X = rng.uniform(-2, 2, size=(n, 3))
        Y = rng.uniform(-2, 2, size=(n, 3))
        c = ((X[:, None, :] - Y[None, :, :]) ** 2).sum(2)
        r, cc = linear_sum_assignment(c)
        assert np.array_equal(_lib.host_best_permutation(X, Y), cc)
        d = X[:, None, :] - Y[None, :, :]
        d -= np.rint(d / box) * box
        cp = np.sqrt((d ** 2).sum(2))
        r, cc = linear_sum_assignment(cp)
        assert np.array_equal(_lib.host_best_permutation(X, Y, None, box), cc)
    Xi = rng.integers(0, 3, size=(n, 3)).astype(float)
    Yi = rng.integers(0, 3, size=(n, 3)).astype(float)
    c = ((Xi[:, None, :] - Yi[None, :, :]) ** 2).sum(2)
    r, cc = linear_sum_assignment(c)
    p = _lib.host_best_permutation(Xi, Yi)
    assert sorted(p) == list(range(n)) and c[np.arange(n), p].sum() == c[r, cc].sum()
    assert _lib.host_lap_isa(0) >= 1


def test_findpeaks_recovers_planted_gaussians():
    """The scipy restatement of the reference's findPeaks (oracle/peaks_host.py), the checker of the device kernel."""
    from peaks_host import findPeaks
    n = 24
    x = np.indices((n, n, n)).astype(float)
    f = np.zeros((n, n, n))
    for c, a in (((5.3, 10.1, 17.6), 3.0), ((15.2, 4.4, 8.9), 2.0)):
        d2 = sum((np.minimum(abs(x[i] - c[i]), n - abs(x[i] - c[i]))) ** 2 for i in range(3))
        f += a * np.exp(-d2 / 4.0)
    peaks, amp, _, _, _ = findPeaks(f, npeaks=2, width=2)
    assert np.allclose(peaks[0], (5.3, 10.1, 17.6), atol=0.05)
    assert np.allclose(peaks[1], (15.2, 4.4, 8.9), atol=0.05)


def test_native_periodic_refine_matches_reference_and_python():
    """fo_host_refine_periodic (C++ JV LAP + mean-displacement loop) on the reference's BLJ256 pair:
    the documented 1.559, the reference's permutation and displacement, and agreement with the
    Python refine on random pairs."""
    from fastoverlap_b200 import _lib
    from fastoverlap_b200.periodic import PeriodicAlign
    g = golden("periodic_blj256.npz")
    perm = [np.arange(204), np.arange(204, 256)]
    p = _lib.Context.per_params(256, g["box"], int(g["n"]), int(g["F"]), float(g["scale"]))
    dist, pm, disp = _lib.host_refine_periodic(p, perm, g["pos1"][None], g["pos2"][None], g["findmax"][None])
    assert abs(dist[0] - 1.5590835031549872) < 1e-9
    assert np.array_equal(pm[0], g["perm"])
    assert np.allclose(disp[0], g["disp"], atol=1e-9)
    rng = np.random.default_rng(4)
    N, box = 30, np.array([4.0, 4.5, 5.0])
    al = PeriodicAlign.__new__(PeriodicAlign)
    al.Natoms, al.boxvec, al.dim = N, box, 3
    al.perm = [np.arange(18), np.arange(18, 30)]
    A = rng.uniform(-0.5, 0.5, size=(6, N, 3)) * box
    shift = rng.uniform(0, 1, size=(6, 1, 3)) * box
    B = A + shift + rng.normal(scale=0.05, size=A.shape)
    for i in range(6):
        B[i] = B[i][np.concatenate([rng.permutation(18), 18 + rng.permutation(12)])]
    F = 24
    frac = (shift[:, 0, :] + rng.normal(scale=0.02, size=(6, 3))) / box * F
    pp = _lib.Context.per_params(N, box, 5, F, 0.3)
    dist, pm, disp = _lib.host_refine_periodic(pp, al.perm, A, B, frac, nthreads=2)
    for i in range(6):
        d, _, _, pyperm, pydisp = al.refine(A[i].copy(), B[i].copy(), (frac[i] * box / F)[None])
        assert abs(d - dist[i]) < 1e-9 and np.array_equal(pyperm, pm[i]) and np.allclose(pydisp, disp[i], atol=1e-9)


@pytest.mark.parametrize("jitter", [0.35, 0.12, 0.04, 0.0])
def test_native_periodic_refine_hard_assignments(jitter):
    """Group sizes that are not multiples of the 8- / 16-column blocks of the cost kernels, at four noise
    levels: large jitter (many rows left free by the column reduction, so the augmentation scans rows whose
    square roots are taken lazily; the single-precision screening fails and the double-precision matrix +
    LAP decides), intermediate (screening passes for some groups and solves, fails for others), small and
    none (screening proves the column minima optimal and the second solve of the loop is skipped by the
    stability margin; coordinates far outside the box).  Permutations identical to scipy's
    linear_sum_assignment on the min-image distance matrix through the reference's full loop
    (periodicAlignment.py:27-80), distances to 1e-12."""
    from scipy.optimize import linear_sum_assignment
    from fastoverlap_b200 import _lib
    rng = np.random.default_rng(11)
    N, box, F = 61, np.array([4.0, 4.5, 5.0]), 24
    groups = [np.arange(37), np.arange(37, 50), np.arange(50, 61)]
    P = 12
    A = rng.uniform(-0.5, 0.5, size=(P, N, 3)) * box
    if jitter < 0.1:
        # spread lattice-like points so that partners are unambiguous, then push them out of the cell
        A = (np.stack(np.unravel_index(rng.permutation(64)[:N], (4, 4, 4)), 1)[None] + 0.5 +
             rng.uniform(-0.15, 0.15, size=(P, N, 3))) / 4 * box + rng.integers(-3, 4, size=(P, N, 3)) * box
    shift = rng.uniform(0, 1, size=(P, 1, 3)) * box
    B = A + shift + rng.normal(scale=jitter, size=A.shape)
    for i in range(P):
        B[i] = B[i][np.concatenate([g[0] + rng.permutation(len(g)) for g in groups])]
    frac = shift[:, 0, :] / box * F
    pp = _lib.Context.per_params(N, box, 5, F, 0.3)
    _lib.host_refine_counters(reset=True)
    dist, pm, disp = _lib.host_refine_periodic(pp, groups, A, B, frac, nthreads=2)
    solved, screened, skipped = _lib.host_refine_counters()
    if jitter >= 0.3:
        assert screened < solved and skipped < P   # the LAP decides
    elif jitter >= 0.1:
        assert solved > 0 and screened > 0         # both paths
    else:
        assert (solved, screened, skipped) == (0, 3 * P, P)   # one screened pass per group, no second solve

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
        x, y, d = A[q], B[q], frac[q] * box / F
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
        assert np.array_equal(perm, pm[q])
        assert abs(ref - dist[q]) < 1e-12


def test_native_periodic_refine_randomized():
    """Random sizes, boxes (cubic and not), 1-3 groups, four kinds of structures (uniform, a blob far outside
    the cell, lattice-like, a thin sheet across the slab axis of the screening) and five noise levels, so that
    every path of fo_host_refine_periodic (slab-pruned single-precision screening, skipped confirming solve,
    double-precision LAP) meets awkward inputs: permutations identical to scipy's through the reference's
    loop (periodicAlignment.py:27-80)."""
    from scipy.optimize import linear_sum_assignment
    from fastoverlap_b200 import _lib
    rng = np.random.default_rng(5)
    tot = np.zeros(3, int)
    for trial in range(40):
        N = int(rng.integers(8, 260))
        box = rng.uniform(3.0, 9.0, 3) if trial % 3 else np.full(3, rng.uniform(3, 9))
        ng = int(rng.integers(1, 4))
        cuts = np.sort(rng.choice(np.arange(1, N), ng - 1, replace=False)) if ng > 1 else []
        groups = np.split(np.arange(N), cuts)
        P, F, mode = 3, 24, trial % 4
        if mode == 0:
            A = rng.uniform(-0.5, 0.5, size=(P, N, 3)) * box
        elif mode == 1:
            A = rng.normal(scale=0.15, size=(P, N, 3)) * box + 7 * box
        elif mode == 2:
            m = int(np.ceil(N ** (1 / 3)))
            A = (np.stack(np.unravel_index(rng.permutation(m ** 3)[:N], (m, m, m)), 1)[None] + 0.5 +
                 rng.uniform(-0.2, 0.2, size=(P, N, 3))) / m * box
        else:
            A = rng.uniform(-0.5, 0.5, size=(P, N, 3)) * box
            A[:, :, np.argmax(box)] *= 0.05
        jitter = [0.0, 0.01, 0.05, 0.1, 0.2][trial % 5] * (np.prod(box) / N) ** (1 / 3)
        shift = rng.uniform(0, 1, size=(P, 1, 3)) * box
        B = A + shift + rng.normal(scale=jitter, size=A.shape)
        for i in range(P):
            B[i] = B[i][np.concatenate([g[0] + rng.permutation(len(g)) for g in groups])]
        frac = shift[:, 0, :] / box * F + rng.normal(scale=0.02, size=(P, 3))
        pp = _lib.Context.per_params(N, box, 5, F, 0.3)
        _lib.host_refine_counters(reset=True)
        dist, pm, disp = _lib.host_refine_periodic(pp, groups, A, B, frac, nthreads=2)
        tot += np.array(_lib.host_refine_counters())

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
            x, y, d = A[q], B[q], frac[q] * box / F
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
            assert np.array_equal(perm, pm[q]), (trial, q, N, mode, jitter)
            assert abs(ref - dist[q]) < 1e-10 * max(1.0, ref)
    assert (tot > 20).all(), tot   # all three paths took part


def test_host_refine_periodic_subset_and_validation():
    """fo_host_refine_periodic_subset (the host stage of fo_per_align_pairs_full for the pairs its device
    screening flags) writes exactly the full call's results at the listed indices and leaves the rest alone;
    malformed permutation groups are rejected with FO_ERR_INVALID instead of being dereferenced."""
    import ctypes
    from fastoverlap_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(5)
    N, box, F, P = 45, np.array([4.0, 4.5, 5.0]), 24, 16
    groups = [np.arange(30), np.arange(30, 45)]
    A = rng.uniform(-0.5, 0.5, size=(P, N, 3)) * box
    shift = rng.uniform(0, 1, size=(P, 1, 3)) * box
    B = A + shift + rng.normal(scale=0.2, size=A.shape)
    for i in range(P):
        B[i] = B[i][np.concatenate([g[0] + rng.permutation(len(g)) for g in groups])]
    frac = shift[:, 0, :] / box * F
    pp = _lib.Context.per_params(N, box, 5, F, 0.3)
    dist, pm, disp = _lib.host_refine_periodic(pp, groups, A, B, frac, nthreads=2)
    off, idx, ng = _lib._group_arrays(groups, N)
    which = np.array([11, 3, 7], np.int64)
    d2 = np.full(P, -1.0)
    p2 = np.full((P, N), -1, np.int32)
    s2 = np.full((P, 3), -1.0)
    rc = lib.fo_host_refine_periodic_subset(ctypes.byref(pp), _lib._ptr(off), ng, _lib._ptr(idx), _lib._ptr(A),
                                            _lib._ptr(B), _lib._ptr(frac), _lib._ptr(which), len(which), 10, 2,
                                            _lib._ptr(d2), _lib._ptr(p2), _lib._ptr(s2))
    assert rc == 0
    rest = np.setdiff1d(np.arange(P), which)
    assert np.array_equal(d2[which], dist[which]) and np.array_equal(p2[which], pm[which])
    assert np.array_equal(s2[which], disp[which])
    assert (d2[rest] == -1).all() and (p2[rest] == -1).all()
    # validation: out-of-range index, decreasing offsets, more grouped atoms than atoms
    for bad_off, bad_idx in ((off, np.where(idx == 44, 45, idx).astype(np.int32)),
                             (np.array([0, 30, 20], np.int32), idx),
                             (np.array([0, 30, 50], np.int32), np.zeros(50, np.int32))):
        rc = lib.fo_host_refine_periodic(ctypes.byref(pp), _lib._ptr(bad_off), ng, _lib._ptr(bad_idx), _lib._ptr(A),
                                         _lib._ptr(B), _lib._ptr(frac), P, 10, 1, _lib._ptr(d2), None, None)
        assert rc == -1
        e = np.zeros((P, 1, 3))
        rc = lib.fo_host_refine_spherical(_lib._ptr(A), _lib._ptr(B), P, N, _lib._ptr(bad_off), ng, _lib._ptr(bad_idx),
                                          _lib._ptr(e), 1, 1, _lib._ptr(d2), None, None, None)
        assert rc == -1


def test_host_refine_spherical_hints():
    """fo_host_refine_spherical_hint: where hint_ok is set the given permutation is used as is (the device
    screening proved it optimal), elsewhere the LAP runs; with the LAP's own permutations as hints the results
    are bit-identical to the un-hinted call."""
    import ctypes
    from fastoverlap_b200 import _lib
    from fastoverlap_b200.utils import EulerM
    lib = _lib.load_library()
    g = golden("spherical_lj38.npz")
    X = g["pos1"] - g["pos1"].mean(0)
    N, P = len(X), 24
    rng = np.random.default_rng(8)
    A = np.repeat(X[None], P, 0) + rng.normal(scale=0.05, size=(P, N, 3))
    A -= A.mean(1, keepdims=True)
    B, eul = np.empty_like(A), np.zeros((P, 2, 3))
    for q in range(P):
        a, b, c = rng.uniform(0, 6), rng.uniform(0.2, 2.9), rng.uniform(0, 6)
        B[q] = (X @ EulerM(a, b, c).T)[rng.permutation(N)]
        eul[q, 0] = (a, b, c)
        eul[q, 1] = rng.uniform(0, 3, 3)
    dist, orient, pm, rmat = _lib.host_refine_spherical(A, B, eul, nthreads=2)
    assert (orient == 0).all() and abs(np.median(dist) - 0.05 * np.sqrt(3 * N)) < 0.05
    off, idx, ng = _lib._group_arrays(None, N)
    hint = np.zeros((P, 2, N), np.int32)
    hint[:, 0] = pm
    ok = np.zeros((P, 2), np.int32)
    ok[::2, 0] = 1   # every other pair carries a hint for the normal orientation; the rest run the LAP
    d2, o2, p2, r2 = np.empty(P), np.empty(P, np.int32), np.empty((P, N), np.int32), np.empty((P, 3, 3))
    _lib.host_refine_counters(reset=True)
    rc = lib.fo_host_refine_spherical_hint(_lib._ptr(A), _lib._ptr(B), P, N, _lib._ptr(off), ng, _lib._ptr(idx),
                                           _lib._ptr(eul), 2, _lib._ptr(hint), _lib._ptr(ok), 2, _lib._ptr(d2),
                                           _lib._ptr(o2), _lib._ptr(p2), _lib._ptr(r2))
    assert rc == 0 and _lib.host_refine_counters()[1] == P // 2
    assert np.array_equal(d2, dist) and np.array_equal(p2, pm) and np.array_equal(r2, rmat)
    # a deliberately wrong hint is used as given (the caller vouches for it): the distance gets worse
    hint[0, 0] = np.roll(pm[0], 1)
    lib.fo_host_refine_spherical_hint(_lib._ptr(A), _lib._ptr(B), P, N, _lib._ptr(off), ng, _lib._ptr(idx),
                                      _lib._ptr(eul), 2, _lib._ptr(hint), _lib._ptr(ok), 2, _lib._ptr(d2),
                                      _lib._ptr(o2), _lib._ptr(p2), _lib._ptr(r2))
    assert d2[0] > dist[0] + 0.1 and np.array_equal(d2[1:], dist[1:])


def test_spherical_align_batch_default_scale():
    """align_batch with the constructor's default scale=None computes the kernel width from the
    batch (the reference's averageSeparation rule) instead of raising AttributeError or reusing a stale one."""
    from fastoverlap_b200.spherical import SphericalAlign
    g = golden("spherical_lj38.npz")
    X = g["pos1"] - g["pos1"].mean(0)
    sa = SphericalAlign.__new__(SphericalAlign)
    sa.calcScale, sa.orientation, sa.perm, sa.Jmax = True, "distance", None, 15
    seen = {}

    def full_batch(X1, X2, perm, invert, nthreads):   # stands in for the native call (no GPU here)
        seen["scale"] = sa.scale
        return np.zeros(len(X1)), np.zeros((len(X1), 2, 3))

    sa._full_batch = full_batch
    A = np.repeat(X[None], 3, 0)
    d, R = sa.align_batch(A, A, nthreads=1)
    expect = 2 * sa.averageSeparation(X) / 6
    assert abs(seen["scale"] - expect) < 1e-12 and d.shape == (3,)


def test_native_spherical_refine_matches_reference_and_python():
    from fastoverlap_b200 import _lib
    from fastoverlap_b200.spherical import SphericalAlign
    from fastoverlap_b200.utils import indtoEuler
    g = golden("spherical_lj38.npz")
    X1 = g["pos1"] - g["pos1"].mean(0)
    X2 = g["pos2"] - g["pos2"].mean(0)
    eul = np.stack([indtoEuler(g["J15_findmax"].real, 32), indtoEuler(g["J15_findmax_inv"].real, 32)])[None]
    dist, orient, pm, rmat = _lib.host_refine_spherical(X1, X2, eul)
    assert abs(dist[0] - 1.4767670631638872) < 1e-9 and orient[0] == 1
    d0 = _lib.host_refine_spherical(X1, X2, eul[:, :1])[0]
    assert abs(d0[0] - float(g["J15_dist_normal_only"])) < 1e-9
    assert abs(abs(np.linalg.det(rmat[0])) - 1) < 1e-9
    sa = SphericalAlign.__new__(SphericalAlign)
    sa.perm, sa.scale, sa.Jmax = None, 0.5, 9
    rng = np.random.default_rng(9)
    A = rng.normal(size=(5, 21, 3))
    B = rng.normal(size=(5, 21, 3))
    A -= A.mean(1, keepdims=True)
    B -= B.mean(1, keepdims=True)
    E = rng.uniform(0, 3, size=(5, 2, 3))
    groups = [np.arange(9), np.arange(9, 21)]
    dist, orient, pm, rmat = _lib.host_refine_spherical(A, B, E, groups, nthreads=2)
    for i in range(5):
        d = min(sa.refine(A[i], B[i], E[i, 0], groups)[0], sa.refine(A[i], -B[i], E[i, 1], groups)[0])
        assert abs(d - dist[i]) < 1e-9
