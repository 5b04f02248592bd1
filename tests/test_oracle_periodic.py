"""CPU: the C oracle (oracle/fo_oracle_periodic.c) against golden vectors produced by the
unmodified reference (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import oracle
from conftest import golden, groups_from


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max()


def test_next_fast_len_table():
    tab = golden("next_fast_len.npz")["table"]
    for i in range(len(tab)):
        assert oracle.next_fast_len(i) == tab[i]


def test_fft1d_against_numpy():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 5, 12, 15, 40, 72, 135, 22, 49):
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        assert rel(oracle.fft1d(x, -1), np.fft.fft(x)) < 1e-13
        assert rel(oracle.fft1d(x, +1), np.fft.ifft(x) * n) < 1e-13


def test_blj256_stages():
    g = golden("periodic_blj256.npz")
    perm = [np.arange(204), np.arange(204, 256)]
    n, F, sc = int(g["n"]), int(g["F"]), float(g["scale"])
    C1 = oracle.per_structure_factors(g["pos1"], g["box"], n, perm)
    C2 = oracle.per_structure_factors(g["pos2"], g["box"], n, perm)
    assert rel(C1, g["C1"]) < 1e-13 and rel(C2, g["C2"]) < 1e-13
    C = oracle.per_cross_spectrum(C1, C2, g["box"], n, sc)
    assert rel(C, g["C"]) < 1e-13
    fabs = oracle.per_fft_abs(C, F)
    assert rel(fabs, g["fabs"]) < 1e-12
    idx, frac = oracle.find_max(fabs)
    assert tuple(idx) == tuple(g["argmax"]) == (10, 38, 32)
    assert np.allclose(frac, g["findmax"], rtol=0, atol=1e-9)
    assert abs(oracle.per_csum(C1, C2, g["box"], n, sc) - float(g["Csum"])) < 1e-9
    assert abs(fabs.max() - 55419.12238387397) < 1e-7


def test_synthetic_cases_whole_path():
    g = golden("periodic_synth.npz")
    for i in range(int(g["ncases"])):
        k = "c%d_" % i
        perm = groups_from(g[k + "groups"], g[k + "gsizes"])
        bi, bv, fr, grids, _ = oracle.per_align_pairs(
            g[k + "pos1"], g[k + "pos2"], g[k + "box"], int(g[k + "n"]), int(g[k + "F"]),
            float(g[k + "scale"]), perm, want_grid=True)
        assert rel(grids[0], g[k + "fabs"]) < 1e-12
        assert tuple(bi[0]) == tuple(g[k + "argmax"])
        assert np.allclose(fr[0], g[k + "findmax"], atol=1e-8)


def test_pairs_threads_deterministic():
    g = golden("periodic_synth.npz")
    k = "c1_"
    perm = groups_from(g[k + "groups"], g[k + "gsizes"])
    A = np.stack([g[k + "pos1"]] * 5)
    B = np.stack([g[k + "pos2"]] * 5)
    r1 = oracle.per_align_pairs(A, B, g[k + "box"], int(g[k + "n"]), int(g[k + "F"]),
                                float(g[k + "scale"]), perm, nthreads=1)
    r4 = oracle.per_align_pairs(A, B, g[k + "box"], int(g[k + "n"]), int(g[k + "F"]),
                                float(g[k + "scale"]), perm, nthreads=4)
    assert np.array_equal(r1[0], r4[0]) and np.array_equal(r1[1], r4[1])
    assert tuple(r1[0][0]) == tuple(g[k + "argmax"])
