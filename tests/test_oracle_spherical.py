"""CPU: the C oracle (oracle/fo_oracle_spherical.c) against golden vectors produced by the
unmodified reference (oracle/make_golden.py) and against scipy special functions."""
import numpy as np
import pytest
from scipy.special import sph_harm_y, ive

import oracle
from conftest import golden, groups_from


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max()


def test_ylm_against_scipy():
    rng = np.random.default_rng(0)
    for p in list(rng.normal(size=(4, 3))) + [np.array([0, 0, 1.0]), np.array([0, 0, -2.0]),
                                               np.array([1.0, 0, 0])]:
        L = 15
        Y, r = oracle.sph_ylm(p, L)
        th, ph = np.arccos(p[2] / r), np.arctan2(p[1], p[0])
        for l in range(L + 1):
            m = np.arange(-l, l + 1)
            assert np.abs(Y[l, m] - sph_harm_y(l, m, th, ph)).max() < 1e-13


def test_bessel_against_scipy():
    for x in [1e-8, 1e-3, 0.1, 0.99, 1.0, 5.0, 22.0, 130.0, 600.0]:
        for L in (7, 15, 31):
            a = oracle.sphi_scaled(L, x)
            b = ive(np.arange(L + 1) + 0.5, x) * np.sqrt(np.pi / 2 / x)
            assert np.abs(a / b - 1).max() < 1e-12, (x, L)


def test_soft_tables_and_isoft():
    s = golden("soft_tables.npz")
    for bw in (4, 8, 11, 16):
        assert np.abs(oracle.wigner_table(bw) - s["Ds_%d" % bw]).max() < 1e-12
        assert np.abs(oracle.soft_weights(bw) - s["weights_%d" % bw]).max() < 1e-15
        o = oracle.isoft(s["flmm_%d" % bw], bw - 1, want_complex=True)
        assert rel(o, s["isoft_%d" % bw]) < 1e-13


def test_lj38_direct_path():
    g = golden("spherical_lj38.npz")
    X1 = g["pos1"] - g["pos1"].mean(0)
    X2 = g["pos2"] - g["pos2"].mean(0)
    for J in (14, 15):
        k = "J%d_" % J
        I = oracle.sph_coeffs_direct(X1, X2, J, 0.3)
        assert rel(I, g[k + "Ilmm"]) < 1e-13
        bi, bv, fr, grids, _ = oracle.sph_align_pairs(X1, X2, J, 0.3, True, want_grid=True)
        assert rel(grids[0, 0], g[k + "grid"]) < 1e-13
        assert rel(grids[0, 1], g[k + "grid_inv"]) < 1e-13
        assert tuple(bi[0, 0]) == tuple(g[k + "argmax"])
        assert tuple(bi[0, 1]) == tuple(g[k + "argmax_inv"])
        assert np.allclose(fr[0, 0], g[k + "findmax"].real, atol=1e-8)
        assert np.allclose(fr[0, 1], g[k + "findmax_inv"].real, atol=1e-8)
    assert tuple(g["J15_argmax"]) == (21, 24, 12) and tuple(g["J15_argmax_inv"]) == (26, 26, 0)
    assert abs(grids[0, 0].max() - 7.2159906786356895) < 1e-12
    assert abs(grids[0, 1].max() - 7.675997319763542) < 1e-12


def test_lj38_harmonic_path():
    """Closed-form radial integrals vs the reference's numpy formulation (which carries ~1e-10
    cancellation noise at nmax=20) and vs the restated Fortran recurrence."""
    g = golden("spherical_lj38.npz")
    X1 = g["pos1"] - g["pos1"].mean(0)
    X2 = g["pos2"] - g["pos2"].mean(0)
    C1 = oracle.sph_harm_coeffs(X1, 20, 15, 1.0, 0.3)
    C1f = oracle.sph_harm_coeffs(X1, 20, 15, 1.0, 0.3, kind="fortran")
    assert rel(C1, g["H_c1"]) < 5e-10
    assert rel(C1f, C1) < 1e-10
    C2 = oracle.sph_harm_coeffs(X2, 20, 15, 1.0, 0.3)
    assert rel(oracle.sph_dot_harm(g["H_c1"], g["H_c2"]), g["H_Ilmm"]) < 1e-14
    assert rel(oracle.sph_dot_harm(g["H_c1"], g["H_c2"], invert=True), g["H_Ilmm_inv"]) < 1e-14
    Ih = oracle.sph_dot_harm(C1, C2)
    assert rel(Ih, g["H_Ilmm"]) < 1e-10
    assert rel(oracle.isoft(Ih, 15), g["H_grid"]) < 1e-10


def test_harmonic_radial_closed_form_vs_mpmath():
    """d_nl(r): closed form (long double) against mpmath quadrature of the defining integral
    4 pi int g_nl(r') exp(-(r'^2+r^2)/2s^2) i_l(r' r/s^2) r'^2 dr' (three spot values)."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    sigma, r0 = mp.mpf("0.3"), mp.mpf(1)

    def exact(n, l, r):
        nu = l + mp.mpf(1) / 2
        N = mp.sqrt(2 * mp.factorial(n) * r0 ** (-2 * l - 3) / mp.gamma(mp.mpf(3) / 2 + n + l))
        gf = lambda x: N * x ** l * mp.exp(-x * x / (2 * r0 * r0)) * mp.laguerre(n, nu, x * x / (r0 * r0))
        il = lambda z: mp.sqrt(mp.pi / (2 * z)) * mp.besseli(nu, z)
        f = lambda x: gf(x) * mp.exp(-(x * x + r * r) / (2 * sigma ** 2)) * il(x * r / sigma ** 2) * x * x
        return 4 * mp.pi * mp.quad(f, [0, r / 2, r, 2 * r, 4 * r + 2])

    for (n, l, r) in [(3, 2, 1.0), (12, 7, 2.0), (20, 15, 2.5)]:
        d = oracle.sph_harm_radial(20, 15, r, 0.3, 1.0)[n, l]
        e = float(exact(n, l, mp.mpf(r)))
        assert abs(d - e) < 1e-13 * max(abs(e), 1e-3)


def test_synthetic_cases():
    g = golden("spherical_synth.npz")
    for i in range(int(g["ncases"])):
        k = "c%d_" % i
        perm = groups_from(g[k + "groups"], g[k + "gsizes"])
        p1, p2 = g[k + "pos1"], g[k + "pos2"]
        X1, X2 = p1 - p1.mean(0), p2 - p2.mean(0)
        J, sc = int(g[k + "Jmax"]), float(g[k + "scale"])
        I = oracle.sph_coeffs_direct(X1, X2, J, sc, perm)
        assert rel(I, g[k + "Ilmm"]) < 1e-13
        assert rel(oracle.isoft(I, J), g[k + "grid"]) < 1e-13
        for gi, idx in enumerate(perm):
            C = oracle.sph_harm_coeffs(X1, 12, J, 1.0, sc, idx=idx)
            assert rel(C, g[k + "H_c1"][gi]) < 1e-9
        assert rel(oracle.sph_dot_harm(g[k + "H_c1"], g[k + "H_c2"]), g[k + "H_Ilmm"]) < 1e-13


def test_oracle_rotation_refinement_vs_reference_golden():
    """oracle.sph_energy_gradient / sph_max_overlap (numpy restatement of sphericalAlignment.py:67-103)
    against the reference's own getEnergyGradient / maxOverlap outputs (tests/golden/refine.npz)."""
    G = golden("refine.npz")
    lj = golden("spherical_lj38.npz")
    sy = golden("spherical_synth.npz")
    src = {"J14": lj["J14_Ilmm"], "J15inv": lj["J15_Ilmm_inv"], "H": lj["H_Ilmm"], "c0": sy["c0_Ilmm"],
           "c1": sy["c1_Ilmm"], "c2": sy["c2_Ilmm"]}
    Ls = dict(zip([str(k) for k in G["keys"]], [int(j) for j in G["Jmax"]]))
    for k, I in src.items():
        L = Ls[k]
        for tag in ("p", "0"):
            E, g = oracle.sph_energy_gradient(G[k + "_R" + tag], I.conj(), L)
            assert abs(E - float(G[k + "_E" + tag])) <= 1e-13 * abs(E), k
            assert np.abs(g - G[k + "_G" + tag]).max() <= 1e-12 * max(1.0, abs(E)), k
        R, f = oracle.sph_max_overlap(G[k + "_R0"], I, L)
        assert abs(f + float(G[k + "_E"])) <= 1e-12 * abs(f), k
        assert np.abs(R - G[k + "_R"]).max() < 1e-9, k
