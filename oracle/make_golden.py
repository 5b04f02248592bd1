"""TEST INFRASTRUCTURE: freeze golden vectors from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, see oracle/refshim.py):

    python oracle/make_golden.py

writes small .npz fixtures under tests/golden/.  The GPU box has no /root/reference, so the
-m gpu tests compare the CUDA path with these files and with the C oracle.

Inputs are the reference's own example data (examples/LJ38, examples/BLJ256 -- copied into the
fixtures as arrays, they are data, not code) plus seeded synthetic structures.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
BOX_BLJ = 5.975206329


def rand_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                     [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                     [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])


def golden_periodic(fo):
    ex = os.path.join(refshim.REFERENCE_ROOT, "examples", "BLJ256")
    pos1 = np.loadtxt(os.path.join(ex, "coords"))
    pos2 = np.loadtxt(os.path.join(ex, "finish"))
    box = np.ones(3) * BOX_BLJ
    perm = [np.arange(204), np.arange(204, 256)]
    al = refshim.periodic_align_pele(fo, 256, box, perm)
    dist, X1, X2, p, disp = al(pos1, pos2)
    disps0 = al.findDisps(pos1, pos2)
    from fastoverlap.utils import findMax
    out = dict(pos1=pos1, pos2=pos2, box=box, nA=204, n=al.n, F=al.fshape[0], scale=al.scale,
               C1=al.C1.copy(), C2=al.C2.copy(), C=al.C.copy(), fabs=al.fabs.copy(),
               argmax=np.array(np.unravel_index(al.fabs.argmax(), al.fabs.shape)),
               findmax=findMax(al.fabs), Csum=al.Csum, disp0=disps0[0], dist=dist, X1=X1, X2=X2,
               perm=np.array(p), disp=disp)
    np.savez_compressed(os.path.join(OUT, "periodic_blj256.npz"), **out)
    print("periodic_blj256: dist", repr(dist), "argmax", out["argmax"], "max", al.fabs.max())

    # small synthetic cases: ragged groups, non-cubic box, explicit n / scale, odd F
    rng = np.random.default_rng(42)
    cases = []
    for (N, groups, bx, n, scale) in [
        (12, [np.arange(12)], [3.1, 3.1, 3.1], 3, 0.4),
        (20, [np.arange(0, 13), np.arange(13, 20)], [4.0, 4.7, 5.3], 4, 0.35),
        (9, [np.array([0, 2, 4, 6, 8]), np.array([1, 3]), np.array([5, 7])], [2.5, 3.5, 3.0], 2, 0.5),
        (30, [np.arange(30)], [5.0, 5.0, 6.0], 5, None),
    ]:
        bx = np.array(bx, float)
        p1 = rng.uniform(-0.5, 0.5, size=(N, 3)) * bx
        shift = rng.uniform(0, 1, 3) * bx
        p2 = p1 + shift + rng.normal(scale=0.03, size=(N, 3))
        prm = np.concatenate([rng.permutation(g) for g in groups])
        order = np.arange(N)
        order[np.concatenate(groups)] = prm
        p2 = p2[order]
        al = refshim.periodic_align_pele(fo, N, bx, groups, scale=scale, n=n)
        dist, X1, X2, p, disp = al(p1, p2)
        cases.append(dict(pos1=p1, pos2=p2, box=bx, groups=np.concatenate(groups),
                          gsizes=np.array([len(g) for g in groups]), n=al.n, F=al.fshape[0],
                          scale=al.scale, C1=al.C1.copy(), C2=al.C2.copy(), C=al.C.copy(),
                          fabs=al.fabs.copy(), findmax=findMax(al.fabs),
                          argmax=np.array(np.unravel_index(al.fabs.argmax(), al.fabs.shape)),
                          dist=dist, disp=disp, Csum=al.Csum))
        print("periodic synth N=%d n=%d F=%d dist %.6g" % (N, al.n, al.fshape[0], dist))
    flat = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat["c%d_%s" % (i, k)] = v
    flat["ncases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "periodic_synth.npz"), **flat)

    # _next_fast_len table (utils.py:278-313)
    from fastoverlap.utils import _next_fast_len
    tab = np.array([_next_fast_len(i) for i in range(0, 401)])
    np.savez_compressed(os.path.join(OUT, "next_fast_len.npz"), table=tab)


def golden_spherical(fo):
    from fastoverlap.utils import findMax
    ex = os.path.join(refshim.REFERENCE_ROOT, "examples", "LJ38")
    pos1 = np.loadtxt(os.path.join(ex, "coords"))
    pos2 = np.loadtxt(os.path.join(ex, "finish"))
    out = dict(pos1=pos1, pos2=pos2)
    for Jmax in (14, 15):
        sa = fo.SphericalAlign(0.3, Jmax)
        X1, X2 = sa.COM_shift(pos1, pos2)
        I = sa.calcSO3Coeffs(X1, X2)
        Iinv = sa.calcSO3Coeffs(X1, -X2)
        g = sa.soft.iSOFT(I)
        ginv = sa.soft.iSOFT(Iinv)
        dist = sa(pos1, pos2)[0]
        dist_inv = sa(pos1, -pos2)[0]
        k = "J%d_" % Jmax
        out.update({k + "Ilmm": I, k + "Ilmm_inv": Iinv, k + "grid": g.real.copy(),
                    k + "grid_imag_max": np.abs(g.imag).max(), k + "grid_inv": ginv.real.copy(),
                    k + "findmax": findMax(g), k + "findmax_inv": findMax(ginv),
                    k + "argmax": np.array(np.unravel_index(g.real.argmax(), g.shape)),
                    k + "argmax_inv": np.array(np.unravel_index(ginv.real.argmax(), g.shape)),
                    k + "dist": dist, k + "dist_inv": dist_inv})
        # per-orientation refined distances
        R = sa.soft.indtoEuler(findMax(g))
        Rinv = sa.soft.indtoEuler(findMax(ginv))
        out[k + "dist_normal_only"] = sa.refine(X1, X2, R)[0]
        out[k + "dist_inverted_only"] = sa.refine(X1, -X2, Rinv)[0]
        print("LJ38 Jmax=%d dist %r inv %r argmax %s %s" % (
            Jmax, dist, dist_inv, out[k + "argmax"], out[k + "argmax_inv"]))
    # harmonic path
    sh = fo.SphericalHarmonicAlign(0.3, 1.0, 20, 15)
    X1, X2 = sh.COM_shift(pos1, pos2)
    c1 = sh.calcHarmCoeff(X1)
    c2 = sh.calcHarmCoeff(X2)
    Ih = sh.calcSO3Harm([c1], [c2])
    Ihinv = sh.calcSO3Harm([c1], [c2], invert=True)
    out.update(dict(H_c1=c1, H_c2=c2, H_Ilmm=Ih, H_Ilmm_inv=Ihinv, H_grid=sh.soft.iSOFT(Ih).real.copy(),
                    H_grid_inv=sh.soft.iSOFT(Ihinv).real.copy(), H_dist=sh(pos1, pos2)[0]))
    print("LJ38 harmonic dist %r" % out["H_dist"])
    np.savez_compressed(os.path.join(OUT, "spherical_lj38.npz"), **out)

    # SOFT tables and a round trip at several bandwidths (soft.py:48-125)
    soft_out = {}
    rng = np.random.default_rng(7)
    for bw in (4, 8, 11, 16):
        s = fo.SOFT(bw)
        soft_out["Ds_%d" % bw] = s.Ds
        soft_out["weights_%d" % bw] = s.weights
        L = bw - 1
        f = np.zeros((bw, 2 * bw - 1, 2 * bw - 1), complex)
        for l in range(bw):
            for m1 in range(-l, l + 1):
                for m2 in range(-l, l + 1):
                    f[l, m1, m2] = rng.normal() + 1j * rng.normal()
        soft_out["flmm_%d" % bw] = f
        soft_out["isoft_%d" % bw] = s.iSOFT(f)
    np.savez_compressed(os.path.join(OUT, "soft_tables.npz"), **soft_out)

    # synthetic clusters: random cloud vs rotated+permuted copy (sphericalAlignment.py:711-732),
    # and perm-group cases; small N so the fixtures stay small
    rng = np.random.default_rng(20171013)
    flat = {}
    cases = [(13, 7, 0.5, None), (20, 10, 0.45, [np.arange(0, 12), np.arange(12, 20)]),
             (50, 9, 0.6, None)]
    for i, (N, Jmax, scale, groups) in enumerate(cases):
        p1 = rng.normal(size=(N, 3)) * 1.3
        R = rand_rotation(rng)
        p2 = p1.dot(R.T) + rng.normal(scale=0.02, size=(N, 3))
        if groups is None:
            order = rng.permutation(N)
        else:
            order = np.arange(N)
            for g in groups:
                order[g] = rng.permutation(g)
        p2 = p2[order]
        sa = fo.SphericalAlign(scale, Jmax, perm=groups)
        X1, X2 = sa.COM_shift(p1, p2)
        perm = groups if groups is not None else [np.arange(N)]
        I = sum(sa.calcSO3Coeffs(X1[p], X2[p]) for p in perm)
        g = sa.soft.iSOFT(I)
        dist = sa(p1, p2)[0]
        flat.update({"c%d_pos1" % i: p1, "c%d_pos2" % i: p2, "c%d_Jmax" % i: Jmax,
                     "c%d_scale" % i: scale, "c%d_Ilmm" % i: I, "c%d_grid" % i: g.real.copy(),
                     "c%d_findmax" % i: findMax(g), "c%d_dist" % i: dist,
                     "c%d_groups" % i: np.concatenate(perm),
                     "c%d_gsizes" % i: np.array([len(q) for q in perm])})
        sh = fo.SphericalHarmonicAlign(scale, 1.0, 12, Jmax, perm=groups)
        cs1 = np.array([sh.calcHarmCoeff(X1[p]) for p in perm])
        cs2 = np.array([sh.calcHarmCoeff(X2[p]) for p in perm])
        flat.update({"c%d_H_c1" % i: cs1, "c%d_H_c2" % i: cs2,
                     "c%d_H_Ilmm" % i: sh.calcSO3Harm(cs1, cs2)})
        print("spherical synth N=%d Jmax=%d dist %.3e" % (N, Jmax, dist))
    flat["ncases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "spherical_synth.npz"), **flat)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    fo = refshim.install()
    which = sys.argv[1:] or ["periodic", "spherical"]
    if "periodic" in which:
        golden_periodic(fo)
    if "spherical" in which:
        golden_spherical(fo)
