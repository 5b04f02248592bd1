"""TEST INFRASTRUCTURE ONLY (container-only; needs /root/reference).

Freezes the outputs of the reference's continuous rotation refinement
(BaseSphericalAlignment.findRotation / maxOverlap / getEnergyGradient,
reference fastoverlap/sphericalAlignment.py:67-113,190-194) into tests/golden/refine.npz:

  for every coefficient set already frozen in spherical_lj38.npz / spherical_synth.npz
    <k>_R0     Euler angles of the interpolated grid maximum (indtoEuler(findMax(iSOFT(I))))
    <k>_E0     getEnergyGradient(R0, conj(I))[0]          (energy at the start point)
    <k>_G0     getEnergyGradient(R0, conj(I))[1]          (gradient at the start point)
    <k>_R      maxOverlap's refined Euler angles (scipy L-BFGS-B)
    <k>_E      res.fun at the optimum
  plus getEnergyGradient at a few fixed off-grid rotations (<k>_Rp, _Ep, _Gp).

Run:  python oracle/make_golden_refine.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refshim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    fo = refshim.install()
    from fastoverlap.utils import findMax
    lj = np.load(os.path.join(OUT, "spherical_lj38.npz"))
    sy = np.load(os.path.join(OUT, "spherical_synth.npz"))
    sets = [("J14", 14, lj["J14_Ilmm"]), ("J14inv", 14, lj["J14_Ilmm_inv"]),
            ("J15", 15, lj["J15_Ilmm"]), ("J15inv", 15, lj["J15_Ilmm_inv"]),
            ("H", 15, lj["H_Ilmm"]), ("Hinv", 15, lj["H_Ilmm_inv"])]
    for i in range(int(sy["ncases"])):
        sets.append(("c%d" % i, int(sy["c%d_Jmax" % i]), sy["c%d_Ilmm" % i]))
    rng = np.random.default_rng(67)
    out = {"keys": np.array([k for k, _, _ in sets]), "Jmax": np.array([j for _, j, _ in sets])}
    for k, Jmax, I in sets:
        sa = fo.SphericalAlign(0.3, Jmax)
        g = sa.soft.iSOFT(I)
        R0 = sa.soft.indtoEuler(findMax(g))
        E0, G0 = sa.getEnergyGradient(R0, I.conj())
        R, res = sa.maxOverlap(R0, I.conj())
        Rp = np.array([rng.uniform(0, 2 * np.pi), rng.uniform(0.05, np.pi - 0.05), rng.uniform(0, 2 * np.pi)])
        Ep, Gp = sa.getEnergyGradient(Rp, I.conj())
        out.update({k + "_R0": R0, k + "_E0": E0, k + "_G0": G0, k + "_R": R, k + "_E": res.fun,
                    k + "_nit": res.nit, k + "_Rp": Rp, k + "_Ep": Ep, k + "_Gp": Gp})
        print("%-7s Jmax %2d  E0 %.12g -> E %.12g  |G0| %.3e  nit %d  dR %s" % (
            k, Jmax, E0, res.fun, np.abs(G0).max(), res.nit, R - R0))
    np.savez_compressed(os.path.join(OUT, "refine.npz"), **out)


if __name__ == "__main__":
    main()
