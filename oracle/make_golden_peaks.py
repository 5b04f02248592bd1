"""TEST INFRASTRUCTURE: golden vectors for row a8 (top-k peaks) from the UNMODIFIED reference.

    python oracle/make_golden_peaks.py      (build container only: needs /root/reference)

Runs fastoverlap.utils.findPeaks (utils.py:366-396, scipy curve_fit) on
  * the reference's own LJ38 overlap grid (Jmax = 15, normal orientation) and BLJ256 |f| grid, read
    from the fixtures make_golden.py froze (tests/golden/spherical_lj38.npz, periodic_blj256.npz);
  * a seeded synthetic grid of planted Gaussians (the test regenerates it from the same seed)
and writes tests/golden/peaks.npz (peak positions, amplitudes, means, packed exponents, and a checksum of
the residual grid)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def planted_grid(seed=5, n=(24, 20, 28), k=4):
    """Sum of k anisotropic Gaussians at fractional positions + a small smooth background."""
    rng = np.random.default_rng(seed)
    idx = np.indices(n).astype(float)
    f = 0.05 * np.cos(2 * np.pi * idx[0] / n[0]) * np.cos(2 * np.pi * idx[2] / n[2])
    for j in range(k):
        x0 = np.array([rng.uniform(4, m - 4) for m in n])
        amp = 10.0 / (1 + 0.6 * j)
        s = rng.uniform(0.15, 0.4, size=3)
        off = rng.uniform(-0.05, 0.05, size=3)
        d = idx - x0[:, None, None, None]
        q = (s[0] * d[0] ** 2 + s[1] * d[1] ** 2 + s[2] * d[2] ** 2 + off[0] * d[0] * d[1] +
             off[1] * d[0] * d[2] + off[2] * d[1] * d[2])
        f += amp * np.exp(-q)
    return f


def run(findPeaks, grid, npeaks, width):
    peaks, amp, mean, sigma, f = findPeaks(grid, npeaks=npeaks, width=width)
    with np.errstate(invalid="ignore", divide="ignore"):
        alpha = np.array([0.5 / np.asarray(s) ** 2 for s in sigma])
    return dict(peaks=np.asarray(peaks), amplitude=np.asarray(amp), mean=np.asarray(mean), alpha=alpha,
                resid_sum=f.sum(), resid_abs=np.abs(f).sum(), resid_max=f.max())


def main():
    refshim.install()
    from fastoverlap.utils import findPeaks
    out = {}
    g = np.load(os.path.join(OUT, "spherical_lj38.npz"))
    for k, v in run(findPeaks, g["J15_grid"], 5, 2).items():
        out["lj38_" + k] = v
    g = np.load(os.path.join(OUT, "periodic_blj256.npz"))
    for k, v in run(findPeaks, g["fabs"], 4, 2).items():
        out["blj256_" + k] = v
    for k, v in run(findPeaks, planted_grid(), 4, 2).items():
        out["planted_" + k] = v
    for k in ("lj38", "blj256", "planted"):
        print(k, "peaks\n", out[k + "_peaks"], "\namp", out[k + "_amplitude"])
    np.savez_compressed(os.path.join(OUT, "peaks.npz"), **out)


if __name__ == "__main__":
    main()
