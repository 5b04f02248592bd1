"""findPeaks -- the reference's top-k peak extraction (utils.py:366-396) on a host grid, run on the device:
arg-max, Levenberg-Marquardt Gaussian fit on the periodic window, subtraction, repeated (fo_grid_find_peaks,
csrc/fo_peaks.cu).  The alignment paths (findRotations, findDisps with npeaks > 1) use the fused forms that never
copy the grid to the host; this entry point serves callers that already hold one."""
import numpy as np

from . import _lib
from .utils import findMax


def findPeaks(a, npeaks=10, width=2, ctx=None):
    """Up to npeaks (fractional indices (k, 3), amplitudes, means, sigmas, residual grid) of a 3-D array.
    As in the reference a search that finds nothing returns the interpolated maximum (findMax)."""
    a = np.asarray(a, float)
    if a.ndim != 3:
        raise NotImplementedError("the device peak search handles 3-D grids (what both alignment paths produce)")
    ctx = ctx or _lib.default_context()
    pk, amp, mean, alpha, nf, res = ctx.grid_find_peaks(a, min(int(npeaks), 64), min(int(width), 4),
                                                        want_residual=True)
    k = int(nf[0])
    if k == 0:
        return findMax(a)[None, :], [float(a.max())], [0], [np.nan], res[0]
    with np.errstate(invalid="ignore", divide="ignore"):
        sigma = [(2 * s) ** -0.5 for s in alpha[0, :k]]
    return pk[0, :k], list(amp[0, :k]), list(mean[0, :k]), sigma, res[0]
