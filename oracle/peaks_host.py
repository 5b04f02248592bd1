"""TEST INFRASTRUCTURE ONLY -- scipy restatement of the reference's top-k peak extraction (iterative Gaussian
fit-and-subtract, fastoverlap/utils.py:347-396; Fortran FINDPEAKS fastutils.f90:475-548), the checker the
device kernel (fo_grid_find_peaks, csrc/fo_peaks.cu) is tested against.  Never imported by the product."""
import numpy as np
from scipy.optimize import curve_fit


def findMax(a):
    """utils.py:319-338"""
    a = np.asanyarray(a)
    ind = np.unravel_index(a.argmax(), a.shape)
    d = np.empty(a.ndim)
    for ax in range(a.ndim):
        ip, im = list(ind), list(ind)
        ip[ax] = (ind[ax] + 1) % a.shape[ax]
        im[ax] = ind[ax] - 1
        y1, y2, y3 = np.abs(a[tuple(ip)]), np.abs(a[tuple(ind)]), np.abs(a[tuple(im)])
        d[ax] = (y3 - y1) / (2 * (2 * y2 - y1 - y3))
    return np.array(ind) - d


def _gaussian(x, A, mu, *alphax0):
    """A exp(-(x-x0)^T S (x-x0)) + mu with S upper-triangular packed (utils.py:347-353)."""
    x = np.atleast_2d(x)
    dim = len(x)
    S = np.zeros((dim, dim))
    S[np.triu_indices(dim)] = alphax0[:-dim]
    x0 = x - np.array(alphax0[-dim:])[:, None]
    return A * np.exp(-np.einsum("ik,jk,ij->k", x0, x0, S)) + mu


def fitPeak(f, ind, n=2):
    """Fit a Gaussian to the (2n+1)^dim periodic window around ind (utils.py:355-364)."""
    dim = f.ndim
    win = f[np.ix_(*[np.arange(i - n, i + n + 1) % s for i, s in zip(ind, f.shape)])]
    coords = np.indices((2 * n + 1,) * dim).reshape((dim, -1)) - n
    p0 = ([win[(n,) * dim], 0.] + [1. if i == j else 0. for i in range(dim) for j in range(i, dim)] +
          [0.] * dim)
    return curve_fit(_gaussian, coords, win.ravel(), p0=p0)


def findPeaks(a, npeaks=10, width=2):
    """Up to npeaks (fractional index, amplitude, mean, sigma) by fit-and-subtract (utils.py:366-396)."""
    f = np.array(a, dtype=float)
    f -= f.min()
    dim = f.ndim
    indices = np.indices(f.shape).reshape((dim, -1))
    peaks, amplitude, mean, sigma = [], [], [], []
    for _ in range(npeaks):
        ind = np.unravel_index(f.argmax(), f.shape)
        try:
            popt = fitPeak(f, ind, width)[0]
            peaks.append(popt[-dim:] + ind)
            amplitude.append(popt[0])
            mean.append(popt[1])
            with np.errstate(invalid="ignore", divide="ignore"):
                sigma.append((2 * popt[2:-dim]) ** -0.5)
            popt[-dim:] += ind
            f.ravel()[:] -= _gaussian(indices, *popt)
        except (RuntimeError, ValueError):
            break
    peaks = np.array(peaks)
    if len(peaks) == 0:
        peaks = findMax(a)[None, :]
        amplitude.append(np.max(a))
        mean.append(0)
        sigma.append(np.nan)
    return peaks, amplitude, mean, sigma, f
