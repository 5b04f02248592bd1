"""SOFT -- drop-in for reference fastoverlap/soft.py:40-130 (SO(3) Fourier transform object).

The Wigner-d table and the inverse transform (the part on the alignment hot path) run on the
GPU; the forward transform (used only by the reference's round-trip self-check) is evaluated on
the host from the same table."""
import numpy as np
from numpy import pi

from . import _lib


class SOFT(object):
    def __init__(self, bw, ctx=None):
        self.bw = bw
        self.bws = np.arange(0, bw * 2)
        self.n = 2 * bw
        self.Jmax = bw - 1
        self._ctx = ctx
        self.weights = self.makeweights(bw)
        self.a = pi / bw * self.bws
        self.b = pi / 4 / bw * (2 * self.bws + 1)
        self.y = pi / bw * self.bws
        self._Ds = None
        self.indFactor = np.array([2 * pi / self.n, pi / self.n, 2 * pi / self.n])

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    @property
    def Ds(self):
        """Ds[l, m1, m2, k] = sqrt((2l+1)/2) d^l_{m1 m2}(beta_k) (soft.py:73-96), from the device."""
        if self._Ds is None:
            self._Ds = self.ctx.sph_wigner_table(self.Jmax)
        return self._Ds

    def calcWignerMatrices(self):
        """The table Ds (soft.py:73-96), computed on the device (fo_sph_wigner_table)."""
        return self.ctx.sph_wigner_table(self.Jmax)

    @classmethod
    def makeweights(cls, bw):
        """Quadrature weights of the 2 bw beta nodes: w_j = (2 / bw) sin(b_j) sum_k sin((2k+1) b_j) / (2k+1),
        b_j = pi (2j+1) / 4bw (the reference's makeweights, soft.py:64-71)."""
        b = pi * (2 * np.arange(2 * bw) + 1.0) / (4 * bw)
        odd = 2 * np.arange(bw) + 1.0
        return 2.0 / bw * np.sin(b) * (np.sin(np.outer(b, odd)) / odd).sum(1)

    def SOFT(self, data):
        """Forward transform (the reference's round-trip self-check, soft.py:98-113; not on the alignment
        path): 2-D FFT over (alpha, gamma), then the weighted Wigner-d quadrature over beta as one contraction
        with the device-computed table (which is zero for l < max(|m1|, |m2|))."""
        data = np.asanyarray(data)
        if data.shape != (self.n,) * 3:
            raise ValueError("expected a (%d, %d, %d) grid" % ((self.n,) * 3))
        S2 = np.fft.fft2(data, axes=(0, 2)) / self.n ** 2
        m = np.r_[0:self.bw, -(self.bw - 1):0]  # the 2 bw - 1 orders in wrapped (negative-index) storage order
        return np.einsum("lack,k,akc->lac", self.Ds, self.weights, S2[np.ix_(m % self.n, np.arange(self.n), m % self.n)])

    def iSOFT(self, flmm):
        """Inverse transform onto the (2bw)^3 Euler grid (soft.py:115-125) -- on the GPU."""
        flmm = np.asarray(flmm)
        assert flmm.shape == (self.bw, self.bw * 2 - 1, self.bw * 2 - 1)
        return self.ctx.sph_isoft(flmm, self.Jmax)[0]

    def indtoEuler(self, ind):
        from .utils import indtoEuler
        return indtoEuler(ind, self.n)
