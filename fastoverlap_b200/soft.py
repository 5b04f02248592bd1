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
        """Quadrature weights (soft.py:64-71)."""
        j = np.arange(0, 2 * bw).astype(float)[:, None]
        k = np.arange(0, bw).astype(float)[None, :]
        fudge = pi / 4 / bw
        return (2 / (2 * k + 1) * np.sin((2 * j + 1) * (2 * k + 1) * fudge) *
                np.sin((2 * j + 1) * fudge) / bw).sum(1)

    def SOFT(self, data):
        """Forward transform (soft.py:98-113); host side, not on the alignment path."""
        Jmax, bw = self.Jmax, self.bw
        data = np.asanyarray(data)
        assert all(n == self.n for n in data.shape)
        S2 = np.fft.fft(np.fft.fft(data, axis=0), axis=2) * (2. * bw) ** -2
        flmm = np.zeros((bw, bw * 2 - 1, bw * 2 - 1), np.complex128)
        for m1 in range(-Jmax, Jmax + 1):
            for m2 in range(-Jmax, Jmax + 1):
                l = max(abs(m1), abs(m2))
                flmm[l:, m1, m2] = self.Ds[l:, m1, m2].dot(self.weights * S2[m1, :, m2])
        return flmm

    def iSOFT(self, flmm):
        """Inverse transform onto the (2bw)^3 Euler grid (soft.py:115-125) -- on the GPU."""
        flmm = np.asarray(flmm)
        assert flmm.shape == (self.bw, self.bw * 2 - 1, self.bw * 2 - 1)
        return self.ctx.sph_isoft(flmm, self.Jmax)[0]

    def indtoEuler(self, ind):
        R = self.indFactor * np.atleast_2d(ind)
        R[:, 1] += 0.5 * pi / self.n
        return R.squeeze()
