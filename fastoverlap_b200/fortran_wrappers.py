"""The reference's f2py wrapper classes (SphericalAlignFortran sphericalAlignment.py:441-533,
SphericalHarmonicAlignFortran :535-663, PeriodicAlignFortran periodicAlignment.py:482-605) with
their constructor / call signatures and return tuples, bound to the GPU-backed facade modules of
fastoverlap_b200.f90 instead of the Fortran extension modules."""
import numpy as np
from numpy import sqrt

from . import f90
from .periodic import BasePeriodicAlignment
from .spherical import BaseSphericalAlignment


class _ClusterWrapper(BaseSphericalAlignment):
    def setPerm(self, perm):
        self.Natoms = sum(map(len, perm))
        self.perm = perm
        self.nperm = len(perm)
        self.npermsize = list(map(len, perm))
        self.permgroup = np.concatenate([np.asanyarray(p) + 1 for p in perm])
        self.fast.fastoverlaputils.setperm(self.Natoms, self.permgroup, self.npermsize)

    def align(self, pos1, pos2, perm=None, invert=True, debug=False):
        return self(pos1, pos2, perm, invert, 1, debug)

    def _prepare(self, pos1, pos2, perm, invert):
        if perm is not None:
            self.setPerm(perm)
        elif len(pos1) != self.Natoms:
            self.Natoms = len(pos1)
            self.setPerm([np.arange(self.Natoms)])
        self.fast.commons.perminvopt = invert
        return np.array(pos1, dtype=float).flatten(), np.array(pos2, dtype=float).flatten()


class SphericalAlignFortran(_ClusterWrapper):
    def __init__(self, scale=0.3, Jmax=15, perm=None, Natoms=None):
        self.scale = scale
        self.Jmax = Jmax
        self.fast = f90.fastclusters
        self.Natoms = Natoms
        self.perm = perm
        if perm is not None:
            self.setPerm(perm)
        elif Natoms is not None:
            self.setPerm([np.arange(Natoms)])
        self.malign = self.__call__

    def __call__(self, pos1, pos2, perm=None, invert=True, nrot=10, debug=False):
        """(distance, X1, X2, rmatbest) as the reference wrapper (:491-533)."""
        coordsb, coordsa = self._prepare(pos1, pos2, perm, invert)
        dist, _, rmatbest = self.fast.clusterfastoverlap.align(coordsb, coordsa, debug, self.Jmax,
                                                               self.scale, nrot)
        return dist, coordsb.reshape(self.Natoms, 3), coordsa.reshape(self.Natoms, 3), rmatbest


class SphericalHarmonicAlignFortran(_ClusterWrapper):
    def __init__(self, scale=0.3, Jmax=15, harmscale=1.0, nmax=20, perm=None, Natoms=None):
        self.scale = scale
        self.Jmax = Jmax
        self.harmscale = harmscale
        self.nmax = nmax
        self.fast = f90.fastclusters
        self.clus = self.fast.clusterfastoverlap
        self.Natoms = Natoms
        self.perm = perm
        if perm is not None:
            self.setPerm(perm)
        elif Natoms is not None:
            self.setPerm([np.arange(Natoms)])
        self.malign = self.__call__

    def __call__(self, pos1, pos2, perm=None, invert=True, nrot=10, debug=False):
        coordsb, coordsa = self._prepare(pos1, pos2, perm, invert)
        dist, _, rmatbest = self.clus.alignharm(coordsb, coordsa, debug, self.nmax, self.Jmax,
                                                self.harmscale, self.scale, nrot)
        return dist, coordsb.reshape(self.Natoms, 3), coordsa.reshape(self.Natoms, 3), rmatbest

    def compareList(self, poslist, perm=None):
        """(avgoverlap, maxoverlap, navgoverlap, nmaxoverlap), reference :622-663."""
        coords = np.array(poslist, dtype=float)
        nlist, Natoms, dim = coords.shape
        assert dim == 3
        if perm is None:
            if Natoms != self.Natoms:
                self.Natoms = Natoms
                self.setPerm([np.arange(self.Natoms)])
        else:
            self.setPerm(perm)
        coords -= coords.mean(1)[:, None, :]
        coordslist = np.rollaxis(coords.reshape(nlist, -1), -1)
        avg, mx = self.clus.calcoverlapmatrices(coordslist, self.nmax, self.Jmax, self.harmscale, self.scale)
        da, dm = avg.diagonal(), mx.diagonal()
        return avg, mx, avg / sqrt(da[:, None] * da[None, :]), mx / sqrt(dm[:, None] * dm[None, :])


class PeriodicAlignFortran(BasePeriodicAlignment):
    def __init__(self, Natoms, boxVec=None, scale=0, perm=None):
        self.Natoms = 1 if Natoms is None else Natoms
        self.boxvec = np.array(boxVec, dtype=float)
        self.scale = scale
        self.fast = f90.fastbulk
        self.bulk = self.fast.bulkfastoverlap
        self.setPerm([np.arange(self.Natoms)] if perm is None else perm)

    def setPerm(self, perm):
        if len(perm):
            self.perm = perm
            self.nperm = len(perm)
            self.npermsize = list(map(len, perm))
            self.permgroup = np.concatenate([np.asanyarray(p) + 1 for p in perm])
            self.Natoms = len(self.permgroup)
        else:
            self.nperm = 1
            self.npermsize = [self.Natoms]
            self.permgroup = np.arange(self.Natoms) + 1
            self.perm = [self.permgroup - 1]
        self.fast.fastoverlaputils.setperm(self.Natoms, self.permgroup, self.npermsize)

    def align(self, pos1, pos2, ndisps=10, perm=None, ohcell=False, debug=False):
        """(distance, X1, X2, perm) as the reference wrapper (periodicAlignment.py:526-568)."""
        coordsb = np.array(pos1, dtype=float).flatten()
        coordsa = np.array(pos2, dtype=float).flatten()
        if perm is not None:
            self.setPerm(perm)
        if coordsa.size != 3 * self.Natoms:
            self.setPerm([np.arange(coordsa.size // 3)])
        self.fast.commons.ohcellt = ohcell
        dist = self.bulk.align(coordsb, coordsa, debug, self.boxvec[0], self.boxvec[1], self.boxvec[2],
                               self.scale, ndisps)[0]
        return (dist, coordsb.reshape(self.Natoms, 3), coordsa.reshape(self.Natoms, 3),
                self.fast.commons.bestperm.copy())

    def alignGroup(self, coords, ndisps=1):
        """(dists, aligned[Natoms,3,nlist,nlist]), reference :570-605."""
        coordslist = np.asanyarray(coords, dtype=float)
        nlist, natoms, dim = coordslist.shape
        assert dim == 3 and natoms == self.Natoms
        coordslist = coordslist.reshape(nlist, -1).T
        s, nwave, ncoeff = self.bulk.calcdefaults(natoms, *self.boxvec)
        dists, aligned = self.bulk.aligngroup(coordslist, coordslist, False, self.boxvec[0], self.boxvec[1],
                                              self.boxvec[2], s, ndisps, nwave, ncoeff, True)
        return dists, aligned.reshape(natoms, 3, nlist, nlist)
