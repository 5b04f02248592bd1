"""SphericalAlignFortran / SphericalHarmonicAlignFortran / PeriodicAlignFortran: the names, constructor and call
signatures and return tuples of the reference's f2py wrapper classes (sphericalAlignment.py:441-663,
periodicAlignment.py:482-605), generated over the facade modules of fastoverlap_b200.f90 -- the objects that
stand where the reference expects its compiled `fastclusters` / `fastbulk` extension modules.  A maintainer of
the reference does not need these classes at all: swapping the modules makes the
reference's own wrapper classes GPU-backed.  They exist so that code written against those classes runs
unchanged on `import fastoverlap_b200 as fastoverlap`."""
import numpy as np

from . import f90


class _OverFacade(object):
    """State the wrappers share: the facade module, the permutation groups last pushed into it."""
    _module = "fastclusters"

    def _bind(self, perm, natoms):
        self.fast = getattr(f90, self._module)
        self.perm, self.Natoms = None, natoms
        if perm is not None or natoms is not None:
            self._groups(natoms, perm)

    def _groups(self, natoms, perm=None):
        """Make `perm` (or the trivial single group of `natoms` atoms) the facade's permutation groups."""
        if perm is None:
            if self.perm is not None and natoms == self.Natoms:
                return
            perm = [np.arange(natoms)]
        self.perm = [np.asarray(g, int) for g in perm]
        self.Natoms = int(sum(len(g) for g in self.perm))
        self.fast.fastoverlaputils.setperm(self.Natoms, np.concatenate(self.perm) + 1, [len(g) for g in self.perm])

    setPerm = lambda self, perm: self._groups(None, perm)  # noqa: E731 -- the reference's public name

    @staticmethod
    def _flat(pos):
        return np.array(pos, dtype=float).ravel()


class _ClusterOverFacade(_OverFacade):
    def _entry(self, coordsb, coordsa, debug, nrot):
        raise NotImplementedError

    def __call__(self, pos1, pos2, perm=None, invert=True, nrot=10, debug=False):
        """-> (distance, X1, X2, rmatbest); pos2 comes back aligned and permuted."""
        b, a = self._flat(pos1), self._flat(pos2)
        self._groups(b.size // 3, perm)
        self.fast.commons.perminvopt = bool(invert)
        dist, _, rmat = self._entry(b, a, debug, nrot)
        return dist, b.reshape(-1, 3), a.reshape(-1, 3), rmat

    malign = __call__

    def align(self, pos1, pos2, perm=None, invert=True, debug=False):
        return self(pos1, pos2, perm, invert, 1, debug)


class SphericalAlignFortran(_ClusterOverFacade):
    def __init__(self, scale=0.3, Jmax=15, perm=None, Natoms=None):
        self.scale, self.Jmax = scale, Jmax
        self._bind(perm, Natoms)

    def _entry(self, b, a, debug, nrot):
        return self.fast.clusterfastoverlap.align(b, a, debug, self.Jmax, self.scale, nrot)


class SphericalHarmonicAlignFortran(_ClusterOverFacade):
    def __init__(self, scale=0.3, Jmax=15, harmscale=1.0, nmax=20, perm=None, Natoms=None):
        self.scale, self.Jmax, self.harmscale, self.nmax = scale, Jmax, harmscale, nmax
        self._bind(perm, Natoms)

    def _entry(self, b, a, debug, nrot):
        return self.fast.clusterfastoverlap.alignharm(b, a, debug, self.nmax, self.Jmax, self.harmscale, self.scale,
                                                      nrot)

    def compareList(self, poslist, perm=None):
        """-> (avgoverlap, maxoverlap, both normalised by their diagonals) of a list of structures."""
        X = np.array(poslist, dtype=float)
        if X.ndim != 3 or X.shape[2] != 3:
            raise ValueError("poslist must be (nlist, natoms, 3)")
        self._groups(X.shape[1], perm)
        X -= X.mean(1, keepdims=True)
        avg, mx = self.fast.clusterfastoverlap.calcoverlapmatrices(X.reshape(len(X), -1).T, self.nmax, self.Jmax,
                                                                   self.harmscale, self.scale)
        unit = lambda m: m / np.sqrt(np.outer(m.diagonal(), m.diagonal()))  # noqa: E731
        return avg, mx, unit(avg), unit(mx)


class PeriodicAlignFortran(_OverFacade):
    _module = "fastbulk"

    def __init__(self, Natoms, boxVec=None, scale=0, perm=None):
        self.boxvec, self.scale = np.array(boxVec, dtype=float), scale
        self._bind(perm if perm is not None and len(perm) else None, 1 if Natoms is None else Natoms)

    def align(self, pos1, pos2, ndisps=10, perm=None, ohcell=False, debug=False):
        """-> (distance, X1, X2, perm (1-based, as the Fortran module reports it))."""
        b, a = self._flat(pos1), self._flat(pos2)
        self._groups(a.size // 3, perm)
        self.fast.commons.ohcellt = bool(ohcell)
        dist = self.fast.bulkfastoverlap.align(b, a, debug, *self.boxvec, self.scale, ndisps)[0]
        return dist, b.reshape(-1, 3), a.reshape(-1, 3), self.fast.commons.bestperm.copy()

    def alignGroup(self, coords, ndisps=1):
        """-> (dists (nlist, nlist), aligned (natoms, 3, nlist, nlist))."""
        X = np.asanyarray(coords, dtype=float)
        nlist, natoms = X.shape[:2]
        self._groups(natoms)
        kw, nwave, nf = self.fast.bulkfastoverlap.calcdefaults(natoms, *self.boxvec)
        flat = X.reshape(nlist, -1).T
        dists, aligned = self.fast.bulkfastoverlap.aligngroup(flat, flat, False, *self.boxvec, kw, ndisps, nwave, nf, True)
        return dists, aligned.reshape(natoms, 3, nlist, nlist)
