"""Facade with the call surface of the reference's f2py extension modules
(reference fastoverlap/f90/__init__.py:3-21): `fastbulk` and `fastclusters` objects whose
sub-objects and functions have the names, argument order, in-place coordinate updates and return
tuples of the Fortran modules, backed by libfastoverlap_b200.so.

A maintainer of the reference makes the Fortran-wrapper classes GPU-backed with

    import fastoverlap.f90 as f90, fastoverlap_b200.f90 as b200
    f90.fastbulk, f90.fastclusters = b200.fastbulk, b200.fastclusters
    f90.have_fastbulk = f90.have_fastclusters = True

(INTEGRATION.md).  Differences from the Fortran modules, all deliberate: state lives
in this object (not in Fortran SAVE variables), errors raise FastOverlapError instead of STOP.
"""
import numpy as np

from .. import _lib
from ..periodic import PeriodicAlign
from ..spherical import SphericalAlign, SphericalHarmonicAlign
from ..utils import EulerM, findrotation

have_fastbulk = True
have_fastclusters = True
have_libbnb = False      # branch-and-bound is a different algorithm, out of scope
have_fortran = False


class _Commons(object):
    """commons.f90:23-28 flags that the wrappers poke through f2py."""
    def __init__(self):
        self.perminvopt = True
        self.ohcellt = False
        self.bestperm = np.zeros(0, dtype=int)


class _Utils(object):
    """fastoverlaputils.setperm (fastutils.f90:117-167): 1-based concatenated groups + sizes."""
    def __init__(self, owner):
        self._owner = owner

    def setperm(self, natoms, permgroup, npermsize):
        permgroup = np.asarray(permgroup, dtype=int).ravel() - 1
        sizes = [int(s) for s in np.atleast_1d(npermsize)]
        off = np.cumsum([0] + sizes)
        self._owner.natoms = int(natoms)
        self._owner.perm = [permgroup[off[i]:off[i + 1]] for i in range(len(sizes))]

    def setnatoms(self, natoms):
        self._owner.natoms = int(natoms)
        self._owner.perm = [np.arange(int(natoms))]


class _Module(object):
    def __init__(self, ctx=None):
        self._ctx = ctx
        self.natoms = 0
        self.perm = None
        self.commons = _Commons()
        self.fastoverlaputils = _Utils(self)

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _perm_for(self, natoms):
        if self.perm is None or self.natoms != natoms:
            self.fastoverlaputils.setnatoms(natoms)
        return self.perm


class _BulkFastOverlap(object):
    """bulkfastoverlap (fastbulk.f90): calcdefaults :253, align :279, aligngroup :336."""
    def __init__(self, mod):
        self._m = mod

    def calcdefaults(self, natoms, boxlx, boxly, boxlz):
        kw = (boxlx * boxly * boxlz / natoms) ** (1. / 3.) / 3.
        nwave = int(np.ceil(2 * 3.14159265359 / min(boxlx, boxly, boxlz) * 1.5 / kw))
        from ..utils import _next_fast_len
        return kw, nwave, _next_fast_len(4 * nwave + 3)

    def _aligner(self, natoms, box, kernelwidth, nwave=None):
        scale = None if kernelwidth is None or kernelwidth <= 0 else kernelwidth
        return PeriodicAlign(natoms, box, self._m._perm_for(natoms), scale=scale, n=nwave, ctx=self._m.ctx)

    def align(self, coordsb, coordsa, debug, boxlx, boxly, boxlz, kernelwidth, ndisplacements):
        """(distance, dist2); coordsa is overwritten with the aligned, permuted structure."""
        natoms = coordsb.size // 3
        al = self._aligner(natoms, [boxlx, boxly, boxlz], kernelwidth)
        if self._m.commons.ohcellt:  # 48 octahedral operations of a cubic cell (ALIGN1, fastbulk.f90:458-480)
            dist, X1, X2, perm, disp, R = al.align_oh(coordsb.reshape(natoms, 3), coordsa.reshape(natoms, 3))
            coordsa[:] = X2.ravel()
            self._m.commons.bestperm = np.asarray(perm, dtype=int) + 1
            return dist, dist ** 2
        nd = 10 if ndisplacements == 0 else int(ndisplacements)
        dist, X1, X2, perm, disp = al.align(coordsb.reshape(natoms, 3), coordsa.reshape(natoms, 3),
                                            npeaks=nd)
        coordsa[:] = X2.ravel()
        self._m.commons.bestperm = np.asarray(perm) + 1
        return dist, dist ** 2

    def aligngroup(self, coords1, coords2, debug, boxlx, boxly, boxlz, kwidth, ndisps, nwave, nfspace, sym):
        """(distmat[n1,n2], aligned[3N,n1,n2]); coords are [3N, nlist] (Fortran order)."""
        c1 = np.asarray(coords1).T.reshape(coords1.shape[1], -1, 3)
        c2 = np.asarray(coords2).T.reshape(coords2.shape[1], -1, 3)
        natoms = c1.shape[1]
        al = self._aligner(natoms, [boxlx, boxly, boxlz], kwidth, nwave)
        p = al._params()
        ctx = self._m.ctx
        bank = ctx.per_bank_create(p, np.concatenate([c1, c2]))
        n1, n2 = len(c1), len(c2)
        pairs = np.array([(i, n1 + j) for i in range(n1) for j in range(n2) if not sym or j >= i])
        _, _, fr, _, _ = ctx.per_align_bank(p, bank, pairs)
        bank.close()
        # host stage for all pairs at once (native pool), then the aligned images of the second structures
        ia, ib = pairs[:, 0], pairs[:, 1] - n1
        d, perms, disps = _lib.host_refine_periodic(p, al.perm, c1[ia], c2[ib], fr, 10, 0)
        X2 = al.periodic(np.take_along_axis(c2[ib], perms[:, :, None].astype(int), axis=1) - disps[:, None, :])
        distmat = np.zeros((n1, n2))
        aligned = np.zeros((3 * natoms, n1, n2))
        distmat[ia, ib] = d
        aligned[:, ia, ib] = X2.reshape(len(pairs), -1).T
        if sym:
            lower = (ib < n1) & (ia < n2)
            distmat[ib[lower], ia[lower]] = d[lower]
        return distmat, aligned


class _ClusterFastOverlap(object):
    """clusterfastoverlap (fastclusters.f90): align :129, alignharm :271, calcoverlapmatrices :853."""
    def __init__(self, mod):
        self._m = mod

    def _finish(self, al, coordsb, coordsa, nrotations):
        natoms = coordsb.size // 3
        pos1, pos2 = coordsb.reshape(natoms, 3), coordsa.reshape(natoms, 3)
        invert = bool(self._m.commons.perminvopt)
        perm = self._m._perm_for(natoms)
        if nrotations == 1:
            dist, X1, X2 = al.align(pos1, pos2, perm, invert)
        else:
            dist, X1, X2 = al.malign(pos1, pos2, perm, invert, nrot=int(nrotations) or 10)
            d1, Y1, Y2 = al.align(pos1, pos2, perm, invert)
            if d1 < dist:
                dist, X1, X2 = d1, Y1, Y2
        # rotation that maps centred pos2 (up to permutation / inversion) onto X2
        c2 = pos2 - pos2.mean(0)
        from ..utils import find_best_permutation
        best = None
        for sgn in ((1.0, -1.0) if invert else (1.0,)):
            _, p = find_best_permutation(X2, sgn * c2, perm)
            d, M = findrotation(X2, (sgn * c2)[p])
            if best is None or d < best[0]:
                best = (d, sgn * M)
        coordsb[:] = X1.ravel()
        coordsa[:] = X2.ravel()
        return dist, dist ** 2, best[1]

    def align(self, coordsb, coordsa, debug, l, kwidth, nrotations):
        al = SphericalAlign(kwidth if kwidth > 0 else None, int(l), ctx=self._m.ctx)
        return self._finish(al, coordsb, coordsa, nrotations)

    def alignharm(self, coordsb, coordsa, debug, n, l, hwidth, kwidth, nrotations):
        al = SphericalHarmonicAlign(kwidth if kwidth > 0 else None, hwidth, int(n), int(l), ctx=self._m.ctx)
        return self._finish(al, coordsb, coordsa, nrotations)

    def calcoverlapmatrices(self, coordslist, n, l, hwidth, kwidth):
        """(norms, maxovers)[nlist,nlist]; coordslist is [3N, nlist] of centred structures."""
        coords = np.asarray(coordslist).T.reshape(coordslist.shape[1], -1, 3)
        al = SphericalHarmonicAlign(kwidth, hwidth, int(n), int(l), ctx=self._m.ctx)
        avg, mx, _, _ = al.compareList(coords, perm=self._m._perm_for(coords.shape[1]))
        return avg, mx


def _make(kind, ctx=None):
    m = _Module(ctx)
    if kind == "bulk":
        m.bulkfastoverlap = _BulkFastOverlap(m)
    else:
        m.clusterfastoverlap = _ClusterFastOverlap(m)
    return m


fastbulk = _make("bulk")
fastclusters = _make("clusters")
