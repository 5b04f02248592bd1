"""PeriodicAlign -- drop-in for reference fastoverlap/periodicAlignment.py:341-479 with the hot
path (structure factors, cross-spectrum, 3-D DFT, arg-max) on the GPU through the C ABI.

Same constructor, call signature, return tuple and attributes as the reference class; the
host-side refinement (PBC Hungarian <-> mean displacement, periodicAlignment.py:27-80) stays on
the CPU.  Additive batched entry points: align_batch, findDisps_batch, alignGroup(pairs=...).
"""
import numpy as np
from numpy.linalg import norm

from . import _lib
from .utils import find_best_permutation, findMax, _next_fast_len


class BasePeriodicAlignment(object):
    """Host refinement shared by the periodic classes (reference periodicAlignment.py:19-110)."""

    def findDisps(self, pos1, pos2):
        raise NotImplementedError

    def align(self, pos1, pos2):
        disps = self.findDisps(pos1, pos2)
        return self.refine(pos1, pos2, disps)

    def refine(self, x, y, disps, niter=10):
        """Permutational alignment <-> mean-displacement update (reference :27-80).
        Returns (distance, X1, X2, perm, disp)."""
        disps = np.atleast_2d(disps)
        distperm = [self.Hungarian(x, y - disp[None, :]) + (disp,) for disp in disps]
        dist, saveperm, disp = min(distperm, key=lambda t: t[0])
        disp = np.array(disp, dtype=float)
        perm = saveperm
        for _ in range(niter):
            dxs = self.get_disp(x, (y - disp)[saveperm])
            disp -= dxs.mean(0)
            perm = self.Hungarian(x, y - disp[None, :])[1]
            if all(p1 == p2 for p1, p2 in zip(saveperm, perm)):
                break
            saveperm = perm
        dxs = self.get_disp(x, (y - disp)[perm])
        disp -= dxs.mean(0)
        pos1 = self.periodic(x, True)
        pos2 = self.periodic(y[perm] - disp)
        dist = self.get_dist(pos1, pos2)
        return dist, pos1, pos2, perm, disp

    def periodic(self, x, copy=False):
        if copy:
            x = x.copy()
        x -= np.round(x / self.boxvec) * self.boxvec
        return x

    def get_disp(self, X1, X2):
        return self.periodic(X1 - X2)

    def get_dist(self, X1, X2):
        return norm(self.get_disp(X1, X2))

    def cost_matrix(self, X1, X2):
        """cost[i, j] = minimum-image distance |X1[i] - X2[j]|.  The reference returns the
        transpose (periodicAlignment.py:94-102) because it targets pele's LAP convention; with the
        row = X1 convention used by utils.lap this is the orientation that reproduces the
        reference's documented result."""
        disps = X1[:, None, :] - X2[None, :, :]
        disps -= np.round(disps / self.boxvec[None, None, :]) * self.boxvec[None, None, :]
        return norm(disps, axis=2)

    def Hungarian(self, X1, X2):
        _, permlist = find_best_permutation(X1, X2, self.perm, user_cost_matrix=self.cost_matrix)
        dist = self.get_dist(X1, X2[permlist])
        return dist, permlist

    def __call__(self, pos1, pos2, *args, **kwargs):
        return self.align(pos1, pos2, *args, **kwargs)


class PeriodicAlign(BasePeriodicAlignment):
    """Best alignment of two configurations of a periodic system (reference :341-479).

    Parameters as the reference: Natoms, boxvec, permlist=None, dim=3, scale=None, n=None.
    Extra keyword: ctx (a fastoverlap_b200.Context; default = process-wide context).
    """

    def __init__(self, Natoms, boxvec, permlist=None, dim=3, scale=None, n=None, ctx=None):
        if dim != 3:
            raise NotImplementedError("only dim=3 is supported (the reference marks dim!=3 "
                                      "'TODO: TEST', periodicAlignment.py:360)")
        self.Natoms = Natoms
        self.boxvec = np.array(boxvec, dtype=float)
        self.dim = dim
        if permlist is None:
            self.perm = [np.arange(self.Natoms)]
        else:
            self.perm = list(map(np.array, permlist))
        self.pos1 = np.zeros((self.Natoms, self.dim))
        self.pos2 = np.zeros((self.Natoms, self.dim))
        if scale is None:
            scale = (np.prod(self.boxvec) / self.Natoms) ** (1. / self.dim) / 3.
        self.n = int(np.ceil(1.3 * Natoms ** (1. / 3.))) if n is None else n
        self.scale = scale
        self.factor = 2 * (np.pi * self.scale ** 2) ** (-self.dim * 0.5) * self.scale ** 2 / np.prod(self.boxvec)
        self._ctx = ctx
        self.setks()

    # -- plumbing
    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _params(self):
        self.ctx.set_perm(self.perm, self.Natoms)
        return _lib.Context.per_params(self.Natoms, self.boxvec, self.n, self.fshape[0], self.scale)

    def setks(self):
        ps = np.indices((self.n * 2 + 1,) * self.dim) - self.n
        self.ks = 2. * np.pi / self.boxvec[(slice(None),) + (None,) * self.dim] * ps
        self.absks = norm(self.ks, axis=0)
        shape = np.array(self.absks.shape) * 2 + 1
        self.fshape = tuple(_next_fast_len(int(d)) for d in shape)
        self.C1 = None
        self.C2 = None
        self.C = None
        self.f = None
        self.fabs = None

    def setScale(self, scale):
        self.scale = scale
        self.factor = 2 * (np.pi * self.scale ** 2) ** (-self.dim * 0.5) * self.scale ** 2 / np.prod(self.boxvec)

    # -- hot path
    def calcFourierCoeff(self, pos, out=None):
        """Structure factors per permutation group, shape (nperm, 2n+1, 2n+1, 2n+1)
        (reference :400-406) -- computed on the GPU."""
        C = self.ctx.per_structure_factors(self._params(), np.asarray(pos, float))[0]
        if out is not None:
            out[...] = C
            return out
        return C

    def setPos(self, pos1=None, pos2=None, Cs=None):
        """Overlap array of two structures (reference :408-440).  Sets self.fabs (the |f| grid),
        self.C1/self.C2 when Cs is given, and the arg-max results used by findDisps."""
        if pos1 is not None:
            self.pos1[:] = np.asanyarray(pos1)
        if pos2 is not None:
            self.pos2[:] = np.asanyarray(pos2)
        p = self._params()
        if Cs is None:
            bi, bv, fr, grid, st = self.ctx.per_align_pairs(p, self.pos1, self.pos2, want_grid=True)
        else:
            self.C1, self.C2 = Cs
            bi, bv, fr, grid, st = self.ctx.per_align_coeffs(p, self.C1, self.C2, want_grid=True)
        self.fabs = grid[0]
        self._best_idx, self._best_val, self._frac_idx = bi[0], bv[0], fr[0]

    def findDisps(self, pos1, pos2, Cs=None, npeaks=1, width=2):
        self.setPos(pos1, pos2, Cs)
        if npeaks > 1:
            # top-k by fit-and-subtract on the device (fo_grid_find_peaks; reference :444-451)
            pk, _, _, _, nf, _ = self.ctx.grid_find_peaks(self.fabs, npeaks, width)
            disps = pk[0, :int(nf[0])]
            if len(disps):
                disps = disps * self.boxvec / self.fabs.shape
            else:
                disps = (self._frac_idx * self.boxvec / self.fabs.shape)[None, :]
            return disps
        disp = self._frac_idx * self.boxvec / self.fabs.shape
        return disp[None, :]

    def align(self, pos1, pos2, Cs=None, npeaks=1, width=2):
        disps = self.findDisps(pos1, pos2, Cs, npeaks, width)
        return self.refine(self.pos1, self.pos2, disps)

    def align_oh(self, pos1, pos2, nthreads=0):
        """Alignment of a CUBIC cell over the 48 octahedral symmetry operations (the reference's
        OHCELLT branch, ALIGN1 fastbulk.f90:458-480, whose own implementation multiplies the
        untransformed coefficients): pos2 is transformed by every operation, the 48
        (pos1, R pos2) pairs go through the GPU hot path in one batch and through the host
        refinement pool, and the operation with the smallest distance is kept.
        Returns (dist, X1, X2, perm, disp, R) -- the reference's tuple plus the winning operation."""
        from .utils import oh_operations
        if np.ptp(self.boxvec) > 1e-12 * self.boxvec.max():
            raise ValueError("O_h cell symmetries need a cubic box, got %s" % (self.boxvec,))
        pos1 = np.asarray(pos1, float).reshape(self.Natoms, 3)
        pos2 = np.asarray(pos2, float).reshape(self.Natoms, 3)
        ops = oh_operations()
        X2s = np.einsum("oij,aj->oai", ops, pos2)
        X1s = np.broadcast_to(pos1, X2s.shape).copy()
        dists, disps, perms = self.align_batch(X1s, X2s, nthreads=nthreads)
        best = int(np.argmin(dists))
        bi, bv, fr, _, st = self.ctx.per_align_pairs(self._params(), pos1, X2s[best])
        res = self.refine(pos1, X2s[best], (fr[0] * self.boxvec / np.array(self.fshape, float))[None, :])
        return tuple(res) + (ops[best],)

    # -- batched, additive API
    def findDisps_batch(self, pos1, pos2):
        """pos1, pos2: (P, Natoms, 3).  Returns (disps (P,3), best_idx (P,3), best_val (P,))."""
        bi, bv, fr, _, st = self.ctx.per_align_pairs(self._params(), pos1, pos2)
        return fr * self.boxvec / np.array(self.fshape, float), bi, bv

    def align_batch(self, pos1, pos2, refine=True, nthreads=0, want_perm=True):
        """Align P independent pairs in one native call (fo_per_align_pairs_full): GPU hot path, on the device
        the nearest-partner screening of the assignment with the permutation <-> mean-displacement loop, and
        the host LAP pool (nthreads; 0 = all cores) only for the pairs whose screening fails, overlapped with
        the GPU's next chunk.  Returns (dists (P,), disps (P,3), perms (P,Natoms) or None); with refine=False
        (None, grid displacements (P,3), None).  self.last_nhost = pairs the host pool took."""
        pos1 = np.asarray(pos1, float).reshape(-1, self.Natoms, 3)
        pos2 = np.asarray(pos2, float).reshape(-1, self.Natoms, 3)
        p = self._params()
        if not refine:
            fr = self.ctx.per_align_pairs(p, pos1, pos2)[2]
            return None, fr * self.boxvec / np.array(self.fshape, float), None
        dists, perms, disps, fr, st, self.last_nhost = self.ctx.per_align_pairs_full(
            p, pos1, pos2, niter=10, nthreads=nthreads, want_perm=want_perm)
        return dists, disps, perms

    def alignGroup(self, coords, keepCoords=False, npeaks=1, width=2):
        """All-vs-all alignment of a list of structures (reference :462-479): structure factors
        once per structure (device-resident bank), then cross-spectrum + DFT + arg-max per pair."""
        coords = np.asarray(coords, float).reshape(-1, self.Natoms, 3)
        nl = len(coords)
        if npeaks > 1:
            coeffs = [self.calcFourierCoeff(p) for p in coords]
        p = self._params()
        bank = self.ctx.per_bank_create(p, coords)
        ii, jj = np.meshgrid(np.arange(nl), np.arange(nl), indexing="ij")
        pairs = np.stack([ii.ravel(), jj.ravel()], axis=1)
        _, _, fr, _, _ = self.ctx.per_align_bank(p, bank, pairs)
        bank.close()
        disps = fr * self.boxvec / np.array(self.fshape, float)
        dists = np.zeros((nl, nl))
        if keepCoords:
            aligned = np.empty((2, nl, nl, self.Natoms, self.dim))
        for k, (i, j) in enumerate(pairs):
            if npeaks > 1:
                dist, x1, x2 = self.align(coords[i], coords[j], [coeffs[i], coeffs[j]], npeaks, width)[:3]
            else:
                dist, x1, x2 = self.refine(coords[i], coords[j], disps[k:k + 1])[:3]
            if keepCoords:
                aligned[0, i, j] = x1
                aligned[1, i, j] = x2
            dists[i, j] = dist
        return (dists, aligned) if keepCoords else dists
