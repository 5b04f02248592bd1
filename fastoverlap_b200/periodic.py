"""PeriodicAlign -- drop-in for reference fastoverlap/periodicAlignment.py:341-479 with the hot
path (structure factors, cross-spectrum, 3-D DFT, arg-max) on the GPU through the C ABI.

Same constructor, call signature, return tuple and attributes as the reference class; the
host-side refinement (PBC Hungarian <-> mean displacement, periodicAlignment.py:27-80) stays on
the CPU.  Additive batched entry points: align_batch, findDisps_batch, alignGroup(pairs=...).
"""
import numpy as np
from numpy.linalg import norm

from . import _lib
from .utils import find_best_permutation, _next_fast_len


class BasePeriodicAlignment(object):
    """Host steps shared by the periodic classes (the reference's base class, periodicAlignment.py:19-110), each a
    call into the native host library: refine = fo_host_refine_periodic, Hungarian = fo_host_best_permutation."""

    def findDisps(self, pos1, pos2):
        raise NotImplementedError

    def align(self, pos1, pos2):
        return self.refine(pos1, pos2, self.findDisps(pos1, pos2))


    def refine(self, x, y, disps, niter=10):
        """Permutational alignment <-> mean-displacement iteration from every candidate displacement, in one
        native call; the candidate with the smallest final distance is kept (the reference refines only the
        candidate with the smallest initial distance, :42-45, which can never come out better).
        Returns (distance, X1, X2, perm, disp)."""
        x = np.asarray(x, float).reshape(-1, 3)
        y = np.asarray(y, float).reshape(-1, 3)
        F = getattr(self, "fshape", (1,))[0]  # the native call takes displacements in units of box / F
        frac = np.atleast_2d(np.asarray(disps, float)) / self.boxvec * F
        k = len(frac)
        p = _lib.Context.per_params(len(x), self.boxvec, 1, F, 1.0)
        dist, perm, disp = _lib.host_refine_periodic(p, self.perm, np.broadcast_to(x, (k,) + x.shape),
                                                     np.broadcast_to(y, (k,) + y.shape), frac, niter, 1)
        b = int(np.argmin(dist))
        return float(dist[b]), self.periodic(x, True), self.periodic(y[perm[b]] - disp[b]), perm[b].tolist(), disp[b]

    def periodic(self, x, copy=False):
        """Wrap into the cell centred on the origin (in place unless copy)."""
        if copy:
            x = np.array(x, float)
        x -= np.round(x / self.boxvec) * self.boxvec
        return x

    def get_disp(self, X1, X2):
        return self.periodic(X1 - X2)

    def get_dist(self, X1, X2):
        return norm(self.get_disp(X1, X2))

    def cost_matrix(self, X1, X2):
        """cost[i, j] = minimum-image distance |X1[i] - X2[j]| (row = X1: the orientation that reproduces the
        reference's documented result)."""
        return norm(self.periodic(X1[:, None, :] - X2[None, :, :]), axis=2)

    def Hungarian(self, X1, X2):
        perm = find_best_permutation(X1, X2, self.perm, box=self.boxvec)[1]
        return self.get_dist(X1, np.asarray(X2)[perm]), perm

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
        self._C1 = self._C2 = None
        self.fabs = None

    # Intermediates the reference keeps as attributes (:362-394, :433-440).  The device path never materialises
    # them -- only `fabs` comes back -- so they are derived on first access: C1 / C2 on the device from the last
    # positions, the rest from those on the host (a 19^3 product and one numpy FFT; not on the alignment path).
    @property
    def C1(self):
        if self._C1 is None:
            self._C1 = self.calcFourierCoeff(self.pos1)
        return self._C1

    @property
    def C2(self):
        if self._C2 is None:
            self._C2 = self.calcFourierCoeff(self.pos2)
        return self._C2

    @property
    def _damp(self):
        return np.exp(-self.absks ** 2 * self.scale ** 2)

    @property
    def C(self):
        return (self.C1 * self.C2.conj() * self._damp).sum(0)

    @property
    def Csum(self):
        return float(((np.abs(self.C1) ** 2 + np.abs(self.C2) ** 2) * self._damp[None]).sum() * 0.5 * self.factor)

    @property
    def f(self):
        return np.fft.fftn(self.C, self.fshape)

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
            self._C1 = self._C2 = None
            bi, bv, fr, grid, st = self.ctx.per_align_pairs(p, self.pos1, self.pos2, want_grid=True)
        else:
            self._C1, self._C2 = np.asarray(Cs[0]), np.asarray(Cs[1])
            bi, bv, fr, grid, st = self.ctx.per_align_coeffs(p, self._C1, self._C2, want_grid=True)
        self.fabs = grid[0]
        self._best_idx, self._best_val, self._frac_idx = bi[0], bv[0], fr[0]

    def findDisps(self, pos1, pos2, Cs=None, npeaks=1, width=2):
        self.setPos(pos1, pos2, Cs)
        if npeaks > 1:
            # top-k by fit-and-subtract on the device (fo_grid_find_peaks; reference :444-451); the kernel's
            # limits (64 peaks, window half-width 4) bound what the reference accepts without limit
            pk, _, _, _, nf, _ = self.ctx.grid_find_peaks(self.fabs, min(int(npeaks), 64), min(int(width), 4))
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
        p = self._params()
        if self.n <= 11:
            # structure factors once per structure; the 48 images of pos2 are index permutations of its bank entry
            # inside the cross-spectrum (OHTRANSFORMCOEFFS, fastbulk.f90:863-1380), then the host pool refines
            bank = self.ctx.per_bank_create(p, np.stack([pos1, pos2]))
            try:
                fr = self.ctx.per_align_bank_ops(p, bank, np.tile([0, 1], (len(ops), 1)), ops)[2]
            finally:
                bank.close()
            dists = _lib.host_refine_periodic(p, self.perm, X1s, X2s, fr, niter=10, nthreads=nthreads)[0]
        else:  # fine k-grids: the images go through the batched hot path as independent pairs
            dists, disps, perms = self.align_batch(X1s, X2s, nthreads=nthreads)
            fr = None
        best = int(np.argmin(dists))
        if fr is None:
            fr = self.ctx.per_align_pairs(p, pos1, X2s[best])[2]
            best_fr = fr[0]
        else:
            best_fr = fr[best]
        res = self.refine(pos1, X2s[best], (best_fr * self.boxvec / np.array(self.fshape, float))[None, :])
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
        """All-vs-all alignment of a list of structures (reference :462-479): structure factors once per
        structure (device-resident bank), cross-spectrum + DFT + arg-max per pair, then the native host pool on
        all pairs at once.  npeaks > 1: per pair through align (top-k displacements from the device)."""
        coords = np.asarray(coords, float).reshape(-1, self.Natoms, 3)
        nl = len(coords)
        ii, jj = (a.ravel() for a in np.meshgrid(np.arange(nl), np.arange(nl), indexing="ij"))
        if npeaks > 1:
            coeffs = [self.calcFourierCoeff(c) for c in coords]
            res = [self.align(coords[i], coords[j], [coeffs[i], coeffs[j]], npeaks, width)[:3] for i, j in zip(ii, jj)]
            dists = np.array([r[0] for r in res]).reshape(nl, nl)
            x1 = np.array([r[1] for r in res])
            x2 = np.array([r[2] for r in res])
        else:
            p = self._params()
            bank = self.ctx.per_bank_create(p, coords)
            try:
                fr = self.ctx.per_align_bank(p, bank, np.stack([ii, jj], axis=1))[2]
            finally:
                bank.close()
            d, perms, disps = _lib.host_refine_periodic(p, self.perm, coords[ii], coords[jj], fr, 10, 0)
            dists = d.reshape(nl, nl)
            if keepCoords:
                x1 = self.periodic(coords[ii], True)
                x2 = self.periodic(np.take_along_axis(coords[jj], perms[:, :, None].astype(int), axis=1) -
                                   disps[:, None, :])
        if not keepCoords:
            return dists
        aligned = np.empty((2, nl, nl, self.Natoms, self.dim))
        aligned[0] = x1.reshape(nl, nl, self.Natoms, self.dim)
        aligned[1] = x2.reshape(nl, nl, self.Natoms, self.dim)
        return dists, aligned
