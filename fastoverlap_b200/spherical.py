"""SphericalAlign / SphericalHarmonicAlign -- drop-ins for reference
fastoverlap/sphericalAlignment.py:29-414 with the hot path (coefficients, inverse SO(3) transform,
arg-max) on the GPU through the C ABI.  Host side (stays on the CPU, north_star): centring,
Hungarian permutation, Kearsley rotation, optional continuous refinement of the rotation.

Orientation rule.  The reference's numpy classes choose between the normal and the
inverted orientation by comparing L-BFGS-refined overlap values; its Fortran path refines both
and keeps the smaller distance.  Default here: orientation="distance" (Fortran rule; never worse
than the numpy rule and needs no optimiser).  orientation="overlap" reproduces the numpy rule.
"""
import numpy as np
from numpy import sin, cos, sqrt, exp, pi
from numpy.linalg import norm

from . import _lib
from .soft import SOFT
from .utils import find_best_permutation, EulerM, indtoEuler


def isoft_executed_flops(Jmax, invert=True):
    """FP64 flop executed by sph_isoft_kernel per pair (DFMA = 2): Wigner contraction shared by both
    orientations + per orientation the two symmetric 1-D transform stages."""
    L = Jmax
    B, W, L1, F, H = L + 1, 2 * L + 1, L + 1, 2 * L + 2, L + 2
    nnz_half = sum((2 * l + 1) * (l + 1) for l in range(L + 1))
    k5 = 2 * nnz_half * F                       # real x complex DFMA pairs
    stage_a = F * L1 * H * L * 4                # lines (k, m2) x outputs x m x 4 DFMA
    stage_b = F * F * H * L * 2                 # lines (a, k) x outputs x m2 x 2 DFMA
    O = 2 if invert else 1
    return 2 * (k5 + O * (stage_a + stage_b))


def isoft_fft_flops(Jmax, invert=True):
    """FP64 flop of the FFT form of the iSOFT (sph_isoft5_kernel, Jmax = 15): the Wigner contraction as above and,
    per orientation and beta plane, F/2 complex transforms of length F = 2 Jmax + 2 along alpha (the columns
    m2 >= 0) and F/2 along gamma (two real-output rows per complex transform), 5 F log2 F flop each."""
    L = Jmax
    F = 2 * L + 2
    nnz_half = sum((2 * l + 1) * (l + 1) for l in range(L + 1))
    k5 = 2 * nnz_half * F
    O = 2 if invert else 1
    return 2 * k5 + O * F * F * 5.0 * F * np.log2(F)


class BaseSphericalAlignment(object):
    calcScale = True
    orientation = "distance"

    # -- plumbing
    @property
    def ctx(self):
        if getattr(self, "_ctx", None) is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def calcSO3Coeffs(self, pos1, pos2):
        raise NotImplementedError

    def averageSeparation(self, pos):
        """Mean nearest-neighbour distance (reference :35-41)."""
        pos = np.asanyarray(pos).reshape(-1, 3)
        d = norm(pos[:, None, :] - pos[None, :, :], axis=2)
        np.fill_diagonal(d, np.inf)
        return d.min(1).sum() / len(pos)

    def setJ(self, Jmax):
        self.Jmax = Jmax
        self.J = np.arange(Jmax + 1)
        self.soft = SOFT(Jmax + 1, ctx=getattr(self, "_ctx", None))
        self.Js, self.m1s, self.m2s = map(np.array, zip(*[(l, m1, m2) for l in range(Jmax + 1)
                                                           for m1 in range(-l, l + 1)
                                                           for m2 in range(-l, l + 1)]))

    def _perm(self, n, perm=None):
        if perm is None:
            perm = [np.arange(n)] if self.perm is None else self.perm
        return perm

    def Hungarian(self, pos1, pos2, perm=None):
        return find_best_permutation(pos1, pos2, permlist=self._perm(len(pos1), perm))

    def rotate(self, X, R):
        return X.dot(EulerM(*R))

    def refine(self, X1, X2, R, permlist=None):
        """Rotate X2 by the Euler angles R, permute, Kearsley fit (reference :118-127) -- one call into the native
        host library (fo_host_refine_spherical).  Returns (distance, X1, aligned X2)."""
        X1 = np.asarray(X1, float)
        X2 = np.asarray(X2, float)
        dist, _, perm, rmat = _lib.host_refine_spherical(X1, X2, np.asarray(R, float).reshape(1, 1, 3),
                                                         self._perm(len(X1), permlist), 1)
        return float(dist[0]), X1, X2.dot(EulerM(*R))[perm[0]].dot(rmat[0].T)

    def sphHarm(self, theta, phi):
        """Y[l, m, j] = Y_lm(theta_j, phi_j) for l <= Jmax (reference :57-65: theta polar, phi azimuth,
        negative m at index m + 2 Jmax + 1) -- evaluated on the GPU (fo_sph_ylm) from the unit vectors."""
        theta = np.asarray(theta, float).ravel()
        phi = np.asarray(phi, float).ravel()
        assert theta.shape == phi.shape
        u = np.stack([sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta)], axis=1)
        return self.ctx.sph_ylm(u, self.Jmax)[0][0]

    def COM_shift(self, pos1, pos2):
        X1 = np.array(pos1, float)
        X2 = np.array(pos2, float)
        X1 -= X1.mean(axis=0)[None, :]
        X2 -= X2.mean(axis=0)[None, :]
        return X1, X2

    # -- continuous refinement of the rotation (reference :67-113) on the device (fo_refine.cu)
    def getEnergyGradient(self, rot, Ilmm):
        """(E, dE/d(a,b,g)) of the reference's objective (:93-96), evaluated on the device
        (fo_sph_overlap_gradient).  Ilmm is the CONJUGATED coefficient array, as the reference passes it."""
        val, grad, _ = self.ctx.sph_overlap_gradient(np.conj(Ilmm), self.Jmax, np.asarray(rot, float))
        return -float(val[0]), -grad[0]

    def maxOverlap(self, R, Ilmm):
        """Refine the Euler angles R by maximising the un-weighted overlap (reference :98-103).
        Ilmm is the CONJUGATED coefficient array, as the reference passes it (:193).  Returns
        (R, res) with res.x, res.fun = -overlap, res.nfev like scipy's OptimizeResult."""
        eu, ov, ne = self.ctx.sph_refine_rotations(np.conj(Ilmm), self.Jmax, np.asarray(R, float))

        class _Res(object):
            pass
        res = _Res()
        res.x, res.fun, res.nfev, res.success = eu[0], -float(ov[0]), int(ne[0]), True
        return eu[0], res

    # -- hot path wrappers
    def _grid_search(self, X1, X2, perm, invert, want_grid=False, calcCoeffs=None):
        """Both orientations: Euler angles of the interpolated grid maximum (GPU)."""
        raise NotImplementedError

    def findRotation(self, Ilmm):
        """Grid arg-max -> Euler angles -> continuous refinement, all on the device (reference :190-194)."""
        bi, bv, fr, _ = self.ctx.sph_isoft_argmax(Ilmm, self.Jmax)
        R = self.soft.indtoEuler(fr[0, 0])
        R, res = self.maxOverlap(R, np.conj(Ilmm))
        return R, res.fun

    def findRotations(self, Ilmm, nrot=10, width=2):
        """Top-nrot rotations by Gaussian fit-and-subtract (reference :196-204): iSOFT and the peak
        search both run on the device (fo_sph_isoft_peaks); the grid is not copied to the host.
        Returns (Euler angles (k,3), amplitude, mean, sigma, None) -- the reference's last element, the
        residual grid, is available from ctx.grid_find_peaks(..., want_residual=True)."""
        from .utils import findMax
        while True:
            if width > 4:  # no fit converged: the reference falls back to the interpolated maximum
                overlap = self.ctx.sph_isoft(Ilmm, self.Jmax, want_imag=False)[0]
                return (np.atleast_2d(self.soft.indtoEuler(findMax(overlap)[None, :])), [overlap.max()], [0],
                        [np.nan], None)
            pk, amp, mean, alpha, nf = self.ctx.sph_isoft_peaks(Ilmm, self.Jmax, npeaks=nrot, width=width)
            k = int(nf[0])
            if k > 0:
                break
            width += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            sigma = [(2 * a) ** -0.5 for a in alpha[0, :k]]
        return np.atleast_2d(self.soft.indtoEuler(pk[0, :k])), list(amp[0, :k]), list(mean[0, :k]), sigma, None

    def _setup(self, pos1, pos2, perm):
        pos1 = np.asanyarray(pos1, dtype=float).reshape(-1, 3)
        pos2 = np.asanyarray(pos2, dtype=float).reshape(-1, 3)
        if self.calcScale:
            self.scale = (self.averageSeparation(pos1) + self.averageSeparation(pos2)) / 6
        perm = self._perm(len(pos1), perm)
        X1, X2 = self.COM_shift(pos1, pos2)
        return X1, X2, perm

    def align(self, pos1, pos2, perm=None, invert=True, calcCoeffs=None):
        """(dist, X1, X2) for the best of the normal / inverted orientation (reference :160-188)."""
        X1, X2, perm = self._setup(pos1, pos2, perm)
        if self.orientation == "overlap" and calcCoeffs is None:
            # numpy rule, fused: coefficients -> iSOFT -> arg-max -> refinement of both orientations
            Rs, ov = self._grid_search_refined(X1, X2, perm, invert)
            if invert and ov[1] > ov[0]:
                return self.refine(X1, -X2, Rs[1], perm)
            return self.refine(X1, X2, Rs[0], perm)
        if self.orientation == "overlap" or calcCoeffs is not None:
            # numpy rule: compare the refined overlaps of the two orientations
            Ilmm = self._coeffs(X1, X2, perm) if calcCoeffs is None else calcCoeffs(False)
            R, res = self.findRotation(Ilmm)
            if invert:
                Ilmm = self._coeffs(X1, -X2, perm) if calcCoeffs is None else calcCoeffs(True)
                invR, invres = self.findRotation(Ilmm)
                if invres < res:
                    R = invR
                    X2 = -X2
            return self.refine(X1, X2, R, perm)
        Rs = self._grid_search(X1, X2, perm, invert)
        best = self.refine(X1, X2, Rs[0], perm)
        if invert:
            inv = self.refine(X1, -X2, Rs[1], perm)
            if inv[0] < best[0]:
                best = inv
        return best

    def _align(self, pos1, pos2, perm=None, invert=True):
        """The reference's older entry point (:139-158): align without the calcCoeffs hook."""
        return self.align(pos1, pos2, perm, invert)

    def malign(self, pos1, pos2, perm=None, invert=True, calcCoeffs=None, nrot=10):
        """Try the nrot best rotations of each orientation (reference :206-235)."""
        X1, X2, perm = self._setup(pos1, pos2, perm)
        Ilmm = self._coeffs(X1, X2, perm) if calcCoeffs is None else calcCoeffs(False)
        Rs = self.findRotations(Ilmm, nrot)[0]
        best = min((self.refine(X1, X2, R, perm) for R in Rs), key=lambda x: x[0])
        if invert:
            Ilmm = self._coeffs(X1, -X2, perm) if calcCoeffs is None else calcCoeffs(True)
            Rs = self.findRotations(Ilmm, nrot)[0]
            inv = min((self.refine(X1, -X2, R, perm) for R in Rs), key=lambda x: x[0])
            if inv[0] < best[0]:
                return inv
        return best

    def __call__(self, pos1, pos2, perm=None, invert=True, calcCoeffs=None, nrot=10, niter=None):
        dist, X1, X2 = self.align(pos1, pos2, perm, invert, calcCoeffs)
        if (norm(X1 - X2, axis=1) > self.scale).sum() > len(X1) / 3:
            try:
                mdist, mX1, mX2 = self.malign(pos1, pos2, perm, invert, calcCoeffs, nrot)
                if mdist < dist:
                    return mdist, mX1, mX2
            except Exception:
                pass
        return dist, X1, X2

    # -- batched, additive API
    def _batch_scale(self, X1, X2):
        """Kernel width of a batch when the constructor left it to be computed (calcScale): the reference's
        rule (:165-166, a third of the mean nearest-neighbour separation of the two structures) averaged
        over the batch, so that one width serves every pair of the call."""
        if not self.calcScale:
            return self.scale
        n = min(len(X1), 64)  # a sample of the batch is enough for a mean
        sep = [self.averageSeparation(X1[i]) + self.averageSeparation(X2[i]) for i in range(n)]
        self.scale = float(np.mean(sep)) / 6
        return self.scale

    def align_batch(self, pos1, pos2, perm=None, invert=True, refine=True, nthreads=0):
        """P independent pairs.  Distance rule (default): one native call (fo_sph_align_pairs_full) -- GPU hot
        path, on the device the rotation by the grid-maximum Euler angles and the nearest-partner screening of
        the assignment for both orientations, on the host pool (nthreads; 0 = all cores, overlapped with the
        GPU's next chunk) the LAP where the screening failed and the Kearsley fit.  orientation="overlap": the
        numpy rule, continuous refinement on the device, then one host LAP + Kearsley per pair.
        Returns dists (P,) and the Euler angles (P, O, 3); self.last_perms / last_orient / last_rmats hold the
        rest of the result."""
        pos1 = np.asarray(pos1, float)
        pos2 = np.asarray(pos2, float)
        X1 = pos1 - pos1.mean(1, keepdims=True)
        X2 = pos2 - pos2.mean(1, keepdims=True)
        perm = self._perm(X1.shape[1], perm)
        self._batch_scale(X1, X2)
        if self.orientation == "overlap":
            # numpy rule (:178-187): the orientation with the larger refined overlap is the only one
            # that goes through the host LAP + Kearsley refinement
            Rs, ov = self._grid_search_refined(X1, X2, perm, invert)
            Rs = Rs.reshape(len(X1), -1, 3)
            if not refine:
                return None, Rs
            pick = ov.reshape(len(X1), -1).argmax(1)
            sign = np.where(pick == 1, -1.0, 1.0)[:, None, None]
            Rp = Rs[np.arange(len(X1)), pick][:, None, :]
            dists, orient, perms, rmats = _lib.host_refine_spherical(X1, sign * X2, Rp, perm, nthreads)
            self.last_orient, self.last_perms, self.last_rmats = pick, perms, rmats
            return dists, Rs
        if not refine or not hasattr(self, "_full_batch"):
            Rs = self._grid_search(X1, X2, perm, invert)
            Rs = Rs.reshape(len(X1), -1, 3)
            if not refine:
                return None, Rs
            dists, self.last_orient, self.last_perms, self.last_rmats = _lib.host_refine_spherical(
                X1, X2, Rs, perm, nthreads)
            return dists, Rs
        return self._full_batch(X1, X2, perm, invert, nthreads)


class SphericalAlign(BaseSphericalAlignment):
    """Direct (N^2 Bessel) overlap coefficients (reference :250-273)."""

    def __init__(self, scale=None, Jmax=15, perm=None, ctx=None, orientation="distance"):
        self._ctx = ctx
        if scale is not None:
            self.scale = scale
            self.calcScale = False
        else:
            self.calcScale = True
        self.orientation = orientation
        self.setJ(Jmax)
        self.perm = perm

    def calcSO3Coeffs(self, pos1, pos2):
        """I[l, m1, m2] of two (already gathered) atom sets (reference :260-273) -- on the GPU."""
        pos1 = np.atleast_2d(pos1)
        pos2 = np.atleast_2d(pos2)
        assert pos1.shape == pos2.shape
        self.ctx.set_perm([np.arange(len(pos1))], len(pos1))
        return self.ctx.sph_coeffs_direct(pos1, pos2, self.Jmax, self.scale)[0][0]

    def _coeffs(self, X1, X2, perm):
        """sum over permutation groups of calcSO3Coeffs (reference :175) in one call."""
        self.ctx.set_perm(perm, len(X1))
        return self.ctx.sph_coeffs_direct(X1, X2, self.Jmax, self.scale)[0][0]

    def _grid_search(self, X1, X2, perm, invert, want_grid=False, calcCoeffs=None):
        n = X1.shape[-2]
        self.ctx.set_perm(perm, n)
        bi, bv, fr, grid, st = self.ctx.sph_align_pairs(X1, X2, self.Jmax, self.scale, invert=invert,
                                                        want_grid=want_grid)
        self._best_idx, self._best_val, self._frac_idx, self._grid = bi, bv, fr, grid
        R = indtoEuler(fr.reshape(-1, 3), self.soft.n).reshape(fr.shape)
        return R[0] if X1.ndim == 2 else R


    def _full_batch(self, X1, X2, perm, invert, nthreads):
        self.ctx.set_perm(perm, X1.shape[-2])
        dists, self.last_orient, self.last_perms, self.last_rmats, Rs, st, self.last_nhost = \
            self.ctx.sph_align_pairs_full(X1, X2, self.Jmax, self.scale, invert=invert, nthreads=nthreads)
        return dists, Rs

    def _grid_search_refined(self, X1, X2, perm, invert):
        self.ctx.set_perm(perm, X1.shape[-2])
        bi, bv, fr, eu, ov, st = self.ctx.sph_align_pairs_refined(X1, X2, self.Jmax, self.scale, invert=invert)
        self._best_idx, self._best_val, self._frac_idx, self._grid = bi, bv, fr, None
        return (eu[0], ov[0]) if X1.ndim == 2 else (eu, ov)


class SphericalHarmonicAlign(BaseSphericalAlignment):
    """Harmonic-oscillator radial basis coefficients C_nlm (reference :276-414)."""

    def __init__(self, scale=None, harmscale=1.0, nmax=15, Jmax=15, perm=None, ctx=None,
                 orientation="distance"):
        self._ctx = ctx
        if scale is not None:
            self.scale = scale
            self.calcScale = False
        else:
            self.calcScale = True
        self.orientation = orientation
        self.harmscale = harmscale
        self.nmax = nmax
        self.setJ(Jmax)
        self.perm = perm

    def setCoeffs(self, nmax=None, Jmax=None, harmscale=None):
        if nmax is not None:
            self.nmax = nmax
        if Jmax is not None:
            self.setJ(Jmax)
        if harmscale is not None:
            self.harmscale = harmscale

    def calcHarmCoeff(self, pos):
        """C[n, l, m] of one atom set (reference :345-361) -- on the GPU."""
        pos = np.atleast_2d(pos)
        self.ctx.set_perm([np.arange(len(pos))], len(pos))
        return self.ctx.sph_harm_coeffs(pos, self.nmax, self.Jmax, self.harmscale, self.scale)[0][0, 0]

    def calcSO3Coeffs(self, pos1, pos2):
        pos1 = np.atleast_2d(pos1)
        pos2 = np.atleast_2d(pos2)
        return self._coeffs(pos1, pos2, [np.arange(len(pos1))])

    def _bank_pair(self, X1, X2, perm, invert, want_grid=False):
        self.ctx.set_perm(perm, len(X1))
        bank = self.ctx.sph_bank_create(np.stack([X1, X2]), self.nmax, self.Jmax, self.harmscale, self.scale)
        try:
            return self.ctx.sph_align_bank(bank, np.array([[0, 1]]), invert=invert, want_grid=want_grid)
        finally:
            bank.close()

    def _coeffs(self, X1, X2, perm):
        """I[l,m1,m2] = sum_groups sum_n conj(C1) C2 (reference :363-372)."""
        self.ctx.set_perm(perm, len(X1))
        C, _ = self.ctx.sph_harm_coeffs(np.stack([X1, X2]), self.nmax, self.Jmax, self.harmscale, self.scale)
        return self.calcSO3Harm(C[0], C[1])

    def calcSO3Harm(self, c1nlms, c2nlms, invert=False):
        """Contraction of host-resident coefficient arrays (reference :368-372).  Host einsum: this
        entry point exists for API compatibility; the alignment path contracts device-resident banks
        (fo_sph_align_bank)."""
        c1 = np.asarray(c1nlms)
        c2 = np.asarray(c2nlms)
        if c1.ndim == 3:
            c1, c2 = c1[None], c2[None]
        if invert:
            c2 = c2 * (-1.0) ** self.J[None, None, :, None]
        return np.einsum("gnlm,gnlo->lmo", c1.conj(), c2)

    def _grid_search(self, X1, X2, perm, invert, want_grid=False, calcCoeffs=None):
        if X1.ndim == 3:
            P, n = X1.shape[:2]
            self.ctx.set_perm(perm, n)
            bank = self.ctx.sph_bank_create(np.concatenate([X1, X2]), self.nmax, self.Jmax, self.harmscale,
                                            self.scale)
            try:
                pairs = np.stack([np.arange(P), P + np.arange(P)], axis=1)
                bi, bv, fr, avg, grid = self.ctx.sph_align_bank(bank, pairs, invert=invert)
            finally:
                bank.close()
        else:
            bi, bv, fr, avg, grid = self._bank_pair(X1, X2, perm, invert, want_grid)
        self._best_idx, self._best_val, self._frac_idx, self._grid = bi, bv, fr, grid
        R = indtoEuler(fr.reshape(-1, 3), self.soft.n).reshape(fr.shape)
        return R[0] if X1.ndim == 2 else R

    def _grid_search_refined(self, X1, X2, perm, invert):
        single = X1.ndim == 2
        A = X1[None] if single else X1
        B = X2[None] if single else X2
        P, n = A.shape[:2]
        self.ctx.set_perm(perm, n)
        bank = self.ctx.sph_bank_create(np.concatenate([A, B]), self.nmax, self.Jmax, self.harmscale, self.scale)
        try:
            pairs = np.stack([np.arange(P), P + np.arange(P)], axis=1)
            bi, bv, fr, avg, eu, ov = self.ctx.sph_align_bank_refined(bank, pairs, invert=invert)
        finally:
            bank.close()
        self._best_idx, self._best_val, self._frac_idx, self._grid = bi, bv, fr, None
        return (eu[0], ov[0]) if single else (eu, ov)

    def compareList(self, poslist, perm=None, invert=False):
        """All-vs-all average / maximum overlap (reference SphericalHarmonicAlignFortran.compareList
        :622-663, CALCOVERLAPMATRICES fastclusters.f90:853-866): coefficients once per structure in a
        device bank, then contraction + inverse SO(3) transform + max per pair."""
        coords = np.array(poslist, dtype=float)
        nlist, natoms, dim = coords.shape
        assert dim == 3
        coords -= coords.mean(1)[:, None, :]
        self.ctx.set_perm(self._perm(natoms, perm), natoms)
        bank = self.ctx.sph_bank_create(coords, self.nmax, self.Jmax, self.harmscale, self.scale)
        iu = np.triu_indices(nlist)
        pairs = np.stack(iu, axis=1)
        try:
            bi, bv, fr, avg, _ = self.ctx.sph_align_bank(bank, pairs, invert=invert)
        finally:
            bank.close()
        avgoverlap = np.zeros((nlist, nlist))
        maxoverlap = np.zeros((nlist, nlist))
        avgoverlap[iu] = avg
        maxoverlap[iu] = bv.max(1)
        avgoverlap = np.triu(avgoverlap) + np.triu(avgoverlap, 1).T
        maxoverlap = np.triu(maxoverlap) + np.triu(maxoverlap, 1).T
        da, dm = avgoverlap.diagonal(), maxoverlap.diagonal()
        return (avgoverlap, maxoverlap, avgoverlap / sqrt(da[:, None] * da[None, :]),
                maxoverlap / sqrt(dm[:, None] * dm[None, :]))

    def alignGroup(self, coords, keepCoords=False):
        """All-vs-all alignment distances (reference :397-414): coefficients once per structure in a device
        bank, contraction + iSOFT + arg-max per pair, then the native host pool on all pairs at once."""
        coords = np.asarray(coords, float)
        nl, natoms = coords.shape[:2]
        X = coords - coords.mean(1)[:, None, :]
        perm = self._perm(natoms)
        self.ctx.set_perm(perm, natoms)
        bank = self.ctx.sph_bank_create(X, self.nmax, self.Jmax, self.harmscale, self.scale)
        ii, jj = (a.ravel() for a in np.meshgrid(np.arange(nl), np.arange(nl), indexing="ij"))
        try:
            fr = self.ctx.sph_align_bank(bank, np.stack([ii, jj], axis=1), invert=True)[2]
        finally:
            bank.close()
        Rs = indtoEuler(fr.reshape(-1, 3), self.soft.n).reshape(fr.shape)
        d, orient, perms, rmats = _lib.host_refine_spherical(X[ii], X[jj], Rs, perm, 0)
        dists = d.reshape(nl, nl)
        if not keepCoords:
            return dists
        aligned = np.empty((2, nl, nl) + coords[0].shape)
        aligned[0] = X[ii].reshape((nl, nl) + coords[0].shape)
        for k in range(len(ii)):
            sg = -1.0 if orient[k] else 1.0
            aligned[1, ii[k], jj[k]] = (sg * X[jj[k]]).dot(EulerM(*Rs[k, orient[k]]))[perms[k]].dot(rmats[k].T)
        return dists, aligned
