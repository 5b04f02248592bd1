"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/fo_oracle_*.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (fastoverlap_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfo_oracle.so")

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_f64 = ctypes.c_double
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".c")]
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_next_fast_len.restype = _i64
        _lib.oracle_next_fast_len.argtypes = [_i64]
        _lib.oracle_per_csum.restype = _f64
        _lib.oracle_per_align_pairs.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _groups(perm, natoms):
    if perm is None:
        perm = [np.arange(natoms)]
    groups = [np.asarray(g, dtype=np.int32).ravel() for g in perm]
    off = np.zeros(len(groups) + 1, np.int32)
    off[1:] = np.cumsum([len(g) for g in groups])
    idx = np.concatenate(groups).astype(np.int32) if len(groups) else np.zeros(1, np.int32)
    return off, idx, len(groups)


def next_fast_len(n):
    return int(lib().oracle_next_fast_len(int(n)))


# ------------------------------------------------------------------ periodic

def per_structure_factors(pos, box, n, perm=None):
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    off, idx, ng = _groups(perm, len(pos))
    box = np.ascontiguousarray(box, dtype=np.float64)
    W = 2 * n + 1
    out = np.empty((ng, W, W, W), np.complex128)
    lib().oracle_per_structure_factors(_p(pos), _i64(len(pos)), _p(off), _i64(ng), _p(idx), _p(box),
                                       _i64(n), _p(out))
    return out


def per_cross_spectrum(C1, C2, box, n, sigma):
    C1 = np.ascontiguousarray(C1, dtype=np.complex128)
    C2 = np.ascontiguousarray(C2, dtype=np.complex128)
    box = np.ascontiguousarray(box, dtype=np.float64)
    W = 2 * n + 1
    C = np.empty((W, W, W), np.complex128)
    lib().oracle_per_cross_spectrum(_p(C1), _p(C2), _i64(C1.shape[0]), _p(box), _i64(n), _f64(sigma),
                                    _p(C))
    return C


def per_csum(C1, C2, box, n, sigma):
    C1 = np.ascontiguousarray(C1, dtype=np.complex128)
    C2 = np.ascontiguousarray(C2, dtype=np.complex128)
    box = np.ascontiguousarray(box, dtype=np.float64)
    return float(lib().oracle_per_csum(_p(C1), _p(C2), _i64(C1.shape[0]), _p(box), _i64(n),
                                       _f64(sigma)))


def per_fft_abs(C, F, want_f=False):
    C = np.ascontiguousarray(C, dtype=np.complex128)
    n = (C.shape[0] - 1) // 2
    fabs = np.empty((F, F, F), np.float64)
    f = np.empty((F, F, F), np.complex128) if want_f else None
    lib().oracle_per_fft_abs(_p(C), _i64(n), _i64(F), _p(fabs), _p(f))
    return (fabs, f) if want_f else fabs


def fft1d(x, sign=-1):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    lib().oracle_fft1d(_i64(len(x)), _p(x), _i64(1), _p(out), ctypes.c_int(sign))
    return out


def find_max(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    shape = np.array(a.shape, np.int64)
    idx = np.empty(3, np.int64)
    frac = np.empty(3, np.float64)
    lib().oracle_find_max(_p(a), _p(shape), _p(idx), _p(frac))
    return idx, frac


def per_align_pairs(posA, posB, box, n, F, sigma, perm=None, nthreads=0, want_grid=False):
    """Whole periodic hot path for P pairs; returns (best_idx, best_val, frac_idx[, grids], threads)."""
    posA = np.ascontiguousarray(posA, dtype=np.float64)
    posB = np.ascontiguousarray(posB, dtype=np.float64)
    if posA.ndim == 2:
        posA, posB = posA[None], posB[None]
    P, N, _ = posA.shape
    off, idx, ng = _groups(perm, N)
    box = np.ascontiguousarray(box, dtype=np.float64)
    bi = np.empty((P, 3), np.int64)
    bv = np.empty(P, np.float64)
    fr = np.empty((P, 3), np.float64)
    if want_grid:
        grids = np.empty((P, F, F, F), np.float64)
        for p in range(P):
            lib().oracle_per_align_pair(_p(posA[p]), _p(posB[p]), _i64(N), _p(off), _i64(ng), _p(idx),
                                        _p(box), _i64(n), _i64(F), _f64(sigma), _p(bi[p]),
                                        _p(bv[p:p + 1]), _p(fr[p]), _p(grids[p]))
        return bi, bv, fr, grids, 1
    used = lib().oracle_per_align_pairs(_p(posA), _p(posB), _i64(P), _i64(N), _p(off), _i64(ng),
                                        _p(idx), _p(box), _i64(n), _i64(F), _f64(sigma), _p(bi),
                                        _p(bv), _p(fr), ctypes.c_int(int(nthreads)))
    return bi, bv, fr, None, int(used)


# ------------------------------------------------------------------ spherical

def sph_ylm(pos, L):
    """Y[l, m] (negative m wrapped) of one point -> (L+1, 2L+1) complex, and r."""
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(3)
    Y = np.empty((L + 1, 2 * L + 1), np.complex128)
    lib().oracle_sph_ylm.restype = _f64
    r = lib().oracle_sph_ylm(_p(pos), _i64(L), _p(Y))
    return Y, float(r)


def sphi_scaled(L, x):
    out = np.empty(L + 1, np.float64)
    lib().oracle_sphi_scaled(_i64(L), _f64(x), _p(out))
    return out


def sph_coeffs_direct(posA, posB, Jmax, sigma, perm=None):
    posA = np.ascontiguousarray(posA, dtype=np.float64).reshape(-1, 3)
    posB = np.ascontiguousarray(posB, dtype=np.float64).reshape(-1, 3)
    off, idx, ng = _groups(perm, len(posA))
    L = int(Jmax)
    I = np.empty((L + 1, 2 * L + 1, 2 * L + 1), np.complex128)
    lib().oracle_sph_coeffs_direct(_p(posA), _p(posB), _i64(len(posA)), _p(off), _i64(ng), _p(idx),
                                   _i64(L), _f64(sigma), _p(I))
    return I


def soft_weights(B):
    w = np.empty(2 * B, np.float64)
    lib().oracle_soft_weights(_i64(B), _p(w))
    return w


def wigner_table(B):
    Ds = np.empty((B, 2 * B - 1, 2 * B - 1, 2 * B), np.float64)
    lib().oracle_wigner_table(_i64(B), _p(Ds))
    return Ds


def isoft(Ilmm, Jmax, want_complex=False):
    B = int(Jmax) + 1
    I = np.ascontiguousarray(Ilmm, dtype=np.complex128)
    assert I.shape == (B, 2 * B - 1, 2 * B - 1)
    out = np.empty((2 * B,) * 3, np.float64)
    outc = np.empty((2 * B,) * 3, np.complex128) if want_complex else None
    lib().oracle_isoft(_p(I), _i64(B), None, _p(out), _p(outc))
    return outc if want_complex else out


def sph_harm_radial(nmax, L, r, sigma, r0, kind="exact"):
    if kind == "exact":
        out = np.empty((nmax + 1, L + 1), np.float64)
        lib().oracle_sph_harm_radial_exact(_i64(nmax), _i64(L), _f64(r), _f64(sigma), _f64(r0), _p(out))
        return out
    Lf = L + 2 * nmax
    out = np.empty((nmax + 1, Lf + 1), np.float64)
    lib().oracle_sph_harm_radial_fortran(_i64(nmax), _i64(Lf), _f64(r), _f64(sigma), _f64(r0), _p(out))
    return out[:, :L + 1].copy()


def sph_harm_coeffs(pos, nmax, Jmax, harmscale, sigma, idx=None, kind="exact"):
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    idx = np.arange(len(pos), dtype=np.int32) if idx is None else np.asarray(idx, dtype=np.int32)
    L = int(Jmax)
    C = np.empty((nmax + 1, L + 1, 2 * L + 1), np.complex128)
    lib().oracle_sph_harm_coeffs(_p(pos), _i64(len(pos)), _p(idx), _i64(len(idx)), _i64(nmax), _i64(L),
                                 _f64(harmscale), _f64(sigma), ctypes.c_int(1 if kind == "fortran" else 0),
                                 _p(C))
    return C


def sph_dot_harm(C1, C2, invert=False):
    C1 = np.ascontiguousarray(C1, dtype=np.complex128)
    C2 = np.ascontiguousarray(C2, dtype=np.complex128)
    if C1.ndim == 3:
        C1, C2 = C1[None], C2[None]
    ng, n1, l1, nj = C1.shape
    I = np.empty((l1, nj, nj), np.complex128)
    lib().oracle_sph_dot_harm(_p(C1), _p(C2), _i64(ng), _i64(n1 - 1), _i64(l1 - 1),
                              ctypes.c_int(int(bool(invert))), _p(I))
    return I


def sph_align_pairs(posA, posB, Jmax, sigma, invert=True, perm=None, nthreads=0, want_grid=False):
    """Whole spherical hot path (direct coefficients) for P pairs of centred structures."""
    posA = np.ascontiguousarray(posA, dtype=np.float64)
    posB = np.ascontiguousarray(posB, dtype=np.float64)
    if posA.ndim == 2:
        posA, posB = posA[None], posB[None]
    P, N, _ = posA.shape
    off, idx, ng = _groups(perm, N)
    O = 2 if invert else 1
    L = int(Jmax)
    nk = 2 * (L + 1)
    bi = np.empty((P, O, 3), np.int64)
    bv = np.empty((P, O), np.float64)
    fr = np.empty((P, O, 3), np.float64)
    if want_grid:
        grids = np.empty((P, O, nk, nk, nk), np.float64)
        for p in range(P):
            lib().oracle_sph_align_pair(_p(posA[p]), _p(posB[p]), _i64(N), _p(off), _i64(ng), _p(idx),
                                        _i64(L), _f64(sigma), ctypes.c_int(int(bool(invert))), None,
                                        _p(bi[p]), _p(bv[p]), _p(fr[p]), _p(grids[p]))
        return bi, bv, fr, grids, 1
    lib().oracle_sph_align_pairs.restype = ctypes.c_int
    used = lib().oracle_sph_align_pairs(_p(posA), _p(posB), _i64(P), _i64(N), _p(off), _i64(ng), _p(idx),
                                        _i64(L), _f64(sigma), ctypes.c_int(int(bool(invert))), _p(bi),
                                        _p(bv), _p(fr), ctypes.c_int(int(nthreads)))
    return bi, bv, fr, None, int(used)


# -- continuous rotation refinement (numpy restatement; small, per-rotation) ---------------------
def sph_wigner_at(rot, Jmax):
    """Un-weighted Wigner D^l_{m1 m2}(a, b, g) and its gradient on the reference's
    [l][m1 wrap][m2 wrap] layout: BaseSphericalAlignment.calcWignerMatrices(rot)
    (sphericalAlignment.py:67-91; d^l by Jacobi polynomials, eval_grad_jacobi utils.py)."""
    from scipy.special import eval_jacobi, gammaln
    a, b, y = rot
    Js, m1s, m2s = map(np.array, zip(*[(l, m1, m2) for l in range(Jmax + 1) for m1 in range(-l, l + 1)
                                       for m2 in range(-l, l + 1)]))
    mu, nu = abs(m1s - m2s), abs(m1s + m2s)
    s = Js - (mu + nu) // 2
    xi = np.where(m2s < m1s, (-1.0) ** (m1s - m2s), 1.0)
    factor = np.exp(0.5 * (gammaln(s + 1) + gammaln(s + mu + nu + 1) - gammaln(s + mu + 1) -
                           gammaln(s + nu + 1))) * xi
    sb2, cb2, cb, sb = np.sin(b / 2), np.cos(b / 2), np.cos(b), np.sin(b)
    jac = eval_jacobi(s, mu, nu, cb)
    d = factor * jac * sb2 ** mu * cb2 ** nu
    gjac = np.where(s > 0, eval_jacobi(np.maximum(s - 1, 0), mu + 1, nu + 1, cb), 0.0) * 0.5 * (mu + nu + s + 1.0) * -sb
    with np.errstate(divide="ignore", invalid="ignore"):
        gd = (factor * gjac * sb2 ** mu * cb2 ** nu +
              np.where(mu > 0, factor * jac * sb2 ** (mu - 1.0) * cb2 ** (nu + 1.0) * mu / 2, 0.0) -
              np.where(nu > 0, factor * jac * sb2 ** (mu + 1.0) * cb2 ** (nu - 1.0) * nu / 2, 0.0))
    Ds = np.zeros((Jmax + 1, 2 * Jmax + 1, 2 * Jmax + 1), np.complex128)
    grad = np.zeros((3,) + Ds.shape, np.complex128)
    ph = np.exp(-1j * m1s * a) * np.exp(-1j * m2s * y)
    Ds[Js, m1s, m2s] = ph * d
    grad[0, Js, m1s, m2s] = -1j * m1s * Ds[Js, m1s, m2s]
    grad[1, Js, m1s, m2s] = ph * gd
    grad[2, Js, m1s, m2s] = -1j * m2s * Ds[Js, m1s, m2s]
    return Ds, grad


def sph_energy_gradient(rot, Ilmm_conj, Jmax):
    """getEnergyGradient (sphericalAlignment.py:93-96): E = -Re sum(Ilmm * D), dE/d(a,b,g)."""
    D, gD = sph_wigner_at(rot, Jmax)
    return -(Ilmm_conj * D).real.sum(), -(Ilmm_conj[None] * gD).real.sum((1, 2, 3))


def sph_max_overlap(rot0, Ilmm, Jmax):
    """findRotation's refinement (sphericalAlignment.py:98-103,190-194): scipy L-BFGS-B on
    getEnergyGradient with conj(Ilmm).  Returns (R, -res.fun)."""
    from scipy.optimize import minimize
    res = minimize(sph_energy_gradient, np.asarray(rot0, float), jac=True, args=(np.conj(Ilmm), Jmax),
                   method="L-BFGS-B")
    return res.x, -res.fun
