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
