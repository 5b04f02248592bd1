"""ctypes binding of libfastoverlap_b200.so (include/fastoverlap_b200.h).

This is the ONLY compute back end of the package: there is no CPU fallback.  Importing the
package works without a GPU (so the host-side logic can be tested), but creating a Context
raises FastOverlapError when the library or a CUDA device is missing.
"""
import ctypes
import os
import threading

import numpy as np

from . import build as _build

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_void_p = ctypes.c_void_p


class FastOverlapError(RuntimeError):
    pass


class PerParams(ctypes.Structure):
    """struct fo_per_params"""
    _fields_ = [("natoms", ctypes.c_int64),
                ("box", ctypes.c_double * 3),
                ("nwave", ctypes.c_int64),
                ("nfspace", ctypes.c_int64),
                ("sigma", ctypes.c_double)]


# name -> (restype, argtypes); every symbol include/fastoverlap_b200.h declares
SIGNATURES = {
    "fo_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_void_p)]),
    "fo_destroy": (None, [c_void_p]),
    "fo_last_error": (ctypes.c_char_p, [c_void_p]),
    "fo_set_stream": (ctypes.c_int, [c_void_p, c_void_p]),
    "fo_reset_stream": (ctypes.c_int, [c_void_p]),
    "fo_sync": (ctypes.c_int, [c_void_p]),
    "fo_profile_begin": (ctypes.c_int, [c_void_p]),
    "fo_profile_end": (ctypes.c_int, [c_void_p, c_f64p, c_i64p]),
    "fo_device_info": (ctypes.c_int, [c_void_p, c_i64p]),
    "fo_launch_count": (ctypes.c_int64, [c_void_p]),
    "fo_set_perm": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64]),
    "fo_next_fast_len": (ctypes.c_int64, [ctypes.c_int64]),
    "fo_measure_fp64_peak": (ctypes.c_int, [c_void_p, c_f64p]),
    "fo_measure_fp64_tensor_peak": (ctypes.c_int, [c_void_p, c_f64p]),
    "fo_set_option": (ctypes.c_int, [c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "fo_per_defaults": (ctypes.c_int, [ctypes.c_int64, c_f64p, c_f64p, c_i64p, c_i64p]),
    "fo_per_structure_factors": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p,
                                                ctypes.c_int64, c_void_p]),
    "fo_per_align_pairs": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p, c_void_p,
                                          ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "fo_per_align_pairs_dev": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p,
                                              c_void_p, ctypes.c_int64, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p]),
    "fo_per_align_coeffs": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p, c_void_p,
                                           ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p]),
    "fo_per_bank_create": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p,
                                          ctypes.c_int64, ctypes.POINTER(c_void_p)]),
    "fo_per_align_bank": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p, c_void_p,
                                         ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p]),
    "fo_per_align_bank_ops": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p, c_void_p, c_void_p,
                                             ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p]),
    "fo_bank_destroy": (None, [c_void_p, c_void_p]),
    "fo_sph_ylm": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_void_p,
                                  c_void_p, c_void_p]),
    "fo_sph_isoft_argmax": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fo_sph_isoft": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p,
                                    c_void_p]),
    "fo_sph_coeffs_direct": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                            ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                            c_void_p, c_void_p]),
    "fo_sph_align_pairs": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                          ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "fo_sph_align_pairs_dev": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                              ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                              ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p]),
    "fo_sph_harm_coeffs": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                          ctypes.c_double, c_void_p, c_void_p]),
    "fo_sph_bank_create": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                          ctypes.c_double, ctypes.POINTER(c_void_p)]),
    "fo_sph_align_bank": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                         ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p]),
    "fo_sph_refine_rotations": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p,
                                               c_void_p, c_void_p, c_void_p]),
    "fo_sph_overlap_gradient": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p,
                                               c_void_p, c_void_p, c_void_p]),
    "fo_sph_align_pairs_refined": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                                  ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                                  ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, c_void_p]),
    "fo_sph_align_pairs_refined_dev": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                                      ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                      c_void_p, c_void_p]),
    "fo_sph_align_bank_refined": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                                 ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_void_p]),
    "fo_sph_wigner_table": (ctypes.c_int, [c_void_p, ctypes.c_int64, c_void_p]),
    "fo_grid_find_peaks": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, c_i64p, ctypes.c_int64,
                                          ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "fo_grid_find_peaks_dev": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, c_i64p, ctypes.c_int64,
                                              ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fo_sph_isoft_peaks": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fo_per_align_pairs_peaks": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p, c_void_p,
                                                ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_void_p,
                                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "fo_host_refine_periodic": (ctypes.c_int, [ctypes.POINTER(PerParams), c_void_p, ctypes.c_int64, c_void_p,
                                               c_void_p, c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int,
                                               ctypes.c_int, c_void_p, c_void_p, c_void_p]),
    "fo_host_refine_periodic_subset": (ctypes.c_int, [ctypes.POINTER(PerParams), c_void_p, ctypes.c_int64, c_void_p,
                                                      c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                                      ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p]),
    "fo_host_refine_spherical_hint": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p,
                                                     ctypes.c_int64, c_void_p, c_void_p, ctypes.c_int, c_void_p,
                                                     c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p,
                                                     c_void_p]),
    "fo_per_align_pairs_full": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p, c_void_p,
                                               ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_i64p]),
    "fo_sph_align_pairs_full": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_i64p]),
    "fo_per_align_pairs_full_dev": (ctypes.c_int, [c_void_p, ctypes.POINTER(PerParams), c_void_p, c_void_p,
                                                   ctypes.c_int64, ctypes.c_int, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fo_sph_align_pairs_screen_dev": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64,
                                                     ctypes.c_int64, ctypes.c_double, ctypes.c_int, c_void_p,
                                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fo_host_best_permutation": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64,
                                                c_void_p, c_void_p, c_void_p]),
    "fo_host_kearsley": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, c_f64p, c_void_p]),
    "fo_host_refine_counters": (None, [c_void_p, ctypes.c_int]),
    "fo_host_lap_isa": (ctypes.c_int, [ctypes.c_int]),
    "fo_host_refine_spherical": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p,
                                                ctypes.c_int64, c_void_p, c_void_p, ctypes.c_int, ctypes.c_int,
                                                c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None
_lib_lock = threading.Lock()


def library_path():
    """In-tree build, or the library named by FASTOVERLAP_B200_LIB (A/B runs of two builds)."""
    return os.environ.get("FASTOVERLAP_B200_LIB") or _build.lib_path()


def load_library():
    """Load (building if absent and nvcc is available) the shared library. Fails loudly."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            try:
                _build.build()
            except Exception as exc:  # no nvcc, compile error ...
                raise FastOverlapError(
                    "libfastoverlap_b200.so is missing and could not be built (%s). "
                    "Run `python -m fastoverlap_b200.build`. There is no CPU fallback." % exc)
        try:
            lib = ctypes.CDLL(path)
        except OSError as exc:
            raise FastOverlapError("cannot load %s: %s (no CPU fallback)" % (path, exc))
        for name, (restype, argtypes) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                raise FastOverlapError("%s does not export %s" % (path, name))
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _group_arrays(perm, natoms):
    groups = [np.asarray(g, dtype=np.int64).ravel() for g in (perm if perm is not None else [np.arange(natoms)])]
    off = np.zeros(len(groups) + 1, np.int32)
    off[1:] = np.cumsum([len(g) for g in groups])
    idx = (np.concatenate(groups) if groups else np.zeros(0, np.int64)).astype(np.int32)
    if idx.size == 0:
        idx = np.zeros(1, np.int32)
    return off, idx, len(groups)


def host_refine_periodic(params, perm, posA, posB, frac_idx, niter=10, nthreads=0):
    """Native host refinement of P periodic pairs (fo_host_refine_periodic; needs no GPU).
    Returns (dist (P,), perm (P,N), disp (P,3))."""
    lib = load_library()
    N = params.natoms
    posA = _f64(posA).reshape(-1, N, 3)
    posB = _f64(posB).reshape(-1, N, 3)
    frac = _f64(frac_idx).reshape(-1, 3)
    P = posA.shape[0]
    off, idx, ng = _group_arrays(perm, N)
    dist = np.empty(P)
    pm = np.empty((P, N), np.int32)
    disp = np.empty((P, 3))
    rc = lib.fo_host_refine_periodic(ctypes.byref(params), _ptr(off), ng, _ptr(idx), _ptr(posA), _ptr(posB),
                                     _ptr(frac), P, int(niter), int(nthreads), _ptr(dist), _ptr(pm), _ptr(disp))
    if rc != 0:
        raise FastOverlapError("fo_host_refine_periodic failed (%d)" % rc)
    return dist, pm, disp


def host_best_permutation(posA, posB, perm=None, box=None):
    """perm (N,) such that posB[perm] best matches posA group by group (fo_host_best_permutation; no GPU)."""
    lib = load_library()
    posA = _f64(posA).reshape(-1, 3)
    posB = _f64(posB).reshape(-1, 3)
    N = posA.shape[0]
    off, idx, ng = _group_arrays(perm, N)
    out = np.empty(N, np.int32)
    b = None if box is None else _f64(box).reshape(3)
    rc = lib.fo_host_best_permutation(_ptr(posA), _ptr(posB), N, _ptr(off), ng, _ptr(idx), _ptr(b), _ptr(out))
    if rc != 0:
        raise FastOverlapError("fo_host_best_permutation failed (%d): invalid permutation groups?" % rc)
    return out


def host_lap_isa(isa=-1):
    """Test hook (fo_host_lap_isa): select / query the instruction set of the host assignment kernels."""
    return int(load_library().fo_host_lap_isa(int(isa)))


def host_kearsley(x1, x2):
    """(distance, rotation matrix (3,3)) of the Kearsley fit of x2 onto x1, both re-centred (fo_host_kearsley)."""
    lib = load_library()
    x1 = _f64(x1).reshape(-1, 3)
    x2 = _f64(x2).reshape(-1, 3)
    if x1.shape != x2.shape:
        raise ValueError("dimension of arrays does not match")
    d = ctypes.c_double()
    R = np.empty((3, 3))
    rc = lib.fo_host_kearsley(_ptr(x1), _ptr(x2), x1.shape[0], ctypes.byref(d), _ptr(R))
    if rc != 0:
        raise FastOverlapError("fo_host_kearsley failed (%d)" % rc)
    return d.value, R


def host_refine_counters(reset=False):
    """(LAP solves, screened assignments, skipped repeat solves) of host_refine_periodic since load / the
    last reset (fo_host_refine_counters)."""
    out = np.zeros(3, np.int64)
    load_library().fo_host_refine_counters(_ptr(out), int(bool(reset)))
    return tuple(int(v) for v in out)


def host_refine_spherical(posA, posB, euler, perm=None, nthreads=0):
    """Native host refinement of P centred cluster pairs (fo_host_refine_spherical; needs no GPU).
    euler (P, O, 3).  Returns (dist (P,), orient (P,), perm (P,N), rmat (P,3,3))."""
    lib = load_library()
    posA = _f64(posA)
    posB = _f64(posB)
    if posA.ndim == 2:
        posA, posB = posA[None], posB[None]
    P, N, _ = posA.shape
    euler = _f64(euler).reshape(P, -1, 3)
    O = euler.shape[1]
    off, idx, ng = _group_arrays(perm, N)
    dist = np.empty(P)
    orient = np.empty(P, np.int32)
    pm = np.empty((P, N), np.int32)
    rmat = np.empty((P, 3, 3))
    rc = lib.fo_host_refine_spherical(_ptr(posA), _ptr(posB), P, N, _ptr(off), ng, _ptr(idx), _ptr(euler), O,
                                      int(nthreads), _ptr(dist), _ptr(orient), _ptr(pm), _ptr(rmat))
    if rc != 0:
        raise FastOverlapError("fo_host_refine_spherical failed (%d)" % rc)
    return dist, orient, pm, rmat


class Context(object):
    """One fo_ctx: one per (host thread, GPU)."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = c_void_p()
        rc = self._lib.fo_create(int(device), ctypes.byref(h))
        if rc != 0:
            msg = self._lib.fo_last_error(None)
            raise FastOverlapError("fo_create(device=%d) failed: %s" %
                                   (device, msg.decode() if msg else rc))
        self._h = h
        self.device = int(device)
        self._perm_key = None

    # -- plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._lib.fo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.fo_last_error(self._h)
            raise FastOverlapError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    def sync(self):
        self._check(self._lib.fo_sync(self._h), "fo_sync")

    def set_stream(self, cuda_stream):
        """Run on an external cudaStream_t handle (int; 0 = legacy default stream)."""
        self._check(self._lib.fo_set_stream(self._h, c_void_p(cuda_stream or 0)), "fo_set_stream")

    def reset_stream(self):
        self._check(self._lib.fo_reset_stream(self._h), "fo_reset_stream")

    PROF_KINDS = ("per_sf", "per_xf", "sph_coef", "sph_harm", "sph_dot", "sph_isoft", "peaks", "sph_refine",
                  "assign")

    def profile_begin(self):
        self._check(self._lib.fo_profile_begin(self._h), "fo_profile_begin")

    def profile_end(self):
        """-> {kernel class: (total ms, launches)} measured with CUDA events on the ctx stream."""
        ms = np.zeros(len(self.PROF_KINDS), np.float64)
        cnt = np.zeros(len(self.PROF_KINDS), np.int64)
        self._check(self._lib.fo_profile_end(self._h, ms.ctypes.data_as(c_f64p),
                                             cnt.ctypes.data_as(c_i64p)), "fo_profile_end")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROF_KINDS) if cnt[i]}

    def device_info(self):
        out = np.zeros(4, np.int64)
        self._check(self._lib.fo_device_info(self._h, out.ctypes.data_as(c_i64p)), "fo_device_info")
        return {"sm_count": int(out[0]), "l2_bytes": int(out[1]), "smem_optin": int(out[2]),
                "cc": int(out[3])}

    def launch_count(self):
        return int(self._lib.fo_launch_count(self._h))

    def measure_fp64_peak(self):
        v = ctypes.c_double()
        self._check(self._lib.fo_measure_fp64_peak(self._h, ctypes.byref(v)), "fo_measure_fp64_peak")
        return v.value

    def measure_fp64_tensor_peak(self):
        v = ctypes.c_double()
        self._check(self._lib.fo_measure_fp64_tensor_peak(self._h, ctypes.byref(v)),
                    "fo_measure_fp64_tensor_peak")
        return v.value

    def set_option(self, name, value):
        self._check(self._lib.fo_set_option(self._h, name.encode(), int(value)), "fo_set_option")

    def set_perm(self, perm, natoms):
        """perm: sequence of index arrays (0-based), as the reference's `perm`/`permlist`."""
        groups = [np.asarray(p, dtype=np.int64).ravel() for p in perm]
        key = (int(natoms), tuple(g.tobytes() for g in groups))
        if key == self._perm_key:
            return
        off = np.zeros(len(groups) + 1, np.int32)
        off[1:] = np.cumsum([len(g) for g in groups])
        idx = (np.concatenate(groups) if groups else np.zeros(0, np.int64)).astype(np.int32)
        if idx.size == 0:
            idx = np.zeros(1, np.int32)
        self._check(self._lib.fo_set_perm(self._h, _ptr(off), len(groups), _ptr(idx), int(natoms)),
                    "fo_set_perm")
        self._perm_key = key

    def _ensure_perm(self, natoms):
        """Default permutation (one group of all atoms) unless groups were set for this natoms."""
        if self._perm_key is None or self._perm_key[0] != int(natoms):
            self.set_perm([np.arange(int(natoms))], natoms)

    # -- periodic
    @staticmethod
    def per_params(natoms, box, nwave, nfspace, sigma):
        p = PerParams()
        p.natoms = int(natoms)
        b = np.asarray(box, dtype=float).ravel()
        p.box[0], p.box[1], p.box[2] = float(b[0]), float(b[1]), float(b[2])
        p.nwave = int(nwave)
        p.nfspace = int(nfspace)
        p.sigma = float(sigma)
        return p

    def per_structure_factors(self, params, pos):
        self._ensure_perm(params.natoms)
        pos = _f64(pos).reshape(-1, params.natoms, 3)
        S = pos.shape[0]
        W = 2 * params.nwave + 1
        ng = len(self._perm_key[1]) if self._perm_key else 1
        out = np.empty((S, ng, W, W, W), np.complex128)
        self._check(self._lib.fo_per_structure_factors(self._h, ctypes.byref(params), _ptr(pos), S,
                                                       _ptr(out)), "fo_per_structure_factors")
        return out

    def _per_outputs(self, P, F, want_grid, lead=()):
        best_idx = np.empty(lead + (P, 3), np.int64)
        best_val = np.empty(lead + (P,), np.float64)
        frac = np.empty(lead + (P, 3), np.float64)
        grid = np.empty((P, F, F, F), np.float64) if want_grid else None
        status = np.zeros(P, np.int32)
        return best_idx, best_val, frac, grid, status

    def per_align_pairs(self, params, posA, posB, want_grid=False):
        self._ensure_perm(params.natoms)
        posA = _f64(posA).reshape(-1, params.natoms, 3)
        posB = _f64(posB).reshape(-1, params.natoms, 3)
        if posA.shape != posB.shape:
            raise ValueError("posA and posB must have the same shape")
        P = posA.shape[0]
        bi, bv, fr, grid, st = self._per_outputs(P, params.nfspace, want_grid)
        self._check(self._lib.fo_per_align_pairs(self._h, ctypes.byref(params), _ptr(posA),
                                                 _ptr(posB), P, _ptr(bi), _ptr(bv), _ptr(fr),
                                                 _ptr(grid), _ptr(st)), "fo_per_align_pairs")
        return bi, bv, fr, grid, st

    def per_align_pairs_dev(self, params, d_posA, d_posB, P, d_best_idx, d_best_val, d_frac,
                            d_grid=0, d_status=0):
        """All arguments are raw device addresses (ints); asynchronous on the ctx stream."""
        self._ensure_perm(params.natoms)
        self._check(self._lib.fo_per_align_pairs_dev(
            self._h, ctypes.byref(params), c_void_p(d_posA), c_void_p(d_posB), int(P),
            c_void_p(d_best_idx), c_void_p(d_best_val), c_void_p(d_frac), c_void_p(d_grid or 0),
            c_void_p(d_status or 0)), "fo_per_align_pairs_dev")

    def per_align_pairs_full_dev(self, params, d_posA, d_posB, P, d_dist, d_perm, d_disp, d_flag, d_best_idx,
                                 d_best_val, d_frac, niter=10, d_status=0):
        """Raw device addresses; hot path + screening / displacement loop on the device, asynchronous."""
        self._ensure_perm(params.natoms)
        self._check(self._lib.fo_per_align_pairs_full_dev(
            self._h, ctypes.byref(params), c_void_p(d_posA), c_void_p(d_posB), int(P), int(niter), c_void_p(d_dist),
            c_void_p(d_perm or 0), c_void_p(d_disp), c_void_p(d_flag), c_void_p(d_best_idx), c_void_p(d_best_val),
            c_void_p(d_frac), c_void_p(d_status or 0)), "fo_per_align_pairs_full_dev")

    def per_align_pairs_full(self, params, posA, posB, niter=10, nthreads=0, want_perm=True, out=None):
        """Hot path + device screening of the assignment + host LAP pool for the flagged pairs
        (fo_per_align_pairs_full).  Returns (dist (P,), perm (P,N)|None, disp (P,3), frac_idx (P,3),
        status (P,), nhost).  out: a previous return tuple of the same batch size whose arrays are written again
        (a caller that aligns batch after batch keeps its result buffers: fresh 67 MB arrays per call cost page
        faults, which eight processes on one host pay dearly)."""
        self._ensure_perm(params.natoms)
        posA = _f64(posA).reshape(-1, params.natoms, 3)
        posB = _f64(posB).reshape(-1, params.natoms, 3)
        if posA.shape != posB.shape:
            raise ValueError("posA and posB must have the same shape")
        P = posA.shape[0]
        if out is not None and out[0].shape == (P,) and (out[1] is not None) == bool(want_perm):
            dist, perm, disp, frac, st = out[:5]
            st[:] = 0
        else:
            dist = np.empty(P)
            perm = np.empty((P, params.natoms), np.int32) if want_perm else None
            disp = np.empty((P, 3))
            frac = np.empty((P, 3))
            st = np.zeros(P, np.int32)
        nhost = ctypes.c_int64(0)
        self._check(self._lib.fo_per_align_pairs_full(self._h, ctypes.byref(params), _ptr(posA), _ptr(posB), P,
                                                      int(niter), int(nthreads), _ptr(dist), _ptr(perm),
                                                      _ptr(disp), _ptr(frac), _ptr(st), ctypes.byref(nhost)),
                    "fo_per_align_pairs_full")
        return dist, perm, disp, frac, st, int(nhost.value)

    def per_align_coeffs(self, params, CA, CB, want_grid=False):
        self._ensure_perm(params.natoms)
        W = 2 * params.nwave + 1
        CA = np.ascontiguousarray(CA, dtype=np.complex128)
        CB = np.ascontiguousarray(CB, dtype=np.complex128)
        ng = len(self._perm_key[1]) if self._perm_key else 1
        CA = CA.reshape(-1, ng, W, W, W)
        CB = CB.reshape(-1, ng, W, W, W)
        P = CA.shape[0]
        bi, bv, fr, grid, st = self._per_outputs(P, params.nfspace, want_grid)
        self._check(self._lib.fo_per_align_coeffs(self._h, ctypes.byref(params), _ptr(CA), _ptr(CB),
                                                  P, _ptr(bi), _ptr(bv), _ptr(fr), _ptr(grid),
                                                  _ptr(st)), "fo_per_align_coeffs")
        return bi, bv, fr, grid, st

    def per_bank_create(self, params, pos):
        self._ensure_perm(params.natoms)
        pos = _f64(pos).reshape(-1, params.natoms, 3)
        h = c_void_p()
        self._check(self._lib.fo_per_bank_create(self._h, ctypes.byref(params), _ptr(pos),
                                                 pos.shape[0], ctypes.byref(h)),
                    "fo_per_bank_create")
        return Bank(self, h, pos.shape[0])

    def per_align_bank(self, params, bank, pairs, want_grid=False):
        pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        P = pairs.shape[0]
        bi, bv, fr, grid, st = self._per_outputs(P, params.nfspace, want_grid)
        self._check(self._lib.fo_per_align_bank(self._h, ctypes.byref(params), bank._h, _ptr(pairs),
                                                P, _ptr(bi), _ptr(bv), _ptr(fr), _ptr(grid),
                                                _ptr(st)), "fo_per_align_bank")
        return bi, bv, fr, grid, st

    def per_align_bank_ops(self, params, bank, pairs, ops, want_grid=False):
        """per_align_bank with a signed permutation matrix per pair (ops: (P, 3, 3), cubic box): structure B of
        pair i is taken as ops[i] @ B -- an index permutation of its bank entry, no new structure factors."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        P = pairs.shape[0]
        ops = np.asarray(ops, float).reshape(P, 3, 3)
        codes = np.zeros(P, np.int32)
        for i in range(3):  # (R r)_i = s_i r_{p_i}
            pi = np.argmax(np.abs(ops[:, i, :]), axis=1)
            si = ops[np.arange(P), i, pi]
            if not np.allclose(np.abs(ops[:, i, :]).sum(1), 1.0) or not np.allclose(np.abs(si), 1.0):
                raise ValueError("ops must be signed permutation matrices")
            codes |= (pi.astype(np.int32) << (2 * i)) | ((si < 0).astype(np.int32) << (6 + i))
        bi, bv, fr, grid, st = self._per_outputs(P, params.nfspace, want_grid)
        self._check(self._lib.fo_per_align_bank_ops(self._h, ctypes.byref(params), bank._h, _ptr(pairs), _ptr(codes),
                                                    P, _ptr(bi), _ptr(bv), _ptr(fr), _ptr(grid), _ptr(st)),
                    "fo_per_align_bank_ops")
        return bi, bv, fr, grid, st

    # -- spherical
    def sph_wigner_table(self, Jmax):
        B = Jmax + 1
        out = np.empty((B, 2 * B - 1, 2 * B - 1, 2 * B), np.float64)
        self._check(self._lib.fo_sph_wigner_table(self._h, int(Jmax), _ptr(out)),
                    "fo_sph_wigner_table")
        return out

    def sph_ylm(self, pos, Jmax):
        """Y[s, l, m (wrapped), atom] and r[s, atom] of S structures (sphHarm layout)."""
        pos = _f64(pos)
        if pos.ndim == 2:
            pos = pos[None]
        S, N, _ = pos.shape
        L = int(Jmax)
        Y = np.empty((S, L + 1, 2 * L + 1, N), np.complex128)
        r = np.empty((S, N), np.float64)
        st = np.zeros(S, np.int32)
        self._check(self._lib.fo_sph_ylm(self._h, _ptr(pos), S, N, L, _ptr(Y), _ptr(r), _ptr(st)), "fo_sph_ylm")
        return Y, r, st

    def _sph_outputs(self, P, Jmax, invert, want_grid):
        O = 2 if invert else 1
        n = 2 * (Jmax + 1)
        bi = np.empty((P, O, 3), np.int64)
        bv = np.empty((P, O), np.float64)
        fr = np.empty((P, O, 3), np.float64)
        grid = np.empty((P, O, n, n, n), np.float64) if want_grid else None
        return bi, bv, fr, grid

    def sph_isoft_argmax(self, Ilmm, Jmax, invert=False, want_grid=False):
        L = int(Jmax)
        Ilmm = np.ascontiguousarray(Ilmm, dtype=np.complex128).reshape(-1, L + 1, 2 * L + 1, 2 * L + 1)
        P = Ilmm.shape[0]
        bi, bv, fr, grid = self._sph_outputs(P, L, invert, want_grid)
        self._check(self._lib.fo_sph_isoft_argmax(self._h, _ptr(Ilmm), P, L, int(bool(invert)),
                                                  _ptr(bi), _ptr(bv), _ptr(fr), _ptr(grid)),
                    "fo_sph_isoft_argmax")
        return bi, bv, fr, grid

    def sph_isoft(self, Ilmm, Jmax, want_imag=True):
        """Inverse SO(3) transform of arbitrary coefficients -> complex (P, 2B, 2B, 2B) grid."""
        L = int(Jmax)
        Ilmm = np.ascontiguousarray(Ilmm, dtype=np.complex128).reshape(-1, L + 1, 2 * L + 1, 2 * L + 1)
        P = Ilmm.shape[0]
        n = 2 * (L + 1)
        re = np.empty((P, n, n, n), np.float64)
        im = np.empty((P, n, n, n), np.float64) if want_imag else None
        self._check(self._lib.fo_sph_isoft(self._h, _ptr(Ilmm), P, L, _ptr(re), _ptr(im)), "fo_sph_isoft")
        return re + 1j * im if want_imag else re

    # -- top-k peaks (a8), on the device
    @staticmethod
    def _peak_outputs(P, npeaks):
        return (np.empty((P, npeaks, 3), np.float64), np.empty((P, npeaks), np.float64),
                np.empty((P, npeaks), np.float64), np.empty((P, npeaks, 6), np.float64),
                np.zeros(P, np.int32))

    def grid_find_peaks(self, grids, npeaks=10, width=2, want_residual=False):
        """findPeaks (utils.py:366-396) of P grids (P, n0, n1, n2) on the device ->
        (peaks (P,k,3), amplitude (P,k), mean (P,k), alpha (P,k,6), nfound (P,), residual|None)."""
        grids = np.ascontiguousarray(grids, dtype=np.float64)
        if grids.ndim == 3:
            grids = grids[None]
        P = grids.shape[0]
        shape = np.array(grids.shape[1:], np.int64)
        pk, amp, mean, alpha, nf = self._peak_outputs(P, npeaks)
        res = np.empty_like(grids) if want_residual else None
        self._check(self._lib.fo_grid_find_peaks(self._h, _ptr(grids), P, shape.ctypes.data_as(c_i64p),
                                                 int(npeaks), int(width), _ptr(pk), _ptr(amp), _ptr(mean),
                                                 _ptr(alpha), _ptr(nf), _ptr(res)), "fo_grid_find_peaks")
        return pk, amp, mean, alpha, nf, res

    def sph_isoft_peaks(self, Ilmm, Jmax, npeaks=10, width=2):
        """Coefficients -> overlap grid -> top-npeaks peaks, all on the device (findRotations)."""
        L = int(Jmax)
        Ilmm = np.ascontiguousarray(Ilmm, dtype=np.complex128).reshape(-1, L + 1, 2 * L + 1, 2 * L + 1)
        P = Ilmm.shape[0]
        pk, amp, mean, alpha, nf = self._peak_outputs(P, npeaks)
        self._check(self._lib.fo_sph_isoft_peaks(self._h, _ptr(Ilmm), P, L, int(npeaks), int(width), _ptr(pk),
                                                 _ptr(amp), _ptr(mean), _ptr(alpha), _ptr(nf)),
                    "fo_sph_isoft_peaks")
        return pk, amp, mean, alpha, nf

    def per_align_pairs_peaks(self, params, posA, posB, npeaks=10, width=2):
        """Positions -> |f| grid -> top-npeaks displacements (fractional grid indices), on the device."""
        self._ensure_perm(params.natoms)
        posA = _f64(posA).reshape(-1, params.natoms, 3)
        posB = _f64(posB).reshape(-1, params.natoms, 3)
        P = posA.shape[0]
        pk, amp, mean, alpha, nf = self._peak_outputs(P, npeaks)
        self._check(self._lib.fo_per_align_pairs_peaks(self._h, ctypes.byref(params), _ptr(posA), _ptr(posB), P,
                                                       int(npeaks), int(width), _ptr(pk), _ptr(amp), _ptr(mean),
                                                       _ptr(alpha), _ptr(nf)), "fo_per_align_pairs_peaks")
        return pk, amp, mean, alpha, nf

    def sph_coeffs_direct(self, posA, posB, Jmax, sigma):
        posA = _f64(posA)
        posB = _f64(posB)
        if posA.ndim == 2:
            posA = posA[None]
            posB = posB[None]
        P, N, _ = posA.shape
        self._ensure_perm(N)
        L = int(Jmax)
        out = np.empty((P, L + 1, 2 * L + 1, 2 * L + 1), np.complex128)
        st = np.zeros(P, np.int32)
        self._check(self._lib.fo_sph_coeffs_direct(self._h, _ptr(posA), _ptr(posB), P, N, L,
                                                   float(sigma), _ptr(out), _ptr(st)),
                    "fo_sph_coeffs_direct")
        return out, st

    def sph_align_pairs(self, posA, posB, Jmax, sigma, invert=True, want_grid=False):
        posA = _f64(posA)
        posB = _f64(posB)
        if posA.ndim == 2:
            posA = posA[None]
            posB = posB[None]
        P, N, _ = posA.shape
        self._ensure_perm(N)
        L = int(Jmax)
        bi, bv, fr, grid = self._sph_outputs(P, L, invert, want_grid)
        st = np.zeros(P, np.int32)
        self._check(self._lib.fo_sph_align_pairs(self._h, _ptr(posA), _ptr(posB), P, N, L,
                                                 float(sigma), int(bool(invert)), _ptr(bi), _ptr(bv),
                                                 _ptr(fr), _ptr(grid), _ptr(st)),
                    "fo_sph_align_pairs")
        return bi, bv, fr, grid, st

    def sph_align_pairs_screen_dev(self, d_posA, d_posB, P, N, Jmax, sigma, invert, d_best_idx, d_best_val,
                                   d_frac, d_perm, d_ok, d_status=0):
        self._ensure_perm(N)
        self._check(self._lib.fo_sph_align_pairs_screen_dev(
            self._h, c_void_p(d_posA), c_void_p(d_posB), int(P), int(N), int(Jmax), float(sigma),
            int(bool(invert)), c_void_p(d_best_idx), c_void_p(d_best_val), c_void_p(d_frac), c_void_p(d_perm),
            c_void_p(d_ok), c_void_p(d_status or 0)), "fo_sph_align_pairs_screen_dev")

    def sph_align_pairs_full(self, posA, posB, Jmax, sigma, invert=True, nthreads=0, want_perm=True, out=None):
        """Hot path + device screening of the assignment + host pool (LAP where needed, Kearsley)
        (fo_sph_align_pairs_full).  Structures must be centred.  Returns (dist (P,), orient (P,),
        perm (P,N)|None, rmat (P,3,3), euler (P,O,3), status (P,), nhost).  out: a previous return tuple of the
        same batch size whose arrays are written again."""
        posA = _f64(posA)
        posB = _f64(posB)
        if posA.ndim == 2:
            posA = posA[None]
            posB = posB[None]
        P, N, _ = posA.shape
        self._ensure_perm(N)
        O = 2 if invert else 1
        if (out is not None and out[0].shape == (P,) and out[4].shape == (P, O, 3) and
                (out[2] is not None) == bool(want_perm)):
            dist, orient, perm, rmat, euler, st = out[:6]
            st[:] = 0
        else:
            dist = np.empty(P)
            orient = np.empty(P, np.int32)
            perm = np.empty((P, N), np.int32) if want_perm else None
            rmat = np.empty((P, 3, 3))
            euler = np.empty((P, O, 3))
            st = np.zeros(P, np.int32)
        nhost = ctypes.c_int64(0)
        self._check(self._lib.fo_sph_align_pairs_full(self._h, _ptr(posA), _ptr(posB), P, N, int(Jmax), float(sigma),
                                                      int(bool(invert)), int(nthreads), _ptr(dist), _ptr(orient),
                                                      _ptr(perm), _ptr(rmat), _ptr(euler), _ptr(st),
                                                      ctypes.byref(nhost)), "fo_sph_align_pairs_full")
        return dist, orient, perm, rmat, euler, st, int(nhost.value)

    def sph_refine_rotations(self, Ilmm, Jmax, euler):
        """maxOverlap on the device: (euler_out (P,3), overlap (P,), nevals (P,))."""
        L = int(Jmax)
        Ilmm = np.ascontiguousarray(Ilmm, dtype=np.complex128).reshape(-1, L + 1, 2 * L + 1, 2 * L + 1)
        P = Ilmm.shape[0]
        euler = _f64(euler).reshape(P, 3)
        out = np.empty((P, 3), np.float64)
        ov = np.empty(P, np.float64)
        ne = np.zeros(P, np.int32)
        self._check(self._lib.fo_sph_refine_rotations(self._h, _ptr(Ilmm), P, L, _ptr(euler), _ptr(out), _ptr(ov),
                                                      _ptr(ne)), "fo_sph_refine_rotations")
        return out, ov, ne

    def sph_overlap_gradient(self, Ilmm, Jmax, euler):
        """(value (P,), grad (P,3), hess (P,6)) of the un-weighted overlap at the Euler angles."""
        L = int(Jmax)
        Ilmm = np.ascontiguousarray(Ilmm, dtype=np.complex128).reshape(-1, L + 1, 2 * L + 1, 2 * L + 1)
        P = Ilmm.shape[0]
        euler = _f64(euler).reshape(P, 3)
        val = np.empty(P, np.float64)
        grad = np.empty((P, 3), np.float64)
        hess = np.empty((P, 6), np.float64)
        self._check(self._lib.fo_sph_overlap_gradient(self._h, _ptr(Ilmm), P, L, _ptr(euler), _ptr(val), _ptr(grad),
                                                      _ptr(hess)), "fo_sph_overlap_gradient")
        return val, grad, hess

    def sph_align_pairs_refined(self, posA, posB, Jmax, sigma, invert=True):
        """Hot path + continuous refinement: (best_idx, best_val, frac_idx, euler (P,O,3), overlap (P,O), status)."""
        posA = _f64(posA)
        posB = _f64(posB)
        if posA.ndim == 2:
            posA = posA[None]
            posB = posB[None]
        P, N, _ = posA.shape
        self._ensure_perm(N)
        L = int(Jmax)
        bi, bv, fr, _ = self._sph_outputs(P, L, invert, False)
        eu = np.empty_like(fr)
        ov = np.empty_like(bv)
        st = np.zeros(P, np.int32)
        self._check(self._lib.fo_sph_align_pairs_refined(self._h, _ptr(posA), _ptr(posB), P, N, L, float(sigma),
                                                         int(bool(invert)), _ptr(bi), _ptr(bv), _ptr(fr), _ptr(eu),
                                                         _ptr(ov), _ptr(st)), "fo_sph_align_pairs_refined")
        return bi, bv, fr, eu, ov, st

    def sph_align_pairs_refined_dev(self, d_posA, d_posB, P, N, Jmax, sigma, invert, d_best_idx, d_best_val,
                                    d_frac, d_euler, d_overlap, d_status=0):
        self._ensure_perm(N)
        self._check(self._lib.fo_sph_align_pairs_refined_dev(
            self._h, c_void_p(d_posA), c_void_p(d_posB), int(P), int(N), int(Jmax), float(sigma),
            int(bool(invert)), c_void_p(d_best_idx), c_void_p(d_best_val), c_void_p(d_frac),
            c_void_p(d_euler), c_void_p(d_overlap), c_void_p(d_status or 0)), "fo_sph_align_pairs_refined_dev")

    def sph_align_pairs_dev(self, d_posA, d_posB, P, N, Jmax, sigma, invert, d_best_idx,
                            d_best_val, d_frac, d_grid=0, d_status=0):
        self._ensure_perm(N)
        self._check(self._lib.fo_sph_align_pairs_dev(
            self._h, c_void_p(d_posA), c_void_p(d_posB), int(P), int(N), int(Jmax), float(sigma),
            int(bool(invert)), c_void_p(d_best_idx), c_void_p(d_best_val), c_void_p(d_frac),
            c_void_p(d_grid or 0), c_void_p(d_status or 0)), "fo_sph_align_pairs_dev")

    def sph_harm_coeffs(self, pos, nmax, Jmax, harmscale, sigma):
        pos = _f64(pos)
        if pos.ndim == 2:
            pos = pos[None]
        S, N, _ = pos.shape
        self._ensure_perm(N)
        ng = len(self._perm_key[1]) if self._perm_key and self._perm_key[0] == N else 1
        L = int(Jmax)
        out = np.empty((S, ng, nmax + 1, L + 1, 2 * L + 1), np.complex128)
        st = np.zeros(S, np.int32)
        self._check(self._lib.fo_sph_harm_coeffs(self._h, _ptr(pos), S, N, int(nmax), L,
                                                 float(harmscale), float(sigma), _ptr(out), _ptr(st)),
                    "fo_sph_harm_coeffs")
        return out, st

    def sph_bank_create(self, pos, nmax, Jmax, harmscale, sigma):
        pos = _f64(pos)
        if pos.ndim == 2:
            pos = pos[None]
        S, N, _ = pos.shape
        self._ensure_perm(N)
        h = c_void_p()
        self._check(self._lib.fo_sph_bank_create(self._h, _ptr(pos), S, N, int(nmax), int(Jmax),
                                                 float(harmscale), float(sigma), ctypes.byref(h)),
                    "fo_sph_bank_create")
        b = Bank(self, h, S)
        b.Jmax = int(Jmax)
        return b

    def sph_align_bank(self, bank, pairs, invert=True, want_grid=False, want_avg=True):
        pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        P = pairs.shape[0]
        bi, bv, fr, grid = self._sph_outputs(P, bank.Jmax, invert, want_grid)
        avg = np.empty(P, np.float64) if want_avg else None
        self._check(self._lib.fo_sph_align_bank(self._h, bank._h, _ptr(pairs), P,
                                                int(bool(invert)), _ptr(bi), _ptr(bv), _ptr(fr),
                                                _ptr(avg), _ptr(grid)), "fo_sph_align_bank")
        return bi, bv, fr, avg, grid


    def sph_align_bank_refined(self, bank, pairs, invert=True, want_avg=True):
        pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        P = pairs.shape[0]
        bi, bv, fr, _ = self._sph_outputs(P, bank.Jmax, invert, False)
        eu = np.empty_like(fr)
        ov = np.empty_like(bv)
        avg = np.empty(P, np.float64) if want_avg else None
        self._check(self._lib.fo_sph_align_bank_refined(self._h, bank._h, _ptr(pairs), P, int(bool(invert)),
                                                        _ptr(bi), _ptr(bv), _ptr(fr), _ptr(avg), _ptr(eu),
                                                        _ptr(ov)), "fo_sph_align_bank_refined")
        return bi, bv, fr, avg, eu, ov


class Bank(object):
    """Device-resident coefficient bank (fo_bank)."""

    def __init__(self, ctx, handle, nstruct):
        self._ctx = ctx
        self._h = handle
        self.nstruct = nstruct

    def close(self):
        if self._h and self._ctx._h:
            self._ctx._lib.fo_bank_destroy(self._ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}
_default_lock = threading.Lock()


def default_context(device=None):
    """Process-wide default Context per device (device from $FASTOVERLAP_DEVICE, LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("FASTOVERLAP_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _default_lock:
        ctx = _default_ctx.get(device)
        if ctx is None or ctx._h is None:
            ctx = Context(device)
            _default_ctx[device] = ctx
        return ctx
