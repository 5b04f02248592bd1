"""Build libfastoverlap_b200.so in-tree with nvcc for sm_100a (B200).

    python -m fastoverlap_b200.build [--force] [--verbose]

The library has no torch dependency: it links only the CUDA runtime (static), so the same
.so is what a maintainer of the reference would load with ctypes.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libfastoverlap_b200.so"
SOURCES = ["fo_context.cu", "fo_periodic.cu", "fo_spherical.cu", "fo_refine.cu", "fo_peaks.cu", "fo_assign.cu", "fo_host.cu"]
HOST_ONLY = {"fo_host.cu"}
HEADERS = [os.path.join(CSRC, "fo_internal.h"), os.path.join(CSRC, "fo_symdft.cuh"), os.path.join(CSRC, "fo_async.cuh"),
           os.path.join(HERE, "..", "include", "fastoverlap_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-fopenmp,-msse4.1",
    "-fmad=true",
    "-prec-div=true", "-prec-sqrt=true",
]


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build %s" % LIBNAME)
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs, cmds = [], []
    for src in srcs:
        obj = os.path.join(LIBDIR, os.path.basename(src).replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            if os.path.basename(src) in HOST_ONLY:
                # no device code: compiled by g++ directly (function multiversioning, OpenMP SIMD)
                cmd = [shutil.which("g++") or "g++", "-x", "c++", "-std=c++17", "-O3", "-fPIC", "-fopenmp",
                       "-msse4.1", "-fno-math-errno", "-Wno-psabi", "-I", os.path.join(os.path.dirname(nvcc), "..", "include"), "-c", src, "-o", obj]
            else:
                cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                if os.path.basename(src) not in HOST_ONLY:
                    cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            cmds.append(cmd)
    if verbose:  # keep the ptxas reports of the translation units apart
        for cmd in cmds:
            subprocess.run(cmd, check=True)
    elif cmds:   # the translation units are independent: compile them concurrently
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(cmds)) as pool:
            list(pool.map(lambda c: subprocess.run(c, check=True), cmds))
    out = lib_path()
    if force or _stale(out, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + \
              ["-Xcompiler", "-fopenmp", "-lgomp"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(p)
