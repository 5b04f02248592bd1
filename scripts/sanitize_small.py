"""Small-shape pass through every kernel family, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
(the mbarrier hand-over of per_sf3_kernel / per_sfx_kernel and its tensor-memory stash, the bulk-copy staging of
per_xf6_kernel / sph_isoft4_kernel / sph_isoft5_kernel, the mbarrier ring of sph_direct2_kernel, the in-place
transforms and named barriers of sph_isoft5_kernel, the cp.async staging and the shared-memory reductions of the
screening kernels are what racecheck is pointed at).  `python scripts/sanitize_small.py sph` runs the cluster
part only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fastoverlap_b200 as fob

ctx = fob.Context(0)
rng = np.random.default_rng(0)
SPH_ONLY = len(sys.argv) > 1 and sys.argv[1] == "sph"

if not SPH_ONLY:
    # periodic: bench configuration (per_sf3, per_cross, per_xf4, per_assign) incl. the grid output
    wl = bench.Blj256()
    wl.setup(ctx)
    A, B, _ = wl.make(6, 0)
    r = ctx.per_align_pairs_full(wl.params, A, B, niter=10, nthreads=2)
    print("blj256 full", r[0][:3], "host", r[-1], flush=True)
    g = ctx.per_align_pairs(wl.params, A[:2], B[:2], want_grid=True)
    print("blj256 grid max", g[3].max(), flush=True)
    # fine k-grid (per_sf2, per_xf5) and a small generic case (per_sf, per_xf), ragged groups
    for N, n, groups in ((36, 16, [np.arange(20), np.arange(20, 36)]), (17, 3, None), (40, 6, [np.arange(23), np.arange(23, 40)])):
        box = np.array([4.0, 4.5, 5.1])
        al = fob.PeriodicAlign(N, box, groups, n=n, ctx=ctx)
        p1 = rng.uniform(-0.5, 0.5, size=(3, N, 3)) * box
        p2 = p1 + rng.uniform(0, 1, size=(3, 1, 3)) * box + rng.normal(scale=0.03, size=p1.shape)
        d = al.align_batch(p1, p2, nthreads=2)[0]
        gg = ctx.per_align_pairs(al._params(), p1[:1], p2[:1], want_grid=True)
        print("periodic N %d n %d dist" % (N, n), d, flush=True)
    ctx.set_option("force_generic", 1)
    al = fob.PeriodicAlign(17, np.array([4.0, 4.5, 5.1]), None, n=3, ctx=ctx)
    p1 = rng.uniform(-0.5, 0.5, size=(2, 17, 3)) * 4
    print("generic periodic", al.align_batch(p1, p1 + 0.3, nthreads=1)[0], flush=True)
    ctx.set_option("force_generic", 0)
    # all-vs-all through the structure-factor bank + top-k peaks
    al = wl.al
    print("alignGroup", al.alignGroup(A[:3, :, :].copy())[0, 1], flush=True)
    print("npeaks", al(A[0], B[0], npeaks=3)[0], flush=True)

# clusters: LJ38 bench configuration (prep2, bessel2, direct2, isoft5, final2, assign), generic Jmax, harmonic bank
wl = bench.Lj38()
wl.setup(ctx)
A, B, _ = wl.make(6, 0)
r = ctx.sph_align_pairs_full(A, B, 15, 0.3, invert=True, nthreads=2)
print("lj38 full", r[0][:3], "host", r[-1], flush=True)
g = ctx.sph_align_pairs(A[:2], B[:2], 15, 0.3, invert=True, want_grid=True)
r = ctx.sph_align_pairs_refined(A[:3], B[:3], 15, 0.3, invert=True)
print("lj38 refined overlap", r[4][0], flush=True)
X = rng.normal(size=(3, 16, 3))
X -= X.mean(1, keepdims=True)
ctx.set_perm([np.arange(16)], 16)
print("Jmax 21", ctx.sph_align_pairs(X, X[::-1].copy(), 21, 0.4, invert=True)[1][0], flush=True)
sh = fob.SphericalHarmonicAlign(0.3, 1.0, 20, 15, ctx=ctx)
print("harmonic compareList", sh.compareList(A[:3])[0][0], flush=True)
print("harmonic alignGroup", sh.alignGroup(A[:3])[0], flush=True)
sa = fob.SphericalAlign(0.3, 15, ctx=ctx)
print("malign", sa.malign(A[0], B[0], nrot=3)[0], flush=True)
# a 70-atom cluster: the GEMM form of the direct coefficients
X = rng.normal(size=(2, 70, 3)) * 1.5
X -= X.mean(1, keepdims=True)
ctx.set_perm([np.arange(70)], 70)
print("N 70", ctx.sph_align_pairs(X, X[::-1].copy(), 12, 0.45, invert=True)[1][0], flush=True)
print("launches", ctx.launch_count(), flush=True)

# older kernel variants kept for A/B and for grids outside the fast paths: bank path (per_sf3 + per_cross6 + per_xf6),
# sph_isoft3 (stage A -> shared memory -> stage B), even Jmax (two planes per CTA)
if not SPH_ONLY:
    wl = bench.Blj256()
    wl.setup(ctx)
    A, B, _ = wl.make(4, 1)
    ctx.set_option("per_pairs_fused", 0)
    print("blj256 bank path", ctx.per_align_pairs(wl.params, A, B)[1][:2], flush=True)
    ctx.set_option("per_pairs_fused", 1)
wl = bench.Lj38()
wl.setup(ctx)
A, B, _ = wl.make(4, 1)
ctx.set_option("sph_isoft_variant", 3)
print("lj38 isoft3", ctx.sph_align_pairs(A, B, 15, 0.3, invert=True)[1][0], flush=True)
ctx.set_option("sph_isoft_variant", 4)
print("lj38 isoft4", ctx.sph_align_pairs(A, B, 15, 0.3, invert=True)[1][0], flush=True)
ctx.set_option("sph_isoft_variant", 0)
ctx.set_option("sph_direct_ring", -1)
print("lj38 direct_mma", ctx.sph_align_pairs(A, B, 15, 0.3, invert=True)[1][0], flush=True)
ctx.set_option("sph_direct_ring", 0)
print("lj38 Jmax 14", ctx.sph_align_pairs(A, B, 14, 0.3, invert=True)[1][0], flush=True)
print("lj38 Jmax 7", ctx.sph_align_pairs(A, B, 7, 0.3, invert=True)[1][0], flush=True)
