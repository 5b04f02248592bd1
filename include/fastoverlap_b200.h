/*
 * fastoverlap_b200 -- C ABI of the B200-native FASTOVERLAP overlap-maximisation hot path.
 *
 * This header is the drop-in boundary.  It replaces the f2py
 * extension modules of the reference (fastoverlap/f90/__init__.py:3-21):
 *
 *   fastbulk.bulkfastoverlap.*      (reference fastoverlap/f90/fastbulk.f90)
 *   fastclusters.clusterfastoverlap.* (reference fastoverlap/f90/fastclusters.f90, DSOFT.f90)
 *   *.fastoverlaputils.setperm      (reference fastoverlap/f90/fastutils.f90:117-167)
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  No torch / numpy types.
 *   - all reals are double, complex numbers are interleaved (re, im) doubles,
 *     arrays are row-major (C order) exactly as numpy lays the reference's arrays out.
 *   - the caller owns every buffer passed in; the library never frees caller memory.
 *   - every function returns FO_OK (0) or a negative error code; the message is
 *     available from fo_last_error().  Nothing ever calls exit/abort (the reference
 *     Fortran STOPs the interpreter, fastclusters.f90:1112-1165).
 *   - no hidden global state (the reference keeps module-level SAVE state,
 *     fastutils.f90:70-75, DSOFT.f90:32-34): everything lives in an fo_ctx.
 *   - one fo_ctx per (host thread, GPU).  Calls on one ctx must not overlap in time.
 *   - functions whose name ends in _dev take DEVICE pointers, are enqueued on the
 *     context's stream and return without synchronising (call fo_sync()).
 *     All other functions take HOST pointers and return with results in place.
 *   - there is NO CPU fallback: if no CUDA device is usable fo_create fails.
 */
#ifndef FASTOVERLAP_B200_H
#define FASTOVERLAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FO_OK 0
#define FO_ERR_INVALID -1   /* bad argument */
#define FO_ERR_CUDA -2      /* CUDA runtime error (message has the detail) */
#define FO_ERR_NOMEM -3     /* host or device allocation failed */
#define FO_ERR_UNSUPPORTED -4 /* size outside what the kernels support */

/* per-pair status bits (written to the optional status[] outputs) */
#define FO_STATUS_OK 0
#define FO_STATUS_NONFINITE 1   /* a coordinate or result was NaN/Inf */
#define FO_STATUS_ATOM_AT_ORIGIN 2 /* spherical: r == 0 (the reference yields NaN: utils.py:444) */

typedef struct fo_ctx fo_ctx;
typedef struct fo_bank fo_bank; /* device-resident coefficient bank (structure factors or C_nlm) */

/* ---------------------------------------------------------------- context */

/* Create a context on CUDA device `device`.  Fails (FO_ERR_CUDA) when no GPU is usable. */
int fo_create(int device, fo_ctx** out);
void fo_destroy(fo_ctx* ctx);
/* Message of the last error on this ctx (or of the last failed fo_create when ctx==NULL). */
const char* fo_last_error(const fo_ctx* ctx);
/* Use an externally owned cudaStream_t (e.g. torch's current stream) for all work; NULL is the
 * legacy default stream.  fo_reset_stream returns to the context's own non-blocking stream. */
int fo_set_stream(fo_ctx* ctx, void* cuda_stream);
int fo_reset_stream(fo_ctx* ctx);
int fo_sync(fo_ctx* ctx);
/* Library / device facts: writes {sm_count, l2_bytes, smem_per_block_optin, cc_major*10+cc_minor}. */
int fo_device_info(fo_ctx* ctx, int64_t out[4]);
/* Number of kernels this ctx has launched since creation (bench.py "gpu_launches"). */
int64_t fo_launch_count(const fo_ctx* ctx);

/* Library options.  "force_generic" (0/1): use the any-size kernels instead of the shared-memory
 * tensor-core fast paths (testing: both families must give the same results).
 * "direct_gemm_min_atoms" (default 64): clusters with at least this many atoms compute the direct
 * coefficients (fo_sph_coeffs_direct / fo_sph_align_pairs) with the tensor-core GEMM kernels. */
int fo_set_option(fo_ctx* ctx, const char* name, int64_t value);

/* Measured throughput of the FP64 tensor pipe (mma.sync.m8n8k4.f64 microbenchmark) in TFLOP/s:
 * the roofline denominator of the kernels that run their contractions as DMMA. */
int fo_measure_fp64_tensor_peak(fo_ctx* ctx, double* tflops);

/* Per-kernel device timing with CUDA events on the ctx stream (bench.py roofline).  Between
 * fo_profile_begin and fo_profile_end every launch of a profiled kernel class is bracketed by
 * an event pair; fo_profile_end synchronises and returns, per class, the summed duration in
 * milliseconds and the launch count.  Classes: FO_PROF_*. */
#define FO_PROF_PER_SF 0      /* periodic structure factors            */
#define FO_PROF_PER_XF 1      /* periodic cross-spectrum + DFT + argmax */
#define FO_PROF_SPH_COEF 2    /* spherical direct coefficients         */
#define FO_PROF_SPH_HARM 3    /* spherical harmonic-basis coefficients */
#define FO_PROF_SPH_DOT 4     /* C_nlm contraction to I_lmm'           */
#define FO_PROF_SPH_ISOFT 5   /* Wigner-d contraction + 2-D DFT + argmax */
#define FO_PROF_PEAKS 6       /* top-k peak extraction (fit-and-subtract) */
#define FO_PROF_SPH_REFINE 7  /* continuous rotation refinement (damped Newton) */
#define FO_PROF_ASSIGN 8      /* nearest-partner screening of the assignment (full alignment) */
#define FO_PROF_NKINDS 9
int fo_profile_begin(fo_ctx* ctx);
int fo_profile_end(fo_ctx* ctx, double ms_out[FO_PROF_NKINDS], int64_t count_out[FO_PROF_NKINDS]);

/* Permutation groups: the atoms of group g are atom_idx[group_offsets[g] .. group_offsets[g+1]).
 * 0-based.  Replaces fastoverlaputils.setperm(natoms, permgroup(1-based), npermsize)
 * (reference fastutils.f90:117-167; called from sphericalAlignment.py:483,
 * periodicAlignment.py:523).  Held per ctx, not globally. */
int fo_set_perm(fo_ctx* ctx, const int32_t* group_offsets, int64_t ngroups,
                const int32_t* atom_idx, int64_t natoms);

/* 5-smooth FFT length >= target: utils.py:278-313 (_next_fast_len) == FASTLEN table
 * fastutils.f90:78-90. */
int64_t fo_next_fast_len(int64_t target);

/* Measured FP64 FMA throughput of the device in TFLOP/s (dependent-chain-free DFMA
 * microbenchmark kernel; used as the roofline denominator because MEASURED_PEAKS.json
 * has no FP64 figure). */
int fo_measure_fp64_peak(fo_ctx* ctx, double* tflops);

/* ---------------------------------------------------------------- periodic (fastbulk) */

/* Geometry of a periodic problem.  natoms atoms in an orthorhombic box; k-grid
 * k = 2*pi/box * (-nwave..nwave)^3 (reference SETWAVEK fastbulk.f90:567-597,
 * periodicAlignment.py:384-386); displacement grid nfspace^3; Gaussian width sigma
 * (`scale` / KWIDTH). */
typedef struct fo_per_params {
  int64_t natoms;
  double box[3];
  int64_t nwave;    /* n  : k index runs -n..n           */
  int64_t nfspace;  /* F  : FFT length per axis, F >= 2(2n+1)+1 is what the reference uses */
  double sigma;     /* kernel width                         */
} fo_per_params;

/* Defaults of the reference for (natoms, box): sigma = (V/N)^(1/3)/3, nwave =
 * ceil(1.3 N^(1/3)), nfspace = next_fast_len(4 nwave + 3)
 * (periodicAlignment.py:375-378,391-392; ALIGN fastbulk.f90:296-315). */
int fo_per_defaults(int64_t natoms, const double box[3], double* sigma, int64_t* nwave,
                    int64_t* nfspace);

/* Structure factors S[s, g, kx, ky, kz] = sum_{j in group g} exp(-i k.r_j) on the full
 * (2n+1)^3 grid, index 0 = k=-n, for S structures.  out is [S, ngroups, W, W, W, 2]
 * doubles, W = 2n+1.  Replaces PeriodicAlign.calcFourierCoeff (periodicAlignment.py:400-406)
 * / PERIODICFOURIERPERM (fastbulk.f90:635-665).  Uses the ctx permutation groups. */
int fo_per_structure_factors(fo_ctx* ctx, const fo_per_params* p, const double* pos /*[S,N,3]*/,
                             int64_t nstruct, double* out);

/* The whole periodic hot path for P independent pairs, host buffers:
 *   structure factors of posA[i], posB[i]  ->  C = sum_g S_A conj(S_B) exp(-k^2 sigma^2)
 *   -> zero-padded forward 3-D DFT to F^3 -> modulus -> arg-max (+ parabolic refinement).
 * Replaces PeriodicAlign.setPos + findDisps(npeaks=1) (periodicAlignment.py:408-456)
 * / ALIGN1 + ALIGNCOEFFS up to FINDPEAKS (fastbulk.f90:414-531).
 *   best_idx [P,3] int64 : arg-max index (dx,dy,dz), first in C order on exact ties (numpy argmax)
 *   best_val [P]         : fabs at the arg-max
 *   frac_idx [P,3]       : findMax()'s parabolically interpolated index (utils.py:319-338)
 *   grid_out [P,F,F,F]   : optional (may be NULL) full |f| grid, = PeriodicAlign.fabs
 *   status   [P] int32   : optional (may be NULL) FO_STATUS_* bits per pair
 * displacement = frac_idx * box / F (periodicAlignment.py:455). */
int fo_per_align_pairs(fo_ctx* ctx, const fo_per_params* p, const double* posA /*[P,N,3]*/,
                       const double* posB /*[P,N,3]*/, int64_t npairs, int64_t* best_idx,
                       double* best_val, double* frac_idx, double* grid_out, int32_t* status);

/* Same, all pointers are DEVICE pointers, enqueued on the ctx stream, no synchronisation.
 * (bench.py `value`: inputs resident in HBM.) */
int fo_per_align_pairs_dev(fo_ctx* ctx, const fo_per_params* p, const double* d_posA,
                           const double* d_posB, int64_t npairs, int64_t* d_best_idx,
                           double* d_best_val, double* d_frac_idx, double* d_grid_out,
                           int32_t* d_status);

/* The whole periodic alignment for P independent pairs, host buffers: the hot path of fo_per_align_pairs,
 * then BasePeriodicAlignment.refine (periodicAlignment.py:27-80) / ITERATIVEALIGN(bulk) (alignutils.f90:111-286).
 * The Hungarian step stays on the host; what runs on the device is the screening that makes it unnecessary for
 * well-aligned pairs: with the arg-max displacement applied every atom of B looks for its nearest same-group
 * atom of A (minimum image); when those nearest partners form a permutation and every runner-up is further away
 * by a clear gap, that permutation is the unique optimum of the assignment problem, and the permutation <->
 * mean-displacement loop and the final distance are finished on the device with the host path's operations in
 * the host path's order (bit-identical results).  Every other pair goes through fo_host_refine_periodic on
 * `nthreads` host threads (<= 0: all cores) while the GPU works on the next chunk.
 *   dist [P]; perm [P,N] int32 (nullable): X2 = posB[perm] - disp; disp [P,3] (nullable);
 *   frac_idx [P,3] (nullable): the hot path's interpolated arg-max; status [P] (nullable);
 *   nhost (nullable): number of pairs the host pool had to take. */
int fo_per_align_pairs_full(fo_ctx* ctx, const fo_per_params* p, const double* posA /*[P,N,3]*/,
                            const double* posB /*[P,N,3]*/, int64_t npairs, int niter, int nthreads,
                            double* dist, int32_t* perm, double* disp, double* frac_idx, int32_t* status,
                            int64_t* nhost);

/* Device-resident form (all pointers are DEVICE pointers, enqueued on the ctx stream, no synchronisation, no
 * host pool): d_flag [P] int32 = 0 where the pair was settled on the device (d_dist / d_disp / d_perm valid),
 * non-zero where the host LAP is needed (the caller then runs fo_host_refine_periodic_subset on those).
 * d_perm [P,N] nullable. */
int fo_per_align_pairs_full_dev(fo_ctx* ctx, const fo_per_params* p, const double* d_posA, const double* d_posB,
                                int64_t npairs, int niter, double* d_dist, int32_t* d_perm, double* d_disp,
                                int32_t* d_flag, int64_t* d_best_idx, double* d_best_val, double* d_frac_idx,
                                int32_t* d_status);

/* Same path starting from caller-supplied structure factors (the reference's
 * `Cs=[c1,c2]` hook, periodicAlignment.py:408-432,458-460): CA, CB are
 * [P, ngroups, W, W, W, 2] as produced by fo_per_structure_factors.  Only the kz>=0 half
 * is read (structure factors of real densities are Hermitian). */
int fo_per_align_coeffs(fo_ctx* ctx, const fo_per_params* p, const double* CA, const double* CB,
                        int64_t npairs, int64_t* best_idx, double* best_val, double* frac_idx,
                        double* grid_out, int32_t* status);

/* Device-resident bank of structure factors for S structures (all-vs-all use:
 * PeriodicAlign.alignGroup periodicAlignment.py:462-479 / ALIGNGROUP fastbulk.f90:336-412
 * compute coefficients once per structure). */
int fo_per_bank_create(fo_ctx* ctx, const fo_per_params* p, const double* pos /*[S,N,3]*/,
                       int64_t nstruct, fo_bank** out);
/* pairs [P,2] int64 = (index of structure A, index of structure B) into the bank. */
int fo_per_align_bank(fo_ctx* ctx, const fo_per_params* p, const fo_bank* bank,
                      const int64_t* pairs, int64_t npairs, int64_t* best_idx, double* best_val,
                      double* frac_idx, double* grid_out, int32_t* status);
/* fo_per_align_bank with a cell-symmetry operation per pair (cubic box; the reference's OHCELLT branch: ALIGN1
 * fastbulk.f90:458-480, OHTRANSFORMCOEFFS fastbulk.f90:863-1380).  Pair i aligns structure A with R_i B, where
 * (R r)_j = s_j r_{p_j} is a signed permutation, encoded as ops[i] = p_0 | p_1 << 2 | p_2 << 4 | (s_j < 0) << (6 + j).
 * The structure factors of R B are an index permutation of B's bank entry (S_RB(k) = S_B(R^T k)), so the 48
 * octahedral images of a structure cost 48 cross-spectra + transforms and no structure-factor pass.
 * ops = NULL: identity for every pair.  FO_ERR_UNSUPPORTED for nwave > 11. */
int fo_per_align_bank_ops(fo_ctx* ctx, const fo_per_params* p, const fo_bank* bank, const int64_t* pairs,
                          const int32_t* ops /*[P] or NULL*/, int64_t npairs, int64_t* best_idx, double* best_val,
                          double* frac_idx, double* grid_out, int32_t* status);
void fo_bank_destroy(fo_ctx* ctx, fo_bank* bank);

/* ---------------------------------------------------------------- spherical (fastclusters) */

/* Spherical harmonics of the atoms of S structures (row a1): Y_out [S, L+1, 2L+1, N] complex in the
 * layout of BaseSphericalAlignment.sphHarm (sphericalAlignment.py:57-65: scipy / Condon-Shortley
 * convention, negative m at index m + 2L+1, zeros for |m| > l) / RYML (fastclusters.f90:602-652);
 * r_out [S, N] (nullable) = |pos| (calcThetaPhiR, utils.py:439-445).  An atom at the origin sets
 * FO_STATUS_ATOM_AT_ORIGIN (the reference yields NaN). */
int fo_sph_ylm(fo_ctx* ctx, const double* pos /*[S,N,3]*/, int64_t nstruct, int64_t natoms, int64_t Jmax,
               double* Y_out, double* r_out, int32_t* status);

/* Inverse SO(3) Fourier transform + arg-max for P coefficient sets, host buffers.
 * Ilmm is [P, L+1, 2L+1, 2L+1, 2] doubles in the reference's numpy layout (negative m stored
 * at index m + 2L+1, i.e. python negative-index wrap; soft.py:115-125).
 *   grid[a,k,g] = sum_l sum_{m1,m2} sqrt((2l+1)/2) d^l_{m1m2}(beta_k) I^l_{m1m2} e^{i(m1 a + m2 g) 2pi/2B}
 * with B = L+1, beta_k = pi(2k+1)/4B (the weighted grid of the reference).
 * Replaces SOFT.iSOFT + findMax (soft.py:115-125, utils.py:319-338) / CALCOVERLAP + ISOFT +
 * FINDROTATIONS arg-max (fastclusters.f90:918-988, DSOFT.f90:265-329).
 *   invert != 0 : also evaluate the inverted orientation I_inv^l = (-1)^l I^l
 *                 (fastclusters.f90:351-353; sphericalAlignment.py:370); outputs then have
 *                 a leading orientation axis of length 2 (0 = normal, 1 = inverted).
 *   best_idx [P,O,3] int64, best_val [P,O], frac_idx [P,O,3], grid_out [P,O,2B,2B,2B] or NULL.
 * Euler angles = frac_idx * (2pi/2B, pi/2B, 2pi/2B) + (0, pi/4B, 0)  (soft.py:127-130). */
int fo_sph_isoft_argmax(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax, int invert,
                        int64_t* best_idx, double* best_val, double* frac_idx, double* grid_out);

/* Plain inverse SO(3) transform of arbitrary complex coefficients (SOFT.iSOFT, soft.py:115-125):
 * grid_re / grid_im [P,2B,2B,2B]; grid_im may be NULL.  Two real-output transforms internally. */
int fo_sph_isoft(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax, double* grid_re,
                 double* grid_im);

/* Direct (N^2 Bessel) SO(3) coefficients of P pairs of already-centred structures:
 *   I[l,m1,m2] = 4 pi^2.5 sigma^3 sum_g sum_{j,k in g} i_l(r_j r_k / 2 sigma^2)
 *                exp(-(r_j^2+r_k^2)/4 sigma^2) Y_lm1(A_j) conj(Y_lm2(B_k))
 * Replaces SphericalAlign.calcSO3Coeffs summed over perm groups
 * (sphericalAlignment.py:175,260-273) / FOURIERCOEFFS (fastclusters.f90:868-916) WITHOUT the
 * Fortran-only 4 sigma pair cut-off.  Ilmm_out [P, L+1, 2L+1, 2L+1, 2]. */
int fo_sph_coeffs_direct(fo_ctx* ctx, const double* posA /*[P,N,3]*/, const double* posB,
                         int64_t npairs, int64_t natoms, int64_t Jmax, double sigma,
                         double* Ilmm_out, int32_t* status);

/* The whole spherical hot path (direct coefficients) for P pairs of centred structures:
 * coefficients -> iSOFT -> arg-max for the normal and (invert!=0) inverted orientation.
 * Replaces BaseSphericalAlignment.align up to findMax (sphericalAlignment.py:160-194) /
 * ALIGN up to FINDROTATIONS (fastclusters.f90:129-269).  Outputs as fo_sph_isoft_argmax. */
int fo_sph_align_pairs(fo_ctx* ctx, const double* posA, const double* posB, int64_t npairs,
                       int64_t natoms, int64_t Jmax, double sigma, int invert, int64_t* best_idx,
                       double* best_val, double* frac_idx, double* grid_out, int32_t* status);
int fo_sph_align_pairs_dev(fo_ctx* ctx, const double* d_posA, const double* d_posB,
                           int64_t npairs, int64_t natoms, int64_t Jmax, double sigma, int invert,
                           int64_t* d_best_idx, double* d_best_val, double* d_frac_idx,
                           double* d_grid_out, int32_t* d_status);

/* The whole cluster alignment for P pairs of centred structures, host buffers: the hot path of
 * fo_sph_align_pairs, then BaseSphericalAlignment.refine (sphericalAlignment.py:118-127) for every orientation
 * with the Fortran orientation rule (smaller distance, fastclusters.f90:243-254).  On the device: rotation by the
 * Euler angles of the grid maximum and the nearest-partner screening of the assignment (see
 * fo_per_align_pairs_full); on the host pool (`nthreads`, <= 0: all cores), overlapped with the GPU's next
 * chunk: the LAP where the screening failed and the Kearsley fit for every (pair, orientation).
 *   dist [P]; orient [P], perm [P,N], rmat [P,9], euler [P,O,3] (grid-maximum angles), status [P] nullable;
 *   nhost (nullable): number of (pair, orientation) assignments the host LAP solved. */
int fo_sph_align_pairs_full(fo_ctx* ctx, const double* posA, const double* posB, int64_t npairs, int64_t natoms,
                            int64_t Jmax, double sigma, int invert, int nthreads, double* dist, int32_t* orient,
                            int32_t* perm, double* rmat, double* euler, int32_t* status, int64_t* nhost);

/* Device-resident form of the GPU stage of fo_sph_align_pairs_full: fo_sph_align_pairs_dev followed by the
 * screening kernel.  d_perm [P,O,N] int32 = nearest-partner permutation of every (pair, orientation) after the
 * rotation by the grid-maximum Euler angles, d_ok [P,O] int32 = 1 where it is the proven optimum of the
 * assignment (fo_host_refine_spherical_hint takes both as they are). */
int fo_sph_align_pairs_screen_dev(fo_ctx* ctx, const double* d_posA, const double* d_posB, int64_t npairs,
                                  int64_t natoms, int64_t Jmax, double sigma, int invert, int64_t* d_best_idx,
                                  double* d_best_val, double* d_frac_idx, int32_t* d_perm, int32_t* d_ok,
                                  int32_t* d_status);

/* Harmonic-basis coefficients C[s, g, n, l, m] = sum_{j in g} d_nl(r_j; sigma, r0) conj(Y_lm(r_j))
 * for S centred structures; out [S, ngroups, nmax+1, L+1, 2L+1, 2] (numpy layout, negative m
 * wrapped).  Replaces SphericalHarmonicAlign.calcHarmCoeff (sphericalAlignment.py:345-361) /
 * HARMONICCOEFFSPERM (fastclusters.f90:689-716). */
int fo_sph_harm_coeffs(fo_ctx* ctx, const double* pos /*[S,N,3]*/, int64_t nstruct, int64_t natoms,
                       int64_t nmax, int64_t Jmax, double harmscale, double sigma, double* out,
                       int32_t* status);

/* Device-resident bank of harmonic coefficients + the all-vs-all / pair-list overlap search:
 *   I[l,m1,m2] = sum_g sum_n conj(C_A[g,n,l,m1]) C_B[g,n,l,m2]   (calcSO3Harm
 *   sphericalAlignment.py:363-372 / DOTHARMONICCOEFFSPERM fastclusters.f90:765-788)
 * -> iSOFT -> arg-max, for both orientations when invert != 0.
 * avg_overlap [P] (may be NULL) = sum |I_lmm'|^2 as CALCSIMILARITY (fastclusters.f90:790-818). */
int fo_sph_bank_create(fo_ctx* ctx, const double* pos /*[S,N,3]*/, int64_t nstruct,
                       int64_t natoms, int64_t nmax, int64_t Jmax, double harmscale, double sigma,
                       fo_bank** out);
int fo_sph_align_bank(fo_ctx* ctx, const fo_bank* bank, const int64_t* pairs /*[P,2]*/,
                      int64_t npairs, int invert, int64_t* best_idx, double* best_val,
                      double* frac_idx, double* avg_overlap, double* grid_out);

/* Continuous refinement of the rotation: maximise the un-weighted overlap
 *   f(a,b,g) = Re sum_{l m1 m2} conj(I^l_{m1m2}) e^{-i m1 a} d^l_{m1m2}(b) e^{-i m2 g}
 * from euler_in [P,3] by a damped Newton iteration on the device, one CTA per rotation (analytic
 * gradient and Hessian from the Wigner-d recurrence).  Replaces BaseSphericalAlignment.maxOverlap /
 * getEnergyGradient / calcWignerMatrices (sphericalAlignment.py:67-113; scipy L-BFGS-B on E = -f):
 * euler_out [P,3] = res.x, overlap_out [P] = -res.fun, nevals_out [P] (nullable) = evaluations used.
 * Ilmm as in fo_sph_isoft_argmax (what findRotation receives; it is conjugated internally as
 * findRotation does, sphericalAlignment.py:190-194). */
int fo_sph_refine_rotations(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax,
                            const double* euler_in, double* euler_out, double* overlap_out,
                            int32_t* nevals_out);

/* One evaluation of that objective: value [P] = f, grad [P,3] = df/d(a,b,g), hess [P,6] (nullable)
 * = second derivatives (aa ab ag bb bg gg).  getEnergyGradient (sphericalAlignment.py:93-96) returns
 * (-value, -grad). */
int fo_sph_overlap_gradient(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax,
                            const double* euler, double* value, double* grad, double* hess);

/* fo_sph_align_pairs / fo_sph_align_bank followed, on the device, by that refinement of the
 * interpolated grid maximum of every (pair, orientation): the numpy classes' findRotation
 * (sphericalAlignment.py:190-194) for whole batches.  euler [P,O,3] are the refined Euler angles,
 * overlap [P,O] the refined un-weighted overlap (= -res.fun): the numpy orientation rule keeps the
 * orientation with the larger overlap (sphericalAlignment.py:178-187). */
int fo_sph_align_pairs_refined(fo_ctx* ctx, const double* posA, const double* posB, int64_t npairs,
                               int64_t natoms, int64_t Jmax, double sigma, int invert, int64_t* best_idx,
                               double* best_val, double* frac_idx, double* euler, double* overlap,
                               int32_t* status);
int fo_sph_align_pairs_refined_dev(fo_ctx* ctx, const double* d_posA, const double* d_posB,
                                   int64_t npairs, int64_t natoms, int64_t Jmax, double sigma, int invert,
                                   int64_t* d_best_idx, double* d_best_val, double* d_frac_idx,
                                   double* d_euler, double* d_overlap, int32_t* d_status);
int fo_sph_align_bank_refined(fo_ctx* ctx, const fo_bank* bank, const int64_t* pairs /*[P,2]*/,
                              int64_t npairs, int invert, int64_t* best_idx, double* best_val,
                              double* frac_idx, double* avg_overlap, double* euler, double* overlap);

/* Wigner-d table of the reference's SOFT object: Ds[l, m1, m2, k] = sqrt((2l+1)/2)
 * d^l_{m1m2}(beta_k), out [B, 2B-1, 2B-1, 2B] doubles, negative m wrapped modulo 2B-1
 * (soft.py:73-96, CALCWIGNERD DSOFT.f90:121-195).  Computed on the device. */
int fo_sph_wigner_table(fo_ctx* ctx, int64_t Jmax, double* out);

/* ---------------------------------------------------------------- top-k peaks (a8)
 * Top-npeaks peaks of P overlap grids [P][n0][n1][n2] by the reference's Gaussian fit-and-subtract,
 * entirely on the device: findPeaks / fitPeak / _gaussian (utils.py:347-396), FINDPEAKS / FINDPEAK /
 * FIT / GAUSSIAN (fastutils.f90:231-548).  Per grid: f = a - min(a); repeat {arg-max; fit
 * A exp(-(x-x0)^T S (x-x0)) + mu to the (2 width + 1)^3 periodic window (Levenberg-Marquardt from
 * (f[ind], 0, identity, 0)); record; subtract the fitted function from the whole grid}.
 * Outputs per grid: peaks [npeaks][3] fractional grid indices (x0 + ind), amplitude [npeaks] (A),
 * mean [npeaks] (mu, nullable), alpha [npeaks][6] (upper triangle of S: s00 s01 s02 s11 s12 s22,
 * nullable; the reference's sigma is (2 alpha)^-1/2), nfound (a failed fit ends the search, as the
 * reference's `except RuntimeError: break`; entries beyond nfound are NaN), residual (nullable): the
 * grid after the subtractions (the reference's returned `f`).  width <= 4, npeaks <= 64. */
int fo_grid_find_peaks(fo_ctx* ctx, const double* grids, int64_t P, const int64_t shape[3], int64_t npeaks,
                       int64_t width, double* peaks, double* amplitude, double* mean, double* alpha,
                       int32_t* nfound, double* residual);
/* Device-resident variant: d_grids is updated in place (becomes the residual). */
int fo_grid_find_peaks_dev(fo_ctx* ctx, double* d_grids, int64_t P, const int64_t shape[3], int64_t npeaks,
                           int64_t width, double* d_peaks, double* d_amplitude, double* d_mean,
                           double* d_alpha, int32_t* d_nfound);
/* Fused forms: the overlap grid is produced and searched on the device and never copied to the host.
 * fo_sph_isoft_peaks: coefficients Ilmm [P][L+1][2L+1][2L+1] complex -> (2L+2)^3 grid -> peaks
 * (findRotations, sphericalAlignment.py:196-204; ALIGN with nrotations > 1, fastclusters.f90:129).
 * fo_per_align_pairs_peaks: positions -> F^3 |f| grid -> peaks (findDisps with npeaks > 1,
 * periodicAlignment.py:442-451; ALIGN with ndisplacements > 1, fastbulk.f90:279). */
int fo_sph_isoft_peaks(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax, int64_t npeaks,
                       int64_t width, double* peaks, double* amplitude, double* mean, double* alpha,
                       int32_t* nfound);
int fo_per_align_pairs_peaks(fo_ctx* ctx, const fo_per_params* p, const double* posA, const double* posB,
                             int64_t npairs, int64_t npeaks, int64_t width, double* peaks, double* amplitude,
                             double* mean, double* alpha, int32_t* nfound);

/* ---------------------------------------------------------------- host refinement (no GPU)
 * The steps the north star keeps on the host, as native code on an OpenMP pool over pairs
 * (nthreads <= 0: all cores).  They take the hot-path outputs (frac_idx / Euler angles) directly. */

/* Periodic: permutation (Jonker-Volgenant LAP on the minimum-image distance matrix, per group) <->
 * mean-displacement iteration, at most niter rounds.  Well-aligned pairs never build the matrix: a
 * single-precision pass proves the nearest-partner assignment optimal, and the loop's confirming second
 * solve is skipped when the displacement update is too small to change it (results bit-identical).  Replaces BasePeriodicAlignment.refine
 * (periodicAlignment.py:27-80) / ITERATIVEALIGN(bulk) (alignutils.f90:111-286).
 * frac_idx [P,3] from fo_per_align_pairs; dist [P]; perm [P,N] (nullable): X2 = posB[perm] - disp;
 * disp [P,3] (nullable). */
int fo_host_refine_periodic(const fo_per_params* p, const int32_t* group_offsets, int64_t ngroups,
                            const int32_t* atom_idx, const double* posA, const double* posB,
                            const double* frac_idx, int64_t npairs, int niter, int nthreads, double* dist,
                            int32_t* perm, double* disp);

/* Same for the pairs pair_idx[0..nidx) only (indices into posA / posB / frac_idx / the outputs): the host stage
 * of fo_per_align_pairs_full for the pairs its device screening flags. */
int fo_host_refine_periodic_subset(const fo_per_params* p, const int32_t* group_offsets, int64_t ngroups,
                                   const int32_t* atom_idx, const double* posA, const double* posB,
                                   const double* frac_idx, const int64_t* pair_idx, int64_t nidx, int niter,
                                   int nthreads, double* dist, int32_t* perm, double* disp);

/* Work counters of fo_host_refine_periodic since load (summed over calls and threads), for tests and
 * profiling: out[0] assignments solved by the double-precision matrix + LAP (per group), out[1] assignments
 * settled by the single-precision screening (column minima proven to be the unique optimum), out[2] repeat
 * solves of the permutation <-> displacement loop skipped because the displacement moved by less than half
 * the stability margin of the last assignment.  reset != 0 zeroes them after reading. */
void fo_host_refine_counters(int64_t out[3], int reset);

/* Test hook of the host pool: instruction set of the assignment kernels (cost matrix, row scan).  isa: 0 = the widest
 * the CPU has (the default), 1 = 2-wide vectors, 2 = AVX2, 3 = AVX-512, < 0 = query only.  Returns the active one, or
 * -1 when the requested one is not available.  Results do not depend on it.  Not thread safe. */
int fo_host_lap_isa(int isa);

/* Clusters: for each orientation o < norient rotate (o = 1: -posB) by the Euler angles
 * euler[P,norient,3], permute (LAP on squared distances per group), Kearsley quaternion fit; keep
 * the orientation with the smaller distance.  Replaces BaseSphericalAlignment.refine
 * (sphericalAlignment.py:118-127) with the Fortran orientation rule (fastclusters.f90:243-254).
 * posA / posB centred; dist [P]; orient [P], perm [P,N], rmat [P,9] nullable. */
int fo_host_refine_spherical(const double* posA, const double* posB, int64_t npairs, int64_t natoms,
                             const int32_t* group_offsets, int64_t ngroups, const int32_t* atom_idx,
                             const double* euler, int norient, int nthreads, double* dist, int32_t* orient,
                             int32_t* perm, double* rmat);

/* Same with permutation hints: perm_hint [P,norient,N] and hint_ok [P,norient]; where hint_ok is non-zero the
 * LAP is skipped and perm_hint used (the device screening of fo_sph_align_pairs_full proved it optimal). */
int fo_host_refine_spherical_hint(const double* posA, const double* posB, int64_t npairs, int64_t natoms,
                                  const int32_t* group_offsets, int64_t ngroups, const int32_t* atom_idx,
                                  const double* euler, int norient, const int32_t* perm_hint,
                                  const int32_t* hint_ok, int nthreads, double* dist, int32_t* orient,
                                  int32_t* perm, double* rmat);

/* The two host steps on their own (one pair), for the single-pair drop-in classes.
 * fo_host_best_permutation: perm [N] such that posB[perm] best matches posA group by group -- minimum-image
 * distance costs when box != NULL (periodicAlignment.py:82-110), squared distances otherwise
 * (find_best_permutation, utils.py:82-167).
 * fo_host_kearsley: distance after the optimal rotation of x2 onto x1 (both re-centred) and the rotation
 * matrix rmat [9] (nullable), aligned = x2 . rmat^T (findrotation, utils.py:169-253; alignutils.f90:304-379). */
int fo_host_best_permutation(const double* posA, const double* posB, int64_t natoms,
                             const int32_t* group_offsets, int64_t ngroups, const int32_t* atom_idx,
                             const double* box /*[3] or NULL*/, int32_t* perm);
int fo_host_kearsley(const double* x1, const double* x2, int64_t natoms, double* dist, double* rmat);

#ifdef __cplusplus
}
#endif
#endif /* FASTOVERLAP_B200_H */
