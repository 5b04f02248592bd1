/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the reference's PERIODIC
 * overlap-maximisation path.  Nothing under fastoverlap_b200/ may call this; it is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
 *
 * Parity status: PINNED.  tests/test_oracle_periodic.py checks every function here against
 * vectors produced by the unmodified reference (oracle/make_golden.py -> tests/golden/), incl.
 * the reference's own known answer for examples/BLJ256 (periodicAlignment.py:609: 1.559).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference/).
 * Plain scalar C, one pair at a time like the reference; OpenMP only over independent pairs.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

/* fastoverlap/utils.py:278-313 (_next_fast_len) == FASTLEN table fastoverlap/f90/fastutils.f90:78-90 */
int64_t oracle_next_fast_len(int64_t target) {
  if (target <= 6) return target;
  if (!(target & (target - 1))) return target;
  int64_t match = INT64_MAX;
  int64_t p5 = 1;
  while (p5 < target) {
    int64_t p35 = p5;
    while (p35 < target) {
      int64_t quotient = (target + p35 - 1) / p35; /* ceil */
      int64_t p2 = 1;
      while (p2 < quotient) p2 *= 2;
      int64_t N = p2 * p35;
      if (N == target) return N;
      if (N < match) match = N;
      p35 *= 3;
      if (p35 == target) return p35;
    }
    if (p35 < match) match = p35;
    p5 *= 5;
    if (p5 == target) return p5;
  }
  if (p5 < match) match = p5;
  return match;
}

/* Structure factors of one structure, one array per permutation group.
 * fastoverlap/f90/fastbulk.f90:599-633 (PERIODICFOURIER: loop over the k grid, inner loop over
 * atoms, KR = sum_k coords*wavek, accumulate EXP(DCMPLX(0,-KR))) and :635-665
 * (PERIODICFOURIERPERM: gather the atoms of each group first);
 * wave vectors as SETWAVEK fastbulk.f90:567-597 / periodicAlignment.py:384-386.
 * out[g][ix][iy][iz], index 0 <-> k = -n (numpy layout of calcFourierCoeff,
 * periodicAlignment.py:400-406). */
void oracle_per_structure_factors(const double* pos, int64_t natoms, const int32_t* goff,
                                  int64_t ngroups, const int32_t* gidx, const double box[3],
                                  int64_t n, cplx* out) {
  (void)natoms;
  const int64_t W = 2 * n + 1;
  const double twopi = 6.283185307179586;
  const double kx = twopi / box[0], ky = twopi / box[1], kz = twopi / box[2];
  for (int64_t g = 0; g < ngroups; ++g) {
    cplx* o = out + g * W * W * W;
    for (int64_t ix = 0; ix < W; ++ix)
      for (int64_t iy = 0; iy < W; ++iy)
        for (int64_t iz = 0; iz < W; ++iz) {
          const double wx = kx * (double)(ix - n), wy = ky * (double)(iy - n),
                       wz = kz * (double)(iz - n);
          cplx acc = 0.0;
          for (int32_t a = goff[g]; a < goff[g + 1]; ++a) {
            const double* r = pos + 3 * (int64_t)gidx[a];
            double kr = 0.0;
            kr += r[0] * wx;
            kr += r[1] * wy;
            kr += r[2] * wz;
            acc += cexp(-I * kr);
          }
          o[(ix * W + iy) * W + iz] = acc;
        }
  }
}

/* C(k) = sum_g C1_g(k) conj(C2_g(k)) exp(-|k|^2 sigma^2)
 * fastoverlap/periodicAlignment.py:437-438; Fortran: conj(A), both times exp(-k^2 sigma^2/2),
 * then the product summed over groups, fastbulk.f90:441-454 and :667-683 (DOTFOURIERCOEFFS). */
void oracle_per_cross_spectrum(const cplx* C1, const cplx* C2, int64_t ngroups, const double box[3],
                               int64_t n, double sigma, cplx* C) {
  const int64_t W = 2 * n + 1, W3 = W * W * W;
  const double twopi = 6.283185307179586;
  for (int64_t ix = 0; ix < W; ++ix)
    for (int64_t iy = 0; iy < W; ++iy)
      for (int64_t iz = 0; iz < W; ++iz) {
        const double wx = twopi / box[0] * (double)(ix - n), wy = twopi / box[1] * (double)(iy - n),
                     wz = twopi / box[2] * (double)(iz - n);
        const double absk = sqrt(wx * wx + wy * wy + wz * wz); /* norm(ks, axis=0), :386 */
        const double damp = exp(-(absk * absk) * (sigma * sigma));
        const int64_t e = (ix * W + iy) * W + iz;
        cplx acc = 0.0;
        for (int64_t g = 0; g < ngroups; ++g) acc += C1[g * W3 + e] * conj(C2[g * W3 + e]) * damp;
        C[e] = acc;
      }
}

/* Csum: self-overlap normaliser, fastoverlap/periodicAlignment.py:433-436 (factor :380). */
double oracle_per_csum(const cplx* C1, const cplx* C2, int64_t ngroups, const double box[3],
                       int64_t n, double sigma) {
  const int64_t W = 2 * n + 1, W3 = W * W * W;
  const double twopi = 6.283185307179586;
  const double factor = 2.0 * pow(M_PI * sigma * sigma, -1.5) * sigma * sigma / (box[0] * box[1] * box[2]);
  double s = 0.0;
  for (int64_t g = 0; g < ngroups; ++g)
    for (int64_t ix = 0; ix < W; ++ix)
      for (int64_t iy = 0; iy < W; ++iy)
        for (int64_t iz = 0; iz < W; ++iz) {
          const double wx = twopi / box[0] * (double)(ix - n), wy = twopi / box[1] * (double)(iy - n),
                       wz = twopi / box[2] * (double)(iz - n);
          const double absk = sqrt(wx * wx + wy * wy + wz * wz);
          const int64_t e = g * W3 + (ix * W + iy) * W + iz;
          s += (creal(C1[e] * conj(C1[e])) + creal(C2[e] * conj(C2[e]))) *
               exp(-(absk * absk) * (sigma * sigma));
        }
  return s * 0.5 * factor;
}

/* ---- a plain mixed-radix complex FFT (stands in for FFTW / pocketfft, which the reference calls:
 * fastoverlap/f90/fastutils.f90:554-569 FFT3D = dfftw_plan_dft_3d FFTW_FORWARD;
 * numpy: np.fft.fftn periodicAlignment.py:439).  sign = -1 forward, +1 backward, unnormalised. */
static void fft_rec(int64_t n, const cplx* in, int64_t istride, cplx* out, const cplx* tw,
                    int64_t N, int64_t tstep, cplx* tmp) {
  if (n == 1) {
    out[0] = in[0];
    return;
  }
  int64_t p = 2;
  while (n % p) ++p; /* smallest prime factor */
  const int64_t m = n / p;
  for (int64_t r = 0; r < p; ++r) fft_rec(m, in + r * istride, istride * p, out + r * m, tw, N, tstep * p, tmp);
  for (int64_t k = 0; k < m; ++k) {
    for (int64_t q = 0; q < p; ++q) {
      cplx acc = 0.0;
      for (int64_t r = 0; r < p; ++r) {
        const int64_t e = (r * (k + q * m)) % n;
        acc += out[r * m + k] * tw[(e * tstep) % N];
      }
      tmp[q] = acc;
    }
    for (int64_t q = 0; q < p; ++q) out[k + q * m] = tmp[q];
  }
}

/* transform with a caller-provided twiddle table tw[t] = exp(sign 2 pi i t/n) (plan reuse) */
void oracle_fft_tw(int64_t n, const cplx* in, int64_t istride, cplx* out, const cplx* tw, cplx* tmp) {
  fft_rec(n, in, istride, out, tw, n, 1, tmp);
}

void oracle_fft1d(int64_t n, const cplx* in, int64_t istride, cplx* out, int sign) {
  cplx* tw = (cplx*)malloc(sizeof(cplx) * (size_t)n);
  cplx* tmp = (cplx*)malloc(sizeof(cplx) * (size_t)(n > 16 ? n : 16));
  for (int64_t t = 0; t < n; ++t) tw[t] = cexp(sign * I * (6.283185307179586476925286766559 * (double)t / (double)n));
  fft_rec(n, in, istride, out, tw, n, 1, tmp);
  free(tw);
  free(tmp);
}

/* f = fftn(C, (F,F,F)): zero-pad each axis to F in turn and transform it (what numpy's fftn
 * does, axis by axis from the last); fabs = |f|.  fastoverlap/periodicAlignment.py:439-440;
 * Fortran ALIGNCOEFFS fastbulk.f90:522-523.  The k=-n corner sits at index 0. */
void oracle_per_fft_abs(const cplx* C, int64_t n, int64_t F, double* fabs_out, cplx* f_out /*nullable*/) {
  const int64_t W = 2 * n + 1;
  cplx* tw = (cplx*)malloc(sizeof(cplx) * (size_t)F);
  cplx* tmp = (cplx*)malloc(sizeof(cplx) * (size_t)(F > 16 ? F : 16));
  cplx* line = (cplx*)calloc((size_t)F, sizeof(cplx));
  cplx* res = (cplx*)malloc(sizeof(cplx) * (size_t)F);
  cplx* A = (cplx*)malloc(sizeof(cplx) * (size_t)(W * W * F));
  cplx* B = (cplx*)malloc(sizeof(cplx) * (size_t)(W * F * F));
  cplx* D = (cplx*)malloc(sizeof(cplx) * (size_t)(F * F * F));
  for (int64_t t = 0; t < F; ++t) tw[t] = cexp(-I * (6.283185307179586476925286766559 * (double)t / (double)F));
  /* axis 2 */
  for (int64_t ix = 0; ix < W; ++ix)
    for (int64_t iy = 0; iy < W; ++iy) {
      memset(line, 0, sizeof(cplx) * (size_t)F);
      for (int64_t iz = 0; iz < W; ++iz) line[iz] = C[(ix * W + iy) * W + iz];
      fft_rec(F, line, 1, res, tw, F, 1, tmp);
      memcpy(A + (ix * W + iy) * F, res, sizeof(cplx) * (size_t)F);
    }
  /* axis 1 */
  for (int64_t ix = 0; ix < W; ++ix)
    for (int64_t dz = 0; dz < F; ++dz) {
      memset(line, 0, sizeof(cplx) * (size_t)F);
      for (int64_t iy = 0; iy < W; ++iy) line[iy] = A[(ix * W + iy) * F + dz];
      fft_rec(F, line, 1, res, tw, F, 1, tmp);
      for (int64_t dy = 0; dy < F; ++dy) B[(ix * F + dy) * F + dz] = res[dy];
    }
  /* axis 0 */
  for (int64_t dy = 0; dy < F; ++dy)
    for (int64_t dz = 0; dz < F; ++dz) {
      memset(line, 0, sizeof(cplx) * (size_t)F);
      for (int64_t ix = 0; ix < W; ++ix) line[ix] = B[(ix * F + dy) * F + dz];
      fft_rec(F, line, 1, res, tw, F, 1, tmp);
      for (int64_t dx = 0; dx < F; ++dx) D[(dx * F + dy) * F + dz] = res[dx];
    }
  for (int64_t e = 0; e < F * F * F; ++e) fabs_out[e] = cabs(D[e]);
  if (f_out) memcpy(f_out, D, sizeof(cplx) * (size_t)(F * F * F));
  free(tw); free(tmp); free(line); free(res); free(A); free(B); free(D);
}

/* findMax: flat arg-max (first in C order), then a 3-point parabola per axis on |a| with
 * periodic wrap.  fastoverlap/utils.py:319-338.  (Fortran: MAXLOC fastutils.f90:420.) */
void oracle_find_max(const double* a, const int64_t shape[3], int64_t idx_out[3], double frac_out[3]) {
  const int64_t n0 = shape[0], n1 = shape[1], n2 = shape[2];
  int64_t best = 0;
  for (int64_t e = 1; e < n0 * n1 * n2; ++e)
    if (a[e] > a[best]) best = e;
  int64_t ind[3] = {best / (n1 * n2), (best / n2) % n1, best % n2};
  for (int ax = 0; ax < 3; ++ax) {
    int64_t ip[3] = {ind[0], ind[1], ind[2]}, im[3] = {ind[0], ind[1], ind[2]};
    ip[ax] = (ind[ax] + 1) % shape[ax];
    im[ax] = (ind[ax] - 1 + shape[ax]) % shape[ax]; /* python negative index wraps */
    const double y1 = fabs(a[(ip[0] * n1 + ip[1]) * n2 + ip[2]]);
    const double y2 = fabs(a[best]);
    const double y3 = fabs(a[(im[0] * n1 + im[1]) * n2 + im[2]]);
    const double d = (y3 - y1) / (2 * (2 * y2 - y1 - y3));
    idx_out[ax] = ind[ax];
    frac_out[ax] = (double)ind[ax] - d;
  }
}

/* The periodic hot path for one pair: coords -> fabs grid -> arg-max.
 * PeriodicAlign.setPos + findDisps(npeaks=1), fastoverlap/periodicAlignment.py:408-456;
 * ALIGN1 + ALIGNCOEFFS up to FINDPEAKS, fastoverlap/f90/fastbulk.f90:414-531. */
void oracle_per_align_pair(const double* posA, const double* posB, int64_t natoms,
                           const int32_t* goff, int64_t ngroups, const int32_t* gidx,
                           const double box[3], int64_t n, int64_t F, double sigma,
                           int64_t best_idx[3], double* best_val, double frac_idx[3],
                           double* grid_out /*nullable [F^3]*/) {
  const int64_t W = 2 * n + 1, W3 = W * W * W;
  cplx* C1 = (cplx*)malloc(sizeof(cplx) * (size_t)(ngroups * W3));
  cplx* C2 = (cplx*)malloc(sizeof(cplx) * (size_t)(ngroups * W3));
  cplx* C = (cplx*)malloc(sizeof(cplx) * (size_t)W3);
  double* fab = grid_out ? grid_out : (double*)malloc(sizeof(double) * (size_t)(F * F * F));
  oracle_per_structure_factors(posA, natoms, goff, ngroups, gidx, box, n, C1);
  oracle_per_structure_factors(posB, natoms, goff, ngroups, gidx, box, n, C2);
  oracle_per_cross_spectrum(C1, C2, ngroups, box, n, sigma, C);
  oracle_per_fft_abs(C, n, F, fab, NULL);
  const int64_t shape[3] = {F, F, F};
  oracle_find_max(fab, shape, best_idx, frac_idx);
  *best_val = fab[(best_idx[0] * F + best_idx[1]) * F + best_idx[2]];
  free(C1); free(C2); free(C);
  if (!grid_out) free(fab);
}

/* P independent pairs, OpenMP over pairs (the reference itself is single-threaded; the fan-out
 * over host cores is the "all host threads" CPU baseline of bench.py). Returns threads used. */
int oracle_per_align_pairs(const double* posA, const double* posB, int64_t npairs, int64_t natoms,
                           const int32_t* goff, int64_t ngroups, const int32_t* gidx,
                           const double box[3], int64_t n, int64_t F, double sigma,
                           int64_t* best_idx, double* best_val, double* frac_idx, int nthreads) {
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < npairs; ++p)
    oracle_per_align_pair(posA + p * natoms * 3, posB + p * natoms * 3, natoms, goff, ngroups, gidx,
                          box, n, F, sigma, best_idx + 3 * p, best_val + p, frac_idx + 3 * p, NULL);
  return used;
}
