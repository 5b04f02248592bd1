/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the reference's SPHERICAL (cluster)
 * overlap-maximisation path.  Nothing under fastoverlap_b200/ may call this; it is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
 *
 * Parity status: PINNED.  tests/test_oracle_spherical.py checks every function against vectors
 * produced by the unmodified reference (oracle/make_golden.py -> tests/golden/), including the
 * reference's own known answer for examples/LJ38 (sphericalAlignment.py:692: 1.4767).
 * One caveat, documented in DESIGN.md: for the harmonic-basis radial integrals the reference's
 * own two formulations (numpy hyp1f1 sum, Fortran recurrence) disagree with each other at the
 * 1e-8..1e-7 level for nmax=20; oracle_sph_harm_radial_exact is the closed form
 * both approximate (validated against mpmath quadrature to 60 digits) and
 * oracle_sph_harm_radial_fortran restates the Fortran recurrence.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference/).
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

void oracle_fft1d(int64_t n, const cplx* in, int64_t istride, cplx* out, int sign);
void oracle_fft_tw(int64_t n, const cplx* in, int64_t istride, cplx* out, const cplx* tw, cplx* tmp);
void oracle_find_max(const double* a, const int64_t shape[3], int64_t idx_out[3], double frac_out[3]);

static inline int64_t wrapm(int64_t m, int64_t n) { return ((m % n) + n) % n; }

/* Spherical harmonics Y_lm(theta, phi) of one point for l <= L, scipy/Condon-Shortley convention,
 * stored Y[l*(2L+1) + (m mod 2L+1)] (numpy negative-index wrap, sphericalAlignment.py:57-65).
 * Follows RYML fastoverlap/f90/fastclusters.f90:602-652: r, phi = atan2(y,x), z = cos(theta),
 * normalised associated Legendre functions (the quantity XDNRMP legendre.f90:143-371 returns,
 * here by the standard stable 3-term recurrence in l), Y_l,-m = (-1)^m conj(Y_lm), times
 * exp(i m phi).  Returns r. */
double oracle_sph_ylm(const double pos[3], int64_t L, cplx* Y) {
  const int64_t NJ = 2 * L + 1;
  const double r = sqrt(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]);
  const double phi = atan2(pos[1], pos[0]);
  const double ct = pos[2] / r;
  const double st = sqrt(fmax(0.0, 1.0 - ct * ct));
  memset(Y, 0, sizeof(cplx) * (size_t)((L + 1) * NJ));
  double* P = (double*)malloc(sizeof(double) * (size_t)((L + 1) * (L + 1)));
  /* P[l*(L+1)+m] = sqrt((2l+1)/(4pi) (l-m)!/(l+m)!) P_l^m(ct), with Condon-Shortley phase */
  double pmm = sqrt(1.0 / (4.0 * M_PI));
  for (int64_t m = 0; m <= L; ++m) {
    if (m > 0) pmm *= -sqrt((2.0 * m + 1.0) / (2.0 * m)) * st;
    P[m * (L + 1) + m] = pmm;
    if (m + 1 <= L) P[(m + 1) * (L + 1) + m] = sqrt(2.0 * m + 3.0) * ct * pmm;
    for (int64_t l = m + 2; l <= L; ++l) {
      const double a = sqrt((4.0 * l * l - 1.0) / ((double)(l * l - m * m)));
      const double b = sqrt((((double)(l - 1) * (l - 1)) - (double)(m * m)) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
      P[l * (L + 1) + m] = a * (ct * P[(l - 1) * (L + 1) + m] - b * P[(l - 2) * (L + 1) + m]);
    }
  }
  for (int64_t l = 0; l <= L; ++l)
    for (int64_t m = 0; m <= l; ++m) {
      const cplx e = cexp(I * ((double)m * phi));
      const cplx y = P[l * (L + 1) + m] * e;
      Y[l * NJ + m] = y;
      if (m > 0) Y[l * NJ + wrapm(-m, NJ)] = ((m & 1) ? -1.0 : 1.0) * conj(y);
    }
  free(P);
  return r;
}

/* exp(-x) i_l(x), l = 0..L: modified spherical Bessel functions of the first kind.
 * Follows SPHI fastoverlap/f90/fastutils.f90:1000-1096 (Zhang & Jin: i_0 = sinh x / x, backward
 * recurrence f_{k} = (2k+3) f_{k+1}/x + f_{k+2} from a start order above L, rescaled to i_0);
 * numpy: iv(l+1/2, x) sqrt(pi/2x), sphericalAlignment.py:270-271.  The exp(-x) scaling only moves
 * the Gaussian factor (see oracle_sph_coeffs_direct) so that large clusters do not overflow. */
void oracle_sphi_scaled(int64_t L, double x, double* si) {
  if (fabs(x) < 1e-100) {
    for (int64_t k = 0; k <= L; ++k) si[k] = 0.0;
    si[0] = 1.0;
    return;
  }
  const double si0 = -expm1(-2.0 * x) / (2.0 * x); /* exp(-x) sinh(x)/x */
  /* start order: high enough that the arbitrary start has decayed below 1e-17 at order L */
  int64_t m = L + 16 + (int64_t)sqrt(50.0 * x + (double)(L * L)) - L;
  double f0 = 0.0, f1 = 1e-100, f = 0.0;
  for (int64_t k = m; k >= 0; --k) {
    f = (2.0 * k + 3.0) * f1 / x + f0;
    if (k <= L) si[k] = f;
    f0 = f1;
    f1 = f;
    if (fabs(f) > 1e200) { /* rescale to stay in range (small x) */
      f0 *= 1e-200;
      f1 *= 1e-200;
      f *= 1e-200;
      for (int64_t q = k; q <= L; ++q) si[q] *= 1e-200;
    }
  }
  const double cs = si0 / f;
  for (int64_t k = 0; k <= L; ++k) si[k] *= cs;
}

/* Direct SO(3) Fourier coefficients of the overlap of two centred structures, summed over
 * permutation groups:
 *   I[l,m1,m2] = 4 pi^2.5 s^3 sum_g sum_{j,k in g} i_l(r_j r_k/2s^2) e^{-(r_j^2+r_k^2)/4s^2}
 *                Y_lm1(A_j) conj(Y_lm2(B_k))
 * numpy SphericalAlign.calcSO3Coeffs fastoverlap/sphericalAlignment.py:260-273 (+ the sum over
 * groups :175); Fortran FOURIERCOEFFS fastoverlap/f90/fastclusters.f90:868-916 -- same pair loop
 * (cost (1/3)(L+1)(2L+1)(2L+3) N^2, :143-145) but WITHOUT its 4 sigma pair cut-off :900-901,
 * which numpy does not have.  A = pos1 (unconjugated), B = pos2 (conjugated).
 * Output I[l][m1 mod 2L+1][m2 mod 2L+1] (numpy layout). */
void oracle_sph_coeffs_direct(const double* posA, const double* posB, int64_t natoms,
                              const int32_t* goff, int64_t ngroups, const int32_t* gidx, int64_t L,
                              double sigma, cplx* Ilmm) {
  const int64_t NJ = 2 * L + 1;
  cplx* YA = (cplx*)malloc(sizeof(cplx) * (size_t)(natoms * (L + 1) * NJ));
  cplx* YB = (cplx*)malloc(sizeof(cplx) * (size_t)(natoms * (L + 1) * NJ));
  double* rA = (double*)malloc(sizeof(double) * (size_t)natoms);
  double* rB = (double*)malloc(sizeof(double) * (size_t)natoms);
  double* il = (double*)malloc(sizeof(double) * (size_t)(L + 1));
  for (int64_t i = 0; i < natoms; ++i) {
    rA[i] = oracle_sph_ylm(posA + 3 * i, L, YA + i * (L + 1) * NJ);
    rB[i] = oracle_sph_ylm(posB + 3 * i, L, YB + i * (L + 1) * NJ);
  }
  const double fact = 4.0 * pow(M_PI, 2.5) * sigma * sigma * sigma;
  memset(Ilmm, 0, sizeof(cplx) * (size_t)((L + 1) * NJ * NJ));
  for (int64_t g = 0; g < ngroups; ++g)
    for (int32_t ja = goff[g]; ja < goff[g + 1]; ++ja)
      for (int32_t kb = goff[g]; kb < goff[g + 1]; ++kb) {
        const int64_t j = gidx[ja], k = gidx[kb];
        const double x = 0.5 * rA[j] * rB[k] / (sigma * sigma);
        oracle_sphi_scaled(L, x, il);
        const double dr = rA[j] - rB[k];
        const double tmp = fact * exp(-0.25 * dr * dr / (sigma * sigma)); /* e^{x} e^{-(ra^2+rb^2)/4s^2} */
        const cplx* ya = YA + j * (L + 1) * NJ;
        const cplx* yb = YB + k * (L + 1) * NJ;
        for (int64_t l = 0; l <= L; ++l) {
          const double w = il[l] * tmp;
          for (int64_t m1 = -l; m1 <= l; ++m1) {
            const cplx a = w * ya[l * NJ + wrapm(m1, NJ)];
            cplx* row = Ilmm + (l * NJ + wrapm(m1, NJ)) * NJ;
            for (int64_t m2 = -l; m2 <= l; ++m2) row[wrapm(m2, NJ)] += a * conj(yb[l * NJ + wrapm(m2, NJ)]);
          }
        }
      }
  free(YA); free(YB); free(rA); free(rB); free(il);
}

/* Quadrature weights of the SOFT sampling: MAKEWEIGHTS fastoverlap/f90/DSOFT.f90:61-79,
 * soft.py:64-71. */
void oracle_soft_weights(int64_t B, double* w) {
  const double fudge = M_PI / 4.0 / (double)B;
  for (int64_t j = 0; j < 2 * B; ++j) {
    double acc = 0.0;
    const double sinj = 2.0 * sin((2.0 * j + 1.0) * fudge) / (double)B;
    for (int64_t k = 0; k < B; ++k) acc += sinj * sin((2.0 * j + 1.0) * (2.0 * k + 1.0) * fudge) / (2.0 * k + 1.0);
    w[j] = acc;
  }
}

/* Coefficients of the 3-term recurrence: RECURRTERMS fastoverlap/f90/DSOFT.f90:81-119. */
static void recurr_terms(int64_t J, int64_t M1, int64_t M2, double* A, double* Bc, double* C) {
  const double dj = (double)J, dm1 = (double)M1, dm2 = (double)M2;
  const double t1 = sqrt((2.0 * dj + 3.0) / (2.0 * dj + 1.0));
  const double t3 = (dj + 1.0) * (2.0 * dj + 1.0);
  const double t5 = 1.0 / sqrt(((dj + 1.0) * (dj + 1.0) - dm1 * dm1) * ((dj + 1.0) * (dj + 1.0) - dm2 * dm2));
  *Bc = t1 * t3 * t5;
  if (J == 0) {
    *A = 0.0;
    *C = 0.0;
  } else {
    const double t2 = sqrt((2.0 * dj + 3.0) / (2.0 * dj - 1.0)) * (dj + 1.0) / dj;
    const double t4 = sqrt((dj * dj - dm1 * dm1) * (dj * dj - dm2 * dm2));
    *A = t2 * t4 * t5;
    *C = dm1 * dm2 / (dj * (dj + 1.0));
  }
}

/* Normalised Wigner little-d table Ds[l][m1][m2][k] = sqrt((2l+1)/2) d^l_{m1 m2}(beta_k),
 * beta_k = pi (2k+1)/(4B), m wrapped modulo 2B-1: CALCWIGNERD fastoverlap/f90/DSOFT.f90:121-195
 * (Kostelec-Rockmore: closed-form edge values d^J_{+-J,m}, d^J_{m,+-J}, then the recurrence in
 * J); numpy closed form soft.py:73-96 gives the same table. */
void oracle_wigner_table(int64_t B, double* Ds) {
  const int64_t NM = 2 * B - 1, NK = 2 * B;
  const double fudge = M_PI / 4.0 / (double)B;
  double* cosb = (double*)malloc(sizeof(double) * (size_t)NK);
  double* cosb2 = (double*)malloc(sizeof(double) * (size_t)NK);
  double* sinb2 = (double*)malloc(sizeof(double) * (size_t)NK);
  double* fct = (double*)malloc(sizeof(double) * (size_t)(3 * B));
  for (int64_t i = 0; i < NK; ++i) {
    const double beta = fudge * (2.0 * i + 1.0);
    cosb[i] = cos(beta);
    cosb2[i] = cos(beta / 2);
    sinb2[i] = sin(beta / 2);
  }
  fct[0] = 1.0;
  for (int64_t i = 1; i < 3 * B; ++i) fct[i] = (double)i * fct[i - 1];
  memset(Ds, 0, sizeof(double) * (size_t)(B * NM * NM * NK));
#define DS(l, a, b, k) Ds[(((l) * NM + wrapm((a), NM)) * NM + wrapm((b), NM)) * NK + (k)]
  for (int64_t M1 = -(B - 1); M1 <= B - 1; ++M1)
    for (int64_t J = llabs(M1); J <= B - 1; ++J) {
      const double factor = sqrt((2.0 * J + 1.0) * fct[2 * J] / fct[J + M1] / fct[J - M1] / 2.0);
      for (int64_t i = 0; i < NK; ++i) {
        DS(J, J, M1, i) = factor * pow(cosb2[i], (double)(J + M1)) * pow(-sinb2[i], (double)(J - M1));
        DS(J, -J, M1, i) = factor * pow(cosb2[i], (double)(J - M1)) * pow(sinb2[i], (double)(J + M1));
        DS(J, M1, J, i) = factor * pow(cosb2[i], (double)(J + M1)) * pow(sinb2[i], (double)(J - M1));
        DS(J, M1, -J, i) = factor * pow(cosb2[i], (double)(J - M1)) * pow(-sinb2[i], (double)(J + M1));
      }
    }
  for (int64_t M2 = -(B - 2); M2 <= B - 2; ++M2)
    for (int64_t M1 = -(B - 2); M1 <= B - 2; ++M1) {
      const int64_t maxm = llabs(M1) > llabs(M2) ? llabs(M1) : llabs(M2);
      for (int64_t J = maxm; J <= B - 2; ++J) {
        double A, Bc, C;
        recurr_terms(J, M1, M2, &A, &Bc, &C);
        for (int64_t i = 0; i < NK; ++i) {
          double v = Bc * (cosb[i] - C) * DS(J, M1, M2, i);
          if (J > 0) v -= A * DS(J - 1, M1, M2, i);
          DS(J + 1, M1, M2, i) = v;
        }
      }
    }
#undef DS
  free(cosb); free(cosb2); free(sinb2); free(fct);
}

/* Inverse SO(3) Fourier transform onto the (2B)^3 Euler grid, real part:
 *   S[m1][k][m2] = sum_{l >= max(|m1|,|m2|)} Ds[l][m1][m2][k] I[l][m1][m2]
 *   out[a][k][g] = sum_{m1,m2} S[m1][k][m2] e^{+2 pi i (m1 a + m2 g)/2B}      (unnormalised)
 * ISOFT fastoverlap/f90/DSOFT.f90:265-329 (Wigner contraction :284-299, backward FFTW 1-D
 * transforms along axis 3 then axis 1 :301-325) with CALCOVERLAP's re-indexing and REAL()
 * fastoverlap/f90/fastclusters.f90:918-960; numpy SOFT.iSOFT soft.py:115-125
 * (ifft * (2B)^2 == unnormalised backward transform).  Ds may be NULL (computed here). */
void oracle_isoft(const cplx* Ilmm, int64_t B, const double* Ds_in, double* out_real, cplx* out_cplx) {
  const int64_t L = B - 1, NJ = 2 * L + 1, NK = 2 * B, NM = 2 * B - 1;
  double* Ds = (double*)Ds_in;
  if (!Ds) {
    Ds = (double*)malloc(sizeof(double) * (size_t)(B * NM * NM * NK));
    oracle_wigner_table(B, Ds);
  }
  cplx* T = (cplx*)calloc((size_t)(NK * NK * NK), sizeof(cplx)); /* T[m1 mod 2B][k][m2 mod 2B] */
  for (int64_t m1 = -L; m1 <= L; ++m1)
    for (int64_t m2 = -L; m2 <= L; ++m2) {
      const int64_t lmin = llabs(m1) > llabs(m2) ? llabs(m1) : llabs(m2);
      for (int64_t k = 0; k < NK; ++k) {
        cplx acc = 0.0;
        for (int64_t l = lmin; l <= L; ++l)
          acc += Ds[((l * NM + wrapm(m1, NM)) * NM + wrapm(m2, NM)) * NK + k] *
                 Ilmm[(l * NJ + wrapm(m1, NJ)) * NJ + wrapm(m2, NJ)];
        T[(wrapm(m1, NK) * NK + k) * NK + wrapm(m2, NK)] = acc;
      }
    }
  cplx* line = (cplx*)malloc(sizeof(cplx) * (size_t)NK);
  cplx* res = (cplx*)malloc(sizeof(cplx) * (size_t)NK);
  cplx* tw = (cplx*)malloc(sizeof(cplx) * (size_t)NK);
  cplx* tmp = (cplx*)malloc(sizeof(cplx) * (size_t)(NK > 16 ? NK : 16));
  for (int64_t t = 0; t < NK; ++t) tw[t] = cexp(I * (6.283185307179586476925286766559 * (double)t / (double)NK));
  for (int64_t a = 0; a < NK; ++a) /* axis 2 (m2 -> gamma) */
    for (int64_t k = 0; k < NK; ++k) {
      oracle_fft_tw(NK, T + (a * NK + k) * NK, 1, res, tw, tmp);
      memcpy(T + (a * NK + k) * NK, res, sizeof(cplx) * (size_t)NK);
    }
  for (int64_t k = 0; k < NK; ++k) /* axis 0 (m1 -> alpha) */
    for (int64_t g = 0; g < NK; ++g) {
      for (int64_t a = 0; a < NK; ++a) line[a] = T[(a * NK + k) * NK + g];
      oracle_fft_tw(NK, line, 1, res, tw, tmp);
      for (int64_t a = 0; a < NK; ++a) T[(a * NK + k) * NK + g] = res[a];
    }
  for (int64_t e = 0; e < NK * NK * NK; ++e) {
    if (out_real) out_real[e] = creal(T[e]);
    if (out_cplx) out_cplx[e] = T[e];
  }
  free(line); free(res); free(T); free(tw); free(tmp);
  if (!Ds_in) free(Ds);
}

/* Harmonic-basis radial overlap integrals d_nl(r), n <= N, l <= L, Fortran formulation:
 * HARMONICNL fastoverlap/f90/fastclusters.f90:492-540 (called with L+2N, :676).
 * ret[n*(Lfull+1)+l], Lfull = L + 2N as in the caller. */
void oracle_sph_harm_radial_fortran(int64_t N, int64_t Lfull, double rj, double sigma, double r0, double* ret) {
  const int64_t S = Lfull + 1;
  double r0s = 1.0 / (r0 * r0 + sigma * sigma);
  memset(ret, 0, sizeof(double) * (size_t)((N + 1) * S));
  ret[0] = sqrt(2.0 * sqrt(M_PI) * pow(r0 * r0s, 3)) * pow(sigma, 3) * exp(-0.5 * rj * rj * r0s) * 4 * M_PI;
  r0s = sqrt(2.0) * r0 * rj * r0s;
  for (int64_t j = 1; j <= Lfull; ++j) ret[j] = r0s / sqrt(1.0 + 2.0 * j) * ret[j - 1];
  const double c = sigma * sigma / rj / r0;
  if (N >= 1)
    for (int64_t j = 0; j <= Lfull - 2; ++j)
      ret[S + j] = (sqrt(1 + j + 0.5) * ret[j] - (2.0 * j + 3.0) * c * ret[j + 1] - sqrt(1 + j + 1.5) * ret[j + 2]);
  for (int64_t i = 2; i <= N; ++i) {
    const double sqi = sqrt((double)i);
    for (int64_t j = 0; j <= Lfull - 2 * i; ++j)
      ret[i * S + j] = (sqrt(i + j + 0.5) * ret[(i - 1) * S + j] - (2.0 * j + 3.0) * c * ret[(i - 1) * S + j + 1] -
                        sqrt(i + j + 1.5) * ret[(i - 1) * S + j + 2] + sqrt(i - 1.0) * ret[(i - 2) * S + j + 2]) / sqi;
  }
}

/* The same integral in closed form (Gradshteyn & Ryzhik 7.421.4 continued to I_nu), long double:
 *   d_nl(r) = 4 pi N_nl sqrt(pi/2) 2^{-l-3/2} beta^{-l-3/2} y^l e^{-r^2/(2(sigma^2+r0^2))} Q_n,
 *   y = r/sigma^2, beta = (1/r0^2 + 1/sigma^2)/2, alpha = 1/r0^2, delta = (beta-alpha)/beta,
 *   Q_n = delta^n L_n^{l+1/2}(alpha y^2 / (4 beta (beta - alpha)))   by the Laguerre recurrence
 *   (n+1) Q_{n+1} = ((2n+l+3/2) delta - kappa) Q_n - (n+l+1/2) delta^2 Q_{n-1},
 *   kappa = alpha y^2/(4 beta^2),
 *   N_nl = sqrt(2 n! r0^{-2l-3} / Gamma(n+l+3/2))   (utils.py:408-412 norm_harmonicBasis).
 * This is what numpy's radialIntegralHarmonic + coeffs_harmonicBasis sum
 * (sphericalAlignment.py:288-299,345-360; utils.py:414-427) and the Fortran recurrence both
 * evaluate, without their cancellation. ret[n*(L+1)+l]. */
void oracle_sph_harm_radial_exact(int64_t N, int64_t L, double rj, double sigma, double r0, double* ret) {
  const long double s2 = (long double)sigma * sigma, r02 = (long double)r0 * r0;
  const long double beta = 0.5L * (1.0L / r02 + 1.0L / s2), alpha = 1.0L / r02;
  const long double delta = (beta - alpha) / beta;
  const long double y = (long double)rj / s2;
  const long double kappa = alpha * y * y / (4.0L * beta * beta);
  const long double ex = expl(-0.5L * (long double)rj * rj / (s2 + r02));
  const long double pi = acosl(-1.0L);
  for (int64_t l = 0; l <= L; ++l) {
    const long double nu = l + 0.5L;
    long double norm = sqrtl(2.0L * powl((long double)r0, -2.0L * l - 3.0L) / tgammal(l + 1.5L));
    const long double pref = 4.0L * pi * sqrtl(pi / 2.0L) * powl(2.0L, -nu - 1.0L) * powl(beta, -nu - 1.0L) *
                             powl(y, (long double)l) * ex;
    long double qm1 = 0.0L, q = 1.0L;
    for (int64_t n = 0; n <= N; ++n) {
      ret[n * (L + 1) + l] = (double)(pref * norm * q);
      const long double qn = (((2.0L * n + 1.0L + nu) * delta - kappa) * q - (n + nu) * delta * delta * qm1) / (n + 1.0L);
      qm1 = q;
      q = qn;
      norm *= sqrtl((n + 1.0L) / (n + 1.0L + nu)); /* N_{n+1,l}/N_{n,l} */
    }
  }
}

/* C[n][l][m] = sum_j d_nl(r_j) conj(Y_lm(r_j)) for one (gathered) set of atoms:
 * HARMONICCOEFFS fastoverlap/f90/fastclusters.f90:654-687; numpy calcHarmCoeff
 * sphericalAlignment.py:345-361.  radial: 0 = closed form, 1 = Fortran recurrence.
 * Output C[(n*(L+1)+l)*(2L+1) + (m mod 2L+1)] (numpy layout of cnlm). */
void oracle_sph_harm_coeffs(const double* pos, int64_t natoms, const int32_t* idx, int64_t nidx,
                            int64_t N, int64_t L, double r0, double sigma, int radial, cplx* C) {
  (void)natoms;
  const int64_t NJ = 2 * L + 1;
  cplx* Y = (cplx*)malloc(sizeof(cplx) * (size_t)((L + 1) * NJ));
  const int64_t Lf = L + 2 * N;
  double* d = (double*)malloc(sizeof(double) * (size_t)((N + 1) * (Lf + 1)));
  memset(C, 0, sizeof(cplx) * (size_t)((N + 1) * (L + 1) * NJ));
  for (int64_t a = 0; a < nidx; ++a) {
    const double* p = pos + 3 * (int64_t)idx[a];
    const double r = oracle_sph_ylm(p, L, Y);
    int64_t S;
    if (radial == 1) {
      oracle_sph_harm_radial_fortran(N, Lf, r, sigma, r0, d);
      S = Lf + 1;
    } else {
      oracle_sph_harm_radial_exact(N, L, r, sigma, r0, d);
      S = L + 1;
    }
    for (int64_t n = 0; n <= N; ++n)
      for (int64_t l = 0; l <= L; ++l)
        for (int64_t m = -l; m <= l; ++m)
          C[(n * (L + 1) + l) * NJ + wrapm(m, NJ)] += d[n * S + l] * conj(Y[l * NJ + wrapm(m, NJ)]);
  }
  free(Y); free(d);
}

/* I[l][m1][m2] = sum_groups sum_n conj(C1[g][n][l][m1]) C2[g][n][l][m2]; invert: C2 <- (-1)^l C2.
 * DOTHARMONICCOEFFSPERM fastoverlap/f90/fastclusters.f90:765-788, inversion :351-353;
 * numpy calcSO3Harm sphericalAlignment.py:368-372. */
void oracle_sph_dot_harm(const cplx* C1, const cplx* C2, int64_t ngroups, int64_t N, int64_t L, int invert,
                         cplx* Ilmm) {
  const int64_t NJ = 2 * L + 1, per = (N + 1) * (L + 1) * NJ;
  memset(Ilmm, 0, sizeof(cplx) * (size_t)((L + 1) * NJ * NJ));
  for (int64_t g = 0; g < ngroups; ++g)
    for (int64_t l = 0; l <= L; ++l) {
      const double sgn = (invert && (l & 1)) ? -1.0 : 1.0;
      for (int64_t m1 = -l; m1 <= l; ++m1)
        for (int64_t m2 = -l; m2 <= l; ++m2) {
          cplx acc = 0.0;
          for (int64_t n = 0; n <= N; ++n)
            acc += conj(C1[g * per + (n * (L + 1) + l) * NJ + wrapm(m1, NJ)]) *
                   C2[g * per + (n * (L + 1) + l) * NJ + wrapm(m2, NJ)];
          Ilmm[(l * NJ + wrapm(m1, NJ)) * NJ + wrapm(m2, NJ)] += sgn * acc;
        }
    }
}

/* Whole spherical hot path (direct coefficients) for one pair of centred structures: coefficients,
 * inverse SOFT and arg-max for the normal and (optionally) the inverted orientation, where the
 * inverted coefficients are those of (A, -B): Y_lm(-r) = (-1)^l Y_lm(r) => I_inv^l = (-1)^l I^l
 * (fastclusters.f90:351-353; numpy recomputes them, sphericalAlignment.py:180-183).
 * BaseSphericalAlignment.align up to findMax (sphericalAlignment.py:160-194); ALIGN up to
 * FINDROTATIONS' arg-max (fastclusters.f90:129-269, :962-988). */
void oracle_sph_align_pair(const double* posA, const double* posB, int64_t natoms, const int32_t* goff,
                           int64_t ngroups, const int32_t* gidx, int64_t L, double sigma, int invert,
                           const double* Ds, int64_t* best_idx, double* best_val, double* frac_idx,
                           double* grid_out /*nullable [O][2B]^3*/) {
  const int64_t B = L + 1, NJ = 2 * L + 1, NK = 2 * B, G3 = NK * NK * NK;
  cplx* Ic = (cplx*)malloc(sizeof(cplx) * (size_t)((L + 1) * NJ * NJ));
  double* grid = (double*)malloc(sizeof(double) * (size_t)G3);
  oracle_sph_coeffs_direct(posA, posB, natoms, goff, ngroups, gidx, L, sigma, Ic);
  const int64_t shape[3] = {NK, NK, NK};
  for (int o = 0; o < (invert ? 2 : 1); ++o) {
    if (o == 1)
      for (int64_t l = 1; l <= L; l += 2)
        for (int64_t e = 0; e < NJ * NJ; ++e) Ic[l * NJ * NJ + e] = -Ic[l * NJ * NJ + e];
    oracle_isoft(Ic, B, Ds, grid, NULL);
    oracle_find_max(grid, shape, best_idx + 3 * o, frac_idx + 3 * o);
    best_val[o] = grid[(best_idx[3 * o] * NK + best_idx[3 * o + 1]) * NK + best_idx[3 * o + 2]];
    if (grid_out) memcpy(grid_out + o * G3, grid, sizeof(double) * (size_t)G3);
  }
  free(Ic); free(grid);
}

int oracle_sph_align_pairs(const double* posA, const double* posB, int64_t npairs, int64_t natoms,
                           const int32_t* goff, int64_t ngroups, const int32_t* gidx, int64_t L,
                           double sigma, int invert, int64_t* best_idx, double* best_val,
                           double* frac_idx, int nthreads) {
  const int64_t B = L + 1, NM = 2 * B - 1, NK = 2 * B;
  const int O = invert ? 2 : 1;
  double* Ds = (double*)malloc(sizeof(double) * (size_t)(B * NM * NM * NK));
  oracle_wigner_table(B, Ds); /* cached once, like SETBANDWIDTH DSOFT.f90:38-59 */
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < npairs; ++p)
    oracle_sph_align_pair(posA + p * natoms * 3, posB + p * natoms * 3, natoms, goff, ngroups, gidx, L,
                          sigma, invert, Ds, best_idx + 3 * O * p, best_val + O * p, frac_idx + 3 * O * p,
                          NULL);
  free(Ds);
  return used;
}
