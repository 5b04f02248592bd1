// Spherical (cluster) hot path.  Replaces reference fastoverlap/f90/fastclusters.f90
// (RYML, FOURIERCOEFFS, HARMONICNL/HARMONICCOEFFS, DOTHARMONICCOEFFS, CALCOVERLAP), DSOFT.f90
// (CALCWIGNERD, ISOFT), fastutils.f90 SPHI + the arg-max of FINDPEAKS, and the numpy kernels
// sphericalAlignment.py:57-65,260-273,288-372, soft.py:73-125, utils.py:319-338.
//
// Kernels
//   sph_prep_kernel     r_j and Y_lm(r_j) (m >= 0) per atom                         [K1]
//   sph_bessel_kernel   B_l[j,k] = 4 pi^2.5 s^3 i_l(r_j r_k/2s^2) e^{-(r_j^2+r_k^2)/4s^2}  [K2a]
//   sph_direct_kernel   I^l = Y_A^l B_l Y_B^l^H per (pair, l)                        [K2a]
//   sph_harm_kernel     C_nlm = sum_j d_nl(r_j) conj(Y_lm(r_j)) per (structure, group) [K2b]
//   sph_dot_kernel      I^l_{mm'} = sum_g sum_n conj(C^A) C^B per pair               [K3]
//   sph_wigner_kernel   table sqrt((2l+1)/2) d^l_{m1m2}(beta_k)  (cached per Jmax)   [K4]
//   sph_isoft_kernel    Wigner contraction + 2-D DFT + arg-max per (pair, beta chunk) [K5-K7]
//   sph_final_kernel    reduce over beta chunks + findMax parabola                   [K7]
//
// Device layouts (complex = double2)
//   Ypk   [struct][atom][lm]            lm = l(l+1)/2 + m, 0 <= m <= l  (Y_{l,-m} = (-1)^m conj Y_lm)
//   Bes   [pair][l][j][k]
//   Ihalf [pair][m2 = 0..L][m1 + L = 0..2L][l = 0..L]     only m2 >= 0 is stored because
//         I[l,-m1,-m2] = (-1)^{m1+m2} conj(I[l,m1,m2]) for real densities => the grid is real
//   Dt    [m2 = 0..L][m1 + L][l][k = 0..2B-1]             Wigner table, same (m2,m1,l) order
//   Cpk   [struct][group][n][lm]        harmonic coefficients, m >= 0
// All arithmetic is FP64.
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "fo_internal.h"
#include "fo_async.cuh"
#include "fo_symdft.cuh"

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kPi25 = 17.49341832762486284626282167987155377876;  // pi^2.5 (pow() in a kernel is ~150 instructions per thread)

__host__ __device__ __forceinline__ int nlm_of(int L) { return (L + 1) * (L + 2) / 2; }

// ------------------------------------------------------------------------------------------
// K1: r and Y_lm (m >= 0), scipy / Condon-Shortley convention.  One thread per (atom, m): the
// stable 3-term recurrence in l of the fully normalised associated Legendre functions
// (what XDNRMP legendre.f90:143-371 returns to RYML fastclusters.f90:602-652).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
sph_prep_kernel(const double* __restrict__ pos, int natoms, int L, size_t nstruct,
                double2* __restrict__ Ypk, double* __restrict__ R, int* status) {
  // Recurrence coefficients once per CTA (they cost two square roots and two divisions per (l, m),
  // which used to be paid by every thread at every step):
  //   P_m^m = cmm[m] sin^m(theta),  P_l^m = ca[lm] (cos(theta) P_{l-1}^m - cb[lm] P_{l-2}^m)
  extern __shared__ double sm_prep[];
  const int NLM = nlm_of(L);
  double* ca = sm_prep;        // [NLM]
  double* cb = ca + NLM;       // [NLM]
  double* cmm = cb + NLM;      // [L + 1]
  for (int e = threadIdx.x; e < NLM; e += blockDim.x) {
    int l = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
    while ((l + 1) * (l + 2) / 2 <= e) ++l;
    while (l * (l + 1) / 2 > e) --l;
    const int m = e - l * (l + 1) / 2;
    double a = 0.0, b = 0.0;
    if (l >= m + 2) {
      a = sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
      b = sqrt((((double)(l - 1) * (l - 1)) - (double)m * m) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
    }
    ca[e] = a;
    cb[e] = b;
  }
  if (threadIdx.x == 0) {
    double c = 0.28209479177387814347403972578039;  // sqrt(1/(4 pi))
    cmm[0] = c;
    for (int q = 1; q <= L; ++q) {
      c *= -sqrt((2.0 * q + 1.0) / (2.0 * q));
      cmm[q] = c;
    }
  }
  __syncthreads();
  const size_t total = nstruct * (size_t)natoms * (L + 1);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(t % (L + 1));
    const size_t sa = t / (L + 1);  // struct*natoms + atom
    const double x = pos[sa * 3 + 0], y = pos[sa * 3 + 1], z = pos[sa * 3 + 2];
    const double r = sqrt(x * x + y * y + z * z);
    const bool fin = isfinite(r);
    double ct = 1.0, st = 0.0, cph = 1.0, sph = 0.0;
    if (r > 0.0 && fin) {
      ct = z / r;
      const double rho = sqrt(x * x + y * y);
      st = rho / r;
      if (rho > 0.0) {
        cph = x / rho;
        sph = y / rho;
      }
    }
    if (m == 0) {
      R[sa] = r;
      if (status) {
        const size_t s = sa / natoms;
        if (!fin) atomicOr(&status[s], FO_STATUS_NONFINITE);
        if (r == 0.0) atomicOr(&status[s], FO_STATUS_ATOM_AT_ORIGIN);
      }
    }
    // exp(i m phi) = (cos phi + i sin phi)^m and sin^m(theta) by m multiplications
    double cm = 1.0, sm = 0.0, stm = 1.0;
    for (int q = 0; q < m; ++q) {
      const double c2 = cm * cph - sm * sph;
      sm = sm * cph + cm * sph;
      cm = c2;
      stm *= st;
    }
    const double pmm = cmm[m] * stm;
    double2* out = Ypk + sa * NLM;
    double pl2 = 0.0, pl1 = pmm;
    out[m * (m + 1) / 2 + m] = make_double2(pmm * cm, pmm * sm);
    if (m + 1 <= L) {
      const double p = sqrt(2.0 * m + 3.0) * ct * pmm;
      out[(m + 1) * (m + 2) / 2 + m] = make_double2(p * cm, p * sm);
      pl2 = pmm;
      pl1 = p;
    }
    for (int l = m + 2; l <= L; ++l) {
      const int lm = l * (l + 1) / 2 + m;
      const double p = ca[lm] * (ct * pl1 - cb[lm] * pl2);
      out[lm] = make_double2(p * cm, p * sm);
      pl2 = pl1;
      pl1 = p;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K2a part 1: Bessel/Gaussian matrix.  One thread per (pair, j, k): exp(-x) i_l(x) for all l by
// Miller's backward recurrence (x >= 1; the algorithm of SPHI fastutils.f90:1000-1096) or the
// ascending series (x < 1), times the Gaussian exp(-(r_j - r_k)^2 / 4 s^2) -- together exactly
// i_l(x) exp(-(r_j^2 + r_k^2)/4 s^2) of sphericalAlignment.py:268-272 without overflow.
// Pairs of atoms in different permutation groups get 0 (the reference sums calcSO3Coeffs over
// groups, sphericalAlignment.py:175).
// ------------------------------------------------------------------------------------------
// exp(-x) i_l(x) g for l = 0..L into out[l * ls] (see sph_bessel_kernel)
__device__ __forceinline__ void bessel_levels(double ra, double rb, int L, double fact, double inv2s2,
                                              double* __restrict__ out, size_t ls) {
  const double x = ra * rb * inv2s2;
  const double dr = ra - rb;
  const double g = fact * exp(-0.5 * dr * dr * inv2s2);
  if (!(x >= 1e-100)) {  // also catches NaN
    out[0] = (x == x) ? g : x;
    for (int l = 1; l <= L; ++l) out[(size_t)l * ls] = (x == x) ? 0.0 : x;
    return;
  }
  if (x < 1.0) {
    // i_l(x) = x^l/(2l+1)!! sum_q (x^2/2)^q / (q! (2l+3)(2l+5)...(2l+2q+1))
    const double emx = exp(-x), h = 0.5 * x * x;
    double pref = emx;  // e^{-x} x^l/(2l+1)!!
    for (int l = 0; l <= L; ++l) {
      double term = 1.0, sum = 1.0;
      for (int q = 1; q <= 14; ++q) {
        term *= h / ((double)q * (2.0 * l + 2.0 * q + 1.0));
        sum += term;
      }
      out[(size_t)l * ls] = g * pref * sum;
      pref *= x / (2.0 * l + 3.0);
    }
    return;
  }
  const double si0 = -expm1(-2.0 * x) / (2.0 * x);
  const int mstart = 16 + (int)sqrt(50.0 * x + (double)L * L);
  const double invx = 1.0 / x;
  // Miller: f_{q} = (2 q + 3) f_{q+1} / x + f_{q+2} downwards from an arbitrary small start; the values are the
  // recurrence's times i_0(x) / f_0.  First pass: down to q = 0 keeping only the state at q = L + 1; second pass:
  // the L + 1 wanted levels again, already scaled (16 more FMAs instead of a read-modify-write of the output).
  // The rescaling test runs every fourth step: a step grows f by less than (2 q + 3) / x + 1 < 300, so four
  // steps after 1e150 stay far below the overflow threshold.
  double f0 = 0.0, f1 = 1e-100, f = 0.0;
  int q = mstart;
  for (; q > L + 3; q -= 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      f = (2.0 * (q - u) + 3.0) * f1 * invx + f0;
      f0 = f1;
      f1 = f;
    }
    if (f > 1e150) {
      f0 *= 1e-150;
      f1 *= 1e-150;
    }
  }
  for (; q > L; --q) {
    f = (2.0 * q + 3.0) * f1 * invx + f0;
    f0 = f1;
    f1 = f;
  }
  if (f1 > 1e150) {
    f0 *= 1e-150;
    f1 *= 1e-150;
  }
  const double s0 = f0, s1 = f1;  // f_{L+2}, f_{L+1}
  for (q = L; q >= 0; --q) {
    f = (2.0 * q + 3.0) * f1 * invx + f0;
    f0 = f1;
    f1 = f;
  }
  const double cs = g * si0 / f;
  f0 = s0 * cs;
  f1 = s1 * cs;
  for (q = L; q >= 0; --q) {
    f = (2.0 * q + 3.0) * f1 * invx + f0;
    out[(size_t)q * ls] = f;
    f0 = f1;
    f1 = f;
  }
}

__global__ void sph_bessel_kernel(const double* __restrict__ RA, const double* __restrict__ RB,
                                  const int* __restrict__ gid, int natoms, int L, double sigma,
                                  size_t npairs, double* __restrict__ Bes) {
  const size_t NN = (size_t)natoms * natoms;
  const size_t total = npairs * NN;
  const double fact = 4.0 * kPi25 * sigma * sigma * sigma;
  const double inv2s2 = 0.5 / (sigma * sigma);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t p = t / NN;
    const int jk = (int)(t - p * NN);
    const int j = jk / natoms, k = jk - j * natoms;
    double* out = Bes + p * (size_t)(L + 1) * NN + jk;
    if (gid[j] != gid[k]) {
      for (int l = 0; l <= L; ++l) out[(size_t)l * NN] = 0.0;
      continue;
    }
    bessel_levels(RA[p * natoms + j], RB[p * natoms + k], L, fact, inv2s2, out, NN);
  }
}

// ------------------------------------------------------------------------------------------
// K2a part 2: per (pair, l):  T[m,k] = sum_j Y^A_{lm}(j) B_l[j,k]  (m >= 0; T[-m] = (-1)^m conj T[m])
//                             I[l,m1,m2] = sum_k T[m1,k] conj(Y^B_{l m2}(k)),  m2 >= 0, all m1.
// FOURIERCOEFFS fastclusters.f90:868-916 factorised into two small GEMMs.
// ------------------------------------------------------------------------------------------
constexpr int DIR_TK = 32;

__global__ void __launch_bounds__(128)
sph_direct_kernel(const double2* __restrict__ YA, const double2* __restrict__ YB,
                  const double* __restrict__ Bes, int natoms, int L, double2* __restrict__ Ihalf) {
  extern __shared__ double2 smd[];
  const int l = blockIdx.x;
  const size_t p = blockIdx.y;
  const int NLM = nlm_of(L), W = 2 * L + 1, L1 = L + 1;
  const int nm = l + 1;
  double2* T = smd;                    // [nm][DIR_TK]
  double2* YBs = T + nm * DIR_TK;      // [DIR_TK][nm]
  double2* acc = YBs + DIR_TK * nm;    // [(2l+1)][nm]  (m1 = -l..l, m2 = 0..l)
  const int nout = (2 * l + 1) * nm;
  const int lbase = l * (l + 1) / 2;
  const size_t NN = (size_t)natoms * natoms;
  const double* Bl = Bes + (p * L1 + l) * NN;
  const double2* ya = YA + p * (size_t)natoms * NLM;
  const double2* yb = YB + p * (size_t)natoms * NLM;
  for (int o = threadIdx.x; o < nout; o += blockDim.x) acc[o] = make_double2(0.0, 0.0);
  for (int k0 = 0; k0 < natoms; k0 += DIR_TK) {
    const int tk = min(DIR_TK, natoms - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < nm * tk; e += blockDim.x) {
      const int kk = e % tk, m = e / tk;
      double tr = 0.0, ti = 0.0;
      for (int j = 0; j < natoms; ++j) {
        const double b = Bl[(size_t)j * natoms + k0 + kk];
        const double2 y = ya[(size_t)j * NLM + lbase + m];
        tr = fma(y.x, b, tr);
        ti = fma(y.y, b, ti);
      }
      T[m * DIR_TK + kk] = make_double2(tr, ti);
    }
    for (int e = threadIdx.x; e < nm * tk; e += blockDim.x) {
      const int m = e % nm, kk = e / nm;
      YBs[kk * nm + m] = yb[(size_t)(k0 + kk) * NLM + lbase + m];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < nout; o += blockDim.x) {
      const int m2 = o % nm, m1 = o / nm - l;
      const int am = m1 < 0 ? -m1 : m1;
      const double sgn = (m1 < 0 && (am & 1)) ? -1.0 : 1.0;
      double ar = 0.0, ai = 0.0;
      for (int kk = 0; kk < tk; ++kk) {
        double2 t = T[am * DIR_TK + kk];
        if (m1 < 0) t.y = -t.y;
        const double2 y = YBs[kk * nm + m2];
        ar += t.x * y.x + t.y * y.y;
        ai += t.y * y.x - t.x * y.y;
      }
      acc[o].x += sgn * ar;
      acc[o].y += sgn * ai;
    }
  }
  __syncthreads();
  double2* out = Ihalf + p * (size_t)L1 * W * L1;
  for (int o = threadIdx.x; o < nout; o += blockDim.x) {
    const int m2 = o % nm, m1 = o / nm - l;
    out[((size_t)m2 * W + (m1 + L)) * L1 + l] = acc[o];
  }
}

// ------------------------------------------------------------------------------------------
// K2a for large clusters (natoms >= ctx->direct_gemm_min): the two contractions of
// sph_direct_kernel as FP64 tensor-core GEMMs (DMMA.8x8x4), batched over (pair, l).  Complex
// quantities are handled as interleaved real rows / columns r = 2 m + (re | im), m = 0..l:
//   MODE 1   T[r][k]   = sum_j  YA[j][r] B_l[j][k]              (2(l+1) x N x N)
//   MODE 2   C2[r][r'] = sum_k  T[r][k]  YB[k][r']              (2(l+1) x 2(l+1) x N)
// and sph_direct_combine_kernel forms I[l,+-m1,m2] from the four real products in C2.
// CTA tile 64 x 64 x 32, 8 warps as 2 (rows) x 4 (cols), warp tile 32 x 16; k-major shared tiles with
// pitch 72 (== 8 mod 32: the four k rows of a fragment fall into disjoint bank groups); a three-stage
// cp.async pipeline keeps two k tiles in flight while the current one is multiplied.
// ------------------------------------------------------------------------------------------
constexpr int DG_TM = 64, DG_TN = 64, DG_TK = 32, DG_LD = 72, DG_THREADS = 256, DG_STAGES = 3;
constexpr size_t DG_SMEM = (size_t)DG_STAGES * 2 * DG_TK * DG_LD * 8;

__host__ __device__ inline size_t dg_c2_off(int l) { return (size_t)2 * l * (l + 1) * (2 * l + 1) / 3; }

template <int MODE>
__global__ void __launch_bounds__(DG_THREADS, 2)
sph_direct_gemm_kernel(const double* __restrict__ Yreal, const double* __restrict__ X, int natoms, int L,
                       double* __restrict__ Cout) {
  extern __shared__ double smg[];  // [DG_STAGES][A | B][DG_TK][DG_LD]
  const int L1 = L + 1, NLM = nlm_of(L);
  const int l = blockIdx.z % L1;
  const size_t p = blockIdx.z / L1;
  const int M = 2 * (l + 1);
  const int N = MODE == 1 ? natoms : M;
  const int K = natoms;
  const int row0 = blockIdx.y * DG_TM, col0 = blockIdx.x * DG_TN;
  if (row0 >= M || col0 >= N) return;
  const int lbase = l * (l + 1) / 2;
  const double* Ay;  // element (r, kk)
  const double* Bx;  // element (kk, n)
  double* C;
  size_t rsA, csA, rsB, ldc;
  if (MODE == 1) {
    Ay = Yreal + (p * (size_t)natoms * NLM + lbase) * 2;
    rsA = 1;
    csA = (size_t)2 * NLM;
    Bx = X + (p * L1 + l) * (size_t)natoms * natoms;
    rsB = natoms;
    C = Cout + (p * (size_t)2 * NLM + 2 * lbase) * natoms;
    ldc = natoms;
  } else {
    Ay = X + (p * (size_t)2 * NLM + 2 * lbase) * natoms;
    rsA = natoms;
    csA = 1;
    Bx = Yreal + (p * (size_t)natoms * NLM + lbase) * 2;
    rsB = (size_t)2 * NLM;
    C = Cout + p * dg_c2_off(L1) + dg_c2_off(l);
    ldc = M;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  // Operand tiles go global -> shared with cp.async (zero fill outside the matrices), three stages of
  // 32 k each: one barrier per 512 DMMA of the CTA, no register staging.
  auto As = [&](int st) { return smg + (size_t)st * 2 * DG_TK * DG_LD; };
  auto Bs = [&](int st) { return smg + (size_t)st * 2 * DG_TK * DG_LD + DG_TK * DG_LD; };
  auto cp16 = [](double* dst, const double* src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
  };
  auto cp8 = [](double* dst, const double* src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
  };
  auto issue = [&](int kt) {
    const int k0 = kt * DG_TK;
    double* as = As(kt % DG_STAGES);
    double* bs = Bs(kt % DG_STAGES);
    if (MODE == 1) {
      // A(r, kk): rows contiguous and 16-byte aligned (interleaved re / im); 64 x 32 doubles = 1024 chunks
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = tid + i * DG_THREADS;      // chunk: kk = c / 32, rows 2 (c % 32) ..
        const int kk = c >> 5, r = (c & 31) * 2;
        const bool ok = row0 + r < M && k0 + kk < K;  // M is even: a chunk is all in or all out
        cp16(as + kk * DG_LD + r, ok ? Ay + (size_t)(k0 + kk) * csA + row0 + r : Ay, ok);
      }
      // B(kk, n): 8-byte aligned only (natoms may be odd); 32 x 64 doubles
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = tid + i * DG_THREADS;
        const int kk = c >> 6, n = c & 63;
        const bool ok = col0 + n < N && k0 + kk < K;
        cp8(bs + kk * DG_LD + n, ok ? Bx + (size_t)(k0 + kk) * rsB + col0 + n : Bx, ok);
      }
    } else {
      // A(r, kk): contiguous along kk in memory, k-major in shared memory: element-wise
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = tid + i * DG_THREADS;
        const int r = c >> 5, kk = c & 31;
        const bool ok = row0 + r < M && k0 + kk < K;
        cp8(as + kk * DG_LD + r, ok ? Ay + (size_t)(row0 + r) * rsA + k0 + kk : Ay, ok);
      }
      // B(kk, n): interleaved re / im columns, 16-byte aligned
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = tid + i * DG_THREADS;
        const int kk = c >> 5, n = (c & 31) * 2;
        const bool ok = col0 + n < N && k0 + kk < K;
        cp16(bs + kk * DG_LD + n, ok ? Bx + (size_t)(k0 + kk) * rsB + col0 + n : Bx, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double acc[4][2][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  bool live[4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) live[mt] = row0 + wm * 32 + mt * 8 < M;
  const int nkt = (K + DG_TK - 1) / DG_TK;
  issue(0);
  if (nkt > 1) issue(1);
  for (int kt = 0; kt < nkt; ++kt) {
    if (kt + 1 < nkt)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    else
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tile kt visible to all; everyone is done with tile kt-1 (its stage is reused next)
    if (kt + 2 < nkt) issue(kt + 2);
    const double* as = As(kt % DG_STAGES) + wm * 32 + g;
    const double* bs = Bs(kt % DG_STAGES) + wn * 16 + g;
#pragma unroll
    for (int k4 = 0; k4 < DG_TK / 4; ++k4) {
      const int krow = (k4 * 4 + t4) * DG_LD;
      const double b0 = bs[krow], b1 = bs[krow + 8];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (live[mt]) {
          const double a = as[krow + mt * 8];
          fo_dmma(acc[mt][0], a, b0);
          fo_dmma(acc[mt][1], a, b1);
        }
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int r = row0 + wm * 32 + mt * 8 + g;
    if (r >= M) continue;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int n = col0 + wn * 16 + nt * 8 + t4 * 2;
      if (n < N) C[(size_t)r * ldc + n] = acc[mt][nt][0];
      if (n + 1 < N) C[(size_t)r * ldc + n + 1] = acc[mt][nt][1];
    }
  }
}

// I[l,m1,m2] (m2 >= 0, all m1) in the Ihalf layout from C2[l][2 m1 + part][2 m2 + part'].
__global__ void sph_direct_combine_kernel(const double* __restrict__ C2, int L, size_t npairs,
                                          double2* __restrict__ Ihalf) {
  const int W = 2 * L + 1, L1 = L + 1;
  const size_t per = (size_t)L1 * W * L1;
  const size_t c2per = dg_c2_off(L1);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < npairs * per;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t p = t / per;
    int r = (int)(t - p * per);
    const int l = r % L1;
    r /= L1;
    const int m1 = r % W - L, m2 = r / W;
    const int am = m1 < 0 ? -m1 : m1;
    if (l < am || l < m2) continue;
    const int n2 = 2 * (l + 1);
    const double* c = C2 + p * c2per + dg_c2_off(l);
    const double rr = c[(size_t)(2 * am) * n2 + 2 * m2], ri = c[(size_t)(2 * am) * n2 + 2 * m2 + 1];
    const double ir = c[(size_t)(2 * am + 1) * n2 + 2 * m2], ii = c[(size_t)(2 * am + 1) * n2 + 2 * m2 + 1];
    double2 v;
    if (m1 >= 0) {
      v = make_double2(rr + ii, ir - ri);
    } else {
      const double sg = (am & 1) ? -1.0 : 1.0;
      v = make_double2(sg * (rr - ii), sg * (-ir - ri));
    }
    Ihalf[t] = v;
  }
}

// ------------------------------------------------------------------------------------------
// K2a for small clusters on the tensor pipe (sph_direct_mma_kernel): the two contractions of
// sph_direct_kernel for one (pair, l) per CTA as DMMA.8x8x4 tiles fed from shared memory, complex
// quantities as interleaved real rows r = 2 m + (re | im) like sph_direct_gemm_kernel:
//   T[r][k]    = sum_j YA[j][r] B_l[j][k]       (2(l+1) x N x N)
//   C2[r][r']  = sum_k T[r][k]  YB[k][r']       (2(l+1) x 2(l+1) x N)
// and I[l,+-m1,m2] is formed in the C fragments (the re / im rows of one m1 sit in lanes g, g^1).
// All operands are k-major with a pitch == 8 (mod 32) doubles: conflict-free fragment loads.
// ncu on the scalar kernel: 16 % FP64 pipe, bound by shared / global operand loads (2 loads per 2 FMA).
// ------------------------------------------------------------------------------------------
constexpr int DS_THREADS = 128;
__host__ __device__ inline int fo_ld8(int w) { return w + ((8 - w) % 32 + 32) % 32; }

__global__ void __launch_bounds__(DS_THREADS)
sph_direct_mma_kernel(const double2* __restrict__ YA, const double2* __restrict__ YB,
                      const double* __restrict__ Bes, int natoms, int L, double2* __restrict__ Ihalf) {
  extern __shared__ double smq[];
  const int l = blockIdx.x;
  const size_t p = blockIdx.y;
  const int NLM = nlm_of(L), W = 2 * L + 1, L1 = L + 1;
  const int R = 2 * (l + 1), R8 = (R + 7) & ~7;
  const int N4 = (natoms + 3) & ~3, N8 = (natoms + 7) & ~7;
  const int LDR = fo_ld8((2 * L1 + 7) & ~7), LDN = fo_ld8(N8);
  double* A1 = smq;                       // [N8][LDR]  YA_l:  A1[j LDR + r]
  double* B2 = A1 + (size_t)N8 * LDR;     // [N8][LDR]  YB_l
  double* Ts = B2 + (size_t)N8 * LDR;     // [N8][LDR]  T:     Ts[k LDR + r]
  double* B1 = Ts + (size_t)N8 * LDR;     // [N8][LDN]  B_l:   B1[j LDN + k]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int lbase = l * (l + 1) / 2;
  const double* ya = reinterpret_cast<const double*>(YA + p * (size_t)natoms * NLM + lbase);
  const double* yb = reinterpret_cast<const double*>(YB + p * (size_t)natoms * NLM + lbase);
  const double* Bl = Bes + (p * L1 + l) * (size_t)natoms * natoms;
  // Operands go global -> shared with cp.async (all copies of the CTA in flight at once: ncu showed
  // the register-staged loads of the first version as 46 % long-scoreboard stalls); the zero padding
  // (rows r >= R, atoms >= natoms) is written with ordinary stores to disjoint addresses.
  {
    // flat loops (the first version walked (row, 16-lane column group) nests: 36 % of the kernel's stall samples
    // sat in their index arithmetic and in the half-empty zero-fill loops, profiles/r02_summary.md)
    const int nm = l + 1;
    for (int e = tid; e < natoms * nm; e += DS_THREADS) {
      const int j = e / nm, m = e - j * nm;
      const unsigned da = (unsigned)__cvta_generic_to_shared(A1 + j * LDR + 2 * m);
      const unsigned db = (unsigned)__cvta_generic_to_shared(B2 + j * LDR + 2 * m);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(ya + (size_t)j * 2 * NLM + 2 * m)
                   : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(db), "l"(yb + (size_t)j * 2 * NLM + 2 * m)
                   : "memory");
    }
    if ((natoms & 1) == 0) {  // rows of B_l are 16-byte multiples: half as many copies
      const int nh = natoms >> 1;
      for (int e = tid; e < natoms * nh; e += DS_THREADS) {
        const int j = e / nh, k = 2 * (e - j * nh);
        const unsigned d1 = (unsigned)__cvta_generic_to_shared(B1 + j * LDN + k);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(Bl + (size_t)j * natoms + k) : "memory");
      }
    } else {
      for (int e = tid; e < natoms * natoms; e += DS_THREADS) {
        const int j = e / natoms, k = e - j * natoms;
        const unsigned d1 = (unsigned)__cvta_generic_to_shared(B1 + j * LDN + k);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d1), "l"(Bl + (size_t)j * natoms + k) : "memory");
      }
    }
    // zero padding: columns r >= R / k >= natoms of the atom rows, and the rows j >= natoms
    const int pr = R8 - R, pk = N8 - natoms;
    for (int e = tid; e < natoms * pr; e += DS_THREADS) {
      const int j = e / pr, r = R + e - j * pr;
      A1[j * LDR + r] = B2[j * LDR + r] = 0.0;
    }
    for (int e = tid; e < natoms * pk; e += DS_THREADS) {
      const int j = e / pk, k = natoms + e - j * pk;
      B1[j * LDN + k] = 0.0;
    }
    for (int e = tid; e < pk * R8; e += DS_THREADS) {
      const int j = natoms + e / R8, r = e % R8;
      A1[j * LDR + r] = B2[j * LDR + r] = 0.0;
    }
    for (int e = tid; e < pk * N8; e += DS_THREADS) {
      const int j = natoms + e / N8, k = e % N8;
      B1[j * LDN + k] = 0.0;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  const int nrt = R8 >> 3, nct = N8 >> 3, nks = N4 >> 2;
  for (int job = warp; job < nrt * nct; job += DS_THREADS / 32) {
    const int rt = job % nrt, ct = job / nrt;
    double acc[2] = {0.0, 0.0};
    const double* a = A1 + t4 * LDR + rt * 8 + g;
    const double* b = B1 + t4 * LDN + ct * 8 + g;
    for (int ks = 0; ks < nks; ++ks) fo_dmma(acc, a[ks * 4 * LDR], b[ks * 4 * LDN]);
    Ts[(ct * 8 + 2 * t4) * LDR + rt * 8 + g] = acc[0];
    Ts[(ct * 8 + 2 * t4 + 1) * LDR + rt * 8 + g] = acc[1];
  }
  __syncthreads();
  double2* out = Ihalf + p * (size_t)L1 * W * L1;
  for (int job = warp; job < nrt * nrt; job += DS_THREADS / 32) {
    const int rt1 = job % nrt, rt2 = job / nrt;
    double acc[2] = {0.0, 0.0};
    const double* a = Ts + t4 * LDR + rt1 * 8 + g;
    const double* b = B2 + t4 * LDR + rt2 * 8 + g;
    for (int ks = 0; ks < nks; ++ks) fo_dmma(acc, a[ks * 4 * LDR], b[ks * 4 * LDR]);
    // this lane: row r1 = 8 rt1 + g (m1 = r1 / 2, part g & 1), columns (re, im) of m2 = 4 rt2 + t4
    const double px0 = __shfl_xor_sync(0xffffffffu, acc[0], 4);
    const double px1 = __shfl_xor_sync(0xffffffffu, acc[1], 4);
    const int m1 = (rt1 * 8 + g) >> 1, m2 = rt2 * 4 + t4;
    if (m1 > l || m2 > l) continue;
    if ((g & 1) == 0) {  // rr, ri here; ir, ii in the partner: I(+m1, m2)
      out[((size_t)m2 * W + (L + m1)) * L1 + l] = make_double2(acc[0] + px1, px0 - acc[1]);
    } else if (m1 > 0) {  // ir, ii here; rr, ri in the partner: I(-m1, m2) = (-1)^m1 (rr - ii, -ir - ri)
      const double sg = (m1 & 1) ? -1.0 : 1.0;
      out[((size_t)m2 * W + (L - m1)) * L1 + l] = make_double2(sg * (px0 - acc[1]), sg * (-acc[0] - px1));
    }
  }
}

size_t direct_mma_smem(int64_t natoms, int L) {
  const int N8 = (int)((natoms + 7) & ~7);
  return ((size_t)3 * N8 * fo_ld8((2 * (L + 1) + 7) & ~7) + (size_t)N8 * fo_ld8(N8)) * 8;
}

// ------------------------------------------------------------------------------------------
// K2a for small clusters, streaming form (sph_direct2_kernel; the pairs API of the flagship configuration).
// sph_direct_mma_kernel spends 97 % of its instructions on moving operands (one CTA per (pair, l): 16-byte
// cp.async with an integer division each, zero fill, two CTA barriers, T through shared memory, two shared loads
// per DMMA).  Here the PRODUCERS write the operands as the shared-memory images the DMMA fragments want:
//   Ysw [structure][l][row 0..N8)[8 nrt(l) doubles]   Y_l: row = atom, column r = 2 m + (re | im), zero padded to
//                                                     whole 8 x 8 tiles (sph_prep2_kernel)
//   Bsw [pair][l][row j 0..N8)[N8 doubles]            B_l[j][k], zero padded (sph_bessel2_kernel)
// with the 8-column groups of a row xor-swizzled by the row (d2_pos) so that the four k rows of any fragment
// fall into disjoint bank groups at pitch 8 ntiles (no padding columns), and the rows of the second structure
// permuted within groups of eight (d2_rowperm) for the chained second product.  A persistent CTA then streams
// (pair, l) items through a ring of shared-memory slots: one elected thread fetches the three operands of an item
// with three bulk copies (cp.async.bulk, 2.5 - 12.8 KB each) completing on the slot's mbarrier, four consumer
// warps take the item's row tiles:
//   T[r][k]   = sum_j YA[j][r] B_l[j][k]     50 DMMA per row tile (N = 38), one A fragment per five B fragments;
//   C2[r][r'] = sum_k T[r][k] YB[k][r']      the C fragments of T ARE the A fragments of this product when a k-step is
//                                            taken as the columns {2 t + h} of a tile (t = lane & 3, h = 0 | 1): T never
//                                            leaves the registers, YB's rows are stored in that order;
// and I[l, +-m1, m2] is formed in the C fragments as in sph_direct_mma_kernel and written both in the Ihalf layout
// and (optionally) in the packed layout of the fast iSOFT kernels.  No CTA barrier, no generic-proxy store to shared
// memory, 1.1 shared loads per DMMA.  A slot is handed back on a named barrier (the consumer warps bar.arrive after
// their last fragment load, the producer warp bar.sync before it refills the slot): compute-sanitizer's racecheck
// follows that, it did not follow an "empty" mbarrier from generic-proxy reads to the async-proxy writes of the copy.
// ------------------------------------------------------------------------------------------
constexpr int D2_CONS = 4;                       // consumer warps
constexpr int D2_THREADS = (D2_CONS + 1) * 32;   // + the producer warp
constexpr int D2_MAXL = 64;

__host__ __device__ __forceinline__ int d2_sw(int ntiles, int p) {
  const int m = ntiles & 3;
  return m == 0 ? (p & 3) : (m == 2 ? ((p >> 1) & 1) : 0);
}
// element (row p, column c) of a block with ntiles 8-column groups per row
__host__ __device__ __forceinline__ int d2_pos(int ntiles, int p, int c) {
  return p * 8 * ntiles + (((c >> 3) ^ d2_sw(ntiles, p)) << 3) + (c & 7);
}
// row of atom k in the second structure's blocks: the k-step (tile ct, h) of the chained product has lane t at
// column 8 ct + 2 t + h of T, i.e. at row 8 ct + 4 h + t here
__host__ __device__ __forceinline__ int d2_rowperm(int k) { return (k & ~7) | ((k & 1) << 2) | ((k & 7) >> 1); }

struct D2Layout {
  int natoms, N8, NCT, L, CW;  // CW = 4 nrt(L): complex columns per row
  int ysz, bsz, ymax, slot;    // doubles: per structure, per (pair, l) of B, largest Y_l block, ring slot
  int nrt[D2_MAXL], yoff[D2_MAXL];
  D2Layout() {}
  D2Layout(int natoms_, int L_) {
    natoms = natoms_;
    L = L_;
    N8 = (natoms + 7) & ~7;
    NCT = N8 >> 3;
    int off = 0;
    for (int l = 0; l <= L && l < D2_MAXL; ++l) {
      nrt[l] = (2 * (l + 1) + 7) >> 3;
      yoff[l] = off;
      off += N8 * 8 * nrt[l];
    }
    ysz = off;
    bsz = N8 * N8;
    ymax = N8 * 8 * nrt[L < D2_MAXL ? L : D2_MAXL - 1];
    CW = 4 * nrt[L < D2_MAXL ? L : D2_MAXL - 1];
    slot = 2 * ymax + bsz;
  }
};

// K1 in the operand layout: one thread per (structure, row, complex column); rows >= natoms and the columns
// beyond m = l of a level are the zero padding.
template <bool BSIDE>
__global__ void __launch_bounds__(128)
sph_prep2_kernel(const double* __restrict__ pos, const __grid_constant__ D2Layout Y, size_t nstruct,
                 double* __restrict__ Ysw, double* __restrict__ R, int* status) {
  extern __shared__ double sm_prep[];
  const int L = Y.L, NLM = nlm_of(L), natoms = Y.natoms;
  double* ca = sm_prep;    // [NLM]
  double* cb = ca + NLM;   // [NLM]
  double* cmm = cb + NLM;  // [L + 1]
  for (int e = threadIdx.x; e < NLM; e += blockDim.x) {
    int l = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
    while ((l + 1) * (l + 2) / 2 <= e) ++l;
    while (l * (l + 1) / 2 > e) --l;
    const int m = e - l * (l + 1) / 2;
    double a = 0.0, b = 0.0;
    if (l >= m + 2) {
      a = sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
      b = sqrt((((double)(l - 1) * (l - 1)) - (double)m * m) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
    }
    ca[e] = a;
    cb[e] = b;
  }
  if (threadIdx.x == 0) {
    double c = 0.28209479177387814347403972578039;  // sqrt(1/(4 pi))
    cmm[0] = c;
    for (int q = 1; q <= L; ++q) {
      c *= -sqrt((2.0 * q + 1.0) / (2.0 * q));
      cmm[q] = c;
    }
  }
  __syncthreads();
  const size_t total = nstruct * (size_t)Y.N8 * Y.CW;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(t % Y.CW);
    const size_t sr = t / Y.CW;
    const int atom = (int)(sr % Y.N8);
    const size_t st_ = sr / Y.N8;
    const int rp = BSIDE ? d2_rowperm(atom) : atom;
    double* out = Ysw + st_ * (size_t)Y.ysz;
    const bool real = atom < natoms && m <= L;
    double ct = 1.0, cm = 1.0, sm = 0.0, pmm = 0.0;
    if (real) {
      const size_t sa = st_ * natoms + atom;
      const double x = pos[sa * 3 + 0], y = pos[sa * 3 + 1], z = pos[sa * 3 + 2];
      const double r = sqrt(x * x + y * y + z * z);
      const bool fin = isfinite(r);
      double st = 0.0, cph = 1.0, sph = 0.0;
      if (r > 0.0 && fin) {
        ct = z / r;
        const double rho = sqrt(x * x + y * y);
        st = rho / r;
        if (rho > 0.0) {
          cph = x / rho;
          sph = y / rho;
        }
      }
      if (m == 0) {
        R[sa] = r;
        if (status) {
          if (!fin) atomicOr(&status[st_], FO_STATUS_NONFINITE);
          if (r == 0.0) atomicOr(&status[st_], FO_STATUS_ATOM_AT_ORIGIN);
        }
      }
      // exp(i m phi) = (cos phi + i sin phi)^m and sin^m(theta) by m multiplications
      double stm = 1.0;
      for (int q = 0; q < m; ++q) {
        const double c2 = cm * cph - sm * sph;
        sm = sm * cph + cm * sph;
        cm = c2;
        stm *= st;
      }
      pmm = cmm[m] * stm;
    }
    double pl2 = 0.0, pl1 = 0.0;
    for (int l = 0; l <= L; ++l) {
      const int nt = Y.nrt[l];
      if (m >= 4 * nt) continue;  // no such column at this level
      double2 v = make_double2(0.0, 0.0);
      if (real && l >= m) {
        double pv;
        if (l == m)
          pv = pmm;
        else if (l == m + 1)
          pv = sqrt(2.0 * m + 3.0) * ct * pmm;
        else {
          const int lm = l * (l + 1) / 2 + m;
          pv = ca[lm] * (ct * pl1 - cb[lm] * pl2);
        }
        pl2 = pl1;
        pl1 = pv;
        v = make_double2(pv * cm, pv * sm);
      }
      *reinterpret_cast<double2*>(out + Y.yoff[l] + d2_pos(nt, rp, 2 * m)) = v;
    }
  }
}

// K2a part 1 in the operand layout: one thread per (pair, row j, column k) of the padded matrix
__global__ void sph_bessel2_kernel(const double* __restrict__ RA, const double* __restrict__ RB,
                                   const int* __restrict__ gid, const __grid_constant__ D2Layout Y, double sigma,
                                   size_t npairs, double* __restrict__ Bsw) {
  const int natoms = Y.natoms, L = Y.L;
  const size_t NN = (size_t)Y.bsz;
  const size_t total = npairs * NN;
  const double fact = 4.0 * kPi25 * sigma * sigma * sigma;
  const double inv2s2 = 0.5 / (sigma * sigma);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t p = t / NN;
    const int jk = (int)(t - p * NN);
    const int j = jk / Y.N8, k = jk - j * Y.N8;
    double* out = Bsw + p * (size_t)(L + 1) * NN + d2_pos(Y.NCT, j, k);
    if (j >= natoms || k >= natoms || gid[j] != gid[k]) {
      for (int l = 0; l <= L; ++l) out[(size_t)l * NN] = 0.0;
      continue;
    }
    bessel_levels(RA[p * natoms + j], RB[p * natoms + k], L, fact, inv2s2, out, NN);
  }
}

template <int NCT>
__global__ void __launch_bounds__(D2_THREADS)
sph_direct2_kernel(const double* __restrict__ YA, const double* __restrict__ YB, const double* __restrict__ Bsw,
                   const __grid_constant__ D2Layout Y, int nslots, size_t npairs, double2* __restrict__ Ihalf,
                   double2* __restrict__ Ipk) {
  // Ipk (optional): the same coefficients in the packed layout of the fast iSOFT kernels as well
  // ([pair][level][shell-ordered entry (a, m2)][+-a], sph_ipack_kernel), saving that kernel's pass over Ihalf
  extern __shared__ __align__(128) double smq[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smq);  // [8]
  double* slots = smq + 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = Y.L, L1 = L + 1, W = 2 * L + 1, N8 = 8 * NCT;
  if (tid == 0) {
    for (int s_ = 0; s_ < nslots; ++s_) fo_mbar_init(full + s_, 1);
    fo_mbar_fence_init();
  }
  __syncthreads();
  if (warp == D2_CONS) {  // producer warp: lane 0 feeds the ring
    unsigned it = 0;
    const unsigned bytesB = (unsigned)Y.bsz * 8;
    for (size_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
      for (int l = L; l >= 0; --l, ++it) {
        const int sl = (int)(it % (unsigned)nslots);
        const unsigned use = it / (unsigned)nslots;
        if (use > 0) {
          // the four consumer warps have released the slot: named barrier 1 + slot, on which they only arrive
          // (after their last fragment load, whose values the DMMAs before it have consumed) and this warp waits
          asm volatile("bar.sync %0, %1;" ::"r"(sl + 1), "r"(D2_THREADS) : "memory");
        }
        if (lane == 0) {
          fo_fence_proxy_async();  // the bulk copies below are async-proxy writes
          const unsigned bytesY = (unsigned)(N8 * 8 * Y.nrt[l]) * 8;
          fo_mbar_arrive_expect_tx(full + sl, 2 * bytesY + bytesB);
          double* dst = slots + (size_t)sl * Y.slot;
          fo_bulk_g2s(dst, YA + pair * (size_t)Y.ysz + Y.yoff[l], bytesY, full + sl);
          fo_bulk_g2s(dst + Y.ymax, YB + pair * (size_t)Y.ysz + Y.yoff[l], bytesY, full + sl);
          fo_bulk_g2s(dst + 2 * Y.ymax, Bsw + (pair * L1 + l) * (size_t)Y.bsz, bytesB, full + sl);
        }
        __syncwarp();
      }
    return;
  }
  const int g = lane >> 2, t4 = lane & 3;
  const size_t nitems = blockIdx.x < npairs ? ((npairs - 1 - blockIdx.x) / gridDim.x + 1) * (size_t)L1 : 0;
  const int swB = d2_sw(NCT, t4);  // (4 ks + t4) & 3 == t4 and ((4 ks + t4) >> 1) & 1 == (t4 >> 1) & 1
  unsigned it = 0;
  for (size_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    double2* out = Ihalf + pair * (size_t)L1 * W * L1;
    double2* opk = Ipk ? Ipk + pair * (size_t)((L + 1) * (L + 2) * (2 * L + 3) / 3) : nullptr;
    for (int l = L; l >= 0; --l, ++it) {
      double2* opl = opk ? opk + (size_t)(l * (l + 1) * (2 * l + 1) / 6) * 2 : nullptr;  // entries below level l
      const int sl = (int)(it % (unsigned)nslots);
      const unsigned use = it / (unsigned)nslots;
      fo_mbar_wait(full + sl, (int)(use & 1));
      const double* A1 = slots + (size_t)sl * Y.slot;
      const double* B2 = A1 + Y.ymax;
      const double* B1 = B2 + Y.ymax;
      const int nrt = Y.nrt[l], R8 = 8 * nrt;
      const int swY = d2_sw(nrt, t4);
      // the row tiles of an item go to the warps in an order rotated by the item: every warp gets a quarter of
      // the row tiles of a pair (L = 15: 10 of 40)
      for (int rt1 = (warp - (int)it) & 3; rt1 < nrt; rt1 += D2_CONS) {
        double c1[NCT][2];
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) c1[ct][0] = c1[ct][1] = 0.0;
        {
          const double* a = A1 + t4 * R8 + ((rt1 ^ swY) << 3) + g;
          const double* b = B1 + t4 * N8 + g;
#pragma unroll 2
          for (int ks = 0; ks < 2 * NCT; ++ks) {
            const double av = a[ks * 4 * R8];
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) fo_dmma(c1[ct], av, b[ks * 4 * N8 + ((ct ^ swB) << 3)]);
          }
        }
        for (int rb = 0; rb < nrt; rb += 4) {
          double c2[4][2];
#pragma unroll
          for (int q = 0; q < 4; ++q) c2[q][0] = c2[q][1] = 0.0;
#pragma unroll
          for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double* b = B2 + (8 * ct + 4 * h + t4) * R8 + g;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (rb + q < nrt) fo_dmma(c2[q], c1[ct][h], b[((rb + q) ^ swY) << 3]);
            }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (rb + q >= nrt) break;  // warp-uniform
            // this lane: row r1 = 8 rt1 + g (m1 = r1 / 2, part g & 1), columns (re, im) of m2 = 4 rt2 + t4
            const double px0 = __shfl_xor_sync(0xffffffffu, c2[q][0], 4);
            const double px1 = __shfl_xor_sync(0xffffffffu, c2[q][1], 4);
            const int m1 = (rt1 * 8 + g) >> 1, m2 = (rb + q) * 4 + t4;
            if (m1 > l || m2 > l) continue;
            // packed entry of (a = m1, m2): shell s = max(a, m2), t = s^2 + (m2 == s ? a : s + 1 + m2)
            const int sh = m1 > m2 ? m1 : m2;
            const int tpk = 2 * (sh * sh + (m2 == sh ? m1 : sh + 1 + m2));
            if ((g & 1) == 0) {  // rr, ri here; ir, ii in the partner: I(+m1, m2)
              const double2 v = make_double2(c2[q][0] + px1, px0 - c2[q][1]);
              out[((size_t)m2 * W + (L + m1)) * L1 + l] = v;
              if (opl) {
                opl[tpk] = v;
                if (m1 == 0) opl[tpk + 1] = v;
              }
            } else if (m1 > 0) {  // ir, ii here; rr, ri in the partner: I(-m1, m2) = (-1)^m1 (rr - ii, -ir - ri)
              const double sg = (m1 & 1) ? -1.0 : 1.0;
              const double2 v = make_double2(sg * (px0 - c2[q][1]), sg * (-c2[q][0] - px1));
              out[((size_t)m2 * W + (L - m1)) * L1 + l] = v;
              if (opl) opl[tpk + 1] = v;
            }
          }
        }
      }
      // release the slot -- unless it is never refilled (the producer does not wait for the last nslots items)
      if ((size_t)it + nslots < nitems) asm volatile("bar.arrive %0, %1;" ::"r"(sl + 1), "r"(D2_THREADS) : "memory");
    }
  }
}

// ring slots of sph_direct2_kernel that fit two CTAs per SM (0: the kernel does not apply); option
// sph_direct_ring: -1 = sph_direct_mma_kernel instead, n > 0 = at most n slots (fewer slots, more CTAs per SM)
int direct2_slots(const fo_ctx* ctx, int64_t natoms, int L) {
  if (ctx->force_generic || ctx->opt("sph_direct_ring") < 0 || natoms >= ctx->direct_gemm_min || natoms > 64 ||
      L >= D2_MAXL)
    return 0;
  const D2Layout Y((int)natoms, L);
  int n = (int)(((size_t)108 * 1024 - 128) / ((size_t)Y.slot * 8));             // two CTAs per SM
  if (n < 2) n = (int)(((size_t)220 * 1024 - 128) / ((size_t)Y.slot * 8));      // large clusters / bandwidths: one
  if (const int64_t o = ctx->opt("sph_direct_ring")) n = std::min<int>(n, (int)o);
  return n >= 2 ? std::min(n, 8) : 0;
}
size_t direct2_smem(const D2Layout& Y, int nslots) { return 128 + (size_t)nslots * Y.slot * 8; }

// The (m2, m1, l < max(|m1|, m2)) entries are never read; no need to clear Ihalf.

// Full numpy layout [l][m1 wrap][m2 wrap] -> Ihalf, keeping the part that generates the REAL grid:
// Ihalf = (I[l,m1,m2] + (-1)^{m1+m2} conj(I[l,-m1,-m2]))/2 (identity for coefficients of real
// densities).  part = 1 selects the part that generates the IMAGINARY grid instead.
__global__ void sph_pack_kernel(const double2* __restrict__ full, int L, size_t npairs, int part,
                                double2* __restrict__ Ihalf) {
  const int W = 2 * L + 1, L1 = L + 1;
  const size_t per = (size_t)L1 * W * L1;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < npairs * per;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t p = t / per;
    int r = (int)(t - p * per);
    const int l = r % L1;
    r /= L1;
    const int m1 = r % W - L, m2 = r / W;
    double2 v = make_double2(0.0, 0.0);
    const int am1 = m1 < 0 ? -m1 : m1;
    if (l >= am1 && l >= m2) {
      const double2* f = full + p * (size_t)L1 * W * W + (size_t)l * W * W;
      const double2 a = f[((m1 + W) % W) * W + m2];
      double2 b = f[((-m1 + W) % W) * W + ((-m2 + W) % W)];
      const double sg = ((m1 + m2) & 1) ? -1.0 : 1.0;
      b = make_double2(sg * b.x, -sg * b.y);  // (-1)^{m1+m2} conj
      if (part == 0)
        v = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
      else  // (a - b) / (2i)
        v = make_double2(0.5 * (a.y - b.y), -0.5 * (a.x - b.x));
    }
    Ihalf[t] = v;
  }
}

__global__ void sph_unpack_kernel(const double2* __restrict__ Ihalf, int L, size_t npairs,
                                  double2* __restrict__ full) {
  const int W = 2 * L + 1, L1 = L + 1;
  const size_t per = (size_t)L1 * W * W;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < npairs * per;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t p = t / per;
    int r = (int)(t - p * per);
    const int i2 = r % W;
    r /= W;
    const int i1 = r % W, l = r / W;
    const int m1 = i1 <= L ? i1 : i1 - W, m2 = i2 <= L ? i2 : i2 - W;
    double2 v = make_double2(0.0, 0.0);
    const int am1 = m1 < 0 ? -m1 : m1, am2 = m2 < 0 ? -m2 : m2;
    if (l >= am1 && l >= am2) {
      const double2* h = Ihalf + p * (size_t)L1 * W * L1;
      if (m2 >= 0) {
        v = h[((size_t)m2 * W + (m1 + L)) * L1 + l];
      } else {
        v = h[((size_t)(-m2) * W + (-m1 + L)) * L1 + l];
        const double sg = ((m1 + m2) & 1) ? -1.0 : 1.0;
        v = make_double2(sg * v.x, -sg * v.y);
      }
    }
    full[t] = v;
  }
}

// Ypk[struct][atom][lm] (m >= 0) -> the reference's sphHarm layout Y[l][m wrap][atom]
// (sphericalAlignment.py:57-65), Y_{l,-m} = (-1)^m conj(Y_lm); entries with |m| > l are 0.
__global__ void sph_ylm_expand_kernel(const double2* __restrict__ Ypk, int natoms, int L, size_t nstruct,
                                      double2* __restrict__ out) {
  const int W = 2 * L + 1, L1 = L + 1, NLM = nlm_of(L);
  const size_t per = (size_t)L1 * W * natoms;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < nstruct * per;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t sidx = t / per;
    size_t r = t - sidx * per;
    const int a = (int)(r % natoms);
    r /= natoms;
    const int mi = (int)(r % W), l = (int)(r / W);
    const int m = mi <= L ? mi : mi - W;
    const int am = m < 0 ? -m : m;
    double2 v = make_double2(0.0, 0.0);
    if (am <= l) {
      v = Ypk[(sidx * natoms + a) * NLM + l * (l + 1) / 2 + am];
      if (m < 0) {
        const double sg = (am & 1) ? -1.0 : 1.0;
        v = make_double2(sg * v.x, -sg * v.y);
      }
    }
    out[t] = v;
  }
}

// ------------------------------------------------------------------------------------------
// K2b: harmonic-basis coefficients.  d_nl(r) in closed form (Gradshteyn-Ryzhik 7.421.4):
//   d_nl = 4 pi N_nl sqrt(pi/2) 2^{-l-3/2} beta^{-l-3/2} y^l exp(-r^2/(2(s^2+r0^2))) Q_n,
//   (n+1) Q_{n+1} = ((2n+l+3/2) delta - kappa) Q_n - (n+l+1/2) delta^2 Q_{n-1}
// which is what radialIntegralHarmonic + coeffs_harmonicBasis (sphericalAlignment.py:288-360)
// and HARMONICNL (fastclusters.f90:492-540) evaluate, without their cancellation.
// One CTA per (structure, group); atoms in tiles; C[n][lm] += d[n][l] conj(Y[lm]).
// ------------------------------------------------------------------------------------------
constexpr int HARM_TA = 8;

__global__ void __launch_bounds__(256)
sph_harm_kernel(const double2* __restrict__ Ypk, const double* __restrict__ R,
                const int32_t* __restrict__ goff, const int32_t* __restrict__ gidx, int ngroups,
                int natoms, int nmax, int L, double r0, double sigma, double2* __restrict__ Cpk) {
  extern __shared__ double smh[];
  const int s = blockIdx.x, g = blockIdx.y;
  const int NLM = nlm_of(L), L1 = L + 1, N1 = nmax + 1;
  double* dnl = smh;                                   // [HARM_TA][N1][L1]
  double2* ys = (double2*)(dnl + HARM_TA * N1 * L1);   // [HARM_TA][NLM]
  const int nout = N1 * NLM;
  // each thread owns outputs o = tid + q*blockDim, o = n*NLM + lm ; keep up to 16 in registers
  constexpr int MAXQ = 16;
  double2 acc[MAXQ];
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) acc[q] = make_double2(0.0, 0.0);
  const double s2 = sigma * sigma, r02 = r0 * r0;
  const double beta = 0.5 * (1.0 / r02 + 1.0 / s2), alpha = 1.0 / r02;
  const double delta = (beta - alpha) / beta;
  const int a_begin = goff[g], a_end = goff[g + 1];
  for (int a0 = a_begin; a0 < a_end; a0 += HARM_TA) {
    const int ta = min(HARM_TA, a_end - a0);
    __syncthreads();
    for (int t = threadIdx.x; t < ta * L1; t += blockDim.x) {
      const int a = t / L1, l = t - a * L1;
      const int atom = gidx[a0 + a];
      const double r = R[(size_t)s * natoms + atom];
      const double y = r / s2;
      const double kappa = alpha * y * y / (4.0 * beta * beta);
      const double nu = l + 0.5;
      double norm = sqrt(2.0 * pow(r0, -2.0 * l - 3.0) / tgamma(l + 1.5));
      const double pref = 4.0 * kPi * sqrt(kPi / 2.0) * pow(2.0, -nu - 1.0) * pow(beta, -nu - 1.0) *
                          pow(y, (double)l) * exp(-0.5 * r * r / (s2 + r02));
      double qm1 = 0.0, q = 1.0;
      for (int n = 0; n <= nmax; ++n) {
        dnl[(a * N1 + n) * L1 + l] = pref * norm * q;
        const double qn = (((2.0 * n + 1.0 + nu) * delta - kappa) * q - (n + nu) * delta * delta * qm1) / (n + 1.0);
        qm1 = q;
        q = qn;
        norm *= sqrt((n + 1.0) / (n + 1.0 + nu));
      }
    }
    for (int t = threadIdx.x; t < ta * NLM; t += blockDim.x) {
      const int a = t / NLM, lm = t - a * NLM;
      ys[t] = Ypk[((size_t)s * natoms + gidx[a0 + a]) * NLM + lm];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      const int o = threadIdx.x + q * blockDim.x;
      if (o < nout) {
        const int n = o / NLM, lm = o - n * NLM;
        // l from lm: l = floor((sqrt(8 lm + 1) - 1)/2)
        int l = (int)((sqrt(8.0 * lm + 1.0) - 1.0) * 0.5);
        while ((l + 1) * (l + 2) / 2 <= lm) ++l;
        while (l * (l + 1) / 2 > lm) --l;
        double ar = acc[q].x, ai = acc[q].y;
        for (int a = 0; a < ta; ++a) {
          const double d = dnl[(a * N1 + n) * L1 + l];
          const double2 yv = ys[a * NLM + lm];
          ar = fma(d, yv.x, ar);
          ai = fma(-d, yv.y, ai);
        }
        acc[q] = make_double2(ar, ai);
      }
    }
  }
  double2* out = Cpk + ((size_t)s * ngroups + g) * nout;
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    const int o = threadIdx.x + q * blockDim.x;
    if (o < nout) out[o] = acc[q];
  }
}

// Cpk (m >= 0) -> numpy layout [n][l][m wrap]: C[n,l,-m] = (-1)^m conj(C[n,l,m]).
__global__ void sph_harm_expand_kernel(const double2* __restrict__ Cpk, int nmax, int L, size_t nsg,
                                       double2* __restrict__ full) {
  const int W = 2 * L + 1, L1 = L + 1, N1 = nmax + 1, NLM = nlm_of(L);
  const size_t per = (size_t)N1 * L1 * W;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < nsg * per;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t sg = t / per;
    int r = (int)(t - sg * per);
    const int im = r % W;
    r /= W;
    const int l = r % L1, n = r / L1;
    const int m = im <= L ? im : im - W;
    const int am = m < 0 ? -m : m;
    double2 v = make_double2(0.0, 0.0);
    if (am <= l) {
      v = Cpk[sg * (size_t)N1 * NLM + (size_t)n * NLM + l * (l + 1) / 2 + am];
      if (m < 0) {
        const double sg2 = (am & 1) ? -1.0 : 1.0;
        v = make_double2(sg2 * v.x, -sg2 * v.y);
      }
    }
    full[t] = v;
  }
}

// ------------------------------------------------------------------------------------------
// K3: I[l,m1,m2] = sum_g sum_n conj(C^A[g,n,l,m1]) C^B[g,n,l,m2] for m2 >= 0, all m1, written in
// the Ihalf layout; also avg = sum_{l,m1,m2 (all)} |I|^2 (CALCSIMILARITY fastclusters.f90:790-818).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sph_dot_kernel(const double2* __restrict__ bank, const long long* __restrict__ pairs, int ngroups,
               int nmax, int L, double2* __restrict__ Ihalf, double* __restrict__ avg) {
  __shared__ double red[8];
  const size_t p = blockIdx.x;
  const int W = 2 * L + 1, L1 = L + 1, N1 = nmax + 1, NLM = nlm_of(L);
  const size_t per_struct = (size_t)ngroups * N1 * NLM;
  const double2* CA = bank + (size_t)pairs[2 * p] * per_struct;
  const double2* CB = bank + (size_t)pairs[2 * p + 1] * per_struct;
  double2* out = Ihalf + p * (size_t)L1 * W * L1;
  const int nitem = L1 * W * L1;
  double norm = 0.0;
  for (int t = threadIdx.x; t < nitem; t += blockDim.x) {
    const int l = t % L1;
    int r = t / L1;
    const int m1 = r % W - L, m2 = r / W;
    const int am1 = m1 < 0 ? -m1 : m1;
    if (l < am1 || l < m2) continue;
    const int lb = l * (l + 1) / 2;
    double ar = 0.0, ai = 0.0;
    for (int g = 0; g < ngroups; ++g)
      for (int n = 0; n < N1; ++n) {
        double2 a = CA[((size_t)g * N1 + n) * NLM + lb + am1];
        if (m1 < 0) {  // C[-m] = (-1)^m conj(C[m])
          const double sg = (am1 & 1) ? -1.0 : 1.0;
          a = make_double2(sg * a.x, -sg * a.y);
        }
        const double2 b = CB[((size_t)g * N1 + n) * NLM + lb + m2];
        // conj(a) * b
        ar += a.x * b.x + a.y * b.y;
        ai += a.x * b.y - a.y * b.x;
      }
    out[t] = make_double2(ar, ai);
    // the (-m1,-m2) partner has the same modulus; m2 = 0 rows are their own partners' mirror
    norm += (m2 == 0 ? 1.0 : 2.0) * (ar * ar + ai * ai);
  }
  if (avg) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) norm += __shfl_down_sync(0xffffffffu, norm, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = norm;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
      avg[p] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K3 on the tensor pipe (sph_dot_mma_kernel): one CTA per pair.  Both structures' packed coefficients
// [k = (group, n)][lm] are copied to shared memory once (cp.async, k-major, pitch == 24 mod 32 doubles);
// for every l the four real products of  I = sum_k conj(C^A[k, l, a]) C^B[k, l, m2]  are one small GEMM
// on interleaved re / im rows and columns (as in sph_direct_mma_kernel), and
//   I(+a, m2) = (RR + II, RI - IR),   I(-a, m2) = (-1)^a (RR - II, RI + IR)
// because C[n, l, -a] = (-1)^a conj(C[n, l, a]).  avg = sum |I|^2 is reduced in a fixed order.
// ------------------------------------------------------------------------------------------
constexpr int DOT_THREADS = 256;

__global__ void __launch_bounds__(DOT_THREADS)
sph_dot_mma_kernel(const double2* __restrict__ bank, const long long* __restrict__ pairs, int ngroups,
                   int nmax, int L, double2* __restrict__ Ihalf, double* __restrict__ avg) {
  extern __shared__ double smd2[];
  const size_t p = blockIdx.x;
  const int W = 2 * L + 1, L1 = L + 1, N1 = nmax + 1, NLM = nlm_of(L);
  const int K = ngroups * N1, K4 = (K + 3) & ~3;
  const int LDC = fo_ld8(2 * NLM) + 16;  // == 24 (mod 32): the four k rows of a fragment in distinct bank groups
  double* As = smd2;                      // [K4][LDC]
  double* Bs = As + (size_t)K4 * LDC;     // [K4][LDC]
  double* red = Bs + (size_t)K4 * LDC;    // [8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const size_t per_struct = (size_t)K * NLM;
  const double2* CA = bank + (size_t)pairs[2 * p] * per_struct;
  const double2* CB = bank + (size_t)pairs[2 * p + 1] * per_struct;
  for (int e = tid; e < K4 * NLM; e += DOT_THREADS) {
    const int k = e / NLM, c = e - k * NLM;
    double* da = As + (size_t)k * LDC + 2 * c;
    double* db = Bs + (size_t)k * LDC + 2 * c;
    if (k < K) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(da), sb = (unsigned)__cvta_generic_to_shared(db);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(CA + (size_t)k * NLM + c) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb), "l"(CB + (size_t)k * NLM + c) : "memory");
    } else {
      da[0] = da[1] = db[0] = db[1] = 0.0;
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  double2* out = Ihalf + p * (size_t)L1 * W * L1;
  double norm = 0.0;
  // jobs: (l, row tile, column tile), enumerated l-major; nrt(l) = ceil(2 (l+1) / 8)
  int job = warp;
  for (int l = 0, j0 = 0; l <= L; ++l) {
    const int nrt = (2 * (l + 1) + 7) >> 3;
    const int lb2 = l * (l + 1);  // 2 * lbase: first double of this l in a row
    for (; job < j0 + nrt * nrt; job += DOT_THREADS / 32) {
      const int q = job - j0;
      const int rt1 = q % nrt, rt2 = q / nrt;
      double acc[2] = {0.0, 0.0};
      // rows / columns beyond 2 (l+1) read the next l's coefficients (finite, results discarded)
      const double* a = As + t4 * LDC + lb2 + rt1 * 8 + g;
      const double* b = Bs + t4 * LDC + lb2 + rt2 * 8 + g;
      for (int ks = 0; ks < (K4 >> 2); ++ks) fo_dmma(acc, a[ks * 4 * LDC], b[ks * 4 * LDC]);
      // this lane: row r1 = 8 rt1 + g (a = r1 / 2, part g & 1), columns (re, im) of m2 = 4 rt2 + t4
      const double px0 = __shfl_xor_sync(0xffffffffu, acc[0], 4);
      const double px1 = __shfl_xor_sync(0xffffffffu, acc[1], 4);
      const int am = (rt1 * 8 + g) >> 1, m2 = rt2 * 4 + t4;
      if (am > l || m2 > l) continue;
      double2 v;
      if ((g & 1) == 0) {  // RR, RI here; IR, II in the partner: I(+a, m2)
        v = make_double2(acc[0] + px1, acc[1] - px0);
        out[((size_t)m2 * W + (L + am)) * L1 + l] = v;
      } else if (am > 0) {  // IR, II here; RR, RI in the partner: I(-a, m2)
        const double sg = (am & 1) ? -1.0 : 1.0;
        v = make_double2(sg * (px0 - acc[1]), sg * (px1 + acc[0]));
        out[((size_t)m2 * W + (L - am)) * L1 + l] = v;
      } else {
        continue;
      }
      // the (-m1, -m2) partner has the same modulus
      norm += (m2 == 0 ? 1.0 : 2.0) * (v.x * v.x + v.y * v.y);
    }
    j0 += nrt * nrt;
  }
  if (avg) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) norm += __shfl_down_sync(0xffffffffu, norm, off);
    if (lane == 0) red[warp] = norm;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < DOT_THREADS / 32; ++w) t += red[w];
      avg[p] = t;
    }
  }
}

size_t dot_mma_smem(int ngroups, int nmax, int L) {
  const int K4 = (ngroups * (nmax + 1) + 3) & ~3;
  return ((size_t)2 * K4 * (fo_ld8(2 * nlm_of(L)) + 16) + 8) * 8;
}

// ------------------------------------------------------------------------------------------
// K4: Wigner-d table in the Dt layout.  One thread per (m2 >= 0, m1, k): closed-form edge value
// at l = max(|m1|, m2) then the Kostelec-Rockmore recurrence in l
// (CALCWIGNERD + RECURRTERMS, DSOFT.f90:81-195).
// ------------------------------------------------------------------------------------------
__device__ double wigner_edge(int J, int m1, int m2, double cb2, double sb2) {
  // d^J_{m1 m2} with max(|m1|,|m2|) = J, times sqrt((2J+1)/2)
  int m;      // the other index
  double c, s;  // bases
  int pc, ps;
  if (m1 == J) {
    m = m2; c = cb2; s = -sb2; pc = J + m; ps = J - m;
  } else if (m1 == -J) {
    m = m2; c = cb2; s = sb2; pc = J - m; ps = J + m;
  } else if (m2 == J) {
    m = m1; c = cb2; s = sb2; pc = J + m; ps = J - m;
  } else {
    m = m1; c = cb2; s = -sb2; pc = J - m; ps = J + m;
  }
  const int am = m < 0 ? -m : m;
  double binom = 1.0;  // C(2J, J+m)
  for (int i = 1; i <= J - am; ++i) binom *= (double)(J + am + i) / (double)i;
  double v = sqrt((2.0 * J + 1.0) * 0.5 * binom);
  for (int i = 0; i < pc; ++i) v *= c;
  for (int i = 0; i < ps; ++i) v *= s;
  return v;
}

// Element (item = m2 W + m1 + L, l, k) of the table lives at item sI + l sL + k sK: the standard layout
// Dt[m2][m1][l][k] has (sI, sL, sK) = (L1 F, F, 1); large bandwidths use the plane-major layout
// DtK[k][m2][m1][l] = (L1, 1, L1 W L1), whose beta planes are contiguous (sph_isoft_big_kernel).
struct DtStride {
  long long sI, sL, sK;
};
__host__ __device__ inline DtStride dt_stride(int L, bool kmajor) {
  const long long L1 = L + 1, W = 2 * L + 1, F = 2 * L + 2;
  DtStride s;
  if (kmajor) {
    s.sI = L1; s.sL = 1; s.sK = L1 * W * L1;
  } else {
    s.sI = L1 * F; s.sL = F; s.sK = 1;
  }
  return s;
}

__global__ void sph_wigner_kernel(int L, DtStride ds, double* __restrict__ Dt) {
  const int B = L + 1, W = 2 * L + 1, L1 = L + 1, NK = 2 * B;
  const int total = L1 * W * NK;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int k = t % NK;
    int r = t / NK;
    const int m1 = r % W - L, m2 = r / W;
    const int am1 = m1 < 0 ? -m1 : m1;
    const int J0 = am1 > m2 ? am1 : m2;
    const double beta = kPi * (2.0 * k + 1.0) / (4.0 * B);
    double sb2, cb2;
    sincos(0.5 * beta, &sb2, &cb2);
    const double cb = cos(beta);
    double* col = Dt + ((size_t)m2 * W + (m1 + L)) * ds.sI + (size_t)k * ds.sK;
    for (int l = 0; l < J0; ++l) col[(size_t)l * ds.sL] = 0.0;
    double dm1 = 0.0, d = wigner_edge(J0, m1, m2, cb2, sb2);
    col[(size_t)J0 * ds.sL] = d;
    for (int J = J0; J < L; ++J) {
      const double dj = J, a1 = m1, a2 = m2;
      const double t1 = sqrt((2.0 * dj + 3.0) / (2.0 * dj + 1.0));
      const double t3 = (dj + 1.0) * (2.0 * dj + 1.0);
      const double t5 = 1.0 / sqrt(((dj + 1.0) * (dj + 1.0) - a1 * a1) * ((dj + 1.0) * (dj + 1.0) - a2 * a2));
      const double Bc = t1 * t3 * t5;
      double A = 0.0, C = 0.0;
      if (J > 0) {
        const double t2 = sqrt((2.0 * dj + 3.0) / (2.0 * dj - 1.0)) * (dj + 1.0) / dj;
        const double t4 = sqrt((dj * dj - a1 * a1) * (dj * dj - a2 * a2));
        A = t2 * t4 * t5;
        C = a1 * a2 / (dj * (dj + 1.0));
      }
      const double dn = Bc * (cb - C) * d - A * dm1;
      dm1 = d;
      d = dn;
      col[(size_t)(J + 1) * ds.sL] = d;
    }
  }
}

// Dt -> the reference's Ds[l][m1 wrap][m2 wrap][k] (soft.py:73-96) using
// d^l_{m1,-m2} = (-1)^{l+m1} d^l_{m1,m2}(pi - beta)  and beta_{2B-1-k} = pi - beta_k.
__global__ void sph_wigner_export_kernel(const double* __restrict__ Dt, DtStride ds, int L,
                                         double* __restrict__ Ds) {
  const int B = L + 1, W = 2 * L + 1, NK = 2 * B;
  const size_t total = (size_t)B * W * W * NK;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(t % NK);
    int r = (int)(t / NK);
    const int i2 = r % W;
    r /= W;
    const int i1 = r % W, l = r / W;
    const int m1 = i1 <= L ? i1 : i1 - W, m2 = i2 <= L ? i2 : i2 - W;
    const int am1 = m1 < 0 ? -m1 : m1, am2 = m2 < 0 ? -m2 : m2;
    double v = 0.0;
    if (l >= am1 && l >= am2) {
      if (m2 >= 0) {
        v = Dt[((size_t)m2 * W + (m1 + L)) * ds.sI + (size_t)l * ds.sL + (size_t)k * ds.sK];
      } else {
        v = Dt[((size_t)(-m2) * W + (m1 + L)) * ds.sI + (size_t)l * ds.sL + (size_t)(NK - 1 - k) * ds.sK];
        if ((l + m1) & 1) v = -v;
      }
    }
    Ds[t] = v;
  }
}

// ------------------------------------------------------------------------------------------
// K5-K7: inverse SO(3) transform of one (pair, chunk of KC beta planes), both orientations.
//   S_par[m1][kk][m2] = sum_{l = par mod 2} Dt[m2][m1][l][k] I[m2][m1][l]      par = even / odd l
//   orientation o: S = S_even + (o ? -1 : +1) S_odd          (I_inv^l = (-1)^l I^l)
//   A  U[a][kk][m2] = sum_m1 S[m1][kk][m2] e^{+2 pi i m1 a/2B}
//   B  g[a][k][gam] = U[a][kk][0] + 2 sum_{m2>=1} Re(U[a][kk][m2] e^{+2 pi i m2 gam/2B})
// with the +-m pairing of the 1-D transforms (see fo_periodic.cu), then the running arg-max.
// ------------------------------------------------------------------------------------------
constexpr int IS_THREADS = 256;
constexpr int IS_DCA = 5;
constexpr int IS_DCB = 9;

struct IsoOut {
  double* part_val;      // [P][O][nchunk]
  long long* part_idx;   // [P][O][nchunk]
  double* grid;          // [P][O][2B][2B][2B] or null
};

__device__ __forceinline__ void better_s(double& bv, long long& bi, double v, long long i) {
  if (v > bv || (v == bv && i < bi)) {
    bv = v;
    bi = i;
  }
}

template <int KC>
__global__ void __launch_bounds__(IS_THREADS)
sph_isoft_kernel(const double2* __restrict__ Ihalf, const double* __restrict__ Dt, int L, int norient,
                 int nchunk, IsoOut out) {
  extern __shared__ double2 smi[];
  const int B = L + 1, W = 2 * L + 1, L1 = L + 1, F = 2 * B, H = B + 1;
  const int M2p = L1 | 1;
  const int NL = KC * L1;  // lines of stage A: (kk, m2)
  double2* tw = smi;                          // [F]
  double2* Se = tw + F;                       // [W][KC][L1]
  double2* So = Se + (size_t)W * NL;          // [W][KC][L1]
  double2* U = So + (size_t)W * NL;           // [F][KC][M2p]
  double* red = (double*)(U + (size_t)F * KC * M2p);  // 32 doubles
  const int tid = threadIdx.x;
  const size_t p = blockIdx.x / nchunk;
  const int chunk = blockIdx.x % nchunk;
  const int k0 = chunk * KC;
  for (int t = tid; t < F; t += IS_THREADS) {
    double sn, cs;
    sincospi(2.0 * (double)t / (double)F, &sn, &cs);
    tw[t] = make_double2(cs, sn);
  }
  // ---- K5
  const double2* Ip = Ihalf + p * (size_t)L1 * W * L1;
  for (int it = tid; it < L1 * W; it += IS_THREADS) {
    // m2 fastest across lanes: conflict-free shared stores below
    const int m2 = it % L1, m1i = it / L1;
    const int item = m2 * W + m1i;  // index in the Ihalf / Dt layouts
    const int m1 = m1i - L;
    const int am1 = m1 < 0 ? -m1 : m1;
    const int l0 = am1 > m2 ? am1 : m2;
    double2 ae[KC], ao[KC];
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      ae[kk] = make_double2(0.0, 0.0);
      ao[kk] = make_double2(0.0, 0.0);
    }
    const double2* ip = Ip + (size_t)item * L1;
    const double* dp = Dt + (size_t)item * L1 * F + k0;
    for (int l = l0; l <= L; ++l) {
      const double2 c = ip[l];
      const double* d = dp + (size_t)l * F;
      if (l & 1) {
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
          const double dv = d[kk];
          ao[kk].x = fma(dv, c.x, ao[kk].x);
          ao[kk].y = fma(dv, c.y, ao[kk].y);
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
          const double dv = d[kk];
          ae[kk].x = fma(dv, c.x, ae[kk].x);
          ae[kk].y = fma(dv, c.y, ae[kk].y);
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      Se[(size_t)m1i * NL + kk * L1 + m2] = ae[kk];
      So[(size_t)m1i * NL + kk * L1 + m2] = ao[kk];
    }
  }
  __syncthreads();
  for (int o = 0; o < norient; ++o) {
    const double so = o ? -1.0 : 1.0;
    // ---- stage A: lines (kk, m2), inputs over m1 = -L..L, outputs a = 0..F-1
    {
      const int nch = (H + IS_DCA - 1) / IS_DCA;
      for (int item = tid; item < NL * nch; item += IS_THREADS) {
        const int line = item % NL, ch = item / NL;
        const int kk = line / L1, m2 = line - kk * L1;
        const int d0 = ch * IS_DCA;
        const double2* se = Se + line;
        const double2* sod = So + line;
        double2 c0 = se[(size_t)L * NL];
        {
          const double2 c0o = sod[(size_t)L * NL];
          c0.x += so * c0o.x;
          c0.y += so * c0o.y;
        }
        double2 P[IS_DCA], Q[IS_DCA];
        int idx[IS_DCA];
#pragma unroll
        for (int t = 0; t < IS_DCA; ++t) {
          P[t] = c0;
          Q[t] = make_double2(0.0, 0.0);
          idx[t] = 0;
        }
        for (int m = 1; m <= L; ++m) {
          double2 a = se[(size_t)(L + m) * NL], b = se[(size_t)(L - m) * NL];
          const double2 a2 = sod[(size_t)(L + m) * NL], b2 = sod[(size_t)(L - m) * NL];
          a.x += so * a2.x; a.y += so * a2.y;
          b.x += so * b2.x; b.y += so * b2.y;
          const double2 E = make_double2(a.x + b.x, a.y + b.y), O = make_double2(a.x - b.x, a.y - b.y);
#pragma unroll
          for (int t = 0; t < IS_DCA; ++t) {
            int k = idx[t] + d0 + t;
            k -= (k >= F) ? F : 0;
            idx[t] = k;
            const double2 w = tw[k];
            P[t].x = fma(E.x, w.x, P[t].x);
            P[t].y = fma(E.y, w.x, P[t].y);
            Q[t].x = fma(O.x, w.y, Q[t].x);
            Q[t].y = fma(O.y, w.y, Q[t].y);
          }
        }
#pragma unroll
        for (int t = 0; t < IS_DCA; ++t) {
          const int d = d0 + t;
          if (d < H) {
            // e^{+i}: U[d] = P + iQ, U[F-d] = P - iQ
            U[((size_t)d * KC + kk) * M2p + m2] = make_double2(P[t].x - Q[t].y, P[t].y + Q[t].x);
            if (d != 0 && 2 * d != F)
              U[((size_t)(F - d) * KC + kk) * M2p + m2] = make_double2(P[t].x + Q[t].y, P[t].y - Q[t].x);
          }
        }
      }
    }
    __syncthreads();
    // ---- stage B: lines (a, kk), half-complex -> real, arg-max
    double bv = -1e300;
    long long bi = 0x7fffffffffffffffLL;
    {
      const int nch = (H + IS_DCB - 1) / IS_DCB;
      for (int item = tid; item < F * KC * nch; item += IS_THREADS) {
        const int line = item % (F * KC), ch = item / (F * KC);
        const int a = line / KC, kk = line - a * KC;
        const int d0 = ch * IS_DCB;
        const double2* vin = U + (size_t)line * M2p;
        const double v0 = vin[0].x;
        double A[IS_DCB], Bq[IS_DCB];
        int idx[IS_DCB];
#pragma unroll
        for (int t = 0; t < IS_DCB; ++t) {
          A[t] = 0.0;
          Bq[t] = 0.0;
          idx[t] = 0;
        }
        for (int l = 1; l <= L; ++l) {
          const double2 v = vin[l];
#pragma unroll
          for (int t = 0; t < IS_DCB; ++t) {
            int k = idx[t] + d0 + t;
            k -= (k >= F) ? F : 0;
            idx[t] = k;
            const double2 w = tw[k];
            A[t] = fma(v.x, w.x, A[t]);
            Bq[t] = fma(v.y, w.y, Bq[t]);
          }
        }
        const long long base = ((long long)a * F + (k0 + kk)) * F;
        double* grow = out.grid ? out.grid + ((p * norient + o) * (size_t)F * F * F + (size_t)base) : nullptr;
#pragma unroll
        for (int t = 0; t < IS_DCB; ++t) {
          const int d = d0 + t;
          if (d < H) {
            const double aa = v0 + 2.0 * A[t], bb = 2.0 * Bq[t];
            const double g1 = aa - bb;  // Re(V e^{+i theta}) = Vr cos - Vi sin
            better_s(bv, bi, g1, base + d);
            if (grow) grow[d] = g1;
            if (d != 0 && 2 * d != F) {
              const double g2 = aa + bb;
              better_s(bv, bi, g2, base + (F - d));
              if (grow) grow[F - d] = g2;
            }
          }
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, bv, off);
      const long long oi = __shfl_down_sync(0xffffffffu, bi, off);
      better_s(bv, bi, ov, oi);
    }
    long long* redi = (long long*)(red + 16);
    if ((tid & 31) == 0) {
      red[tid >> 5] = bv;
      redi[tid >> 5] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < IS_THREADS / 32; ++w) better_s(bv, bi, red[w], redi[w]);
      out.part_val[(p * norient + o) * nchunk + chunk] = bv;
      out.part_idx[(p * norient + o) * nchunk + chunk] = bi;
    }
    __syncthreads();
  }
}

// Grid value at integer point (a, k, g) by the direct Wigner sum (orientation sign so on odd l):
// partial sum over the items tid, tid + nthreads, ... of one thread.
__device__ double iso_point_partial(const double2* __restrict__ Ip, const double* __restrict__ Dt, DtStride ds,
                                    int L, int a, int k, int g, double so, int tid, int nthreads) {
  const int W = 2 * L + 1, L1 = L + 1, F = 2 * L1;
  double acc = 0.0;
  for (int item = tid; item < L1 * W; item += nthreads) {
    const int m1i = item % W, m2 = item / W;
    const int m1 = m1i - L;
    const int am1 = m1 < 0 ? -m1 : m1;
    const int l0 = am1 > m2 ? am1 : m2;
    double sr = 0.0, si = 0.0;
    const double2* ip = Ip + (size_t)item * L1;
    const double* dp = Dt + (size_t)item * ds.sI + (size_t)k * ds.sK;
    for (int l = l0; l <= L; ++l) {
      const double dv = ((l & 1) ? so : 1.0) * dp[(size_t)l * ds.sL];
      const double2 c = ip[l];
      sr = fma(dv, c.x, sr);
      si = fma(dv, c.y, si);
    }
    int e = (m1 * a + m2 * g) % F;
    e += e < 0 ? F : 0;
    double sn, cs;
    sincospi(2.0 * (double)e / (double)F, &sn, &cs);
    const double term = sr * cs - si * sn;
    acc += (m2 == 0) ? term : 2.0 * term;
  }
  return acc;
}

// findMax (utils.py:319-338) of the generic / large-bandwidth paths.  One CTA per ((pair,
// orientation), neighbour w): every CTA reduces the chunk maxima (cheap), then all 256 threads share
// the Wigner sum of neighbour w (axis w >> 1, step +1 / -1) -- at Jmax = 63 that is 8128 (m1, m2)
// items of up to 64 levels, which a single warp per neighbour took 2.2 ms for (profiles/r01_summary.md).
__global__ void __launch_bounds__(256)
sph_final_kernel(const double2* __restrict__ Ihalf, const double* __restrict__ Dt, DtStride ds, int L,
                 int norient, int nchunk, const double* __restrict__ part_val,
                 const long long* __restrict__ part_idx, long long* __restrict__ best_idx,
                 double* __restrict__ best_val, double* __restrict__ nbv) {
  __shared__ double red[8];
  const size_t po = blockIdx.x;
  const int w = blockIdx.y;
  const size_t p = po / norient;
  const int o = (int)(po % norient);
  const int F = 2 * (L + 1);
  double bv = -1e300;
  long long bi = 0x7fffffffffffffffLL;
  for (int c = 0; c < nchunk; ++c) better_s(bv, bi, part_val[po * nchunk + c], part_idx[po * nchunk + c]);
  const bool ok = bi != 0x7fffffffffffffffLL;
  const int b3[3] = {ok ? (int)(bi / ((long long)F * F)) : 0, ok ? (int)((bi / F) % F) : 0,
                     ok ? (int)(bi % F) : 0};
  const int ax = w >> 1, sgn = (w & 1) ? -1 : 1;
  int q[3] = {b3[0], b3[1], b3[2]};
  q[ax] = (q[ax] + sgn + F) % F;
  double v = iso_point_partial(Ihalf + p * (size_t)(L + 1) * (2 * L + 1) * (L + 1), Dt, ds, L, q[0], q[1], q[2],
                               o ? -1.0 : 1.0, threadIdx.x, blockDim.x);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    nbv[po * 6 + w] = fabs(t);
    if (w == 0) {
      best_val[po] = bv;
      for (int a3 = 0; a3 < 3; ++a3) best_idx[po * 3 + a3] = b3[a3];
    }
  }
}

// parabola through the maximum and its two neighbours per axis -> fractional index
__global__ void sph_parabola_kernel(size_t npo, const long long* __restrict__ best_idx,
                                    const double* __restrict__ best_val, const double* __restrict__ nbv,
                                    double* __restrict__ frac_idx) {
  const size_t po = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (po >= npo) return;
  for (int ax = 0; ax < 3; ++ax) {
    const double y1 = nbv[po * 6 + 2 * ax], y3 = nbv[po * 6 + 2 * ax + 1], y2 = fabs(best_val[po]);
    frac_idx[po * 3 + ax] = (double)best_idx[po * 3 + ax] - (y3 - y1) / (2.0 * (2.0 * y2 - y1 - y3));
  }
}

// ------------------------------------------------------------------------------------------
// K5-K7 fast path (sph_isoft3_kernel): same mathematics as sph_isoft_kernel, organised so that the
// FP64 tensor pipe, not L2 / shared memory / index arithmetic, is the limit:
//   * persistent CTAs; a CTA owns KC beta planes in mirror pairs {k, F-1-k} (KC = 2: {c, F-1-c};
//     KC = 4: {2c, 2c+1, F-2-2c, F-1-2c}) and keeps its slice of the Wigner table in shared memory for
//     all its pairs (the generic kernel streams 0.7 MB of table per pair from L2).  Only m1 >= 0
//     entries are stored:   d^l_{-m1,m2}(beta_k) = (-1)^{l+m2} d^l_{m1,m2}(beta_{F-1-k})
//     so the mirrored plane supplies the -m1 values.  Entries are ordered by shell
//     lmin = max(m1, m2): level l is the prefix of (l+1)^2 entries.
//   * the packed coefficients of the next pair are staged into shared memory with cp.async while the
//     current pair is transformed;
//   * K5 (Wigner contraction) in registers in the A-fragment layout, stage A (m1 -> alpha) and stage
//     B (m2 -> gamma, half-complex -> real) as DMMA tiles (SymMma), arg-max per orientation.
// ------------------------------------------------------------------------------------------
#ifndef FO_I3_KC
#define FO_I3_KC 2  // beta planes per CTA of sph_isoft3_kernel (2: two CTAs per SM; 4: one)
#endif

struct I2Layout {
  int L, L1, W, F, H, KC, nchunk, NP, RA, RB, RBp, nent, dts, ipk;
  int o_ent[65];                                   // entries below level l: sum_{l' < l} (l'+1)^2
  int o_dts, o3_iks, o3_br, o3_bi, o3_red, total3;  // shared-memory offsets in doubles
  I2Layout() {}
  I2Layout(int L_, int KC_) {
    L = L_;
    KC = KC_;
    L1 = L + 1;
    W = 2 * L + 1;
    F = 2 * L1;
    H = F / 2 + 1;
    nchunk = F / KC;
    NP = L1 * L1;
    RA = KC * L1 * 2;
    RB = F * KC;
    RBp = RB | 1;   // 8-byte stores of consecutive m2 land in distinct banks
    int off = 0;
    for (int l = 0; l <= 64; ++l) {
      o_ent[l] = off;
      if (l <= L) off += (l + 1) * (l + 1);
    }
    nent = off;
    dts = nent * KC;  // doubles of one chunk's table slice: [level][entry][kk]
    ipk = nent * 2;   // double2 elements of the packed coefficients of one pair: [level][entry][+-]
    o_dts = 0;
    o3_iks = o_dts + ((dts + 1) & ~1);
    o3_br = o3_iks + 2 * ipk;
    o3_bi = o3_br + 2 * L1 * RBp;
    o3_red = o3_bi + 2 * L1 * RBp;
    o3_red = (o3_red + 1) & ~1;
    total3 = o3_red + 96;
  }
};

__host__ __device__ __forceinline__ int i2_plane(int F, int KC, int chunk, int kk) {
  // first half of the kk range: planes c KC/2 + kk; second half: their mirrors in reverse order, so that
  // the mirror of kk is KC - 1 - kk
  const int h = KC >> 1;
  return kk < h ? h * chunk + kk : F - KC + kk - h * chunk;
}

// DtP[chunk][level l][pair entry t < (l+1)^2][kk] from the dense table Dt[m2][m1+L][l][k]
__global__ void sph_wigner_pack_kernel(const double* __restrict__ Dt, const __grid_constant__ I2Layout Y,
                                       double* __restrict__ DtP) {
  const int L = Y.L, L1 = Y.L1, W = Y.W, F = Y.F;
  const int total = Y.nchunk * Y.dts;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e / Y.dts;
    int r = e - c * Y.dts;
    int l = 0;
    while (l < L && r >= Y.o_ent[l + 1] * Y.KC) ++l;
    r -= Y.o_ent[l] * Y.KC;
    const int t = r / Y.KC, kk = r - t * Y.KC;
    int s = (int)sqrt((double)t);
    while ((s + 1) * (s + 1) <= t) ++s;
    while (s * s > t) --s;
    const int q = t - s * s;
    const int a = q <= s ? q : s, m2 = q <= s ? s : q - s - 1;
    DtP[e] = Dt[(((size_t)m2 * W + (a + L)) * L1 + l) * F + i2_plane(F, Y.KC, c, kk)];
  }
}

// Ipk[pair][level l][entry t < (l+1)^2][sign +-] from the dense Ihalf[m2][m1+L][l]
__global__ void sph_ipack_kernel(const double2* __restrict__ Ihalf, const __grid_constant__ I2Layout Y,
                                 size_t npairs, double2* __restrict__ Ipk) {
  const int L = Y.L, L1 = Y.L1, W = Y.W;
  const size_t total = npairs * (size_t)Y.ipk;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t p = e / Y.ipk;
    int r = (int)(e - p * Y.ipk);
    int l = 0;
    while (l < L && r >= Y.o_ent[l + 1] * 2) ++l;
    r -= Y.o_ent[l] * 2;
    const int t = r >> 1, sg = r & 1;
    int s = (int)sqrtf((float)t);
    while ((s + 1) * (s + 1) <= t) ++s;
    while (s * s > t) --s;
    const int q = t - s * s;
    const int a = q <= s ? q : s, m2 = q <= s ? s : q - s - 1;
    Ipk[e] = Ihalf[p * (size_t)L1 * W * L1 + ((size_t)m2 * W + (L + (sg ? -a : a))) * L1 + l];
  }
}

struct Iso2Out {
  double* part_val;   // [P][O][nchunk]
  int* part_idx;      // [P][O][nchunk]  flat (a F + k) F + g
  double* grid;       // [P][O][F][F][F] or null
};

// ------------------------------------------------------------------------------------------
// sph_isoft3_kernel<KC, KS, NT, NYQ, WANT_GRID>.  KC = beta planes per CTA (the planes k and F-1-k
// come in mirror pairs), KS = ceil(L/4) k-steps, NT = number of 8-wide output tiles handled by DMMA;
// NYQ: H = 8 NT + 1, the last output (alpha or gamma = F/2) is the alternating sum
// c0 + sum (-1)^m E_m, done on the side.  4 KC warps per CTA.
// The Wigner contraction (K5) is done in registers, directly in the A-fragment layout of stage A
// (the first fast kernel kept E / O arrays in shared memory between K5 and stage A: 70 KB, one more
// barrier, and 97 % of its issue slots were index arithmetic -- profiles/r01_summary.md).
// A warp owns stage-A tile w of both orientations: its 8 rows
// are 4 consecutive (kk, m2) lines x (re | im); lane (g, t) holds A[row g][m1 = 4 ks + t + 1].  The
// lane evaluates S(+-m1) for exactly those entries, the two lanes of one line (g, g ^ 1) splitting the
// l sum by parity -- the even-l / odd-l partial sums are what the two orientations need
// (I_inv^l = (-1)^l I^l) and the re / im component each lane lacks comes from its partner by one
// shuffle.  No AE / AO arrays (70 KB), no barrier between K5 and stage A; the freed shared memory
// holds the Wigner slice again (no L1 misses) next to the cp.async-staged coefficients.
// ------------------------------------------------------------------------------------------
template <int KC, int KS, int NT, bool NYQ, bool WANT_GRID>
__global__ void __launch_bounds__(KC * 128, 4 / KC)
sph_isoft3_kernel(const __grid_constant__ I2Layout Y, const double2* __restrict__ Ipk,
                  const double* __restrict__ DtP, int npairs, int norient, Iso2Out out) {
  extern __shared__ double smj[];
  constexpr int NTHREADS = KC * 128;
  const int L = Y.L, L1 = Y.L1, F = Y.F, H = Y.H;
  const int RA = Y.RA, RB = Y.RB, RBp = Y.RBp;
  double* DtS = smj + Y.o_dts;
  double2* IkS = reinterpret_cast<double2*>(smj + Y.o3_iks);
  double* BR = smj + Y.o3_br;   // [o][m2 = 0..L][RBp]   row = a * KC + kk
  double* BI = smj + Y.o3_bi;
  double* red = smj + Y.o3_red;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int chunk = blockIdx.x % Y.nchunk;
  const int jstart = blockIdx.x / Y.nchunk, jstride = gridDim.x / Y.nchunk;
  const int nvalid = NYQ ? H - 1 : H;  // outputs produced by the DMMA tiles
  for (int e = tid; e < Y.dts; e += NTHREADS) DtS[e] = DtP[(size_t)chunk * Y.dts + e];
  auto stage_coeffs = [&](int pr) {
    const double2* src = Ipk + (size_t)pr * Y.ipk;
    for (int e = tid; e < Y.ipk; e += NTHREADS) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(IkS + e);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (jstart < npairs) stage_coeffs(jstart);
  SymMma<KS, NT> mm;
  mm.init(L, F, nvalid, lane);

  // ---- geometry of this warp's stage-A tile (the same tile index in both orientations)
  constexpr int NW = NTHREADS / 32;
  const bool ta_on = warp * 8 < RA;  // ceil(RA / 8) <= 4 KC tiles: one per warp (L1 <= 16)
  const int part = g & 1;
  // KC L1 lines; with KC = 2 and an even Jmax the last tile is only half full: lane_on masks its rows
  const bool lane_on = ta_on && warp * 4 + (g >> 1) < KC * L1;
  const int line = lane_on ? warp * 4 + (g >> 1) : 0;
  // lines are ordered (m2, kk), kk fastest: the KC lanes of one m2 read the same coefficient / table
  // entries in the K5 loop, so those shared-memory loads are broadcasts (half the wavefronts at KC = 2)
  const int kkA = line % KC, m2A = line / KC;
  const double sm2 = (m2A & 1) ? -1.0 : 1.0;
  // entry ks: a = 4 ks + t4 + 1; entry KS: a = 0 (the row constant c0, evaluated by the t4 == 0 lanes).
  // e_tt: shell-ordered entry index; e_l0: first level >= max(a, m2) with this lane's parity
  int e_tt[KS + 1], e_l0[KS + 1];
#pragma unroll
  for (int ks = 0; ks <= KS; ++ks) {
    const int a = ks < KS ? 4 * ks + t4 + 1 : 0;
    const bool on = lane_on && a <= L && (ks < KS || t4 == 0);
    const int sh = a > m2A ? a : m2A;
    e_tt[ks] = m2A >= a ? m2A * m2A + a : a * a + a + 1 + m2A;
    e_l0[ks] = on ? sh + ((sh ^ part) & 1) : L + 1;  // L + 1: empty level loop
  }
  const int outA = (m2A)*RBp + kkA;  // + (part ? o3_bi : o3_br) + o L1 RBp
  // ---- stage-B geometry (tile = warp + 16 j)
  bool tb_v[2], tb_o[2], tb_lane[2];
  int tb_in[2], tb_base[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int tile = warp + NW * j;
    tb_v[j] = tile * 8 < norient * RB;
    tb_lane[j] = tile * 8 + g < norient * RB;  // the last tile can be partial (RB = F KC, F = 2 mod 4)
    const int r = tb_lane[j] ? tile * 8 + g : 0;
    const int o = r / RB, rowb = r - o * RB;
    const int a = rowb / KC, kk = rowb - a * KC;
    tb_in[j] = o * L1 * RBp + rowb;
    tb_base[j] = (a * F + i2_plane(F, KC, chunk, kk)) * F;
    tb_o[j] = o != 0;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  for (int pair = jstart; pair < npairs; pair += jstride) {
    // ---- K5 in registers + stage A
    if (ta_on) {
      double fe[2][KS], fo[2][KS], c0o[2];  // A fragments E / O per orientation, row constants
      // partial sums over the levels of this lane's parity, all KS + 1 entries advanced together (one
      // running level offset, 4 (KS + 1) independent loads in flight per step):
      // P = sum d(+a) I(+a), Mn = sum d(-a) I(-a)
      double2 Ps[KS + 1], Ms[KS + 1];
#pragma unroll
      for (int ks = 0; ks <= KS; ++ks) Ps[ks] = Ms[ks] = make_double2(0.0, 0.0);
      {
        int base = Y.o_ent[part];  // entries below level lv
        for (int lv = part; lv <= L; lv += 2) {
#pragma unroll
          for (int ks = 0; ks <= KS; ++ks) {
            if (lv >= e_l0[ks]) {
              const int idx = base + e_tt[ks];
              double dp, dm;
              if (KC == 2) {  // the plane and its mirror are one 16-byte entry
                const double2 dd = *reinterpret_cast<const double2*>(DtS + idx * 2);
                dp = kkA ? dd.y : dd.x;
                dm = kkA ? dd.x : dd.y;
              } else {
                dp = DtS[idx * KC + kkA];
                dm = DtS[idx * KC + (KC - 1 - kkA)];
              }
              const double2 cp = IkS[idx * 2], cm = IkS[idx * 2 + 1];
              Ps[ks].x = fma(dp, cp.x, Ps[ks].x);
              Ps[ks].y = fma(dp, cp.y, Ps[ks].y);
              Ms[ks].x = fma(dm, cm.x, Ms[ks].x);
              Ms[ks].y = fma(dm, cm.y, Ms[ks].y);
            }
          }
          base += (lv + 1) * (lv + 1) + (lv + 2) * (lv + 2);
        }
      }
#pragma unroll
      for (int ks = 0; ks <= KS; ++ks) {
        const double2 P = Ps[ks], Mn = Ms[ks];
        // component c = part of the even / odd sums: own one, partner (lane ^ 4) supplies the other
        const double sendP = part ? P.x : P.y, sendM = part ? Mn.x : Mn.y;
        const double recvP = __shfl_xor_sync(0xffffffffu, sendP, 4);
        const double recvM = __shfl_xor_sync(0xffffffffu, sendM, 4);
        const double pe = part ? recvP : P.x, po = part ? P.y : recvP;
        const double me = part ? recvM : Mn.x, mo = part ? Mn.y : recvM;
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const double so = o ? -1.0 : 1.0;
          // S(+a) = sum_l so^l d I ; S(-a) = (-1)^m2 sum_l (-so)^l d_mirror I_-
          const double sp = fma(so, po, pe), sn = sm2 * fma(-so, mo, me);
          if (ks < KS) {
            fe[o][ks] = sp + sn;
            fo[o][ks] = sp - sn;
          } else {
            c0o[o] = __shfl_sync(0xffffffffu, sp, lane & ~3);  // from the t4 == 0 lane of this row
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        if (o >= norient) break;
        double Pq[NT][2], Qq[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          Pq[nt][0] = Pq[nt][1] = c0o[o];
          Qq[nt][0] = Qq[nt][1] = 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            fo_dmma(Pq[nt], fe[o][ks], mm.bc[ks][nt]);
            fo_dmma(Qq[nt], fo[o][ks], mm.bs[ks][nt]);
          }
        double* Bout = smj + (part ? Y.o3_bi : Y.o3_br) + o * L1 * RBp + outA;
        const double sgn = part ? 1.0 : -1.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int d = nt * 8 + t4 * 2 + q;
            const double qx = __shfl_xor_sync(0xffffffffu, Qq[nt][q], 4);
            if (lane_on && (NYQ || d < nvalid)) {
              Bout[d * KC] = fma(sgn, qx, Pq[nt][q]);
              if (d != 0 && (NYQ || 2 * d != F)) Bout[(F - d) * KC] = fma(-sgn, qx, Pq[nt][q]);
            }
          }
        if (NYQ) {  // a = F/2: U = c0 + sum_m (-1)^m E_m  (sin terms vanish); m = 4 ks + t4 + 1
          double acc = 0.0;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) acc += (t4 & 1) ? fe[o][ks] : -fe[o][ks];
          acc += __shfl_xor_sync(0xffffffffu, acc, 1);
          acc += __shfl_xor_sync(0xffffffffu, acc, 2);
          if (t4 == 0 && lane_on) Bout[(F / 2) * KC] = c0o[o] + acc;
        }
      }
    }
    __syncthreads();
    if (pair + jstride < npairs) stage_coeffs(pair + jstride);  // lands during stage B
    // ---- stage B: tiles of 8 rows (o, a, kk): g[gam] = aa - bb, g[F-gam] = aa + bb; arg-max at half
    // scale: accumulators start at v0/2, the grid value is 2 (A -+ B), and the larger of the two outputs
    // of a column is A + |B| -- one DADD + one max per column; only a tile that beats the running
    // maximum is looked at in detail
    double bvh[2] = {-1e300, -1e300};
    int bix[2] = {0x7fffffff, 0x7fffffff};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (!tb_v[j]) continue;
      const double* br = BR + tb_in[j];
      const double* bi1p = BI + tb_in[j] + RBp;
      const int base = tb_base[j];
      const int o = tb_o[j] ? 1 : 0;
      const double v0h = 0.5 * br[0];
      double A[NT][2], Bq[NT][2];
      mm.run(br + RBp - g, bi1p - g, RBp, L, v0h, lane, A, Bq);
      double cmax = -1e300;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int d = nt * 8 + t4 * 2 + q;
          const double c = A[nt][q] + fabs(Bq[nt][q]);
          if (NYQ || d < nvalid) cmax = fmax(cmax, c);
          if (WANT_GRID && tb_lane[j]) {
            double* grow = out.grid + (((size_t)pair * norient + o) * F * F * F + (size_t)base);
            if (NYQ || d < nvalid) {
              grow[d] = 2.0 * (A[nt][q] - Bq[nt][q]);
              if (d != 0 && (NYQ || 2 * d != F)) grow[F - d] = 2.0 * (A[nt][q] + Bq[nt][q]);
            }
          }
        }
      double gny = -1e300;
      if (NYQ) {  // gamma = F/2: g = v0 + 2 sum_m2 (-1)^m2 Vr_m2
        double acc = 0.0;
        for (int m = 1 + t4; m <= L; m += 4) {
          const double ev = br[(size_t)m * RBp];
          acc += (m & 1) ? -ev : ev;
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (t4 == 0) {
          gny = v0h + acc;
          cmax = fmax(cmax, gny);
          if (WANT_GRID && tb_lane[j])
            out.grid[((size_t)pair * norient + o) * F * F * F + (size_t)base + F / 2] = 2.0 * gny;
        }
      }
      const double cur = o ? bvh[1] : bvh[0];
      if (tb_lane[j] && cmax >= cur) {
        double tbv = cur;
        int tbi = o ? bix[1] : bix[0];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int d = nt * 8 + t4 * 2 + q;
            const bool in = NYQ || d < nvalid;
            const double g1 = in ? A[nt][q] - Bq[nt][q] : -1e300;
            const double g2 = (in && d != 0 && (NYQ || 2 * d != F)) ? A[nt][q] + Bq[nt][q] : -1e300;
            if (g1 > tbv || (g1 == tbv && base + d < tbi)) { tbv = g1; tbi = base + d; }
            if (g2 > tbv || (g2 == tbv && base + F - d < tbi)) { tbv = g2; tbi = base + F - d; }
          }
        if (NYQ && (gny > tbv || (gny == tbv && base + F / 2 < tbi))) { tbv = gny; tbi = base + F / 2; }
        if (o) { bvh[1] = tbv; bix[1] = tbi; } else { bvh[0] = tbv; bix[0] = tbi; }
      }
    }
    int* redi = reinterpret_cast<int*>(red + 48);
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      double m = bvh[o];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
      const int cand = (bvh[o] == m) ? bix[o] : 0x7fffffff;
      const int imin = __reduce_min_sync(0xffffffffu, cand);
      if (lane == 0) {
        red[o * 16 + warp] = m;
        redi[o * 16 + warp] = imin;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // next pair's coefficients: visible after the barrier
    __syncthreads();
    if (tid < norient) {
      double v = red[tid * 16];
      int i = redi[tid * 16];
      for (int w = 1; w < NW; ++w) {
        const double ov = red[tid * 16 + w];
        const int oi = redi[tid * 16 + w];
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
      }
      out.part_val[((size_t)pair * norient + tid) * Y.nchunk + chunk] = 2.0 * v;
      out.part_idx[((size_t)pair * norient + tid) * Y.nchunk + chunk] = i;
    }
    // no barrier: red is rewritten only after the next pair's stage-A barrier
  }
}

// ------------------------------------------------------------------------------------------
// sph_isoft4_kernel<KC, KS, NT, NYQ, WANT_GRID>: sph_isoft3_kernel with stages A and B chained in registers
// (the scheme of per_xf6_kernel, fo_periodic.cu).  Per pair a CTA (KC beta planes in mirror pairs; 4 KC warps):
//   phase 1  K5 (Wigner contraction) in registers; E = S(+a) + S(-a) and O'' = i (S(+a) - S(-a)) of both
//            orientations go into the plane's S block in shared memory, rows [c0 | E_a | O''_a], columns
//            (m2, re | im) -- the layout of a YIN slab of per_xf6_kernel.
//            KC = 4 (odd Jmax; one CTA of 16 warps per SM): one lane per (|m1| = a, m2, level parity) in shell
//            order -- a warp's lanes have the same number of levels to within one -- carrying all four planes:
//            the table entry of a level is one 32-byte load (two planes and their mirrors, which supply -a), the
//            two coefficients I(+a), I(-a) another, for 16 DFMA.  sph_isoft3's K5 (one lane per (plane, m2, a
//            mod 4)) read 12 bytes of shared memory per DFMA: ~60 k wavefronts per pair and SM, more cycles than
//            all the DMMAs; this one reads 4.
//            KC = 2 (even Jmax; two CTAs of 8 warps per SM): the same with two planes per lane;
//   phase 2  one work item (plane, orientation, alpha row tile) per warp: stage A transposed (twiddles = A
//            operand, S block = B operand): V[alpha] = c0 + sum_a cos E_a + sin O''_a as ONE DMMA chain per column
//            tile (P, then V = P + sin O''), V[F - alpha] = 2 P - V; a lane's C fragment of column tile ct is the A
//            fragment (row alpha, k = m2 slot) of stage B's k-step ks = ct, so stage B consumes the accumulators
//            of stage A directly: no BR / BI arrays (sph_isoft3: 2 x 33 KB written and re-read per pair), no
//            shuffles, no scalar FP64 between the stages.  m2 = 0 rides in slot L + 1 with weight 1/2 (half
//            scale).  NYQ (F/2 = 8 NT, Jmax = 7, 15): the row alpha = F/2 is one extra DMMA chain with the
//            twiddle (-1)^a in row 0 and takes the free slot of the mirror tile (the mirror of alpha = 0); the
//            column gamma = F/2 is a one-column tile with the twiddle (-1)^m2.
//   arg-max  the larger output of a column is A + |B|: bounded from above for the whole item by one addition of
//            integer-found maxima; only items that may beat the running maximum are examined exactly.
// ------------------------------------------------------------------------------------------
// signed-comparable integer key of a double (monotone: a < b <=> key(a) < key(b); -0 < +0) and its inverse
__device__ __forceinline__ long long fo_dkey(double x) {
  const long long b = __double_as_longlong(x);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double fo_dunkey(long long k) {
  return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}

struct I4Layout {
  int SPc, sblk, o4_s, o4_red, total4;  // S block: [kk][o][2 L + 1 rows][SPc]
  I4Layout() {}
  I4Layout(const I2Layout& Y) {
    const int cols = 2 * Y.L1;
    SPc = ((cols + 11) / 16) * 16 + 4;  // == 4 (mod 16): conflict-free B fragments (rows t4 32 bytes apart mod 128)
    sblk = (2 * Y.L + 1) * SPc;
    o4_s = Y.o3_br;                     // in place of BR / BI
    o4_red = (o4_s + 2 * Y.KC * sblk + 1) & ~1;
    total4 = o4_red + 96;
  }
};

template <int KC, int KS, int NT, bool NYQ, bool WANT_GRID>
__global__ void __launch_bounds__(KC * 128, 4 / KC)
sph_isoft4_kernel(const __grid_constant__ I2Layout Y, const __grid_constant__ I4Layout Z,
                  const double2* __restrict__ Ipk, const double* __restrict__ DtP, int npairs, int norient,
                  Iso2Out out) {
  extern __shared__ double smk[];
  constexpr int NTHREADS = KC * 128, NW = KC * 4;
  const int L = Y.L, L1 = Y.L1, F = Y.F;
  const int SPc = Z.SPc;
  double* DtS = smk + Y.o_dts;
  double2* IkS = reinterpret_cast<double2*>(smk + Y.o3_iks);
  double* SB = smk + Z.o4_s;  // [kk][o] blocks
  double* red = smk + Z.o4_red;
  int* sbest = reinterpret_cast<int*>(red + 80);  // [o]: high word (signed key) of a lower bound of the maximum
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int chunk = blockIdx.x % Y.nchunk;
  const int jstart = blockIdx.x / Y.nchunk, jstride = gridDim.x / Y.nchunk;
  const int nvalid = NYQ ? F / 2 : F / 2 + 1;  // alpha / gamma values covered by the 8-wide tiles
  for (int e = tid; e < Y.dts; e += NTHREADS) DtS[e] = DtP[(size_t)chunk * Y.dts + e];
  // the packed coefficients of a pair (48 KB at Jmax = 15) arrive by one bulk asynchronous copy (cp.async.bulk,
  // SASS UBLKCP) completing on an mbarrier: issued by one thread, no LDGSTS slots, no issue slots of the others
  uint64_t* cbar = reinterpret_cast<uint64_t*>(red + 88);
  const unsigned ipk_bytes = (unsigned)Y.ipk * 16u;
  auto stage_coeffs = [&](int pr) {
    if (tid == 0) {
      fo_fence_proxy_async();
      fo_mbar_arrive_expect_tx(cbar, ipk_bytes);
      fo_bulk_g2s(IkS, Ipk + (size_t)pr * Y.ipk, ipk_bytes, cbar);
    }
  };
  if (tid == 0) {
    fo_mbar_init(cbar, 1);
    fo_mbar_fence_init();
  }
  int cph = 0;
  if (jstart < npairs) stage_coeffs(jstart);
  SymMma<KS, NT> mm;
  mm.init(L, F, nvalid, lane);
  // stage-B cos fragment of the k-step that carries m2 = 0 (slot L + 1): weight 1/2 at half scale
  const int ks0 = L >> 2, t0 = L & 3;
  double bcz[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    bcz[nt] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      if (ks == ks0) bcz[nt] = (t4 == t0 && nt * 8 + g < nvalid) ? 0.5 : mm.bc[ks][nt];
  }
  // NYQ: A fragment of the row alpha = F/2 ((-1)^a in row 0), B fragment of the column gamma = F/2 ((-1)^m2 in
  // column 0, 1/2 for m2 = 0)
  double nyqa[KS], nyqz[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int a = 4 * ks + t4 + 1;
    nyqa[ks] = (NYQ && g == 0 && a <= L) ? ((a & 1) ? -1.0 : 1.0) : 0.0;
    nyqz[ks] = (NYQ && g == 0) ? (a <= L ? ((a & 1) ? -1.0 : 1.0) : (a == L + 1 ? 0.5 : 0.0)) : 0.0;
  }

  // ---- K5 geometry: task = (shell-ordered entry t = (a, m2), level parity); KC = 4: one task per thread, hoisted
  auto k5_geom = [&](int task, int& t, int& sq, int& a, int& m2) {
    t = task >> 1;
    sq = (int)sqrtf((float)t);
    while ((sq + 1) * (sq + 1) <= t) ++sq;
    while (sq * sq > t) --sq;
    const int q = t - sq * sq;
    a = q <= sq ? q : sq;
    m2 = q <= sq ? sq : q - sq - 1;
  };
  int k5t, k5s, k5a, k5m2;
  k5_geom(tid, k5t, k5s, k5a, k5m2);
  const int k5par = tid & 1;
  const bool k5on = k5t < L1 * L1;
  const int k5l0 = k5on ? k5s + ((k5s ^ k5par) & 1) : L + 1;  // first level >= max(a, m2) of this lane's parity
  // ---- phase-2 geometry: item (kk, o, mt) of this warp; B-fragment rows / columns of the lane
  int ycol[KS], ycol0[KS], yrow[KS];
#pragma unroll
  for (int ct = 0; ct < KS; ++ct) {
    const int sb = 4 * ct + (g >> 1) + 1, sc = 4 * ct + t4 + 1;
    ycol[ct] = 2 * (sb <= L ? sb : 0) + (g & 1);
    ycol0[ct] = 2 * (sc <= L ? sc : 0);
    const int j = 4 * ct + t4 + 1;
    yrow[ct] = (j <= L ? j : L) * SPc;
  }
  const int nq = KC * norient;  // (kk, o) combinations; items = nq NT
  __syncthreads();  // the initialised mbarrier is visible to every waiter

  for (int pair = jstart; pair < npairs; pair += jstride) {
    if (tid < 2) sbest[tid] = (int)0x80000000;
    fo_mbar_wait(cbar, cph);  // this pair's coefficients have landed
    cph ^= 1;
    // ---- phase 1: K5 in registers -> S blocks.  Task = (shell-ordered entry t = (a, m2), level parity): the
    // lanes of a warp have the same number of levels to within one
    auto k5_task = [&](int t, int a, int m2, int l0, bool on) {
      // partial sums over the levels of this lane's parity for the KC planes: P = sum d(+a) I(+a),
      // Mn = sum d_mirror I(-a)  (d^l_{-a,m2}(beta_kk) = (-1)^(l + m2) d^l_{a,m2}(beta_{KC - 1 - kk}))
      double2 Pp[KC], Mp[KC];
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) Pp[kk] = Mp[kk] = make_double2(0.0, 0.0);
      for (int lv = l0; lv <= L; lv += 2) {
        const int idx = Y.o_ent[lv] + t;
        double dd[KC];
#pragma unroll
        for (int kk = 0; kk < KC; kk += 2) {
          const double2 d2 = *reinterpret_cast<const double2*>(DtS + idx * KC + kk);
          dd[kk] = d2.x;
          dd[kk + 1] = d2.y;
        }
        const double2 cp = IkS[idx * 2], cm = IkS[idx * 2 + 1];
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
          Pp[kk].x = fma(dd[kk], cp.x, Pp[kk].x);
          Pp[kk].y = fma(dd[kk], cp.y, Pp[kk].y);
          Mp[kk].x = fma(dd[KC - 1 - kk], cm.x, Mp[kk].x);
          Mp[kk].y = fma(dd[KC - 1 - kk], cm.y, Mp[kk].y);
        }
      }
      // the lane keeps component c = parity of every sum and gets that component of the other parity's sums
      // from its partner (lane ^ 1): then it holds the even (e) and odd (o) level sums of component c
      const double sm2q = (m2 & 1) ? -1.0 : 1.0;
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const double sendP = k5par ? Pp[kk].x : Pp[kk].y, sendM = k5par ? Mp[kk].x : Mp[kk].y;
        const double recvP = __shfl_xor_sync(0xffffffffu, sendP, 1);
        const double recvM = __shfl_xor_sync(0xffffffffu, sendM, 1);
        const double pe = k5par ? recvP : Pp[kk].x, po = k5par ? Pp[kk].y : recvP;
        const double me = k5par ? recvM : Mp[kk].x, mo = k5par ? Mp[kk].y : recvM;
        if (on) {
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            if (o >= norient) break;
            const double so = o ? -1.0 : 1.0;
            // S(+a) = sum_l so^l d I ; S(-a) = (-1)^m2 sum_l (-so)^l d_mirror I_-
            const double sp = fma(so, po, pe), sn = sm2q * fma(-so, mo, me);
            double* blk = SB + (kk * 2 + o) * Z.sblk;
            if (a > 0) {
              blk[a * SPc + 2 * m2 + k5par] = sp + sn;                                     // E
              blk[(L + a) * SPc + 2 * m2 + (k5par ^ 1)] = k5par ? (sn - sp) : (sp - sn);   // O'' = i O
            } else {
              blk[2 * m2 + k5par] = sp;  // c0 = S(0, m2)
            }
          }
        }
      }
    };
    k5_task(k5t, k5a, k5m2, k5l0, k5on);
    if (KC == 2) {  // 512 tasks on 256 threads: the second one (warp-uniform: the partner shuffles need every lane)
      int t2, s2, a2, m22;
      k5_geom(tid + NTHREADS, t2, s2, a2, m22);
      const bool on2 = t2 < L1 * L1;
      k5_task(t2, a2, m22, on2 ? s2 + ((s2 ^ k5par) & 1) : L + 1, on2);
    }
    __syncthreads();
    if (pair + jstride < npairs) stage_coeffs(pair + jstride);  // lands during phase 2
    // ---- phase 2: item = (kk, o, mt): stage A -> stage B in registers
    // running maximum per orientation as a signed-comparable 64-bit key (monotone in the value)
    long long bkey[2] = {fo_dkey(-1e300), fo_dkey(-1e300)};
    int bix[2] = {0x7fffffff, 0x7fffffff};
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
      const int q4 = warp - mt * 2 * KC;  // warps 2 KC mt .. 2 KC mt + nq - 1 take row tile mt
      if (q4 < 0 || q4 >= nq) continue;
      const int kk = q4 % KC, o = q4 / KC;
      const double* S = SB + (kk * 2 + o) * Z.sblk;
      const int plane = i2_plane(F, KC, chunk, kk);
      const int al0 = 8 * mt + g;
      const bool valid0 = al0 < nvalid;
      // mirror tile: alpha = F - al0; NYQ: the free slot (mirror of alpha = 0) carries alpha = F/2
      const bool nyqrow = NYQ && mt == 0 && g == 0;
      const bool valid1 = nyqrow || (valid0 && al0 != 0 && 2 * al0 != F);
      const int al1 = nyqrow ? F / 2 : F - al0;
      double A[2][NT][2], B[2][NT][2], An[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        An[h][0] = An[h][1] = 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) A[h][nt][0] = A[h][nt][1] = B[h][nt][0] = B[h][nt][1] = 0.0;
      }
#pragma unroll
      for (int c2 = 0; c2 < KS; c2 += 2) {  // two column tiles (m2 slots 4 ct + 1 .. 4 ct + 4) at a time
        double P[2][2], V[2][2], Vn[2][2];
        {
          double be[KS][2], bo[KS][2];
#pragma unroll
          for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int ct = (c2 + c < KS) ? c2 + c : KS - 1;
              be[ks][c] = S[yrow[ks] + ycol[ct]];
              bo[ks][c] = S[yrow[ks] + L * SPc + ycol[ct]];
            }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int ct = (c2 + c < KS) ? c2 + c : KS - 1;
            const double2 z = *reinterpret_cast<const double2*>(S + ycol0[ct]);
            fo_dmma3(P[c], mm.bc[0][mt], be[0][c], z.x, z.y);
            if (NYQ && mt == 0) fo_dmma3(Vn[c], nyqa[0], be[0][c], z.x, z.y);
          }
#pragma unroll
          for (int ks = 1; ks < KS; ++ks)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              fo_dmma(P[c], mm.bc[ks][mt], be[ks][c]);
              if (NYQ && mt == 0) fo_dmma(Vn[c], nyqa[ks], be[ks][c]);
            }
#pragma unroll
          for (int c = 0; c < 2; ++c) fo_dmma3(V[c], mm.bs[0][mt], bo[0][c], P[c][0], P[c][1]);
#pragma unroll
          for (int ks = 1; ks < KS; ++ks)
#pragma unroll
            for (int c = 0; c < 2; ++c) fo_dmma(V[c], mm.bs[ks][mt], bo[ks][c]);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c2 + c >= KS) break;
          const int ks = c2 + c;  // stage-B k-step fed by this column tile
          double wr = fma(2.0, P[c][0], -V[c][0]), wi = fma(2.0, P[c][1], -V[c][1]);  // V[F - alpha]
          if (NYQ && mt == 0 && g == 0) {
            wr = Vn[c][0];
            wi = Vn[c][1];
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double zc = (ks == ks0) ? bcz[nt] : mm.bc[ks][nt];
            fo_dmma(A[0][nt], V[c][0], zc);
            fo_dmma(B[0][nt], V[c][1], mm.bs[ks][nt]);
            fo_dmma(A[1][nt], wr, zc);
            fo_dmma(B[1][nt], wi, mm.bs[ks][nt]);
          }
          if (NYQ) {
            fo_dmma(An[0], V[c][0], nyqz[ks]);
            fo_dmma(An[1], wr, nyqz[ks]);
          }
        }
      }
      // outputs (half scale): g[gam] = A - B, g[F - gam] = A + B, gam = 8 nt + 2 t4 + q; NYQ: g[F/2] = An (t4 = 0, q = 0)
      if (WANT_GRID) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const bool valid = h == 0 ? valid0 : valid1;
          const int al = h == 0 ? al0 : al1;
          double* grow = out.grid + (((size_t)pair * norient + o) * F * F * F + (size_t)(al * F + plane) * F);
          if (valid) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int d = nt * 8 + t4 * 2 + q;
                if (d < nvalid) {
                  grow[d] = 2.0 * (A[h][nt][q] - B[h][nt][q]);
                  if (d != 0 && 2 * d != F) grow[F - d] = 2.0 * (A[h][nt][q] + B[h][nt][q]);
                }
              }
            if (NYQ && t4 == 0) grow[F / 2] = 2.0 * An[h][0];
          }
        }
      }
      // Filter off the FP64 pipe: signed keys of the high words (monotone in the value), largest A (and An) and
      // largest |B| of the lane's accumulators; one addition of their upper bounds bounds every A + |B|.
      int ka = (int)0x80000000;
      unsigned kb = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int ha = __double2hiint(A[h][nt][q]);
            ka = max(ka, ha ^ ((ha >> 31) & 0x7fffffff));
            kb = max(kb, (unsigned)__double2hiint(B[h][nt][q]) << 1);
          }
        if (NYQ) {
          const int ha = __double2hiint(An[h][0]);
          ka = max(ka, ha ^ ((ha >> 31) & 0x7fffffff));
        }
      }
      // upper bounds as doubles: A <= ua (0 for a negative maximum), |B| <= ub
      const double ua = ka >= 0 ? __hiloint2double(min(ka + 1, 0x7ff00000), 0) : 0.0;
      const double ub = __hiloint2double(min((int)(kb >> 1) + 1, 0x7ff00000), 0);
      const int hs = __double2hiint(ua + ub);
      const int thr = max(sbest[o], (int)(bkey[o] >> 32));
      if ((hs ^ ((hs >> 31) & 0x7fffffff)) >= thr) {
        // the larger output of a column is A + |B| (bit for bit A - B or A + B): at gamma = d when B < 0 or
        // B = 0 (a tie of the two outputs: the smaller index), at F - d when B > 0; compared as integer keys
        long long tk = bkey[o];
        int tbi = bix[o];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const bool valid = h == 0 ? valid0 : valid1;
          const int al = h == 0 ? al0 : al1;
          const int base = (al * F + plane) * F;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int d = nt * 8 + t4 * 2 + q;
              const long long ck = fo_dkey(A[h][nt][q] + fabs(B[h][nt][q]));
              const int idx = base + (__double_as_longlong(B[h][nt][q]) > 0 ? F - d : d);  // > 0 as integers: B > +0
              if (valid && d < nvalid && (ck > tk || (ck == tk && idx < tbi))) {
                tk = ck;
                tbi = idx;
              }
            }
          if (NYQ && valid && t4 == 0) {
            const long long ck = fo_dkey(An[h][0]);
            if (ck > tk || (ck == tk && base + F / 2 < tbi)) {
              tk = ck;
              tbi = base + F / 2;
            }
          }
        }
        if (tk != bkey[o] || tbi != bix[o]) {
          bkey[o] = tk;
          bix[o] = tbi;
          atomicMax(&sbest[o], (int)(tk >> 32));
        }
      }
    }
    int* redi = reinterpret_cast<int*>(red + 48);
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      // warp maximum of the keys: high words by REDUX, then the low words among the lanes that hold it
      const int hi = (int)(bkey[o] >> 32);
      const int mhi = __reduce_max_sync(0xffffffffu, hi);
      const unsigned lo = hi == mhi ? (unsigned)bkey[o] : 0u;
      const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
      const bool top = hi == mhi && (unsigned)bkey[o] == mlo;
      const int imin = __reduce_min_sync(0xffffffffu, top ? bix[o] : 0x7fffffff);
      if (lane == 0) {
        red[o * 16 + warp] = fo_dunkey(((long long)mhi << 32) | (long long)mlo);
        redi[o * 16 + warp] = imin;
      }
    }
    __syncthreads();
    if (tid < norient) {
      double v = red[tid * 16];
      int i = redi[tid * 16];
      for (int w = 1; w < NW; ++w) {
        const double ov = red[tid * 16 + w];
        const int oi = redi[tid * 16 + w];
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
      }
      out.part_val[((size_t)pair * norient + tid) * Y.nchunk + chunk] = 2.0 * v;
      out.part_idx[((size_t)pair * norient + tid) * Y.nchunk + chunk] = i;
    }
    // no barrier: red / sbest are rewritten only after the next pair's phase-1 barrier... sbest is reset at the
    // top of the loop by tid < 2, read in phase 2 (after the barrier)
  }
}

// ------------------------------------------------------------------------------------------
// sph_isoft5_kernel<WANT_GRID> (Jmax = 15: F = 32): the two inverse transforms of a beta plane as radix FFTs on
// the FP64 vector pipe instead of DFT matrices on the tensor pipe.  On B200 the two pipes have the same FP64 peak
// (DFMA 34 TFLOP/s, DMMA.8x8x4 37 TFLOP/s, bench.py `fp64_peaks`) and share the same units; a 32-point FFT needs a
// tenth of the operations of the 32 x 31 DFT product, so the transform that takes 1280 DMMA (41 k cycles of the
// pipe per pair and SM for all 16 plane quartets) becomes ~123 FP64 instructions per thread and transform step.
// Same frame as sph_isoft4_kernel<4, ...>: persistent CTA of 512 threads per (chunk of four planes), Wigner slice
// resident, coefficients by bulk copy, phase 1 = K5 (one lane per (a, m2, level parity) carrying four planes),
// which here stores S(+a, m2) and S(-a, m2) themselves: block (kk, o) = [row m1 mod 32][m2 = 0..15] complex.
//   step A  per column m2:  V[al][m2] = sum_m1 S(m1, m2) e^{+2 pi i m1 al / 32}, in place;
//   step B  per pair of rows (al, al + 16):  Z(m2) = Vh_al(m2) + i Vh_al'(m2) over m2 = -15 .. 15 with
//           Vh(-m2) = conj V(m2); one complex transform yields g(al, .) in its real and g(al', .) in its
//           imaginary parts (the grid is real);
// each 32-point transform by FOUR lanes: lane n1 takes the inputs 4 n2 + n1, an 8-point DFT in registers, the
// twiddles e^{2 pi i n1 k2 / 32}, a 4 x 8 -> 8 x 4 exchange by two rounds of xor shuffles, two 4-point DFTs: lane t
// ends with the outputs k2 + 8 k1, k2 = 2 t, 2 t + 1.  All 512 threads have a task in both steps (8 blocks x 16
// transforms x 4 lanes); one CTA barrier between K5, step A and step B.  The arg-max runs on integer keys.
// ------------------------------------------------------------------------------------------
struct I5Layout {
  int SP, sblk, o5_s, o5_red, total5;
  I5Layout() {}
  I5Layout(const I2Layout& Y) {
    SP = 36;  // doubles per row: 16-byte accesses of the four lanes of a transform (rows n1 apart) and of two
              // neighbouring columns fall into eight distinct 16-byte bank groups
    sblk = 32 * SP;
    o5_s = Y.o3_br;
    o5_red = (o5_s + 2 * Y.KC * sblk + 1) & ~1;
    total5 = o5_red + 96;
  }
};

__device__ __forceinline__ double2 fo_cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 fo_csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 fo_cmul(double2 a, double2 t) {
  return make_double2(fma(a.x, t.x, -a.y * t.y), fma(a.x, t.y, a.y * t.x));
}
__device__ __forceinline__ double2 fo_cshx(double2 v, int m) {
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
// X[k] = sum_n x[n] i^{n k} (sign +), in place
__device__ __forceinline__ void fo_fft4(double2& a, double2& b, double2& c, double2& d) {
  const double2 s0 = fo_cadd(a, c), d0 = fo_csub(a, c), s1 = fo_cadd(b, d), d1 = fo_csub(b, d);
  a = fo_cadd(s0, s1);
  c = fo_csub(s0, s1);
  b = make_double2(d0.x - d1.y, d0.y + d1.x);
  d = make_double2(d0.x + d1.y, d0.y - d1.x);
}
// sign flips on the integer pipe: m = 0 or 0x80000000
__device__ __forceinline__ double fo_flip(double v, unsigned m) {
  return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v));
}
__device__ __forceinline__ double2 fo_cflip(double2 v, unsigned m) { return make_double2(fo_flip(v.x, m), fo_flip(v.y, m)); }

// One 32-point transform X[k] = sum_n x[n] e^{+2 pi i n k / 32} on the four lanes n1 = lane & 3 = 2 b1 + b0 of a
// quad; lane n1 brings x[n2] = input 4 n2 + n1 and ends with out[jj][k1] = X[(2 n1 + jj) + 8 k1].
//  * 8-point DFT of the lane's inputs, with its outputs in the order z[k] = y[k ^ r], r = 2 b0 + 4 b1, so that the
//    exchange below keeps and sends the same registers on every lane: negating the odd inputs swaps y[k] and
//    y[k + 4], negating the inputs 2, 3, 6, 7 swaps y[k] and y[k ^ 2] up to the constants (1, w, i, w^3) <->
//    (i, w^3, 1, w), w = e^{i pi / 4}, that multiply the odd half.  The negations are xors of the sign bits.
//  * twiddles e^{2 pi i n1 k2 / 32} (per-lane table in the same order), exchange by xor shuffles over 2 and 1
//    (the lane keeps z[0], z[1] of itself and of lane n1 ^ 2 and receives the same from n1 ^ 1, n1 ^ 3),
//  * 4-point DFT F over the sources (n1, n1 ^ 1, n1 ^ 2, n1 ^ 3): X[k1] = (-1)^{b1 k1} F[k1] for b0 = 0 and
//    (-1)^{b1 k1} i^{k1} F[-k1] for b0 = 1.
struct FftLane {
  bool b0;
  unsigned mb0, mb1;   // sign masks of b0, b1
  double a1;           // b0 ? -1/sqrt 2 : 1/sqrt 2
  double2 tw[8];       // tw[k] = e^{2 pi i n1 (k ^ r) / 32}
  __device__ __forceinline__ void init(int n1) {
    b0 = n1 & 1;
    mb0 = (unsigned)(n1 & 1) << 31;
    mb1 = (unsigned)((n1 >> 1) & 1) << 31;
    a1 = b0 ? -0.70710678118654752440084436210485 : 0.70710678118654752440084436210485;
    const int r = 2 * (n1 & 1) + 4 * ((n1 >> 1) & 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double sn, cs;
      sincospi((double)(n1 * (k ^ r)) / 16.0, &sn, &cs);
      tw[k] = make_double2(cs, sn);
    }
  }
  __device__ __forceinline__ void run(double2 (&x)[8], double2 (&out)[2][4]) const {
    const double r = 0.70710678118654752440084436210485;
    x[1] = fo_cflip(x[1], mb1);
    x[2] = fo_cflip(x[2], mb0);
    x[3] = fo_cflip(x[3], mb0 ^ mb1);
    x[5] = fo_cflip(x[5], mb1);
    x[6] = fo_cflip(x[6], mb0);
    x[7] = fo_cflip(x[7], mb0 ^ mb1);
    fo_fft4(x[0], x[2], x[4], x[6]);
    fo_fft4(x[1], x[3], x[5], x[7]);
    // odd half times (1, w, i, w^3) or (i, w^3, 1, w)
    const double2 o0 = make_double2(b0 ? -x[1].y : x[1].x, b0 ? x[1].x : x[1].y);
    const double2 o2 = make_double2(b0 ? x[5].x : -x[5].y, b0 ? x[5].y : x[5].x);
    const double2 o1 = make_double2(fma(a1, x[3].x, -r * x[3].y), fma(r, x[3].x, a1 * x[3].y));
    const double2 o3 = make_double2(fma(-a1, x[7].x, -r * x[7].y), fma(r, x[7].x, -a1 * x[7].y));
    double2 z[8];
    z[0] = fo_cmul(fo_cadd(x[0], o0), tw[0]);
    z[4] = fo_cmul(fo_csub(x[0], o0), tw[4]);
    z[1] = fo_cmul(fo_cadd(x[2], o1), tw[1]);
    z[5] = fo_cmul(fo_csub(x[2], o1), tw[5]);
    z[2] = fo_cmul(fo_cadd(x[4], o2), tw[2]);
    z[6] = fo_cmul(fo_csub(x[4], o2), tw[6]);
    z[3] = fo_cmul(fo_cadd(x[6], o3), tw[3]);
    z[7] = fo_cmul(fo_csub(x[6], o3), tw[7]);
    double2 R[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) R[j] = fo_cshx(z[j + 4], 2);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      double2 f0 = z[jj], f1 = fo_cshx(z[jj + 2], 1), f2 = R[jj], f3 = fo_cshx(R[jj + 2], 1);
      fo_fft4(f0, f1, f2, f3);
      out[jj][0] = f0;
      out[jj][1] = make_double2(fo_flip(b0 ? f3.y : f1.x, mb1 ^ mb0), fo_flip(b0 ? f3.x : f1.y, mb1));
      out[jj][2] = fo_cflip(f2, mb0);
      out[jj][3] = make_double2(fo_flip(b0 ? f1.y : f3.x, mb1), fo_flip(b0 ? f1.x : f3.y, mb1 ^ mb0));
    }
  }
};

template <bool WANT_GRID>
__global__ void __launch_bounds__(512, 1)
sph_isoft5_kernel(const __grid_constant__ I2Layout Y, const __grid_constant__ I5Layout Z,
                  const double2* __restrict__ Ipk, const double* __restrict__ DtP, int npairs, int norient,
                  Iso2Out out) {
  extern __shared__ double smk[];
  constexpr int KC = 4, NTHREADS = 512, NW = 16, F = 32;
  const int L = Y.L, L1 = Y.L1, SP = Z.SP;
  double* DtS = smk + Y.o_dts;
  double2* IkS = reinterpret_cast<double2*>(smk + Y.o3_iks);
  double* SB = smk + Z.o5_s;  // [kk][o] blocks
  double* red = smk + Z.o5_red;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x % Y.nchunk;
  const int jstart = blockIdx.x / Y.nchunk, jstride = gridDim.x / Y.nchunk;
  for (int e = tid; e < Y.dts; e += NTHREADS) DtS[e] = DtP[(size_t)chunk * Y.dts + e];
  uint64_t* cbar = reinterpret_cast<uint64_t*>(red + 88);
  const unsigned ipk_bytes = (unsigned)Y.ipk * 16u;
  auto stage_coeffs = [&](int pr) {
    if (tid == 0) {
      fo_fence_proxy_async();
      fo_mbar_arrive_expect_tx(cbar, ipk_bytes);
      fo_bulk_g2s(IkS, Ipk + (size_t)pr * Y.ipk, ipk_bytes, cbar);
    }
  };
  if (tid == 0) {
    fo_mbar_init(cbar, 1);
    fo_mbar_fence_init();
  }
  int cph = 0;
  if (jstart < npairs) stage_coeffs(jstart);
  // ---- K5 geometry (as sph_isoft4_kernel<4, ...>): task = (shell-ordered entry t = (a, m2), level parity)
  // The shell-ordered groups of 16 entries take (8, 6, 6, 5, 4, 4, 4, 3, 3, 2, 2, 2, 2, 1, 1, 1) level trips; they are
  // dealt to the warps so that the four warps of an SM sub-partition (warp & 3) get 14, 14, 14 and 12 of them
  // (in warp order: 17, 13, 13, 11; tests/test_lane_emulation.py)
  const int k5grp = (int)((0xfdce98ba65473210ull >> (4 * warp)) & 15);
  int k5t = (k5grp * 32 + lane) >> 1, k5s, k5a, k5m2;
  {
    k5s = (int)sqrtf((float)k5t);
    while ((k5s + 1) * (k5s + 1) <= k5t) ++k5s;
    while (k5s * k5s > k5t) --k5s;
    const int q = k5t - k5s * k5s;
    k5a = q <= k5s ? q : k5s;
    k5m2 = q <= k5s ? k5s : q - k5s - 1;
  }
  const int k5par = tid & 1;
  const bool k5on = k5t < L1 * L1;
  const int k5l0 = k5on ? k5s + ((k5s ^ k5par) & 1) : L + 1;
  // ---- transform geometry: block (kk, o) = tid >> 6, transform f = (tid >> 2) & 15, lane n1 = tid & 3
  const int blk = tid >> 6, kk = blk >> 1, o = blk & 1;
  const int f = (tid >> 2) & 15, n1 = tid & 3;
  FftLane fl;
  fl.init(n1);
  double* S = SB + blk * Z.sblk;
  const int plane = i2_plane(F, KC, chunk, kk);
  // step B: rows al = (f with bits 0 and 1 swapped) and al + 16 -- two neighbouring transforms are two rows apart
  const int alB = (f & ~3) | ((f & 1) << 1) | ((f >> 1) & 1);
  const bool active = o < norient;
  __syncthreads();  // the initialised mbarrier is visible to every waiter

  // combine the 16 warp results of pair pr (buffer pb) and write the chunk's partial maximum
  auto flush = [&](int pr, int pb) {
    if (pr >= 0 && tid < norient) {
      const double* rv = red + 16 * pb;
      const int* ri = reinterpret_cast<const int*>(red + 48 + 8 * pb);
      double bv = -1e300;
      int bi = 0x7fffffff;
      for (int w = 0; w < NW; ++w) {
        if (((w >> 1) & 1) != tid) continue;
        const double ov = rv[w];
        const int oi = ri[w];
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      out.part_val[((size_t)pr * norient + tid) * Y.nchunk + chunk] = bv;
      out.part_idx[((size_t)pr * norient + tid) * Y.nchunk + chunk] = bi;
    }
  };
  int prev = -1, par = 0;
  for (int pair = jstart; pair < npairs; pair += jstride) {
    fo_mbar_wait(cbar, cph);  // this pair's coefficients have landed
    cph ^= 1;
    // ---- phase 1: K5 in registers -> S blocks
    {
      double2 Pp[KC], Mp[KC];
#pragma unroll
      for (int q = 0; q < KC; ++q) Pp[q] = Mp[q] = make_double2(0.0, 0.0);
#pragma unroll 2
      for (int lv = k5l0; lv <= L; lv += 2) {
        const int idx = Y.o_ent[lv] + k5t;
        double dd[KC];
#pragma unroll
        for (int q = 0; q < KC; q += 2) {
          const double2 d2 = *reinterpret_cast<const double2*>(DtS + idx * KC + q);
          dd[q] = d2.x;
          dd[q + 1] = d2.y;
        }
        const double2 cp = IkS[idx * 2], cm = IkS[idx * 2 + 1];
#pragma unroll
        for (int q = 0; q < KC; ++q) {
          Pp[q].x = fma(dd[q], cp.x, Pp[q].x);
          Pp[q].y = fma(dd[q], cp.y, Pp[q].y);
          Mp[q].x = fma(dd[KC - 1 - q], cm.x, Mp[q].x);
          Mp[q].y = fma(dd[KC - 1 - q], cm.y, Mp[q].y);
        }
      }
      const double sm2q = (k5m2 & 1) ? -1.0 : 1.0;
      double pe[KC], po[KC], me[KC], mo[KC];
#pragma unroll
      for (int q = 0; q < KC; ++q) {
        const double sendP = k5par ? Pp[q].x : Pp[q].y, sendM = k5par ? Mp[q].x : Mp[q].y;
        const double recvP = __shfl_xor_sync(0xffffffffu, sendP, 1);
        const double recvM = __shfl_xor_sync(0xffffffffu, sendM, 1);
        pe[q] = k5par ? recvP : Pp[q].x;
        po[q] = k5par ? Pp[q].y : recvP;
        me[q] = k5par ? recvM : Mp[q].x;
        mo[q] = k5par ? Mp[q].y : recvM;
      }
      // the S blocks are free once every warp has finished step B of the previous pair (whose results are combined
      // behind the same barrier); the sums above did not need them
      __syncthreads();
#pragma unroll
      for (int q = 0; q < KC; ++q) {
        if (k5on) {
#pragma unroll
          for (int oo = 0; oo < 2; ++oo) {
            if (oo >= norient) break;
            const double so = oo ? -1.0 : 1.0;
            // S(+a) = sum_l so^l d I ; S(-a) = (-1)^m2 sum_l (-so)^l d_mirror I_-   (component k5par of each)
            const double sp = fma(so, po[q], pe[q]), sn = sm2q * fma(-so, mo[q], me[q]);
            double* bq = SB + (q * 2 + oo) * Z.sblk;
            bq[k5a * SP + 2 * k5m2 + k5par] = sp;
            if (k5a > 0) bq[(F - k5a) * SP + 2 * k5m2 + k5par] = sn;
          }
        }
      }
    }
    __syncthreads();
    if (pair + jstride < npairs) stage_coeffs(pair + jstride);  // lands during the transforms
    flush(prev, par ^ 1);  // (behind the second barrier: the other warps do not wait for it)
    // ---- step A: column m2 = f, in place (row 16 of the input is zero: |m1| <= 15)
    if (active) {
      double2 y[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) y[n2] = *reinterpret_cast<const double2*>(S + (4 * n2 + n1) * SP + 2 * f);
      if (n1 == 0) y[4] = make_double2(0.0, 0.0);
      double2 v[2][4];
      fl.run(y, v);
      // every lane of the transform has loaded its inputs before any of them can have received all it stores (the
      // shuffles carry the dependence); the warp barrier states that order for the tools
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1)
          *reinterpret_cast<double2*>(S + (2 * n1 + jj + 8 * k1) * SP + 2 * f) = v[jj][k1];
    }
    // step B of a block needs step A of the same block only: the two warps of the block meet on a named barrier
    asm volatile("bar.sync %0, 64;" ::"r"(blk + 1) : "memory");
    // ---- step B: rows alB, alB + 16; arg-max on integer keys
    long long bkey = fo_dkey(-1e300);
    int bix = 0x7fffffff;
    if (active) {
      const double* ra = S + alB * SP;
      const double* rb = ra + 16 * SP;
      double2 y[8];
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        const double2 va = *reinterpret_cast<const double2*>(ra + 2 * (4 * n2 + n1));
        const double2 vb = *reinterpret_cast<const double2*>(rb + 2 * (4 * n2 + n1));
        y[n2] = make_double2(va.x - vb.y, va.y + vb.x);  // V_al + i V_al'
      }
#pragma unroll
      for (int n2 = 4; n2 < 8; ++n2) {
        const int mneg = 32 - (4 * n2 + n1);  // 1 .. 16
        const double2 va = *reinterpret_cast<const double2*>(ra + 2 * (mneg & 15));
        const double2 vb = *reinterpret_cast<const double2*>(rb + 2 * (mneg & 15));
        y[n2] = make_double2(va.x + vb.y, vb.x - va.y);  // conj V_al + i conj V_al'
      }
      if (n1 == 0) y[4] = make_double2(0.0, 0.0);  // m2 = 16
      double2 v[2][4];
      fl.run(y, v);
      // v[jj][k1] = g(alB, d) + i g(alB + 16, d), d = 2 n1 + jj + 8 k1
      if (WANT_GRID) {
        double* g0 = out.grid + (((size_t)pair * norient + o) * F * F * F + (size_t)(alB * F + plane) * F);
        double* g1 = g0 + (size_t)16 * F * F;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int k1 = 0; k1 < 4; ++k1) {
            g0[2 * n1 + jj + 8 * k1] = v[jj][k1].x;
            g1[2 * n1 + jj + 8 * k1] = v[jj][k1].y;
          }
      }
      // thread maximum in index order (row alB before row alB + 16, d increasing): a strict > keeps the smallest
      // index of equal values
      double bv = v[0][0].x;
      int bd = 2 * n1;
#pragma unroll
      for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
          if ((k1 | jj) && v[jj][k1].x > bv) {
            bv = v[jj][k1].x;
            bd = 2 * n1 + jj + 8 * k1;
          }
#pragma unroll
      for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
          if (v[jj][k1].y > bv) {
            bv = v[jj][k1].y;
            bd = 2 * n1 + jj + 8 * k1 + 16 * F * F;
          }
      bkey = fo_dkey(bv);
      bix = (alB * F + plane) * F + bd;
    }
    // ---- reduction: a warp belongs to one orientation (o = (warp >> 1) & 1)
    double* redv = red + 16 * par;
    int* redi = reinterpret_cast<int*>(red + 48 + 8 * par);
    {
      const int hi = (int)(bkey >> 32);
      const int mhi = __reduce_max_sync(0xffffffffu, hi);
      const unsigned lo = hi == mhi ? (unsigned)bkey : 0u;
      const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
      const bool top = hi == mhi && (unsigned)bkey == mlo;
      const int imin = __reduce_min_sync(0xffffffffu, top ? bix : 0x7fffffff);
      if (lane == 0) {
        redv[warp] = fo_dunkey(((long long)mhi << 32) | (long long)mlo);
        redi[warp] = imin;
      }
    }
    // the warps' results are combined behind the NEXT CTA barrier (the one in the next pair's K5, or the final one):
    // no barrier of its own; the scratch alternates between two buffers
    prev = pair;
    par ^= 1;
  }
  __syncthreads();
  flush(prev, par ^ 1);
}

// One CTA per pair (all orientations): best chunk, then findMax's parabola (utils.py:319-338).  The six
// neighbours are evaluated from the coefficients: one pass over the (m1, m2 >= 0) entries forms
// S_k(m1, m2) for the three planes k0-1, k0, k0+1 (adjacent table columns) of each orientation and accumulates
// the four in-plane neighbours and the two out-of-plane ones; all 256 threads share the pass, and both
// orientations share the coefficient loads (the pass is bound by reading the pair's coefficients: one CTA per
// (pair, orientation) read them twice -- 248 -> ~150 us per 2236 LJ38 pairs).
__global__ void __launch_bounds__(256)
sph_final2_kernel(const double2* __restrict__ Ihalf, const double* __restrict__ Dt, int L, int norient,
                  int nchunk, const double* __restrict__ part_val, const int* __restrict__ part_idx,
                  long long* __restrict__ best_idx, double* __restrict__ best_val,
                  double* __restrict__ frac_idx) {
  __shared__ double2 tw[128];
  __shared__ double nbs[8][12];
  const size_t p = blockIdx.x;
  const int L1 = L + 1, W = 2 * L + 1, F = 2 * L1;
  for (int t = threadIdx.x; t < F; t += blockDim.x) {
    double sn, cs;
    sincospi(2.0 * (double)t / (double)F, &sn, &cs);
    tw[t] = make_double2(cs, sn);
  }
  double bv[2];
  int a0[2], k0[2], g0[2], km[2], kp[2];
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    bv[o] = -1e300;
    int bi = 0x7fffffff;
    if (o < norient) {
      const size_t po = p * norient + o;
      for (int c = 0; c < nchunk; ++c) {
        const double v = part_val[po * nchunk + c];
        const int i = part_idx[po * nchunk + c];
        if (v > bv[o] || (v == bv[o] && i < bi)) { bv[o] = v; bi = i; }
      }
    }
    const bool ok = bi != 0x7fffffff;
    a0[o] = ok ? bi / (F * F) : 0;
    k0[o] = ok ? (bi / F) % F : 0;
    g0[o] = ok ? bi % F : 0;
    km[o] = (k0[o] + F - 1) % F;
    kp[o] = (k0[o] + 1) % F;
  }
  __syncthreads();
  // accumulators per orientation: 0,1 = a0 -+ 1; 2,3 = k0 -+ 1; 4,5 = g0 -+ 1   (order of the parabola below)
  double acc[2][6];
#pragma unroll
  for (int o = 0; o < 2; ++o)
#pragma unroll
    for (int q = 0; q < 6; ++q) acc[o][q] = 0.0;
  const double2* Ip = Ihalf + p * (size_t)L1 * W * L1;
  for (int item = threadIdx.x; item < L1 * W; item += blockDim.x) {
    const int m1i = item % W, m2 = item / W;
    const int m1 = m1i - L;
    const int am1 = m1 < 0 ? -m1 : m1;
    const int l0 = am1 > m2 ? am1 : m2;
    double sr[2][3], si[2][3];
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int j = 0; j < 3; ++j) sr[o][j] = si[o][j] = 0.0;
    const double2* ip = Ip + (size_t)item * L1;
    const double* dp = Dt + (size_t)item * L1 * F;
    for (int l = l0; l <= L; ++l) {
      const double2 c = ip[l];
      const double* d = dp + (size_t)l * F;
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        if (o >= norient) break;
        const double sg = ((l & 1) && o) ? -1.0 : 1.0;
        const double cx = sg * c.x, cy = sg * c.y;
        const double d0 = d[km[o]], d1 = d[k0[o]], d2 = d[kp[o]];
        sr[o][0] = fma(d0, cx, sr[o][0]); si[o][0] = fma(d0, cy, si[o][0]);
        sr[o][1] = fma(d1, cx, sr[o][1]); si[o][1] = fma(d1, cy, si[o][1]);
        sr[o][2] = fma(d2, cx, sr[o][2]); si[o][2] = fma(d2, cy, si[o][2]);
      }
    }
    const double wgt = (m2 == 0) ? 1.0 : 2.0;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      if (o >= norient) break;
      const int e0 = ((m1 * a0[o] + m2 * g0[o]) % F + F) % F;
      // term(plane j, phase e) = S_r cos - S_i sin
      auto term = [&](int j, int e) {
        const double2 w = tw[e];
        return wgt * (sr[o][j] * w.x - si[o][j] * w.y);
      };
      acc[o][0] += term(1, ((e0 - m1) % F + F) % F);
      acc[o][1] += term(1, ((e0 + m1) % F + F) % F);
      acc[o][2] += term(0, e0);
      acc[o][3] += term(2, e0);
      acc[o][4] += term(1, ((e0 - m2) % F + F) % F);
      acc[o][5] += term(1, (e0 + m2) % F);
    }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 0; o < 2; ++o)
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[o][q] += __shfl_down_sync(0xffffffffu, acc[o][q], off);
      if (lane == 0) nbs[w][o * 6 + q] = acc[o][q];
    }
  __syncthreads();
  if ((int)threadIdx.x < norient) {
    const int o = threadIdx.x;
    const size_t po = p * norient + o;
    double nb[6];
    for (int q = 0; q < 6; ++q) {
      double s = 0.0;
      for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) s += nbs[ww][o * 6 + q];
      nb[q] = fabs(s);
    }
    const double bvo = o ? bv[1] : bv[0];
    best_val[po] = bvo;
    const int b3[3] = {o ? a0[1] : a0[0], o ? k0[1] : k0[0], o ? g0[1] : g0[0]};
    for (int ax = 0; ax < 3; ++ax) {
      best_idx[po * 3 + ax] = b3[ax];
      const double y1 = nb[2 * ax + 1], y3 = nb[2 * ax], y2 = fabs(bvo);
      frac_idx[po * 3 + ax] = (double)b3[ax] - (y3 - y1) / (2.0 * (2.0 * y2 - y1 - y3));
    }
  }
}

// ------------------------------------------------------------------------------------------
// K5-K7 for large bandwidths (Jmax > IB_LMIN, up to 63: grid up to 128^3).  One CTA per
// (beta plane k, pair, orientation).  The m2 range is processed in blocks of IB_MB so that only the
// stage-A output U[a][m2] of the plane (133 KB at Jmax = 63) and one block of S[m1][m2] stay in
// shared memory; the Wigner table is plane-major (DtK[k][m2][m1][l]: a CTA streams one contiguous
// plane, and consecutive CTAs share it through L2).  Same mathematics as sph_isoft_kernel.
// ------------------------------------------------------------------------------------------
constexpr int IB_LMIN = 32;
constexpr int IB_THREADS = 256;
constexpr int IB_MB = 16;

__global__ void __launch_bounds__(IB_THREADS)
sph_isoft_big_kernel(const double2* __restrict__ Ihalf, const double* __restrict__ DtK, int L, int norient,
                     size_t npairs, IsoOut out) {
  extern __shared__ double2 smb[];
  const int W = 2 * L + 1, L1 = L + 1, F = 2 * L1, H = L1 + 1;
  const int M2p = L1 | 1;
  double2* tw = smb;                       // [F]
  double2* S = tw + F;                     // [W][IB_MB]
  double2* U = S + (size_t)W * IB_MB;      // [F][M2p]
  double* red = (double*)(U + (size_t)F * M2p);
  const int tid = threadIdx.x;
  const int o = (int)(blockIdx.x % norient);
  const size_t p = (blockIdx.x / norient) % npairs;
  const int k = (int)(blockIdx.x / ((size_t)norient * npairs));
  const double so = o ? -1.0 : 1.0;
  for (int t = tid; t < F; t += IB_THREADS) {
    double sn, cs;
    sincospi(2.0 * (double)t / (double)F, &sn, &cs);
    tw[t] = make_double2(cs, sn);
  }
  const double2* Ip = Ihalf + p * (size_t)L1 * W * L1;
  const double* Dk = DtK + (size_t)k * L1 * W * L1;
  for (int m2a = 0; m2a < L1; m2a += IB_MB) {
    const int nb = min(IB_MB, L1 - m2a);
    __syncthreads();  // S of the previous block is consumed; tw is ready
    // ---- K5 for m2 in [m2a, m2a + nb)
    for (int it = tid; it < W * nb; it += IB_THREADS) {
      const int mm = it % nb, m1i = it / nb;
      const int m2 = m2a + mm, m1 = m1i - L;
      const int am1 = m1 < 0 ? -m1 : m1;
      const int l0 = am1 > m2 ? am1 : m2;
      const size_t item = (size_t)m2 * W + m1i;
      const double2* ip = Ip + item * L1;
      const double* dp = Dk + item * L1;
      double2 ae = make_double2(0.0, 0.0), ao = ae;
      for (int l = l0; l <= L; ++l) {
        const double dv = dp[l];
        const double2 c = ip[l];
        if (l & 1) {
          ao.x = fma(dv, c.x, ao.x);
          ao.y = fma(dv, c.y, ao.y);
        } else {
          ae.x = fma(dv, c.x, ae.x);
          ae.y = fma(dv, c.y, ae.y);
        }
      }
      S[(size_t)m1i * IB_MB + mm] = make_double2(fma(so, ao.x, ae.x), fma(so, ao.y, ae.y));
    }
    __syncthreads();
    // ---- stage A: lines m2 of the block, inputs m1 = -L..L, outputs a = 0..F-1 into U[a][m2]
    const int nch = (H + IS_DCA - 1) / IS_DCA;
    for (int item = tid; item < nb * nch; item += IB_THREADS) {
      const int mm = item % nb, ch = item / nb;
      const int m2 = m2a + mm;
      const int d0 = ch * IS_DCA;
      const double2* se = S + mm;
      const double2 c0 = se[(size_t)L * IB_MB];
      double2 P[IS_DCA], Q[IS_DCA];
      int idx[IS_DCA];
#pragma unroll
      for (int t = 0; t < IS_DCA; ++t) {
        P[t] = c0;
        Q[t] = make_double2(0.0, 0.0);
        idx[t] = 0;
      }
      for (int m = 1; m <= L; ++m) {
        const double2 a = se[(size_t)(L + m) * IB_MB], b = se[(size_t)(L - m) * IB_MB];
        const double2 E = make_double2(a.x + b.x, a.y + b.y), O = make_double2(a.x - b.x, a.y - b.y);
#pragma unroll
        for (int t = 0; t < IS_DCA; ++t) {
          int kq = idx[t] + d0 + t;
          kq -= (kq >= F) ? F : 0;
          idx[t] = kq;
          const double2 w = tw[kq];
          P[t].x = fma(E.x, w.x, P[t].x);
          P[t].y = fma(E.y, w.x, P[t].y);
          Q[t].x = fma(O.x, w.y, Q[t].x);
          Q[t].y = fma(O.y, w.y, Q[t].y);
        }
      }
#pragma unroll
      for (int t = 0; t < IS_DCA; ++t) {
        const int d = d0 + t;
        if (d < H) {
          U[(size_t)d * M2p + m2] = make_double2(P[t].x - Q[t].y, P[t].y + Q[t].x);
          if (d != 0 && 2 * d != F)
            U[(size_t)(F - d) * M2p + m2] = make_double2(P[t].x + Q[t].y, P[t].y - Q[t].x);
        }
      }
    }
  }
  __syncthreads();
  // ---- stage B: lines a, half-complex -> real, arg-max
  double bv = -1e300;
  long long bi = 0x7fffffffffffffffLL;
  {
    const int nch = (H + IS_DCB - 1) / IS_DCB;
    for (int item = tid; item < F * nch; item += IB_THREADS) {
      const int a = item % F, ch = item / F;
      const int d0 = ch * IS_DCB;
      const double2* vin = U + (size_t)a * M2p;
      const double v0 = vin[0].x;
      double A[IS_DCB], Bq[IS_DCB];
      int idx[IS_DCB];
#pragma unroll
      for (int t = 0; t < IS_DCB; ++t) {
        A[t] = 0.0;
        Bq[t] = 0.0;
        idx[t] = 0;
      }
      for (int l = 1; l <= L; ++l) {
        const double2 v = vin[l];
#pragma unroll
        for (int t = 0; t < IS_DCB; ++t) {
          int kq = idx[t] + d0 + t;
          kq -= (kq >= F) ? F : 0;
          idx[t] = kq;
          const double2 w = tw[kq];
          A[t] = fma(v.x, w.x, A[t]);
          Bq[t] = fma(v.y, w.y, Bq[t]);
        }
      }
      const long long base = ((long long)a * F + k) * F;
      double* grow = out.grid ? out.grid + ((p * norient + o) * (size_t)F * F * F + (size_t)base) : nullptr;
#pragma unroll
      for (int t = 0; t < IS_DCB; ++t) {
        const int d = d0 + t;
        if (d < H) {
          const double aa = v0 + 2.0 * A[t], bb = 2.0 * Bq[t];
          const double g1 = aa - bb;
          better_s(bv, bi, g1, base + d);
          if (grow) grow[d] = g1;
          if (d != 0 && 2 * d != F) {
            const double g2 = aa + bb;
            better_s(bv, bi, g2, base + (F - d));
            if (grow) grow[F - d] = g2;
          }
        }
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, bv, off);
    const long long oi = __shfl_down_sync(0xffffffffu, bi, off);
    better_s(bv, bi, ov, oi);
  }
  long long* redi = (long long*)(red + 16);
  if ((tid & 31) == 0) {
    red[tid >> 5] = bv;
    redi[tid >> 5] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < IB_THREADS / 32; ++w) better_s(bv, bi, red[w], redi[w]);
    out.part_val[(p * norient + o) * (size_t)F + k] = bv;
    out.part_idx[(p * norient + o) * (size_t)F + k] = bi;
  }
}

// ------------------------------------------------------------------------------------------
// sph_isoft_big2_kernel: the large-bandwidth transform with stages A and B on the FP64 tensor pipe.
// Same decomposition as sph_isoft_big_kernel (CTA per (beta plane, pair, orientation), K5 per block of
// IB2_MB values of m2), but
//   * stage A reads its E / O fragments straight from the S block (E = S(+m1) + S(-m1), O = S(+) - S(-),
//     pitch 2 IB2_MB + 8 doubles: conflict-free) and writes U re / im k-major (BR / BI [m2][a]);
//   * stage B: rows a, half-complex -> real, as in sph_isoft3_kernel;
//   * a work unit is (8-row tile, group of 3 column tiles): no template parameter depends on Jmax; the
//     twiddles come from a one-period table TW[F] through the running index (m d) mod F.
// ------------------------------------------------------------------------------------------
constexpr int IB2_THREADS = 512;
constexpr int IB2_MB = 16;
constexpr int IB2_SP = 2 * IB2_MB + 8;  // doubles per m1 row of the S block (== 8 mod 32 for MB = 16)
constexpr int IB2_NS = 3;               // column tiles per unit

struct Ib2Layout {
  int L, L1, W, F, H, NT, NG, KS, FBp;
  int o_tw, o_s, o_br, o_bi, o_red, total;  // doubles
  Ib2Layout() {}
  explicit Ib2Layout(int L_) {
    L = L_;
    L1 = L + 1;
    W = 2 * L + 1;
    F = 2 * L1;
    H = L1 + 1;
    NT = (H + 7) / 8;
    NG = (NT + IB2_NS - 1) / IB2_NS;
    KS = (L + 3) / 4;
    FBp = F + ((8 - F) % 32 + 32) % 32;  // == 8 (mod 32)
    o_tw = 0;
    o_s = o_tw + 2 * F;
    o_br = o_s + W * IB2_SP;
    o_bi = o_br + L1 * FBp;
    o_red = o_bi + L1 * FBp;
    total = o_red + 48;
  }
};

__global__ void __launch_bounds__(IB2_THREADS, 1)
sph_isoft_big2_kernel(const __grid_constant__ Ib2Layout Y, const double2* __restrict__ Ihalf,
                      const double* __restrict__ DtK, int norient, size_t npairs, IsoOut out) {
  extern __shared__ double smb2[];
  const int L = Y.L, L1 = Y.L1, W = Y.W, F = Y.F, H = Y.H, NT = Y.NT, NG = Y.NG, KS = Y.KS, FBp = Y.FBp;
  double2* TW = reinterpret_cast<double2*>(smb2 + Y.o_tw);  // [F] (cos, sin)(2 pi t / F)
  double* Sd = smb2 + Y.o_s;    // S block as doubles: [m1 + L][IB2_SP], (mm, part) -> 2 mm + part
  double* BR = smb2 + Y.o_br;   // [m2][FBp]  Re U[a][m2]
  double* BI = smb2 + Y.o_bi;
  double* red = smb2 + Y.o_red;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  constexpr int NWARP = IB2_THREADS / 32;
  const int o = (int)(blockIdx.x % norient);
  const size_t p = (blockIdx.x / norient) % npairs;
  const int k = (int)(blockIdx.x / ((size_t)norient * npairs));
  const double so = o ? -1.0 : 1.0;
  for (int t = tid; t < F; t += IB2_THREADS) {
    double sn, cs;
    sincospi(2.0 * (double)t / (double)F, &sn, &cs);
    TW[t] = make_double2(cs, sn);
  }
  const double2* Ip = Ihalf + p * (size_t)L1 * W * L1;
  const double* Dk = DtK + (size_t)k * L1 * W * L1;
  for (int m2a = 0; m2a < L1; m2a += IB2_MB) {
    const int nb = min(IB2_MB, L1 - m2a);
    __syncthreads();  // S of the previous block is consumed; TW is ready
    // ---- K5 for m2 in [m2a, m2a + nb): S[m1][mm] = sum_l so^l d I
    for (int it = tid; it < W * nb; it += IB2_THREADS) {
      const int mm = it % nb, m1i = it / nb;
      const int m2 = m2a + mm, m1 = m1i - L;
      const int am1 = m1 < 0 ? -m1 : m1;
      const int l0 = am1 > m2 ? am1 : m2;
      const size_t item = (size_t)m2 * W + m1i;
      const double2* ip = Ip + item * L1;
      const double* dp = Dk + item * L1;
      double2 ae = make_double2(0.0, 0.0), ao = ae;
      for (int l = l0; l <= L; ++l) {
        const double dv = dp[l];
        const double2 c = ip[l];
        if (l & 1) {
          ao.x = fma(dv, c.x, ao.x);
          ao.y = fma(dv, c.y, ao.y);
        } else {
          ae.x = fma(dv, c.x, ae.x);
          ae.y = fma(dv, c.y, ae.y);
        }
      }
      *reinterpret_cast<double2*>(Sd + (size_t)m1i * IB2_SP + 2 * mm) =
          make_double2(fma(so, ao.x, ae.x), fma(so, ao.y, ae.y));
    }
    // rows of a partial last block: keep them finite
    for (int it = tid; it < W * (IB2_MB - nb); it += IB2_THREADS) {
      const int mm = nb + it % (IB2_MB - nb), m1i = it / (IB2_MB - nb);
      *reinterpret_cast<double2*>(Sd + (size_t)m1i * IB2_SP + 2 * mm) = make_double2(0.0, 0.0);
    }
    __syncthreads();
    // ---- stage A: rows (mm, part) of the block, k = m1 = 1..L, outputs a; units (row tile, column group)
    const int ntile = (2 * nb + 7) >> 3;
    for (int unit = warp; unit < ntile * NG; unit += NWARP) {
      const int tile = unit % ntile, cg = unit / ntile;
      const int row = tile * 8 + g;                 // 2 mm + part
      const int part = row & 1, mm = row >> 1;
      const bool rvalid = mm < nb;
      const double c0 = Sd[(size_t)L * IB2_SP + row];
      double P[IB2_NS][2], Q[IB2_NS][2];
      int idx[IB2_NS], dstep[IB2_NS];
      bool dval[IB2_NS];
#pragma unroll
      for (int u = 0; u < IB2_NS; ++u) {
        const int d = (cg * IB2_NS + u) * 8 + g;    // output of this lane's B-fragment column
        dval[u] = cg * IB2_NS + u < NT && d < H;
        idx[u] = ((t4 + 1) * d) % F;
        dstep[u] = (4 * d) % F;
        P[u][0] = P[u][1] = c0;
        Q[u][0] = Q[u][1] = 0.0;
      }
      for (int ks = 0; ks < KS; ++ks) {
        const int m = 4 * ks + t4 + 1;
        const int mc = m <= L ? m : L;
        const double sp = Sd[(size_t)(L + mc) * IB2_SP + row], sm = Sd[(size_t)(L - mc) * IB2_SP + row];
        const double fe = sp + sm, fo = sp - sm;
#pragma unroll
        for (int u = 0; u < IB2_NS; ++u) {
          const double2 w = TW[idx[u]];
          const bool on = dval[u] && m <= L;
          fo_dmma(P[u], fe, on ? w.x : 0.0);
          fo_dmma(Q[u], fo, on ? w.y : 0.0);
          idx[u] += dstep[u];
          idx[u] -= idx[u] >= F ? F : 0;
        }
      }
      double* Bout = (part ? BI : BR) + (size_t)(m2a + mm) * FBp;
      const double sgn = part ? 1.0 : -1.0;
#pragma unroll
      for (int u = 0; u < IB2_NS; ++u)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int d = (cg * IB2_NS + u) * 8 + t4 * 2 + q;  // output of this lane's C-fragment column
          const double qx = __shfl_xor_sync(0xffffffffu, Q[u][q], 4);
          if (rvalid && cg * IB2_NS + u < NT && d < H) {
            Bout[d] = fma(sgn, qx, P[u][q]);
            if (d != 0 && 2 * d != F) Bout[F - d] = fma(-sgn, qx, P[u][q]);
          }
        }
    }
  }
  __syncthreads();
  // ---- stage B: rows a, k = m2 = 1..L (re with cos, im with sin), outputs gamma; arg-max
  double bv = -1e300;
  long long bi = 0x7fffffffffffffffLL;
  {
    const int ntile = (F + 7) >> 3;  // F = 2 (L + 1) is a multiple of 8 only when L + 1 is a multiple of 4
    for (int unit = warp; unit < ntile * NG; unit += NWARP) {
      const int tile = unit % ntile, cg = unit / ntile;
      const int a = tile * 8 + g;
      const bool avalid = a < F;
      const int ac = avalid ? a : 0;
      const double v0h = 0.5 * BR[ac];
      double A[IB2_NS][2], Bq[IB2_NS][2];
      int idx[IB2_NS], dstep[IB2_NS];
      bool dval[IB2_NS];
#pragma unroll
      for (int u = 0; u < IB2_NS; ++u) {
        const int d = (cg * IB2_NS + u) * 8 + g;
        dval[u] = cg * IB2_NS + u < NT && d < H;
        idx[u] = ((t4 + 1) * d) % F;
        dstep[u] = (4 * d) % F;
        A[u][0] = A[u][1] = v0h;
        Bq[u][0] = Bq[u][1] = 0.0;
      }
      for (int ks = 0; ks < KS; ++ks) {
        const int m = 4 * ks + t4 + 1;
        const int mc = m <= L ? m : L;
        const double vr = BR[(size_t)mc * FBp + ac], vi = BI[(size_t)mc * FBp + ac];
#pragma unroll
        for (int u = 0; u < IB2_NS; ++u) {
          const double2 w = TW[idx[u]];
          const bool on = dval[u] && m <= L;
          fo_dmma(A[u], vr, on ? w.x : 0.0);
          fo_dmma(Bq[u], vi, on ? w.y : 0.0);
          idx[u] += dstep[u];
          idx[u] -= idx[u] >= F ? F : 0;
        }
      }
      const long long base = ((long long)ac * F + k) * F;
      double* grow = out.grid ? out.grid + ((p * norient + o) * (size_t)F * F * F + (size_t)base) : nullptr;
#pragma unroll
      for (int u = 0; u < IB2_NS; ++u)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int d = (cg * IB2_NS + u) * 8 + t4 * 2 + q;
          if (avalid && cg * IB2_NS + u < NT && d < H) {
            const double g1 = 2.0 * (A[u][q] - Bq[u][q]);  // Re(V e^{+i theta})
            better_s(bv, bi, g1, base + d);
            if (grow) grow[d] = g1;
            if (d != 0 && 2 * d != F) {
              const double g2 = 2.0 * (A[u][q] + Bq[u][q]);
              better_s(bv, bi, g2, base + (F - d));
              if (grow) grow[F - d] = g2;
            }
          }
        }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, bv, off);
    const long long oi = __shfl_down_sync(0xffffffffu, bi, off);
    better_s(bv, bi, ov, oi);
  }
  long long* redi = (long long*)(red + 16);
  if (lane == 0) {
    red[warp] = bv;
    redi[warp] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NWARP; ++w) better_s(bv, bi, red[w], redi[w]);
    out.part_val[(p * norient + o) * (size_t)F + k] = bv;
    out.part_idx[(p * norient + o) * (size_t)F + k] = bi;
  }
}

size_t isoft_big_smem(int L) {
  const int W = 2 * L + 1, L1 = L + 1, F = 2 * L1, M2p = L1 | 1;
  return ((size_t)F + (size_t)W * IB_MB + (size_t)F * M2p) * 16 + 32 * 8;
}

// ---------------------------------------------------------------------------------- host side

int grid_for(size_t total, int threads, int cap = 1 << 16) {
  size_t b = (total + threads - 1) / threads;
  if (b > (size_t)cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int ensure_wigner(fo_ctx* ctx, int L) {
  if (ctx->wig.Jmax == L && ctx->wig.d_table) return FO_OK;
  if (ctx->wig.d_table) {
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->wig.d_table);
    ctx->wig.d_table = nullptr;
    ctx->wig.Jmax = -1;
  }
  const size_t n = (size_t)(L + 1) * (2 * L + 1) * (L + 1) * (2 * L + 2);
  cudaError_t e = cudaMalloc(&ctx->wig.d_table, n * 8);
  if (e != cudaSuccess)
    return fo_fail(ctx, FO_ERR_NOMEM, "cudaMalloc of the Wigner table (%zu bytes) failed", n * 8);
  ctx->wig.kmajor = L > IB_LMIN;
  sph_wigner_kernel<<<grid_for((size_t)(L + 1) * (2 * L + 1) * (2 * L + 2), 128), 128, 0, ctx->stream>>>(
      L, dt_stride(L, ctx->wig.kmajor), ctx->wig.d_table);
  FO_LAUNCH_CHECK(ctx);
  ctx->wig.Jmax = L;
  ctx->wig.bytes = n * 8;
  // packed per-chunk slices for the fast iSOFT kernel (bandwidths with 2B % 4 == 0)
  if (ctx->wig.d_packed) {
    cudaFree(ctx->wig.d_packed);
    ctx->wig.d_packed = nullptr;
  }
  ctx->wig.packed_kc = 0;
  if (ctx->wig.d_packed4) {
    cudaFree(ctx->wig.d_packed4);
    ctx->wig.d_packed4 = nullptr;
  }
  if (L <= 15 && !ctx->wig.kmajor) {  // bandwidths of the fast iSOFT kernel (one stage-A tile per warp)
    ctx->wig.packed_kc = FO_I3_KC;
    const I2Layout Y(L, FO_I3_KC);
    if (cudaMalloc(&ctx->wig.d_packed, (size_t)Y.nchunk * Y.dts * 8) != cudaSuccess)
      return fo_fail(ctx, FO_ERR_NOMEM, "cudaMalloc of the packed Wigner table failed");
    sph_wigner_pack_kernel<<<grid_for((size_t)Y.nchunk * Y.dts, 256), 256, 0, ctx->stream>>>(
        ctx->wig.d_table, Y, ctx->wig.d_packed);
    FO_LAUNCH_CHECK(ctx);
    if ((L & 1) && L >= 1) {  // four planes per chunk (sph_isoft4_kernel<4, ...>): 2 (L + 1) divisible by 4
      const I2Layout Y4(L, 4);
      if (cudaMalloc(&ctx->wig.d_packed4, (size_t)Y4.nchunk * Y4.dts * 8) != cudaSuccess)
        return fo_fail(ctx, FO_ERR_NOMEM, "cudaMalloc of the packed Wigner table failed");
      sph_wigner_pack_kernel<<<grid_for((size_t)Y4.nchunk * Y4.dts, 256), 256, 0, ctx->stream>>>(
          ctx->wig.d_table, Y4, ctx->wig.d_packed4);
      FO_LAUNCH_CHECK(ctx);
    }
  }
  return FO_OK;
}

int check_L(fo_ctx* ctx, int64_t L) {
  if (L < 0 || L > 63) return fo_fail(ctx, FO_ERR_UNSUPPORTED, "Jmax=%lld outside 0..63", (long long)L);
  return FO_OK;
}

size_t isoft_smem(int L, int KC) {
  const int B = L + 1, W = 2 * L + 1, L1 = L + 1, F = 2 * B, M2p = L1 | 1;
  return ((size_t)F + 2 * (size_t)W * KC * L1 + (size_t)F * KC * M2p) * 16 + 32 * 8;
}

// iSOFT + arg-max for npairs coefficient sets already in the Ihalf layout (device).
// whether run_isoft takes a kernel that reads the packed coefficients (call after ensure_wigner)
bool isoft_packed_applies(const fo_ctx* ctx, int L) {
  if (!ctx->wig.d_packed || ctx->force_generic || L < 1) return false;
  const I2Layout Y(L, ctx->wig.packed_kc);
  const int KSq = (L + 3) / 4;
  const bool nyq = (Y.H % 8) == 1;
  const int NTq = nyq ? (Y.H - 1) / 8 : (Y.H + 7) / 8;
  const int code = KSq * 100 + NTq * 10 + (nyq ? 1 : 0);
  const bool have = code == 421 || code == 420 || code == 320 || code == 220 || code == 211 || code == 210 ||
                    code == 110;
  return have && (size_t)Y.total3 * 8 <= ctx->prop.sharedMemPerBlockOptin;
}

// ipk_ready: the packed coefficients of these pairs are already in the FO_SCR_IPK scratch (written by
// sph_direct2_kernel): sph_ipack_kernel is skipped
int run_isoft(fo_ctx* ctx, const double2* d_Ihalf, int64_t npairs, int L, int norient,
              long long* d_best_idx, double* d_best_val, double* d_frac, double* d_grid, bool ipk_ready = false) {
  if (npairs == 0) return FO_OK;
  FO_CHECK(ensure_wigner(ctx, L));
  const int F = 2 * (L + 1);
  if (ctx->wig.d_packed && !ctx->force_generic && L >= 1) {
    const int KC = ctx->wig.packed_kc;
    const I2Layout Y(L, KC);
    const size_t smem3 = (size_t)Y.total3 * 8;
    const int KSq = (L + 3) / 4;
    const bool nyq = (Y.H % 8) == 1;
    const int NTq = nyq ? (Y.H - 1) / 8 : (Y.H + 7) / 8;
    const int code = KSq * 100 + NTq * 10 + (nyq ? 1 : 0);
    const bool have = code == 421 || code == 420 || code == 320 || code == 220 || code == 211 || code == 210 ||
                      code == 110;
    if (have && smem3 <= ctx->prop.sharedMemPerBlockOptin) {
      const int nch = Y.nchunk;
      void* part = nullptr;
      FO_CHECK(fo_scratch(ctx, FO_SCR_MISC, (size_t)npairs * norient * nch * (8 + 4) + 64, &part));
      Iso2Out o2;
      o2.part_val = (double*)part;
      o2.part_idx = (int*)(o2.part_val + (size_t)npairs * norient * nch);
      o2.grid = d_grid;
      void* ipk = nullptr;
      FO_CHECK(fo_scratch(ctx, FO_SCR_IPK, (size_t)npairs * Y.ipk * 16, &ipk));
      const double2* d_Ipk = (const double2*)ipk;
      // persistent grid: (CTAs per SM) x SMs, a multiple of the chunk count
      int per = ctx->prop.multiProcessorCount * (4 / KC) / nch;
      if (per < 1) per = 1;
      if ((int64_t)per > npairs) per = (int)npairs;
      const unsigned blocks = (unsigned)(per * nch);
      fo_prof_scope prof(ctx, FO_PROF_SPH_ISOFT);
      if (!ipk_ready) {
        sph_ipack_kernel<<<grid_for((size_t)npairs * Y.ipk, 256), 256, 0, ctx->stream>>>(d_Ihalf, Y, (size_t)npairs,
                                                                                          (double2*)ipk);
        FO_LAUNCH_CHECK(ctx);
      }
      {  // stages A -> B chained in registers (sph_isoft4_kernel); four planes per CTA when 2 (L + 1) % 4 == 0
        const int KC4 = (ctx->wig.d_packed4 && ctx->isoft_variant != 42) ? 4 : 2;
        const I2Layout Y4(L, KC4);
        const I4Layout Z4(Y4);
        const size_t smem4 = (size_t)Z4.total4 * 8;
        const int KS4 = L / 4 + 1;
        const int code4 = KS4 * 100 + NTq * 10 + (nyq ? 1 : 0);
        const bool have4 = code4 == 421 || code4 == 420 || code4 == 320 || code4 == 211 || code4 == 210 || code4 == 110;
        if (KC == 2 && have4 && ctx->isoft_variant != 3 && smem4 <= ctx->prop.sharedMemPerBlockOptin) {
          const int nch4 = Y4.nchunk;
          if (KC4 == 4) FO_CHECK(fo_scratch(ctx, FO_SCR_MISC, (size_t)npairs * norient * nch * (8 + 4) + 64, &part));
          o2.part_val = (double*)part;
          o2.part_idx = (int*)(o2.part_val + (size_t)npairs * norient * nch4);
          int per4 = ctx->prop.multiProcessorCount * (4 / KC4) / nch4;
          if (per4 < 1) per4 = 1;
          if ((int64_t)per4 > npairs) per4 = (int)npairs;
          const unsigned blocks4 = (unsigned)(per4 * nch4);
          const double* dtp = KC4 == 4 ? ctx->wig.d_packed4 : ctx->wig.d_packed;
#define FO_I4_GO(KC_, KS_, NT_, NYQ_, G_)                                                                     \
  do {                                                                                                        \
    FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft4_kernel<KC_, KS_, NT_, NYQ_, G_>,                             \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));              \
    sph_isoft4_kernel<KC_, KS_, NT_, NYQ_, G_><<<blocks4, KC_ * 128, smem4, ctx->stream>>>(                   \
        Y4, Z4, d_Ipk, dtp, (int)npairs, norient, o2);                                                        \
  } while (0)
#define FO_I4_LAUNCH(KS_, NT_, NYQ_)                      \
  do {                                                    \
    if (KC4 == 4) {                                       \
      if (d_grid) FO_I4_GO(4, KS_, NT_, NYQ_, true);      \
      else FO_I4_GO(4, KS_, NT_, NYQ_, false);            \
    } else {                                              \
      if (d_grid) FO_I4_GO(2, KS_, NT_, NYQ_, true);      \
      else FO_I4_GO(2, KS_, NT_, NYQ_, false);            \
    }                                                     \
  } while (0)
          const I5Layout Z5(Y4);
          const size_t smem5 = (size_t)Z5.total5 * 8;
          if (KC4 == 4 && L == 15 && ctx->isoft_variant != 4 && smem5 <= ctx->prop.sharedMemPerBlockOptin) {
            // the transforms as FFTs on the FP64 vector pipe (sph_isoft5_kernel)
            int per5 = ctx->prop.multiProcessorCount / nch4;
            if (per5 < 1) per5 = 1;
            if ((int64_t)per5 > npairs) per5 = (int)npairs;
            if (d_grid) {
              FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));
              sph_isoft5_kernel<true><<<(unsigned)(per5 * nch4), 512, smem5, ctx->stream>>>(Y4, Z5, d_Ipk, dtp, (int)npairs, norient, o2);
            } else {
              FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));
              sph_isoft5_kernel<false><<<(unsigned)(per5 * nch4), 512, smem5, ctx->stream>>>(Y4, Z5, d_Ipk, dtp, (int)npairs, norient, o2);
            }
          } else
          switch (code4) {
            case 421: FO_I4_LAUNCH(4, 2, true); break;    // Jmax 15
            case 420: FO_I4_LAUNCH(4, 2, false); break;   // Jmax 12 .. 14
            case 320: FO_I4_LAUNCH(3, 2, false); break;   // Jmax 8 .. 11
            case 211: FO_I4_LAUNCH(2, 1, true); break;    // Jmax 7
            case 210: FO_I4_LAUNCH(2, 1, false); break;   // Jmax 4 .. 6
            default: FO_I4_LAUNCH(1, 1, false); break;    // Jmax 1 .. 3
          }
#undef FO_I4_LAUNCH
#undef FO_I4_GO
          FO_LAUNCH_CHECK(ctx);
          sph_final2_kernel<<<(unsigned)npairs, 256, 0, ctx->stream>>>(
              d_Ihalf, ctx->wig.d_table, L, norient, nch4, o2.part_val, o2.part_idx, d_best_idx,
              d_best_val, d_frac);
          FO_LAUNCH_CHECK(ctx);
          return FO_OK;
        }
      }
#define FO_I3_GO(KC_, KS_, NT_, NYQ_, G_)                                                                    \
  do {                                                                                                       \
    FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft3_kernel<KC_, KS_, NT_, NYQ_, G_>,                            \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));             \
    sph_isoft3_kernel<KC_, KS_, NT_, NYQ_, G_><<<blocks, KC_ * 128, smem3, ctx->stream>>>(                   \
        Y, d_Ipk, ctx->wig.d_packed, (int)npairs, norient, o2);                                              \
  } while (0)
#define FO_I3_LAUNCH(KS_, NT_, NYQ_)                         \
  do {                                                       \
    if (d_grid) FO_I3_GO(FO_I3_KC, KS_, NT_, NYQ_, true);    \
    else FO_I3_GO(FO_I3_KC, KS_, NT_, NYQ_, false);          \
  } while (0)
      switch (code) {
        case 421: FO_I3_LAUNCH(4, 2, true); break;    // Jmax 15
        case 420: FO_I3_LAUNCH(4, 2, false); break;   // Jmax 13, 14
        case 320: FO_I3_LAUNCH(3, 2, false); break;   // Jmax 9 .. 12
        case 220: FO_I3_LAUNCH(2, 2, false); break;   // Jmax 8
        case 211: FO_I3_LAUNCH(2, 1, true); break;    // Jmax 7
        case 210: FO_I3_LAUNCH(2, 1, false); break;   // Jmax 5, 6
        default: FO_I3_LAUNCH(1, 1, false); break;    // Jmax 1 .. 4
      }
#undef FO_I3_LAUNCH
#undef FO_I3_GO
      FO_LAUNCH_CHECK(ctx);
      sph_final2_kernel<<<(unsigned)npairs, 256, 0, ctx->stream>>>(
          d_Ihalf, ctx->wig.d_table, L, norient, nch, o2.part_val, o2.part_idx, d_best_idx,
          d_best_val, d_frac);
      FO_LAUNCH_CHECK(ctx);
      return FO_OK;
    }
  }
  if (ctx->wig.kmajor) {
    const size_t smem = isoft_big_smem(L);
    if (smem > ctx->prop.sharedMemPerBlockOptin)
      return fo_fail(ctx, FO_ERR_UNSUPPORTED, "Jmax=%d needs %zu bytes of shared memory per block", L, smem);
    // keep the per-launch partial arrays bounded: at most ~2^20 CTAs per launch
    void* part = nullptr;
    FO_CHECK(fo_scratch(ctx, FO_SCR_MISC, (size_t)npairs * norient * F * 16, &part));
    IsoOut o;
    o.part_val = (double*)part;
    o.part_idx = (long long*)(o.part_val + (size_t)npairs * norient * F);
    o.grid = d_grid;
    fo_prof_scope prof(ctx, FO_PROF_SPH_ISOFT);
    const size_t blocks = (size_t)F * npairs * norient;
    if (blocks > 0x7fffffffULL) return fo_fail(ctx, FO_ERR_UNSUPPORTED, "too many pairs in one iSOFT launch");
    const Ib2Layout Y2(L);
    const size_t smem2 = (size_t)Y2.total * 8;
    if (!ctx->force_generic && smem2 <= ctx->prop.sharedMemPerBlockOptin) {  // stages A / B on the tensor pipe
      FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft_big2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      sph_isoft_big2_kernel<<<(unsigned)blocks, IB2_THREADS, smem2, ctx->stream>>>(Y2, d_Ihalf, ctx->wig.d_table,
                                                                                  norient, (size_t)npairs, o);
    } else {
      FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sph_isoft_big_kernel<<<(unsigned)blocks, IB_THREADS, smem, ctx->stream>>>(d_Ihalf, ctx->wig.d_table, L, norient,
                                                                               (size_t)npairs, o);
    }
    FO_LAUNCH_CHECK(ctx);
    void* nbv = nullptr;
    FO_CHECK(fo_scratch(ctx, FO_SCR_DBG, (size_t)npairs * norient * 48 + 64, &nbv));
    sph_final_kernel<<<dim3((unsigned)(npairs * norient), 6), 256, 0, ctx->stream>>>(
        d_Ihalf, ctx->wig.d_table, dt_stride(L, true), L, norient, F, o.part_val, o.part_idx, d_best_idx,
        d_best_val, (double*)nbv);
    FO_LAUNCH_CHECK(ctx);
    sph_parabola_kernel<<<grid_for((size_t)npairs * norient, 128), 128, 0, ctx->stream>>>(
        (size_t)npairs * norient, d_best_idx, d_best_val, (const double*)nbv, d_frac);
    FO_LAUNCH_CHECK(ctx);
    return FO_OK;
  }
  int KC = 4;
  while (KC > 1 && (isoft_smem(L, KC) > 100 * 1024 || F % KC)) KC >>= 1;
  const size_t smem = isoft_smem(L, KC);
  if (smem > ctx->prop.sharedMemPerBlockOptin)
    return fo_fail(ctx, FO_ERR_UNSUPPORTED,
                   "Jmax=%d needs %zu bytes of shared memory per block in the iSOFT kernel (> %zu); "
                   "large-bandwidth transforms are not supported yet", L, smem,
                   (size_t)ctx->prop.sharedMemPerBlockOptin);
  const int nchunk = F / KC;
  void* part = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_MISC, (size_t)npairs * norient * nchunk * 16, &part));
  IsoOut o;
  o.part_val = (double*)part;
  o.part_idx = (long long*)(o.part_val + (size_t)npairs * norient * nchunk);
  o.grid = d_grid;
  {
    fo_prof_scope prof(ctx, FO_PROF_SPH_ISOFT);
    const unsigned blocks = (unsigned)(npairs * nchunk);
    if (KC == 4) {
      FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sph_isoft_kernel<4><<<blocks, IS_THREADS, smem, ctx->stream>>>(d_Ihalf, ctx->wig.d_table, L, norient, nchunk, o);
    } else if (KC == 2) {
      FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sph_isoft_kernel<2><<<blocks, IS_THREADS, smem, ctx->stream>>>(d_Ihalf, ctx->wig.d_table, L, norient, nchunk, o);
    } else {
      FO_CUDA(ctx, cudaFuncSetAttribute(sph_isoft_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sph_isoft_kernel<1><<<blocks, IS_THREADS, smem, ctx->stream>>>(d_Ihalf, ctx->wig.d_table, L, norient, nchunk, o);
    }
    FO_LAUNCH_CHECK(ctx);
    void* nbv = nullptr;
    FO_CHECK(fo_scratch(ctx, FO_SCR_DBG, (size_t)npairs * norient * 48 + 64, &nbv));
    sph_final_kernel<<<dim3((unsigned)(npairs * norient), 6), 256, 0, ctx->stream>>>(
        d_Ihalf, ctx->wig.d_table, dt_stride(L, false), L, norient, nchunk, o.part_val, o.part_idx, d_best_idx,
        d_best_val, (double*)nbv);
    FO_LAUNCH_CHECK(ctx);
    sph_parabola_kernel<<<grid_for((size_t)npairs * norient, 128), 128, 0, ctx->stream>>>(
        (size_t)npairs * norient, d_best_idx, d_best_val, (const double*)nbv, d_frac);
    FO_LAUNCH_CHECK(ctx);
  }
  return FO_OK;
}

size_t ihalf_elems(int L) { return (size_t)(L + 1) * (2 * L + 1) * (L + 1); }

// group id per atom on the device (from the ctx permutation groups); atoms in no group get -1-i.  Cached in
// the ctx until fo_set_perm changes the groups: the per-chunk calls of the host-buffer pipeline must not
// synchronise the stream (that would serialise the double-buffered copies against the kernels).
int upload_gid(fo_ctx* ctx, int64_t natoms, int** d_gid) {
  if (ctx->gid_natoms == natoms && ctx->scratch[FO_SCR_GID].ptr) {
    *d_gid = (int*)ctx->scratch[FO_SCR_GID].ptr;
    return FO_OK;
  }
  std::vector<int> gid(natoms);
  for (int64_t i = 0; i < natoms; ++i) gid[i] = -1 - (int)i;
  const int ng = (int)ctx->h_goff.size() - 1;
  for (int g = 0; g < ng; ++g)
    for (int a = ctx->h_goff[g]; a < ctx->h_goff[g + 1]; ++a) gid[ctx->h_gidx[a]] = g;
  void* p = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_GID, (size_t)natoms * 4 + 64, &p));
  FO_CUDA(ctx, cudaMemcpyAsync(p, gid.data(), (size_t)natoms * 4, cudaMemcpyHostToDevice, ctx->stream));
  FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // gid is a stack vector
  ctx->gid_natoms = natoms;
  *d_gid = (int*)p;
  return FO_OK;
}

// direct coefficients of np pairs (device positions) into d_Ihalf
// d_Ipk (optional; only honoured by the streaming form, see ipk_fused_applies): packed copy for run_isoft
int run_direct(fo_ctx* ctx, const double* d_posA, const double* d_posB, int64_t np, int64_t natoms, int L,
               double sigma, const int* d_gid, double2* d_Ihalf, int* d_status, char* work,
               double2* d_Ipk = nullptr) {
  if (np == 0) return FO_OK;
  const int NLM = nlm_of(L);
  if (const int nslots = direct2_slots(ctx, natoms, L)) {
    // streaming form: operands in the fragment layouts (work: YswA | YswB | RA | RB | Bsw)
    const D2Layout Y((int)natoms, L);
    double* YswA = (double*)work;
    double* YswB = YswA + (size_t)np * Y.ysz;
    double* RA = YswB + (size_t)np * Y.ysz;
    double* RB = RA + (size_t)np * natoms;
    double* Bsw = RB + (size_t)np * natoms;
    fo_prof_scope prof(ctx, FO_PROF_SPH_COEF);
    const size_t tot = (size_t)np * Y.N8 * Y.CW;
    const size_t smem_prep = ((size_t)2 * NLM + L + 1) * 8;
    FO_CUDA(ctx, cudaFuncSetAttribute(sph_prep2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
    FO_CUDA(ctx, cudaFuncSetAttribute(sph_prep2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
    sph_prep2_kernel<false><<<grid_for(tot, 128, 148 * 8), 128, smem_prep, ctx->stream>>>(d_posA, Y, (size_t)np, YswA, RA, d_status);
    FO_LAUNCH_CHECK(ctx);
    sph_prep2_kernel<true><<<grid_for(tot, 128, 148 * 8), 128, smem_prep, ctx->stream>>>(d_posB, Y, (size_t)np, YswB, RB, d_status);
    FO_LAUNCH_CHECK(ctx);
    sph_bessel2_kernel<<<grid_for((size_t)np * Y.bsz, 128), 128, 0, ctx->stream>>>(RA, RB, d_gid, Y, sigma, (size_t)np, Bsw);
    FO_LAUNCH_CHECK(ctx);
    const size_t smem = direct2_smem(Y, nslots);
    const int per_sm = (int)std::min<size_t>(4, ((size_t)227 * 1024) / (smem + 1024));
    const int blocks = (int)std::min<int64_t>(np, (int64_t)per_sm * ctx->prop.multiProcessorCount);
#define FO_D2_CASE(N_)                                                                                              \
  case N_:                                                                                                          \
    FO_CUDA(ctx, cudaFuncSetAttribute(sph_direct2_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    sph_direct2_kernel<N_><<<blocks, D2_THREADS, smem, ctx->stream>>>(YswA, YswB, Bsw, Y, nslots, (size_t)np, d_Ihalf, d_Ipk); \
    break;
    switch (Y.NCT) {
      FO_D2_CASE(1) FO_D2_CASE(2) FO_D2_CASE(3) FO_D2_CASE(4) FO_D2_CASE(5) FO_D2_CASE(6) FO_D2_CASE(7) FO_D2_CASE(8)
      default: return fo_fail(ctx, FO_ERR_UNSUPPORTED, "sph_direct2: %d atoms", (int)natoms);
    }
#undef FO_D2_CASE
    FO_LAUNCH_CHECK(ctx);
    return FO_OK;
  }
  // work: YA | YB | RA | RB | Bes
  double2* YA = (double2*)work;
  double2* YB = YA + (size_t)np * natoms * NLM;
  double* RA = (double*)(YB + (size_t)np * natoms * NLM);
  double* RB = RA + (size_t)np * natoms;
  double* Bes = RB + (size_t)np * natoms;
  fo_prof_scope prof(ctx, FO_PROF_SPH_COEF);
  const size_t tot = (size_t)np * natoms * (L + 1);
  const size_t smem_prep = ((size_t)2 * NLM + L + 1) * 8;
  FO_CUDA(ctx, cudaFuncSetAttribute(sph_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
  // YA and YB are adjacent in the work buffer and posA / posB need not be: two launches
  sph_prep_kernel<<<grid_for(tot, 128, 148 * 16), 128, smem_prep, ctx->stream>>>(d_posA, (int)natoms, L, (size_t)np, YA, RA, d_status);
  FO_LAUNCH_CHECK(ctx);
  sph_prep_kernel<<<grid_for(tot, 128, 148 * 16), 128, smem_prep, ctx->stream>>>(d_posB, (int)natoms, L, (size_t)np, YB, RB, d_status);
  FO_LAUNCH_CHECK(ctx);
  sph_bessel_kernel<<<grid_for((size_t)np * natoms * natoms, 128), 128, 0, ctx->stream>>>(
      RA, RB, d_gid, (int)natoms, L, sigma, (size_t)np, Bes);
  FO_LAUNCH_CHECK(ctx);
  if (natoms >= ctx->direct_gemm_min) {
    double* T = Bes + (size_t)np * (L + 1) * natoms * natoms;
    double* C2 = T + (size_t)np * 2 * NLM * natoms;
    const unsigned rowt = (unsigned)((2 * (L + 1) + DG_TM - 1) / DG_TM);
    dim3 g1((unsigned)((natoms + DG_TN - 1) / DG_TN), rowt, (unsigned)(np * (L + 1)));
    FO_CUDA(ctx, cudaFuncSetAttribute(sph_direct_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DG_SMEM));
    FO_CUDA(ctx, cudaFuncSetAttribute(sph_direct_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DG_SMEM));
    sph_direct_gemm_kernel<1><<<g1, DG_THREADS, DG_SMEM, ctx->stream>>>((const double*)YA, Bes, (int)natoms, L, T);
    FO_LAUNCH_CHECK(ctx);
    dim3 g2(rowt, rowt, (unsigned)(np * (L + 1)));
    sph_direct_gemm_kernel<2><<<g2, DG_THREADS, DG_SMEM, ctx->stream>>>((const double*)YB, T, (int)natoms, L, C2);
    FO_LAUNCH_CHECK(ctx);
    sph_direct_combine_kernel<<<grid_for((size_t)np * ihalf_elems(L), 256), 256, 0, ctx->stream>>>(
        C2, L, (size_t)np, d_Ihalf);
    FO_LAUNCH_CHECK(ctx);
    return FO_OK;
  }
  dim3 grid((unsigned)(L + 1), (unsigned)np);
  const size_t smem_mma = direct_mma_smem(natoms, L);
  if (!ctx->force_generic && smem_mma <= 110 * 1024) {
    FO_CUDA(ctx, cudaFuncSetAttribute(sph_direct_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
    sph_direct_mma_kernel<<<grid, DS_THREADS, smem_mma, ctx->stream>>>(YA, YB, Bes, (int)natoms, L, d_Ihalf);
    FO_LAUNCH_CHECK(ctx);
    return FO_OK;
  }
  const size_t smem = ((size_t)(L + 1) * DIR_TK * 2 + (size_t)(2 * L + 1) * (L + 1)) * 16;
  FO_CUDA(ctx, cudaFuncSetAttribute(sph_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sph_direct_kernel<<<grid, 128, smem, ctx->stream>>>(YA, YB, Bes, (int)natoms, L, d_Ihalf);
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}

size_t direct_work_bytes(const fo_ctx* ctx, int64_t np, int64_t natoms, int L) {
  const size_t NLM = nlm_of(L);
  // YA, YB | RA, RB | Bes | T, C2 (GEMM path for large clusters only)
  if (direct2_slots(ctx, natoms, L)) {
    const D2Layout Y((int)natoms, L);
    return (size_t)np * Y.ysz * 16 + (size_t)np * natoms * 16 + (size_t)np * (L + 1) * Y.bsz * 8 + 256;
  }
  size_t b = (size_t)np * natoms * NLM * 32 + (size_t)np * natoms * 16 +
             (size_t)np * (L + 1) * natoms * natoms * 8 + 256;
  if (natoms >= ctx->direct_gemm_min)
    b += (size_t)np * 2 * NLM * natoms * 8 + (size_t)np * dg_c2_off(L + 1) * 8;
  return b;
}

int64_t direct_chunk(const fo_ctx* ctx, int64_t npairs, int64_t natoms, int L, bool want_grid,
                     bool device_resident = false) {
  const size_t per = direct_work_bytes(ctx, 1, natoms, L) + ihalf_elems(L) * 16;
  // 4 GB of scratch per chunk (LJ38: ~7800 pairs).  Through host buffers, end of round 2 (1 / 2 / 3 / 4 / 6 GB):
  // 1.32 / 1.39 / 1.42 / 1.44 / 1.41 M aligned pairs/s -- with the kernels of the last session the per-chunk tails
  // weigh more than the copy / compute overlap gains from shorter chunks (the 2 GB of the first sessions: 0.987 / 1.006 /
  // 1.002 M at 1 / 2 / 4 GB); the device-resident path has no copies to overlap at all
  (void)device_resident;
  size_t budget = (size_t)4 << 30;
  if (const int64_t mb = ctx->opt("sph_chunk_mb")) budget = (size_t)mb << 20;  // tuning hook of the A/B scripts
  int64_t c = (int64_t)(budget / per);
  if (want_grid) {
    const size_t g = (size_t)8 * 2 * (2 * L + 2) * (2 * L + 2) * (2 * L + 2);
    int64_t cg = (int64_t)(((size_t)512 << 20) / g);
    if (cg < c) c = cg;
  }
  if (c > 65535) c = 65535;  // gridDim.y
  if (natoms >= ctx->direct_gemm_min && c * (L + 1) > 65535) c = 65535 / (L + 1);  // gridDim.z of the GEMM kernels
  if (c < 1) c = 1;
  if (c > npairs) c = npairs;
  return c;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------

extern "C" int fo_sph_wigner_table(fo_ctx* ctx, int64_t Jmax, double* out) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (!out) return fo_fail(ctx, FO_ERR_INVALID, "out is NULL");
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)Jmax;
  FO_CHECK(ensure_wigner(ctx, L));
  const size_t n = (size_t)(L + 1) * (2 * L + 1) * (2 * L + 1) * (2 * L + 2);
  void* d = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, n * 8, &d));
  sph_wigner_export_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(
      ctx->wig.d_table, dt_stride(L, ctx->wig.kmajor), L, (double*)d);
  FO_LAUNCH_CHECK(ctx);
  FO_CUDA(ctx, cudaMemcpyAsync(out, d, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FO_OK;
}

extern "C" int fo_sph_ylm(fo_ctx* ctx, const double* pos, int64_t nstruct, int64_t natoms, int64_t Jmax,
                          double* Y_out, double* r_out, int32_t* status) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (natoms < 1) return fo_fail(ctx, FO_ERR_INVALID, "natoms >= 1 required");
  if (nstruct < 0 || (nstruct > 0 && (!pos || !Y_out))) return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_ylm: NULL argument");
  if (nstruct == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)Jmax, NLM = nlm_of(L);
  const size_t full = (size_t)(L + 1) * (2 * L + 1) * natoms;
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full * 16 + (size_t)natoms * (NLM * 16 + 40)));
  if (chunk < 1) chunk = 1;
  if (chunk > nstruct) chunk = nstruct;
  void *dpos, *work, *dfull, *dst;
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSA, (size_t)chunk * natoms * 24, &dpos));
  FO_CHECK(fo_scratch(ctx, FO_SCR_WORK, (size_t)chunk * natoms * (NLM * 16 + 8) + 256, &work));
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * full * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_MISC, (size_t)chunk * 4 + 64, &dst));
  const size_t smem_prep = ((size_t)2 * NLM + L + 1) * 8;
  FO_CUDA(ctx, cudaFuncSetAttribute(sph_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
  for (int64_t s0 = 0; s0 < nstruct; s0 += chunk) {
    const int64_t ns = std::min(chunk, nstruct - s0);
    double2* Y = (double2*)work;
    double* R = (double*)(Y + (size_t)ns * natoms * NLM);
    FO_CUDA(ctx, cudaMemcpyAsync(dpos, pos + (size_t)s0 * natoms * 3, (size_t)ns * natoms * 24, cudaMemcpyHostToDevice, ctx->stream));
    FO_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)ns * 4, ctx->stream));
    sph_prep_kernel<<<grid_for((size_t)ns * natoms * (L + 1), 128, 148 * 16), 128, smem_prep, ctx->stream>>>(
        (const double*)dpos, (int)natoms, L, (size_t)ns, Y, R, (int*)dst);
    FO_LAUNCH_CHECK(ctx);
    sph_ylm_expand_kernel<<<grid_for((size_t)ns * full, 256), 256, 0, ctx->stream>>>(Y, (int)natoms, L, (size_t)ns,
                                                                                   (double2*)dfull);
    FO_LAUNCH_CHECK(ctx);
    FO_CUDA(ctx, cudaMemcpyAsync(Y_out + (size_t)s0 * full * 2, dfull, (size_t)ns * full * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (r_out)
      FO_CUDA(ctx, cudaMemcpyAsync(r_out + (size_t)s0 * natoms, R, (size_t)ns * natoms * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) FO_CUDA(ctx, cudaMemcpyAsync(status + s0, dst, (size_t)ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

extern "C" int fo_sph_isoft_argmax(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax,
                                   int invert, int64_t* best_idx, double* best_val, double* frac_idx,
                                   double* grid_out) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (npairs < 0 || (npairs > 0 && (!Ilmm || !best_idx || !best_val || !frac_idx)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_isoft_argmax: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)Jmax, O = invert ? 2 : 1;
  const size_t full = (size_t)(L + 1) * (2 * L + 1) * (2 * L + 1);
  const size_t G3 = (size_t)(2 * L + 2) * (2 * L + 2) * (2 * L + 2);
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full * 16 + ihalf_elems(L) * 16 + (grid_out ? O * G3 * 8 : 0)));
  if (chunk < 1) chunk = 1;
  if (chunk > npairs) chunk = npairs;
  void *dfull, *dhalf, *dout, *dgrid = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * full * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * ihalf_elems(L) * 16, &dhalf));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * O * 64, &dout));
  if (grid_out) FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * O * G3 * 8, &dgrid));
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    FO_CUDA(ctx, cudaMemcpyAsync(dfull, Ilmm + (size_t)p0 * full * 2, (size_t)np * full * 16,
                                 cudaMemcpyHostToDevice, ctx->stream));
    sph_pack_kernel<<<grid_for((size_t)np * ihalf_elems(L), 256), 256, 0, ctx->stream>>>(
        (const double2*)dfull, L, (size_t)np, 0, (double2*)dhalf);
    FO_LAUNCH_CHECK(ctx);
    long long* d_bi = (long long*)dout;
    double* d_bv = (double*)((char*)dout + (size_t)np * O * 24);
    double* d_fr = (double*)((char*)dout + (size_t)np * O * 32);
    FO_CHECK(run_isoft(ctx, (const double2*)dhalf, np, L, O, d_bi, d_bv, d_fr, (double*)dgrid));
    FO_CUDA(ctx, cudaMemcpyAsync(best_idx + (size_t)p0 * O * 3, d_bi, (size_t)np * O * 24, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(best_val + (size_t)p0 * O, d_bv, (size_t)np * O * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(frac_idx + (size_t)p0 * O * 3, d_fr, (size_t)np * O * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (grid_out)
      FO_CUDA(ctx, cudaMemcpyAsync(grid_out + (size_t)p0 * O * G3, dgrid, (size_t)np * O * G3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

extern "C" int fo_sph_isoft(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax, double* grid_re,
                            double* grid_im) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (npairs < 0 || (npairs > 0 && (!Ilmm || !grid_re)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_isoft: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)Jmax;
  const size_t full = (size_t)(L + 1) * (2 * L + 1) * (2 * L + 1);
  const size_t G3 = (size_t)(2 * L + 2) * (2 * L + 2) * (2 * L + 2);
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full * 16 + ihalf_elems(L) * 16 + G3 * 8));
  if (chunk < 1) chunk = 1;
  if (chunk > npairs) chunk = npairs;
  void *dfull, *dhalf, *dout, *dgrid;
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * full * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * ihalf_elems(L) * 16, &dhalf));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * 64, &dout));
  FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * G3 * 8, &dgrid));
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    FO_CUDA(ctx, cudaMemcpyAsync(dfull, Ilmm + (size_t)p0 * full * 2, (size_t)np * full * 16,
                                 cudaMemcpyHostToDevice, ctx->stream));
    for (int part = 0; part < (grid_im ? 2 : 1); ++part) {
      sph_pack_kernel<<<grid_for((size_t)np * ihalf_elems(L), 256), 256, 0, ctx->stream>>>(
          (const double2*)dfull, L, (size_t)np, part, (double2*)dhalf);
      FO_LAUNCH_CHECK(ctx);
      FO_CHECK(run_isoft(ctx, (const double2*)dhalf, np, L, 1, (long long*)dout,
                         (double*)((char*)dout + (size_t)np * 24), (double*)((char*)dout + (size_t)np * 32),
                         (double*)dgrid));
      FO_CUDA(ctx, cudaMemcpyAsync((part ? grid_im : grid_re) + (size_t)p0 * G3, dgrid, (size_t)np * G3 * 8,
                                   cudaMemcpyDeviceToHost, ctx->stream));
      FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  return FO_OK;
}

// a8 fused: coefficients -> (2L+2)^3 overlap grid on the device -> top-npeaks rotations (fractional
// grid indices; indtoEuler on the host).  findRotations, sphericalAlignment.py:196-204.
extern "C" int fo_sph_isoft_peaks(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax, int64_t npeaks,
                                  int64_t width, double* peaks, double* amplitude, double* mean, double* alpha,
                                  int32_t* nfound) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (npairs < 0 || (npairs > 0 && (!Ilmm || !peaks || !amplitude || !nfound)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_isoft_peaks: NULL argument");
  if (npeaks < 1 || npeaks > 64 || width < 1 || width > 4)
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_isoft_peaks: npeaks in 1..64, width in 1..4");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)Jmax;
  const size_t full = (size_t)(L + 1) * (2 * L + 1) * (2 * L + 1);
  const int64_t shape[3] = {2 * L + 2, 2 * L + 2, 2 * L + 2};
  const size_t G3 = (size_t)shape[0] * shape[1] * shape[2];
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full * 16 + ihalf_elems(L) * 16 + G3 * 8));
  if (chunk < 1) chunk = 1;
  if (chunk > npairs) chunk = npairs;
  void *dfull, *dhalf, *dout, *dgrid;
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * full * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * ihalf_elems(L) * 16, &dhalf));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * 64, &dout));
  FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * G3 * 8, &dgrid));
  double *pk, *amp, *mn, *al;
  int32_t* nf;
  FO_CHECK(fo_peaks_outputs(ctx, chunk, npeaks, &pk, &amp, &mn, &al, &nf));
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    FO_CUDA(ctx, cudaMemcpyAsync(dfull, Ilmm + (size_t)p0 * full * 2, (size_t)np * full * 16,
                                 cudaMemcpyHostToDevice, ctx->stream));
    sph_pack_kernel<<<grid_for((size_t)np * ihalf_elems(L), 256), 256, 0, ctx->stream>>>(
        (const double2*)dfull, L, (size_t)np, 0, (double2*)dhalf);
    FO_LAUNCH_CHECK(ctx);
    FO_CHECK(run_isoft(ctx, (const double2*)dhalf, np, L, 1, (long long*)dout,
                       (double*)((char*)dout + (size_t)np * 24), (double*)((char*)dout + (size_t)np * 32),
                       (double*)dgrid));
    FO_CHECK(fo_peaks_run_dev(ctx, (double*)dgrid, np, shape, npeaks, width, pk, amp, mn, al, nf));
    FO_CHECK(fo_peaks_copy_out(ctx, p0, np, npeaks, pk, amp, mn, al, nf, peaks, amplitude, mean, alpha, nfound));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

extern "C" int fo_sph_coeffs_direct(fo_ctx* ctx, const double* posA, const double* posB, int64_t npairs,
                                    int64_t natoms, int64_t Jmax, double sigma, double* Ilmm_out,
                                    int32_t* status) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (natoms < 1 || !(sigma > 0.0)) return fo_fail(ctx, FO_ERR_INVALID, "natoms >= 1 and sigma > 0 required");
  if (npairs < 0 || (npairs > 0 && (!posA || !posB || !Ilmm_out)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_coeffs_direct: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, natoms));
  const int L = (int)Jmax;
  int* d_gid = nullptr;
  FO_CHECK(upload_gid(ctx, natoms, &d_gid));
  const size_t full = (size_t)(L + 1) * (2 * L + 1) * (2 * L + 1);
  int64_t chunk = direct_chunk(ctx, npairs, natoms, L, false);
  void *dA, *dB, *work, *dhalf, *dfull, *dst;
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSA, (size_t)chunk * natoms * 24, &dA));
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSB, (size_t)chunk * natoms * 24, &dB));
  FO_CHECK(fo_scratch(ctx, FO_SCR_WORK, direct_work_bytes(ctx, chunk, natoms, L), &work));
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * ihalf_elems(L) * 16, &dhalf));
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * full * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_MISC, (size_t)chunk * 4 + 64, &dst));
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    FO_CUDA(ctx, cudaMemcpyAsync(dA, posA + (size_t)p0 * natoms * 3, (size_t)np * natoms * 24, cudaMemcpyHostToDevice, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(dB, posB + (size_t)p0 * natoms * 3, (size_t)np * natoms * 24, cudaMemcpyHostToDevice, ctx->stream));
    FO_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)np * 4, ctx->stream));
    FO_CHECK(run_direct(ctx, (const double*)dA, (const double*)dB, np, natoms, L, sigma, d_gid,
                        (double2*)dhalf, (int*)dst, (char*)work));
    sph_unpack_kernel<<<grid_for((size_t)np * full, 256), 256, 0, ctx->stream>>>((const double2*)dhalf, L, (size_t)np, (double2*)dfull);
    FO_LAUNCH_CHECK(ctx);
    FO_CUDA(ctx, cudaMemcpyAsync(Ilmm_out + (size_t)p0 * full * 2, dfull, (size_t)np * full * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (status)
      FO_CUDA(ctx, cudaMemcpyAsync(status + p0, dst, (size_t)np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

namespace {
// d_euler [P,O,3] / d_overlap [P,O] non-null: continuous refinement of both orientations (fo_refine.cu)
int align_pairs_dev_impl(fo_ctx* ctx, const double* d_posA, const double* d_posB, int64_t npairs, int64_t natoms,
                         int64_t Jmax, double sigma, int invert, int64_t* d_best_idx, double* d_best_val,
                         double* d_frac_idx, double* d_grid_out, int32_t* d_status, double* d_euler,
                         double* d_overlap) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (natoms < 1 || !(sigma > 0.0)) return fo_fail(ctx, FO_ERR_INVALID, "natoms >= 1 and sigma > 0 required");
  if (npairs < 0 || (npairs > 0 && (!d_posA || !d_posB || !d_best_idx || !d_best_val || !d_frac_idx)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_align_pairs_dev: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, natoms));
  const int L = (int)Jmax, O = invert ? 2 : 1;
  int* d_gid = nullptr;
  FO_CHECK(upload_gid(ctx, natoms, &d_gid));
  const size_t G3 = (size_t)(2 * L + 2) * (2 * L + 2) * (2 * L + 2);
  const int64_t chunk = direct_chunk(ctx, npairs, natoms, L, false, true);
  void *work, *dhalf;
  FO_CHECK(fo_scratch(ctx, FO_SCR_WORK, direct_work_bytes(ctx, chunk, natoms, L), &work));
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * ihalf_elems(L) * 16, &dhalf));
  if (d_status) FO_CUDA(ctx, cudaMemsetAsync(d_status, 0, (size_t)npairs * 4, ctx->stream));
  // the streaming coefficient kernel also writes the packed copy the fast iSOFT kernels read (no sph_ipack pass)
  FO_CHECK(ensure_wigner(ctx, L));
  const bool fuse_ipk = isoft_packed_applies(ctx, L) && direct2_slots(ctx, natoms, L) > 0;
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    void* ipk = nullptr;
    if (fuse_ipk) FO_CHECK(fo_scratch(ctx, FO_SCR_IPK, (size_t)np * I2Layout(L, ctx->wig.packed_kc).ipk * 16, &ipk));
    FO_CHECK(run_direct(ctx, d_posA + (size_t)p0 * natoms * 3, d_posB + (size_t)p0 * natoms * 3, np, natoms,
                        L, sigma, d_gid, (double2*)dhalf, d_status ? d_status + p0 : nullptr, (char*)work,
                        (double2*)ipk));
    FO_CHECK(run_isoft(ctx, (const double2*)dhalf, np, L, O, (long long*)d_best_idx + (size_t)p0 * O * 3,
                       d_best_val + (size_t)p0 * O, d_frac_idx + (size_t)p0 * O * 3,
                       d_grid_out ? d_grid_out + (size_t)p0 * O * G3 : nullptr, ipk != nullptr));
    if (d_euler)
      FO_CHECK(fo_refine_run_dev(ctx, dhalf, np, L, O, d_frac_idx + (size_t)p0 * O * 3, 1,
                                 d_euler + (size_t)p0 * O * 3, d_overlap + (size_t)p0 * O, nullptr));
  }
  return FO_OK;
}
}  // namespace

extern "C" int fo_sph_align_pairs_dev(fo_ctx* ctx, const double* d_posA, const double* d_posB,
                                      int64_t npairs, int64_t natoms, int64_t Jmax, double sigma,
                                      int invert, int64_t* d_best_idx, double* d_best_val,
                                      double* d_frac_idx, double* d_grid_out, int32_t* d_status) {
  return align_pairs_dev_impl(ctx, d_posA, d_posB, npairs, natoms, Jmax, sigma, invert, d_best_idx, d_best_val,
                              d_frac_idx, d_grid_out, d_status, nullptr, nullptr);
}

extern "C" int fo_sph_align_pairs_screen_dev(fo_ctx* ctx, const double* d_posA, const double* d_posB,
                                             int64_t npairs, int64_t natoms, int64_t Jmax, double sigma,
                                             int invert, int64_t* d_best_idx, double* d_best_val,
                                             double* d_frac_idx, int32_t* d_perm, int32_t* d_ok,
                                             int32_t* d_status) {
  if (ctx && npairs > 0 && (!d_perm || !d_ok))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_align_pairs_screen_dev: NULL argument");
  FO_CHECK(align_pairs_dev_impl(ctx, d_posA, d_posB, npairs, natoms, Jmax, sigma, invert, d_best_idx, d_best_val,
                                d_frac_idx, nullptr, d_status, nullptr, nullptr));
  return fo_sph_assign_run_dev(ctx, d_posA, d_posB, d_frac_idx, npairs, natoms, (int)Jmax, invert ? 2 : 1, d_perm,
                               d_ok);
}

extern "C" int fo_sph_align_pairs_refined_dev(fo_ctx* ctx, const double* d_posA, const double* d_posB,
                                              int64_t npairs, int64_t natoms, int64_t Jmax, double sigma,
                                              int invert, int64_t* d_best_idx, double* d_best_val,
                                              double* d_frac_idx, double* d_euler, double* d_overlap,
                                              int32_t* d_status) {
  if (ctx && npairs > 0 && (!d_euler || !d_overlap))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_align_pairs_refined_dev: NULL argument");
  return align_pairs_dev_impl(ctx, d_posA, d_posB, npairs, natoms, Jmax, sigma, invert, d_best_idx, d_best_val,
                              d_frac_idx, nullptr, d_status, d_euler, d_overlap);
}

namespace {
// Outputs of the full alignment (fo_sph_align_pairs_full): the device screening proposes a permutation per
// (pair, orientation); the host pool runs the Kearsley fit (and the LAP where the screening failed).
struct SphFull {
  int nthreads;
  double* dist;        // [P]
  int32_t* orient;     // [P] or null
  int32_t* perm;       // [P,N] or null
  double* rmat;        // [P,9] or null
  double* euler_grid;  // [P,O,3] or null: Euler angles of the interpolated grid maxima
  int64_t nhost = 0;   // (pair, orientation) assignments the host LAP solved
};

int align_pairs_impl(fo_ctx* ctx, const double* posA, const double* posB, int64_t npairs, int64_t natoms,
                     int64_t Jmax, double sigma, int invert, int64_t* best_idx, double* best_val,
                     double* frac_idx, double* grid_out, int32_t* status, double* euler, double* overlap,
                     SphFull* full = nullptr) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (natoms < 1 || !(sigma > 0.0)) return fo_fail(ctx, FO_ERR_INVALID, "natoms >= 1 and sigma > 0 required");
  if (npairs < 0 || (npairs > 0 && (!posA || !posB || !best_idx || !best_val || !frac_idx)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_align_pairs: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)Jmax, O = invert ? 2 : 1;
  const size_t G3 = (size_t)(2 * L + 2) * (2 * L + 2) * (2 * L + 2);
  const int64_t chunk = direct_chunk(ctx, npairs, natoms, L, grid_out != nullptr);
  const size_t pos_bytes = (size_t)chunk * natoms * 24;
  void *dA, *dB, *dout, *dgrid = nullptr, *hA, *hB;
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSA, 2 * pos_bytes, &dA));
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSB, 2 * pos_bytes, &dB));
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * (O * 88 + 8) + 256, &dout));
  if (grid_out) FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * O * G3 * 8, &dgrid));
  // full alignment, per chunk: kdist[np O] f64 | krot[np O 9] f64 | ok[np O] i32 | perm[np O N] i32
  void *dfull = nullptr, *hFull = nullptr;
  const size_t full_stride = (size_t)O * (80 + 4 * (1 + natoms));
  if (full) {
    FO_CHECK(fo_ensure_perm(ctx, natoms));
    FO_CHECK(fo_scratch(ctx, FO_SCR_FULL, (size_t)chunk * full_stride, &dfull));
    FO_CHECK(fo_pinned(ctx, 3, 2 * (size_t)chunk * full_stride, &hFull));
  }
  std::vector<double> eul;
  const bool pinnedA = fo_is_pinned(posA), pinnedB = fo_is_pinned(posB);
  hA = hB = nullptr;
  if (!pinnedA) FO_CHECK(fo_pinned(ctx, 0, 2 * pos_bytes, &hA));
  if (!pinnedB) FO_CHECK(fo_pinned(ctx, 1, 2 * pos_bytes, &hB));
  const std::vector<int64_t> starts = fo_chunk_starts(npairs, chunk);  // short first chunk
  const int64_t nchunks = (int64_t)starts.size() - 1;
  auto stage_in = [&](int64_t c) -> int {
    const int64_t p0 = starts[c];
    const int64_t np = starts[c + 1] - p0;
    const int buf = (int)(c & 1);
    const size_t nb = (size_t)np * natoms * 24;
    if (c >= 2) FO_CUDA(ctx, cudaEventSynchronize(ctx->ev[buf]));
    const char* srcA = (const char*)(posA + (size_t)p0 * natoms * 3);
    const char* srcB = (const char*)(posB + (size_t)p0 * natoms * 3);
    if (!pinnedA) {
      fo_host_copy((char*)hA + buf * pos_bytes, srcA, nb);
      srcA = (const char*)hA + buf * pos_bytes;
    }
    if (!pinnedB) {
      fo_host_copy((char*)hB + buf * pos_bytes, srcB, nb);
      srcB = (const char*)hB + buf * pos_bytes;
    }
    if (c >= 2) FO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[2 + buf], 0));
    FO_CUDA(ctx, cudaMemcpyAsync((char*)dA + buf * pos_bytes, srcA, nb, cudaMemcpyHostToDevice, ctx->copy_stream));
    FO_CUDA(ctx, cudaMemcpyAsync((char*)dB + buf * pos_bytes, srcB, nb, cudaMemcpyHostToDevice, ctx->copy_stream));
    FO_CUDA(ctx, cudaEventRecord(ctx->ev[buf], ctx->copy_stream));
    return FO_OK;
  };
  // results: one D2H of the chunk's output block into a pinned ring, unpacked into the caller's arrays
  // one chunk later (a D2H into pageable arrays would block the host and idle the GPU between chunks)
  void* hOut = nullptr;
  const size_t out_bytes = (size_t)chunk * (O * 88 + 8);
  if (!grid_out) FO_CHECK(fo_pinned(ctx, 2, 2 * out_bytes, &hOut));
  auto deliver = [&](int64_t c) -> int {
    const int64_t p0 = starts[c], np = starts[c + 1] - p0;
    const char* src = (const char*)hOut + (c & 1) * out_bytes;
    memcpy(best_idx + (size_t)p0 * O * 3, src, (size_t)np * O * 24);
    memcpy(best_val + (size_t)p0 * O, src + (size_t)np * O * 24, (size_t)np * O * 8);
    memcpy(frac_idx + (size_t)p0 * O * 3, src + (size_t)np * O * 32, (size_t)np * O * 24);
    if (euler) {
      memcpy(euler + (size_t)p0 * O * 3, src + (size_t)np * O * 56, (size_t)np * O * 24);
      memcpy(overlap + (size_t)p0 * O, src + (size_t)np * O * 80, (size_t)np * O * 8);
    }
    if (status) memcpy(status + p0, src + (size_t)np * O * 88, (size_t)np * 4);
    if (!full) return FO_OK;
    // host stage of the full alignment, while the GPU works on the next chunk: Euler angles of the grid
    // maxima (indtoEuler, utils.py:340-345), then rotate + permutation (device hint or LAP) + Kearsley
    const double kPi = 3.14159265358979323846, n2 = (double)(2 * (L + 1));
    eul.resize((size_t)np * O * 3);
    const double* fr = frac_idx + (size_t)p0 * O * 3;
    for (int64_t i = 0; i < np * O; ++i) {
      eul[3 * i] = (2 * kPi / n2) * fr[3 * i];
      eul[3 * i + 1] = (kPi / n2) * fr[3 * i + 1] + 0.5 * kPi / n2;
      eul[3 * i + 2] = (2 * kPi / n2) * fr[3 * i + 2];
    }
    if (full->euler_grid) memcpy(full->euler_grid + (size_t)p0 * O * 3, eul.data(), (size_t)np * O * 24);
    const char* fsrc = (const char*)hFull + (c & 1) * (size_t)chunk * full_stride;
    const double* kd = (const double*)fsrc;
    const double* kR = kd + (size_t)np * O;
    const int32_t* ok = (const int32_t*)(kR + (size_t)np * O * 9);
    const int32_t* hint = ok + (size_t)np * O;
    // pairs whose orientations were all settled on the device (assignment proven optimal, Kearsley fit done there):
    // pick the orientation with the smaller distance, first one on ties (the host loop's rule), and copy; the
    // others go through the host pool (LAP where the screening failed + Kearsley)
    std::vector<int64_t> hard;
    for (int64_t q = 0; q < np; ++q) {
      bool all = true;
      for (int o = 0; o < O; ++o) all = all && ok[q * O + o];
      if (!all) {
        for (int o = 0; o < O; ++o) full->nhost += ok[q * O + o] ? 0 : 1;
        hard.push_back(q);
        continue;
      }
      int bo = 0;
      for (int o = 1; o < O; ++o)
        if (kd[q * O + o] < kd[q * O + bo]) bo = o;
      full->dist[p0 + q] = kd[q * O + bo];
      if (full->orient) full->orient[p0 + q] = bo;
      if (full->perm) memcpy(full->perm + (size_t)(p0 + q) * natoms, hint + ((size_t)q * O + bo) * natoms, (size_t)natoms * 4);
      if (full->rmat) memcpy(full->rmat + 9 * (p0 + q), kR + ((size_t)q * O + bo) * 9, 72);
    }
    int rc = FO_OK;
    if (!hard.empty())
      rc = fo_host_refine_spherical_subset(
          posA + (size_t)p0 * natoms * 3, posB + (size_t)p0 * natoms * 3, natoms, ctx->h_goff.data(),
          (int64_t)ctx->h_goff.size() - 1, ctx->h_gidx.data(), eul.data(), O, hint, ok, hard.data(), (int64_t)hard.size(),
          full->nthreads, full->dist + p0, full->orient ? full->orient + p0 : nullptr,
          full->perm ? full->perm + (size_t)p0 * natoms : nullptr, full->rmat ? full->rmat + 9 * p0 : nullptr, kd, kR);
    if (rc != FO_OK) return fo_fail(ctx, rc, "host refinement of chunk %lld failed", (long long)c);
    return FO_OK;
  };
  FO_CHECK(stage_in(0));
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t p0 = starts[c];
    const int64_t np = starts[c + 1] - p0;
    const int buf = (int)(c & 1);
    FO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev[buf], 0));
    long long* d_bi = (long long*)dout;
    double* d_bv = (double*)((char*)dout + (size_t)np * O * 24);
    double* d_fr = (double*)((char*)dout + (size_t)np * O * 32);
    double* d_eu = (double*)((char*)dout + (size_t)np * O * 56);
    double* d_ov = (double*)((char*)dout + (size_t)np * O * 80);
    int* d_st = (int*)((char*)dout + (size_t)np * O * 88);
    FO_CHECK(align_pairs_dev_impl(ctx, (const double*)((char*)dA + buf * pos_bytes),
                                  (const double*)((char*)dB + buf * pos_bytes), np, natoms, Jmax, sigma,
                                  invert, (int64_t*)d_bi, d_bv, d_fr, (double*)dgrid, d_st,
                                  euler ? d_eu : nullptr, euler ? d_ov : nullptr));
    if (full)  // nearest-partner screening of both orientations on the device (reads the positions again)
      FO_CHECK(fo_sph_assign_run_dev(ctx, (const double*)((char*)dA + buf * pos_bytes),
                                     (const double*)((char*)dB + buf * pos_bytes), d_fr, np, natoms, L, O,
                                     (int32_t*)((double*)dfull + (size_t)np * O * 10) + (size_t)np * O,
                                     (int32_t*)((double*)dfull + (size_t)np * O * 10), (double*)dfull,
                                     (double*)dfull + (size_t)np * O));
    FO_CUDA(ctx, cudaEventRecord(ctx->ev[2 + buf], ctx->stream));
    if (grid_out) {  // test / single-pair path: straight into the caller's arrays
      if (euler) {
        FO_CUDA(ctx, cudaMemcpyAsync(euler + (size_t)p0 * O * 3, d_eu, (size_t)np * O * 24, cudaMemcpyDeviceToHost, ctx->stream));
        FO_CUDA(ctx, cudaMemcpyAsync(overlap + (size_t)p0 * O, d_ov, (size_t)np * O * 8, cudaMemcpyDeviceToHost, ctx->stream));
      }
      FO_CUDA(ctx, cudaMemcpyAsync(best_idx + (size_t)p0 * O * 3, d_bi, (size_t)np * O * 24, cudaMemcpyDeviceToHost, ctx->stream));
      FO_CUDA(ctx, cudaMemcpyAsync(best_val + (size_t)p0 * O, d_bv, (size_t)np * O * 8, cudaMemcpyDeviceToHost, ctx->stream));
      FO_CUDA(ctx, cudaMemcpyAsync(frac_idx + (size_t)p0 * O * 3, d_fr, (size_t)np * O * 24, cudaMemcpyDeviceToHost, ctx->stream));
      if (status)
        FO_CUDA(ctx, cudaMemcpyAsync(status + p0, d_st, (size_t)np * 4, cudaMemcpyDeviceToHost, ctx->stream));
      FO_CUDA(ctx, cudaMemcpyAsync(grid_out + (size_t)p0 * O * G3, dgrid, (size_t)np * O * G3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
      if (c + 1 < nchunks) FO_CHECK(stage_in(c + 1));
      continue;
    }
    FO_CUDA(ctx, cudaMemcpyAsync((char*)hOut + buf * out_bytes, dout, (size_t)np * (O * 88 + 4), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    if (full)
      FO_CUDA(ctx, cudaMemcpyAsync((char*)hFull + buf * (size_t)chunk * full_stride, dfull, (size_t)np * full_stride,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaEventRecord(ctx->ev[4 + buf], ctx->stream));
    if (c + 1 < nchunks) FO_CHECK(stage_in(c + 1));  // host staging runs while the GPU works on chunk c
    if (c >= 1) {
      FO_CUDA(ctx, cudaEventSynchronize(ctx->ev[4 + (buf ^ 1)]));
      FO_CHECK(deliver(c - 1));
    }
  }
  FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (!grid_out) FO_CHECK(deliver(nchunks - 1));
  return FO_OK;
}
}  // namespace

extern "C" int fo_sph_align_pairs_full(fo_ctx* ctx, const double* posA, const double* posB, int64_t npairs,
                                       int64_t natoms, int64_t Jmax, double sigma, int invert, int nthreads,
                                       double* dist, int32_t* orient, int32_t* perm, double* rmat, double* euler,
                                       int32_t* status, int64_t* nhost) {
  if (!ctx) return FO_ERR_INVALID;
  if (npairs > 0 && !dist) return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_align_pairs_full: dist is NULL");
  if (npairs <= 0) {
    if (nhost) *nhost = 0;
    return npairs < 0 ? fo_fail(ctx, FO_ERR_INVALID, "npairs < 0") : FO_OK;
  }
  const int O = invert ? 2 : 1;
  std::vector<int64_t> bi((size_t)npairs * O * 3);
  std::vector<double> bv((size_t)npairs * O), fr((size_t)npairs * O * 3);
  SphFull full;
  full.nthreads = nthreads;
  full.dist = dist;
  full.orient = orient;
  full.perm = perm;
  full.rmat = rmat;
  full.euler_grid = euler;
  const int rc = align_pairs_impl(ctx, posA, posB, npairs, natoms, Jmax, sigma, invert, bi.data(), bv.data(),
                                  fr.data(), nullptr, status, nullptr, nullptr, &full);
  if (nhost) *nhost = full.nhost;
  return rc;
}

extern "C" int fo_sph_align_pairs(fo_ctx* ctx, const double* posA, const double* posB, int64_t npairs,
                                  int64_t natoms, int64_t Jmax, double sigma, int invert,
                                  int64_t* best_idx, double* best_val, double* frac_idx, double* grid_out,
                                  int32_t* status) {
  return align_pairs_impl(ctx, posA, posB, npairs, natoms, Jmax, sigma, invert, best_idx, best_val, frac_idx,
                          grid_out, status, nullptr, nullptr);
}

extern "C" int fo_sph_align_pairs_refined(fo_ctx* ctx, const double* posA, const double* posB, int64_t npairs,
                                          int64_t natoms, int64_t Jmax, double sigma, int invert,
                                          int64_t* best_idx, double* best_val, double* frac_idx, double* euler,
                                          double* overlap, int32_t* status) {
  if (ctx && npairs > 0 && (!euler || !overlap))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_align_pairs_refined: NULL argument");
  return align_pairs_impl(ctx, posA, posB, npairs, natoms, Jmax, sigma, invert, best_idx, best_val, frac_idx,
                          nullptr, status, euler, overlap);
}

// Continuous refinement of P rotations from caller-supplied coefficients (maxOverlap,
// sphericalAlignment.py:98-103): Ilmm as fo_sph_isoft_argmax, euler_in [P,3] the starting angles.
extern "C" int fo_sph_refine_rotations(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax,
                                       const double* euler_in, double* euler_out, double* overlap_out,
                                       int32_t* nevals_out) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (npairs < 0 || (npairs > 0 && (!Ilmm || !euler_in || !euler_out || !overlap_out)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_refine_rotations: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)Jmax;
  const size_t full = (size_t)(L + 1) * (2 * L + 1) * (2 * L + 1);
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full * 16 + ihalf_elems(L) * 16));
  if (chunk < 1) chunk = 1;
  if (chunk > npairs) chunk = npairs;
  void *dfull, *dhalf, *dout;
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * full * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * ihalf_elems(L) * 16, &dhalf));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * 64, &dout));
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    double* d_in = (double*)dout;
    double* d_eu = d_in + (size_t)np * 3;
    double* d_ov = d_eu + (size_t)np * 3;
    int* d_ne = (int*)(d_ov + np);
    FO_CUDA(ctx, cudaMemcpyAsync(dfull, Ilmm + (size_t)p0 * full * 2, (size_t)np * full * 16,
                                 cudaMemcpyHostToDevice, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(d_in, euler_in + (size_t)p0 * 3, (size_t)np * 24, cudaMemcpyHostToDevice, ctx->stream));
    sph_pack_kernel<<<grid_for((size_t)np * ihalf_elems(L), 256), 256, 0, ctx->stream>>>(
        (const double2*)dfull, L, (size_t)np, 0, (double2*)dhalf);
    FO_LAUNCH_CHECK(ctx);
    FO_CHECK(fo_refine_run_dev(ctx, dhalf, np, L, 1, d_in, 0, d_eu, d_ov, d_ne));
    FO_CUDA(ctx, cudaMemcpyAsync(euler_out + (size_t)p0 * 3, d_eu, (size_t)np * 24, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(overlap_out + p0, d_ov, (size_t)np * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nevals_out)
      FO_CUDA(ctx, cudaMemcpyAsync(nevals_out + p0, d_ne, (size_t)np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

// One evaluation of the refinement objective and its derivatives (getEnergyGradient,
// sphericalAlignment.py:93-96, with the sign of the overlap: value = -E, grad = -dE).
extern "C" int fo_sph_overlap_gradient(fo_ctx* ctx, const double* Ilmm, int64_t npairs, int64_t Jmax,
                                       const double* euler, double* value, double* grad, double* hess) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (npairs < 0 || (npairs > 0 && (!Ilmm || !euler || !value || !grad)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_overlap_gradient: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)Jmax;
  const size_t full = (size_t)(L + 1) * (2 * L + 1) * (2 * L + 1);
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full * 16 + ihalf_elems(L) * 16));
  if (chunk < 1) chunk = 1;
  if (chunk > npairs) chunk = npairs;
  void *dfull, *dhalf, *dout;
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * full * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * ihalf_elems(L) * 16, &dhalf));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * 104, &dout));
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    double* d_in = (double*)dout;
    double* d_v = d_in + (size_t)np * 3;
    double* d_g = d_v + np;
    double* d_h = d_g + (size_t)np * 3;
    FO_CUDA(ctx, cudaMemcpyAsync(dfull, Ilmm + (size_t)p0 * full * 2, (size_t)np * full * 16,
                                 cudaMemcpyHostToDevice, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(d_in, euler + (size_t)p0 * 3, (size_t)np * 24, cudaMemcpyHostToDevice, ctx->stream));
    sph_pack_kernel<<<grid_for((size_t)np * ihalf_elems(L), 256), 256, 0, ctx->stream>>>(
        (const double2*)dfull, L, (size_t)np, 0, (double2*)dhalf);
    FO_LAUNCH_CHECK(ctx);
    FO_CHECK(fo_refine_eval_dev(ctx, dhalf, np, L, d_in, d_v, d_g, d_h));
    FO_CUDA(ctx, cudaMemcpyAsync(value + p0, d_v, (size_t)np * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(grad + (size_t)p0 * 3, d_g, (size_t)np * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (hess)
      FO_CUDA(ctx, cudaMemcpyAsync(hess + (size_t)p0 * 6, d_h, (size_t)np * 48, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

namespace {
// harmonic coefficients of nstruct structures (device positions) into d_Cpk
int run_harm(fo_ctx* ctx, const double* d_pos, int64_t ns, int64_t natoms, int nmax, int L, double r0,
             double sigma, double2* d_Cpk, int* d_status, char* work) {
  if (ns == 0) return FO_OK;
  const int NLM = nlm_of(L);
  const int ng = (int)ctx->h_goff.size() - 1;
  if ((nmax + 1) * NLM > 16 * 256)
    return fo_fail(ctx, FO_ERR_UNSUPPORTED, "(nmax+1)*(Jmax+1)(Jmax+2)/2 = %d exceeds 4096",
                   (nmax + 1) * NLM);
  double2* Y = (double2*)work;
  double* R = (double*)(Y + (size_t)ns * natoms * NLM);
  fo_prof_scope prof(ctx, FO_PROF_SPH_HARM);
  const size_t smem_prep = ((size_t)2 * NLM + L + 1) * 8;
  FO_CUDA(ctx, cudaFuncSetAttribute(sph_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
  sph_prep_kernel<<<grid_for((size_t)ns * natoms * (L + 1), 128, 148 * 16), 128, smem_prep, ctx->stream>>>(
      d_pos, (int)natoms, L, (size_t)ns, Y, R, d_status);
  FO_LAUNCH_CHECK(ctx);
  const size_t smem = (size_t)HARM_TA * (nmax + 1) * (L + 1) * 8 + (size_t)HARM_TA * NLM * 16;
  FO_CUDA(ctx, cudaFuncSetAttribute(sph_harm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ns, (unsigned)ng);
  sph_harm_kernel<<<grid, 256, smem, ctx->stream>>>(Y, R, ctx->d_goff, ctx->d_gidx, ng, (int)natoms, nmax, L,
                                                    r0, sigma, d_Cpk);
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}
}  // namespace

extern "C" int fo_sph_harm_coeffs(fo_ctx* ctx, const double* pos, int64_t nstruct, int64_t natoms,
                                  int64_t nmax, int64_t Jmax, double harmscale, double sigma, double* out,
                                  int32_t* status) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (natoms < 1 || nmax < 0 || nmax > 255 || !(sigma > 0.0) || !(harmscale > 0.0))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_harm_coeffs: bad size / scale");
  if (nstruct < 0 || (nstruct > 0 && (!pos || !out))) return fo_fail(ctx, FO_ERR_INVALID, "NULL argument");
  if (nstruct == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, natoms));
  const int L = (int)Jmax, NLM = nlm_of(L);
  const int ng = (int)ctx->h_goff.size() - 1;
  const size_t cpk = (size_t)ng * (nmax + 1) * NLM, full = (size_t)ng * (nmax + 1) * (L + 1) * (2 * L + 1);
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full * 16 + cpk * 16 + (size_t)natoms * NLM * 16 + 64));
  if (chunk < 1) chunk = 1;
  if (chunk > nstruct) chunk = nstruct;
  void *dpos, *work, *dc, *dfull, *dst;
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSA, (size_t)chunk * natoms * 24, &dpos));
  FO_CHECK(fo_scratch(ctx, FO_SCR_WORK, (size_t)chunk * natoms * (NLM * 16 + 8) + 256, &work));
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * cpk * 16, &dc));
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * full * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_MISC, (size_t)chunk * 4 + 64, &dst));
  for (int64_t s0 = 0; s0 < nstruct; s0 += chunk) {
    const int64_t ns = std::min(chunk, nstruct - s0);
    FO_CUDA(ctx, cudaMemcpyAsync(dpos, pos + (size_t)s0 * natoms * 3, (size_t)ns * natoms * 24, cudaMemcpyHostToDevice, ctx->stream));
    FO_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)ns * 4, ctx->stream));
    FO_CHECK(run_harm(ctx, (const double*)dpos, ns, natoms, (int)nmax, L, harmscale, sigma, (double2*)dc, (int*)dst, (char*)work));
    sph_harm_expand_kernel<<<grid_for((size_t)ns * full, 256), 256, 0, ctx->stream>>>(
        (const double2*)dc, (int)nmax, L, (size_t)ns * ng, (double2*)dfull);
    FO_LAUNCH_CHECK(ctx);
    FO_CUDA(ctx, cudaMemcpyAsync(out + (size_t)s0 * full * 2, dfull, (size_t)ns * full * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) FO_CUDA(ctx, cudaMemcpyAsync(status + s0, dst, (size_t)ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

extern "C" int fo_sph_bank_create(fo_ctx* ctx, const double* pos, int64_t nstruct, int64_t natoms,
                                  int64_t nmax, int64_t Jmax, double harmscale, double sigma, fo_bank** out) {
  if (!ctx) return FO_ERR_INVALID;
  FO_CHECK(check_L(ctx, Jmax));
  if (!out || !pos || nstruct < 1 || natoms < 1 || nmax < 0 || nmax > 255 || !(sigma > 0.0) || !(harmscale > 0.0))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_bank_create: bad argument");
  *out = nullptr;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, natoms));
  const int L = (int)Jmax, NLM = nlm_of(L);
  const int ng = (int)ctx->h_goff.size() - 1;
  fo_bank* b = new (std::nothrow) fo_bank();
  if (!b) return fo_fail(ctx, FO_ERR_NOMEM, "out of host memory");
  b->kind = 2;
  b->nstruct = nstruct;
  b->ngroups = ng;
  b->nmax = nmax;
  b->Jmax = Jmax;
  b->natoms = natoms;
  b->sigma = sigma;
  b->harmscale = harmscale;
  b->per_struct_elems = (int64_t)ng * (nmax + 1) * NLM;
  if (cudaMalloc(&b->d_data, (size_t)nstruct * b->per_struct_elems * 16) != cudaSuccess) {
    delete b;
    return fo_fail(ctx, FO_ERR_NOMEM, "cudaMalloc of the coefficient bank failed");
  }
  int64_t chunk = (int64_t)(((size_t)256 << 20) / ((size_t)natoms * (NLM * 16 + 32)));
  if (chunk < 1) chunk = 1;
  if (chunk > nstruct) chunk = nstruct;
  void *dpos = nullptr, *work = nullptr;
  int rc = fo_scratch(ctx, FO_SCR_POSA, (size_t)chunk * natoms * 24, &dpos);
  if (rc == FO_OK) rc = fo_scratch(ctx, FO_SCR_WORK, (size_t)chunk * natoms * (NLM * 16 + 8) + 256, &work);
  for (int64_t s0 = 0; rc == FO_OK && s0 < nstruct; s0 += chunk) {
    const int64_t ns = std::min(chunk, nstruct - s0);
    if (cudaMemcpyAsync(dpos, pos + (size_t)s0 * natoms * 3, (size_t)ns * natoms * 24, cudaMemcpyHostToDevice,
                        ctx->stream) != cudaSuccess) {
      rc = fo_fail(ctx, FO_ERR_CUDA, "H2D copy failed");
      break;
    }
    rc = run_harm(ctx, (const double*)dpos, ns, natoms, (int)nmax, L, harmscale, sigma,
                  (double2*)b->d_data + (size_t)s0 * b->per_struct_elems, nullptr, (char*)work);
    if (rc == FO_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      rc = fo_fail(ctx, FO_ERR_CUDA, "harmonic coefficient kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (rc != FO_OK) {
    cudaFree(b->d_data);
    delete b;
    return rc;
  }
  *out = b;
  return FO_OK;
}

namespace {
int align_bank_impl(fo_ctx* ctx, const fo_bank* bank, const int64_t* pairs, int64_t npairs, int invert,
                    int64_t* best_idx, double* best_val, double* frac_idx, double* avg_overlap, double* grid_out,
                    double* euler, double* overlap) {
  if (!ctx) return FO_ERR_INVALID;
  if (!bank || bank->kind != 2) return fo_fail(ctx, FO_ERR_INVALID, "not a spherical coefficient bank");
  if (npairs < 0 || (npairs > 0 && (!pairs || !best_idx || !best_val || !frac_idx)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_align_bank: NULL argument");
  for (int64_t i = 0; i < 2 * npairs; ++i)
    if (pairs[i] < 0 || pairs[i] >= bank->nstruct)
      return fo_fail(ctx, FO_ERR_INVALID, "pair index %lld out of range", (long long)pairs[i]);
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = (int)bank->Jmax, O = invert ? 2 : 1;
  const size_t G3 = (size_t)(2 * L + 2) * (2 * L + 2) * (2 * L + 2);
  int64_t chunk = (int64_t)(((size_t)512 << 20) / (ihalf_elems(L) * 16 + (grid_out ? O * G3 * 8 : 0)));
  if (chunk < 1) chunk = 1;
  if (chunk > npairs) chunk = npairs;
  void *dhalf, *dout, *dpairs, *dgrid = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * ihalf_elems(L) * 16, &dhalf));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * (O * 88 + 8) + 256, &dout));
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, (size_t)chunk * 16, &dpairs));
  if (grid_out) FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * O * G3 * 8, &dgrid));
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = std::min(chunk, npairs - p0);
    FO_CUDA(ctx, cudaMemcpyAsync(dpairs, pairs + 2 * p0, (size_t)np * 16, cudaMemcpyHostToDevice, ctx->stream));
    long long* d_bi = (long long*)dout;
    double* d_bv = (double*)((char*)dout + (size_t)np * O * 24);
    double* d_fr = (double*)((char*)dout + (size_t)np * O * 32);
    double* d_eu = (double*)((char*)dout + (size_t)np * O * 56);
    double* d_ov = (double*)((char*)dout + (size_t)np * O * 80);
    double* d_avg = (double*)((char*)dout + (size_t)np * O * 88);
    {
      fo_prof_scope prof(ctx, FO_PROF_SPH_DOT);
      const size_t smem_dot = dot_mma_smem((int)bank->ngroups, (int)bank->nmax, L);
      if (!ctx->force_generic && smem_dot <= ctx->prop.sharedMemPerBlockOptin) {
        FO_CUDA(ctx, cudaFuncSetAttribute(sph_dot_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dot));
        sph_dot_mma_kernel<<<(unsigned)np, DOT_THREADS, smem_dot, ctx->stream>>>(
            (const double2*)bank->d_data, (const long long*)dpairs, (int)bank->ngroups, (int)bank->nmax, L,
            (double2*)dhalf, d_avg);
      } else {
        sph_dot_kernel<<<(unsigned)np, 256, 0, ctx->stream>>>((const double2*)bank->d_data, (const long long*)dpairs,
                                                              (int)bank->ngroups, (int)bank->nmax, L, (double2*)dhalf, d_avg);
      }
      FO_LAUNCH_CHECK(ctx);
    }
    FO_CHECK(run_isoft(ctx, (const double2*)dhalf, np, L, O, d_bi, d_bv, d_fr, (double*)dgrid));
    if (euler) {
      FO_CHECK(fo_refine_run_dev(ctx, dhalf, np, L, O, d_fr, 1, d_eu, d_ov, nullptr));
      FO_CUDA(ctx, cudaMemcpyAsync(euler + (size_t)p0 * O * 3, d_eu, (size_t)np * O * 24, cudaMemcpyDeviceToHost, ctx->stream));
      FO_CUDA(ctx, cudaMemcpyAsync(overlap + (size_t)p0 * O, d_ov, (size_t)np * O * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FO_CUDA(ctx, cudaMemcpyAsync(best_idx + (size_t)p0 * O * 3, d_bi, (size_t)np * O * 24, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(best_val + (size_t)p0 * O, d_bv, (size_t)np * O * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(frac_idx + (size_t)p0 * O * 3, d_fr, (size_t)np * O * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (avg_overlap)
      FO_CUDA(ctx, cudaMemcpyAsync(avg_overlap + p0, d_avg, (size_t)np * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (grid_out)
      FO_CUDA(ctx, cudaMemcpyAsync(grid_out + (size_t)p0 * O * G3, dgrid, (size_t)np * O * G3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}
}  // namespace

extern "C" int fo_sph_align_bank(fo_ctx* ctx, const fo_bank* bank, const int64_t* pairs, int64_t npairs,
                                 int invert, int64_t* best_idx, double* best_val, double* frac_idx,
                                 double* avg_overlap, double* grid_out) {
  return align_bank_impl(ctx, bank, pairs, npairs, invert, best_idx, best_val, frac_idx, avg_overlap, grid_out,
                         nullptr, nullptr);
}

extern "C" int fo_sph_align_bank_refined(fo_ctx* ctx, const fo_bank* bank, const int64_t* pairs, int64_t npairs,
                                         int invert, int64_t* best_idx, double* best_val, double* frac_idx,
                                         double* avg_overlap, double* euler, double* overlap) {
  if (ctx && npairs > 0 && (!euler || !overlap))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_sph_align_bank_refined: NULL argument");
  return align_bank_impl(ctx, bank, pairs, npairs, invert, best_idx, best_val, frac_idx, avg_overlap, nullptr,
                         euler, overlap);
}
