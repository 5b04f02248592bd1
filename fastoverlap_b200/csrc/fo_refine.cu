// Continuous refinement of the rotation on the device.
//
// Replaces BaseSphericalAlignment.maxOverlap / getEnergyGradient / calcWignerMatrices(rot)
// (reference fastoverlap/sphericalAlignment.py:67-113): the reference minimises
//     E(a, b, g) = -Re sum_{l m1 m2} conj(I^l_{m1 m2}) e^{-i m1 a} d^l_{m1 m2}(b) e^{-i m2 g}
// (un-weighted Wigner-d: no sqrt((2l+1)/2)) with scipy's L-BFGS-B starting from the interpolated
// grid maximum.  Here one CTA per (pair, orientation) maximises f = -E by a damped Newton
// (Levenberg-Marquardt) iteration with the analytic gradient and Hessian:
//   * item = (m2 >= 0, m1): d^l, d/db d^l, d2/db2 d^l for l = max(|m1|, m2) .. L by the
//     Kostelec-Rockmore three-term recurrence (DSOFT.f90:81-195) and its first two beta
//     derivatives, started from the closed-form edge value; recurrence coefficients are tabulated
//     once per bandwidth (sph_refine_tab_kernel, cached in the ctx);
//   * (m1, m2) <-> (-m1, -m2) symmetry of the coefficients of real densities: only m2 >= 0 is
//     stored (Ihalf layout of fo_spherical.cu), weight 2 for m2 > 0;
//   * ten block-wide sums (f, 3 gradient, 6 Hessian entries) in a fixed order -> results are
//     bit-identical for any batch split.
// The stopping rule is tighter than L-BFGS-B's (pgtol 1e-5, factr 1e7), so the value reached is
// >= the reference's to ~1e-9 relative and the angles agree to ~1e-6 (tests/test_refine_gpu.py).
#include <math.h>

#include <algorithm>

#include "fo_internal.h"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr int RF_THREADS = 256;
constexpr int RF_MAXIT = 60;
constexpr int RF_PW = 132;  // 2 Jmax + 1 <= 127 powers, padded

__device__ __forceinline__ double powi_dev(double x, int n) {
  if (n < 0) return 0.0;  // only ever multiplied by a zero coefficient
  double r = 1.0;
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// Per item (m2 W + m1 + L): edge constants; per (item, l): coefficients of the l -> l+1 step
//   d_{l+1} = Bc (cos b - Cc) d_l - A d_{l-1}       (weighted by sqrt((2l+1)/2) as in the table kernel)
// and w_l = 1 / sqrt((2l+1)/2) that removes the weight in the objective.
struct RfEdge {
  double K;   // sqrt((2 l0 + 1)/2 binom) * sign^q
  int p, q;   // cos(b/2)^p sin(b/2)^q
};

__global__ void sph_refine_tab_kernel(int L, RfEdge* __restrict__ edge, double4* __restrict__ coef) {
  const int W = 2 * L + 1, L1 = L + 1;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= L1 * W) return;
  const int m1 = item % W - L, m2 = item / W;
  const int am1 = m1 < 0 ? -m1 : m1;
  const int J = am1 > m2 ? am1 : m2;
  int m, p, q;
  double sg;
  if (m1 == J) {
    m = m2; sg = -1.0; p = J + m; q = J - m;
  } else if (m1 == -J) {
    m = m2; sg = 1.0; p = J - m; q = J + m;
  } else if (m2 == J) {
    m = m1; sg = 1.0; p = J + m; q = J - m;
  } else {
    m = m1; sg = -1.0; p = J - m; q = J + m;
  }
  const int am = m < 0 ? -m : m;
  double binom = 1.0;
  for (int i = 1; i <= J - am; ++i) binom *= (double)(J + am + i) / (double)i;
  RfEdge e;
  e.K = sqrt((2.0 * J + 1.0) * 0.5 * binom) * ((q & 1) ? sg : 1.0);
  e.p = p;
  e.q = q;
  edge[item] = e;
  for (int l = 0; l <= L; ++l) {
    double4 c = make_double4(0.0, 0.0, 0.0, 1.0 / sqrt((2.0 * l + 1.0) * 0.5));
    if (l >= J && l < L) {
      const double dj = l, a1 = m1, a2 = m2;
      const double t1 = sqrt((2.0 * dj + 3.0) / (2.0 * dj + 1.0));
      const double t3 = (dj + 1.0) * (2.0 * dj + 1.0);
      const double t5 = 1.0 / sqrt(((dj + 1.0) * (dj + 1.0) - a1 * a1) * ((dj + 1.0) * (dj + 1.0) - a2 * a2));
      c.x = t1 * t3 * t5;
      if (l > 0) {
        const double t2 = sqrt((2.0 * dj + 3.0) / (2.0 * dj - 1.0)) * (dj + 1.0) / dj;
        const double t4 = sqrt((dj * dj - a1 * a1) * (dj * dj - a2 * a2));
        c.y = t2 * t4 * t5;
        c.z = a1 * a2 / (dj * (dj + 1.0));
      }
    }
    coef[(size_t)item * L1 + l] = c;
  }
}

// f, gradient (a, b, g) and Hessian (aa, ab, ag, bb, bg, gg) at x; every thread of the CTA takes
// part, the ten sums land in res[0..9] (shared).  so = -1 for the inverted orientation (odd l flip).
__device__ void rf_eval(const double2* __restrict__ Ip, const RfEdge* __restrict__ edge,
                        const double4* __restrict__ coef, int L, double so, const double x[3],
                        double* red /*[RF_THREADS/32][10]*/, double* res /*[10]*/,
                        double* pw /*[2][RF_PW]: cos(b/2)^t, sin(b/2)^t*/) {
  const int W = 2 * L + 1, L1 = L + 1;
  double sb2, cb2, sb, cb;
  sincos(0.5 * x[1], &sb2, &cb2);
  sincos(x[1], &sb, &cb);
  // powers 0 .. 2L of cos(b/2) and sin(b/2) for the edge values (one short product per thread)
  for (int t = threadIdx.x; t < 2 * (2 * L + 1); t += RF_THREADS) {
    const int e = t >> 1;
    pw[(t & 1) * RF_PW + e] = powi_dev((t & 1) ? sb2 : cb2, e);
  }
  __syncthreads();
  const double* pc = pw;
  const double* ps = pw + RF_PW;
  double acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.0;
  for (int item = threadIdx.x; item < L1 * W; item += RF_THREADS) {
    const int m1 = item % W - L, m2 = item / W;
    const int am1 = m1 < 0 ? -m1 : m1;
    const int l0 = am1 > m2 ? am1 : m2;
    const RfEdge e = edge[item];
    const int p = e.p, q = e.q;
    // C^a S^b and the two derivatives of the edge value (d/db C^a S^b = -a/2 C^{a-1} S^{b+1} + b/2 C^{a+1} S^{b-1})
    // (negative exponents only ever meet a zero coefficient)
    const double Cp2 = p >= 2 ? pc[p - 2] : 0.0, Sq2 = q >= 2 ? ps[q - 2] : 0.0;
    const double Cp1 = p >= 1 ? pc[p - 1] : 0.0, Sq1 = q >= 1 ? ps[q - 1] : 0.0;
    const double Cp = pc[p], Sq = ps[q];
    double d = e.K * Cp * Sq;
    double d1 = e.K * (-0.5 * p * Cp1 * (Sq * sb2) + 0.5 * q * (Cp * cb2) * Sq1);
    double d2 = e.K * (0.25 * p * (p - 1) * Cp2 * (Sq * sb2 * sb2) - 0.25 * (p * (q + 1) + q * (p + 1)) * Cp * Sq +
                       0.25 * q * (q - 1) * (Cp * cb2 * cb2) * Sq2);
    double dm = 0.0, dm1 = 0.0, dm2 = 0.0;
    double s0r = 0.0, s0i = 0.0, s1r = 0.0, s1i = 0.0, s2r = 0.0, s2i = 0.0;
    const double2* ip = Ip + (size_t)item * L1;
    const double4* cp = coef + (size_t)item * L1;
#pragma unroll 4
    for (int l = l0; l <= L; ++l) {  // the loads do not depend on the recurrence: unrolled, they overlap
      const double4 c = cp[l];
      const double2 v = ip[l];
      const double w = (l & 1) ? so * c.w : c.w;
      const double wr = w * v.x, wi = w * v.y;
      s0r = fma(d, wr, s0r);
      s0i = fma(d, wi, s0i);
      s1r = fma(d1, wr, s1r);
      s1i = fma(d1, wi, s1i);
      s2r = fma(d2, wr, s2r);
      s2i = fma(d2, wi, s2i);
      const double u = cb - c.z;
      const double n0 = c.x * (u * d) - c.y * dm;
      const double n1 = c.x * (u * d1 - sb * d) - c.y * dm1;
      const double n2 = c.x * (u * d2 - 2.0 * sb * d1 - cb * d) - c.y * dm2;
      dm = d; dm1 = d1; dm2 = d2;
      d = n0; d1 = n1; d2 = n2;
    }
    double sn, cs;
    sincos((double)m1 * x[0] + (double)m2 * x[2], &sn, &cs);
    const double wt = m2 == 0 ? 1.0 : 2.0;
    const double T0 = wt * (s0r * cs - s0i * sn), U0 = wt * (-s0r * sn - s0i * cs);
    const double T1 = wt * (s1r * cs - s1i * sn), U1 = wt * (-s1r * sn - s1i * cs);
    const double T2 = wt * (s2r * cs - s2i * sn);
    const double fm1 = m1, fm2 = m2;
    acc[0] += T0;
    acc[1] += fm1 * U0;
    acc[2] += T1;
    acc[3] += fm2 * U0;
    acc[4] -= fm1 * fm1 * T0;
    acc[5] += fm1 * U1;
    acc[6] -= fm1 * fm2 * T0;
    acc[7] += T2;
    acc[8] += fm2 * U1;
    acc[9] -= fm2 * fm2 * T0;
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    double v = acc[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) red[(threadIdx.x >> 5) * 10 + i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.0;
    for (int w = 0; w < RF_THREADS / 32; ++w) t += red[w * 10 + threadIdx.x];
    res[threadIdx.x] = t;
  }
  __syncthreads();
}

// (-H + lam I) step = g by Cholesky; false when the matrix is not positive definite
__device__ bool rf_solve(const double* r /*f g3 H6*/, double lam, double step[3]) {
  const double a00 = -r[4] + lam, a01 = -r[5], a02 = -r[6], a11 = -r[7] + lam, a12 = -r[8], a22 = -r[9] + lam;
  if (!(a00 > 0.0)) return false;
  const double l00 = sqrt(a00), l10 = a01 / l00, l20 = a02 / l00;
  const double t11 = a11 - l10 * l10;
  if (!(t11 > 0.0)) return false;
  const double l11 = sqrt(t11), l21 = (a12 - l20 * l10) / l11;
  const double t22 = a22 - l20 * l20 - l21 * l21;
  if (!(t22 > 0.0)) return false;
  const double l22 = sqrt(t22);
  const double y0 = r[1] / l00, y1 = (r[2] - l10 * y0) / l11, y2 = (r[3] - l20 * y0 - l21 * y1) / l22;
  step[2] = y2 / l22;
  step[1] = (y1 - l21 * step[2]) / l11;
  step[0] = (y0 - l10 * step[1] - l20 * step[2]) / l00;
  return isfinite(step[0]) && isfinite(step[1]) && isfinite(step[2]);
}

// start[po*3..]: fractional grid index (from_frac != 0; soft.py:127-130 indtoEuler) or Euler angles.
__global__ void __launch_bounds__(RF_THREADS)
sph_refine_kernel(const double2* __restrict__ Ihalf, const RfEdge* __restrict__ edge,
                  const double4* __restrict__ coef, int L, int norient, const double* __restrict__ start,
                  int from_frac, double* __restrict__ euler_out, double* __restrict__ overlap_out,
                  int* __restrict__ iters_out) {
  __shared__ double red[(RF_THREADS / 32) * 10];
  __shared__ double cur[10], tri[10];
  __shared__ double pw[2 * RF_PW];
  __shared__ double xs[3], xt[3];
  __shared__ int flag;  // 0: evaluate xt, 1: finished
  const size_t po = blockIdx.x;
  const size_t p = po / norient;
  const int o = (int)(po % norient);
  const double so = o ? -1.0 : 1.0;
  const double2* Ip = Ihalf + p * (size_t)(L + 1) * (2 * L + 1) * (L + 1);
  if (threadIdx.x == 0) {
    const double F = 2.0 * (L + 1);
    for (int a = 0; a < 3; ++a) xs[a] = start[po * 3 + a];
    if (from_frac) {
      xs[0] *= 2.0 * kPi / F;
      xs[1] = xs[1] * kPi / F + 0.5 * kPi / F;
      xs[2] *= 2.0 * kPi / F;
    }
  }
  __syncthreads();
  rf_eval(Ip, edge, coef, L, so, xs, red, cur, pw);
  double lam = 0.0, hs = 0.0;
  int it = 0, nev = 1;
  if (threadIdx.x == 0) hs = fmax(fmax(fabs(cur[4]), fabs(cur[7])), fmax(fabs(cur[9]), 1e-300));
  for (; it < RF_MAXIT; ++it) {
    if (threadIdx.x == 0) {
      int fl = 0;
      const double gmax = fmax(fabs(cur[1]), fmax(fabs(cur[2]), fabs(cur[3])));
      if (!(gmax > 1e-12 * fmax(1.0, fabs(cur[0])))) {  // also ends on NaN
        fl = 1;
      } else {
        double step[3];
        // raise the damping until the step exists and is shorter than half a radian
        for (int k = 0; k < 40; ++k) {
          if (rf_solve(cur, lam, step) && fmax(fabs(step[0]), fmax(fabs(step[1]), fabs(step[2]))) <= 0.5) break;
          lam = fmax(10.0 * lam, 1e-3 * hs);
          step[0] = step[1] = step[2] = 0.0;
        }
        for (int a = 0; a < 3; ++a) xt[a] = xs[a] + step[a];
        if (fmax(fabs(step[0]), fmax(fabs(step[1]), fabs(step[2]))) < 1e-15) fl = 1;
      }
      flag = fl;
    }
    __syncthreads();
    if (flag) break;
    rf_eval(Ip, edge, coef, L, so, xt, red, tri, pw);
    ++nev;
    if (threadIdx.x == 0) {
      if (isfinite(tri[0]) && tri[0] >= cur[0] - 1e-15 * fabs(cur[0])) {
        for (int a = 0; a < 3; ++a) xs[a] = xt[a];
        for (int i = 0; i < 10; ++i) cur[i] = tri[i];
        lam = lam > 1e-8 * hs ? 0.1 * lam : 0.0;
      } else {
        lam = fmax(10.0 * lam, 1e-3 * hs);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a) euler_out[po * 3 + a] = xs[a];
    overlap_out[po] = cur[0];
    if (iters_out) iters_out[po] = nev;
  }
}

// one evaluation: value, gradient (a, b, g), Hessian (aa ab ag bb bg gg) at euler[po]
__global__ void __launch_bounds__(RF_THREADS)
sph_refine_eval_kernel(const double2* __restrict__ Ihalf, const RfEdge* __restrict__ edge,
                       const double4* __restrict__ coef, int L, const double* __restrict__ euler,
                       double* __restrict__ value, double* __restrict__ grad, double* __restrict__ hess) {
  __shared__ double red[(RF_THREADS / 32) * 10];
  __shared__ double cur[10];
  __shared__ double pw[2 * RF_PW];
  const size_t p = blockIdx.x;
  const double x[3] = {euler[p * 3], euler[p * 3 + 1], euler[p * 3 + 2]};
  rf_eval(Ihalf + p * (size_t)(L + 1) * (2 * L + 1) * (L + 1), edge, coef, L, 1.0, x, red, cur, pw);
  if (threadIdx.x == 0) value[p] = cur[0];
  if (threadIdx.x < 3) grad[p * 3 + threadIdx.x] = cur[1 + threadIdx.x];
  if (hess && threadIdx.x < 6) hess[p * 6 + threadIdx.x] = cur[4 + threadIdx.x];
}

}  // namespace

int fo_refine_ensure_table(fo_ctx* ctx, int L) {
  if (ctx->refine_L == L && ctx->refine_tab.ptr) return FO_OK;
  const size_t items = (size_t)(L + 1) * (2 * L + 1);
  const size_t bytes = items * sizeof(RfEdge) + items * (L + 1) * sizeof(double4) + 64;
  if (ctx->refine_tab.bytes < bytes) {
    if (ctx->refine_tab.ptr) cudaFree(ctx->refine_tab.ptr);
    ctx->refine_tab.ptr = nullptr;
    ctx->refine_tab.bytes = 0;
    ctx->refine_L = -1;
    if (cudaMalloc(&ctx->refine_tab.ptr, bytes) != cudaSuccess) {
      ctx->refine_tab.ptr = nullptr;
      return fo_fail(ctx, FO_ERR_NOMEM, "cudaMalloc of the rotation-refinement table (%zu bytes) failed", bytes);
    }
    ctx->refine_tab.bytes = bytes;
  }
  double4* coef = (double4*)ctx->refine_tab.ptr;
  RfEdge* edge = (RfEdge*)(coef + items * (L + 1));
  sph_refine_tab_kernel<<<(unsigned)((items + 127) / 128), 128, 0, ctx->stream>>>(L, edge, coef);
  FO_LAUNCH_CHECK(ctx);
  ctx->refine_L = L;
  return FO_OK;
}

int fo_refine_run_dev(fo_ctx* ctx, const void* d_Ihalf, int64_t np, int L, int norient, const double* d_start,
                      int from_frac, double* d_euler, double* d_overlap, int* d_iters) {
  if (np == 0) return FO_OK;
  FO_CHECK(fo_refine_ensure_table(ctx, L));
  const size_t items = (size_t)(L + 1) * (2 * L + 1);
  const double4* coef = (const double4*)ctx->refine_tab.ptr;
  const RfEdge* edge = (const RfEdge*)(coef + items * (L + 1));
  fo_prof_scope prof(ctx, FO_PROF_SPH_REFINE);
  sph_refine_kernel<<<(unsigned)(np * norient), RF_THREADS, 0, ctx->stream>>>(
      (const double2*)d_Ihalf, edge, coef, L, norient, d_start, from_frac, d_euler, d_overlap, d_iters);
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}

int fo_refine_eval_dev(fo_ctx* ctx, const void* d_Ihalf, int64_t np, int L, const double* d_euler, double* d_value,
                       double* d_grad, double* d_hess) {
  if (np == 0) return FO_OK;
  FO_CHECK(fo_refine_ensure_table(ctx, L));
  const size_t items = (size_t)(L + 1) * (2 * L + 1);
  const double4* coef = (const double4*)ctx->refine_tab.ptr;
  const RfEdge* edge = (const RfEdge*)(coef + items * (L + 1));
  sph_refine_eval_kernel<<<(unsigned)np, RF_THREADS, 0, ctx->stream>>>((const double2*)d_Ihalf, edge, coef, L,
                                                                       d_euler, d_value, d_grad, d_hess);
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}
