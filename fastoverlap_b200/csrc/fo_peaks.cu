// Top-k peak extraction on device-resident overlap grids.  Replaces the reference's
// findPeaks / fitPeak / _gaussian (fastoverlap/utils.py:347-396) and FINDPEAKS / FINDPEAK / FIT /
// GAUSSIAN (fastoverlap/f90/fastutils.f90:231-548):
//
//   f = a - min(a)
//   repeat up to npeaks times
//     ind  = arg-max of f (C order, first on ties)                          utils.py:315-317
//     fit  A exp(-(x-x0)^T S (x-x0)) + mu, S upper triangular (each off-diagonal term once), to the
//          (2w+1)^3 window around ind taken with periodic wrap, starting from
//          (A, mu, S, x0) = (f[ind], 0, identity, 0)                           utils.py:355-364
//     peak = x0 + ind, amplitude A, mean mu
//     f   -= the fitted function evaluated on the integer grid WITHOUT wrap    utils.py:385-386
//
// One CTA per grid; the grid stays in HBM and is updated in place (it becomes the reference's residual
// `f`).  The non-linear fit is Levenberg-Marquardt with the analytic Jacobian: the window's residuals
// and Jacobian rows live in shared memory, the 11 x 11 normal equations are formed by 77 threads and
// solved by Cholesky.  The reference fits with scipy curve_fit (MINPACK lmdif, ftol = xtol = 1.5e-8);
// this solver iterates to the local optimum well below that tolerance, so the two agree to the
// reference's own fit tolerance (tests: 1e-5 on positions, relative 1e-5 on amplitudes).
#include <math.h>

#include <algorithm>

#include "fo_internal.h"

namespace {

constexpr int PK_THREADS = 256;
constexpr int PK_NP = 11;        // A, mu, s00 s01 s02 s11 s12 s22, x0 x1 x2
constexpr int PK_MAXIT = 400;
constexpr int PK_MAXW = 4;

struct PkOut {
  double* peaks;      // [P][npeaks][3]
  double* amplitude;  // [P][npeaks]
  double* mean;       // [P][npeaks]
  double* alpha;      // [P][npeaks][6]
  int* nfound;        // [P]
};

__device__ __forceinline__ double pk_block_sum(double v, double* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < PK_THREADS / 32; ++w) s += red[w];
  return s;
}

// model value and (optionally) Jacobian row at window offset d = (d0, d1, d2) - x0
__device__ __forceinline__ double pk_model(const double* p, double c0, double c1, double c2, double* jrow) {
  const double d0 = c0 - p[8], d1 = c1 - p[9], d2 = c2 - p[10];
  const double q = p[2] * d0 * d0 + p[3] * d0 * d1 + p[4] * d0 * d2 + p[5] * d1 * d1 + p[6] * d1 * d2 +
                   p[7] * d2 * d2;
  const double e = exp(-q);
  if (jrow) {
    const double ae = p[0] * e;
    jrow[0] = e;
    jrow[1] = 1.0;
    jrow[2] = -ae * d0 * d0;
    jrow[3] = -ae * d0 * d1;
    jrow[4] = -ae * d0 * d2;
    jrow[5] = -ae * d1 * d1;
    jrow[6] = -ae * d1 * d2;
    jrow[7] = -ae * d2 * d2;
    jrow[8] = ae * (2.0 * p[2] * d0 + p[3] * d1 + p[4] * d2);
    jrow[9] = ae * (p[3] * d0 + 2.0 * p[5] * d1 + p[6] * d2);
    jrow[10] = ae * (p[4] * d0 + p[6] * d1 + 2.0 * p[7] * d2);
  }
  return p[0] * e + p[1];
}

__global__ void __launch_bounds__(PK_THREADS)
grid_peaks_kernel(double* __restrict__ grids, int n0, int n1, int n2, int npeaks, int width, PkOut out) {
  extern __shared__ double smp[];
  const int tid = threadIdx.x;
  const size_t pair = blockIdx.x;
  const int G = n0 * n1 * n2;
  const int ww = 2 * width + 1, NW = ww * ww * ww;
  double* win = smp;                    // [NW] window values
  double* res = win + NW;               // [NW] residuals at the current parameters
  double* jac = res + NW;               // [NW][PK_NP] Jacobian at the current parameters
  double* rtr = jac + (size_t)NW * PK_NP;  // [NW] residuals at the trial parameters
  double* jtr = rtr + NW;               // [NW][PK_NP]
  double* H = jtr + (size_t)NW * PK_NP; // [PK_NP][PK_NP] J^T J
  double* gv = H + PK_NP * PK_NP;       // [PK_NP] J^T r
  double* par = gv + PK_NP;             // [PK_NP] current parameters
  double* ptry = par + PK_NP;           // [PK_NP]
  double* red = ptry + PK_NP;           // [16]
  double* ctl = red + 16;               // [8] control words written by thread 0
  int* redi = reinterpret_cast<int*>(ctl + 8);  // [16]
  double* f = grids + pair * (size_t)G;

  // ---- f -= min(f)
  double mn = 1e300;
  for (int i = tid; i < G; i += PK_THREADS) mn = fmin(mn, f[i]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
  if ((tid & 31) == 0) red[tid >> 5] = mn;
  __syncthreads();
  mn = red[0];
  for (int w = 1; w < PK_THREADS / 32; ++w) mn = fmin(mn, red[w]);
  __syncthreads();
  for (int i = tid; i < G; i += PK_THREADS) f[i] -= mn;
  __syncthreads();

  int found = 0;
  for (int it = 0; it < npeaks; ++it) {
    // ---- arg-max (first flat index on ties)
    double bv = -1e300;
    int bi = 0x7fffffff;
    for (int i = tid; i < G; i += PK_THREADS) {
      const double v = f[i];
      if (v > bv) {
        bv = v;
        bi = i;
      }
    }
    {
      double m = bv;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
      const int cand = (bv == m) ? bi : 0x7fffffff;
      const int imin = __reduce_min_sync(0xffffffffu, cand);
      if ((tid & 31) == 0) {
        red[tid >> 5] = m;
        redi[tid >> 5] = imin;
      }
    }
    __syncthreads();
    bv = red[0];
    bi = redi[0];
    for (int w = 1; w < PK_THREADS / 32; ++w)
      if (red[w] > bv || (red[w] == bv && redi[w] < bi)) {
        bv = red[w];
        bi = redi[w];
      }
    __syncthreads();
    if (bi == 0x7fffffff) break;  // nothing finite
    const int i0 = bi / (n1 * n2), i1 = (bi / n2) % n1, i2 = bi % n2;
    // ---- window with periodic wrap, initial parameters
    for (int t = tid; t < NW; t += PK_THREADS) {
      const int a = t / (ww * ww) - width, b = (t / ww) % ww - width, c = t % ww - width;
      const int j0 = ((i0 + a) % n0 + n0) % n0, j1 = ((i1 + b) % n1 + n1) % n1, j2 = ((i2 + c) % n2 + n2) % n2;
      win[t] = f[((size_t)j0 * n1 + j1) * n2 + j2];
    }
    if (tid < PK_NP) par[tid] = (tid == 0) ? bv : ((tid == 2 || tid == 5 || tid == 7) ? 1.0 : 0.0);
    __syncthreads();
    // ---- Levenberg-Marquardt
    double cost = 0.0;
    {
      double c = 0.0;
      for (int t = tid; t < NW; t += PK_THREADS) {
        const double a = t / (ww * ww) - width, b = (t / ww) % ww - width, cc = t % ww - width;
        const double r = win[t] - pk_model(par, a, b, cc, jac + (size_t)t * PK_NP);
        res[t] = r;
        c += r * r;
      }
      cost = pk_block_sum(c, red);
    }
    double lambda = 1e-3;
    bool ok = isfinite(cost), need_normal = true;
    int iter = 0;
    for (; ok && iter < PK_MAXIT; ++iter) {
      if (need_normal) {  // J^T J (upper triangle) and J^T r: one entry per thread
        if (tid < PK_NP * (PK_NP + 1) / 2 + PK_NP) {
          int i, j;
          if (tid < PK_NP) {
            i = tid;
            j = -1;
          } else {
            int e = tid - PK_NP;
            i = 0;
            while (e >= PK_NP - i) {
              e -= PK_NP - i;
              ++i;
            }
            j = i + e;
          }
          double s = 0.0;
          if (j < 0)
            for (int t = 0; t < NW; ++t) s += jac[(size_t)t * PK_NP + i] * res[t];
          else
            for (int t = 0; t < NW; ++t) s += jac[(size_t)t * PK_NP + i] * jac[(size_t)t * PK_NP + j];
          if (j < 0) {
            gv[i] = s;
          } else {
            H[i * PK_NP + j] = s;
            H[j * PK_NP + i] = s;
          }
        }
        need_normal = false;
      }
      __syncthreads();
      if (tid == 0) {  // (H + lambda diag H) delta = g by Cholesky; trial parameters
        double Lc[PK_NP][PK_NP], y[PK_NP], dl[PK_NP];
        bool pd = true;
        for (int i = 0; i < PK_NP && pd; ++i)
          for (int j = 0; j <= i; ++j) {
            double s = H[i * PK_NP + j];
            if (i == j) s += lambda * (H[i * PK_NP + i] > 0.0 ? H[i * PK_NP + i] : 1.0);
            for (int k = 0; k < j; ++k) s -= Lc[i][k] * Lc[j][k];
            if (i == j) {
              if (!(s > 0.0)) {
                pd = false;
                break;
              }
              Lc[i][i] = sqrt(s);
            } else {
              Lc[i][j] = s / Lc[j][j];
            }
          }
        double dn = 0.0, pn = 0.0;
        if (pd) {
          for (int i = 0; i < PK_NP; ++i) {
            double s = gv[i];
            for (int k = 0; k < i; ++k) s -= Lc[i][k] * y[k];
            y[i] = s / Lc[i][i];
          }
          for (int i = PK_NP - 1; i >= 0; --i) {
            double s = y[i];
            for (int k = i + 1; k < PK_NP; ++k) s -= Lc[k][i] * dl[k];
            dl[i] = s / Lc[i][i];
          }
          for (int i = 0; i < PK_NP; ++i) {
            ptry[i] = par[i] + dl[i];
            // step and parameter norms in the scaled variables sqrt(H_ii) p_i
            const double sc = H[i * PK_NP + i] > 0.0 ? H[i * PK_NP + i] : 1.0;
            dn += sc * dl[i] * dl[i];
            pn += sc * par[i] * par[i];
          }
        }
        ctl[0] = pd ? 1.0 : 0.0;
        ctl[1] = dn;
        ctl[2] = pn;
      }
      __syncthreads();
      if (ctl[0] == 0.0) {  // not positive definite: more damping
        lambda *= 10.0;
        if (!(lambda < 1e15)) break;  // stalled: the current parameters are the answer
        continue;
      }
      const double dn = ctl[1], pn = ctl[2];
      double c = 0.0;
      for (int t = tid; t < NW; t += PK_THREADS) {
        const double a = t / (ww * ww) - width, b = (t / ww) % ww - width, cc = t % ww - width;
        const double r = win[t] - pk_model(ptry, a, b, cc, jtr + (size_t)t * PK_NP);
        rtr[t] = r;
        c += r * r;
      }
      const double ctry = pk_block_sum(c, red);
      if (ctry < cost) {  // accept (uniform: every thread holds the same sums)
        for (int t = tid; t < NW * PK_NP; t += PK_THREADS) jac[t] = jtr[t];
        for (int t = tid; t < NW; t += PK_THREADS) res[t] = rtr[t];
        if (tid < PK_NP) par[tid] = ptry[tid];
        const double dcost = cost - ctry;
        cost = ctry;
        lambda = fmax(lambda * 0.1, 1e-15);
        need_normal = true;
        __syncthreads();
        if (dn <= 1e-24 * (pn + 1e-300) || dcost <= 1e-16 * cost || cost == 0.0) {
          ++iter;
          break;
        }
      } else {
        if (!isfinite(ctry) && !(lambda < 1e15)) {
          ok = false;
          break;
        }
        lambda *= 10.0;
        if (!(lambda < 1e15)) break;  // no further descent possible: converged
      }
    }
    __syncthreads();
    if (!ok || iter >= PK_MAXIT) break;  // curve_fit would raise (maxfev): the reference stops here
    bool fin = true;
    for (int i = 0; i < PK_NP; ++i) fin = fin && isfinite(par[i]);
    if (!fin) break;
    // ---- record the peak, subtract the fitted function (no wrap) from the whole grid
    const double px0 = par[8] + i0, px1 = par[9] + i1, px2 = par[10] + i2;
    if (tid == 0) {
      double* pk = out.peaks + (pair * npeaks + it) * 3;
      pk[0] = px0;
      pk[1] = px1;
      pk[2] = px2;
      out.amplitude[pair * npeaks + it] = par[0];
      out.mean[pair * npeaks + it] = par[1];
      for (int q = 0; q < 6; ++q) out.alpha[(pair * npeaks + it) * 6 + q] = par[2 + q];
    }
    {
      const double A = par[0], mu = par[1];
      const double s00 = par[2], s01 = par[3], s02 = par[4], s11 = par[5], s12 = par[6], s22 = par[7];
      for (int i = tid; i < G; i += PK_THREADS) {
        const int a = i / (n1 * n2), b = (i / n2) % n1, c = i % n2;
        const double d0 = a - px0, d1 = b - px1, d2 = c - px2;
        const double q = s00 * d0 * d0 + s01 * d0 * d1 + s02 * d0 * d2 + s11 * d1 * d1 + s12 * d1 * d2 +
                         s22 * d2 * d2;
        f[i] -= A * exp(-q) + mu;
      }
    }
    ++found;
    __syncthreads();
  }
  if (tid == 0) out.nfound[pair] = found;
}

size_t peaks_smem(int width) {
  const size_t NW = (size_t)(2 * width + 1) * (2 * width + 1) * (2 * width + 1);
  return (NW * (3 + 2 * PK_NP) + PK_NP * PK_NP + 3 * PK_NP + 16 + 8 + 16) * 8;
}

int check_peaks_args(fo_ctx* ctx, int64_t P, const int64_t shape[3], int64_t npeaks, int64_t width) {
  if (!ctx) return FO_ERR_INVALID;
  if (P < 0 || !shape || npeaks < 1 || npeaks > 64) return fo_fail(ctx, FO_ERR_INVALID, "find_peaks: bad P / npeaks");
  if (width < 1 || width > PK_MAXW)
    return fo_fail(ctx, FO_ERR_UNSUPPORTED, "find_peaks: width=%lld outside 1..%d", (long long)width, PK_MAXW);
  for (int i = 0; i < 3; ++i)
    if (shape[i] < 1 || shape[i] > 1024) return fo_fail(ctx, FO_ERR_INVALID, "find_peaks: bad grid shape");
  if (shape[0] * shape[1] * shape[2] > (int64_t)0x7fffffff)
    return fo_fail(ctx, FO_ERR_UNSUPPORTED, "find_peaks: grid too large");
  return FO_OK;
}

}  // namespace

int fo_peaks_run_dev(fo_ctx* ctx, double* d_grids, int64_t P, const int64_t shape[3], int64_t npeaks,
                     int64_t width, double* d_peaks, double* d_amp, double* d_mean, double* d_alpha,
                     int32_t* d_nfound) {
  if (P == 0) return FO_OK;
  const size_t smem = peaks_smem((int)width);
  if (smem > ctx->prop.sharedMemPerBlockOptin)
    return fo_fail(ctx, FO_ERR_UNSUPPORTED, "find_peaks: window too large for shared memory");
  FO_CUDA(ctx, cudaFuncSetAttribute(grid_peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PkOut o{d_peaks, d_amp, d_mean, d_alpha, d_nfound};
  // peaks that are not found stay NaN
  FO_CUDA(ctx, cudaMemsetAsync(d_peaks, 0xff, (size_t)P * npeaks * 3 * 8, ctx->stream));
  FO_CUDA(ctx, cudaMemsetAsync(d_amp, 0xff, (size_t)P * npeaks * 8, ctx->stream));
  FO_CUDA(ctx, cudaMemsetAsync(d_mean, 0xff, (size_t)P * npeaks * 8, ctx->stream));
  FO_CUDA(ctx, cudaMemsetAsync(d_alpha, 0xff, (size_t)P * npeaks * 6 * 8, ctx->stream));
  fo_prof_scope prof(ctx, FO_PROF_PEAKS);
  grid_peaks_kernel<<<(unsigned)P, PK_THREADS, smem, ctx->stream>>>(d_grids, (int)shape[0], (int)shape[1],
                                                                   (int)shape[2], (int)npeaks, (int)width, o);
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}

// device scratch for the per-peak outputs of np grids; returns the five device pointers
int fo_peaks_outputs(fo_ctx* ctx, int64_t np, int64_t npeaks, double** pk, double** amp, double** mean,
                     double** alpha, int32_t** nf) {
  void* p = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_PEAKS, (size_t)np * npeaks * 11 * 8 + (size_t)np * 4 + 64, &p));
  *pk = (double*)p;
  *amp = *pk + (size_t)np * npeaks * 3;
  *mean = *amp + (size_t)np * npeaks;
  *alpha = *mean + (size_t)np * npeaks;
  *nf = (int32_t*)(*alpha + (size_t)np * npeaks * 6);
  return FO_OK;
}

int fo_peaks_copy_out(fo_ctx* ctx, int64_t p0, int64_t np, int64_t npeaks, const double* pk, const double* amp,
                      const double* mean, const double* alpha, const int32_t* nf, double* peaks,
                      double* amplitude, double* meanv, double* alphav, int32_t* nfound) {
  FO_CUDA(ctx, cudaMemcpyAsync(peaks + (size_t)p0 * npeaks * 3, pk, (size_t)np * npeaks * 24, cudaMemcpyDeviceToHost, ctx->stream));
  FO_CUDA(ctx, cudaMemcpyAsync(amplitude + (size_t)p0 * npeaks, amp, (size_t)np * npeaks * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (meanv)
    FO_CUDA(ctx, cudaMemcpyAsync(meanv + (size_t)p0 * npeaks, mean, (size_t)np * npeaks * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (alphav)
    FO_CUDA(ctx, cudaMemcpyAsync(alphav + (size_t)p0 * npeaks * 6, alpha, (size_t)np * npeaks * 48, cudaMemcpyDeviceToHost, ctx->stream));
  FO_CUDA(ctx, cudaMemcpyAsync(nfound + p0, nf, (size_t)np * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return FO_OK;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------

extern "C" int fo_grid_find_peaks_dev(fo_ctx* ctx, double* d_grids, int64_t P, const int64_t shape[3],
                                      int64_t npeaks, int64_t width, double* d_peaks, double* d_amplitude,
                                      double* d_mean, double* d_alpha, int32_t* d_nfound) {
  FO_CHECK(check_peaks_args(ctx, P, shape, npeaks, width));
  if (P > 0 && (!d_grids || !d_peaks || !d_amplitude || !d_mean || !d_alpha || !d_nfound))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_grid_find_peaks_dev: NULL argument");
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  return fo_peaks_run_dev(ctx, d_grids, P, shape, npeaks, width, d_peaks, d_amplitude, d_mean, d_alpha, d_nfound);
}

extern "C" int fo_grid_find_peaks(fo_ctx* ctx, const double* grids, int64_t P, const int64_t shape[3],
                                  int64_t npeaks, int64_t width, double* peaks, double* amplitude, double* mean,
                                  double* alpha, int32_t* nfound, double* residual) {
  FO_CHECK(check_peaks_args(ctx, P, shape, npeaks, width));
  if (P > 0 && (!grids || !peaks || !amplitude || !nfound))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_grid_find_peaks: NULL argument");
  if (P == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t G = (size_t)shape[0] * shape[1] * shape[2];
  int64_t chunk = (int64_t)(((size_t)512 << 20) / (G * 8));
  if (chunk < 1) chunk = 1;
  if (chunk > P) chunk = P;
  void* dg = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * G * 8, &dg));
  double *pk, *amp, *mn, *al;
  int32_t* nf;
  FO_CHECK(fo_peaks_outputs(ctx, chunk, npeaks, &pk, &amp, &mn, &al, &nf));
  for (int64_t p0 = 0; p0 < P; p0 += chunk) {
    const int64_t np = std::min(chunk, P - p0);
    FO_CUDA(ctx, cudaMemcpyAsync(dg, grids + (size_t)p0 * G, (size_t)np * G * 8, cudaMemcpyHostToDevice, ctx->stream));
    FO_CHECK(fo_peaks_run_dev(ctx, (double*)dg, np, shape, npeaks, width, pk, amp, mn, al, nf));
    FO_CHECK(fo_peaks_copy_out(ctx, p0, np, npeaks, pk, amp, mn, al, nf, peaks, amplitude, mean, alpha, nfound));
    if (residual)
      FO_CUDA(ctx, cudaMemcpyAsync(residual + (size_t)p0 * G, dg, (size_t)np * G * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}
