// Device-side screening of the permutational assignment that follows the hot path.
//
// The Hungarian / LAP step stays on the host (BASELINE.json north_star).  What runs here is the test that
// makes it unnecessary for well-aligned pairs -- the role of screen_periodic / best_perm in fo_host.cu:
// after the arg-max displacement (periodic) or rotation (clusters) has been applied, every atom of the
// second structure looks for its nearest same-group atom of the first.  When those nearest partners form a
// permutation and every column's runner-up is further away by a clear gap, that permutation is the unique
// optimum of the assignment problem (it attains the lower bound sum_j min_i c_ij), so the LAP solver would
// return exactly it.  Pairs that fail the test are flagged and go to the host pool unchanged.
//
//   per_assign_kernel   periodic: screening + the permutation <-> mean-displacement loop of
//                       BasePeriodicAlignment.refine (periodicAlignment.py:27-80; ITERATIVEALIGN
//                       alignutils.f90:111-286, FINDDISPLACEMENT :381-419) + the final distance.
//                       The displacement updates and the distance use the same operations in the same
//                       order as fo_host_refine_periodic (no FMA contraction, canonical summation order),
//                       so device-settled and host-settled pairs agree bit for bit.
//   sph_assign_kernel   clusters: rotation by the Euler angles of the grid maximum (utils.py:447-460),
//                       screening per orientation; the Kearsley fit (utils.py:169-253) stays on the host
//                       and receives the permutation as a hint (fo_host_refine_spherical_hint).
#include <math.h>

#include "fo_internal.h"

namespace {

constexpr int AS_THREADS = 256;

struct PerAssignParams {
  double box[3], ibox[3];
  double F;       // nfspace
  double tolgap;  // smallest accepted gap between a column's nearest and second nearest row
  float tol32;    // relative bound on the error of a single-precision gap: tol32 (max |coordinate| + max box)
  int natoms, ncols, ngroups, niter;
};

// min_image of fo_host.cu, operation by operation
__device__ __forceinline__ double mi_exact(double d, double box, double ibox) {
  return __dsub_rn(d, __dmul_rn(rint(__dmul_rn(d, ibox)), box));
}

// Canonical summation order shared with fo_host.cu (canon_sum there): 32 strided partial sums, each
// accumulated in index order, combined by an xor butterfly (every lane ends with the same bits because
// IEEE addition is commutative).  Called by one full warp.
__device__ __forceinline__ double canon_sum(const double* t, int n, int lane) {
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s = __dadd_rn(s, t[i]);
#pragma unroll
  for (int off = 16; off; off >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, off));
  return s;
}

struct PerAssignSmem {
  double *xa, *yb, *xg, *term;  // [3][N] atom order, [3][N], [3][ncols] group order, [3][N]
  float* xf;                    // [3][ncols] group order, single precision (the nearest-partner search)
  int *perm, *save, *cnt, *cs, *ce;
  double* sc;  // disp[3] dref[3] margin red[8]
};

__device__ __forceinline__ PerAssignSmem per_assign_carve(unsigned char* base, int N, int ncols) {
  PerAssignSmem s;
  double* d = reinterpret_cast<double*>(base);
  s.xa = d;
  s.yb = s.xa + 3 * N;
  s.term = s.yb + 3 * N;
  s.xg = s.term + 3 * N;
  s.sc = s.xg + 3 * ncols;
  int* i = reinterpret_cast<int*>(s.sc + 16);
  s.perm = i;
  s.save = s.perm + N;
  s.cnt = s.save + N;
  s.cs = s.cnt + ncols;
  s.ce = s.cs + ncols;
  s.xf = reinterpret_cast<float*>(s.ce + ncols);
  return s;
}

size_t per_assign_smem(int N, int ncols) { return ((size_t)9 * N + 3 * ncols + 16) * 8 + ((size_t)2 * N + 6 * ncols) * 4; }

// One nearest-partner solve at the displacement in S.sc[0..2]: fills pm (grouped atoms only), returns
// (block-uniform) whether the column minima form a permutation with every gap > tolgap; S.sc[6] = margin.
__device__ bool per_assign_solve(const PerAssignParams& P, const PerAssignSmem& S, const int* __restrict__ gidx,
                                 int* pm) {
  const int tid = threadIdx.x, N = P.natoms;
  for (int r = tid; r < P.ncols; r += AS_THREADS) S.cnt[r] = 0;
  __syncthreads();
  // The search runs in SINGLE precision (the FP32 pipes have twice the FP64 rate and are otherwise idle), like the
  // host pool's screening pass (colmin_periodic_f32, fo_host.cu): coordinates and the shifted structure are rounded
  // to float once, every distance carries an absolute error below tol32 / 4, so nearest partners whose runner-up is
  // more than tol32 further away are the exact nearest partners, and gap - tol32 is a lower bound of the exact gap.
  const double d0 = S.sc[0], d1 = S.sc[1], d2 = S.sc[2];
  const float b0 = (float)P.box[0], b1 = (float)P.box[1], b2 = (float)P.box[2];
  const float i0 = (float)P.ibox[0], i1 = (float)P.ibox[1], i2 = (float)P.ibox[2];
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23: (t + MAGIC) - MAGIC == rintf(t) for |t| < 2^22
  const float inf = __int_as_float(0x7f800000);
  const float tol32 = (float)S.sc[7];
  double mygap = __longlong_as_double(0x7ff0000000000000LL);
  int bad = 0;
  for (int c = tid; c < P.ncols; c += AS_THREADS) {
    const int j = gidx[c];
    // the shifted structure of the host loop: ys = y - disp
    const float y0 = (float)__dsub_rn(S.yb[j], d0), y1 = (float)__dsub_rn(S.yb[N + j], d1),
                y2 = (float)__dsub_rn(S.yb[2 * N + j], d2);
    const int r0 = S.cs[c], r1 = S.ce[c];
    float best = inf, second = inf;
    int br = -1;
    const float* x0 = S.xf;
    const float* x1 = S.xf + P.ncols;
    const float* x2 = S.xf + 2 * P.ncols;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      float dx = x0[r] - y0, dy = x1[r] - y1, dz = x2[r] - y2;
      dx -= (__fadd_rn(dx * i0, MAGIC) - MAGIC) * b0;
      dy -= (__fadd_rn(dy * i1, MAGIC) - MAGIC) * b1;
      dz -= (__fadd_rn(dz * i2, MAGIC) - MAGIC) * b2;
      const float d = dx * dx + dy * dy + dz * dz;
      const bool lt = d < best;
      const float hi = lt ? best : d;
      second = hi < second ? hi : second;
      best = lt ? d : best;
      br = lt ? r : br;
    }
    if (br < 0) {  // NaN coordinates (or an empty group): the host decides
      bad = 1;
    } else {
      atomicAdd(&S.cnt[br], 1);
      pm[gidx[br]] = j;  // X atom of row br <- Y atom of this column
      const float gapf = sqrtf(second) - sqrtf(best) - tol32;  // lower bound of the exact gap
      const double gap = (double)gapf;
      mygap = gap < mygap ? gap : mygap;
      if (!(gap > P.tolgap)) bad = 1;
    }
  }
  // margin = min gap over the block
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, mygap, off);
    mygap = o < mygap ? o : mygap;
  }
  if ((tid & 31) == 0) S.sc[8 + (tid >> 5)] = mygap;
  __syncthreads();  // cnt complete, warp minima visible
  for (int r = tid; r < P.ncols; r += AS_THREADS)
    if (S.cnt[r] != 1) bad = 1;
  const int anybad = __syncthreads_or(bad);
  if (tid == 0) {
    double m = S.sc[8];
    for (int w = 1; w < AS_THREADS / 32; ++w) m = S.sc[8 + w] < m ? S.sc[8 + w] : m;
    S.sc[6] = m;
  }
  __syncthreads();
  return !anybad;
}

// disp -= mean_i min_image(x_i - (y_pm[i] - disp))     (fo_host_refine_periodic: recentre)
__device__ void per_assign_recentre(const PerAssignParams& P, const PerAssignSmem& S, const int* pm) {
  const int tid = threadIdx.x, N = P.natoms;
  for (int i = tid; i < N; i += AS_THREADS) {
    const int j = pm[i];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      S.term[k * N + i] =
          mi_exact(__dsub_rn(S.xa[k * N + i], __dsub_rn(S.yb[k * N + j], S.sc[k])), P.box[k], P.ibox[k]);
  }
  __syncthreads();
  const int w = tid >> 5, lane = tid & 31;
  if (w < 3) {
    const double m = canon_sum(S.term + w * N, N, lane);
    if (lane == 0) S.sc[w] = __dsub_rn(S.sc[w], __ddiv_rn(m, (double)N));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(AS_THREADS)
per_assign_kernel(const __grid_constant__ PerAssignParams P, const double* __restrict__ posA,
                  const double* __restrict__ posB, const double* __restrict__ frac,
                  const int* __restrict__ goff, const int* __restrict__ gidx, double* __restrict__ dist,
                  double* __restrict__ disp_out, void* __restrict__ perm_out, int perm_elt, int* __restrict__ flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = P.natoms, tid = threadIdx.x;
  const PerAssignSmem S = per_assign_carve(smem_raw, N, P.ncols);
  const size_t pair = blockIdx.x;
  const double* xA = posA + pair * (size_t)N * 3;
  const double* yB = posB + pair * (size_t)N * 3;
  double amax = 0.0;
  for (int e = tid; e < 3 * N; e += AS_THREADS) {
    const int i = e / 3, k = e - 3 * i;
    const double xv = xA[e], yv = yB[e];
    S.xa[k * N + i] = xv;
    S.yb[k * N + i] = yv;
    amax = fmax(amax, fmax(fabs(xv), fabs(yv)));
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  if ((tid & 31) == 0) S.sc[8 + (tid >> 5)] = amax;
  for (int c = tid; c < P.ncols; c += AS_THREADS) {
    int lo = 0, hi = P.ngroups;  // group g with goff[g] <= c < goff[g+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (goff[mid] <= c) lo = mid; else hi = mid;
    }
    S.cs[c] = goff[lo];
    S.ce[c] = goff[lo + 1];
  }
  for (int i = tid; i < N; i += AS_THREADS) S.perm[i] = S.save[i] = i;
  if (tid < 3) S.sc[tid] = __ddiv_rn(__dmul_rn(frac[pair * 3 + tid], P.box[tid]), P.F);
  __syncthreads();
  if (tid == 0) {  // error bound of the single-precision search for this pair's coordinate range
    double m = S.sc[8];
    for (int w = 1; w < AS_THREADS / 32; ++w) m = fmax(m, S.sc[8 + w]);
    S.sc[7] = (double)P.tol32 * (m + fmax(P.box[0], fmax(P.box[1], P.box[2])));
  }
  for (int r = tid; r < P.ncols; r += AS_THREADS) {
    const int i = gidx[r];
    S.xg[r] = S.xa[i];
    S.xg[P.ncols + r] = S.xa[N + i];
    S.xg[2 * P.ncols + r] = S.xa[2 * N + i];
    S.xf[r] = (float)S.xa[i];
    S.xf[P.ncols + r] = (float)S.xa[N + i];
    S.xf[2 * P.ncols + r] = (float)S.xa[2 * N + i];
  }
  __syncthreads();
  bool ok = per_assign_solve(P, S, gidx, S.save);
  if (!ok) {
    if (tid == 0) flag[pair] = 1;
    return;
  }
  double margin = S.sc[6];
  if (tid < 3) S.sc[3 + tid] = S.sc[tid];  // dref
  for (int i = tid; i < N; i += AS_THREADS) S.perm[i] = S.save[i];
  __syncthreads();
  for (int it = 0; it < P.niter; ++it) {
    per_assign_recentre(P, S, S.save);
    // the confirming second solve of the reference loop (periodicAlignment.py:66-75) cannot change the
    // assignment when the displacement moved by less than half the stability margin (fo_host.cu: best_perm)
    const double dx = S.sc[0] - S.sc[3], dy = S.sc[1] - S.sc[4], dz = S.sc[2] - S.sc[5];
    if (margin > 0 && 2.0 * sqrt(dx * dx + dy * dy + dz * dz) + 1e-9 < margin) break;
    __syncthreads();  // every thread has read dref before it is overwritten below
    ok = per_assign_solve(P, S, gidx, S.perm);
    if (!ok) {
      if (tid == 0) flag[pair] = 1;
      return;
    }
    margin = S.sc[6];
    if (tid < 3) S.sc[3 + tid] = S.sc[tid];
    int diff = 0;
    for (int i = tid; i < N; i += AS_THREADS) diff |= S.perm[i] != S.save[i];
    const int changed = __syncthreads_or(diff);
    if (!changed) break;
    for (int i = tid; i < N; i += AS_THREADS) S.save[i] = S.perm[i];
    __syncthreads();
  }
  per_assign_recentre(P, S, S.perm);
  for (int i = tid; i < N; i += AS_THREADS) {
    const int j = S.perm[i];
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double a = mi_exact(S.xa[k * N + i], P.box[k], P.ibox[k]);
      const double b = mi_exact(__dsub_rn(S.yb[k * N + j], S.sc[k]), P.box[k], P.ibox[k]);
      const double d = mi_exact(__dsub_rn(a, b), P.box[k], P.ibox[k]);
      s = k == 0 ? __dmul_rn(d, d) : __dadd_rn(s, __dmul_rn(d, d));
    }
    S.term[i] = s;
    if (perm_out) {  // element width of the output: 4 (int32), 2 or 1 bytes (the host pipeline's D2H format)
      const size_t e = pair * (size_t)N + i;
      if (perm_elt == 4) static_cast<int*>(perm_out)[e] = j;
      else if (perm_elt == 2) static_cast<unsigned short*>(perm_out)[e] = (unsigned short)j;
      else static_cast<unsigned char*>(perm_out)[e] = (unsigned char)j;
    }
  }
  __syncthreads();
  if (tid < 32) {
    const double d2 = canon_sum(S.term, N, tid);
    if (tid == 0) {
      dist[pair] = __dsqrt_rn(d2);
      flag[pair] = 0;
    }
    if (tid < 3) disp_out[pair * 3 + tid] = S.sc[tid];
  }
}

// ------------------------------------------------------------------------------------------ clusters

struct SphAssignParams {
  int natoms, ncols, ngroups, norient, n2B;
  double tolgap;
};

__global__ void __launch_bounds__(128)
sph_assign_kernel(const __grid_constant__ SphAssignParams P, const double* __restrict__ posA,
                  const double* __restrict__ posB, const double* __restrict__ frac,
                  const int* __restrict__ goff, const int* __restrict__ gidx, int* __restrict__ perm_out,
                  int* __restrict__ ok_out, double* __restrict__ kdist, double* __restrict__ krot) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = P.natoms, tid = threadIdx.x, T = blockDim.x;
  double* xg = reinterpret_cast<double*>(smem_raw);  // [3][ncols] rows (structure A) in group order
  double* yr = xg + 3 * P.ncols;                       // [3][N] rotated structure B, atom order
  double* M = yr + 3 * N;                              // [9] + pad
  int* cnt = reinterpret_cast<int*>(M + 10);
  int* cs = cnt + P.ncols;
  int* ce = cs + P.ncols;
  __shared__ int s_bad;
  const size_t po = blockIdx.x;  // pair * norient + orientation
  const size_t pair = po / P.norient;
  const int o = (int)(po - pair * P.norient);
  const double* xA = posA + pair * (size_t)N * 3;
  const double* xB = posB + pair * (size_t)N * 3;
  if (tid == 0) {
    // indtoEuler (utils.py:340-345) and EulerM = My Mb Ma (utils.py:447-460)
    const double pi = 3.14159265358979323846, n = (double)P.n2B;
    const double a = (2 * pi / n) * frac[po * 3], b = (pi / n) * frac[po * 3 + 1] + 0.5 * pi / n,
                 y = (2 * pi / n) * frac[po * 3 + 2];
    double sa, ca, sb, cb, sy, cy;
    sincos(a, &sa, &ca);
    sincos(b, &sb, &cb);
    sincos(y, &sy, &cy);
    const double Ma[9] = {ca, -sa, 0, sa, ca, 0, 0, 0, 1};
    const double Mb[9] = {cb, 0, -sb, 0, 1, 0, sb, 0, cb};
    const double My[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    double Tm[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Tm[3 * i + j] = My[3 * i] * Mb[j] + My[3 * i + 1] * Mb[3 + j] + My[3 * i + 2] * Mb[6 + j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[3 * i + j] = Tm[3 * i] * Ma[j] + Tm[3 * i + 1] * Ma[3 + j] + Tm[3 * i + 2] * Ma[6 + j];
    s_bad = 0;
  }
  for (int c = tid; c < P.ncols; c += T) {
    int lo = 0, hi = P.ngroups;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (goff[mid] <= c) lo = mid; else hi = mid;
    }
    cs[c] = goff[lo];
    ce[c] = goff[lo + 1];
    cnt[c] = 0;
    const int i = gidx[c];
    xg[c] = xA[3 * i];
    xg[P.ncols + c] = xA[3 * i + 1];
    xg[2 * P.ncols + c] = xA[3 * i + 2];
  }
  __syncthreads();
  const double sg = o ? -1.0 : 1.0;  // orientation 1: the inverted structure -X2
  for (int i = tid; i < N; i += T) {
    const double b0 = sg * xB[3 * i], b1 = sg * xB[3 * i + 1], b2 = sg * xB[3 * i + 2];
#pragma unroll
    for (int j = 0; j < 3; ++j) yr[j * N + i] = b0 * M[j] + b1 * M[3 + j] + b2 * M[6 + j];  // X2 . M
    perm_out[po * (size_t)N + i] = i;  // atoms in no group keep their place
  }
  __syncthreads();
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  int bad = 0;
  for (int c = tid; c < P.ncols; c += T) {
    const int j = gidx[c];
    const double y0 = yr[j], y1 = yr[N + j], y2 = yr[2 * N + j];
    double best = inf, second = inf;
    int br = -1;
    for (int r = cs[c]; r < ce[c]; ++r) {
      const double dx = xg[r] - y0, dy = xg[P.ncols + r] - y1, dz = xg[2 * P.ncols + r] - y2;
      const double d = dx * dx + dy * dy + dz * dz;
      const bool lt = d < best;
      const double hi = lt ? best : d;
      second = hi < second ? hi : second;
      best = lt ? d : best;
      br = lt ? r : br;
    }
    if (br < 0 || !(second - best > P.tolgap)) {
      bad = 1;
    } else {
      atomicAdd(&cnt[br], 1);
      perm_out[po * (size_t)N + gidx[br]] = j;
    }
  }
  if (bad) s_bad = 1;
  __syncthreads();
  for (int r = tid; r < P.ncols; r += T)
    if (cnt[r] != 1) s_bad = 1;
  __syncthreads();
  if (tid == 0) ok_out[po] = s_bad ? 0 : 1;
  // Kearsley fit of the settled assignment (utils.py:169-253 findrotation_kearsley; kearsley() in fo_host.cu): the
  // rotated structure is already in shared memory, the 4 x 4 quaternion matrix is 16 sums over the atoms, its
  // smallest eigenpair comes from cyclic Jacobi in one thread.  Round 2 measured the host pool as the limit of
  // the end-to-end LJ38 rate at 8 GPUs (32 cores: 3.7 M pairs/s of fits for 8.9 M pairs/s of hot path); with the
  // fit here the host only picks the orientation and copies.
  if (kdist == nullptr || s_bad || tid >= 32) return;
  const int lane = tid;
  // x1 = structure A (atom order, global), x2[i] = rotated B atom perm[i]
  const int* pm = perm_out + po * (size_t)N;
  auto wsum = [&](double v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  };
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  for (int i = lane; i < N; i += 32) {
    const int j = pm[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      c1[k] += xA[3 * i + k];
      c2[k] += yr[k * N + j];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    c1[k] = wsum(c1[k]) / (double)N;
    c2[k] = wsum(c2[k]) / (double)N;
  }
  double q[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // Q00 Q01 Q02 Q03 Q11 Q12 Q13 Q22 Q23 Q33
  for (int i = lane; i < N; i += 32) {
    const int j = pm[i];
    const double a0 = xA[3 * i] - c1[0], a1 = xA[3 * i + 1] - c1[1], a2 = xA[3 * i + 2] - c1[2];
    const double b0 = yr[j] - c2[0], b1 = yr[N + j] - c2[1], b2 = yr[2 * N + j] - c2[2];
    const double xm = a0 - b0, ym = a1 - b1, zm = a2 - b2;
    const double xp = a0 + b0, yp = a1 + b1, zp = a2 + b2;
    q[0] += xm * xm + ym * ym + zm * zm;
    q[1] += ym * zp - yp * zm;
    q[2] += xp * zm - xm * zp;
    q[3] += xm * yp - xp * ym;
    q[4] += yp * yp + zp * zp + xm * xm;
    q[5] += xm * ym - xp * yp;
    q[6] += xm * zm - xp * zp;
    q[7] += xp * xp + zp * zp + ym * ym;
    q[8] += ym * zm - yp * zp;
    q[9] += xp * xp + yp * yp + zm * zm;
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) q[k] = wsum(q[k]);
  if (lane != 0) return;
  double A[4][4] = {{q[0], q[1], q[2], q[3]}, {q[1], q[4], q[5], q[6]}, {q[2], q[5], q[7], q[8]}, {q[3], q[6], q[8], q[9]}};
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      diag += A[a][a] * A[a][a];
#pragma unroll
      for (int r = a + 1; r < 4; ++r) off += A[a][r] * A[a][r];
    }
    if (off <= 1e-36 * diag) break;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int r = a + 1; r < 4; ++r) {
        if (fabs(A[a][r]) < 1e-300) continue;
        const double theta = (A[r][r] - A[a][a]) / (2 * A[a][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), sn = t * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double akp = A[k][a], akr = A[k][r];
          A[k][a] = c * akp - sn * akr;
          A[k][r] = sn * akp + c * akr;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double apk = A[a][k], ark = A[r][k];
          A[a][k] = c * apk - sn * ark;
          A[r][k] = sn * apk + c * ark;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][a], vkr = V[k][r];
          V[k][a] = c * vkp - sn * vkr;
          V[k][r] = sn * vkp + c * vkr;
        }
      }
  }
  double eig = A[0][0], qv[4] = {V[0][0], V[1][0], V[2][0], V[3][0]};
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (A[k][k] < eig) {
      eig = A[k][k];
      qv[0] = V[0][k]; qv[1] = V[1][k]; qv[2] = V[2][k]; qv[3] = V[3][k];
    }
  if (eig < 0) eig = fabs(eig) < 1e-6 ? 0.0 : -eig;
  const double nq = sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2] + qv[3] * qv[3]);
  const double q0 = qv[0] / nq, q1 = qv[1] / nq, q2 = qv[2] / nq, q3 = qv[3] / nq;
  double* R = krot + po * 9;
  R[0] = 2 * (0.5 - q2 * q2 - q3 * q3); R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
  R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = 2 * (0.5 - q1 * q1 - q3 * q3); R[5] = 2 * (q2 * q3 - q0 * q1);
  R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = 2 * (0.5 - q1 * q1 - q2 * q2);
  kdist[po] = sqrt(eig);
}

}  // namespace

// Launchers (declared in fo_internal.h).  d_flag [np]: 0 = settled on the device, 1 = host LAP needed.
int fo_per_assign_run_dev(fo_ctx* ctx, const fo_per_params* p, const double* d_posA, const double* d_posB,
                          const double* d_frac, int64_t np, int niter, double* d_dist, double* d_disp,
                          void* d_perm, int32_t* d_flag, int perm_elt) {
  if (np == 0) return FO_OK;
  const int N = (int)p->natoms, ng = (int)ctx->h_goff.size() - 1, ncols = ctx->h_goff[ng];
  const size_t smem = per_assign_smem(N, ncols);
  if (ncols < 1 || smem > ctx->prop.sharedMemPerBlockOptin) {  // too large for the shared-memory form: all to the host
    FO_CUDA(ctx, cudaMemsetAsync(d_flag, 0xff, (size_t)np * 4, ctx->stream));
    return FO_OK;
  }
  PerAssignParams P;
  double bmax = 0;
  for (int k = 0; k < 3; ++k) {
    P.box[k] = p->box[k];
    P.ibox[k] = 1.0 / p->box[k];
    bmax = std::max(bmax, p->box[k]);
  }
  P.F = (double)p->nfspace;
  P.tolgap = 1e-7 * bmax;
  // float rounding of the coordinates, their differences and the minimum images: a component is off by less than
  // 8 x 2^-24 M (M = max |coordinate| + box), a distance by 1.1e-6 M, a gap by 2.2e-6 M; 3.8e-6 M is subtracted
  P.tol32 = 3.8e-6f;
  P.natoms = N;
  P.ncols = ncols;
  P.ngroups = ng;
  P.niter = niter;
  fo_prof_scope prof(ctx, FO_PROF_ASSIGN);
  FO_CUDA(ctx, cudaFuncSetAttribute(per_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  per_assign_kernel<<<(unsigned)np, AS_THREADS, smem, ctx->stream>>>(P, d_posA, d_posB, d_frac, ctx->d_goff,
                                                                     ctx->d_gidx, d_dist, d_disp, d_perm, perm_elt, d_flag);
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}

// d_perm [np, norient, natoms], d_ok [np, norient] (1 = the permutation is the proven optimum)
int fo_sph_assign_run_dev(fo_ctx* ctx, const double* d_posA, const double* d_posB, const double* d_frac,
                          int64_t np, int64_t natoms, int L, int norient, int32_t* d_perm, int32_t* d_ok,
                          double* d_kdist, double* d_krot) {
  if (np == 0) return FO_OK;
  const int N = (int)natoms, ng = (int)ctx->h_goff.size() - 1, ncols = ctx->h_goff[ng];
  const size_t smem = ((size_t)3 * ncols + 3 * N + 10) * 8 + (size_t)3 * ncols * 4;
  if (smem > ctx->prop.sharedMemPerBlockOptin) {
    FO_CUDA(ctx, cudaMemsetAsync(d_ok, 0, (size_t)np * norient * 4, ctx->stream));
    return FO_OK;
  }
  SphAssignParams P;
  P.natoms = N;
  P.ncols = ncols;
  P.ngroups = ng;
  P.norient = norient;
  P.n2B = 2 * (L + 1);
  P.tolgap = 1e-9;
  const int threads = N <= 64 ? 64 : 128;
  fo_prof_scope prof(ctx, FO_PROF_ASSIGN);
  FO_CUDA(ctx, cudaFuncSetAttribute(sph_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sph_assign_kernel<<<(unsigned)(np * norient), threads, smem, ctx->stream>>>(P, d_posA, d_posB, d_frac, ctx->d_goff,
                                                                             ctx->d_gidx, d_perm, d_ok, d_kdist, d_krot);
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}
