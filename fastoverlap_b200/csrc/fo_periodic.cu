// Periodic hot path (replaces reference fastoverlap/f90/fastbulk.f90 ALIGN1/ALIGNCOEFFS up to
// FINDPEAKS and fastoverlap/periodicAlignment.py:384-456).
//
//   K_sf  per_sf_kernel   structure factors S_g(k) = sum_{j in g} exp(-i k.r_j)
//                         (PERIODICFOURIER fastbulk.f90:599-633, calcFourierCoeff
//                         periodicAlignment.py:400-406)
//   K_xf  per_xf_kernel   C(k) = sum_g S^A_g conj(S^B_g) exp(-k^2 sigma^2)      (setPos :437-438,
//                         DOTFOURIERCOEFFS fastbulk.f90:667-683), zero-padded forward 3-D DFT
//                         to F^3 (fftn :439, FFT3D fastutils.f90:554-569), modulus, arg-max and
//                         findMax's parabola (utils.py:319-338) -- fused, the F^3 grid never
//                         leaves the SM unless the caller asks for it.
//
// Data layout in HBM ("bank"): per structure and permutation group the HALF grid
//   S[ix][iy][l]   ix,iy = 0..2n (k = ix-n, iy-n), l = 0..n (kz >= 0), complex128,
// because S(-k) = conj(S(k)).  All arithmetic is FP64.
#include <math.h>
#include <stdlib.h>

#include <string.h>
#include <omp.h>

#include "fo_internal.h"
#include "fo_async.cuh"
#include "fo_symdft.cuh"

namespace {

constexpr double kTwoPi = 6.283185307179586476925286766559;

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// ------------------------------------------------------------------------------------------
// K_sf: structure factors.
//
// exp(-i k.r) factorises per axis; with c = cos(m theta), s = sin(m theta) for m = 0..n the
// (2n+1)^3 complex sums reduce to 8 REAL sums over (|kx|,|ky|,|kz|) in [0,n]^3:
//   S(rho i, sig j, l) = [ccc - rho sig ssc - sig css - rho scs]
//                      + i[-sig csc - rho scc - ccs + rho sig sss]
// (names: x,y,z factor each c or s).  A thread owns one (i,j) and TL consecutive l: per atom
// 4 DMUL + 8*TL DFMA against 2+TL 16-byte shared loads.  This is 4x fewer flops than the
// complex accumulation over the full grid that the reference performs.
// ------------------------------------------------------------------------------------------
constexpr int SF_TA = 64;  // atoms per shared-memory tile
// smallest row pitch >= m with pitch = 2 (mod 4), in double2 elements (DMMA structure-factor kernels)
__host__ __device__ constexpr int sf_pitch(int m) { return m + ((6 - (m & 3)) & 3); }

template <int TL>
__global__ void __launch_bounds__(512)
per_sf_kernel(const double* __restrict__ pos, const int32_t* __restrict__ goff,
              const int32_t* __restrict__ gidx, int ngroups, int natoms, int n, double kx,
              double ky, double kz, double2* __restrict__ bank, int nitems) {
  extern __shared__ double2 sm_ph[];
  const int M = n + 1;
  const int Mp = M | 1;           // odd row pitch: conflict-free 16 B accesses
  const int LC = (M + TL - 1) / TL;
  const int Mz = (LC * TL) | 1;   // z rows are zero padded up to LC*TL
  double2* phx = sm_ph;
  double2* phy = phx + SF_TA * Mp;
  double2* phz = phy + SF_TA * Mp;

  const int s = blockIdx.x;
  const int g = blockIdx.y;
  const int item = blockIdx.z * blockDim.x + threadIdx.x;
  const bool active = item < nitems;
  int lc = 0, j = 0, i = 0;
  if (active) {
    lc = item % LC;
    j = (item / LC) % M;
    i = item / (LC * M);
  }
  const int l0 = lc * TL;
  const int a_begin = goff[g], a_end = goff[g + 1];
  const double* spos = pos + (size_t)s * natoms * 3;
  const double kax[3] = {kx, ky, kz};

  double acc[TL][8];
#pragma unroll
  for (int t = 0; t < TL; ++t)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[t][q] = 0.0;

  for (int a0 = a_begin; a0 < a_end; a0 += SF_TA) {
    const int ta = min(SF_TA, a_end - a0);
    __syncthreads();
    // phasor tables: one thread per (atom, axis); powers by complex recurrence
    for (int t = threadIdx.x; t < 3 * ta; t += blockDim.x) {
      const int a = t / 3, ax = t - 3 * a;
      const int atom = gidx[a0 + a];
      const double th = kax[ax] * spos[atom * 3 + ax];
      double sn, cs;
      sincos(th, &sn, &cs);
      const int pitch = (ax == 2) ? Mz : Mp;
      double2* row = (ax == 0 ? phx : (ax == 1 ? phy : phz)) + a * pitch;
      double c = 1.0, sv = 0.0;
      row[0] = make_double2(1.0, 0.0);
      for (int m = 1; m <= n; ++m) {
        const double cn = c * cs - sv * sn;
        const double snn = sv * cs + c * sn;
        c = cn;
        sv = snn;
        row[m] = make_double2(c, sv);
      }
      if (ax == 2)
        for (int m = M; m < Mz; ++m) row[m] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (active) {
      const double2* px = phx + i;
      const double2* py = phy + j;
      const double2* pz = phz + l0;
#pragma unroll 2
      for (int a = 0; a < ta; ++a) {
        const double2 X = px[a * Mp];
        const double2 Y = py[a * Mp];
        const double cc = X.x * Y.x, cs = X.x * Y.y, sc = X.y * Y.x, ss = X.y * Y.y;
#pragma unroll
        for (int t = 0; t < TL; ++t) {
          const double2 Z = pz[a * Mz + t];
          acc[t][0] = fma(cc, Z.x, acc[t][0]);  // ccc
          acc[t][1] = fma(cc, Z.y, acc[t][1]);  // ccs
          acc[t][2] = fma(cs, Z.x, acc[t][2]);  // csc
          acc[t][3] = fma(cs, Z.y, acc[t][3]);  // css
          acc[t][4] = fma(sc, Z.x, acc[t][4]);  // scc
          acc[t][5] = fma(sc, Z.y, acc[t][5]);  // scs
          acc[t][6] = fma(ss, Z.x, acc[t][6]);  // ssc
          acc[t][7] = fma(ss, Z.y, acc[t][7]);  // sss
        }
      }
    }
  }
  if (!active) return;
  const int W = 2 * n + 1;
  double2* out = bank + ((size_t)s * ngroups + g) * ((size_t)W * W * M);
#pragma unroll
  for (int r = 0; r < 2; ++r) {      // rho = +1, -1
    if (r == 1 && i == 0) continue;
    const double rho = r ? -1.0 : 1.0;
    const int ix = n + (r ? -i : i);
#pragma unroll
    for (int q = 0; q < 2; ++q) {    // sig = +1, -1
      if (q == 1 && j == 0) continue;
      const double sig = q ? -1.0 : 1.0;
      const int iy = n + (q ? -j : j);
      double2* o = out + ((size_t)ix * W + iy) * M + l0;
#pragma unroll
      for (int t = 0; t < TL; ++t) {
        if (l0 + t < M) {
          const double re = acc[t][0] - rho * sig * acc[t][6] - sig * acc[t][3] - rho * acc[t][5];
          const double im = -sig * acc[t][2] - rho * acc[t][4] - acc[t][1] + rho * sig * acc[t][7];
          o[t] = make_double2(re, im);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K_sf on the FP64 tensor pipe (per_sf2_kernel).  The 8 real sums are one GEMM per (structure,
// group):   C[(i,j,c4)][(l,c2)] = sum_a XY[(i,j,c4)][a] Z[a][(l,c2)]
//   rows r = (i M + j) 4 + c4, c4 = (x: c|s) 2 + (y: c|s);  cols q = 2 l + (z: c|s);  K = atoms.
// A warp owns MT row tiles (8 rows each) and all NT column tiles: per k-step (4 atoms) it builds MT
// A elements (two 8-byte shared loads + one DMUL each) and NT B elements (one load each) for
// MT NT DMMA.8x8x4, i.e. ~1 shared load per DMMA (256 MAC) -- the scalar kernel needs 7 loads per
// 44 FP64 instructions and is held at ~59 % pipe utilisation by LSU return bandwidth and issue.
// The C fragment layout puts the two z components of one l in one lane and the four (x,y)
// combinations of one (i,j) in the 4 lanes {g = 4u..4u+3}: three shuffle rounds gather the 8 sums
// and each of the 4 lanes writes one of the four sign combinations (rho, sig).
// ------------------------------------------------------------------------------------------
template <int MT, int NT>
__global__ void __launch_bounds__(320, 2)
per_sf2_kernel(const double* __restrict__ pos, const int32_t* __restrict__ goff,
               const int32_t* __restrict__ gidx, int ngroups, int natoms, int n, int TA, double kx, double ky,
               double kz, double2* __restrict__ bank) {
  extern __shared__ double2 sm_ph2[];
  const int M = n + 1;
  // row pitches = 2 mod 4 (in double2): the four atoms (t4) of a k-step start 8 banks apart, so the
  // 8-byte operand loads of a warp are free of bank conflicts (odd pitches cost ~1.6 wavefronts per load)
  const int Mp = sf_pitch(M);
  const int Mz = sf_pitch(4 * NT);  // z rows zero padded up to 4 NT values of l
  double2* phx = sm_ph2;
  double2* phy = phx + TA * Mp;
  double2* phz = phy + TA * Mp;
  const int s = blockIdx.x, gq = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  // blockIdx.z splits the row tiles of fine k-grids over several CTAs (each rebuilds the phasors)
  const int warp = blockIdx.z * (blockDim.x >> 5) + (tid >> 5);
  const int g = lane >> 2, t4 = lane & 3;
  const int nrows = M * M * 4;
  const int a_begin = goff[gq], a_end = goff[gq + 1];
  const double* spos = pos + (size_t)s * natoms * 3;
  const double kax[3] = {kx, ky, kz};

  // per-lane constants of the MT row tiles: offsets of the x and y components this lane multiplies
  int offx[MT], offy[MT];
  bool rvalid[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r = (warp * MT + mt) * 8 + g;
    rvalid[mt] = r < nrows;
    const int rr = rvalid[mt] ? r : 0;
    const int c4 = rr & 3, ij = rr >> 2;
    const int i = ij / M, j = ij - i * M;
    offx[mt] = i * 2 + ((c4 >> 1) & 1);  // in doubles within an atom row of phx
    offy[mt] = j * 2 + (c4 & 1);
  }
  int offz[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) offz[nt] = nt * 8 + g;  // = 2 l + c2 (z rows are padded to 4 NT)

  double acc[MT][NT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  // Phasor tables are double buffered: the table of atom tile t+1 is built (by whichever threads get
  // there first) before a thread starts its DMMA share of tile t, so one barrier per tile suffices and
  // the sincos / recurrence work overlaps the tensor-pipe work of the other warps (ncu on the
  // single-buffered version: 15 % of the samples at the two barriers around the phasor phase).
  const size_t buf_elems = (size_t)TA * (2 * Mp + Mz);  // double2 elements per buffer
  auto build_phasors = [&](int a0, int buf) {
    const int ta = min(TA, a_end - a0);
    const int ta4 = (ta + 3) & ~3;
    double2* bx = phx + buf * buf_elems;
    double2* by = phy + buf * buf_elems;
    double2* bz = phz + buf * buf_elems;
    // (items packed into the first warps: spreading them over all warps costs more FP64 issue slots --
    // a partly filled warp occupies the pipe like a full one -- and measured 2.6 % slower)
    for (int t = tid; t < 3 * ta4; t += blockDim.x) {
      const int a = t / 3, ax = t - 3 * a;
      const int pitch = (ax == 2) ? Mz : Mp;
      double2* row = (ax == 0 ? bx : (ax == 1 ? by : bz)) + a * pitch;
      if (a < ta) {
        const int atom = gidx[a0 + a];
        const double th = kax[ax] * spos[atom * 3 + ax];
        double sn, cs;
        sincos(th, &sn, &cs);
        double c = 1.0, sv = 0.0;
        row[0] = make_double2(1.0, 0.0);
        for (int m = 1; m <= n; ++m) {
          const double cn = c * cs - sv * sn;
          const double snn = sv * cs + c * sn;
          c = cn;
          sv = snn;
          row[m] = make_double2(c, sv);
        }
        if (ax == 2)
          for (int m = M; m < Mz; ++m) row[m] = make_double2(0.0, 0.0);
      } else {  // padding atoms of the last k-step contribute nothing
        for (int m = 0; m < pitch; ++m) row[m] = make_double2(0.0, 0.0);
      }
    }
  };
  if (a_begin < a_end) build_phasors(a_begin, 0);
  __syncthreads();
  int buf = 0;
  for (int a0 = a_begin; a0 < a_end; a0 += TA, buf ^= 1) {
    const int ta = min(TA, a_end - a0);
    const int ta4 = (ta + 3) & ~3;
    if (a0 + TA < a_end) build_phasors(a0 + TA, buf ^ 1);
    const double* dx_ = reinterpret_cast<const double*>(phx + buf * buf_elems);
    const double* dy_ = reinterpret_cast<const double*>(phy + buf * buf_elems);
    const double* dz_ = reinterpret_cast<const double*>(phz + buf * buf_elems);
    for (int k0 = 0; k0 < ta4; k0 += 4) {
      const int a = k0 + t4;
      double bz[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bz[nt] = dz_[(size_t)a * Mz * 2 + offz[nt]];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const double av = dx_[(size_t)a * Mp * 2 + offx[mt]] * dy_[(size_t)a * Mp * 2 + offy[mt]];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) fo_dmma(acc[mt][nt], av, bz[nt]);
      }
    }
    __syncthreads();
  }
  // ---- epilogue: gather the 8 sums of (i, j, l) from the 4 lanes g = 4u + c4 and emit S(rho i, sig j, l)
  const int W = 2 * n + 1;
  double2* out = bank + ((size_t)s * ngroups + gq) * ((size_t)W * W * M);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r = (warp * MT + mt) * 8 + g;
    const int rr = rvalid[mt] ? r : 0;
    const int c4 = rr & 3, ij = rr >> 2;
    const int i = ij / M, j = ij - i * M;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      // v[c4'][z]: sums of the row with x,y combination c4' (this lane holds c4' = c4)
      double v[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // lane holding combination q of the same (i,j): same t4, g' = (g & 4) | q
        const int src = (((g & 4) | q) << 2) | t4;
        v[q][0] = __shfl_sync(0xffffffffu, acc[mt][nt][0], src);
        v[q][1] = __shfl_sync(0xffffffffu, acc[mt][nt][1], src);
      }
      const int l = nt * 4 + t4;
      // this lane writes the sign combination (rho, sig) = (c4 & 2 ? -1 : +1, c4 & 1 ? -1 : +1)
      const double rho = (c4 & 2) ? -1.0 : 1.0, sig = (c4 & 1) ? -1.0 : 1.0;
      // names: v[cc=0][c]=ccc v[0][s]=ccs v[cs=1][c]=csc v[1][s]=css v[sc=2][c]=scc v[2][s]=scs v[ss=3][c]=ssc v[3][s]=sss
      const double re = v[0][0] - rho * sig * v[3][0] - sig * v[1][1] - rho * v[2][1];
      const double im = -sig * v[1][0] - rho * v[2][0] - v[0][1] + rho * sig * v[3][1];
      const bool dup = ((c4 & 2) && i == 0) || ((c4 & 1) && j == 0);  // -0 duplicates +0
      if (rvalid[mt] && l < M && !dup) {
        const int ix = n + ((c4 & 2) ? -i : i), iy = n + ((c4 & 1) ? -j : j);
        out[((size_t)ix * W + iy) * M + l] = make_double2(re, im);
      }
    }
  }
}

// mbarrier primitives: fo_async.cuh

// ------------------------------------------------------------------------------------------
// per_sf3_kernel: per_sf2 for n = 9 (M = 10, the default k-grid of a 256-atom cell) without the
// 20 -> 24 column padding.  The 2M = 20 columns split into 16 (l = 0..7: two full column tiles, as in
// per_sf2) and 4 left over (l = 8, 9).  The left-over block X(20) x Y(20) x Z(4) is matricised the
// other way round:   C'[(i, cx, zl)][(j, cy)] = sum_a (X[a][(i,cx)] Z[a][zl]) Y[a][(j,cy)]
// 80 rows = one 8-row tile per i (= per warp: 10 warps), 20 -> 24 columns (3 tiles): 30 DMMA per
// k-step instead of the 50 of a third column tile; 130 instead of 150 DMMA per k-step in total.
// ------------------------------------------------------------------------------------------
template <bool MB>
__global__ void __launch_bounds__(320, 2)
per_sf3_kernel(const double* __restrict__ pos, const int32_t* __restrict__ goff,
               const int32_t* __restrict__ gidx, int ngroups, int natoms, int TA, double kx, double ky,
               double kz, double2* __restrict__ bank) {
  extern __shared__ double2 sm_ph3[];
  constexpr int n = 9, M = 10, MT = 5, NT = 2, NTL = 3;
  constexpr int Mp = sf_pitch(M);        // x rows (pitches = 2 mod 4: conflict-free operand loads)
  constexpr int My = sf_pitch(4 * NTL);  // y rows zero padded to 4 NTL values of j
  constexpr int Mz = sf_pitch(M);
  const int s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const double* spos = pos + (size_t)s * natoms * 3;
  const double kax[3] = {kx, ky, kz};

  // rows r = (warp MT + mt) 8 + g: (i M + j) = warp 10 + 2 mt + (g >> 2), i.e. i = warp for every tile of
  // the warp (MT = 5, M = 10) -- one x operand per k-step serves all five A elements
  const int offx0 = warp * 2 + ((g >> 1) & 1);
  int offy[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) offy[mt] = (2 * mt + (g >> 2)) * 2 + (g & 1);
  // left-over tile of this warp: i = warp, row g = cx 4 + zl, zl = (l - 8) 2 + cz
  const int offxl = warp * 2 + (g >> 2), offzl = 16 + (g & 3);

  const size_t buf_elems = (size_t)TA * (Mp + My + Mz);  // double2 elements per buffer
  double2* phx = sm_ph3;
  double2* phy = phx + (size_t)TA * Mp;
  double2* phz = phy + (size_t)TA * My;
  // (items packed into the first warps: spreading them over all warps costs more FP64 issue slots --
  // a partly filled warp occupies the pipe like a full one -- and measured 2.6 % slower)
  auto build_phasors = [&](int a0, int a_end, int buf) {
    const int ta = min(TA, a_end - a0);
    const int ta4 = (ta + 3) & ~3;
    double2* bx = phx + buf * buf_elems;
    double2* by = phy + buf * buf_elems;
    double2* bz = phz + buf * buf_elems;
    for (int t = tid; t < 3 * ta4; t += blockDim.x) {
      const int a = t / 3, ax = t - 3 * a;
      const int pitch = ax == 0 ? Mp : (ax == 1 ? My : Mz);
      double2* row = (ax == 0 ? bx : (ax == 1 ? by : bz)) + a * pitch;
      if (a < ta) {
        const int atom = gidx[a0 + a];
        const double th = kax[ax] * spos[atom * 3 + ax];
        double sn, cs;
        sincos(th, &sn, &cs);
        double c = 1.0, sv = 0.0;
        row[0] = make_double2(1.0, 0.0);
        for (int m = 1; m <= n; ++m) {
          const double cn = c * cs - sv * sn;
          const double snn = sv * cs + c * sn;
          c = cn;
          sv = snn;
          row[m] = make_double2(c, sv);
        }
        for (int m = M; m < pitch; ++m) row[m] = make_double2(0.0, 0.0);
      } else {  // padding atoms of the last k-step contribute nothing
        for (int m = 0; m < pitch; ++m) row[m] = make_double2(0.0, 0.0);
      }
    }
  };
  // One CTA per structure runs through the permutation groups in turn: the table of the next tile -- of
  // the same group or the first tile of the next non-empty one -- is built while the current tile is on the
  // tensor pipe, so only the first tile of the structure has an exposed build, and a warp's epilogue of one
  // group overlaps the other warps' DMMA of the next.  (One CTA per (structure, group) left the short
  // groups -- 52 of 256 atoms in BLJ256 -- with a build and an epilogue as long as their DMMA phase.)
  auto next_group = [&](int q) {  // first non-empty group >= q, or ngroups
    while (q < ngroups && goff[q + 1] == goff[q]) ++q;
    return q;
  };
  // Tile hand-over.  MB: mbarriers instead of a CTA barrier per tile -- full[b] (every thread arrives after
  // its share of the table of buffer b is written) and empty[b] (every thread arrives when it has finished
  // reading buffer b).  A thread waits for full[b] before the tile, and builds the next tile in the MIDDLE
  // of the current one, after waiting for empty of the buffer it overwrites -- by then a formality: all
  // warps left that buffer half a tile ago.  So warps never meet; they can drift by up to half a tile
  // (a __syncthreads per tile cost ~2.5 % of the kernel each: profiles/r01_summary.md).
  __shared__ uint64_t mbar[4];  // full[0], full[1], empty[0], empty[1]
  if (MB) {
    if (tid == 0)
      for (int i = 0; i < 4; ++i) fo_mbar_init(&mbar[i], blockDim.x);
    __syncthreads();
  }
  {
    const int q0 = next_group(0);
    if (q0 < ngroups) build_phasors(goff[q0], goff[q0 + 1], 0);
  }
  if (MB) fo_mbar_arrive(&mbar[0]); else __syncthreads();
  int buf = 0, tile = 0;
  constexpr int W = 2 * n + 1;
  for (int gq = 0; gq < ngroups; ++gq) {
    const int a_begin = goff[gq], a_end = goff[gq + 1];
    double acc[MT][NT][2], accl[NTL][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) accl[nt][0] = accl[nt][1] = 0.0;
    for (int a0 = a_begin; a0 < a_end; a0 += TA, buf ^= 1, ++tile) {
      const int ta = min(TA, a_end - a0);
      const int ta4 = (ta + 3) & ~3;
      // next tile (CTA-uniform): same group, or the first tile of the next non-empty group
      int nb = -1, ne = 0;
      if (a0 + TA < a_end) {
        nb = a0 + TA;
        ne = a_end;
      } else {
        const int q = next_group(gq + 1);
        if (q < ngroups) {
          nb = goff[q];
          ne = goff[q + 1];
        }
      }
      const bool more = nb >= 0;
      if (!MB && more) build_phasors(nb, ne, buf ^ 1);
      if (MB) fo_mbar_wait(&mbar[buf], (tile >> 1) & 1);
      const double* dx_ = reinterpret_cast<const double*>(phx + buf * buf_elems);
      const double* dy_ = reinterpret_cast<const double*>(phy + buf * buf_elems);
      const double* dz_ = reinterpret_cast<const double*>(phz + buf * buf_elems);
      auto mma_ksteps = [&](int kb, int ke) {
        for (int k0 = kb; k0 < ke; k0 += 4) {
          const int a = k0 + t4;
          const double* xr = dx_ + (size_t)a * Mp * 2;
          const double* yr = dy_ + (size_t)a * My * 2;
          const double* zr = dz_ + (size_t)a * Mz * 2;
          double bz[NT];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) bz[nt] = zr[nt * 8 + g];
          const double xv = xr[offx0];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const double av = xv * yr[offy[mt]];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) fo_dmma(acc[mt][nt], av, bz[nt]);
          }
          const double avl = xr[offxl] * zr[offzl];
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) fo_dmma(accl[nt], avl, yr[nt * 8 + g]);
        }
      };
      if (MB) {
        const int kh = ((ta4 >> 2) >> 1) << 2;  // first half of the k-steps
        mma_ksteps(0, kh);
        if (more) {
          if (tile >= 1) fo_mbar_wait(&mbar[2 + (buf ^ 1)], ((tile - 1) >> 1) & 1);
          build_phasors(nb, ne, buf ^ 1);
          fo_mbar_arrive(&mbar[buf ^ 1]);
        }
        mma_ksteps(kh, ta4);
        fo_mbar_arrive(&mbar[2 + buf]);
      } else {
        mma_ksteps(0, ta4);
        if (more) __syncthreads();  // the last tile of the structure needs no barrier behind it
      }
    }
    // ---- epilogue of group gq (zeros for an empty group)
    double2* out = bank + ((size_t)s * ngroups + gq) * ((size_t)W * W * M);
    // main block: as per_sf2 (lanes g = 4u + c4 of one (i, j); this lane writes the signs of its own c4)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r = (warp * MT + mt) * 8 + g;
      const int c4 = r & 3, ij = r >> 2;
      const int i = ij / M, j = ij - i * M;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        double v[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int src = (((g & 4) | q) << 2) | t4;
          v[q][0] = __shfl_sync(0xffffffffu, acc[mt][nt][0], src);
          v[q][1] = __shfl_sync(0xffffffffu, acc[mt][nt][1], src);
        }
        const int l = nt * 4 + t4;
        const double rho = (c4 & 2) ? -1.0 : 1.0, sig = (c4 & 1) ? -1.0 : 1.0;
        const double re = v[0][0] - rho * sig * v[3][0] - sig * v[1][1] - rho * v[2][1];
        const double im = -sig * v[1][0] - rho * v[2][0] - v[0][1] + rho * sig * v[3][1];
        const bool dup = ((c4 & 2) && i == 0) || ((c4 & 1) && j == 0);
        if (!dup) {
          const int ix = n + ((c4 & 2) ? -i : i), iy = n + ((c4 & 1) ? -j : j);
          out[((size_t)ix * W + iy) * M + l] = make_double2(re, im);
        }
      }
    }
    // left-over block: i = warp; lane (g, t4) of column tile nt holds (j = 4 nt + t4; cy = 0, 1) of row
    // (cx = g >> 2, l = 8 + ((g >> 1) & 1), cz = g & 1).  The four lanes (cx, cz) of one (j, l) gather all
    // eight sums v[cx][cy][cz] and write the sign combination rho = (cx ? - : +), sig = (cz ? - : +).
    {
      const int i = warp;
      const int cxw = g >> 2, lb = (g >> 1) & 1, czw = g & 1;
      const int l = 8 + lb;
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        double v[2][2][2];  // [cx][cy][cz]
#pragma unroll
        for (int cx = 0; cx < 2; ++cx)
#pragma unroll
          for (int cz = 0; cz < 2; ++cz) {
            const int src = ((cx * 4 + lb * 2 + cz) << 2) | t4;
            v[cx][0][cz] = __shfl_sync(0xffffffffu, accl[nt][0], src);
            v[cx][1][cz] = __shfl_sync(0xffffffffu, accl[nt][1], src);
          }
        const int j = nt * 4 + t4;
        const double rho = cxw ? -1.0 : 1.0, sig = czw ? -1.0 : 1.0;
        // re = ccc - rho sig ssc - sig css - rho scs ; im = -sig csc - rho scc - ccs + rho sig sss   (x y z)
        const double re = v[0][0][0] - rho * sig * v[1][1][0] - sig * v[0][1][1] - rho * v[1][0][1];
        const double im = -sig * v[0][1][0] - rho * v[1][0][0] - v[0][0][1] + rho * sig * v[1][1][1];
        const bool dup = (cxw && i == 0) || (czw && j == 0);
        if (j < M && !dup) {
          const int ix = n + (cxw ? -i : i), iy = n + (czw ? -j : j);
          out[((size_t)ix * W + iy) * M + l] = make_double2(re, im);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// per_sfx_kernel: structure factors of BOTH structures of a pair + damped cross-spectrum, written directly as the
// stage-X image of per_xf6_kernel -- the independent-pairs path never materialises the structure-factor bank
// (per_sf3 wrote 2 x 116 KB per pair, per_cross6 read them back: a pure HBM pass, 10 % of the step).
//
// One CTA per pair runs the tile pipeline of per_sf3_kernel over the units (A, g0), (B, g0), (A, g1), (B, g1), ...
// A lane's epilogue of per_sf3 yields 13 complex values S(rho i, sig j, l); the same lane holds the same k for
// every unit, so the cross-spectrum is elementwise per lane:
//   unit (A, g): S_A -> the lane's stash in TENSOR MEMORY (tcgen05.st; TMEM is otherwise idle in this FP64 path:
//                13 x 16 bytes per thread = 52 columns per warp in the warp's own lane quarter, 256 columns per
//                CTA, two CTAs per SM; reading it back costs ~12 cycles where an L2 stash cost ~700);
//   unit (B, g): C_g = S_A conj(S_B) exp(-|k|^2 sigma^2), paired over +-ky and +-kx by two warp shuffles with the
//                lanes that hold the other signs (C_E = C(+) + C(-), C_O' = -i (C(+) - C(-))): stored to the
//                image by the first group, added by the later ones with fire-and-forget reductions (red.add.f64;
//                always the same thread per address, so the sum has a fixed order).
// ------------------------------------------------------------------------------------------
constexpr int SFX_NVAL = 13;
constexpr int SFX_TMEM_COLS = 256;  // >= 3 warps x 52 columns per lane quarter, power of two

__device__ __forceinline__ void fo_tmem_st_d2(uint32_t taddr, double x, double y) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr),
               "r"((uint32_t)__double2loint(x)), "r"((uint32_t)__double2hiint(x)), "r"((uint32_t)__double2loint(y)),
               "r"((uint32_t)__double2hiint(y))
               : "memory");
}
__device__ __forceinline__ double2 fo_tmem_ld_d2(uint32_t taddr) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return make_double2(__hiloint2double((int)r1, (int)r0), __hiloint2double((int)r3, (int)r2));
}

__global__ void __launch_bounds__(320, 2)
per_sfx_kernel(const double* __restrict__ posA, const double* __restrict__ posB, const int32_t* __restrict__ goff,
               const int32_t* __restrict__ gidx, int ngroups, int natoms, int TA, double kx, double ky, double kz,
               double sigma, int RXp, double* __restrict__ ximg, size_t ximg_stride) {
  extern __shared__ double2 sm_phx[];
  constexpr int n = 9, M = 10, MT = 5, NT = 2, NTL = 3, RY = 2 * M;
  constexpr int Mp = sf_pitch(M);
  constexpr int My = sf_pitch(4 * NTL);
  constexpr int Mz = sf_pitch(M);
  const size_t pair = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const double* sposA = posA + pair * natoms * 3;
  const double* sposB = posB + pair * natoms * 3;
  const double kax[3] = {kx, ky, kz};

  const int offx0 = warp * 2 + ((g >> 1) & 1);
  int offy[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) offy[mt] = (2 * mt + (g >> 2)) * 2 + (g & 1);
  const int offxl = warp * 2 + (g >> 2), offzl = 16 + (g & 3);

  const size_t buf_elems = (size_t)TA * (Mp + My + Mz);
  double2* phx = sm_phx;
  double2* phy = phx + (size_t)TA * Mp;
  double2* phz = phy + (size_t)TA * My;
  auto build_phasors = [&](const double* spos, int a0, int a_end, int buf) {
    const int ta = min(TA, a_end - a0);
    const int ta4 = (ta + 3) & ~3;
    double2* bx = phx + buf * buf_elems;
    double2* by = phy + buf * buf_elems;
    double2* bz = phz + buf * buf_elems;
    for (int t = tid; t < 3 * ta4; t += blockDim.x) {
      const int a = t / 3, ax = t - 3 * a;
      const int pitch = ax == 0 ? Mp : (ax == 1 ? My : Mz);
      double2* row = (ax == 0 ? bx : (ax == 1 ? by : bz)) + a * pitch;
      if (a < ta) {
        const int atom = gidx[a0 + a];
        const double th = kax[ax] * spos[atom * 3 + ax];
        double sn, cs;
        sincos(th, &sn, &cs);
        double c = 1.0, sv = 0.0;
        row[0] = make_double2(1.0, 0.0);
        for (int m = 1; m <= n; ++m) {
          const double cn = c * cs - sv * sn;
          const double snn = sv * cs + c * sn;
          c = cn;
          sv = snn;
          row[m] = make_double2(c, sv);
        }
        for (int m = M; m < pitch; ++m) row[m] = make_double2(0.0, 0.0);
      } else {
        for (int m = 0; m < pitch; ++m) row[m] = make_double2(0.0, 0.0);
      }
    }
  };
  // units u = 2 group + structure (0 = A, 1 = B); empty groups are skipped (they contribute nothing)
  const int nunits = 2 * ngroups;
  auto next_unit = [&](int u) {
    while (u < nunits && goff[(u >> 1) + 1] == goff[u >> 1]) ++u;
    return u;
  };
  __shared__ uint64_t mbar[4];  // full[0], full[1], empty[0], empty[1]
  __shared__ uint32_t s_tmem;
  __shared__ int s_done;
  __shared__ double s_damp[3 * M];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fo_smem_addr(&s_tmem)),
                 "n"(SFX_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) fo_mbar_init(&mbar[i], blockDim.x);
    s_done = 0;
  }
  if (tid >= 32 && tid < 32 + 3 * M) {
    const int ax = (tid - 32) / M, m = (tid - 32) - ax * M;
    const double k = kax[ax] * (double)m;
    s_damp[tid - 32] = exp(-(k * k) * (sigma * sigma));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // the warp's stash: its own lane quarter (warp % 4), 52 columns per warp of the quarter
  const uint32_t tstash = s_tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (4 * SFX_NVAL);
  const int u_first = next_unit(0);
  if (u_first < nunits) build_phasors(sposA, goff[u_first >> 1], goff[(u_first >> 1) + 1], 0);
  fo_mbar_arrive(&mbar[0]);
  int buf = 0, tile = 0;
  double* XE = ximg + pair * ximg_stride;
  double* XO = XE + (size_t)M * RXp;
  for (int u = u_first; u < nunits; u = next_unit(u + 1)) {
    const int gq = u >> 1, st = u & 1;
    const int a_begin = goff[gq], a_end = goff[gq + 1];
    double acc[MT][NT][2], accl[NTL][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) accl[nt][0] = accl[nt][1] = 0.0;
    for (int a0 = a_begin; a0 < a_end; a0 += TA, buf ^= 1, ++tile) {
      const int ta = min(TA, a_end - a0);
      const int ta4 = (ta + 3) & ~3;
      // next tile (CTA-uniform): same unit, or the first tile of the next non-empty unit
      int nb = -1, ne = 0;
      const double* nspos = st ? sposB : sposA;
      if (a0 + TA < a_end) {
        nb = a0 + TA;
        ne = a_end;
      } else {
        const int q = next_unit(u + 1);
        if (q < nunits) {
          nb = goff[q >> 1];
          ne = goff[(q >> 1) + 1];
          nspos = (q & 1) ? sposB : sposA;
        }
      }
      const bool more = nb >= 0;
      fo_mbar_wait(&mbar[buf], (tile >> 1) & 1);
      const double* dx_ = reinterpret_cast<const double*>(phx + buf * buf_elems);
      const double* dy_ = reinterpret_cast<const double*>(phy + buf * buf_elems);
      const double* dz_ = reinterpret_cast<const double*>(phz + buf * buf_elems);
      auto mma_ksteps = [&](int kb, int ke) {
        for (int k0 = kb; k0 < ke; k0 += 4) {
          const int a = k0 + t4;
          const double* xr = dx_ + (size_t)a * Mp * 2;
          const double* yr = dy_ + (size_t)a * My * 2;
          const double* zr = dz_ + (size_t)a * Mz * 2;
          double bz[NT];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) bz[nt] = zr[nt * 8 + g];
          const double xv = xr[offx0];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const double av = xv * yr[offy[mt]];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) fo_dmma(acc[mt][nt], av, bz[nt]);
          }
          const double avl = xr[offxl] * zr[offzl];
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) fo_dmma(accl[nt], avl, yr[nt * 8 + g]);
        }
      };
      const int kh = ((ta4 >> 2) >> 1) << 2;  // first half of the k-steps
      mma_ksteps(0, kh);
      if (more) {
        if (tile >= 1) fo_mbar_wait(&mbar[2 + (buf ^ 1)], ((tile - 1) >> 1) & 1);
        build_phasors(nspos, nb, ne, buf ^ 1);
        fo_mbar_arrive(&mbar[buf ^ 1]);
      }
      mma_ksteps(kh, ta4);
      fo_mbar_arrive(&mbar[2 + buf]);
    }
    // ---- epilogue of the unit: S of this lane's 13 k-points -> stash (A) / cross-spectrum into the image (B)
    const bool firstB = (u == u_first + 1);
#pragma unroll
    for (int vi = 0; vi < SFX_NVAL; ++vi) {
      // value vi: S = (re, im) at k = (rho i, sig j, l); rmask: lane xor to the opposite kx sign (ky: lane ^ 4)
      double re, im;
      int i, j, l, rmask;
      bool rneg, sneg, ok;
      if (vi < MT * NT) {
        // main block: lanes g = 4 u + c4 of one (i, j); this lane takes the signs of its own c4
        const int mt = vi / NT, nt = vi % NT;
        const int r = (warp * MT + mt) * 8 + g;
        const int c4 = r & 3, ij = r >> 2;
        i = ij / M;
        j = ij - i * M;
        double v[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int src = (((g & 4) | q) << 2) | t4;
          v[q][0] = __shfl_sync(0xffffffffu, acc[mt][nt][0], src);
          v[q][1] = __shfl_sync(0xffffffffu, acc[mt][nt][1], src);
        }
        l = nt * 4 + t4;
        const double rho = (c4 & 2) ? -1.0 : 1.0, sig = (c4 & 1) ? -1.0 : 1.0;
        re = v[0][0] - rho * sig * v[3][0] - sig * v[1][1] - rho * v[2][1];
        im = -sig * v[1][0] - rho * v[2][0] - v[0][1] + rho * sig * v[3][1];
        rneg = (c4 & 2) != 0;
        sneg = (c4 & 1) != 0;
        ok = !((rneg && i == 0) || (sneg && j == 0));
        rmask = 8;
      } else {
        // left-over block: i = warp; lane (g, t4) of column tile nt holds (j = 4 nt + t4; cy = 0, 1) of row
        // (cx = g >> 2, l = 8 + ((g >> 1) & 1), cz = g & 1); the four lanes (cx, cz) of one (j, l) gather all
        // eight sums v[cx][cy][cz]; this lane takes rho = (cx ? - : +), sig = (cz ? - : +)
        const int nt = vi - MT * NT;
        i = warp;
        const int cxw = g >> 2, lb = (g >> 1) & 1, czw = g & 1;
        l = 8 + lb;
        double v[2][2][2];  // [cx][cy][cz]
#pragma unroll
        for (int cx = 0; cx < 2; ++cx)
#pragma unroll
          for (int cz = 0; cz < 2; ++cz) {
            const int src = ((cx * 4 + lb * 2 + cz) << 2) | t4;
            v[cx][0][cz] = __shfl_sync(0xffffffffu, accl[nt][0], src);
            v[cx][1][cz] = __shfl_sync(0xffffffffu, accl[nt][1], src);
          }
        const int jj = nt * 4 + t4;
        const double rho = cxw ? -1.0 : 1.0, sig = czw ? -1.0 : 1.0;
        re = v[0][0][0] - rho * sig * v[1][1][0] - sig * v[0][1][1] - rho * v[1][0][1];
        im = -sig * v[0][1][0] - rho * v[1][0][0] - v[0][0][1] + rho * sig * v[1][1][1];
        rneg = cxw != 0;
        sneg = czw != 0;
        ok = jj < M && !((rneg && i == 0) || (sneg && jj == 0));
        j = jj < M ? jj : M - 1;
        rmask = 16;
      }
      if (st == 0) {
        fo_tmem_st_d2(tstash + 4 * vi, re, im);
        continue;
      }
      const double2 sa = fo_tmem_ld_d2(tstash + 4 * vi);
      const double dmp = s_damp[i] * s_damp[M + j] * s_damp[2 * M + l];
      const double2 c = make_double2((sa.x * re + sa.y * im) * dmp, (sa.y * re - sa.x * im) * dmp);  // S_A conj(S_B)
      const double wx = __shfl_xor_sync(0xffffffffu, c.x, 4), wy = __shfl_xor_sync(0xffffffffu, c.y, 4);
      double2 a = c;
      if (j > 0) a = sneg ? make_double2(wy - c.y, c.x - wx) : make_double2(c.x + wx, c.y + wy);
      const double px = __shfl_xor_sync(0xffffffffu, a.x, rmask), py = __shfl_xor_sync(0xffffffffu, a.y, rmask);
      double2 r = a;
      if (i > 0) r = rneg ? make_double2(py - a.y, a.x - px) : make_double2(a.x + px, a.y + py);
      if (ok) {
        double* dst = (rneg ? XO : XE) + (size_t)i * RXp + ((sneg ? n + j : j) * RY + 2 * l);
        if (firstB) {
          *reinterpret_cast<double2*>(dst) = r;
        } else {
          atomicAdd(dst, r.x);
          atomicAdd(dst + 1, r.y);
        }
      }
    }
    if (st == 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // the last warp to finish returns the CTA's tensor memory (no CTA-wide barrier: warps drift by half a tile)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  int last = 0;
  if (lane == 0) last = atomicAdd(&s_done, 1) == (int)(blockDim.x >> 5) - 1;
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(SFX_TMEM_COLS) : "memory");
  }
}

// Expand the half-grid bank to the reference's full (2n+1)^3 layout (calcFourierCoeff output).
__global__ void per_expand_kernel(const double2* __restrict__ bank, double2* __restrict__ full,
                                  int n, size_t nsg) {
  const int W = 2 * n + 1, M = n + 1;
  const size_t per = (size_t)W * W * W;
  const size_t total = nsg * per;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const size_t sg = e / per;
    int r = (int)(e - sg * per);
    const int iz = r % W;
    r /= W;
    const int iy = r % W;
    const int ix = r / W;
    const double2* b = bank + sg * ((size_t)W * W * M);
    double2 v;
    if (iz >= n) {
      v = b[((size_t)ix * W + iy) * M + (iz - n)];
    } else {
      v = b[((size_t)(2 * n - ix) * W + (2 * n - iy)) * M + (n - iz)];
      v.y = -v.y;
    }
    full[e] = v;
  }
}

// Inverse: take the kz >= 0 half of caller-supplied full-grid coefficients (Cs= hook).
__global__ void per_compress_kernel(const double2* __restrict__ full, double2* __restrict__ bank,
                                    int n, size_t nsg) {
  const int W = 2 * n + 1, M = n + 1;
  const size_t per = (size_t)W * W * M;
  const size_t total = nsg * per;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const size_t sg = e / per;
    int r = (int)(e - sg * per);
    const int l = r % M;
    r /= M;
    const int iy = r % W;
    const int ix = r / W;
    bank[e] = full[sg * ((size_t)W * W * W) + ((size_t)ix * W + iy) * W + (n + l)];
  }
}

// ------------------------------------------------------------------------------------------
// K_xf: cross-spectrum + pruned, symmetry-reduced 3-D DFT + |.| + arg-max + parabola.
//
//   f[d] = sum_k C[k] exp(-2 pi i (k_idx . d)/F),  k_idx = k + n  (reference: the k=-n corner
//   sits at index 0)  =>  f[d] = phase(d) * g[d],  g[d] = sum_k C[k] w^(k.d) REAL
//   because C(-k) = conj(C(k)); the reference only uses |f| = |g|.
//
// Stages (all in shared memory; T = trig table of 2 pi t/F):
//   X  U[dx][iy][l] = sum_mx C[mx][iy][l] w^(mx dx)          (2n+1 -> F, complex)
//   per slab dx (one group of XF_THREADS / xf_ng threads each):
//   Y  V[dy][l]     = sum_my U[dx][my][l] w^(my dy)          (2n+1 -> F, complex)
//   Z  g[dy][dz]    = V[dy][0] + 2 sum_{l>=1} Re(V[dy][l] w^(l dz))   (n+1 complex -> F real)
// Each 1-D transform uses the +-m pairing (E = c_m + c_-m with cos, O = c_m - c_-m with sin) so
// that outputs d and F-d share all products: a quarter of the dense-DFT flops, same order as a
// radix FFT at these lengths, with no bit reversal and exact handling of any F.
// ------------------------------------------------------------------------------------------
constexpr int XF_THREADS = 512;
constexpr int XF_NG = 4;                       // slab groups (upper bound; fewer for large F (n+1))
constexpr int XF_DCX = 11;                     // outputs per thread, stage X
constexpr int XF_DCY = 3;                      // stage Y
constexpr int XF_DCZ = 7;                      // stage Z

struct XfOut {
  long long* best_idx;  // [P,3]
  double* best_val;     // [P]
  double* frac_idx;     // [P,3]
  double* grid;         // [P,F,F,F] or null
  int* status;          // [P] or null
};

__device__ __forceinline__ void group_sync(int grp, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void better32(double& bv, int& bi, double v, int i) {
  if (v > bv || (v == bv && i < bi)) {
    bv = v;
    bi = i;
  }
}

__device__ __forceinline__ void better(double& bv, long long& bi, double v, long long i) {
  if (v > bv || (v == bv && i < bi)) {
    bv = v;
    bi = i;
  }
}

__global__ void __launch_bounds__(XF_THREADS, 1)
per_xf_kernel(const double2* __restrict__ bankA, const double2* __restrict__ bankB,
              const long long* __restrict__ pairs,  // [P,2] or null (then pair p = (p, p))
              int npairs, int ngroups, int n, int F, double kx, double ky, double kz, double sigma,
              double2* __restrict__ gscratch,  // null: C and U live in shared memory
              int xf_ng,                       // slab groups (4, 2 or 1: fewer when F (n+1) is large)
              XfOut out) {
  extern __shared__ double2 sm[];
  const int W = 2 * n + 1, M = n + 1, Mp = M | 1;
  const int WM = W * M;        // lines of stage X
  const int H = F / 2 + 1;     // outputs d = 0..F/2, partner F-d
  const size_t c_elems = (size_t)W * WM, u_elems = (size_t)F * WM;
  // shared layout: tw[F] | damp[3W doubles -> ceil] | V[NG][F*Mp] | red | (C | U)
  double2* tw = sm;
  double* damp = (double*)(tw + F);
  double2* Vall = (double2*)(damp + ((3 * W + 1) & ~1));
  double* red = (double*)(Vall + (size_t)xf_ng * F * Mp);  // 64 doubles of reduction scratch
  double2* C;
  double2* U;
  if (gscratch) {
    C = gscratch + (size_t)blockIdx.x * (c_elems + u_elems);
    U = C + c_elems;
  } else {
    C = (double2*)(red + 64);
    U = C + c_elems;
  }
  const int tid = threadIdx.x;
  const int xf_gt = XF_THREADS / xf_ng;  // threads per slab group
  const int grp = tid / xf_gt, gtid = tid - grp * xf_gt;
  double2* V = Vall + (size_t)grp * F * Mp;

  for (int t = tid; t < F; t += XF_THREADS) {
    double sn, cs;
    sincospi(2.0 * (double)t / (double)F, &sn, &cs);
    tw[t] = make_double2(cs, sn);
  }
  for (int t = tid; t < 3 * W; t += XF_THREADS) {
    const int ax = t / W, m = t - ax * W - n;
    const double k = (ax == 0 ? kx : (ax == 1 ? ky : kz)) * (double)m;
    damp[t] = exp(-(k * k) * (sigma * sigma));
  }
  __syncthreads();

  const size_t bank_stride = (size_t)ngroups * c_elems;
  for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const long long ia = pairs ? pairs[2 * pair] : pair;
    const long long ib = pairs ? pairs[2 * pair + 1] : pair;
    const double2* SA = bankA + (size_t)ia * bank_stride;
    const double2* SB = bankB + (size_t)ib * bank_stride;
    // ---- cross spectrum C = sum_g SA conj(SB) * damp
    for (int e = tid; e < (int)c_elems; e += XF_THREADS) {
      double re = 0.0, im = 0.0;
      for (int g = 0; g < ngroups; ++g) {
        const double2 a = SA[(size_t)g * c_elems + e];
        const double2 b = SB[(size_t)g * c_elems + e];
        re += a.x * b.x + a.y * b.y;
        im += a.y * b.x - a.x * b.y;
      }
      const int l = e % M;
      const int r = e / M;
      const int iy = r % W, ix = r / W;
      const double dmp = damp[ix] * damp[W + iy] * damp[2 * W + n + l];
      C[e] = make_double2(re * dmp, im * dmp);
    }
    __syncthreads();
    // ---- stage X: lines = (iy,l), item = (line, chunk of XF_DCX outputs)
    {
      const int nch = (H + XF_DCX - 1) / XF_DCX;
      for (int item = tid; item < WM * nch; item += XF_THREADS) {
        const int line = item % WM, ch = item / WM;
        const int d0 = ch * XF_DCX;
        const double2* cin = C + line;
        const double2 c0 = cin[(size_t)n * WM];
        double2 P[XF_DCX], Q[XF_DCX];
        int idx[XF_DCX];
#pragma unroll
        for (int t = 0; t < XF_DCX; ++t) {
          P[t] = c0;
          Q[t] = make_double2(0.0, 0.0);
          idx[t] = 0;
        }
        for (int m = 1; m <= n; ++m) {
          const double2 a = cin[(size_t)(n + m) * WM], b = cin[(size_t)(n - m) * WM];
          const double2 E = cadd(a, b), O = csub(a, b);
#pragma unroll
          for (int t = 0; t < XF_DCX; ++t) {
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
        for (int t = 0; t < XF_DCX; ++t) {
          const int d = d0 + t;
          if (d < H) {
            U[(size_t)d * WM + line] = make_double2(P[t].x + Q[t].y, P[t].y - Q[t].x);
            if (d != 0 && 2 * d != F)
              U[(size_t)(F - d) * WM + line] = make_double2(P[t].x - Q[t].y, P[t].y + Q[t].x);
          }
        }
      }
    }
    __syncthreads();
    // ---- slabs: group grp takes dx = grp, grp + NG, ...
    double bv = -1.0;
    long long bi = 0x7fffffffffffffffLL;
    const int nslab_iter = (F + xf_ng - 1) / xf_ng;
    for (int it = 0; it < nslab_iter; ++it) {
      const int dx = it * xf_ng + grp;
      if (dx < F) {
        const double2* Us = U + (size_t)dx * WM;
        // stage Y: lines l, input over iy (stride M), output V[dy][l]
        const int nchy = (H + XF_DCY - 1) / XF_DCY;
        for (int item = gtid; item < M * nchy; item += xf_gt) {
          const int l = item % M, ch = item / M;
          const int d0 = ch * XF_DCY;
          const double2* cin = Us + l;
          const double2 c0 = cin[(size_t)n * M];
          double2 P[XF_DCY], Q[XF_DCY];
          int idx[XF_DCY];
#pragma unroll
          for (int t = 0; t < XF_DCY; ++t) {
            P[t] = c0;
            Q[t] = make_double2(0.0, 0.0);
            idx[t] = 0;
          }
          for (int m = 1; m <= n; ++m) {
            const double2 a = cin[(size_t)(n + m) * M], b = cin[(size_t)(n - m) * M];
            const double2 E = cadd(a, b), O = csub(a, b);
#pragma unroll
            for (int t = 0; t < XF_DCY; ++t) {
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
          for (int t = 0; t < XF_DCY; ++t) {
            const int d = d0 + t;
            if (d < H) {
              V[(size_t)d * Mp + l] = make_double2(P[t].x + Q[t].y, P[t].y - Q[t].x);
              if (d != 0 && 2 * d != F)
                V[(size_t)(F - d) * Mp + l] = make_double2(P[t].x - Q[t].y, P[t].y + Q[t].x);
            }
          }
        }
      }
      group_sync(grp, xf_gt);
      if (dx < F) {
        // stage Z (complex half-line -> real line), fused with |.| and the running arg-max
        const int nchz = (H + XF_DCZ - 1) / XF_DCZ;
        for (int item = gtid; item < F * nchz; item += xf_gt) {
          const int dy = item % F, ch = item / F;
          const int d0 = ch * XF_DCZ;
          const double2* vin = V + (size_t)dy * Mp;
          const double v0 = vin[0].x;
          double A[XF_DCZ], B[XF_DCZ];
          int idx[XF_DCZ];
#pragma unroll
          for (int t = 0; t < XF_DCZ; ++t) {
            A[t] = 0.0;
            B[t] = 0.0;
            idx[t] = 0;
          }
          for (int l = 1; l <= n; ++l) {
            const double2 v = vin[l];
#pragma unroll
            for (int t = 0; t < XF_DCZ; ++t) {
              int k = idx[t] + d0 + t;
              k -= (k >= F) ? F : 0;
              idx[t] = k;
              const double2 w = tw[k];
              A[t] = fma(v.x, w.x, A[t]);
              B[t] = fma(v.y, w.y, B[t]);
            }
          }
          const long long base = ((long long)dx * F + dy) * F;
          double* grow = out.grid ? out.grid + ((size_t)pair * F * F * F + (size_t)base) : nullptr;
#pragma unroll
          for (int t = 0; t < XF_DCZ; ++t) {
            const int d = d0 + t;
            if (d < H) {
              const double a = v0 + 2.0 * A[t], b = 2.0 * B[t];
              const double g1 = fabs(a + b);
              better(bv, bi, g1, base + d);
              if (grow) grow[d] = g1;
              if (d != 0 && 2 * d != F) {
                const double g2 = fabs(a - b);
                better(bv, bi, g2, base + (F - d));
                if (grow) grow[F - d] = g2;
              }
            }
          }
        }
      }
      group_sync(grp, xf_gt);
    }
    // ---- block arg-max (numpy order: first flat index on exact ties)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, bv, off);
      const long long oi = __shfl_down_sync(0xffffffffu, bi, off);
      better(bv, bi, ov, oi);
    }
    long long* redi = (long long*)(red + 32);
    if ((tid & 31) == 0) {
      red[tid >> 5] = bv;
      redi[tid >> 5] = bi;
    }
    __syncthreads();
    if (tid < 32) {
      bv = (tid < XF_THREADS / 32) ? red[tid] : -1.0;
      bi = (tid < XF_THREADS / 32) ? redi[tid] : 0x7fffffffffffffffLL;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const long long oi = __shfl_down_sync(0xffffffffu, bi, off);
        better(bv, bi, ov, oi);
      }
      if (tid == 0) {
        red[0] = bv;
        redi[0] = bi;
      }
    }
    __syncthreads();
    bv = red[0];
    bi = redi[0];
    const bool ok = (bi != 0x7fffffffffffffffLL) && isfinite(bv);
    const int bx = ok ? (int)(bi / ((long long)F * F)) : 0;
    const int by = ok ? (int)((bi / F) % F) : 0;
    const int bz = ok ? (int)(bi % F) : 0;
    __syncthreads();
    // ---- findMax parabola: the six periodic neighbours, evaluated from U (utils.py:327-337)
    {
      const int w = tid >> 5, lane = tid & 31;
      if (w < 6) {
        const int ax = w >> 1, sgn = (w & 1) ? -1 : 1;  // w even: +1 neighbour, odd: -1
        int px = bx, py = by, pz = bz;
        if (ax == 0) px = (bx + sgn + F) % F;
        if (ax == 1) py = (by + sgn + F) % F;
        if (ax == 2) pz = (bz + sgn + F) % F;
        const double2* Us = U + (size_t)px * WM;
        double acc = 0.0;
        for (int e = lane; e < WM; e += 32) {
          const int l = e % M, iy = e / M;
          int k = ((iy - n) * py + l * pz) % F;
          k += (k < 0) ? F : 0;
          const double2 wv = tw[k];
          const double2 u = Us[e];
          const double term = u.x * wv.x + u.y * wv.y;
          acc += (l == 0) ? term : 2.0 * term;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if (lane == 0) red[2 + w] = fabs(acc);
      }
    }
    __syncthreads();
    if (tid == 0) {
      out.best_idx[3 * (size_t)pair + 0] = bx;
      out.best_idx[3 * (size_t)pair + 1] = by;
      out.best_idx[3 * (size_t)pair + 2] = bz;
      out.best_val[pair] = bv;
      const int b3[3] = {bx, by, bz};
      for (int ax = 0; ax < 3; ++ax) {
        const double y1 = red[2 + 2 * ax], y3 = red[2 + 2 * ax + 1], y2 = bv;
        const double d = (y3 - y1) / (2.0 * (2.0 * y2 - y1 - y3));
        out.frac_idx[3 * (size_t)pair + ax] = (double)b3[ax] - d;
      }
      if (out.status) out.status[pair] = ok ? FO_STATUS_OK : FO_STATUS_NONFINITE;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// K_xf tensor-core paths: the stages of per_xf_kernel with every 1-D transform done per REAL row (re / im
// parts are separate rows, inputs in E / O form) on the FP64 tensor pipe (DMMA.8x8x4, SymMma in fo_symdft.cuh),
// twiddle fragments in registers.  per_xf6_kernel (below) is the resident form for the default k-grids,
// per_xf5_kernel the form for fine k-grids.  (per_xf4_kernel, the round-1 form with a shared-memory hand-over
// between stages Y and Z on named half-CTA barriers, was removed: per_xf6 covers every grid it covered and
// is 1.5x faster.  Earlier scalar forms were limited by operand delivery and issue slots, not by the FP64
// pipe: profiles/r01_summary.md.)
// ------------------------------------------------------------------------------------------
// geometry of the +-kx-paired stage-X image of per_cross_kernel (read by per_xf5_kernel): pitch RXp == 8 (mod 16)
struct X4Layout {
  int M, W, RXp;
  X4Layout() {}
  __host__ __device__ int ximg_doubles() const { return 2 * M * RXp; }
};

// Cross-spectrum of one pair in the E/O form stage X consumes (a10, periodicAlignment.py:433-438 /
// fastbulk.f90:441-454,667-683): C[k] = sum_g SA_g[k] conj(SB_g[k]) exp(-|k|^2 sigma^2), then for
// +-kx: E = C(+) + C(-), O = C(+) - C(-), written as the shared-memory image [E | O][kx = 0..n][RXp]
// of per_xf5_kernel.  This is the only phase that reads the structure-factor bank (231 KB per
// BLJ256 pair from HBM); it used to be phase 1 of the transform kernel, where ncu showed its load
// latency exposed (one CTA per SM) -- as a separate kernel it runs at full occupancy and the
// transform kernel reads the image of its next pair from L2.
__global__ void __launch_bounds__(256, 4)
per_cross_kernel(const __grid_constant__ X4Layout L, const double2* __restrict__ bankA,
                 const double2* __restrict__ bankB, const long long* __restrict__ pairs, int ngroups, int n,
                 double kx, double ky, double kz, double sigma, double* __restrict__ ximg) {
  __shared__ double damp[3 * 129];
  const int M = L.M, W = L.W, RXp = L.RXp;
  const int tid = threadIdx.x;
  const size_t pair = blockIdx.x;
  for (int t = tid; t < 3 * W; t += blockDim.x) {
    const int ax = t / W, m = t - ax * W - n;
    const double k = (ax == 0 ? kx : (ax == 1 ? ky : kz)) * (double)m;
    damp[t] = exp(-(k * k) * (sigma * sigma));
  }
  __syncthreads();
  const size_t c_elems = (size_t)W * W * M;
  const size_t bank_stride = (size_t)ngroups * c_elems;
  const long long ia = pairs ? pairs[2 * pair] : (long long)pair;
  const long long ib = pairs ? pairs[2 * pair + 1] : (long long)pair;
  const double2* SA = bankA + (size_t)ia * bank_stride;
  const double2* SB = bankB + (size_t)ib * bank_stride;
  double* XE = ximg + pair * (size_t)L.ximg_doubles();
  double* XO = XE + (size_t)M * RXp;
  for (int item = tid; item < M * W * M; item += blockDim.x) {
    const int l = item % M;
    const int iy = (item / M) % W;
    const int i = item / (M * W);
    const size_t ep = ((size_t)(n + i) * W + iy) * M + l, em = ((size_t)(n - i) * W + iy) * M + l;
    double pr = 0.0, pi = 0.0, mr = 0.0, mi = 0.0;
    for (int gq = 0; gq < ngroups; ++gq) {
      const double2 a = SA[(size_t)gq * c_elems + ep], b = SB[(size_t)gq * c_elems + ep];
      pr += a.x * b.x + a.y * b.y;
      pi += a.y * b.x - a.x * b.y;
      if (i) {
        const double2 a2 = SA[(size_t)gq * c_elems + em], b2 = SB[(size_t)gq * c_elems + em];
        mr += a2.x * b2.x + a2.y * b2.y;
        mi += a2.y * b2.x - a2.x * b2.y;
      }
    }
    const double dmp = damp[n + i] * damp[W + iy] * damp[2 * W + n + l];
    pr *= dmp; pi *= dmp; mr *= dmp; mi *= dmp;
    const int j = iy >= n ? iy - n : n - iy;
    const int s = iy >= n ? 0 : 1;
    const int row = ((j * M + l) * 2 + s) * 2;
    double er, ei, orr, oi;
    if (i == 0) {
      er = pr; ei = pi; orr = 0.0; oi = 0.0;
    } else {
      er = pr + mr; ei = pi + mi; orr = pr - mr; oi = pi - mi;
    }
    *reinterpret_cast<double2*>(XE + (size_t)i * RXp + row) = make_double2(er, ei);
    *reinterpret_cast<double2*>(XO + (size_t)i * RXp + row) = make_double2(orr, oi);
    if (j == 0) {  // ky = 0 has no s = 1 partner: keep those rows finite (their output is unused)
      *reinterpret_cast<double2*>(XE + (size_t)i * RXp + row + 2) = make_double2(er, ei);
      *reinterpret_cast<double2*>(XO + (size_t)i * RXp + row + 2) = make_double2(orr, oi);
    }
  }
}

// ------------------------------------------------------------------------------------------
// per_cross6_kernel + per_xf6_kernel: the pruned 3-D DFT + |.| + arg-max with every stage in the "transposed" DMMA
// form (twiddles = A operand, data = B operand) and stages Y and Z chained in REGISTERS.
//
// * Transposed form.  The C fragment of a tile holds (row d = 8 mt + g, columns 2 t, 2 t + 1) = (re, im) of ONE
//   complex entry, so P - iQ is formed inside a lane -- and it is never formed by FP64 adds: the previous stage
//   stores the odd part already multiplied by -i (O' = -iO, a swap and a sign folded into a subtraction order),
//       U[d] = c0 + sum_m cos(m d) E[m] + sin(m d) O'[m]        (one DMMA chain: P, then V = P + sin O')
//       U[F - d] = 2 P - U[d]                                   (one DFMA per element)
//   ncu on the first version (profiles/r02_summary.md): a scalar FP64 instruction issued while other warps stream
//   DMMAs waits ~45 cycles for the shared FP64 pipe, 28 % of all warp time; what is left are those DFMAs.
// * The +-ky pairing of stage Y's input is linear, so it is applied to the cross-spectrum (per_cross6_kernel writes
//   C_E = C(+ky) + C(-ky), C_O = -i (C(+ky) - C(-ky)) in the +-kx-paired form): stage X is a plain GEMM
//   [F/2+1 rows dx] x [K2 RY columns = YIN slab] with no shuffles, its C fragments go to YIN as 16-byte stores.
// * Stage Y -> Z: a lane's C fragment of row tile mt, column tile ct holds V at (row dy, l-slot 4 ct + t + 1, re | im)
//   = exactly the A fragment (row g, k = t) of stage Z's k-step ks = ct.  The Z-stage DMMAs take the Y-stage
//   accumulators as operands: nothing is written to shared memory between Y and Z, no warp waits for another one
//   in the slab phase (round 1: 14 half-CTA barriers per pair), and the 40 KB Z-input buffers are gone.  l-slot
//   s = 1..n is harmonic l = s; l = 0 rides in slot n + 1, whose X / Y twiddles are zero (m > K) and whose Z
//   twiddle is the weight of the l = 0 term, so all three stages use ONE register-resident twiddle table
//   (cos / sin (2 pi m d / F) at m = 4 ks + t + 1, d = 8 nt + g is symmetric in the roles of m and d).
// * Arg-max filter on integer pipes: max(|A + B|, |A - B|) = |A| + |B| <= 2 max(|A|, |B|), so a tile whose largest
//   |A|, |B| high word (sign shifted out) is more than one binade below the running maximum is skipped without
//   touching the FP64 pipe; the running maximum is shared as a high word (native 32-bit shared atomic).
// * Twelve warps (three per scheduler, 168 registers): stage-X items (pair of column tiles, row tile) and slab
//   items (slab, row tile) are dealt round-robin -- n = 9, F = 40: 72 and 120 items, 6 and 10 per warp.
// * The image of the next pair arrives by one bulk asynchronous copy (cp.async.bulk, SASS UBLKCP) completing on
//   an mbarrier: no thread spends issue slots on it.
// ------------------------------------------------------------------------------------------
constexpr int X6_WARPS = 12;
constexpr int X6_THREADS = X6_WARPS * 32;

struct X6Layout {
  int M, H, RY, K2, NC, NCP, RXp, SP;
  int o_red, o_tw, o_x, o_yin, total;  // in doubles
  X6Layout() {}
  X6Layout(int n, int F) {
    M = n + 1;
    H = F / 2 + 1;
    RY = 2 * M;
    K2 = 2 * n + 1;
    NC = K2 * RY;                // columns of stage X = doubles of one YIN slab [K2][RY]
    NCP = (NC + 15) / 16;        // pairs of 8-column tiles
    RXp = 16 * NCP + 8;          // image row pitch, == 8 (mod 16): conflict-free B fragments
    SP = RXp;                    // YIN slab pitch, == 8 (mod 16): conflict-free 16-byte C-fragment stores
    o_red = 0;                   // 64 doubles: reduction scratch, running maximum, mbarrier
    o_tw = 64;                   // F double2: twiddles of the parabola neighbours
    o_x = o_tw + 2 * F;          // image [E | O'][M][RXp]
    o_yin = o_x + ximg_doubles();
    total = o_yin + F * SP;
  }
  __host__ __device__ int ximg_doubles() const { return 2 * M * RXp; }
};

// Cross-spectrum of one pair in the form stage X of per_xf6_kernel consumes (a10, periodicAlignment.py:433-438 /
// fastbulk.f90:441-454,667-683): C[k] = sum_g SA_g[k] conj(SB_g[k]) exp(-|k|^2 sigma^2); paired over +-ky into
// C_E = C(+j) + C(-j) (row j), C_O = -i (C(+j) - C(-j)) (row n + j), each then paired over +-kx into
// E[m] = C_s(+m) + C_s(-m), O'[m] = -i (C_s(+m) - C_s(-m)); image [E | O'][m = 0..n][RXp], column = offset in a
// YIN slab = (row, l, re | im).
__global__ void __launch_bounds__(256, 3)
per_cross6_kernel(const __grid_constant__ X6Layout L, const double2* __restrict__ bankA,
                  const double2* __restrict__ bankB, const long long* __restrict__ pairs,
                  const int32_t* __restrict__ ops, int ngroups, int n, double kx, double ky, double kz, double sigma,
                  double* __restrict__ ximg) {
  __shared__ double damp[3 * 129];
  const int M = L.M, W = 2 * n + 1, RXp = L.RXp, RY = L.RY;
  const int tid = threadIdx.x;
  const size_t pair = blockIdx.x;
  for (int t = tid; t < 3 * W; t += blockDim.x) {
    const int ax = t / W, m = t - ax * W - n;
    const double k = (ax == 0 ? kx : (ax == 1 ? ky : kz)) * (double)m;
    damp[t] = exp(-(k * k) * (sigma * sigma));
  }
  __syncthreads();
  const size_t c_elems = (size_t)W * W * M;
  const size_t bank_stride = (size_t)ngroups * c_elems;
  const long long ia = pairs ? pairs[2 * pair] : (long long)pair;
  const long long ib = pairs ? pairs[2 * pair + 1] : (long long)pair;
  const double2* SA = bankA + (size_t)ia * bank_stride;
  const double2* SB = bankB + (size_t)ib * bank_stride;
  double* XE = ximg + pair * (size_t)L.ximg_doubles();
  double* XO = XE + (size_t)M * RXp;
  // Cell symmetry of a cubic box (O_h, fastbulk.f90:863-1380 OHTRANSFORMCOEFFS): structure B is taken as R B with
  // the signed permutation (R r)_i = s_i r_{p_i}.  S_{RB}(k) = S_B(R^T k), (R^T k)_{p_i} = s_i k_i: an index
  // permutation of B's bank entry (conjugated where the image has kz < 0: the bank holds kz >= 0), no new
  // structure factors.  op = p_0 | p_1 << 2 | p_2 << 4 | (s_i < 0) << (6 + i); 0x24 = identity.
  const int op = ops ? ops[pair] : 0x24;
  const int opp[3] = {op & 3, (op >> 2) & 3, (op >> 4) & 3};
  const int ops_[3] = {(op >> 6) & 1, (op >> 7) & 1, (op >> 8) & 1};
  auto cross = [&](int ix, int iy, int l) {
    const size_t e = ((size_t)ix * W + iy) * M + l;
    size_t eb = e;
    double cj = 1.0;
    if (ops) {
      const int k[3] = {ix - n, iy - n, l};
      int kb[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int v = ops_[i] ? -k[i] : k[i];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (opp[i] == j) kb[j] = v;
      }
      if (kb[2] < 0) {
        kb[0] = -kb[0];
        kb[1] = -kb[1];
        kb[2] = -kb[2];
        cj = -1.0;
      }
      eb = ((size_t)(kb[0] + n) * W + (kb[1] + n)) * M + kb[2];
    }
    double re = 0.0, im = 0.0;
    for (int gq = 0; gq < ngroups; ++gq) {
      const double2 a = SA[(size_t)gq * c_elems + e];
      double2 b = SB[(size_t)gq * c_elems + eb];
      b.y *= cj;
      re += a.x * b.x + a.y * b.y;
      im += a.y * b.x - a.x * b.y;
    }
    const double dmp = damp[ix] * damp[W + iy] * damp[2 * W + n + l];
    return make_double2(re * dmp, im * dmp);
  };
  for (int item = tid; item < M * M * M; item += blockDim.x) {
    const int l = item % M;
    const int j = (item / M) % M;
    const int m = item / (M * M);
    const double2 pp = cross(n + m, n + j, l);
    double2 pm = pp, mp = pp, mm = pp;  // (kx sign, ky sign)
    if (j) pm = cross(n + m, n - j, l);
    if (m) mp = cross(n - m, n + j, l);
    if (m && j) mm = cross(n - m, n - j, l);
    // ky pairing at kx = +m (cp) and kx = -m (cm): s = E, O
    double2 ep = pp, em = mp, op = make_double2(0.0, 0.0), om = op;
    if (j) {
      ep = make_double2(pp.x + pm.x, pp.y + pm.y);
      em = make_double2(mp.x + mm.x, mp.y + mm.y);
      op = make_double2(pp.y - pm.y, pm.x - pp.x);  // -i (pp - pm)
      om = make_double2(mp.y - mm.y, mm.x - mp.x);
    }
    const int colE = j * RY + 2 * l, colO = (n + j) * RY + 2 * l;
    if (m == 0) {
      *reinterpret_cast<double2*>(XE + colE) = ep;
      if (j) *reinterpret_cast<double2*>(XE + colO) = op;
    } else {
      *reinterpret_cast<double2*>(XE + (size_t)m * RXp + colE) = make_double2(ep.x + em.x, ep.y + em.y);
      *reinterpret_cast<double2*>(XO + (size_t)m * RXp + colE) = make_double2(ep.y - em.y, em.x - ep.x);
      if (j) {
        *reinterpret_cast<double2*>(XE + (size_t)m * RXp + colO) = make_double2(op.x + om.x, op.y + om.y);
        *reinterpret_cast<double2*>(XO + (size_t)m * RXp + colO) = make_double2(op.y - om.y, om.x - op.x);
      }
    }
  }
}

template <int KS, int NT, bool WANT_GRID>
__global__ void __launch_bounds__(X6_THREADS, 1)
per_xf6_kernel(const __grid_constant__ X6Layout L, const double* __restrict__ ximg, int npairs, int n, int F,
               XfOut out) {
  extern __shared__ double sm6[];
  const int M = L.M, H = L.H, RXp = L.RXp, RY = L.RY, SP = L.SP, NCP = L.NCP;
  double* red = sm6 + L.o_red;
  double2* twz = reinterpret_cast<double2*>(sm6 + L.o_tw);
  double* XE = sm6 + L.o_x;           // [M][RXp] (row 0: c0)
  double* XO = XE + (size_t)M * RXp;  // [M][RXp] (row 0 unused)
  double* YIN = sm6 + L.o_yin;        // [F][SP]: slab = [K2][RY], row j: c0 / E, row n + j: O'
  int* sbest = reinterpret_cast<int*>(red + 48);
  uint64_t* xbar = reinterpret_cast<uint64_t*>(red + 56);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const unsigned ximg_bytes = (unsigned)L.ximg_doubles() * 8u;

  if (tid == 0) {
    fo_mbar_init(xbar, 1);
    fo_mbar_fence_init();
    fo_mbar_arrive_expect_tx(xbar, ximg_bytes);
    fo_bulk_g2s(XE, ximg + (size_t)blockIdx.x * L.ximg_doubles(), ximg_bytes, xbar);
  }
  for (int t = tid; t < F; t += X6_THREADS) {
    double sn, cs;
    sincospi(2.0 * (double)t / (double)F, &sn, &cs);
    twz[t] = make_double2(cs, sn);
  }
  SymMma<KS, NT> mm;
  mm.init(n, F, H, lane);
  // stage-Z cos fragment of the k-step that carries l = 0 (slot n + 1): weight 1/2 at half scale
  const int ks0 = n >> 2, t0 = n & 3;
  double bcz[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    bcz[nt] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      if (ks == ks0) bcz[nt] = (t4 == t0 && nt * 8 + g < H) ? 0.5 : mm.bc[ks][nt];
  }
  // per-lane geometry: B-fragment rows (ks, t4) of stages X / Y, stage-Y B-fragment column (ct, g) and C-init
  // column pair (ct, t4)
  int ycol[KS], ycol0[KS], yrow[KS], xrow[KS];
#pragma unroll
  for (int ct = 0; ct < KS; ++ct) {
    const int sb = 4 * ct + (g >> 1) + 1, sc = 4 * ct + t4 + 1;
    ycol[ct] = 2 * (sb <= n ? sb : 0) + (g & 1);
    ycol0[ct] = 2 * (sc <= n ? sc : 0);
    const int j = 4 * ct + t4 + 1;
    yrow[ct] = (j <= n ? j : n) * RY;
    xrow[ct] = (j <= n ? j : n) * RXp;
  }
  int xph = 0;
  __syncthreads();  // the initialised mbarrier is visible to every waiter

  for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    if (tid == 0) *sbest = 0;  // high word of +0.0: |f| >= 0, and only strictly smaller tiles are filtered
    fo_mbar_wait(xbar, xph);
    xph ^= 1;
    __syncthreads();
    // ---- stage X: item = (pair of column tiles cp, row tile mt); rows dx = 8 mt + g and their mirrors F - dx
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
      const int dx0 = 8 * mt + g;
      const bool v0 = dx0 < H, v1 = v0 && dx0 != 0 && 2 * dx0 != F;
      for (int cp = (warp + X6_WARPS - (mt * NCP) % X6_WARPS) % X6_WARPS; cp < NCP; cp += X6_WARPS) {
        const double* xe = XE + 16 * cp;
        const double* xo = XO + 16 * cp;
        double be[KS][2], bo[KS][2], P[2][2], V[2][2];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            be[ks][c] = xe[xrow[ks] + 8 * c + g];
            bo[ks][c] = xo[xrow[ks] + 8 * c + g];
          }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const double2 z = *reinterpret_cast<const double2*>(xe + 8 * c + 2 * t4);
          fo_dmma3(P[c], mm.bc[0][mt], be[0][c], z.x, z.y);
        }
#pragma unroll
        for (int ks = 1; ks < KS; ++ks)
#pragma unroll
          for (int c = 0; c < 2; ++c) fo_dmma(P[c], mm.bc[ks][mt], be[ks][c]);
#pragma unroll
        for (int c = 0; c < 2; ++c) fo_dmma3(V[c], mm.bs[0][mt], bo[0][c], P[c][0], P[c][1]);
#pragma unroll
        for (int ks = 1; ks < KS; ++ks)
#pragma unroll
          for (int c = 0; c < 2; ++c) fo_dmma(V[c], mm.bs[ks][mt], bo[ks][c]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = 16 * cp + 8 * c + 2 * t4;
          if (v0) *reinterpret_cast<double2*>(YIN + (size_t)dx0 * SP + col) = make_double2(V[c][0], V[c][1]);
          if (v1)
            *reinterpret_cast<double2*>(YIN + (size_t)(F - dx0) * SP + col) =
                make_double2(fma(2.0, P[c][0], -V[c][0]), fma(2.0, P[c][1], -V[c][1]));
        }
      }
    }
    __syncthreads();
    if (tid == 0 && pair + (int)gridDim.x < npairs) {  // next pair's image lands during the slab phase
      fo_fence_proxy_async();
      fo_mbar_arrive_expect_tx(xbar, ximg_bytes);
      fo_bulk_g2s(XE, ximg + (size_t)(pair + gridDim.x) * L.ximg_doubles(), ximg_bytes, xbar);
    }
    // ---- slabs: stage Y -> stage Z in registers, arg-max at half scale: the two outputs of a column are
    // 2 |A + B| and 2 |A - B|, their maximum is 2 (|A| + |B|).
    // Work item = (slab dx, row tile mt), dealt round-robin to the warps.
    long long bbits = __double_as_longlong(-1.0);  // running maximum (half scale) as a bit pattern
    int bi = 0x7fffffff;
    int thr = 0;  // high word of max(bvh, CTA lower bound): tiles strictly below it cannot hold the maximum
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
      const int dy0 = 8 * mt + g;
      const bool valid0 = dy0 < H, valid1 = valid0 && dy0 != 0 && 2 * dy0 != F;
      for (int dx = (warp + X6_WARPS - (mt * F) % X6_WARPS) % X6_WARPS; dx < F; dx += X6_WARPS) {
        const double* Y = YIN + (size_t)dx * SP;
        double P[KS][2], V[KS][2];
        {
          double be[KS][KS], bo[KS][KS];
#pragma unroll
          for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int ct = 0; ct < KS; ++ct) {
              be[ks][ct] = Y[yrow[ks] + ycol[ct]];
              bo[ks][ct] = Y[yrow[ks] + n * RY + ycol[ct]];
            }
#pragma unroll
          for (int ct = 0; ct < KS; ++ct) {
            const double2 c = *reinterpret_cast<const double2*>(Y + ycol0[ct]);
            fo_dmma3(P[ct], mm.bc[0][mt], be[0][ct], c.x, c.y);
          }
#pragma unroll
          for (int ks = 1; ks < KS; ++ks)
#pragma unroll
            for (int ct = 0; ct < KS; ++ct) fo_dmma(P[ct], mm.bc[ks][mt], be[ks][ct]);
#pragma unroll
          for (int ct = 0; ct < KS; ++ct) fo_dmma3(V[ct], mm.bs[0][mt], bo[0][ct], P[ct][0], P[ct][1]);
#pragma unroll
          for (int ks = 1; ks < KS; ++ks)
#pragma unroll
            for (int ct = 0; ct < KS; ++ct) fo_dmma(V[ct], mm.bs[ks][mt], bo[ks][ct]);
        }
        thr = max(thr, *sbest);
        // V[dy] (tile 0) and V[F - dy] = 2 P - V[dy] (tile 1) are the A fragments of stage Z
        double W2[KS][2];
#pragma unroll
        for (int ct = 0; ct < KS; ++ct) {
          W2[ct][0] = fma(2.0, P[ct][0], -V[ct][0]);
          W2[ct][1] = fma(2.0, P[ct][1], -V[ct][1]);
        }
        double A[2][NT][2], B[2][NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double zc = (0 == ks0) ? bcz[nt] : mm.bc[0][nt];
          fo_dmma3(A[0][nt], V[0][0], zc, 0.0, 0.0);
          fo_dmma3(B[0][nt], V[0][1], mm.bs[0][nt], 0.0, 0.0);
          fo_dmma3(A[1][nt], W2[0][0], zc, 0.0, 0.0);
          fo_dmma3(B[1][nt], W2[0][1], mm.bs[0][nt], 0.0, 0.0);
        }
#pragma unroll
        for (int ks = 1; ks < KS; ++ks)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double zc = (ks == ks0) ? bcz[nt] : mm.bc[ks][nt];
            fo_dmma(A[0][nt], V[ks][0], zc);
            fo_dmma(B[0][nt], V[ks][1], mm.bs[ks][nt]);
            fo_dmma(A[1][nt], W2[ks][0], zc);
            fo_dmma(B[1][nt], W2[ks][1], mm.bs[ks][nt]);
          }
        if (WANT_GRID) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const bool valid = h == 0 ? valid0 : valid1;
            const int dy = h == 0 ? dy0 : F - dy0;
            double* grow = out.grid + ((size_t)pair * F * F * F + (size_t)(dx * F + dy) * F);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int d = nt * 8 + t4 * 2 + q;
                if (valid && d < H) {
                  grow[d] = 2.0 * fabs(A[h][nt][q] + B[h][nt][q]);
                  if (d != 0 && 2 * d != F) grow[F - d] = 2.0 * fabs(A[h][nt][q] - B[h][nt][q]);
                }
              }
          }
        }
        // Arg-max filter off the FP64 pipe (a scalar FP64 instruction issued among other warps' DMMAs waits ~45
        // cycles for the pipe).  Non-negative doubles order like their bit patterns, so:
        //  1. the largest |A| and the largest |B| of the lane's 24 accumulators are found as high words with the
        //     sign shifted out (integer max); ONE addition of their upper bounds (high word + 1) bounds every
        //     |A| + |B| = max(|A + B|, |A - B|) of the item from above;
        //  2. an item that may reach the running maximum (high word thr) forms |A| + |B| per column (bit for bit
        //     the larger of the two outputs) and compares bit patterns: the larger output is |A + B| (index d)
        //     when A and B have the same sign or B = 0, else |A - B| (index F - d); ties take the smaller index.
        unsigned ma = 0, mb = 0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            ma = __vimax3_u32(ma, (unsigned)__double2hiint(A[0][nt][q]) << 1, (unsigned)__double2hiint(A[1][nt][q]) << 1);
            mb = __vimax3_u32(mb, (unsigned)__double2hiint(B[0][nt][q]) << 1, (unsigned)__double2hiint(B[1][nt][q]) << 1);
          }
        const double ubound = __hiloint2double(min((int)(ma >> 1) + 1, 0x7ff00000), 0) +
                              __hiloint2double(min((int)(mb >> 1) + 1, 0x7ff00000), 0);
        if (__double2hiint(ubound) >= thr) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const bool valid = h == 0 ? valid0 : valid1;
            const int dy = h == 0 ? dy0 : F - dy0;
            const int base = (dx * F + dy) * F;
            double c[NT][2];
            int chi = 0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                c[nt][q] = fabs(A[h][nt][q]) + fabs(B[h][nt][q]);  // columns d >= H: zero twiddles, A = B = 0
                chi = max(chi, __double2hiint(c[nt][q]));
              }
            if (valid && chi >= thr) {
#pragma unroll
              for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const int d = nt * 8 + t4 * 2 + q;
                  const long long cb = __double_as_longlong(c[nt][q]);
                  const bool plus = ((__double2hiint(A[h][nt][q]) ^ __double2hiint(B[h][nt][q])) >= 0) ||
                                    ((__double_as_longlong(B[h][nt][q]) << 1) == 0);
                  const int idx = base + (plus ? d : F - d);
                  if (d < H && (cb > bbits || (cb == bbits && idx < bi))) {
                    bbits = cb;
                    bi = idx;
                  }
                }
              const int bh = (int)(bbits >> 32);
              if (bh > thr) {
                thr = bh;
                atomicMax(sbest, bh);
              }
            }
          }
        }
      }
    }
    const double bvh = __longlong_as_double(bbits);
    double bv = 2.0 * bvh;
    if (bvh < 0.0) bv = -1.0;
    // ---- block arg-max (numpy order) and parabola neighbours
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, bv, off);
      const int oi = __shfl_down_sync(0xffffffffu, bi, off);
      better32(bv, bi, ov, oi);
    }
    int* redi = reinterpret_cast<int*>(red + 32);
    if (lane == 0) {
      red[warp] = bv;
      redi[warp] = bi;
    }
    __syncthreads();
    if (tid < 32) {
      bv = (tid < X6_WARPS) ? red[tid] : -1.0;
      bi = (tid < X6_WARPS) ? redi[tid] : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        better32(bv, bi, ov, oi);
      }
      if (tid == 0) {
        red[16] = bv;
        *reinterpret_cast<int*>(red + 17) = bi;
      }
    }
    __syncthreads();
    bv = red[16];
    bi = *reinterpret_cast<const int*>(red + 17);
    const bool ok = (bi != 0x7fffffff) && isfinite(bv);
    const int bx = ok ? bi / (F * F) : 0;
    const int by = ok ? (bi / F) % F : 0;
    const int bz = ok ? bi % F : 0;
    if (warp < 6) {
      const int ax = warp >> 1, sgn = (warp & 1) ? -1 : 1;
      int px = bx, py = by, pz = bz;
      if (ax == 0) px = (bx + sgn + F) % F;
      if (ax == 1) py = (by + sgn + F) % F;
      if (ax == 2) pz = (bz + sgn + F) % F;
      const double* Y = YIN + (size_t)px * SP;
      double acc = 0.0;
      for (int e = lane; e < M * M; e += 32) {
        const int j = e / M, l = e - j * M;
        const double2 wj = twz[(j * py) % F], wl = twz[(l * pz) % F];
        double2 v = *reinterpret_cast<const double2*>(Y + l * 2);  // j = 0: U(ky = 0)
        if (j) {  // V = E cos + O' sin
          const double2 ev = *reinterpret_cast<const double2*>(Y + (size_t)j * RY + l * 2);
          const double2 ov = *reinterpret_cast<const double2*>(Y + (size_t)(n + j) * RY + l * 2);
          v = make_double2(ev.x * wj.x + ov.x * wj.y, ev.y * wj.x + ov.y * wj.y);
        }
        const double term = v.x * wl.x + v.y * wl.y;
        acc += (l == 0) ? term : 2.0 * term;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
      if (lane == 0) red[20 + warp] = fabs(acc);
    }
    __syncthreads();
    if (tid == 0) {
      out.best_idx[3 * (size_t)pair + 0] = bx;
      out.best_idx[3 * (size_t)pair + 1] = by;
      out.best_idx[3 * (size_t)pair + 2] = bz;
      out.best_val[pair] = bv;
      const int b3[3] = {bx, by, bz};
      for (int ax = 0; ax < 3; ++ax) {
        const double y1 = red[20 + 2 * ax], y3 = red[20 + 2 * ax + 1], y2 = bv;
        const double d = (y3 - y1) / (2.0 * (2.0 * y2 - y1 - y3));
        out.frac_idx[3 * (size_t)pair + ax] = (double)b3[ax] - d;
      }
      if (out.status) out.status[pair] = ok ? FO_STATUS_OK : FO_STATUS_NONFINITE;
    }
    // no barrier here: the barrier at the top of the loop orders the reuse of red / YIN
  }
}

// ------------------------------------------------------------------------------------------
// K_xf for fine k-grids (per_xf5_kernel<NT>): the stages with a shared-memory hand-over when neither the stage-X
// image nor YIN (646 KB per pair at n = 16, 4.6 MB at n = 32) fit in shared memory.
//   * stage X reads its E / O fragments straight from the image per_cross_kernel wrote (L2) and writes
//     YIN[dx][ky][l] to a per-CTA global scratch;
//   * slabs: YIN[dx] (K2 RY doubles) is staged into shared memory with cp.async, double buffered, SPI
//     slabs at a time; stages Y and Z as SymMma row tiles; arg-max with the high-word filter;
//   * the twiddle matrices live in shared memory (B fragments: one 8-byte load per DMMA) -- with up to
//     8 x 9 fragment pairs they do not fit the register file.
// One persistent 512-thread CTA per SM.  NT = ceil((F/2+1)/8) column tiles (template), any n.
// ------------------------------------------------------------------------------------------
constexpr int X5_THREADS = 512;
constexpr int X5_WARPS = X5_THREADS / 32;

struct X5Layout {
  int M, W, H, KS, FP, RX, RXp, RY, K2, SPI, LDT;
  int o_red, o_tc, o_ts, o_yb, o_z, total;  // shared-memory offsets in doubles
  X5Layout() {}
  X5Layout(int n, int F, int NT, size_t smem_limit_bytes) {
    M = n + 1;
    W = 2 * n + 1;
    H = F / 2 + 1;
    KS = (n + 3) / 4;
    FP = ((F + 3) / 4) * 4 + 2;
    RX = M * M * 4;
    RXp = ((RX + 7) / 16) * 16 + 8;
    RY = 2 * M;
    K2 = 2 * n + 1;
    LDT = 8 * NT;
    LDT += ((8 - LDT) % 32 + 32) % 32;  // == 8 (mod 32)
    o_red = 0;
    o_tc = 64;
    o_ts = o_tc + 4 * KS * LDT;
    o_yb = o_ts + 4 * KS * LDT;
    SPI = 8;
    while (SPI > 1 && (size_t)(o_yb + SPI * (2 * K2 * RY + 2 * M * FP)) * 8 > smem_limit_bytes) --SPI;
    o_z = o_yb + 2 * SPI * K2 * RY;
    total = o_z + SPI * 2 * M * FP;
  }
  __host__ __device__ int zin_per_slab() const { return 2 * M * FP; }
  __host__ __device__ int ximg_doubles() const { return 2 * M * RXp; }
  __host__ __device__ size_t yin_doubles(int F) const { return (size_t)F * K2 * RY; }
};

// P[8 x 8NT] = c0 + E[8 x K] COS[K x 8NT], Q = O SIN with the twiddles in shared memory (TC / TS:
// [4 KS][LDT], zero beyond K harmonics / H outputs).  e1 / o1: E[m = 1][first row of the tile].
template <int NT>
__device__ __forceinline__ void sym_run_smem(const double* __restrict__ e1, const double* __restrict__ o1,
                                             size_t kstride, int K, int KS, const double* __restrict__ TC,
                                             const double* __restrict__ TS, int LDT, double c0, int lane,
                                             double (&P)[NT][2], double (&Q)[NT][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    P[nt][0] = P[nt][1] = c0;
    Q[nt][0] = Q[nt][1] = 0.0;
  }
  for (int ks = 0; ks < KS; ++ks) {
    int m = ks * 4 + t;
    const int mc = m < K ? m : K - 1;  // rows beyond K: twiddles are zero
    const double ae = e1[(size_t)mc * kstride + g];
    const double ao = o1[(size_t)mc * kstride + g];
    const double* tc = TC + m * LDT + g;
    const double* ts = TS + m * LDT + g;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      fo_dmma(P[nt], ae, tc[nt * 8]);
      fo_dmma(Q[nt], ao, ts[nt * 8]);
    }
  }
}

template <int NT, bool WANT_GRID>
__global__ void __launch_bounds__(X5_THREADS, 1)
per_xf5_kernel(const __grid_constant__ X5Layout L, const double* __restrict__ ximg, double* __restrict__ yin_all,
               int npairs, int n, int F, XfOut out) {
  extern __shared__ double sm5[];
  const int M = L.M, H = L.H, KS = L.KS, FP = L.FP, RX = L.RX, RXp = L.RXp, RY = L.RY, K2 = L.K2;
  const int SPI = L.SPI, LDT = L.LDT;
  double* red = sm5 + L.o_red;
  double* TC = sm5 + L.o_tc;
  double* TS = sm5 + L.o_ts;
  double* YB = sm5 + L.o_yb;            // [2][SPI][K2][RY]
  double* ZIN = sm5 + L.o_z;            // [SPI][2][M][FP]
  double* YIN = yin_all + (size_t)blockIdx.x * L.yin_doubles(F);  // [F][K2][RY]: k = j (c0 / E), n + j (O)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  unsigned long long* sbest = reinterpret_cast<unsigned long long*>(red + 48);
  for (int e = tid; e < 4 * KS * LDT; e += X5_THREADS) {
    const int m = e / LDT + 1, d = e - (m - 1) * LDT;
    double sn = 0.0, cs = 0.0;
    if (m <= n && d < H) sincospi(2.0 * (double)((m * d) % F) / (double)F, &sn, &cs);
    TC[e] = cs;
    TS[e] = sn;
  }
  const int slab_d = K2 * RY;  // doubles of one YIN slab
  auto stage_slabs = [&](int x0, int buf) {
    const int ns = min(SPI, F - x0);
    const double2* src = reinterpret_cast<const double2*>(YIN + (size_t)x0 * slab_d);
    double2* dst = reinterpret_cast<double2*>(YB + (size_t)buf * SPI * slab_d);
    for (int e = tid; e < ns * (slab_d >> 1); e += X5_THREADS) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + e);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  __syncthreads();

  for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    if (tid == 0) *sbest = 0ull;
    const double* XE = ximg + (size_t)pair * L.ximg_doubles();  // [M][RXp] (index 0: c0)
    const double* XO = XE + (size_t)M * RXp;
    // ---- stage X; tile = 8 rows; row bits: 0 = part, 1 = s  (partners: lane ^ 4, lane ^ 8)
    for (int tile = warp; tile * 8 < RX; tile += X5_WARPS) {
      const int row = tile * 8 + g;
      const bool valid = row < RX;
      const int r = valid ? row : 0;
      const int part = r & 1, s = (r >> 1) & 1, jl = r >> 2;
      const int j = jl / M, l = jl - j * M;
      const double sgn = part ? -1.0 : 1.0;
      const int krow = (s == 0) ? j : n + j;  // s = 0 lanes store c0 / E, s = 1 lanes store O
      const bool store = valid && !(j == 0 && s == 1);
      double P[NT][2], Q[NT][2];
      // rows of a partial last tile read the (finite) padding of the image; their outputs are not stored
      sym_run_smem<NT>(XE + RXp + tile * 8, XO + RXp + tile * 8, RXp, n, KS, TC, TS, LDT, XE[r], lane, P, Q);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int d = nt * 8 + t4 * 2 + q;
          const double qx = __shfl_xor_sync(0xffffffffu, Q[nt][q], 4);
          const double ud = fma(sgn, qx, P[nt][q]);   // U[d]   = P - iQ
          const double um = fma(-sgn, qx, P[nt][q]);  // U[F-d] = P + iQ
          const double xd = __shfl_xor_sync(0xffffffffu, ud, 8);
          const double xm = __shfl_xor_sync(0xffffffffu, um, 8);
          const double vd = (j == 0) ? ud : (s == 0 ? ud + xd : xd - ud);
          const double vm = (j == 0) ? um : (s == 0 ? um + xm : xm - um);
          if (store && d < H) {
            YIN[((size_t)d * K2 + krow) * RY + l * 2 + part] = vd;
            if (d != 0 && 2 * d != F) YIN[((size_t)(F - d) * K2 + krow) * RY + l * 2 + part] = vm;
          }
        }
    }
    __syncthreads();  // YIN complete (global writes of this CTA are visible to it after the barrier)
    stage_slabs(0, 0);
    // ---- slabs, SPI at a time
    double bvh = -1.0;
    int bi = 0x7fffffff;
    int buf = 0;
    for (int x0 = 0; x0 < F; x0 += SPI, buf ^= 1) {
      const int ns = min(SPI, F - x0);
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();  // slabs of this iteration staged; ZIN and the other buffer free
      if (x0 + SPI < F) stage_slabs(x0 + SPI, buf ^ 1);
      const double* Yb = YB + (size_t)buf * SPI * slab_d;
      // stage Y: rows (slab, l, part)
      for (int tile = warp; tile * 8 < ns * RY; tile += X5_WARPS) {
        const int row = tile * 8 + g;
        const bool valid = row < ns * RY;
        const int r = valid ? row : 0;
        const int sl = r / RY, lp = r - sl * RY;
        const int l = lp >> 1, part = lp & 1;
        const double* Y = Yb + (size_t)sl * slab_d + lp;
        double* Zrow = ZIN + (size_t)sl * L.zin_per_slab() + (size_t)part * M * FP + (size_t)l * FP;
        const double sgn = part ? -1.0 : 1.0;
        double P[NT][2], Q[NT][2];
        // per-lane rows (a tile can straddle two slabs): e1 / o1 are given per lane, g offset removed
        sym_run_smem<NT>(Y + RY - g, Y + (size_t)(n + 1) * RY - g, RY, n, KS, TC, TS, LDT, Y[0], lane, P, Q);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int d = nt * 8 + t4 * 2 + q;
            const double qx = __shfl_xor_sync(0xffffffffu, Q[nt][q], 4);
            if (valid && d < H) {
              Zrow[d] = fma(sgn, qx, P[nt][q]);
              if (d != 0 && 2 * d != F) Zrow[F - d] = fma(-sgn, qx, P[nt][q]);
            }
          }
      }
      __syncthreads();
      // stage Z: rows (slab, dy)
      for (int tile = warp; tile * 8 < ns * F; tile += X5_WARPS) {
        const int row = tile * 8 + g;
        const bool valid = row < ns * F;
        const int r = valid ? row : 0;
        const int sl = r / F, dy = r - sl * F;
        const double sb = __longlong_as_double(*sbest);
        const double* ZR = ZIN + (size_t)sl * L.zin_per_slab() + dy;
        const double* ZI = ZR + (size_t)M * FP;
        const int base = ((x0 + sl) * F + dy) * F;
        const double v0h = 0.5 * ZR[0];
        double A[NT][2], B[NT][2];
        sym_run_smem<NT>(ZR + FP - g, ZI + FP - g, FP, n, KS, TC, TS, LDT, v0h, lane, A, B);
        int chi = 0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int d = nt * 8 + t4 * 2 + q;
            const double c = fabs(A[nt][q]) + fabs(B[nt][q]);
            if (d < H) chi = max(chi, __double2hiint(c));
            if (WANT_GRID) {
              if (valid && d < H) {
                double* grow = out.grid + ((size_t)pair * F * F * F + (size_t)base);
                grow[d] = 2.0 * fabs(A[nt][q] + B[nt][q]);
                if (d != 0 && 2 * d != F) grow[F - d] = 2.0 * fabs(A[nt][q] - B[nt][q]);
              }
            }
          }
        if (valid && chi >= __double2hiint(fmax(sb, bvh))) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int d = nt * 8 + t4 * 2 + q;
              const double g1 = d < H ? fabs(A[nt][q] + B[nt][q]) : -2.0;
              const double g2 = (d < H && d != 0 && 2 * d != F) ? fabs(A[nt][q] - B[nt][q]) : -2.0;
              better32(bvh, bi, g1, base + d);
              better32(bvh, bi, g2, base + (F - d));
            }
          if (bvh > sb) atomicMax(sbest, (unsigned long long)__double_as_longlong(bvh));
        }
      }
    }
    double bv = 2.0 * bvh;
    if (bvh < 0.0) bv = -1.0;
    // ---- block arg-max (numpy order) and parabola neighbours
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, bv, off);
      const int oi = __shfl_down_sync(0xffffffffu, bi, off);
      better32(bv, bi, ov, oi);
    }
    int* redi = reinterpret_cast<int*>(red + 32);
    if ((tid & 31) == 0) {
      red[tid >> 5] = bv;
      redi[tid >> 5] = bi;
    }
    __syncthreads();
    if (tid < 32) {
      bv = (tid < X5_WARPS) ? red[tid] : -1.0;
      bi = (tid < X5_WARPS) ? redi[tid] : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        better32(bv, bi, ov, oi);
      }
      if (tid == 0) {
        red[0] = bv;
        redi[0] = bi;
      }
    }
    __syncthreads();
    bv = red[0];
    bi = redi[0];
    const bool ok = (bi != 0x7fffffff) && isfinite(bv);
    const int bx = ok ? bi / (F * F) : 0;
    const int by = ok ? (bi / F) % F : 0;
    const int bz = ok ? bi % F : 0;
    __syncthreads();
    {
      if (warp < 6) {
        const int ax = warp >> 1, sgn = (warp & 1) ? -1 : 1;
        int px = bx, py = by, pz = bz;
        if (ax == 0) px = (bx + sgn + F) % F;
        if (ax == 1) py = (by + sgn + F) % F;
        if (ax == 2) pz = (bz + sgn + F) % F;
        const double* Y = YIN + (size_t)px * slab_d;
        double acc = 0.0;
        for (int e = lane; e < M * M; e += 32) {
          const int j = e / M, l = e - j * M;
          double sj, cj, sl_, cl;
          sincospi(2.0 * (double)((j * py) % F) / (double)F, &sj, &cj);
          sincospi(2.0 * (double)((l * pz) % F) / (double)F, &sl_, &cl);
          double vr, vi;
          if (j == 0) {
            vr = Y[l * 2];
            vi = Y[l * 2 + 1];
          } else {
            const double er = Y[(size_t)j * RY + l * 2], ei = Y[(size_t)j * RY + l * 2 + 1];
            const double orr = Y[(size_t)(n + j) * RY + l * 2], oi = Y[(size_t)(n + j) * RY + l * 2 + 1];
            vr = er * cj + oi * sj;   // Re(E cos - i O sin)
            vi = ei * cj - orr * sj;  // Im
          }
          const double term = vr * cl + vi * sl_;
          acc += (l == 0) ? term : 2.0 * term;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        if (lane == 0) red[2 + warp] = fabs(acc);
      }
    }
    __syncthreads();
    if (tid == 0) {
      out.best_idx[3 * (size_t)pair + 0] = bx;
      out.best_idx[3 * (size_t)pair + 1] = by;
      out.best_idx[3 * (size_t)pair + 2] = bz;
      out.best_val[pair] = bv;
      const int b3[3] = {bx, by, bz};
      for (int ax = 0; ax < 3; ++ax) {
        const double y1 = red[2 + 2 * ax], y3 = red[2 + 2 * ax + 1], y2 = bv;
        const double d = (y3 - y1) / (2.0 * (2.0 * y2 - y1 - y3));
        out.frac_idx[3 * (size_t)pair + ax] = (double)b3[ax] - d;
      }
      if (out.status) out.status[pair] = ok ? FO_STATUS_OK : FO_STATUS_NONFINITE;
    }
    __syncthreads();  // red / YIN are reused by the next pair
  }
}

size_t xf_smem_bytes(int n, int F, bool with_grids, int xf_ng = XF_NG) {
  const int W = 2 * n + 1, M = n + 1, Mp = M | 1;
  size_t b = (size_t)F * 16;                       // tw
  b += (size_t)((3 * W + 1) & ~1) * 8;             // damp
  b += (size_t)xf_ng * F * Mp * 16;                // V
  b += 64 * 8;                                     // red
  if (with_grids) b += ((size_t)W * W * M + (size_t)F * W * M) * 16;
  return b;
}

int check_params(fo_ctx* ctx, const fo_per_params* p) {
  if (!ctx) return FO_ERR_INVALID;
  if (!p) return fo_fail(ctx, FO_ERR_INVALID, "params is NULL");
  if (p->natoms < 1) return fo_fail(ctx, FO_ERR_INVALID, "natoms must be >= 1");
  if (p->nwave < 1 || p->nwave > 64)
    return fo_fail(ctx, FO_ERR_UNSUPPORTED, "nwave=%lld outside supported range 1..64",
                   (long long)p->nwave);
  if (p->nfspace < 2 * p->nwave + 1 || p->nfspace > 1024)
    return fo_fail(ctx, FO_ERR_INVALID, "nfspace=%lld must be in [2*nwave+1, 1024]",
                   (long long)p->nfspace);
  if (!(p->sigma > 0.0)) return fo_fail(ctx, FO_ERR_INVALID, "sigma must be > 0");
  for (int i = 0; i < 3; ++i)
    if (!(p->box[i] > 0.0)) return fo_fail(ctx, FO_ERR_INVALID, "box lengths must be > 0");
  return FO_OK;
}

// Launch K_sf for nstruct structures (device positions) into bank.
int launch_sf(fo_ctx* ctx, const fo_per_params* p, const double* d_pos, int64_t nstruct,
              double2* d_bank) {
  if (nstruct == 0) return FO_OK;
  const int n = (int)p->nwave, M = n + 1;
  const int ngroups = (int)ctx->h_goff.size() - 1;
  {  // tensor-core path: NT column tiles cover the 2M columns, MT row tiles per warp, <= 10 warps per
     // CTA; more row tiles than 10 MT (fine k-grids) are split over blockIdx.z
    const int NTq = (2 * M + 7) / 8;
    const int mtiles = (M * M * 4 + 7) / 8;
    static const int mt_for_nt[10] = {0, 5, 5, 5, 4, 3, 2, 2, 2, 2};
    if (NTq <= 9 && !ctx->force_generic && !ctx->opt("per_sf_scalar")) {  // n <= 35; beyond, the scalar kernel
      const int MTq = mt_for_nt[NTq];
      int warps = (mtiles + MTq - 1) / MTq;
      if (warps > 10) warps = 10;
      const int nz = (mtiles + MTq * warps - 1) / (MTq * warps);
      const int Mp = sf_pitch(M), Mz = sf_pitch(4 * NTq);
      // double-buffered phasor tables of TA atoms.  Every tile costs one CTA barrier (measured ~2.5 % of the
      // kernel each at n = 9), so TA is the largest tile that still lets two CTAs share an SM (113 KB each),
      // then evened out over the tiles of the largest group: 204 atoms -> 2 tiles of 104 (64-atom tiles: 4)
      const bool sf3 = M == 10 && warps == 10 && !ctx->opt("per_sf_padded");  // per_sf3_kernel: no column padding
      const size_t row_bytes = (sf3 ? sf_pitch(10) + sf_pitch(12) + sf_pitch(10) : 2 * Mp + Mz) * (size_t)16;
      int TA = (int)((size_t)113 * 1024 / (2 * row_bytes)) & ~3;
      TA = TA < 16 ? 16 : (TA > 128 ? 128 : TA);
      int gmax = 1;
      for (int q = 0; q < ngroups; ++q) gmax = std::max(gmax, (int)(ctx->h_goff[q + 1] - ctx->h_goff[q]));
      const int ntile = (gmax + TA - 1) / TA;
      TA = std::min(TA, (((gmax + ntile - 1) / ntile) + 3) & ~3);
      if (const int v = (int)ctx->opt("per_sf_tile_atoms")) {  // tuning override (multiple of 4)
        if (v >= 4 && v <= 256 && v % 4 == 0) TA = v;
      }
      const size_t smem = (size_t)2 * TA * row_bytes;
      if (smem <= ctx->prop.sharedMemPerBlockOptin) {
        const double kx = kTwoPi / p->box[0], ky = kTwoPi / p->box[1], kz = kTwoPi / p->box[2];
        dim3 grid((unsigned)nstruct, (unsigned)ngroups, (unsigned)nz);
        fo_prof_scope prof(ctx, FO_PROF_PER_SF);
#define FO_SF2_LAUNCH(MT_, NT_)                                                                              \
  do {                                                                                                       \
    FO_CUDA(ctx, cudaFuncSetAttribute(per_sf2_kernel<MT_, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                           \
    per_sf2_kernel<MT_, NT_><<<grid, warps * 32, smem, ctx->stream>>>(                                       \
        d_pos, ctx->d_goff, ctx->d_gidx, ngroups, (int)p->natoms, n, TA, kx, ky, kz, d_bank);                \
  } while (0)
        if (sf3) {  // default k-grid of 256 atoms
          if (!ctx->opt("per_sf_syncthreads")) {  // tile hand-over through mbarriers (default)
            FO_CUDA(ctx, cudaFuncSetAttribute(per_sf3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            per_sf3_kernel<true><<<(unsigned)nstruct, 320, smem, ctx->stream>>>(
                d_pos, ctx->d_goff, ctx->d_gidx, ngroups, (int)p->natoms, TA, kx, ky, kz, d_bank);
          } else {
            FO_CUDA(ctx, cudaFuncSetAttribute(per_sf3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            per_sf3_kernel<false><<<(unsigned)nstruct, 320, smem, ctx->stream>>>(
                d_pos, ctx->d_goff, ctx->d_gidx, ngroups, (int)p->natoms, TA, kx, ky, kz, d_bank);
          }
          FO_LAUNCH_CHECK(ctx);
          return FO_OK;
        }
        switch (NTq) {
          case 1: FO_SF2_LAUNCH(5, 1); break;
          case 2: FO_SF2_LAUNCH(5, 2); break;
          case 3: FO_SF2_LAUNCH(5, 3); break;
          case 4: FO_SF2_LAUNCH(4, 4); break;
          case 5: FO_SF2_LAUNCH(3, 5); break;
          case 6: FO_SF2_LAUNCH(2, 6); break;
          case 7: FO_SF2_LAUNCH(2, 7); break;
          case 8: FO_SF2_LAUNCH(2, 8); break;
          default: FO_SF2_LAUNCH(2, 9); break;
        }
#undef FO_SF2_LAUNCH
        FO_LAUNCH_CHECK(ctx);
        return FO_OK;
      }
    }
  }
  // TL = 5 gives the best FMA : (mul + load) ratio when it divides n+1 well, else TL = 2.
  const int waste5 = ((M + 4) / 5) * 5 - M, waste2 = ((M + 1) / 2) * 2 - M;
  const bool use5 = (waste5 * 2 <= M / 5 + waste2 * 2) || (waste5 == 0);
  const int TL = use5 ? 5 : 2;
  const int LC = (M + TL - 1) / TL;
  const int nitems = M * M * LC;
  int threads = ((nitems + 31) / 32) * 32;
  const int max_threads = use5 ? 256 : 512;
  if (threads > max_threads) threads = max_threads;
  const int nblk = (nitems + threads - 1) / threads;
  const int Mp = M | 1, Mz = (LC * TL) | 1;
  const size_t smem = (size_t)SF_TA * (2 * Mp + Mz) * 16;
  const double kx = kTwoPi / p->box[0], ky = kTwoPi / p->box[1], kz = kTwoPi / p->box[2];
  // gridDim.x carries the structures (up to 2^31-1)
  dim3 grid((unsigned)nstruct, (unsigned)ngroups, (unsigned)nblk);
  fo_prof_scope prof(ctx, FO_PROF_PER_SF);
  if (use5) {
    FO_CUDA(ctx, cudaFuncSetAttribute(per_sf_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    per_sf_kernel<5><<<grid, threads, smem, ctx->stream>>>(d_pos, ctx->d_goff, ctx->d_gidx, ngroups,
                                                           (int)p->natoms, n, kx, ky, kz, d_bank,
                                                           nitems);
  } else {
    FO_CUDA(ctx, cudaFuncSetAttribute(per_sf_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    per_sf_kernel<2><<<grid, threads, smem, ctx->stream>>>(d_pos, ctx->d_goff, ctx->d_gidx, ngroups,
                                                           (int)p->natoms, n, kx, ky, kz, d_bank,
                                                           nitems);
  }
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}

// per_xf6_kernel applies: n <= 11, F <= 46, everything resident in shared memory
bool xf6_applies(fo_ctx* ctx, int n, int F, X6Layout* lay, int* code_out) {
  const X6Layout lay6(n, F);
  const int KS = n / 4 + 1, NT = (lay6.H + 7) / 8;
  const int code = KS * 10 + NT;
  if (lay) *lay = lay6;
  if (code_out) *code_out = code;
  return (size_t)lay6.total * 8 <= ctx->prop.sharedMemPerBlockOptin && !ctx->force_generic &&
         2 * n + 1 <= 129 && (code == 11 || code == 12 || code == 22 || code == 23 || code == 33);
}

// transform + arg-max of npairs stage-X images (per_cross6_kernel / per_sfx_kernel wrote them)
int launch_xf6_image(fo_ctx* ctx, const X6Layout& lay6, int code, const double* ximg, int64_t npairs, int n, int F,
                     XfOut out) {
  const size_t smem6 = (size_t)lay6.total * 8;
  int blocks = ctx->prop.multiProcessorCount;
  if ((int64_t)blocks > npairs) blocks = (int)npairs;
#define FO_X6_LAUNCH(KS_, NT_)                                                                           \
  do {                                                                                                   \
    if (out.grid) {                                                                                      \
      FO_CUDA(ctx, cudaFuncSetAttribute(per_xf6_kernel<KS_, NT_, true>,                                  \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6));       \
      per_xf6_kernel<KS_, NT_, true><<<blocks, X6_THREADS, smem6, ctx->stream>>>(                        \
          lay6, ximg, (int)npairs, n, F, out);                                                           \
    } else {                                                                                             \
      FO_CUDA(ctx, cudaFuncSetAttribute(per_xf6_kernel<KS_, NT_, false>,                                 \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6));       \
      per_xf6_kernel<KS_, NT_, false><<<blocks, X6_THREADS, smem6, ctx->stream>>>(                       \
          lay6, ximg, (int)npairs, n, F, out);                                                           \
    }                                                                                                    \
  } while (0)
  if (code == 11) FO_X6_LAUNCH(1, 1);
  else if (code == 12) FO_X6_LAUNCH(1, 2);
  else if (code == 22) FO_X6_LAUNCH(2, 2);
  else if (code == 23) FO_X6_LAUNCH(2, 3);
  else FO_X6_LAUNCH(3, 3);
#undef FO_X6_LAUNCH
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}

int launch_xf(fo_ctx* ctx, const fo_per_params* p, const double2* d_bankA, const double2* d_bankB,
              const long long* d_pairs, int64_t npairs, XfOut out, const int32_t* d_ops = nullptr) {
  if (npairs == 0) return FO_OK;
  const int n = (int)p->nwave, F = (int)p->nfspace;
  const int ngroups = (int)ctx->h_goff.size() - 1;
  const size_t optin = ctx->prop.sharedMemPerBlockOptin;
  int blocks = ctx->prop.multiProcessorCount;
  if ((int64_t)blocks > npairs) blocks = (int)npairs;
  const double kx = kTwoPi / p->box[0], ky = kTwoPi / p->box[1], kz = kTwoPi / p->box[2];
  {  // tensor-core path, stages Y -> Z chained in registers (per_xf6_kernel)
    X6Layout lay6;
    int code = 0;
    if (xf6_applies(ctx, n, F, &lay6, &code)) {
      void* ximg = nullptr;
      FO_CHECK(fo_scratch(ctx, FO_SCR_IPK, (size_t)npairs * lay6.ximg_doubles() * 8, &ximg));
      fo_prof_scope prof(ctx, FO_PROF_PER_XF);
      per_cross6_kernel<<<(unsigned)npairs, 256, 0, ctx->stream>>>(lay6, d_bankA, d_bankB, d_pairs, d_ops, ngroups, n,
                                                                 kx, ky, kz, p->sigma, (double*)ximg);
      FO_LAUNCH_CHECK(ctx);
      return launch_xf6_image(ctx, lay6, code, (const double*)ximg, npairs, n, F, out);
    }
  }
  if (d_ops)
    return fo_fail(ctx, FO_ERR_UNSUPPORTED, "cell-symmetry operations need the resident transform (nwave <= 11)");
  {  // tensor-core path for fine k-grids: stage-X image and YIN in global memory (L2), slabs staged
    const int NT5 = (F / 2 + 1 + 7) / 8;
    const X5Layout lay5(n, F, NT5, optin);
    const size_t smem5 = (size_t)lay5.total * 8;
    if (NT5 >= 3 && NT5 <= 9 && smem5 <= optin && !ctx->force_generic && 2 * n + 1 <= 129 && !ctx->opt("per_xf_generic")) {
      void *ximg = nullptr, *yin = nullptr;
      FO_CHECK(fo_scratch(ctx, FO_SCR_IPK, (size_t)npairs * lay5.ximg_doubles() * 8, &ximg));
      FO_CHECK(fo_scratch(ctx, FO_SCR_WORK, (size_t)blocks * lay5.yin_doubles(F) * 8, &yin));
      // per_cross_kernel only needs M, W, RXp of the layout: build the X4 form of it
      X4Layout lx;
      lx.M = lay5.M;
      lx.W = lay5.W;
      lx.RXp = lay5.RXp;
      fo_prof_scope prof(ctx, FO_PROF_PER_XF);
      per_cross_kernel<<<(unsigned)npairs, 256, 0, ctx->stream>>>(lx, d_bankA, d_bankB, d_pairs, ngroups, n, kx,
                                                                ky, kz, p->sigma, (double*)ximg);
      FO_LAUNCH_CHECK(ctx);
#define FO_X5_LAUNCH(NT_)                                                                                  \
  do {                                                                                                     \
    if (out.grid) {                                                                                        \
      FO_CUDA(ctx, cudaFuncSetAttribute(per_xf5_kernel<NT_, true>,                                         \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));         \
      per_xf5_kernel<NT_, true><<<blocks, X5_THREADS, smem5, ctx->stream>>>(                               \
          lay5, (const double*)ximg, (double*)yin, (int)npairs, n, F, out);                                \
    } else {                                                                                               \
      FO_CUDA(ctx, cudaFuncSetAttribute(per_xf5_kernel<NT_, false>,                                        \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));         \
      per_xf5_kernel<NT_, false><<<blocks, X5_THREADS, smem5, ctx->stream>>>(                              \
          lay5, (const double*)ximg, (double*)yin, (int)npairs, n, F, out);                                \
    }                                                                                                      \
  } while (0)
      switch (NT5) {
        case 3: FO_X5_LAUNCH(3); break;
        case 4: FO_X5_LAUNCH(4); break;
        case 5: FO_X5_LAUNCH(5); break;
        case 6: FO_X5_LAUNCH(6); break;
        case 7: FO_X5_LAUNCH(7); break;
        case 8: FO_X5_LAUNCH(8); break;
        default: FO_X5_LAUNCH(9); break;
      }
#undef FO_X5_LAUNCH
      FO_LAUNCH_CHECK(ctx);
      return FO_OK;
    }
  }
  size_t smem = xf_smem_bytes(n, F, true);
  double2* gscratch = nullptr;
  int xf_ng = XF_NG;
  if (smem > optin) {
    smem = xf_smem_bytes(n, F, false);
    while (smem > optin && xf_ng > 1) {  // fine k-grids: fewer, larger slab groups
      xf_ng >>= 1;
      smem = xf_smem_bytes(n, F, false, xf_ng);
    }
    if (smem > optin)
      return fo_fail(ctx, FO_ERR_UNSUPPORTED,
                     "nwave=%d nfspace=%d needs %zu bytes of shared memory (> %zu)", n, F, smem,
                     optin);
    const int W = 2 * n + 1, M = n + 1;
    const size_t per_cta = ((size_t)W * W * M + (size_t)F * W * M) * 16;
    void* ptr = nullptr;
    FO_CHECK(fo_scratch(ctx, FO_SCR_WORK, per_cta * blocks, &ptr));
    gscratch = (double2*)ptr;
  }
  FO_CUDA(ctx, cudaFuncSetAttribute(per_xf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  fo_prof_scope prof(ctx, FO_PROF_PER_XF);
  per_xf_kernel<<<blocks, XF_THREADS, smem, ctx->stream>>>(d_bankA, d_bankB, d_pairs, (int)npairs,
                                                           ngroups, n, F, kx, ky, kz, p->sigma,
                                                           gscratch, xf_ng, out);
  FO_LAUNCH_CHECK(ctx);
  return FO_OK;
}

// Independent pairs at the default k-grid of a 256-atom cell (n = 9): per_sfx_kernel (structure factors of both
// structures + cross-spectrum, no bank) -> per_xf6_kernel.
bool pairs_fused_applies(fo_ctx* ctx, const fo_per_params* p) {
  return p->nwave == 9 && ctx->pairs_fused && xf6_applies(ctx, (int)p->nwave, (int)p->nfspace, nullptr, nullptr);
}

int launch_pairs_fused(fo_ctx* ctx, const fo_per_params* p, const double* d_posA, const double* d_posB,
                       int64_t npairs, XfOut out) {
  if (npairs == 0) return FO_OK;
  const int n = (int)p->nwave, F = (int)p->nfspace;
  const int ngroups = (int)ctx->h_goff.size() - 1;
  X6Layout lay6;
  int code = 0;
  xf6_applies(ctx, n, F, &lay6, &code);
  // atom tile as in launch_sf (per_sf3_kernel): two CTAs per SM, evened out over the tiles of the largest group
  const size_t row_bytes = (size_t)(sf_pitch(10) + sf_pitch(12) + sf_pitch(10)) * 16;
  int TA = (int)((size_t)113 * 1024 / (2 * row_bytes)) & ~3;
  int gmax = 1;
  for (int q = 0; q < ngroups; ++q) gmax = std::max(gmax, (int)(ctx->h_goff[q + 1] - ctx->h_goff[q]));
  const int ntile = (gmax + TA - 1) / TA;
  TA = std::max(16, std::min(TA, (((gmax + ntile - 1) / ntile) + 3) & ~3));
  const size_t smem = (size_t)2 * TA * row_bytes;
  void* ximg = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_IPK, (size_t)npairs * lay6.ximg_doubles() * 8, &ximg));
  const double kx = kTwoPi / p->box[0], ky = kTwoPi / p->box[1], kz = kTwoPi / p->box[2];
  {
    fo_prof_scope prof(ctx, FO_PROF_PER_SF);
    FO_CUDA(ctx, cudaFuncSetAttribute(per_sfx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    per_sfx_kernel<<<(unsigned)npairs, 320, smem, ctx->stream>>>(
        d_posA, d_posB, ctx->d_goff, ctx->d_gidx, ngroups, (int)p->natoms, TA, kx, ky, kz, p->sigma, lay6.RXp,
        (double*)ximg, (size_t)lay6.ximg_doubles());
    FO_LAUNCH_CHECK(ctx);
  }
  fo_prof_scope prof(ctx, FO_PROF_PER_XF);
  return launch_xf6_image(ctx, lay6, code, (const double*)ximg, npairs, n, F, out);
}

size_t bank_elems_per_struct(fo_ctx* ctx, const fo_per_params* p) {
  const size_t W = 2 * p->nwave + 1, M = p->nwave + 1;
  return (ctx->h_goff.size() - 1) * W * W * M;
}

// pairs per chunk: bounded by a bank budget, a multiple of the SM count where possible
int64_t chunk_pairs(fo_ctx* ctx, const fo_per_params* p, int64_t npairs, bool want_grid,
                    size_t budget_mb = 768) {
  const size_t per_pair = 2 * bank_elems_per_struct(ctx, p) * 16;
  size_t budget = budget_mb << 20;
  if (const int64_t mb = ctx->opt("per_chunk_mb")) budget = (size_t)mb << 20;  // tuning hook of the A/B scripts
  int64_t c = (int64_t)(budget / per_pair);
  if (want_grid) {
    const size_t g = (size_t)p->nfspace * p->nfspace * p->nfspace * 8;
    int64_t cg = (int64_t)(((size_t)512 << 20) / g);
    if (cg < c) c = cg;
  }
  const int sms = ctx->prop.multiProcessorCount;
  if (c > sms) c = (c / sms) * sms;
  if (c < 1) c = 1;
  if (c > npairs) c = npairs;
  return c;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------

extern "C" int fo_per_align_pairs_dev(fo_ctx* ctx, const fo_per_params* p, const double* d_posA,
                                      const double* d_posB, int64_t npairs, int64_t* d_best_idx,
                                      double* d_best_val, double* d_frac_idx, double* d_grid_out,
                                      int32_t* d_status) {
  FO_CHECK(check_params(ctx, p));
  if (npairs < 0 || (npairs > 0 && (!d_posA || !d_posB || !d_best_idx || !d_best_val || !d_frac_idx)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_align_pairs_dev: NULL argument");
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, p->natoms));
  const size_t per_struct = bank_elems_per_struct(ctx, p);
  // device-resident input: no copies to overlap, so the chunks only bound the scratch -- three times the bank
  // budget of the host-buffer pipeline (fewer kernel boundaries: BLJ256 9.80 -> 9.63 ms per 16384 pairs)
  const bool fused = pairs_fused_applies(ctx, p);  // no bank: structure factors -> cross-spectrum in one kernel
  // (the fused path keeps only the 62.7 KB image per pair: four times the pairs per chunk, 31.9 -> 31.6 ms per 65536)
  const int64_t chunk = chunk_pairs(ctx, p, npairs, false, fused ? 9216 : 2304);
  void* bank = nullptr;
  if (npairs > 0 && !fused) FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, 2 * (size_t)chunk * per_struct * 16, &bank));
  double2* bankA = (double2*)bank;
  double2* bankB = bankA + (size_t)chunk * per_struct;
  const size_t F3 = (size_t)p->nfspace * p->nfspace * p->nfspace;
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = (npairs - p0 < chunk) ? npairs - p0 : chunk;
    XfOut out;
    out.best_idx = (long long*)d_best_idx + 3 * p0;
    out.best_val = d_best_val + p0;
    out.frac_idx = d_frac_idx + 3 * p0;
    out.grid = d_grid_out ? d_grid_out + (size_t)p0 * F3 : nullptr;
    out.status = d_status ? d_status + p0 : nullptr;
    if (fused) {
      FO_CHECK(launch_pairs_fused(ctx, p, d_posA + (size_t)p0 * p->natoms * 3, d_posB + (size_t)p0 * p->natoms * 3,
                                  np, out));
      continue;
    }
    FO_CHECK(launch_sf(ctx, p, d_posA + (size_t)p0 * p->natoms * 3, np, bankA));
    FO_CHECK(launch_sf(ctx, p, d_posB + (size_t)p0 * p->natoms * 3, np, bankB));
    FO_CHECK(launch_xf(ctx, p, bankA, bankB, nullptr, np, out));
  }
  return FO_OK;
}

extern "C" int fo_per_align_pairs_full_dev(fo_ctx* ctx, const fo_per_params* p, const double* d_posA,
                                           const double* d_posB, int64_t npairs, int niter, double* d_dist,
                                           int32_t* d_perm, double* d_disp, int32_t* d_flag, int64_t* d_best_idx,
                                           double* d_best_val, double* d_frac_idx, int32_t* d_status) {
  if (ctx && npairs > 0 && (!d_dist || !d_disp || !d_flag))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_align_pairs_full_dev: NULL argument");
  FO_CHECK(fo_per_align_pairs_dev(ctx, p, d_posA, d_posB, npairs, d_best_idx, d_best_val, d_frac_idx, nullptr,
                                  d_status));
  return fo_per_assign_run_dev(ctx, p, d_posA, d_posB, d_frac_idx, npairs, niter, d_dist, d_disp, d_perm, d_flag);
}

namespace {
// Shared host-buffer driver: `stage(p0, np)` must enqueue whatever fills bankA/bankB for pairs
// [p0, p0+np) and return the (bankA, bankB, d_pairs) to use.
struct HostOut {
  int64_t* best_idx;
  double* best_val;
  double* frac_idx;
  double* grid_out;
  int32_t* status;
};

int copy_out(fo_ctx* ctx, const fo_per_params* p, int64_t p0, int64_t np, const char* d_out,
             const double* d_grid, const HostOut& h) {
  // d_out layout per chunk: best_idx[np*3] i64 | best_val[np] | frac[np*3] | status[np] i32
  const size_t o_idx = 0, o_val = (size_t)np * 24, o_frac = o_val + (size_t)np * 8,
               o_st = o_frac + (size_t)np * 24;
  FO_CUDA(ctx, cudaMemcpyAsync(h.best_idx + 3 * p0, d_out + o_idx, (size_t)np * 24,
                               cudaMemcpyDeviceToHost, ctx->stream));
  FO_CUDA(ctx, cudaMemcpyAsync(h.best_val + p0, d_out + o_val, (size_t)np * 8,
                               cudaMemcpyDeviceToHost, ctx->stream));
  FO_CUDA(ctx, cudaMemcpyAsync(h.frac_idx + 3 * p0, d_out + o_frac, (size_t)np * 24,
                               cudaMemcpyDeviceToHost, ctx->stream));
  if (h.status)
    FO_CUDA(ctx, cudaMemcpyAsync(h.status + p0, d_out + o_st, (size_t)np * 4,
                                 cudaMemcpyDeviceToHost, ctx->stream));
  if (h.grid_out) {
    const size_t F3 = (size_t)p->nfspace * p->nfspace * p->nfspace;
    FO_CUDA(ctx, cudaMemcpyAsync(h.grid_out + (size_t)p0 * F3, d_grid, (size_t)np * F3 * 8,
                                 cudaMemcpyDeviceToHost, ctx->stream));
  }
  return FO_OK;
}

// Deferred delivery of the per-pair results: one D2H of the whole chunk block into a pinned ring
// (never blocks the host, unlike a D2H into the caller's pageable arrays), unpacked into the caller's
// arrays one chunk later while the GPU already works on the next chunk.
void deliver_out(const HostOut& h, int64_t p0, int64_t np, const char* src) {
  memcpy(h.best_idx + 3 * p0, src, (size_t)np * 24);
  memcpy(h.best_val + p0, src + (size_t)np * 24, (size_t)np * 8);
  memcpy(h.frac_idx + 3 * p0, src + (size_t)np * 32, (size_t)np * 24);
  if (h.status) memcpy(h.status + p0, src + (size_t)np * 56, (size_t)np * 4);
}

XfOut make_out(char* d_out, int64_t np, double* d_grid, bool want_status) {
  XfOut o;
  o.best_idx = (long long*)d_out;
  o.best_val = (double*)(d_out + (size_t)np * 24);
  o.frac_idx = (double*)(d_out + (size_t)np * 32);
  o.status = want_status ? (int*)(d_out + (size_t)np * 56) : nullptr;
  o.grid = d_grid;
  return o;
}
}  // namespace

namespace {
// Outputs of the full alignment (fo_per_align_pairs_full): the device screening settles a pair (dist, disp,
// perm written by per_assign_kernel) or flags it for the host LAP pool.
struct FullOut {
  int niter, nthreads;
  double* dist;      // [P]
  int32_t* perm;     // [P,N] or null
  double* disp;      // [P,3] or null
  int64_t nhost = 0; // pairs that went through the host pool
};

int per_align_pairs_impl(fo_ctx* ctx, const fo_per_params* p, const double* posA, const double* posB,
                         int64_t npairs, int64_t* best_idx, double* best_val, double* frac_idx, double* grid_out,
                         int32_t* status, FullOut* full) {
  FO_CHECK(check_params(ctx, p));
  if (npairs < 0 || (npairs > 0 && (!posA || !posB || !best_idx || !best_val || !frac_idx)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_align_pairs: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, p->natoms));
  const size_t per_struct = bank_elems_per_struct(ctx, p);
  const bool fused = pairs_fused_applies(ctx, p);  // no bank: structure factors -> cross-spectrum in one kernel
  // chunk of the host pipeline: 3256 BLJ256 pairs on the bank path; the fused path takes four times as many
  // (measured 1628 / 3256 / 6512 / 13024 / 26048 pairs per chunk: 1.52 / 1.60 / 1.66 / 1.68 / 1.66 M aligned pairs/s)
  const int64_t chunk = chunk_pairs(ctx, p, npairs, grid_out != nullptr, fused ? 3072 : 768);
  const int64_t N = p->natoms;
  const size_t pos_bytes = (size_t)chunk * N * 3 * 8;
  const size_t F3 = (size_t)p->nfspace * p->nfspace * p->nfspace;
  void *bank, *dA, *dB, *dOut, *dGrid = nullptr, *hA, *hB, *dFull = nullptr;
  bank = nullptr;
  if (!fused) FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, 2 * (size_t)chunk * per_struct * 16, &bank));
  // two position buffers per side so the copy of chunk c+1 overlaps the kernels of chunk c
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSA, 2 * pos_bytes, &dA));
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSB, 2 * pos_bytes, &dB));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * 64, &dOut));
  if (grid_out) FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * F3 * 8, &dGrid));
  // full alignment: dist[np] | disp[3 np] | flag[np] i32 (padded to 8 np) | perm[np N] i32, per chunk
  const bool want_perm = full && full->perm;
  // permutations travel device -> host in the narrowest index type that holds natoms (BLJ256: one byte per atom
  // instead of four: the D2H of the permutations was 94 % of the result bytes) and are widened on delivery
  const int pelt = N <= 256 ? 1 : (N <= 65536 ? 2 : 4);
  const size_t full_stride = 40 + (want_perm ? (size_t)N * pelt : 0);
  if (full) FO_CHECK(fo_scratch(ctx, FO_SCR_FULL, (size_t)chunk * full_stride, &dFull));
  const bool pinnedA = fo_is_pinned(posA), pinnedB = fo_is_pinned(posB);
  hA = hB = nullptr;
  if (!pinnedA) FO_CHECK(fo_pinned(ctx, 0, 2 * pos_bytes, &hA));
  if (!pinnedB) FO_CHECK(fo_pinned(ctx, 1, 2 * pos_bytes, &hB));
  double2* bankA = (double2*)bank;
  double2* bankB = bankA + (size_t)chunk * per_struct;
  HostOut h = {best_idx, best_val, frac_idx, grid_out, status};
  // chunk boundaries: a short first chunk, so that the only H2D copy nothing can hide is small
  const std::vector<int64_t> starts = fo_chunk_starts(npairs, chunk);
  const int64_t nchunks = (int64_t)starts.size() - 1;
  // pipeline: [host memcpy -> pinned] -> H2D on copy_stream -> kernels + D2H on stream
  auto stage_in = [&](int64_t c) -> int {
    const int64_t p0 = starts[c];
    const int64_t np = starts[c + 1] - p0;
    const int buf = (int)(c & 1);
    const size_t nb = (size_t)np * N * 3 * 8;
    // the pinned buffer `buf` was last read by the H2D of chunk c-2
    if (c >= 2) FO_CUDA(ctx, cudaEventSynchronize(ctx->ev[buf]));
    // caller buffers that are already page-locked are DMA'd directly; pageable ones are staged
    const char* srcA = (const char*)(posA + (size_t)p0 * N * 3);
    const char* srcB = (const char*)(posB + (size_t)p0 * N * 3);
    if (!pinnedA) {
      fo_host_copy((char*)hA + buf * pos_bytes, srcA, nb, full ? full->nthreads : 0);
      srcA = (const char*)hA + buf * pos_bytes;
    }
    if (!pinnedB) {
      fo_host_copy((char*)hB + buf * pos_bytes, srcB, nb, full ? full->nthreads : 0);
      srcB = (const char*)hB + buf * pos_bytes;
    }
    // device buffer `buf` was last read by the kernels of chunk c-2
    if (c >= 2) FO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[2 + buf], 0));
    FO_CUDA(ctx, cudaMemcpyAsync((char*)dA + buf * pos_bytes, srcA, nb, cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
    FO_CUDA(ctx, cudaMemcpyAsync((char*)dB + buf * pos_bytes, srcB, nb, cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
    FO_CUDA(ctx, cudaEventRecord(ctx->ev[buf], ctx->copy_stream));
    return FO_OK;
  };
  void *hOut = nullptr, *hFull = nullptr;
  const size_t out_bytes = (size_t)chunk * 64, full_bytes = (size_t)chunk * full_stride;
  if (!grid_out) FO_CHECK(fo_pinned(ctx, 2, 2 * out_bytes, &hOut));
  if (full) FO_CHECK(fo_pinned(ctx, 3, 2 * full_bytes, &hFull));
  std::vector<int64_t> hard;
  // unpack chunk c from the pinned ring into the caller's arrays; full alignment: the flagged pairs go through
  // the host pool here, while the GPU works on the next chunk
  auto deliver = [&](int64_t c) -> int {
    const int64_t p0 = starts[c], np = starts[c + 1] - p0;
    deliver_out(h, p0, np, (const char*)hOut + (c & 1) * out_bytes);
    if (!full) return FO_OK;
    const char* src = (const char*)hFull + (c & 1) * full_bytes;
    const int32_t* flag = (const int32_t*)(src + (size_t)np * 32);
    memcpy(full->dist + p0, src, (size_t)np * 8);
    if (full->disp) memcpy(full->disp + 3 * p0, src + (size_t)np * 8, (size_t)np * 24);
    if (want_perm) {
      int32_t* dstp = full->perm + (size_t)p0 * N;
      const size_t ne = (size_t)np * N;
      if (pelt == 4) {
        fo_host_copy(dstp, src + (size_t)np * 40, ne * 4, full->nthreads);
      } else {
        int nt = full->nthreads > 0 ? full->nthreads / 2 : omp_get_max_threads() / 2;
        nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
        const unsigned char* s1 = (const unsigned char*)(src + (size_t)np * 40);
        const unsigned short* s2 = (const unsigned short*)s1;
#pragma omp parallel for schedule(static) num_threads(nt) if (nt > 1)
        for (long long b = 0; b < (long long)((ne + 65535) >> 16); ++b) {
          const size_t e0 = (size_t)b << 16, e1 = std::min(ne, e0 + 65536);
          if (pelt == 1)
            for (size_t e = e0; e < e1; ++e) dstp[e] = s1[e];
          else
            for (size_t e = e0; e < e1; ++e) dstp[e] = s2[e];
        }
      }
    }
    hard.clear();
    for (int64_t q = 0; q < np; ++q)
      if (flag[q]) hard.push_back(p0 + q);
    if (!hard.empty()) {
      full->nhost += (int64_t)hard.size();
      const int rc = fo_host_refine_periodic_subset(p, ctx->h_goff.data(), (int64_t)ctx->h_goff.size() - 1,
                                                    ctx->h_gidx.data(), posA, posB, frac_idx, hard.data(),
                                                    (int64_t)hard.size(), full->niter, full->nthreads, full->dist,
                                                    full->perm, full->disp);
      if (rc != FO_OK) return fo_fail(ctx, rc, "host refinement of %zu flagged pairs failed", hard.size());
    }
    return FO_OK;
  };
  FO_CHECK(stage_in(0));
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t p0 = starts[c];
    const int64_t np = starts[c + 1] - p0;
    const int buf = (int)(c & 1);
    const double* cA = (const double*)((char*)dA + buf * pos_bytes);
    const double* cB = (const double*)((char*)dB + buf * pos_bytes);
    FO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev[buf], 0));
    XfOut out = make_out((char*)dOut, np, (double*)dGrid, status != nullptr);
    if (fused) {
      FO_CHECK(launch_pairs_fused(ctx, p, cA, cB, np, out));
      if (!full) FO_CUDA(ctx, cudaEventRecord(ctx->ev[2 + buf], ctx->stream));
    } else {
      FO_CHECK(launch_sf(ctx, p, cA, np, bankA));
      FO_CHECK(launch_sf(ctx, p, cB, np, bankB));
      if (!full) FO_CUDA(ctx, cudaEventRecord(ctx->ev[2 + buf], ctx->stream));
      FO_CHECK(launch_xf(ctx, p, bankA, bankB, nullptr, np, out));
    }
    if (full) {  // screening + permutation <-> displacement loop on the device (reads the positions again)
      char* f = (char*)dFull;
      FO_CHECK(fo_per_assign_run_dev(ctx, p, cA, cB, out.frac_idx, np, full->niter, (double*)f,
                                     (double*)(f + (size_t)np * 8), want_perm ? (void*)(f + (size_t)np * 40) : nullptr,
                                     (int32_t*)(f + (size_t)np * 32), pelt));
      FO_CUDA(ctx, cudaEventRecord(ctx->ev[2 + buf], ctx->stream));
    }
    if (grid_out) {  // test / single-pair path: large grids straight into the caller's array
      FO_CHECK(copy_out(ctx, p, p0, np, (const char*)dOut, (const double*)dGrid, h));
      if (c + 1 < nchunks) FO_CHECK(stage_in(c + 1));
      continue;
    }
    FO_CUDA(ctx, cudaMemcpyAsync((char*)hOut + buf * out_bytes, dOut, (size_t)np * 60, cudaMemcpyDeviceToHost,
                                 ctx->stream));
    if (full)
      FO_CUDA(ctx, cudaMemcpyAsync((char*)hFull + buf * full_bytes, dFull, (size_t)np * full_stride,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaEventRecord(ctx->ev[4 + buf], ctx->stream));
    // host work of this iteration runs while the GPU is busy with chunk c
    if (c + 1 < nchunks) FO_CHECK(stage_in(c + 1));
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

extern "C" int fo_per_align_pairs(fo_ctx* ctx, const fo_per_params* p, const double* posA,
                                  const double* posB, int64_t npairs, int64_t* best_idx,
                                  double* best_val, double* frac_idx, double* grid_out,
                                  int32_t* status) {
  return per_align_pairs_impl(ctx, p, posA, posB, npairs, best_idx, best_val, frac_idx, grid_out, status, nullptr);
}

extern "C" int fo_per_align_pairs_full(fo_ctx* ctx, const fo_per_params* p, const double* posA, const double* posB,
                                       int64_t npairs, int niter, int nthreads, double* dist, int32_t* perm,
                                       double* disp, double* frac_idx, int32_t* status, int64_t* nhost) {
  if (!ctx) return FO_ERR_INVALID;
  if (npairs > 0 && !dist) return fo_fail(ctx, FO_ERR_INVALID, "fo_per_align_pairs_full: dist is NULL");
  if (npairs <= 0) {
    if (nhost) *nhost = 0;
    return npairs < 0 ? fo_fail(ctx, FO_ERR_INVALID, "npairs < 0") : FO_OK;
  }
  // best_idx / best_val (and frac_idx when the caller does not want it) live in library-owned host memory
  std::vector<int64_t> bi((size_t)npairs * 3);
  std::vector<double> bv((size_t)npairs), fr;
  if (!frac_idx) {
    fr.resize((size_t)npairs * 3);
    frac_idx = fr.data();
  }
  FullOut full;
  full.niter = niter;
  full.nthreads = nthreads;
  full.dist = dist;
  full.perm = perm;
  full.disp = disp;
  const int rc = per_align_pairs_impl(ctx, p, posA, posB, npairs, bi.data(), bv.data(), frac_idx, nullptr, status,
                                      &full);
  if (nhost) *nhost = full.nhost;
  return rc;
}

// a8 fused: positions -> F^3 |f| grid on the device -> top-npeaks displacements; the grid never
// leaves HBM (findDisps with npeaks > 1, periodicAlignment.py:442-451).
extern "C" int fo_per_align_pairs_peaks(fo_ctx* ctx, const fo_per_params* p, const double* posA,
                                        const double* posB, int64_t npairs, int64_t npeaks, int64_t width,
                                        double* peaks, double* amplitude, double* mean, double* alpha,
                                        int32_t* nfound) {
  FO_CHECK(check_params(ctx, p));
  if (npairs < 0 || (npairs > 0 && (!posA || !posB || !peaks || !amplitude || !nfound)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_align_pairs_peaks: NULL argument");
  if (npeaks < 1 || npeaks > 64 || width < 1 || width > 4)
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_align_pairs_peaks: npeaks in 1..64, width in 1..4");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, p->natoms));
  const size_t per_struct = bank_elems_per_struct(ctx, p);
  const int64_t chunk = chunk_pairs(ctx, p, npairs, true);
  const size_t pos_bytes = (size_t)chunk * p->natoms * 3 * 8;
  const size_t F3 = (size_t)p->nfspace * p->nfspace * p->nfspace;
  const int64_t shape[3] = {p->nfspace, p->nfspace, p->nfspace};
  void *bank, *dA, *dB, *dOut, *dGrid;
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, 2 * (size_t)chunk * per_struct * 16, &bank));
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSA, pos_bytes, &dA));
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSB, pos_bytes, &dB));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * 64, &dOut));
  FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * F3 * 8, &dGrid));
  double *pk, *amp, *mn, *al;
  int32_t* nf;
  FO_CHECK(fo_peaks_outputs(ctx, chunk, npeaks, &pk, &amp, &mn, &al, &nf));
  double2* bankA = (double2*)bank;
  double2* bankB = bankA + (size_t)chunk * per_struct;
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = (npairs - p0 < chunk) ? npairs - p0 : chunk;
    const size_t nb = (size_t)np * p->natoms * 3 * 8;
    FO_CUDA(ctx, cudaMemcpyAsync(dA, posA + (size_t)p0 * p->natoms * 3, nb, cudaMemcpyHostToDevice, ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(dB, posB + (size_t)p0 * p->natoms * 3, nb, cudaMemcpyHostToDevice, ctx->stream));
    FO_CHECK(launch_sf(ctx, p, (const double*)dA, np, bankA));
    FO_CHECK(launch_sf(ctx, p, (const double*)dB, np, bankB));
    XfOut out = make_out((char*)dOut, np, (double*)dGrid, false);
    FO_CHECK(launch_xf(ctx, p, bankA, bankB, nullptr, np, out));
    FO_CHECK(fo_peaks_run_dev(ctx, (double*)dGrid, np, shape, npeaks, width, pk, amp, mn, al, nf));
    FO_CHECK(fo_peaks_copy_out(ctx, p0, np, npeaks, pk, amp, mn, al, nf, peaks, amplitude, mean, alpha, nfound));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

extern "C" int fo_per_structure_factors(fo_ctx* ctx, const fo_per_params* p, const double* pos,
                                        int64_t nstruct, double* out) {
  FO_CHECK(check_params(ctx, p));
  if (nstruct < 0 || (nstruct > 0 && (!pos || !out)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_structure_factors: NULL argument");
  if (nstruct == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, p->natoms));
  const size_t per_struct = bank_elems_per_struct(ctx, p);
  const size_t ng = ctx->h_goff.size() - 1;
  const size_t W = 2 * p->nwave + 1;
  const size_t full_per_struct = ng * W * W * W;
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full_per_struct * 16));
  if (chunk < 1) chunk = 1;
  if (chunk > nstruct) chunk = nstruct;
  void *bank, *dpos, *dfull;
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, (size_t)chunk * per_struct * 16, &bank));
  FO_CHECK(fo_scratch(ctx, FO_SCR_POSA, (size_t)chunk * p->natoms * 24, &dpos));
  FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * full_per_struct * 16, &dfull));
  for (int64_t s0 = 0; s0 < nstruct; s0 += chunk) {
    const int64_t ns = (nstruct - s0 < chunk) ? nstruct - s0 : chunk;
    FO_CUDA(ctx, cudaMemcpyAsync(dpos, pos + (size_t)s0 * p->natoms * 3, (size_t)ns * p->natoms * 24,
                                 cudaMemcpyHostToDevice, ctx->stream));
    FO_CHECK(launch_sf(ctx, p, (const double*)dpos, ns, (double2*)bank));
    const size_t total = (size_t)ns * full_per_struct;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 65535) blocks = 65535;
    per_expand_kernel<<<blocks, 256, 0, ctx->stream>>>((const double2*)bank, (double2*)dfull,
                                                       (int)p->nwave, (size_t)ns * ng);
    FO_LAUNCH_CHECK(ctx);
    FO_CUDA(ctx, cudaMemcpyAsync(out + (size_t)s0 * full_per_struct * 2, dfull, total * 16,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

extern "C" int fo_per_align_coeffs(fo_ctx* ctx, const fo_per_params* p, const double* CA,
                                   const double* CB, int64_t npairs, int64_t* best_idx,
                                   double* best_val, double* frac_idx, double* grid_out,
                                   int32_t* status) {
  FO_CHECK(check_params(ctx, p));
  if (npairs < 0 || (npairs > 0 && (!CA || !CB || !best_idx || !best_val || !frac_idx)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_align_coeffs: NULL argument");
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, p->natoms));
  const size_t per_struct = bank_elems_per_struct(ctx, p);
  const size_t ng = ctx->h_goff.size() - 1;
  const size_t W = 2 * p->nwave + 1;
  const size_t full_per_struct = ng * W * W * W;
  const size_t F3 = (size_t)p->nfspace * p->nfspace * p->nfspace;
  int64_t chunk = (int64_t)(((size_t)256 << 20) / (full_per_struct * 32));
  if (grid_out) {
    int64_t cg = (int64_t)(((size_t)512 << 20) / (F3 * 8));
    if (cg < chunk) chunk = cg;
  }
  if (chunk < 1) chunk = 1;
  if (chunk > npairs) chunk = npairs;
  void *bank, *dfull, *dOut, *dGrid = nullptr;
  FO_CHECK(fo_scratch(ctx, FO_SCR_BANK, 2 * (size_t)chunk * per_struct * 16, &bank));
  FO_CHECK(fo_scratch(ctx, FO_SCR_COEF, 2 * (size_t)chunk * full_per_struct * 16, &dfull));
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * 64, &dOut));
  if (grid_out) FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * F3 * 8, &dGrid));
  double2* bankA = (double2*)bank;
  double2* bankB = bankA + (size_t)chunk * per_struct;
  double2* fullA = (double2*)dfull;
  double2* fullB = fullA + (size_t)chunk * full_per_struct;
  HostOut h = {best_idx, best_val, frac_idx, grid_out, status};
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = (npairs - p0 < chunk) ? npairs - p0 : chunk;
    FO_CUDA(ctx, cudaMemcpyAsync(fullA, CA + (size_t)p0 * full_per_struct * 2,
                                 (size_t)np * full_per_struct * 16, cudaMemcpyHostToDevice,
                                 ctx->stream));
    FO_CUDA(ctx, cudaMemcpyAsync(fullB, CB + (size_t)p0 * full_per_struct * 2,
                                 (size_t)np * full_per_struct * 16, cudaMemcpyHostToDevice,
                                 ctx->stream));
    const size_t total = (size_t)np * per_struct;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 65535) blocks = 65535;
    per_compress_kernel<<<blocks, 256, 0, ctx->stream>>>(fullA, bankA, (int)p->nwave, (size_t)np * ng);
    FO_LAUNCH_CHECK(ctx);
    per_compress_kernel<<<blocks, 256, 0, ctx->stream>>>(fullB, bankB, (int)p->nwave, (size_t)np * ng);
    FO_LAUNCH_CHECK(ctx);
    XfOut out = make_out((char*)dOut, np, (double*)dGrid, status != nullptr);
    FO_CHECK(launch_xf(ctx, p, bankA, bankB, nullptr, np, out));
    FO_CHECK(copy_out(ctx, p, p0, np, (const char*)dOut, (const double*)dGrid, h));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

extern "C" int fo_per_bank_create(fo_ctx* ctx, const fo_per_params* p, const double* pos,
                                  int64_t nstruct, fo_bank** out) {
  FO_CHECK(check_params(ctx, p));
  if (!out || nstruct < 1 || !pos)
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_bank_create: bad argument");
  *out = nullptr;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  FO_CHECK(fo_ensure_perm(ctx, p->natoms));
  fo_bank* b = new (std::nothrow) fo_bank();
  if (!b) return fo_fail(ctx, FO_ERR_NOMEM, "out of host memory");
  b->kind = 1;
  b->nstruct = nstruct;
  b->ngroups = (int64_t)ctx->h_goff.size() - 1;
  b->nwave = p->nwave;
  b->natoms = p->natoms;
  b->per_struct_elems = (int64_t)bank_elems_per_struct(ctx, p);
  cudaError_t e = cudaMalloc(&b->d_data, (size_t)nstruct * b->per_struct_elems * 16);
  if (e != cudaSuccess) {
    delete b;
    return fo_fail(ctx, FO_ERR_NOMEM, "cudaMalloc of the structure-factor bank failed: %s",
                   cudaGetErrorString(e));
  }
  int64_t chunk = (int64_t)(((size_t)64 << 20) / ((size_t)p->natoms * 24));
  if (chunk < 1) chunk = 1;
  if (chunk > nstruct) chunk = nstruct;
  void* dpos;
  int rc = fo_scratch(ctx, FO_SCR_POSA, (size_t)chunk * p->natoms * 24, &dpos);
  for (int64_t s0 = 0; rc == FO_OK && s0 < nstruct; s0 += chunk) {
    const int64_t ns = (nstruct - s0 < chunk) ? nstruct - s0 : chunk;
    cudaError_t ce = cudaMemcpyAsync(dpos, pos + (size_t)s0 * p->natoms * 3,
                                     (size_t)ns * p->natoms * 24, cudaMemcpyHostToDevice, ctx->stream);
    if (ce != cudaSuccess) {
      rc = fo_fail(ctx, FO_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(ce));
      break;
    }
    rc = launch_sf(ctx, p, (const double*)dpos, ns,
                   (double2*)b->d_data + (size_t)s0 * b->per_struct_elems);
    if (rc == FO_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      rc = fo_fail(ctx, FO_ERR_CUDA, "structure-factor kernel failed");
  }
  if (rc != FO_OK) {
    cudaFree(b->d_data);
    delete b;
    return rc;
  }
  *out = b;
  return FO_OK;
}

extern "C" int fo_per_align_bank_ops(fo_ctx* ctx, const fo_per_params* p, const fo_bank* bank, const int64_t* pairs,
                                     const int32_t* ops, int64_t npairs, int64_t* best_idx, double* best_val,
                                     double* frac_idx, double* grid_out, int32_t* status);

extern "C" int fo_per_align_bank(fo_ctx* ctx, const fo_per_params* p, const fo_bank* bank,
                                 const int64_t* pairs, int64_t npairs, int64_t* best_idx,
                                 double* best_val, double* frac_idx, double* grid_out,
                                 int32_t* status) {
  return fo_per_align_bank_ops(ctx, p, bank, pairs, nullptr, npairs, best_idx, best_val, frac_idx, grid_out, status);
}

extern "C" int fo_per_align_bank_ops(fo_ctx* ctx, const fo_per_params* p, const fo_bank* bank, const int64_t* pairs,
                                     const int32_t* ops, int64_t npairs, int64_t* best_idx, double* best_val,
                                     double* frac_idx, double* grid_out, int32_t* status) {
  FO_CHECK(check_params(ctx, p));
  if (ops) {
    if (!(p->box[0] == p->box[1] && p->box[1] == p->box[2]))
      return fo_fail(ctx, FO_ERR_INVALID, "cell-symmetry operations need a cubic box");
    for (int64_t i = 0; i < npairs; ++i) {
      const int o = ops[i], a = o & 3, b = (o >> 2) & 3, c = (o >> 4) & 3;
      if (o < 0 || o >= 512 || a > 2 || b > 2 || c > 2 || a == b || a == c || b == c)
        return fo_fail(ctx, FO_ERR_INVALID, "operation %lld is not a signed permutation code", (long long)i);
    }
  }
  if (!bank || bank->kind != 1) return fo_fail(ctx, FO_ERR_INVALID, "not a periodic bank");
  if (bank->nwave != p->nwave || bank->ngroups != (int64_t)ctx->h_goff.size() - 1)
    return fo_fail(ctx, FO_ERR_INVALID, "bank was built with different nwave / perm groups");
  if (npairs < 0 || (npairs > 0 && (!pairs || !best_idx || !best_val || !frac_idx)))
    return fo_fail(ctx, FO_ERR_INVALID, "fo_per_align_bank: NULL argument");
  for (int64_t i = 0; i < 2 * npairs; ++i)
    if (pairs[i] < 0 || pairs[i] >= bank->nstruct)
      return fo_fail(ctx, FO_ERR_INVALID, "pair index %lld out of range", (long long)pairs[i]);
  if (npairs == 0) return FO_OK;
  FO_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t F3 = (size_t)p->nfspace * p->nfspace * p->nfspace;
  int64_t chunk = 1 << 16;
  if (grid_out) {
    chunk = (int64_t)(((size_t)512 << 20) / (F3 * 8));
    if (chunk < 1) chunk = 1;
  }
  if (chunk > npairs) chunk = npairs;
  void *dOut, *dGrid = nullptr, *dPairs;
  FO_CHECK(fo_scratch(ctx, FO_SCR_OUT, (size_t)chunk * 64, &dOut));
  FO_CHECK(fo_scratch(ctx, FO_SCR_MISC, (size_t)chunk * 20, &dPairs));
  int32_t* dOps = ops ? (int32_t*)((char*)dPairs + (size_t)chunk * 16) : nullptr;
  if (grid_out) FO_CHECK(fo_scratch(ctx, FO_SCR_GRID, (size_t)chunk * F3 * 8, &dGrid));
  HostOut h = {best_idx, best_val, frac_idx, grid_out, status};
  for (int64_t p0 = 0; p0 < npairs; p0 += chunk) {
    const int64_t np = (npairs - p0 < chunk) ? npairs - p0 : chunk;
    FO_CUDA(ctx, cudaMemcpyAsync(dPairs, pairs + 2 * p0, (size_t)np * 16, cudaMemcpyHostToDevice,
                                 ctx->stream));
    if (ops)
      FO_CUDA(ctx, cudaMemcpyAsync(dOps, ops + p0, (size_t)np * 4, cudaMemcpyHostToDevice, ctx->stream));
    XfOut out = make_out((char*)dOut, np, (double*)dGrid, status != nullptr);
    FO_CHECK(launch_xf(ctx, p, (const double2*)bank->d_data, (const double2*)bank->d_data,
                       (const long long*)dPairs, np, out, dOps));
    FO_CHECK(copy_out(ctx, p, p0, np, (const char*)dOut, (const double*)dGrid, h));
    FO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FO_OK;
}

extern "C" void fo_bank_destroy(fo_ctx* ctx, fo_bank* bank) {
  if (!bank) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  if (bank->d_data) cudaFree(bank->d_data);
  delete bank;
}
