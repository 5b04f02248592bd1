// Symmetric 1-D DFT row primitive shared by the periodic (fo_periodic.cu) and spherical
// (fo_spherical.cu) transform kernels.
//
// Every 1-D transform of the hot path has inputs c_m, m = -K..K, and is evaluated per REAL row
// (re / im parts are separate rows) as
//     P[d] = c0 + sum_{m=1..K} E[m] cos(2 pi m d/F),    Q[d] = sum_{m=1..K} O[m] sin(2 pi m d/F)
// with E = c_m + c_-m, O = c_m - c_-m formed by the previous stage; outputs d and F-d then follow
// from P -+ iQ.  A thread owns one row and walks over the outputs in chunks of DC: per (m, chunk)
// it loads two doubles from shared memory for 2 DC DFMA.  The twiddle matrices live in the
// kernel-parameter constant bank (__grid_constant__ FoTw): their indices are warp-uniform, so they
// reach the DFMA as uniform-register operands (SASS: LDCU.64 + DFMA R, R, UR, R) and cost no
// shared-memory or register-file bandwidth.  Why not a shared-memory-fed register tile: the LSU
// returns 128 B/clk/SM = 16 doubles against 64 DFMA/clk/SM, so every loaded double has to feed
// >= 4 DFMA (measured: profiles/r01_summary.md).
//
// Rules that keep ptxas on the uniform path (found the hard way):
//   * all layout integers must come from the parameter bank (no integer division in the kernel);
//   * control flow around sym_row must be warp-uniform (clamp invalid rows, mask the stores);
//   * no per-thread global stores and no 64-bit index arithmetic in the epilogue.
#pragma once

constexpr int FO_TWMAX = 1024;  // doubles per twiddle table in the parameter bank

struct FoTw {
  double c[FO_TWMAX];  // [m-1][HP]  cos(2 pi m d / F), zero for d >= H
  double s[FO_TWMAX];  // [m-1][HP]  sin
};

// e / o point at E[1][row] / O[1][row]; consecutive m are kstride doubles apart.
template <int DC, class Epi>
__device__ __forceinline__ void sym_row(const FoTw& tw, const double* __restrict__ e,
                                        const double* __restrict__ o, int kstride, int K, int HP,
                                        double c0, int d_begin, int d_end, Epi&& epi) {
  // The epilogue gets its own per-thread output counter dv, started from an opaque copy of d_begin:
  // the compiler cannot merge it with the loop counter d0, so d0 (and with it every table index)
  // stays in uniform registers even when the epilogue uses dv for per-thread addressing.
  int dv;
  asm volatile("mov.s32 %0, %1;" : "=r"(dv) : "r"(d_begin));
  for (int d0 = d_begin; d0 < d_end; d0 += DC, dv += DC) {
    double P[DC], Q[DC];
#pragma unroll
    for (int t = 0; t < DC; ++t) {
      P[t] = c0;
      Q[t] = 0.0;
    }
    const double* ep = e;
    const double* op = o;
    int ti = d0;
    for (int k = 0; k < K; ++k) {
      const double ev = *ep, ov = *op;
#pragma unroll
      for (int t = 0; t < DC; ++t) {
        P[t] = fma(ev, tw.c[ti + t], P[t]);
        Q[t] = fma(ov, tw.s[ti + t], Q[t]);
      }
      ep += kstride;
      op += kstride;
      ti += HP;
    }
    epi(dv, P, Q);
  }
}

// host: fill the tables for transform length F, K harmonics, H = F/2+1 outputs padded to HP
inline void fo_fill_tw(FoTw& tw, int K, int F, int H, int HP) {
  const double twopi = 6.283185307179586476925286766559;
  for (int m = 1; m <= K; ++m)
    for (int d = 0; d < HP; ++d) {
      const double ang = twopi * (double)((m * d) % F) / (double)F;
      tw.c[(m - 1) * HP + d] = d < H ? cos(ang) : 0.0;
      tw.s[(m - 1) * HP + d] = d < H ? sin(ang) : 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// FP64 tensor-core form of the same stage.  ncu (profiles/r01_summary.md) shows the scalar forms
// are bound by operand delivery / issue slots, not by the FP64 pipe: a shared-memory-fed register
// tile is capped by the 128 B/clk LSU return path, the uniform-operand form by one LDCU per DFMA
// and by latency at 8-16 warps per SM.  mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) runs at the same
// 37 TFLOP/s as DFMA on B200 (scripts/dmma_peak.cu: 37.2 vs 33.9 TFLOP/s measured) while one
// instruction carries 256 MACs and the operands are spread over the warp, so the contraction-bound
// stages go to the tensor pipe (BASELINE.json north_star: "... placed on the FP64 tensor-core pipe
// only where ncu shows it is contraction-bound").
//
// A warp owns a tile of 8 real rows and all NT*8 outputs:
//     P[8 x 8NT] = c0 + E[8 x 4KS] COS[4KS x 8NT],    Q = O SIN
// Fragment layout of m8n8k4 (g = lane >> 2, t = lane & 3):
//     A: lane holds A[row g][k t]      B: lane holds B[k t][col g]      C: lane holds C[row g][cols 2t, 2t+1]
// The twiddle fragments stay in registers for the lifetime of the CTA (2 KS NT doubles per lane).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void fo_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <int KS, int NT>
struct SymMma {
  double bc[KS][NT], bs[KS][NT];

  // K harmonics (m = 1..K), transform length F, outputs d = 0..H-1 (H = F/2+1 <= 8 NT)
  __device__ __forceinline__ void init(int K, int F, int H, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int m = ks * 4 + t + 1, d = nt * 8 + g;
        double sn = 0.0, cs = 0.0;
        if (m <= K && d < H) sincospi(2.0 * (double)((m * d) % F) / (double)F, &sn, &cs);
        bc[ks][nt] = cs;
        bs[ks][nt] = sn;
      }
  }

  // e1 / o1 point at E[m=1][first row of the tile] / O[m=1][..]; consecutive m are kstride doubles
  // apart, consecutive rows 1 double.  Rows beyond K are clamped (their twiddle fragments are 0).
  // c0 is the lane's row constant (row g).  Result: P[nt][j], Q[nt][j] at (row g, d = 8nt + 2t + j).
  __device__ __forceinline__ void run(const double* __restrict__ e1, const double* __restrict__ o1,
                                      int kstride, int K, double c0, int lane, double (&P)[NT][2],
                                      double (&Q)[NT][2]) const {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      P[nt][0] = P[nt][1] = c0;
      Q[nt][0] = Q[nt][1] = 0.0;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      int m = ks * 4 + t;
      m = m < K ? m : K - 1;
      const double ae = e1[(size_t)m * kstride + g];
      const double ao = o1[(size_t)m * kstride + g];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        fo_dmma(P[nt], ae, bc[ks][nt]);
        fo_dmma(Q[nt], ao, bs[ks][nt]);
      }
    }
  }
};
