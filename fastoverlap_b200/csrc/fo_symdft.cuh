// Symmetric 1-D DFT row primitive shared by the periodic (fo_periodic.cu) and spherical
// (fo_spherical.cu) transform kernels.
//
// Every 1-D transform of the hot path has inputs c_m, m = -K..K, and is evaluated per REAL row
// (re / im parts are separate rows) as
//     P[d] = c0 + sum_{m=1..K} E[m] cos(2 pi m d/F),    Q[d] = sum_{m=1..K} O[m] sin(2 pi m d/F)
// with E = c_m + c_-m, O = c_m - c_-m formed by the previous stage; outputs d and F-d then follow
// from P -+ iQ: a quarter of the dense-DFT flops, exact for any F.
#pragma once

// ------------------------------------------------------------------------------------------
// FP64 tensor-core form.  ncu (profiles/r01_summary.md) showed the scalar forms of this stage bound
// by operand delivery / issue slots, not by the FP64 pipe: a shared-memory-fed register tile is
// capped by the 128 B/clk LSU return path (16 doubles against 64 DFMA per clock and SM), a
// uniform-constant-operand form by one LDCU per DFMA and by latency at 8-16 warps per SM.  mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) runs at the same
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

// d = a b + (c0, c1): the first k-step of a chain takes its initial value as the C operand (no register moves)
__device__ __forceinline__ void fo_dmma3(double (&d)[2], double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d[0]), "=d"(d[1])
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
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
