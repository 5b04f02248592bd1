// Microbenchmark: the instruction mix of per_sf3_kernel's k-loop (13 DMMA + 6 DMUL + 13 LDS.64 per warp and
// k-step, 10 warps per CTA, 2 CTAs per SM) with the pieces switched on one at a time:
//   mode 0: DMMA only, operands in registers     mode 1: + the DMULs that form A
//   mode 2: + operands from shared memory        mode 3: + one __syncthreads per 16 k-steps
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/dmma_mix.cu -o scripts/dmma_mix
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(320, 2) k_mix(double* out, int ksteps, double seed) {
  extern __shared__ double sm[];
  constexpr int PX = 20, PY = 28, PZ = 20, TA = 64;  // doubles per atom row
  double* dx = sm; double* dy = dx + TA * PX; double* dz = dy + TA * PY;
  for (int i = threadIdx.x; i < TA * (PX + PY + PZ); i += blockDim.x) sm[i] = seed + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  double acc[5][2][2], accl[3][2];
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = lane;
  for (int j = 0; j < 3; ++j) accl[j][0] = accl[j][1] = lane;
  const int offx0 = warp * 2 + ((g >> 1) & 1), offxl = warp * 2 + (g >> 2), offzl = 16 + (g & 3);
  int offy[5];
  for (int mt = 0; mt < 5; ++mt) offy[mt] = (2 * mt + (g >> 2)) * 2 + (g & 1);
  double rx = seed * (lane + 1), ry[5], rz[2], rl[3];
  for (int i = 0; i < 5; ++i) ry[i] = seed + i;
  for (int i = 0; i < 2; ++i) rz[i] = seed - i;
  for (int i = 0; i < 3; ++i) rl[i] = seed * i;
  for (int k0 = 0; k0 < ksteps; k0 += 4) {
    const int a = (k0 & (TA - 1)) + t4;
    if (MODE >= 2) {
      const double* xr = dx + a * PX; const double* yr = dy + a * PY; const double* zr = dz + a * PZ;
      double bz[2];
      for (int nt = 0; nt < 2; ++nt) bz[nt] = zr[nt * 8 + g];
      const double xv = xr[offx0];
#pragma unroll
      for (int mt = 0; mt < 5; ++mt) {
        const double av = xv * yr[offy[mt]];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) dmma(acc[mt][nt], av, bz[nt]);
      }
      const double avl = xr[offxl] * zr[offzl];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) dmma(accl[nt], avl, yr[nt * 8 + g]);
    } else {
#pragma unroll
      for (int mt = 0; mt < 5; ++mt) {
        const double av = MODE >= 1 ? rx * ry[mt] : ry[mt];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) dmma(acc[mt][nt], av, rz[nt]);
      }
      const double avl = MODE >= 1 ? rx * rz[0] : rz[1];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) dmma(accl[nt], avl, rl[nt]);
      if (MODE >= 1) rx = -rx;  // keep the DMULs in the loop
    }
    if (MODE >= 3 && ((k0 >> 2) & 15) == 15) __syncthreads();
  }
  double s = 0;
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 2; ++j) s += acc[i][j][0] + acc[i][j][1];
  for (int j = 0; j < 3; ++j) s += accl[j][0] + accl[j][1];
  if (s == 123.456) out[0] = s;
}
template <int MODE> void run(double* d, int sms) {
  const int ksteps = 4 * 20000;  // 20000 k-steps
  const size_t smem = 64 * (20 + 28 + 20) * 8 * 2;  // as per_sf3: two buffers
  cudaFuncSetAttribute(k_mix<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_mix<MODE><<<sms * 2, 320, smem>>>(d, ksteps, 1.0000001); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 256 * 13 * 20000.0 * 10 * sms * 2;
    if (rep) printf("mode %d: %.2f TFLOP/s in DMMA (%.3f ms)\n", MODE, flop / ms / 1e9, ms);
  }
}
int main() {
  double* d; cudaMalloc(&d, 64);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>(d, sms); run<1>(d, sms); run<2>(d, sms); run<3>(d, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
