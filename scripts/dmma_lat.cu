#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc, double a, double b) {
  double c0 = threadIdx.x, c1 = 1;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) dmma(c0, c1, a, b);
  long long t1 = clock64();
  double x = c0;
#pragma unroll
  for (int i = 0; i < 64; ++i) x = fma(x, a, b);
  long long t2 = clock64();
  double y = x;
#pragma unroll
  for (int i = 0; i < 64; ++i) y = __shfl_xor_sync(0xffffffffu, y, 4) + 1.0;
  long long t3 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
  out[threadIdx.x] = c0 + c1 + y;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 64);
  for (int threads : {32, 128, 512}) {
    k<<<1, threads>>>(d, c, 1.0000001, 0.5); k<<<1, threads>>>(d, c, 1.0000001, 0.5);
    long long h[3]; cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
    printf("threads %d: dependent DMMA %.1f cyc, dependent DFMA %.1f cyc, SHFL+DADD %.1f cyc\n", threads, h[0] / 64.0, h[1] / 64.0, h[2] / 64.0);
  }
  return 0;
}
