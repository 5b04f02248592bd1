// Microbenchmark: FP64 tensor-core (mma.sync.m8n8k4.f64) vs vector DFMA throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[16]; for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}
int main() {
  double* d; cudaMalloc(&d, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0); k_dmma<8><<<sms, threads>>>(d, iters, 1.0000001, 0.999999); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flop = 2.0 * 256 * 8 * (double)iters * (threads / 32) * sms;
      if (rep) printf("DMMA m8n8k4 x8 acc, %4d thr/SM: %.2f TFLOP/s (%.3f ms)\n", threads, flop / ms / 1e9, ms);
    }
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0); k_dfma<<<sms, threads>>>(d, iters, 1.0000001, 0.999999); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flop = 2.0 * 16 * (double)iters * threads * sms;
      if (rep) printf("DFMA x16 chains,        %4d thr/SM: %.2f TFLOP/s (%.3f ms)\n", threads, flop / ms / 1e9, ms);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
