// Probe: which hardware warp slots (%warpid; SMSP = %warpid % 4) the warps of two co-resident
// 320-thread CTAs get.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/smsp_probe.cu -o scripts/smsp_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 2) probe(int* out, int regs_pad) {
  extern __shared__ double sm[];
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
  if ((threadIdx.x & 31) == 0) {
    int w = threadIdx.x >> 5;
    out[(blockIdx.x * 10 + w) * 2] = smid;
    out[(blockIdx.x * 10 + w) * 2 + 1] = wid;
  }
  if (regs_pad == 12345) sm[threadIdx.x] = 1.0;
}
int main() {
  const int nb = 296 * 2;
  int* d; cudaMalloc(&d, nb * 10 * 2 * sizeof(int));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
  probe<<<nb, 320, 72 * 1024>>>(d, 0);
  int* h = new int[nb * 20];
  cudaMemcpy(h, d, nb * 20 * sizeof(int), cudaMemcpyDeviceToHost);
  // per CTA: SMSP histogram; print the first 12 CTAs and a summary of distinct patterns
  int pat[4][4] = {}; 
  for (int b = 0; b < nb; ++b) {
    int c[4] = {0, 0, 0, 0};
    for (int w = 0; w < 10; ++w) c[h[(b * 10 + w) * 2 + 1] & 3]++;
    if (b < 8 || (b >= 296 && b < 304)) {
      printf("cta %d sm %d wids:", b, h[b * 20]);
      for (int w = 0; w < 10; ++w) printf(" %d", h[(b * 10 + w) * 2 + 1]);
      printf("  smsp counts %d %d %d %d\n", c[0], c[1], c[2], c[3]);
    }
    for (int q = 0; q < 4; ++q) pat[q][c[q] > 3 ? 3 : c[q]]++;
  }
  for (int q = 0; q < 4; ++q) printf("smsp %d: CTAs with 2 warps %d, 3 warps %d, other %d\n", q, pat[q][2], pat[q][3], pat[q][0] + pat[q][1]);
  return 0;
}
