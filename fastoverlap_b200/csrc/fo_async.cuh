// mbarrier + bulk asynchronous copy (the Blackwell / Hopper copy engine, SASS UBLKCP) primitives shared by the
// periodic and spherical transform kernels.
//
// Staging pattern used by per_xf6_kernel / sph_isoft3_kernel: one thread arms an mbarrier with the byte count
// (arrive.expect_tx) and issues cp.async.bulk.shared::cluster.global copies that complete on it; every thread
// that is about to read the staged block waits on the barrier's phase parity.  No thread spends issue slots or
// LDGSTS requests on the copy itself (the cp.async.cg loops this replaces issued one 16-byte request per
// thread and iteration), and the data is visible to a waiter without a CTA barrier.
#pragma once
#include <stdint.h>

__device__ __forceinline__ unsigned fo_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fo_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fo_smem_addr(bar)), "r"(count) : "memory");
}
// make the initialised barrier visible to the async proxy (before the first bulk copy names it)
__device__ __forceinline__ void fo_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fo_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fo_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void fo_mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fo_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fo_mbar_wait(uint64_t* bar, int parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "FO_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra FO_MBAR_DONE;\n"
      "bra FO_MBAR_WAIT;\n"
      "FO_MBAR_DONE:\n"
      "}" ::"r"(fo_smem_addr(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy accesses to shared memory made before this fence are ordered before later async-proxy
// (bulk copy) accesses: needed before a bulk copy overwrites a buffer that threads have just read
__device__ __forceinline__ void fo_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared, completing (complete_tx) on `bar`.  dst, src 16-byte aligned, bytes a
// multiple of 16.  Issued by ONE thread; split so that a single instruction stays well below the 2^20 limit.
__device__ __forceinline__ void fo_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  const unsigned piece = 32768;
  for (unsigned off = 0; off < bytes; off += piece) {
    const unsigned nb = bytes - off < piece ? bytes - off : piece;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     fo_smem_addr((const char*)dst + off)),
                 "l"((const char*)src + off), "r"(nb), "r"(fo_smem_addr(bar))
                 : "memory");
  }
}
