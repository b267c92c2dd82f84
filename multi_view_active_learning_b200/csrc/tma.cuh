// mbarrier / bulk-copy (TMA) primitives shared by the persistent kernels (fused.cu, mapstream.cu).
#pragma once
#include "common.cuh"

namespace mval {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Orders this thread's generic-proxy accesses to shared memory before later async-proxy (TMA) accesses: needed when a
// consumer has WRITTEN into a stage that the producer is about to refill.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Watchdog: a wait that lasts longer than ~5 s of SM clocks is a protocol bug, not load.  Instead of hanging the
// device the waiter records who/what/where in `abort_rec` (8 x u64 in global memory: [0] flag, [1] code, [2] block,
// [3] warp, [4] iteration, [5] index), raises the flag, and every role drains out of its loops; the host reports
// MVAL_ERR_CUDA with the record.
// kBackoff: the waiter expects to wait long; it sleeps between polls so that it does not take issue slots from the
// other warps of its scheduler.
template <bool kBackoff = false>
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, unsigned long long* abort_rec, uint32_t code,
                                          long long iter, int index) {
  uint32_t ok;
  long long t0 = 0;
  uint32_t polls = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
    if (kBackoff) __nanosleep(256);
    if ((++polls & 255u) == 0u) {
      if (*((volatile unsigned long long*)&abort_rec[0]) != 0ull) return false;
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 10000000000ll) {
        if (atomicCAS(&abort_rec[0], 0ull, 1ull) == 0ull) {
          abort_rec[1] = code;
          abort_rec[2] = blockIdx.x;
          abort_rec[3] = threadIdx.x >> 5;
          abort_rec[4] = (unsigned long long)iter;
          abort_rec[5] = (unsigned long long)index;
          __threadfence();
        }
        return false;
      }
    }
  }
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace mval
