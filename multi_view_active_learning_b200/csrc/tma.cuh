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
// device the waiter records who/what/where in `abort_rec` (kWdWords x u64 in global memory: [0] flag, [1] code,
// [2] block, [3] warp, [4] iteration, [5] index), raises the flag, and every role drains out of its loops.  The record
// is mirrored into pinned host memory ([kWdMirror] holds its address), which the host reads -- without touching the
// device -- at the next entry point of the library and in mval_check_async: a tripped launch is reported as
// MVAL_ERR_CUDA and the record is cleared, so that later launches run normally again (WatchdogHost below).
// [kWdTimeout] overrides the time-out in SM clocks (0 = default) and [kWdStall] != 0 makes the producers of the
// persistent kernels issue nothing: test hooks behind mval_debug_watchdog.
constexpr int kWdWords = 16, kWdTimeout = 8, kWdMirror = 9, kWdStall = 10;
constexpr long long kWdDefaultCycles = 10000000000ll;

__device__ __forceinline__ bool watchdog_stalled(const unsigned long long* abort_rec) {
  return *((const volatile unsigned long long*)&abort_rec[kWdStall]) != 0ull;
}

// kBackoff (1 long, 2 short): the waiter expects to wait; it sleeps between polls so that it does not take issue slots from the
// other warps of its scheduler.
// The poll loop with its watchdog is ONE out-of-line function per flavour: inlined at every wait site it was 9 KB of the
// fused kernels' SASS (round 2), and those kernels stall on instruction fetch.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok;
}

template <int kBackoff>
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar_addr, uint32_t parity, unsigned long long* abort_rec, uint32_t code,
                                            long long iter, int index) {
  long long t0 = 0;
  uint32_t polls = 0;
  for (;;) {
    // A long waiter (the RANSAC warps wait ~50 us for the key-points of a frame, with up to three frames of slack) sleeps
    // between polls: try_wait alone comes back every ~0.1 us whatever its suspend-time hint says (measured, round 2: the poll
    // loop was 25 % of all executed warp instructions of the fused kernel), a 2-4 us sleep cuts that by ~20x.
    // kBackoff == 2: the producer of an issue-bound kernel (MPE / BSB scoring).  It waits for a free stage almost always,
    // and its bare poll loop was 440 of the 2 970 warp instructions per map of the fused MPE pass (profiles/r2u_*); a stage
    // frees every ~0.6 us, so a short sleep costs no bandwidth and gives the slots to the decode warps.
    if (kBackoff == 1) __nanosleep(polls < 2u ? 500u : 3000u);
    if (kBackoff == 2) __nanosleep(250u);
    if (mbar_try_wait(bar_addr, parity)) return true;
    if ((++polls & (kBackoff ? 15u : 255u)) == 0u) {
      if (*((volatile unsigned long long*)&abort_rec[0]) != 0ull) return false;
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      const long long lim = (long long)*((volatile unsigned long long*)&abort_rec[kWdTimeout]);
      if (now - t0 > (lim > 0 ? lim : kWdDefaultCycles)) {
        if (atomicCAS(&abort_rec[0], 0ull, 1ull) == 0ull) {
          abort_rec[1] = code;
          abort_rec[2] = blockIdx.x;
          abort_rec[3] = threadIdx.x >> 5;
          abort_rec[4] = (unsigned long long)iter;
          abort_rec[5] = (unsigned long long)index;
          __threadfence();
          volatile unsigned long long* mirror =
              reinterpret_cast<volatile unsigned long long*>(*((volatile unsigned long long*)&abort_rec[kWdMirror]));
          if (mirror != nullptr) {
            mirror[1] = code;
            mirror[2] = blockIdx.x;
            mirror[3] = threadIdx.x >> 5;
            mirror[4] = (unsigned long long)iter;
            mirror[5] = (unsigned long long)index;
            __threadfence_system();
            mirror[0] = 1ull;
            __threadfence_system();
          }
        }
        return false;
      }
    }
  }
}

template <int kBackoff = 0>
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, unsigned long long* abort_rec, uint32_t code,
                                          long long iter, int index) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return true;
  return mbar_wait_slow<kBackoff>(addr, parity, abort_rec, code, iter, index);
}

// Host side of the watchdog of one persistent kernel family (one instance per translation unit, next to its
// __device__ record).  prepare() runs before every launch: the first time on a device it allocates the pinned mirror
// and plants its address in the device record; every time it looks at the mirror (a host memory read) and, if an
// earlier launch tripped, drains the device, clears both copies and fails THIS call with the record in the message.
struct WatchdogHost {
  static constexpr int kMaxDevices = 64;
  unsigned long long* mirror[kMaxDevices] = {};

  template <typename Symbol>
  int prepare(const Symbol& symbol, const char* who) {
    int dev = 0;
    MVAL_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return MVAL_OK;
    if (mirror[dev] == nullptr) {
      unsigned long long* m = nullptr;
      MVAL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&m), sizeof(unsigned long long) * kWdWords,
                              cudaHostAllocMapped | cudaHostAllocPortable));
      for (int i = 0; i < kWdWords; ++i) m[i] = 0ull;
      const unsigned long long addr = (unsigned long long)reinterpret_cast<uintptr_t>(m);
      MVAL_CUDA(cudaMemcpyToSymbol(symbol, &addr, sizeof(addr), sizeof(unsigned long long) * kWdMirror));
      mirror[dev] = m;
    }
    return poll(symbol, who, dev);
  }

  // Reports (and clears) a trip recorded on the current device; does not wait for running work.
  template <typename Symbol>
  int poll(const Symbol& symbol, const char* who, int dev = -1) {
    if (dev < 0) MVAL_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || mirror[dev] == nullptr) return MVAL_OK;
    volatile unsigned long long* m = mirror[dev];
    if (m[0] == 0ull) return MVAL_OK;
    cudaDeviceSynchronize();  // let the tripped launch drain before its record is cleared
    unsigned long long rec[6];
    for (int i = 0; i < 6; ++i) rec[i] = m[i];
    unsigned long long zero[6] = {0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(symbol, zero, sizeof(zero));
    for (int i = 0; i < 6; ++i) m[i] = 0ull;
    (void)cudaGetLastError();
    set_error("%s watchdog: an mbarrier wait (code %llu) timed out in block %llu warp %llu at iteration %llu index %llu; "
              "the outputs of that launch are incomplete", who, rec[1], rec[2], rec[3], rec[4], rec[5]);
    return MVAL_ERR_CUDA;
  }

  template <typename Symbol>
  int debug_set(const Symbol& symbol, unsigned long long timeout_cycles, int stall) {
    const unsigned long long v[3] = {timeout_cycles, 0ull, stall ? 1ull : 0ull};
    MVAL_CUDA(cudaMemcpyToSymbol(symbol, &v[0], sizeof(unsigned long long), sizeof(unsigned long long) * kWdTimeout));
    MVAL_CUDA(cudaMemcpyToSymbol(symbol, &v[2], sizeof(unsigned long long), sizeof(unsigned long long) * kWdStall));
    return MVAL_OK;
  }
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace mval
