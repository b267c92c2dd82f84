// Shared plumbing for the mval_b200 C-ABI library: error reporting, launch accounting, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mval_b200.h"

namespace mval {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
int require_device();
int num_sms();

#define MVAL_CUDA(call)                                        \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return mval::cuda_fail(e__, #call); \
  } while (0)

#define MVAL_LAUNCH_CHECK(name)                                      \
  do {                                                               \
    mval::count_launch();                                            \
    cudaError_t e__ = cudaGetLastError();                            \
    if (e__ != cudaSuccess) return mval::cuda_fail(e__, "launch " name); \
  } while (0)

#define MVAL_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      mval::set_error(__VA_ARGS__);      \
      return MVAL_ERR_INVALID_ARGUMENT;  \
    }                                    \
  } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// 128-bit streaming load: read-only path, do not allocate in L1 (each heat-map byte is read exactly once).
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// Monotone key of a float for arg-max with torch.argmax semantics: NaN is the maximum, -0.0 == +0.0.
// x + 0.0f turns -0.0 into +0.0 and any NaN into the canonical 0x7fffffff, which maps to 0xffffffff.
__device__ __forceinline__ uint32_t argmax_key(float x) {
  uint32_t u = __float_as_uint(x + 0.0f);
  uint32_t m = (uint32_t)((int32_t)u >> 31) | 0x80000000u;
  return u ^ m;
}
__device__ __forceinline__ float argmax_key_to_float(uint32_t k) {
  uint32_t m = (k & 0x80000000u) ? 0x80000000u : 0xffffffffu;
  return __uint_as_float(k ^ m);
}

}  // namespace mval
