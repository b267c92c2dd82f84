// Shared plumbing for the mval_b200 C-ABI library: error reporting, launch accounting, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
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

// ---------------------------------------------------------------------------------------------------------------
// Arg-max of one map by one warp with torch.argmax semantics (first index of the maximum, NaN is the maximum,
// -0.0 == +0.0).  `load(i)` returns the i-th float4 of the map (global streaming load or shared memory).
//
// Fast path, branch-free, ~2.75 instructions per element: per 128-bit vector 3 FMNMX (vector max), one compare +
// two selects keeping the lane's running (best value, index of the best VECTOR), and a NaN/inf sentinel
// (sentinel = fma(x+y+z+w, 0, sentinel) turns NaN as soon as a vector holds a NaN or an infinity).  The component
// inside the winning vector is recovered afterwards by re-loading that one vector.  If any lane's sentinel fired
// the warp re-scans the map with the exact monotone-key compare (slow path, ~7 instructions per element; never
// taken on finite heat maps).  Returns the flat index to all lanes; *peak_key (optional) receives the monotone key
// of the maximum.
// ---------------------------------------------------------------------------------------------------------------
// kReload: recover the winning component by re-loading the best vector after the scan (cheap from shared memory);
// otherwise the best vector's four values ride along in registers (4 more selects per vector, but no dependent
// global load at the tail of a warp that only lives for one map).
template <int kUnroll, bool kReload, typename LoadFn>
__device__ __forceinline__ uint32_t warp_argmax_map(LoadFn load, int hw4, int lane, uint32_t* peak_key) {
  float best = -INFINITY;
  int best_vec = lane;  // stays put when every element of the lane is -inf
  float4 best_v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  float sentinel = 0.0f;
  for (int base = lane; base < hw4; base += kWarp * kUnroll) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kWarp;
      if (i < hw4) v[u] = load(i);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kWarp;
      if (i < hw4) {
        const float m = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
        sentinel = fmaf((v[u].x + v[u].y) + (v[u].z + v[u].w), 0.0f, sentinel);
        const bool gt = m > best;  // strict: the earlier vector is kept on ties
        best = gt ? m : best;
        best_vec = gt ? i : best_vec;
        if (!kReload) {
          best_v.x = gt ? v[u].x : best_v.x;
          best_v.y = gt ? v[u].y : best_v.y;
          best_v.z = gt ? v[u].z : best_v.z;
          best_v.w = gt ? v[u].w : best_v.w;
        }
      }
    }
  }
  uint32_t best_idx;
  {
    const float4 w = kReload ? load(best_vec < hw4 ? best_vec : 0) : best_v;
    best_idx = (uint32_t)best_vec * 4u + (w.x == best ? 0u : (w.y == best ? 1u : (w.z == best ? 2u : 3u)));
  }
  const bool suspicious = (sentinel != sentinel);
  uint32_t best_key = argmax_key(best);
  if (__any_sync(kFull, suspicious)) {
    best_key = 0u;  // every real key is >= 0x007fffff, so the first element always wins
    best_idx = 0u;
    for (int base = lane; base < hw4; base += kWarp) {
      const float4 v = load(base);
      const uint32_t e = (uint32_t)base * 4u;
      uint32_t k;
      k = argmax_key(v.x); if (k > best_key) { best_key = k; best_idx = e; }
      k = argmax_key(v.y); if (k > best_key) { best_key = k; best_idx = e + 1; }
      k = argmax_key(v.z); if (k > best_key) { best_key = k; best_idx = e + 2; }
      k = argmax_key(v.w); if (k > best_key) { best_key = k; best_idx = e + 3; }
    }
  }
  // each lane holds the first maximum of its own elements; across lanes: highest key, then lowest index
  const uint32_t top = __reduce_max_sync(kFull, best_key);
  const uint32_t idx = __reduce_min_sync(kFull, best_key == top ? best_idx : 0xffffffffu);
  if (peak_key) *peak_key = top;
  return idx;
}

}  // namespace mval
