// Coreset k-center greedy (reference utils/coreset.py:49-95) on float32 features, sm_100a.
//
// One greedy step of the reference is  dist = pairwise_distances(features, features[[ind]])  followed by
// np.minimum(min_distances, dist)  and the next  np.argmax(min_distances): a GEMV, an element-wise min and an
// arg-max.  kcenter_update_kernel fuses all three: one warp per feature row streams the row once (128-bit
// read-only loads), forms the distance to the new centre, updates the running minimum in place and carries the
// running (max value, lowest index) through warp -> block -> grid reductions in the same launch (the last block
// to finish folds the per-block partials; no second launch, no host round trip).  Bytes per step = n*d*4 (+ 12 n).
//
// Arithmetic is float32 in the *canonical summation order* of oracle/coreset_oracle.py so that the selected
// indices are bit-reproducible on the CPU: element e -> (chunk e/128, lane (e%128)/4, slot e%4); per (lane, slot)
// sequential accumulation over chunks with separately rounded multiply and add (no FMA contraction); lane total
// (a0+a1)+(a2+a3); xor-butterfly 16,8,4,2,1; d2 = ((-2*dot) + |x|^2) + |c|^2; d = sqrt(max(d2, 0)).
#include <math.h>

#include "common.cuh"

namespace mval {

constexpr int kKcThreads = 256;
constexpr int kKcWarps = kKcThreads / kWarp;

template <bool kVec>
__device__ __forceinline__ float canonical_dot(const float* __restrict__ x, const float* __restrict__ c, int d, int lane) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (kVec) {
    const int n4 = d >> 2;  // d % 4 == 0 and both pointers 16-byte aligned
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    const float4* __restrict__ c4 = reinterpret_cast<const float4*>(c);
#pragma unroll 8
    for (int i = lane; i < n4; i += kWarp) {
      const float4 xv = ld_stream_f4(x4 + i);
      const float4 cv = c4[i];
      a0 = __fadd_rn(a0, __fmul_rn(xv.x, cv.x));
      a1 = __fadd_rn(a1, __fmul_rn(xv.y, cv.y));
      a2 = __fadd_rn(a2, __fmul_rn(xv.z, cv.z));
      a3 = __fadd_rn(a3, __fmul_rn(xv.w, cv.w));
    }
  } else {
    for (int e = lane * 4; e < d; e += kWarp * 4) {
      if (e < d) a0 = __fadd_rn(a0, __fmul_rn(x[e], c[e]));
      if (e + 1 < d) a1 = __fadd_rn(a1, __fmul_rn(x[e + 1], c[e + 1]));
      if (e + 2 < d) a2 = __fadd_rn(a2, __fmul_rn(x[e + 2], c[e + 2]));
      if (e + 3 < d) a3 = __fadd_rn(a3, __fmul_rn(x[e + 3], c[e + 3]));
    }
  }
  float s = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(kFull, s, o));
  return s;
}

template <bool kVec>
__global__ void __launch_bounds__(kKcThreads)
kcenter_norms_kernel(const float* __restrict__ feat, int64_t n, int d, float* __restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * kKcWarps;
  for (int64_t row = (int64_t)blockIdx.x * kKcWarps + (threadIdx.x >> 5); row < n; row += warps) {
    const float* x = feat + row * d;
    const float s = canonical_dot<kVec>(x, x, d, lane);
    if (lane == 0) norms[row] = s;
  }
}

struct KcPartial {
  float val;
  int64_t idx;
};

__device__ __forceinline__ bool kc_better(float v, int64_t i, float bv, int64_t bi) {
  return v > bv || (v == bv && i < bi);
}

// Candidate record exchanged between ranks in the multi-GPU loop: {float val; int32 pad; int64 idx; float row[d4]}
// with d4 = d rounded up to a multiple of 4 (16-byte aligned records).  idx < 0 marks an empty shard.
__host__ __device__ inline size_t kc_record_bytes(int d) { return 16 + sizeof(float) * (size_t)((d + 3) & ~3); }

struct KcArgs {
  const float* feat;
  const float* norms;
  int64_t n;
  int d;
  const float* centre;        // explicit centre (d floats), or
  const int64_t* centre_idx;  // row (*centre_idx - index_offset) of feat, or
  const char* cands_in;       // the best of n_cands candidate records (highest val, lowest idx)
  int n_cands;
  float* min_dist;
  int64_t index_offset;
  KcPartial* partials;
  unsigned int* done_counter;
  float* out_best_val;    // local arg-max of the updated min_dist (may be null)
  int64_t* out_best_idx;  // global index = local + index_offset (may be null)
  int64_t* also_idx;      // second copy of out_best_idx, or -- with cands_in -- the index of the chosen centre
  char* cand_out;         // candidate record of the local arg-max, row included (may be null)
};

template <bool kVec>
__global__ void __launch_bounds__(kKcThreads) kcenter_update_kernel(const KcArgs a) {
  extern __shared__ __align__(16) float s_centre[];  // d floats (+ padding)
  __shared__ float s_cc;
  __shared__ KcPartial s_part[kKcWarps];
  __shared__ bool s_last;
  __shared__ int s_win;
  __shared__ KcPartial s_best;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = a.d;
  const float* c = a.centre;
  if (a.cands_in != nullptr) {
    if (threadIdx.x == 0) {
      const size_t rb = kc_record_bytes(d);
      int win = -1;
      float bv = 0.f;
      int64_t bi = 0;
      for (int r = 0; r < a.n_cands; ++r) {
        const float v = *reinterpret_cast<const float*>(a.cands_in + r * rb);
        const int64_t i = *reinterpret_cast<const int64_t*>(a.cands_in + r * rb + 8);
        if (i >= 0 && (win < 0 || kc_better(v, i, bv, bi))) { win = r; bv = v; bi = i; }
      }
      s_win = win < 0 ? 0 : win;
      if (blockIdx.x == 0 && a.also_idx) *a.also_idx = win < 0 ? -1 : bi;
    }
    __syncthreads();
    c = reinterpret_cast<const float*>(a.cands_in + s_win * kc_record_bytes(d) + 16);
  } else if (a.centre_idx != nullptr) {
    c = a.feat + (*a.centre_idx - a.index_offset) * d;
  }
  for (int i = threadIdx.x; i < d; i += kKcThreads) s_centre[i] = c[i];
  __syncthreads();
  if (warp == 0) {
    const float cc = canonical_dot<false>(s_centre, s_centre, d, lane);
    if (lane == 0) s_cc = cc;
  }
  __syncthreads();
  const float cc = s_cc;
  float best_v = -INFINITY;
  int64_t best_i = INT64_MAX;
  const int64_t warps = (int64_t)gridDim.x * kKcWarps;
  for (int64_t row = (int64_t)blockIdx.x * kKcWarps + warp; row < a.n; row += warps) {
    const float dot = canonical_dot<kVec>(a.feat + row * d, s_centre, d, lane);
    if (lane == 0) {
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), a.norms[row]), cc);
      const float dist = __fsqrt_rn(fmaxf(d2, 0.0f));
      const float m = fminf(a.min_dist[row], dist);
      a.min_dist[row] = m;
      if (kc_better(m, row, best_v, best_i)) { best_v = m; best_i = row; }
    }
  }
  if (lane == 0) s_part[warp] = KcPartial{best_v, best_i};
  __syncthreads();
  if (threadIdx.x == 0) {
    KcPartial b = s_part[0];
    for (int w = 1; w < kKcWarps; ++w)
      if (kc_better(s_part[w].val, s_part[w].idx, b.val, b.idx)) b = s_part[w];
    a.partials[blockIdx.x] = b;
    __threadfence();
    const unsigned int ticket = atomicAdd(a.done_counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (warp == 0) {
    __threadfence();
    KcPartial b{-INFINITY, INT64_MAX};
    for (int i = lane; i < (int)gridDim.x; i += kWarp) {
      const float pv = __ldcg(&a.partials[i].val);
      const int64_t pi = __ldcg(&a.partials[i].idx);
      if (kc_better(pv, pi, b.val, b.idx)) { b.val = pv; b.idx = pi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(kFull, b.val, o);
      const int64_t oi = __shfl_xor_sync(kFull, b.idx, o);
      if (kc_better(ov, oi, b.val, b.idx)) { b.val = ov; b.idx = oi; }
    }
    if (lane == 0) {
      const int64_t g = (b.idx == INT64_MAX) ? -1 : b.idx + a.index_offset;
      if (a.out_best_val) *a.out_best_val = b.val;
      if (a.out_best_idx) *a.out_best_idx = g;
      if (a.also_idx && a.cands_in == nullptr) *a.also_idx = g;
      if (a.cand_out) {
        *reinterpret_cast<float*>(a.cand_out) = b.val;
        *reinterpret_cast<int64_t*>(a.cand_out + 8) = g;
      }
      s_best = b;
      *a.done_counter = 0u;  // ready for the next launch on this stream
    }
  }
  __syncthreads();
  if (a.cand_out != nullptr && s_best.idx != INT64_MAX) {  // ship the local winner's row with its record
    const float* row = a.feat + s_best.idx * d;
    float* dst = reinterpret_cast<float*>(a.cand_out + 16);
    for (int i = threadIdx.x; i < d; i += kKcThreads) dst[i] = row[i];
  }
}

__global__ void kcenter_fill_kernel(float* __restrict__ p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// Per-device scratch for the grid-level arg-max (partials + ticket counter).  Launches that share it must be
// stream-ordered with respect to each other (one selection loop per device at a time), which is how the
// reference's single-threaded loop behaves anyway.
struct KcScratch {
  KcPartial* partials = nullptr;
  unsigned int* counter = nullptr;
  int capacity = 0;
};
static KcScratch g_scratch[64];

static int get_scratch(int grid, KcScratch** out) {
  int dev = 0;
  MVAL_CUDA(cudaGetDevice(&dev));
  MVAL_REQUIRE(dev < 64, "kcenter: device ordinal too large");
  KcScratch& s = g_scratch[dev];
  if (s.capacity < grid) {
    if (s.partials) cudaFree(s.partials);
    if (s.counter) cudaFree(s.counter);
    s = KcScratch();
    MVAL_CUDA(cudaMalloc(&s.partials, sizeof(KcPartial) * grid));
    MVAL_CUDA(cudaMalloc(&s.counter, sizeof(unsigned int)));
    MVAL_CUDA(cudaMemset(s.counter, 0, sizeof(unsigned int)));
    s.capacity = grid;
  }
  *out = &s;
  return MVAL_OK;
}

static int kc_grid(int64_t n) {
  const int64_t want = (n + kKcWarps - 1) / kKcWarps;
  const int64_t cap = (int64_t)num_sms() * 8;  // 8 resident 256-thread blocks per SM
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

static bool vec_ok(const float* feat, int d) { return d % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0; }

static int kcenter_launch(KcArgs a, cudaStream_t stream) {
  const int grid = kc_grid(a.n);
  KcScratch* s = nullptr;
  if (int rc = get_scratch(num_sms() * 8, &s)) return rc;
  a.partials = s->partials;
  a.done_counter = s->counter;
  const size_t smem = sizeof(float) * ((a.d + 3) & ~3);
  if (vec_ok(a.feat, a.d)) {
    if (smem > 48 * 1024)
      MVAL_CUDA(cudaFuncSetAttribute(kcenter_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kcenter_update_kernel<true><<<grid, kKcThreads, smem, stream>>>(a);
  } else {
    if (smem > 48 * 1024)
      MVAL_CUDA(cudaFuncSetAttribute(kcenter_update_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kcenter_update_kernel<false><<<grid, kKcThreads, smem, stream>>>(a);
  }
  MVAL_LAUNCH_CHECK("kcenter_update");
  return MVAL_OK;
}

int kcenter_update(const float* feat, const float* norms, int64_t n, int d, const float* centre,
                   const int64_t* centre_idx, float* min_dist, int64_t index_offset, float* out_best_val,
                   int64_t* out_best_idx, int64_t* also_idx, cudaStream_t stream) {
  KcArgs a{};
  a.feat = feat; a.norms = norms; a.n = n; a.d = d; a.centre = centre; a.centre_idx = centre_idx;
  a.min_dist = min_dist; a.index_offset = index_offset; a.out_best_val = out_best_val; a.out_best_idx = out_best_idx;
  a.also_idx = also_idx;
  return kcenter_launch(a, stream);
}

}  // namespace mval

extern "C" int mval_kcenter_norms(const float* features, int64_t n, int d, float* row_norms, void* stream) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && d > 0, "mval_kcenter_norms: bad shape");
  if (n == 0) return MVAL_OK;
  MVAL_REQUIRE(features && row_norms, "mval_kcenter_norms: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec_ok(features, d))
    kcenter_norms_kernel<true><<<kc_grid(n), kKcThreads, 0, st>>>(features, n, d, row_norms);
  else
    kcenter_norms_kernel<false><<<kc_grid(n), kKcThreads, 0, st>>>(features, n, d, row_norms);
  MVAL_LAUNCH_CHECK("kcenter_norms");
  return MVAL_OK;
}

extern "C" int mval_kcenter_update(const float* features, const float* row_norms, int64_t n, int d, const float* centre,
                                   float* min_dist, int64_t index_offset, float* out_best_val, int64_t* out_best_idx,
                                   void* stream) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && d > 0, "mval_kcenter_update: bad shape");
  MVAL_REQUIRE(centre && out_best_val && out_best_idx, "mval_kcenter_update: null pointer");
  MVAL_REQUIRE(n == 0 || (features && row_norms && min_dist), "mval_kcenter_update: null pointer");
  MVAL_REQUIRE((size_t)d * 4 <= 200 * 1024, "mval_kcenter_update: feature dimension too large for shared memory");
  return kcenter_update(features, row_norms, n, d, centre, nullptr, min_dist, index_offset, out_best_val, out_best_idx,
                        nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int mval_kcenter_greedy(const float* features, int64_t n, int64_t n_unlabeled, int d, int32_t budget,
                                   float* min_dist, int64_t* out_selected, void* stream_) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n > 0 && d > 0 && budget >= 0, "mval_kcenter_greedy: bad shape");
  // utils/coreset.py: with no labeled centre min_distances stays None and the reference's argmax is undefined
  MVAL_REQUIRE(n_unlabeled >= 0 && n_unlabeled < n, "mval_kcenter_greedy: need at least one labeled row (n_unlabeled < n)");
  MVAL_REQUIRE(features && min_dist && (budget == 0 || out_selected), "mval_kcenter_greedy: null pointer");
  MVAL_REQUIRE((size_t)d * 4 <= 200 * 1024, "mval_kcenter_greedy: feature dimension too large for shared memory");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  char* ws = nullptr;
  const size_t sz_norms = (sizeof(float) * n + 255) & ~size_t(255);
  MVAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), sz_norms + 256, stream));
  float* norms = reinterpret_cast<float*>(ws);
  float* best_val = reinterpret_cast<float*>(ws + sz_norms);
  int64_t* best_idx = reinterpret_cast<int64_t*>(ws + sz_norms + 64);
  auto run = [&]() -> int {
    if (int rc = mval_kcenter_norms(features, n, d, norms, stream)) return rc;
    kcenter_fill_kernel<<<num_sms() * 4, 256, 0, stream>>>(min_dist, n, INFINITY);
    MVAL_LAUNCH_CHECK("kcenter_fill");
    // coreset.py:83-84  update_distances(al_indices): one pass per labeled centre
    for (int64_t ci = n_unlabeled; ci < n; ++ci)
      if (int rc = kcenter_update(features, norms, n, d, features + ci * d, nullptr, min_dist, 0, best_val, best_idx,
                                  (budget > 0 && ci == n - 1) ? out_selected : nullptr, stream))
        return rc;
    // coreset.py:86-93  ind = argmax(min_distances); update_distances([ind]); B times
    for (int32_t t = 0; t < budget; ++t)
      if (int rc = kcenter_update(features, norms, n, d, nullptr, best_idx, min_dist, 0, best_val, best_idx,
                                  (t + 1 < budget) ? out_selected + t + 1 : nullptr, stream))
        return rc;
    return MVAL_OK;
  };
  const int rc = run();
  cudaError_t e = cudaFreeAsync(ws, stream);
  if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  return rc;
}

extern "C" size_t mval_kcenter_record_bytes(int d) { return mval::kc_record_bytes(d); }

extern "C" int mval_kcenter_update_exchange(const float* features, const float* row_norms, int64_t n, int d,
                                            const float* centre, const void* cands_in, int n_cands, float* min_dist,
                                            int64_t index_offset, void* cand_out, int64_t* out_selected, void* stream) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && d > 0, "mval_kcenter_update_exchange: bad shape");
  MVAL_REQUIRE((centre != nullptr) != (cands_in != nullptr), "mval_kcenter_update_exchange: give exactly one of centre / cands_in");
  MVAL_REQUIRE(cands_in == nullptr || n_cands > 0, "mval_kcenter_update_exchange: n_cands must be positive");
  MVAL_REQUIRE(n == 0 || (features && row_norms && min_dist), "mval_kcenter_update_exchange: null pointer");
  MVAL_REQUIRE((size_t)d * 4 <= 200 * 1024, "mval_kcenter_update_exchange: feature dimension too large for shared memory");
  KcArgs a{};
  a.feat = features; a.norms = row_norms; a.n = n; a.d = d; a.centre = centre;
  a.cands_in = static_cast<const char*>(cands_in); a.n_cands = n_cands; a.min_dist = min_dist;
  a.index_offset = index_offset; a.also_idx = out_selected; a.cand_out = static_cast<char*>(cand_out);
  return kcenter_launch(a, static_cast<cudaStream_t>(stream));
}
