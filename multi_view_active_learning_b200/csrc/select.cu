// Ranking: top-k of float64 scores, descending, ties by ascending pool index, NaN dropped
// (reference strategy.py:932-949: NaN filter + heapq.nlargest(n, dict, key=dict.get), which is a stable
// descending sort truncated to n).
//
// Implemented as a stable LSD radix sort (8 passes x 8 bits) of an order-reversing 64-bit key with the local index
// as payload; stability + initial index order give the reference's tie-break for free.  The pool's score vector is
// tiny next to its heat maps (8 B vs 2.5 MB per frame), so this is bookkeeping, not a roofline kernel.
#include "common.cuh"

namespace mval {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / kWarp;
constexpr int kSortItems = 8;                                // per lane
constexpr int kSortTile = kSortThreads * kSortItems;         // 2048 elements per block
constexpr int kWarpSeg = kWarp * kSortItems;                 // contiguous elements per warp

// ascending order of this key == descending order of the score; NaN -> all ones (sorted last); -0.0 == +0.0
__device__ __forceinline__ uint64_t desc_key(double s) {
  if (s != s) return ~0ull;
  const uint64_t u = (uint64_t)__double_as_longlong(s + 0.0);
  const uint64_t asc = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  const uint64_t k = ~asc;
  return k == ~0ull ? k - 1 : k;  // keep the all-ones pattern for NaN only (-inf would collide)
}

__global__ void __launch_bounds__(kSortThreads)
topk_make_keys_kernel(const double* __restrict__ scores, int64_t n, uint64_t* __restrict__ keys,
                      uint32_t* __restrict__ idx, int32_t* __restrict__ n_valid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    const double s = scores[i];
    keys[i] = desc_key(s);
    idx[i] = (uint32_t)i;
    ok = (s == s);
  }
  const int c = __syncthreads_count(ok);
  if (threadIdx.x == 0 && c) atomicAdd(n_valid, c);
}

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int n_blocks, uint32_t* __restrict__ counts) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int k = 0; k < kSortItems; ++k) {
    const int64_t i = base + k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  counts[(int64_t)threadIdx.x * n_blocks + blockIdx.x] = hist[threadIdx.x];
}

// exclusive scan of counts[256 * n_blocks] (digit-major) by one block
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ counts, int64_t total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < total; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = (i < total) ? counts[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(kFull, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(kFull, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    const uint32_t before = carry + (warp ? warp_sums[warp - 1] : 0u) + (x - v);
    if (i < total) counts[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
                     uint64_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out, int64_t n, int shift, int n_blocks,
                     const uint32_t* __restrict__ offsets) {
  __shared__ uint32_t whist[kSortWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < kSortWarps * 256; d += kSortThreads) (&whist[0][0])[d] = 0;
  __syncthreads();
  // element order inside the tile is (warp, item, lane): contiguous per warp, so ranks follow the input order
  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + warp * kWarpSeg;
  uint64_t key[kSortItems];
  uint32_t val[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int64_t i = wbase + k * kWarp + lane;
    key[k] = (i < n) ? keys_in[i] : ~0ull;
    val[k] = (i < n) ? idx_in[i] : 0u;
  }
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const bool in = wbase + k * kWarp + lane < n;
    const uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(kFull, in ? d : 256u + lane);
    if (in && (peers & ((1u << lane) - 1u)) == 0u) whist[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // kSortThreads == 256 digits
    uint32_t running = offsets[(int64_t)d * n_blocks + blockIdx.x];
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = whist[w][d];
      whist[w][d] = running;
      running += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const bool in = wbase + k * kWarp + lane < n;
    const uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(kFull, in ? d : 256u + lane);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t dst = 0;
    if (in) dst = whist[warp][d] + rank;
    __syncwarp();
    if (in && rank == 0) whist[warp][d] += __popc(peers);
    __syncwarp();
    if (in) {
      keys_out[dst] = key[k];
      idx_out[dst] = val[k];
    }
  }
}

__global__ void __launch_bounds__(256)
topk_gather_kernel(const double* __restrict__ scores, const uint32_t* __restrict__ sorted_idx,
                   const int32_t* __restrict__ n_valid, int32_t k, int64_t index_offset, int64_t* __restrict__ out_idx,
                   double* __restrict__ out_val, int32_t* __restrict__ out_count) {
  const int take = min(k, *n_valid);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && out_count) *out_count = take;
  if (i < take) {
    const uint32_t s = sorted_idx[i];
    out_idx[i] = index_offset + (int64_t)s;
    if (out_val) out_val[i] = scores[s];
  }
}

}  // namespace mval

extern "C" int mval_topk_desc(const double* scores, int64_t n, int64_t index_offset, int32_t k, int64_t* out_idx,
                              double* out_val, int32_t* out_count, void* stream_) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && k >= 0, "mval_topk_desc: bad sizes");
  MVAL_REQUIRE(n < (1ll << 32), "mval_topk_desc: more than 2^32 scores in one call");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0 || k == 0) {
    if (out_count) MVAL_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), stream));
    return MVAL_OK;
  }
  MVAL_REQUIRE(scores && out_idx, "mval_topk_desc: null pointer");
  const int n_blocks = (int)((n + kSortTile - 1) / kSortTile);
  const size_t sz_keys = (sizeof(uint64_t) * n + 255) & ~size_t(255);
  const size_t sz_idx = (sizeof(uint32_t) * n + 255) & ~size_t(255);
  const size_t sz_counts = (sizeof(uint32_t) * 256 * (size_t)n_blocks + 255) & ~size_t(255);
  char* ws = nullptr;
  MVAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), 2 * sz_keys + 2 * sz_idx + sz_counts + 256, stream));
  uint64_t* keys[2] = {reinterpret_cast<uint64_t*>(ws), reinterpret_cast<uint64_t*>(ws + sz_keys)};
  uint32_t* idx[2] = {reinterpret_cast<uint32_t*>(ws + 2 * sz_keys), reinterpret_cast<uint32_t*>(ws + 2 * sz_keys + sz_idx)};
  uint32_t* counts = reinterpret_cast<uint32_t*>(ws + 2 * sz_keys + 2 * sz_idx);
  int32_t* n_valid = reinterpret_cast<int32_t*>(ws + 2 * sz_keys + 2 * sz_idx + sz_counts);
  int rc = MVAL_OK;
  auto run = [&]() -> int {
    MVAL_CUDA(cudaMemsetAsync(n_valid, 0, sizeof(int32_t), stream));
    topk_make_keys_kernel<<<(unsigned)((n + kSortThreads - 1) / kSortThreads), kSortThreads, 0, stream>>>(scores, n, keys[0],
                                                                                                        idx[0], n_valid);
    MVAL_LAUNCH_CHECK("topk_make_keys");
    int cur = 0;
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = pass * 8;
      radix_hist_kernel<<<n_blocks, kSortThreads, 0, stream>>>(keys[cur], n, shift, n_blocks, counts);
      MVAL_LAUNCH_CHECK("radix_hist");
      radix_scan_kernel<<<1, 1024, 0, stream>>>(counts, 256ll * n_blocks);
      MVAL_LAUNCH_CHECK("radix_scan");
      radix_scatter_kernel<<<n_blocks, kSortThreads, 0, stream>>>(keys[cur], idx[cur], keys[cur ^ 1], idx[cur ^ 1], n, shift,
                                                                 n_blocks, counts);
      MVAL_LAUNCH_CHECK("radix_scatter");
      cur ^= 1;
    }
    const int64_t kk = k < n ? k : n;
    topk_gather_kernel<<<(unsigned)((kk + 255) / 256), 256, 0, stream>>>(scores, idx[cur], n_valid, k, index_offset, out_idx,
                                                                       out_val, out_count);
    MVAL_LAUNCH_CHECK("topk_gather");
    return MVAL_OK;
  };
  rc = run();
  cudaError_t e = cudaFreeAsync(ws, stream);
  if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  return rc;
}
