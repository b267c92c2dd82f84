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
                   const int32_t* __restrict__ n_valid, int32_t k, int64_t index_offset, const int64_t* __restrict__ index_map,
                   int64_t* __restrict__ out_idx, double* __restrict__ out_val, int32_t* __restrict__ out_count) {
  const int take = min(k, *n_valid);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && out_count) *out_count = take;
  if (i < take) {
    const uint32_t s = sorted_idx[i];
    out_idx[i] = index_map ? index_map[s] : index_offset + (int64_t)s;
    if (out_val) out_val[i] = scores[s];
  } else if (i < k) {
    out_idx[i] = -1;  // slots beyond the count are defined: -1 / NaN (fixed-size exchange buffers rely on it)
    if (out_val) out_val[i] = __longlong_as_double(0x7ff8000000000000ll);
  }
}

// Workspace of one stable LSD radix sort of n (uint64 key, uint32 payload) pairs, carved out of one stream-ordered
// allocation; sort() leaves the result in keys[cur] / idx[cur].
struct RadixSort {
  char* ws = nullptr;
  uint64_t* keys[2] = {nullptr, nullptr};
  uint32_t* idx[2] = {nullptr, nullptr};
  uint32_t* counts = nullptr;
  int32_t* scalar = nullptr;  // 64 spare bytes for the caller (counters)
  int n_blocks = 0;
  int cur = 0;

  int alloc(int64_t n, cudaStream_t stream) {
    n_blocks = (int)((n + kSortTile - 1) / kSortTile);
    const size_t sz_keys = (sizeof(uint64_t) * n + 255) & ~size_t(255);
    const size_t sz_idx = (sizeof(uint32_t) * n + 255) & ~size_t(255);
    const size_t sz_counts = (sizeof(uint32_t) * 256 * (size_t)n_blocks + 255) & ~size_t(255);
    MVAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), 2 * sz_keys + 2 * sz_idx + sz_counts + 256, stream));
    keys[0] = reinterpret_cast<uint64_t*>(ws);
    keys[1] = reinterpret_cast<uint64_t*>(ws + sz_keys);
    idx[0] = reinterpret_cast<uint32_t*>(ws + 2 * sz_keys);
    idx[1] = reinterpret_cast<uint32_t*>(ws + 2 * sz_keys + sz_idx);
    counts = reinterpret_cast<uint32_t*>(ws + 2 * sz_keys + 2 * sz_idx);
    scalar = reinterpret_cast<int32_t*>(ws + 2 * sz_keys + 2 * sz_idx + sz_counts);
    return MVAL_OK;
  }
  // keys[0] / idx[0] hold the input; first_pass..last_pass-1 are the 8-bit digits to sort on (LSD)
  int sort(int64_t n, cudaStream_t stream, int first_pass = 0, int last_pass = 8) {
    cur = 0;
    for (int pass = first_pass; pass < last_pass; ++pass) {
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
    return MVAL_OK;
  }
  int release(cudaStream_t stream, int rc) {
    if (ws == nullptr) return rc;
    cudaError_t e = cudaFreeAsync(ws, stream);
    ws = nullptr;
    if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
    return rc;
  }
};

// ---- dict-insertion semantics for the guid-keyed tables of strategy.py:1115-1133 ----------------------------------
// Row i carries the key (pose_i, frame_i) = the guid "%s-%s" % (pose, frame).  Inserting the rows in order into an
// OrderedDict keeps a key at the position of its FIRST row and with the value of its LAST row.
__global__ void __launch_bounds__(256)
guid_keys_kernel(const int64_t* __restrict__ pose, const int64_t* __restrict__ frame, int64_t n, uint64_t* __restrict__ keys,
                 uint32_t* __restrict__ idx, int32_t* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t p = (uint64_t)pose[i], f = (uint64_t)frame[i];
  if ((p >> 32) != 0ull || (f >> 32) != 0ull) *bad = 1;  // does not fit the packed 64-bit key
  keys[i] = (p << 32) | (f & 0xffffffffull);
  idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
guid_groups_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, int64_t n, uint8_t* __restrict__ keep,
                   int32_t* __restrict__ src, int32_t* __restrict__ n_unique) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool first = false;
  if (p < n) {
    const uint64_t k = keys[p];
    first = (p == 0) || keys[p - 1] != k;
    const uint32_t row = idx[p];
    if (first) {
      int64_t q = p;  // the sort is stable: rows of one key are in ascending order, the last one carries the value
      while (q + 1 < n && keys[q + 1] == k) ++q;
      keep[row] = 1;
      src[row] = (int32_t)idx[q];
    } else {
      keep[row] = 0;
      src[row] = -1;
    }
  }
  const int c = __syncthreads_count(first);
  if (threadIdx.x == 0 && c) atomicAdd(n_unique, c);
}

__global__ void guid_finish_kernel(const int32_t* __restrict__ scalar, int32_t* __restrict__ out_unique) {
  // scalar[0] = number of unique keys, scalar[1] = a key did not fit 32 + 32 bits
  *out_unique = scalar[1] ? -1 : scalar[0];
}

}  // namespace mval

static int topk_impl(const double* scores, int64_t n, int64_t index_offset, const int64_t* index_map, int32_t k,
                     int64_t* out_idx, double* out_val, int32_t* out_count, cudaStream_t stream, const char* who) {
  using namespace mval;
  MVAL_REQUIRE(n >= 0 && k >= 0, "%s: bad sizes", who);
  MVAL_REQUIRE(n < (1ll << 32), "%s: more than 2^32 scores in one call", who);
  if (n == 0 || k == 0) {
    if (out_count) MVAL_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), stream));
    if (k > 0 && out_idx) MVAL_CUDA(cudaMemsetAsync(out_idx, 0xff, sizeof(int64_t) * k, stream));
    if (k > 0 && out_val) MVAL_CUDA(cudaMemsetAsync(out_val, 0xff, sizeof(double) * k, stream));  // all ones = NaN
    return MVAL_OK;
  }
  MVAL_REQUIRE(scores && out_idx, "%s: null pointer", who);
  RadixSort rs;
  if (int rc = rs.alloc(n, stream)) return rc;
  auto run = [&]() -> int {
    MVAL_CUDA(cudaMemsetAsync(rs.scalar, 0, sizeof(int32_t), stream));
    topk_make_keys_kernel<<<(unsigned)((n + kSortThreads - 1) / kSortThreads), kSortThreads, 0, stream>>>(scores, n, rs.keys[0],
                                                                                                        rs.idx[0], rs.scalar);
    MVAL_LAUNCH_CHECK("topk_make_keys");
    if (int rc = rs.sort(n, stream)) return rc;
    topk_gather_kernel<<<(unsigned)((k + 255) / 256), 256, 0, stream>>>(scores, rs.idx[rs.cur], rs.scalar, k, index_offset,
                                                                      index_map, out_idx, out_val, out_count);
    MVAL_LAUNCH_CHECK("topk_gather");
    return MVAL_OK;
  };
  return rs.release(stream, run());
}

extern "C" int mval_topk_desc(const double* scores, int64_t n, int64_t index_offset, int32_t k, int64_t* out_idx,
                              double* out_val, int32_t* out_count, void* stream) {
  if (int rc = mval::require_device()) return rc;
  return topk_impl(scores, n, index_offset, nullptr, k, out_idx, out_val, out_count, static_cast<cudaStream_t>(stream),
                   "mval_topk_desc");
}

extern "C" int mval_topk_merge(const double* scores, const int64_t* indices, int64_t n, int32_t k, int64_t* out_idx,
                               double* out_val, int32_t* out_count, void* stream) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(n == 0 || indices != nullptr, "mval_topk_merge: null pointer");
  return topk_impl(scores, n, 0, indices, k, out_idx, out_val, out_count, static_cast<cudaStream_t>(stream), "mval_topk_merge");
}

extern "C" int mval_first_occurrence(const int64_t* pose, const int64_t* frame, int64_t n, uint8_t* out_keep, int32_t* out_src,
                                     int32_t* out_unique, void* stream_) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && n < (1ll << 31), "mval_first_occurrence: bad size");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MVAL_REQUIRE(out_unique != nullptr, "mval_first_occurrence: null pointer");
  if (n == 0) {
    MVAL_CUDA(cudaMemsetAsync(out_unique, 0, sizeof(int32_t), stream));
    return MVAL_OK;
  }
  MVAL_REQUIRE(pose && frame && out_keep && out_src, "mval_first_occurrence: null pointer");
  RadixSort rs;
  if (int rc = rs.alloc(n, stream)) return rc;
  auto run = [&]() -> int {
    MVAL_CUDA(cudaMemsetAsync(rs.scalar, 0, 2 * sizeof(int32_t), stream));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    guid_keys_kernel<<<blocks, 256, 0, stream>>>(pose, frame, n, rs.keys[0], rs.idx[0], rs.scalar + 1);
    MVAL_LAUNCH_CHECK("guid_keys");
    if (int rc = rs.sort(n, stream)) return rc;
    guid_groups_kernel<<<blocks, 256, 0, stream>>>(rs.keys[rs.cur], rs.idx[rs.cur], n, out_keep, out_src, rs.scalar);
    MVAL_LAUNCH_CHECK("guid_groups");
    guid_finish_kernel<<<1, 1, 0, stream>>>(rs.scalar, out_unique);
    MVAL_LAUNCH_CHECK("guid_finish");
    return MVAL_OK;
  };
  return rs.release(stream, run());
}
