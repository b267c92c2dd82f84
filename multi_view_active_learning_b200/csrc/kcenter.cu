// Coreset k-center greedy (reference utils/coreset.py:49-95) on float32 features, sm_100a.
//
// The reference's loop is B dependent steps  {ind = argmax(min_d); min_d = minimum(min_d, dist(X, X[ind]))} , each a
// full pass over the n x d feature matrix.  Here the same selection -- bit for bit -- is produced in ROUNDS that read
// the feature matrix once per round instead of once per step:
//
//   1. select   : radix-select the K-th largest running minimum (kappa) and compact every row above it into a
//                 candidate set (value, global index, feature row).  All other rows are <= kappa and stay <= kappa
//                 for the rest of the round because running minima only decrease.
//   2. resolve  : exact pairwise distances among the candidates (K x K, tiny), then ONE CTA replays the greedy loop on
//                 the candidates only: the arg-max of the candidates is the arg-max of the whole pool as long as its
//                 value is strictly above kappa (first pick of a round: always, the set holds the global arg-max).
//                 This yields T >= 1 picks that are exactly the next T picks of the sequential algorithm.
//   3. update   : min_d[i] = min(min_d[i], dist(x_i, c_t)) for all T new centres in one register-tiled FFMA pass
//                 (kc_batch_kernel) -- or, for large d, one tcgen05 TF32 screening GEMM that proves most (i, t)
//                 pairs cannot lower min_d[i] plus an exact re-evaluation of the few that might (kcenter_tc.cu).
//
// Across GPUs (rows sharded contiguously) step 1 runs per shard, the candidate records are all-gathered once per
// ROUND (not per step), every rank replays step 2 on the union and updates its own shard.
//
// Arithmetic (shared with oracle/coreset_oracle.c, which is the definition):
//     dot(x, c)  = fma chain over k ascending, one accumulator, starting from +0
//     d2         = ((-2 * dot) + |x|^2) + |c|^2 ;  dist = sqrt(max(d2, 0)) + 0      x = pool row, c = centre
// Any tiling that keeps one accumulator per (row, centre) and walks k in ascending order reproduces it exactly, which
// is what makes the batched / tiled kernels below legal.  Tie-break: lowest global index among equal maxima
// (np.argmax, coreset.py:90).
#include <math.h>
#include <string.h>

#include "kcenter.cuh"

namespace mval {

// ------------------------------------------------------------------------------------------------------------------
// per-device scratch
// ------------------------------------------------------------------------------------------------------------------
static KcDeviceScratch g_kc[64];

int kc_scratch(KcDeviceScratch** out) {
  int dev = 0;
  MVAL_CUDA(cudaGetDevice(&dev));
  MVAL_REQUIRE(dev < 64, "kcenter: device ordinal too large");
  KcDeviceScratch& s = g_kc[dev];
  if (s.sel == nullptr) {
    MVAL_CUDA(cudaMalloc(&s.sel, sizeof(KcSelectState)));
    MVAL_CUDA(cudaMemset(s.sel, 0, sizeof(KcSelectState)));
    MVAL_CUDA(cudaMalloc(&s.cand_idx, sizeof(uint32_t) * kKcMaxSlots));
    MVAL_CUDA(cudaMalloc(&s.partials, sizeof(KcPartial) * 2048));
    MVAL_CUDA(cudaMalloc(&s.counter, sizeof(unsigned int)));
    MVAL_CUDA(cudaMemset(s.counter, 0, sizeof(unsigned int)));
    MVAL_CUDA(cudaMallocHost(&s.host_i32, 64));
  }
  *out = &s;
  return MVAL_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// row-dot kernel: one lane per row, sequential fma chain, coalesced through a per-warp shared-memory transpose.
//   kMode 0: norms[i] = dot(x_i, x_i)
//   kMode 1: min_dist[i] = min(min_dist[i], dist(x_i, centre))            (single-centre update, HBM bound)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kRdWarps = 8;

template <int kMode>
__global__ void __launch_bounds__(kRdWarps * 32)
kc_rowdot_kernel(const float* __restrict__ X, int64_t n, int d, const float* __restrict__ centre,
                 const float* __restrict__ cc_ptr, const float* __restrict__ xx, float* __restrict__ out) {
  __shared__ float tile[kRdWarps][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float(*t)[33] = tile[warp];
  const float cc = (kMode == 1) ? __ldg(cc_ptr) : 0.0f;
  const int64_t n_groups = (n + 31) / 32;
  for (int64_t g = (int64_t)blockIdx.x * kRdWarps + warp; g < n_groups; g += (int64_t)gridDim.x * kRdWarps) {
    const int64_t row0 = g * 32;
    const int rows = (int)((n - row0) < 32 ? (n - row0) : 32);
    float acc = 0.0f;
    for (int k0 = 0; k0 < d; k0 += 32) {
      const int k = k0 + lane;
      float v[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) v[r] = (r < rows && k < d) ? __ldg(X + (row0 + r) * d + k) : 0.0f;
      float cv = 0.0f;
      if (kMode == 1) cv = (k < d) ? __ldg(centre + k) : 0.0f;
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; ++r) t[r][lane] = v[r];
      __syncwarp();
      const int kk = (d - k0) < 32 ? (d - k0) : 32;
      if (kk == 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = t[lane][j];
          const float c = (kMode == 1) ? __shfl_sync(kFull, cv, j) : x;
          acc = __fmaf_rn(x, c, acc);
        }
      } else {
        for (int j = 0; j < kk; ++j) {
          const float x = t[lane][j];
          const float c = (kMode == 1) ? __shfl_sync(kFull, cv, j) : x;
          acc = __fmaf_rn(x, c, acc);
        }
      }
    }
    if (lane < rows) {
      const int64_t i = row0 + lane;
      if (kMode == 0) {
        out[i] = acc;
      } else {
        const float dist = kc_dist(acc, xx[i], cc);
        if (dist < out[i]) out[i] = dist;
      }
    }
  }
}

// norm of one row (the centre of a single-centre update), one warp, canonical order
__global__ void kc_one_norm_kernel(const float* __restrict__ c, int d, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    float acc = 0.0f;
    for (int k = 0; k < d; ++k) acc = __fmaf_rn(c[k], c[k], acc);
    *out = acc;
  }
}

static int rowdot_grid(int64_t n) {
  const int64_t want = ((n + 31) / 32 + kRdWarps - 1) / kRdWarps;
  const int64_t cap = (int64_t)num_sms() * 6;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// ------------------------------------------------------------------------------------------------------------------
// batched exact kernel: register-tiled FFMA "SGEMM" with one accumulator per (row, centre), k ascending.
//   CTA tile 128 rows x (16*TN) centres, BK = 16, 256 threads, thread tile 8 x TN.
//   kStore = false: min_dist[i] = min(min_dist[i], min_t dist(x_i, c_t))          (atomicMin on the float bits)
//   kStore = true : out[t * ld_out + i] = dist(x_i, c_t)                          (candidate pairwise matrix)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBM = 128, kBK = 16, kBatchThreads = 256;

template <int TN>
__device__ __forceinline__ int kc_col_of(int tx, int c) {
  return TN == 8 ? ((c >> 2) * 64 + tx * 4 + (c & 3)) : (TN == 4 ? tx * 4 + c : tx);
}

template <int TN, bool kStore, bool kVec>
__global__ void __launch_bounds__(kBatchThreads, 2)
kc_batch_kernel(const float* __restrict__ X, const float* __restrict__ xx, int64_t n, int d, const float* __restrict__ C,
                const float* __restrict__ cc, int T, int n_col_tiles, float* __restrict__ min_dist, float* __restrict__ out,
                int64_t ld_out, const unsigned int* __restrict__ gate, unsigned int gate_capacity, KcCount cnt) {
  constexpr int BN = 16 * TN;
  if (gate != nullptr && *gate <= gate_capacity) return;  // fallback pass of the tensor-core path: nothing overflowed
  T = kc_effective_T(T, cnt);
  if ((int)(blockIdx.x % n_col_tiles) * BN >= T) return;  // (device-side batch size: nothing to do for this column tile)
  constexpr int XS = kBM + 4, CS = BN + 4;
  __shared__ __align__(16) float xs[2][kBK][XS];
  __shared__ __align__(16) float cs[2][kBK][CS];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row_tile = blockIdx.x / n_col_tiles;
  const int col_tile = (int)(blockIdx.x % n_col_tiles);
  const int64_t row0 = row_tile * kBM;
  const int col0 = col_tile * BN;

  float acc[8][TN];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = 0.0f;

  // ---- tile loaders (global -> registers -> transposed shared)
  constexpr int kXv = kVec ? 2 : 8;                       // float4 / scalar loads of X per thread per tile
  constexpr int kCv = kVec ? (BN * kBK / 4 + 255) / 256   // float4 loads of C per thread per tile
                           : (BN * kBK + 255) / 256;
  float4 xr4[kVec ? 2 : 1];
  float xr1[kVec ? 1 : 8];
  float4 cr4[kVec ? kCv : 1];
  float cr1[kVec ? 1 : kCv];

  auto load_tiles = [&](int k0) {
    if constexpr (kVec) {
#pragma unroll
      for (int j = 0; j < kXv; ++j) {
        const int lr = (tid >> 2) + 64 * j, kq = (tid & 3) * 4;
        const int64_t row = row0 + lr;
        xr4[j] = (row < n && k0 + kq < d) ? *reinterpret_cast<const float4*>(X + row * d + k0 + kq)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < kCv; ++j) {
        const int e = tid + 256 * j;  // float4 index inside the BN x 4 tile of float4s
        const int lc = e >> 2, kq = (e & 3) * 4;
        const int col = col0 + lc;
        cr4[j] = (lc < BN && col < T && k0 + kq < d) ? *reinterpret_cast<const float4*>(C + (int64_t)col * d + k0 + kq)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kXv; ++j) {
        const int lr = (tid >> 4) + 16 * j, k = tid & 15;
        const int64_t row = row0 + lr;
        xr1[j] = (row < n && k0 + k < d) ? __ldg(X + row * d + k0 + k) : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < kCv; ++j) {
        const int e = tid + 256 * j;
        const int lc = e >> 4, k = e & 15;
        const int col = col0 + lc;
        cr1[j] = (lc < BN && col < T && k0 + k < d) ? __ldg(C + (int64_t)col * d + k0 + k) : 0.0f;
      }
    }
  };
  auto store_tiles = [&](int buf) {
    if constexpr (kVec) {
#pragma unroll
      for (int j = 0; j < kXv; ++j) {
        const int lr = (tid >> 2) + 64 * j, kq = (tid & 3) * 4;
        xs[buf][kq + 0][lr] = xr4[j].x;
        xs[buf][kq + 1][lr] = xr4[j].y;
        xs[buf][kq + 2][lr] = xr4[j].z;
        xs[buf][kq + 3][lr] = xr4[j].w;
      }
#pragma unroll
      for (int j = 0; j < kCv; ++j) {
        const int e = tid + 256 * j;
        const int lc = e >> 2, kq = (e & 3) * 4;
        if (lc < BN) {
          cs[buf][kq + 0][lc] = cr4[j].x;
          cs[buf][kq + 1][lc] = cr4[j].y;
          cs[buf][kq + 2][lc] = cr4[j].z;
          cs[buf][kq + 3][lc] = cr4[j].w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < kXv; ++j) xs[buf][tid & 15][(tid >> 4) + 16 * j] = xr1[j];
#pragma unroll
      for (int j = 0; j < kCv; ++j) {
        const int e = tid + 256 * j;
        if ((e >> 4) < BN) cs[buf][e & 15][e >> 4] = cr1[j];
      }
    }
  };

  const int nk = (d + kBK - 1) / kBK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * kBK);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 xa = *reinterpret_cast<const float4*>(&xs[cur][k][ty * 4]);
      const float4 xb = *reinterpret_cast<const float4*>(&xs[cur][k][64 + ty * 4]);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      float cv[TN];
      if constexpr (TN == 8) {
        const float4 ca = *reinterpret_cast<const float4*>(&cs[cur][k][tx * 4]);
        const float4 cb = *reinterpret_cast<const float4*>(&cs[cur][k][64 + tx * 4]);
        cv[0] = ca.x; cv[1] = ca.y; cv[2] = ca.z; cv[3] = ca.w;
        cv[4] = cb.x; cv[5] = cb.y; cv[6] = cb.z; cv[7] = cb.w;
      } else if constexpr (TN == 4) {
        const float4 ca = *reinterpret_cast<const float4*>(&cs[cur][k][tx * 4]);
        cv[0] = ca.x; cv[1] = ca.y; cv[2] = ca.z; cv[3] = ca.w;
      } else {
        cv[0] = cs[cur][k][tx];
      }
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = __fmaf_rn(xv[r], cv[c], acc[r][c]);
    }
    if (kt + 1 < nk) store_tiles(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  float ccv[TN];
  bool cok[TN];
#pragma unroll
  for (int c = 0; c < TN; ++c) {
    const int col = col0 + kc_col_of<TN>(tx, c);
    cok[c] = col < T;
    ccv[c] = cok[c] ? __ldg(cc + col) : 0.0f;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int lr = (r >> 2) * 64 + ty * 4 + (r & 3);
    const int64_t row = row0 + lr;
    const bool rok = row < n;
    const float xr = rok ? __ldg(xx + row) : 0.0f;
    float best = INFINITY;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const float dist = kc_dist(acc[r][c], xr, ccv[c]);
      if (kStore) {
        if (rok && cok[c]) out[(int64_t)(col0 + kc_col_of<TN>(tx, c)) * ld_out + row] = dist;
      } else {
        best = cok[c] ? fminf(best, dist) : best;
      }
    }
    if (!kStore) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(kFull, best, o));
      if (tx == 0 && rok && best < min_dist[row])
        atomicMin(reinterpret_cast<unsigned int*>(min_dist + row), __float_as_uint(best));
    }
  }
}

template <int TN, bool kStore>
static int launch_batch_tn(const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T,
                           float* min_dist, float* out, int64_t ld_out, cudaStream_t stream,
                           const unsigned int* gate = nullptr, unsigned int gate_capacity = 0, KcCount cnt = KcCount{nullptr, 0}) {
  constexpr int BN = 16 * TN;
  const int n_col_tiles = (T + BN - 1) / BN;
  const int64_t n_row_tiles = (n + kBM - 1) / kBM;
  const int64_t blocks = n_row_tiles * n_col_tiles;
  MVAL_REQUIRE(blocks < (int64_t)INT32_MAX, "kcenter batch: problem too large for one launch");
  const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  if (vec)
    kc_batch_kernel<TN, kStore, true><<<(unsigned)blocks, kBatchThreads, 0, stream>>>(X, xx, n, d, C, cc, T, n_col_tiles,
                                                                                        min_dist, out, ld_out, gate, gate_capacity, cnt);
  else
    kc_batch_kernel<TN, kStore, false><<<(unsigned)blocks, kBatchThreads, 0, stream>>>(X, xx, n, d, C, cc, T, n_col_tiles,
                                                                                         min_dist, out, ld_out, gate, gate_capacity, cnt);
  MVAL_LAUNCH_CHECK("kc_batch");
  return MVAL_OK;
}

int kc_update_batch_exact(const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T,
                          float* min_dist, cudaStream_t stream, KcCount cnt) {
  if (n == 0 || T == 0) return MVAL_OK;
  if (cnt.ptr != nullptr) {  // batch size known to the device only: the tiled kernel clamps it
    if (T <= 64) return launch_batch_tn<4, false>(X, xx, n, d, C, cc, T, min_dist, nullptr, 0, stream, nullptr, 0, cnt);
    return launch_batch_tn<8, false>(X, xx, n, d, C, cc, T, min_dist, nullptr, 0, stream, nullptr, 0, cnt);
  }
  if (T == 1) {  // one centre: HBM-bound row-dot kernel
    kc_rowdot_kernel<1><<<rowdot_grid(n), kRdWarps * 32, 0, stream>>>(X, n, d, C, cc, xx, min_dist);
    MVAL_LAUNCH_CHECK("kc_rowdot_update");
    return MVAL_OK;
  }
  if (T <= 16) return launch_batch_tn<1, false>(X, xx, n, d, C, cc, T, min_dist, nullptr, 0, stream);
  if (T <= 64) return launch_batch_tn<4, false>(X, xx, n, d, C, cc, T, min_dist, nullptr, 0, stream);
  return launch_batch_tn<8, false>(X, xx, n, d, C, cc, T, min_dist, nullptr, 0, stream);
}

// The same pass, executed only if *count > capacity on the device (overflow of the tensor-core path's pair list).
int kc_update_batch_exact_if(const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T,
                             float* min_dist, const unsigned int* count, unsigned int capacity, cudaStream_t stream, KcCount cnt) {
  if (T <= 64) return launch_batch_tn<4, false>(X, xx, n, d, C, cc, T, min_dist, nullptr, 0, stream, count, capacity, cnt);
  return launch_batch_tn<8, false>(X, xx, n, d, C, cc, T, min_dist, nullptr, 0, stream, count, capacity, cnt);
}

int kc_pairwise_exact(const float* X, const float* xx, int n, int d, float* out_t, cudaStream_t stream) {
  if (n <= 64) return launch_batch_tn<4, true>(X, xx, n, d, X, xx, n, nullptr, out_t, n, stream);
  return launch_batch_tn<8, true>(X, xx, n, d, X, xx, n, nullptr, out_t, n, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// select: K-th largest running minimum by 3-pass radix select (12 + 10 + 10 bits of the float bit pattern; running
// minima are >= +0 or +inf, so the unsigned bit pattern is monotone), then compaction of everything above it.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 256;

// Suffix search over a histogram: finds the largest bin b with count(bins >= b) >= want; returns b and writes the
// number still wanted inside bin b (want - count(bins > b)).  If the total is < want returns -1.  Whole CTA calls.
template <int kBins>
__device__ int kc_find_bin(const uint32_t* __restrict__ hist, uint32_t want, uint32_t* rem_out, uint32_t* sh /*[kSelThreads+2]*/) {
  constexpr int kPer = kBins / kSelThreads;
  static_assert(kBins % kSelThreads == 0, "bins must be a multiple of the block size");
  const int tid = threadIdx.x;
  // thread `tid` owns bins [hi - kPer + 1, hi] counted from the TOP: chunk index tid covers the tid-th highest chunk
  const int top = kBins - 1 - tid * kPer;
  uint32_t local = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) local += hist[top - j];
  sh[tid] = local;
  __syncthreads();
  if (tid == 0) {  // 256 sequential adds: negligible
    uint32_t run = 0;
    int chunk = -1;
    uint32_t before = 0;
    for (int c = 0; c < kSelThreads; ++c) {
      if (run + sh[c] >= want) { chunk = c; before = run; break; }
      run += sh[c];
    }
    sh[kSelThreads] = (uint32_t)chunk;
    sh[kSelThreads + 1] = before;
  }
  __syncthreads();
  const int chunk = (int)sh[kSelThreads];
  if (chunk < 0) return -1;
  uint32_t run = sh[kSelThreads + 1];
  const int ctop = kBins - 1 - chunk * kPer;
  int bin = ctop;
  for (int j = 0; j < kPer; ++j) {
    const uint32_t h = hist[ctop - j];
    if (run + h >= want) { bin = ctop - j; break; }
    run += h;
  }
  *rem_out = want - run;
  __syncthreads();
  return bin;
}

// pass 0: histogram of bits 31..20
__global__ void __launch_bounds__(kSelThreads) kc_sel_hist1_kernel(const float* __restrict__ m, int64_t n, KcSelectState* st) {
  __shared__ uint32_t h[4096];
  for (int i = threadIdx.x; i < 4096; i += kSelThreads) h[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSelThreads)
    atomicAdd(&h[__float_as_uint(m[i]) >> 20], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += kSelThreads)
    if (h[i]) atomicAdd(&st->hist1[i], h[i]);
}

// pass 1: find b1, histogram bits 19..10 of the elements inside it
__global__ void __launch_bounds__(kSelThreads) kc_sel_hist2_kernel(const float* __restrict__ m, int64_t n, uint32_t K, KcSelectState* st) {
  __shared__ uint32_t h[1024];
  __shared__ uint32_t sh[kSelThreads + 2];
  uint32_t rem = 0;
  const int b1 = kc_find_bin<4096>(st->hist1, K, &rem, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->b1 = b1; st->k1 = rem; }
  if (b1 < 0) return;  // fewer than K rows: everything is a candidate
  for (int i = threadIdx.x; i < 1024; i += kSelThreads) h[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSelThreads) {
    const uint32_t u = __float_as_uint(m[i]);
    if ((int)(u >> 20) == b1) atomicAdd(&h[(u >> 10) & 1023u], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += kSelThreads)
    if (h[i]) atomicAdd(&st->hist2[i], h[i]);
}

// pass 2: find b2, histogram bits 9..0
__global__ void __launch_bounds__(kSelThreads) kc_sel_hist3_kernel(const float* __restrict__ m, int64_t n, KcSelectState* st) {
  __shared__ uint32_t h[1024];
  __shared__ uint32_t sh[kSelThreads + 2];
  const int b1 = st->b1;
  if (b1 < 0) return;
  uint32_t rem = 0;
  const int b2 = kc_find_bin<1024>(st->hist2, st->k1, &rem, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->b2 = b2; st->k2 = rem; }
  const uint32_t prefix = ((uint32_t)b1 << 10) | (uint32_t)b2;
  for (int i = threadIdx.x; i < 1024; i += kSelThreads) h[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSelThreads) {
    const uint32_t u = __float_as_uint(m[i]);
    if ((u >> 10) == prefix) atomicAdd(&h[u & 1023u], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += kSelThreads)
    if (h[i]) atomicAdd(&st->hist3[i], h[i]);
}

// pass 3: find kappa, compact rows with key > kappa, remember the first row with key == kappa
__global__ void __launch_bounds__(kSelThreads)
kc_sel_compact_kernel(const float* __restrict__ m, int64_t n, KcSelectState* st, uint32_t* __restrict__ cand_idx, uint32_t K) {
  __shared__ uint32_t sh[kSelThreads + 2];
  const int b1 = st->b1;
  uint32_t kappa = 0;
  bool all = false;
  if (b1 < 0) {
    all = true;
  } else {
    uint32_t rem = 0;
    const int b3 = kc_find_bin<1024>(st->hist3, st->k2, &rem, sh);
    kappa = ((uint32_t)b1 << 20) | ((uint32_t)st->b2 << 10) | (uint32_t)b3;
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->kappa = kappa; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) st->all = all ? 1u : 0u;
  for (int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSelThreads) {
    const uint32_t u = __float_as_uint(m[i]);
    if (all || u > kappa) {
      const uint32_t slot = atomicAdd(&st->n_cand, 1u);
      if (slot < K) cand_idx[slot] = (uint32_t)i;
    } else if (u == kappa) {
      atomicMin(&st->first_eq, (uint32_t)i);
    }
  }
}

// Writes the record block of this shard (layout: kcenter.cuh) and resets the select state for the next round.
__global__ void __launch_bounds__(128)
kc_sel_records_kernel(const float* __restrict__ X, const float* __restrict__ xx, const float* __restrict__ m, int64_t n, int d,
                      int64_t index_offset, int K, KcSelectState* st, const uint32_t* __restrict__ cand_idx, char* __restrict__ rec) {
  const int slot = blockIdx.x;
  uint32_t n_cand = st->n_cand;
  const bool all = st->all != 0u;
  uint32_t first_eq = st->first_eq;
  const uint32_t kappa = st->kappa;
  // n_cand < K by construction unless all (n < K) -> also < K... keep a clamp for safety
  if (n_cand > (uint32_t)K) n_cand = (uint32_t)K;
  int64_t idx = -1;
  if ((uint32_t)slot < n_cand) idx = cand_idx[slot];
  else if (n_cand == 0 && slot == 0 && !all && first_eq != 0xffffffffu) idx = first_eq;
  KcRecordView v = kc_record_view(rec, K, d);
  float* row = v.rows + (int64_t)slot * d;
  if (idx >= 0) {
    const float* src = X + idx * d;
    for (int k = threadIdx.x; k < d; k += blockDim.x) row[k] = src[k];
  } else {
    for (int k = threadIdx.x; k < d; k += blockDim.x) row[k] = 0.0f;
  }
  if (threadIdx.x == 0) {
    v.val[slot] = idx >= 0 ? m[idx] : -1.0f;
    v.xx[slot] = idx >= 0 ? xx[idx] : 0.0f;
    v.gidx[slot] = idx >= 0 ? idx + index_offset : INT64_MAX;
    if (slot == 0) {
      v.head->count = (int32_t)((n_cand == 0 && idx >= 0) ? 1u : n_cand);
      v.head->tau = all ? -1.0f : __uint_as_float(kappa);
    }
  }
}

__global__ void kc_sel_reset_kernel(KcSelectState* st) {
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) st->hist1[i] = 0;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) { st->hist2[i] = 0; st->hist3[i] = 0; }
  if (threadIdx.x == 0) {
    st->n_cand = 0; st->first_eq = 0xffffffffu; st->kappa = 0; st->all = 0; st->b1 = -1; st->b2 = 0; st->k1 = 0; st->k2 = 0;
  }
}

int kc_select(const float* X, const float* xx, const float* m, int64_t n, int d, int64_t index_offset, int K, void* records,
              cudaStream_t stream) {
  MVAL_REQUIRE(K >= 4 && K <= kKcMaxSlots && K % 4 == 0, "kcenter select: k_slots must be a multiple of 4 in [4, %d]", kKcMaxSlots);
  MVAL_REQUIRE(n < (int64_t)0xffffffffll, "kcenter select: shard too large (rows must fit 32 bits)");
  KcDeviceScratch* s = nullptr;
  if (int rc = kc_scratch(&s)) return rc;
  kc_sel_reset_kernel<<<1, 256, 0, stream>>>(s->sel);
  MVAL_LAUNCH_CHECK("kc_sel_reset");
  if (n > 0) {
    const int64_t want = (n + kSelThreads * 8 - 1) / (kSelThreads * 8);
    const int grid = (int)(want < (int64_t)num_sms() * 4 ? want : (int64_t)num_sms() * 4);
    kc_sel_hist1_kernel<<<grid, kSelThreads, 0, stream>>>(m, n, s->sel);
    MVAL_LAUNCH_CHECK("kc_sel_hist1");
    kc_sel_hist2_kernel<<<grid, kSelThreads, 0, stream>>>(m, n, (uint32_t)K, s->sel);
    MVAL_LAUNCH_CHECK("kc_sel_hist2");
    kc_sel_hist3_kernel<<<grid, kSelThreads, 0, stream>>>(m, n, s->sel);
    MVAL_LAUNCH_CHECK("kc_sel_hist3");
    kc_sel_compact_kernel<<<grid, kSelThreads, 0, stream>>>(m, n, s->sel, s->cand_idx, (uint32_t)K);
    MVAL_LAUNCH_CHECK("kc_sel_compact");
  }
  kc_sel_records_kernel<<<K, 128, 0, stream>>>(X, xx, m, n, d, index_offset, K, s->sel, s->cand_idx, static_cast<char*>(records));
  MVAL_LAUNCH_CHECK("kc_sel_records");
  return MVAL_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// resolve: unpack the gathered record blocks, pairwise distances, greedy replay on the candidates
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
kc_unpack_kernel(const char* __restrict__ recs, int n_blocks, int K, int d, size_t block_bytes, float* __restrict__ rows,
                 float* __restrict__ val, float* __restrict__ xx, int64_t* __restrict__ gidx, float* __restrict__ tau_out) {
  const int u = blockIdx.x;  // union slot
  const int b = u / K, slot = u % K;
  KcRecordView v = kc_record_view(const_cast<char*>(recs) + (size_t)b * block_bytes, K, d);
  const float* src = v.rows + (int64_t)slot * d;
  float* dst = rows + (int64_t)u * d;
  for (int k = threadIdx.x; k < d; k += blockDim.x) dst[k] = src[k];
  if (threadIdx.x == 0) {
    val[u] = v.val[slot];
    xx[u] = v.xx[slot];
    gidx[u] = v.gidx[slot];
    if (u == 0) {
      float tau = -1.0f;
      for (int i = 0; i < n_blocks; ++i) {
        KcRecordView w = kc_record_view(const_cast<char*>(recs) + (size_t)i * block_bytes, K, d);
        tau = fmaxf(tau, w.head->tau);
      }
      *tau_out = tau;
    }
  }
}

// One CTA replays the greedy loop on the Kc <= 1024 candidates.  dt[s * Kc + j] = dist(candidate j, centre = candidate s).
//
// A pick is a block arg-max followed, for every candidate, by ONE element of the winner's row of dt.  Two things made the
// round-1 kernel (1024 threads, one candidate each) cost 0.65-0.8 us per pick: every one of its 32 warps issued the whole
// bookkeeping (~3 000 warp instructions per pick on one SM), and the row element is a dependent L2 access.  Here
//   * 16 warps own 2 candidates per thread (candidate j = k * 512 + thread, coalesced rows), one barrier per pick (the
//     per-warp partials are double-buffered).  Measured per pick on one 125k x 2048 shard (~1000 picks per launch,
//     profiles/r2_summary.md): 1024 threads x 1 candidate (round 1) 0.82 us -- all 32 warps repeat the bookkeeping, 3 300
//     warp instructions per pick; 512 x 2 (this kernel) 0.72 us; 128 x 8: 1.1 us -- one warp per scheduler, every latency
//     exposed; ONE warp with 32 candidates per lane in registers and no barrier at all: 2.9 us with 4-byte cp.async staging
//     (1024 LSU passes per pick), 1.95 us with 16-byte staging -- a single warp's dependent chain (32 compare-selects, two
//     REDUX pairs, three dependent shared loads) is slower than 16 warps sharing the work across a barrier;
//   * rows are staged ahead of time in shared memory: the kReplayStage rows of the candidates with the highest initial
//     values are resident (running minima only decrease, so candidates are picked roughly in that order) and every
//     thread copies exactly the elements it will read itself (cp.async, 4 B each), so staged data needs no block-level
//     synchronisation.  A winner whose row is not (yet) there is read from global memory: the result never depends on
//     the staging.
constexpr int kReplayThreads = 512;
constexpr int kReplayPer = kKcMaxSlots / kReplayThreads;  // 2 candidates per thread
constexpr int kReplayStage = 40;   // staged rows (40 x 4 KiB at Kc = 1024)
constexpr int kReplayDepth = 8;    // cp.async groups in flight: a staged row is usable kReplayDepth picks after its issue

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kN>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory"); }

// Ranks and the sorted order of the Kc candidates, computed by a small grid instead of inside the replay CTA:
//   rank[j]          position of slot j among all slots by (global index, slot): the tie-break of the arg-max
//   order[p]         the slot at position p by (value desc, rank asc): the order in which candidates are picked as long as
//                    picks do not lower other candidates; also the staging order of the replay
//   slot_of_rank[r]  inverse of rank
constexpr int kOrderThreads = 128;
__global__ void __launch_bounds__(kOrderThreads)
kc_order_kernel(const float* __restrict__ val, const int64_t* __restrict__ gidx, int Kc, int32_t* __restrict__ rank,
                int32_t* __restrict__ order, int32_t* __restrict__ slot_of_rank, int32_t* __restrict__ prefix_m) {
  __shared__ int64_t s_g[kKcMaxSlots];
  __shared__ float s_v[kKcMaxSlots];
  for (int i = threadIdx.x; i < Kc; i += kOrderThreads) {
    s_g[i] = gidx[i];
    s_v[i] = val[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *prefix_m = Kc;
  __syncthreads();
  const int j = blockIdx.x * kOrderThreads + threadIdx.x;
  if (j >= Kc) return;
  const int64_t g = s_g[j];
  const float v = s_v[j];
  int r = 0, ord = 0;
  for (int i = 0; i < Kc; ++i) {
    const int64_t gi = s_g[i];
    const bool before = gi < g || (gi == g && i < j);
    r += before ? 1 : 0;
    const float vi = s_v[i];
    ord += (vi > v || (vi == v && before)) ? 1 : 0;
  }
  rank[j] = r;
  slot_of_rank[r] = j;
  order[ord] = j;
}

// Speculative look-ahead over a whole round.  As long as no pick lowers a later candidate, the greedy loop simply walks the
// sorted order: the first m picks are order[0..m) with m the first position p such that
//   some earlier candidate is closer to candidate p than p's own running minimum (dt[order[i]][order[p]] < val[order[p]], i < p),
//   or p may not be picked at all (value not above tau, or not positive: a zero value means "picked before", see the replay).
// That is Kc^2 / 2 independent comparisons -- one warp per position, a few microseconds on the whole GPU -- instead of m
// dependent trips through the replay CTA (0.7 us each).  On near-orthogonal high-dimensional features (C4) m is the whole
// round; on clustered pools it is short and the replay CTA takes over from pick m with the exact state.
__global__ void __launch_bounds__(256)
kc_prefix_kernel(const float* __restrict__ val, const float* __restrict__ dt, const int32_t* __restrict__ order, int Kc,
                 const float* __restrict__ tau_ptr, int32_t* __restrict__ prefix_m) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p < 1 || p >= Kc) return;
  const int sp = order[p];
  const float vp = val[sp];
  bool stop = !(vp > 0.0f) || !(vp > *tau_ptr);
  if (!stop) {
    bool lowered = false;
    for (int i = lane; i < p; i += 32) lowered |= dt[(int64_t)order[i] * Kc + sp] < vp;
    stop = __any_sync(kFull, lowered);
  }
  if (stop && lane == 0) atomicMin(prefix_m, p);
}

__global__ void __launch_bounds__(kReplayThreads)
kc_replay_kernel(const float* __restrict__ val, const int64_t* __restrict__ gidx, const float* __restrict__ dt, int Kc,
                 const float* __restrict__ tau_ptr, int max_picks, const int32_t* __restrict__ g_rank,
                 const int32_t* __restrict__ g_order, const int32_t* __restrict__ g_slot_of_rank,
                 const int32_t* __restrict__ prefix_m, int64_t* __restrict__ selected_out, int32_t* __restrict__ pick_slots,
                 int32_t* __restrict__ n_picks_out, int32_t* __restrict__ state) {
  extern __shared__ float stage[];  // [kReplayStage][Kc]
  __shared__ uint32_t w_val[2][kReplayThreads / 32];
  __shared__ uint32_t w_rank[2][kReplayThreads / 32];
  __shared__ int64_t s_g[kKcMaxSlots];
  __shared__ int16_t s_slot_of_rank[kKcMaxSlots];
  __shared__ int16_t s_order[kKcMaxSlots];       // candidates by (initial value desc, rank asc): the staging order
  __shared__ int16_t s_stage_slot[kKcMaxSlots];  // candidate -> stage slot holding its row, -1 = none
  __shared__ int32_t s_stage_time[kKcMaxSlots];  // pick counter at which that row was issued
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float tau = *tau_ptr;
  if (state != nullptr) {  // device-side bookkeeping of the rounds: [0] picks so far, [2] budget
    const int done = state[0], left = state[2] - done;
    max_picks = left < max_picks ? (left < 0 ? 0 : left) : max_picks;
    selected_out += done;
  }
  // the look-ahead's picks: order[0 .. m)
  int m = *prefix_m;
  m = m < 1 ? 1 : m;  // the first pick of a round is always valid (the set holds the global arg-max)
  m = m < max_picks ? m : max_picks;
  float v[kReplayPer];
  uint32_t rank[kReplayPer];
#pragma unroll
  for (int k = 0; k < kReplayPer; ++k) {
    const int j = k * kReplayThreads + tid;
    v[k] = (j < Kc) ? val[j] : -1.0f;
    rank[k] = (j < Kc) ? (uint32_t)g_rank[j] : 0xffffffffu;
    s_g[j] = (j < Kc) ? gidx[j] : INT64_MAX;
    s_slot_of_rank[j] = (j < Kc) ? (int16_t)g_slot_of_rank[j] : (int16_t)0;
    s_order[j] = (j < Kc) ? (int16_t)g_order[j] : (int16_t)0;
    s_stage_slot[j] = -1;
    s_stage_time[j] = 0;
  }
  __syncthreads();
  if (!(val[s_order[0]] >= 0.0f)) m = 0;  // no valid candidate at all (uniform: every thread reads the same element)
  for (int t = tid; t < m; t += kReplayThreads) {
    const int sl = s_order[t];
    selected_out[t] = s_g[sl];
    pick_slots[t] = sl;
  }
  // the state after those m picks: every candidate folded against the m winners' rows (coalesced, independent loads)
  for (int i = 0; i < m; ++i) {
    const float* grow = dt + (int64_t)s_order[i] * Kc;
#pragma unroll
    for (int k = 0; k < kReplayPer; ++k) {
      const int j = k * kReplayThreads + tid;
      if (j < Kc) v[k] = fminf(v[k], __ldg(grow + j));
    }
  }
  // stage the rows of the next candidates in the sorted order
  const int n_stage = (Kc - m) < kReplayStage ? (Kc - m) : kReplayStage;
  for (int r = 0; r < n_stage; ++r) {
    const int c = s_order[m + r];
#pragma unroll
    for (int k = 0; k < kReplayPer; ++k) {
      const int j = k * kReplayThreads + tid;
      if (j < Kc) cp_async4(&stage[r * Kc + j], dt + (int64_t)c * Kc + j);
    }
    if (tid == 0) {
      s_stage_slot[c] = (int16_t)r;
      s_stage_time[c] = m - kReplayDepth;
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  int next_stage = m + n_stage;
  int t = m;
  for (; t < max_picks; ++t) {
    // arg-max of (value desc, rank asc): own candidates, warp, block.  Invalid slots (v < 0) carry key 0 and never beat a
    // valid one.
    uint32_t bk = 0u, br = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < kReplayPer; ++k) {
      const uint32_t key = (v[k] >= 0.0f) ? (__float_as_uint(v[k]) + 1u) : 0u;
      const bool better = key > bk || (key == bk && rank[k] < br);
      bk = better ? key : bk;
      br = better ? rank[k] : br;
    }
    const uint32_t wmax = __reduce_max_sync(kFull, bk);
    const uint32_t wrank = __reduce_min_sync(kFull, bk == wmax ? br : 0xffffffffu);
    const int buf = t & 1;
    if (lane == 0) { w_val[buf][warp] = wmax; w_rank[buf][warp] = wrank; }
    __syncthreads();
    const uint32_t k2 = (lane < kReplayThreads / 32) ? w_val[buf][lane] : 0u;
    const uint32_t r2 = (lane < kReplayThreads / 32) ? w_rank[buf][lane] : 0xffffffffu;
    const uint32_t bmax = __reduce_max_sync(kFull, k2);
    const uint32_t brank = __reduce_min_sync(kFull, k2 == bmax ? r2 : 0xffffffffu);
    if (bmax == 0u) break;  // no valid candidate at all
    const float best_v = __uint_as_float(bmax - 1u);
    if (t > 0 && !(best_v > tau)) break;
    const int s = s_slot_of_rank[brank];
    if (tid == 0) {
      selected_out[t] = s_g[s];
      pick_slots[t] = s;
    }
    // A candidate that was picked before has value 0 (its distance to itself), so it can only win again when EVERY value
    // is 0 (degenerate pools; the reference then repeats an index as well) -- and then no row can lower anything.  Only a
    // first-time winner (best_v > 0) reads its row; the staging entries of picked candidates are never looked at again.
    const bool fetch = best_v > 0.0f;
    const int slot = fetch ? (int)s_stage_slot[s] : -1;
    const bool ready = slot >= 0 && (t - s_stage_time[s]) >= kReplayDepth;
    if (fetch) {
      const float* row = ready ? (stage + slot * Kc) : nullptr;
      const float* grow = dt + (int64_t)s * Kc;
#pragma unroll
      for (int k = 0; k < kReplayPer; ++k) {
        const int j = k * kReplayThreads + tid;
        if (j < Kc) v[k] = fminf(v[k], ready ? row[j] : __ldg(grow + j));
      }
    }
    // the winner's row is no longer needed: its slot takes the next row of the staging order.  (A slot whose copy is
    // still in flight is not reused -- the new copy could be overtaken by the old one -- it simply stays unused.)
    if (ready && next_stage < Kc) {
      const int c = s_order[next_stage];
#pragma unroll
      for (int k = 0; k < kReplayPer; ++k) {
        const int j = k * kReplayThreads + tid;
        if (j < Kc) cp_async4(&stage[slot * Kc + j], dt + (int64_t)c * Kc + j);
      }
      if (tid == 0) {
        s_stage_slot[c] = (int16_t)slot;
        s_stage_time[c] = t;
      }
    }
    if (ready) ++next_stage;
    cp_async_commit();
    cp_async_wait<kReplayDepth - 1>();
  }
  cp_async_wait<0>();
  if (tid == 0) {
    *n_picks_out = t;
    if (state != nullptr) {
      state[1] = t;
      state[0] += t;
    }
  }
}

__global__ void __launch_bounds__(128)
kc_gather_centres_kernel(const float* __restrict__ rows, const float* __restrict__ xx, const int32_t* __restrict__ pick_slots,
                         const int32_t* __restrict__ n_picks, int d, float* __restrict__ centres, float* __restrict__ centre_norms) {
  const int t = blockIdx.x;
  if (t >= *n_picks) return;
  const int s = pick_slots[t];
  const float* src = rows + (int64_t)s * d;
  float* dst = centres + (int64_t)t * d;
  for (int k = threadIdx.x; k < d; k += blockDim.x) dst[k] = src[k];
  if (threadIdx.x == 0) centre_norms[t] = xx[s];
}

size_t kc_resolve_workspace_bytes(int n_blocks, int K, int d) {
  const size_t Kc = (size_t)n_blocks * K;
  return kc_align256(Kc * d * 4) + kc_align256(Kc * 4) * 2 + kc_align256(Kc * 8) + kc_align256(Kc * Kc * 4) +
         kc_align256(Kc * 4) * 4 + 768;
}

int kc_resolve(const void* records, int n_blocks, int K, int d, int max_picks, void* workspace, float* centres,
               float* centre_norms, int64_t* selected_out, int32_t* n_picks_host, cudaStream_t stream, int32_t* state) {
  const int Kc = n_blocks * K;
  MVAL_REQUIRE(K % 4 == 0 && Kc >= 4 && Kc <= kKcMaxSlots, "kcenter resolve: k_slots %% 4 == 0 and n_blocks * k_slots <= %d required", kKcMaxSlots);
  MVAL_REQUIRE(max_picks >= 1, "kcenter resolve: max_picks must be >= 1");
  if (max_picks > Kc) max_picks = Kc;
  KcDeviceScratch* s = nullptr;
  if (int rc = kc_scratch(&s)) return rc;
  char* w = static_cast<char*>(workspace);
  float* rows = reinterpret_cast<float*>(w);            w += kc_align256((size_t)Kc * d * 4);
  float* val = reinterpret_cast<float*>(w);             w += kc_align256((size_t)Kc * 4);
  float* xx = reinterpret_cast<float*>(w);              w += kc_align256((size_t)Kc * 4);
  int64_t* gidx = reinterpret_cast<int64_t*>(w);        w += kc_align256((size_t)Kc * 8);
  float* dt = reinterpret_cast<float*>(w);              w += kc_align256((size_t)Kc * Kc * 4);
  int32_t* pick_slots = reinterpret_cast<int32_t*>(w);  w += kc_align256((size_t)Kc * 4);
  int32_t* rank = reinterpret_cast<int32_t*>(w);        w += kc_align256((size_t)Kc * 4);
  int32_t* order = reinterpret_cast<int32_t*>(w);       w += kc_align256((size_t)Kc * 4);
  int32_t* slot_of_rank = reinterpret_cast<int32_t*>(w); w += kc_align256((size_t)Kc * 4);
  float* tau = reinterpret_cast<float*>(w);
  int32_t* n_picks = reinterpret_cast<int32_t*>(w + 256);
  int32_t* prefix_m = reinterpret_cast<int32_t*>(w + 512);
  kc_unpack_kernel<<<Kc, 128, 0, stream>>>(static_cast<const char*>(records), n_blocks, K, d, kc_records_bytes(K, d), rows, val,
                                           xx, gidx, tau);
  MVAL_LAUNCH_CHECK("kc_unpack");
  if (kc_pairwise_tc_applicable(rows, Kc, d)) {
    if (int rc = kc_pairwise_tc(rows, xx, val, Kc, d, dt, stream)) return rc;
  } else {
    if (int rc = kc_pairwise_exact(rows, xx, Kc, d, dt, stream)) return rc;
  }
  const int threads = kReplayThreads;
  const size_t replay_smem = sizeof(float) * (size_t)(Kc < kReplayStage ? Kc : kReplayStage) * Kc;
  MVAL_CUDA(cudaFuncSetAttribute(kc_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * kReplayStage * kKcMaxSlots)));
  kc_order_kernel<<<(Kc + kOrderThreads - 1) / kOrderThreads, kOrderThreads, 0, stream>>>(val, gidx, Kc, rank, order, slot_of_rank,
                                                                                         prefix_m);
  MVAL_LAUNCH_CHECK("kc_order");
  kc_prefix_kernel<<<(Kc + 7) / 8, 256, 0, stream>>>(val, dt, order, Kc, tau, prefix_m);
  MVAL_LAUNCH_CHECK("kc_prefix");
  kc_replay_kernel<<<1, threads, replay_smem, stream>>>(val, gidx, dt, Kc, tau, max_picks, rank, order, slot_of_rank, prefix_m,
                                                        selected_out, pick_slots, n_picks, state);
  MVAL_LAUNCH_CHECK("kc_replay");
  kc_gather_centres_kernel<<<max_picks, 128, 0, stream>>>(rows, xx, pick_slots, n_picks, d, centres, centre_norms);
  MVAL_LAUNCH_CHECK("kc_gather_centres");
  if (state != nullptr) return MVAL_OK;  // the caller reads the counters when it needs them; nothing waits here
  MVAL_CUDA(cudaMemcpyAsync(s->host_i32, n_picks, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  MVAL_CUDA(cudaStreamSynchronize(stream));
  *n_picks_host = s->host_i32[0];
  return MVAL_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// misc kernels
// ------------------------------------------------------------------------------------------------------------------
__global__ void kcenter_fill_kernel(float* __restrict__ p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__device__ __forceinline__ bool kc_better(float v, int64_t i, float bv, int64_t bi) {
  return v > bv || (v == bv && i < bi);
}

// grid arg-max (value desc, index asc) of m[0..n); the last block folds the per-block partials
__global__ void __launch_bounds__(256)
kc_argmax_kernel(const float* __restrict__ m, int64_t n, int64_t index_offset, KcPartial* partials, unsigned int* done_counter,
                 float* out_val, int64_t* out_idx) {
  __shared__ KcPartial s_part[8];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float bv = -INFINITY;
  int64_t bi = INT64_MAX;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float v = m[i];
    if (kc_better(v, i, bv, bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(kFull, bv, o);
    const int64_t oi = __shfl_xor_sync(kFull, bi, o);
    if (kc_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) s_part[warp] = KcPartial{bv, bi};
  __syncthreads();
  if (threadIdx.x == 0) {
    KcPartial b = s_part[0];
    for (int w = 1; w < 8; ++w)
      if (kc_better(s_part[w].val, s_part[w].idx, b.val, b.idx)) b = s_part[w];
    partials[blockIdx.x] = b;
    __threadfence();
    s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || warp != 0) return;
  __threadfence();
  KcPartial b{-INFINITY, INT64_MAX};
  for (int i = lane; i < (int)gridDim.x; i += kWarp) {
    const float pv = __ldcg(&partials[i].val);
    const int64_t pi = __ldcg(&partials[i].idx);
    if (kc_better(pv, pi, b.val, b.idx)) { b.val = pv; b.idx = pi; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(kFull, b.val, o);
    const int64_t oi = __shfl_xor_sync(kFull, b.idx, o);
    if (kc_better(ov, oi, b.val, b.idx)) { b.val = ov; b.idx = oi; }
  }
  if (lane == 0) {
    *out_val = b.val;
    *out_idx = (b.idx == INT64_MAX) ? -1 : b.idx + index_offset;
    *done_counter = 0u;
  }
}

int kc_norms(const float* X, int64_t n, int d, float* out, cudaStream_t stream) {
  if (n == 0) return MVAL_OK;
  kc_rowdot_kernel<0><<<rowdot_grid(n), kRdWarps * 32, 0, stream>>>(X, n, d, nullptr, nullptr, nullptr, out);
  MVAL_LAUNCH_CHECK("kc_norms");
  return MVAL_OK;
}

}  // namespace mval

// ====================================================================================================================
// C ABI
// ====================================================================================================================
using namespace mval;

extern "C" int mval_kcenter_norms(const float* features, int64_t n, int d, float* row_norms, void* stream) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && d > 0, "mval_kcenter_norms: bad shape");
  if (n == 0) return MVAL_OK;
  MVAL_REQUIRE(features && row_norms, "mval_kcenter_norms: null pointer");
  return kc_norms(features, n, d, row_norms, static_cast<cudaStream_t>(stream));
}

extern "C" int mval_kcenter_update(const float* features, const float* row_norms, int64_t n, int d, const float* centre,
                                   float* min_dist, int64_t index_offset, float* out_best_val, int64_t* out_best_idx,
                                   void* stream_) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && d > 0, "mval_kcenter_update: bad shape");
  MVAL_REQUIRE(centre && out_best_val && out_best_idx, "mval_kcenter_update: null pointer");
  MVAL_REQUIRE(n == 0 || (features && row_norms && min_dist), "mval_kcenter_update: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  KcDeviceScratch* s = nullptr;
  if (int rc = kc_scratch(&s)) return rc;
  float* cc = nullptr;
  MVAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&cc), 256, stream));
  kc_one_norm_kernel<<<1, 32, 0, stream>>>(centre, d, cc);
  MVAL_LAUNCH_CHECK("kc_one_norm");
  int rc = kc_update_batch_exact(features, row_norms, n, d, centre, cc, 1, min_dist, stream);
  if (rc == MVAL_OK) {
    const int64_t want = (n + 2047) / 2048;
    const int grid = (int)(want < 1 ? 1 : (want < 1024 ? want : 1024));
    kc_argmax_kernel<<<grid, 256, 0, stream>>>(min_dist, n, index_offset, s->partials, s->counter, out_best_val, out_best_idx);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "launch kc_argmax");
  }
  cudaError_t e = cudaFreeAsync(cc, stream);
  if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  return rc;
}

extern "C" int mval_kcenter_update_batch(const float* features, const float* row_norms, int64_t n, int d, const float* centres,
                                         const float* centre_norms, int n_centres, float* min_dist, int flags, void* stream) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && d > 0 && n_centres >= 0, "mval_kcenter_update_batch: bad shape");
  if (n == 0 || n_centres == 0) return MVAL_OK;
  MVAL_REQUIRE(features && row_norms && centres && centre_norms && min_dist, "mval_kcenter_update_batch: null pointer");
  return kc_update_batch(features, row_norms, n, d, centres, centre_norms, n_centres, min_dist, flags,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int mval_kcenter_tc_stats(uint64_t* survivors, uint64_t* capacity, void* stream) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(survivors && capacity, "mval_kcenter_tc_stats: null pointer");
  return kc_tc_last_stats(survivors, capacity, static_cast<cudaStream_t>(stream));
}

extern "C" size_t mval_kcenter_records_bytes(int k_slots, int d) { return kc_records_bytes(k_slots, d); }

extern "C" int mval_kcenter_select(const float* features, const float* row_norms, const float* min_dist, int64_t n, int d,
                                   int64_t index_offset, int k_slots, void* records_out, void* stream) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && d > 0, "mval_kcenter_select: bad shape");
  MVAL_REQUIRE(records_out && (n == 0 || (features && row_norms && min_dist)), "mval_kcenter_select: null pointer");
  return kc_select(features, row_norms, min_dist, n, d, index_offset, k_slots, records_out, static_cast<cudaStream_t>(stream));
}

extern "C" size_t mval_kcenter_resolve_workspace_bytes(int n_blocks, int k_slots, int d) {
  return kc_resolve_workspace_bytes(n_blocks, k_slots, d);
}

extern "C" int mval_kcenter_resolve(const void* records, int n_blocks, int k_slots, int d, int max_picks, void* workspace,
                                    float* centres_out, float* centre_norms_out, int64_t* selected_out, int32_t* n_picks_host,
                                    void* stream) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_blocks >= 1 && k_slots >= 2 && d > 0, "mval_kcenter_resolve: bad shape");
  MVAL_REQUIRE(records && workspace && centres_out && centre_norms_out && selected_out && n_picks_host,
               "mval_kcenter_resolve: null pointer");
  return kc_resolve(records, n_blocks, k_slots, d, max_picks, workspace, centres_out, centre_norms_out, selected_out,
                    n_picks_host, static_cast<cudaStream_t>(stream));
}

extern "C" int mval_kcenter_resolve_async(const void* records, int n_blocks, int k_slots, int d, void* workspace, float* centres_out,
                                          float* centre_norms_out, int64_t* selected_out, int32_t* state, void* stream) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_blocks >= 1 && k_slots >= 2 && d > 0, "mval_kcenter_resolve_async: bad shape");
  MVAL_REQUIRE(records && workspace && centres_out && centre_norms_out && selected_out && state,
               "mval_kcenter_resolve_async: null pointer");
  return kc_resolve(records, n_blocks, k_slots, d, n_blocks * k_slots, workspace, centres_out, centre_norms_out, selected_out,
                    nullptr, static_cast<cudaStream_t>(stream), state);
}

extern "C" int mval_kcenter_update_batch_dev(const float* features, const float* row_norms, int64_t n, int d, const float* centres,
                                             const float* centre_norms, int max_centres, const int32_t* n_centres, float* min_dist,
                                             int flags, void* stream) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && d > 0 && max_centres >= 0, "mval_kcenter_update_batch_dev: bad shape");
  if (n == 0 || max_centres == 0) return MVAL_OK;
  MVAL_REQUIRE(features && row_norms && centres && centre_norms && min_dist && n_centres, "mval_kcenter_update_batch_dev: null pointer");
  return kc_update_batch(features, row_norms, n, d, centres, centre_norms, max_centres, min_dist, flags,
                         static_cast<cudaStream_t>(stream), n_centres);
}

extern "C" int mval_kcenter_greedy(const float* features, int64_t n, int64_t n_unlabeled, int d, int32_t budget,
                                   float* min_dist, int64_t* out_selected, void* stream_) {
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n > 0 && d > 0 && budget >= 0, "mval_kcenter_greedy: bad shape");
  // utils/coreset.py: with no labeled centre min_distances stays None and the reference's argmax is undefined
  MVAL_REQUIRE(n_unlabeled >= 0 && n_unlabeled < n, "mval_kcenter_greedy: need at least one labeled row (n_unlabeled < n)");
  MVAL_REQUIRE(features && min_dist && (budget == 0 || out_selected), "mval_kcenter_greedy: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int K = kKcGreedySlots;
  const size_t sz_norms = kc_align256(sizeof(float) * (size_t)n);
  const size_t sz_rec = kc_align256(kc_records_bytes(K, d));
  const size_t sz_ws = kc_align256(kc_resolve_workspace_bytes(1, K, d));
  const size_t sz_centres = kc_align256((size_t)K * d * 4 + 64);  // + the rounds' device counters
  const size_t sz_cn = kc_align256((size_t)K * 4);
  char* ws = nullptr;
  MVAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), sz_norms + sz_rec + sz_ws + sz_centres + sz_cn, stream));
  float* norms = reinterpret_cast<float*>(ws);
  char* rec = ws + sz_norms;
  char* rws = rec + sz_rec;
  float* centres = reinterpret_cast<float*>(rws + sz_ws);
  float* cnorms = reinterpret_cast<float*>(rws + sz_ws + sz_centres);
  auto run = [&]() -> int {
    if (int rc = kc_norms(features, n, d, norms, stream)) return rc;
    kcenter_fill_kernel<<<num_sms() * 4, 256, 0, stream>>>(min_dist, n, INFINITY);
    MVAL_LAUNCH_CHECK("kcenter_fill");
    // coreset.py:83-84  update_distances(al_indices): the labeled rows are contiguous, fold them in in chunks
    const int64_t L = n - n_unlabeled;
    for (int64_t c0 = 0; c0 < L; c0 += kKcInitChunk) {
      const int T = (int)((L - c0) < kKcInitChunk ? (L - c0) : kKcInitChunk);
      if (int rc = kc_update_batch(features, norms, n, d, features + (n_unlabeled + c0) * d, norms + n_unlabeled + c0, T,
                                   min_dist, 0, stream))
        return rc;
    }
    // coreset.py:86-93 in rounds.  The number of picks of a round is only known on the device: resolve advances the
    // counters there and the update reads the round's count there, so no round waits for the host.  The host learns the
    // total one round late (pinned copy + event) and stops launching when it sees the budget reached; the one round that
    // was launched in the meantime finds budget - done = 0 and does nothing.
    KcDeviceScratch* sc = nullptr;
    if (int rc = kc_scratch(&sc)) return rc;
    int32_t* state = reinterpret_cast<int32_t*>(centres + (size_t)K * d);  // behind the centre rows: 4 x int32
    const int32_t init[4] = {0, 0, budget, 0};
    MVAL_CUDA(cudaMemcpyAsync(state, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    cudaEvent_t ev[2];
    MVAL_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    MVAL_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int rc = MVAL_OK;
    for (int r = 0; rc == MVAL_OK && budget > 0; ++r) {
      rc = kc_select(features, norms, min_dist, n, d, 0, K, rec, stream);
      if (rc == MVAL_OK) rc = kc_resolve(rec, 1, K, d, K, rws, centres, cnorms, out_selected, nullptr, stream, state);
      if (rc == MVAL_OK) rc = kc_update_batch(features, norms, n, d, centres, cnorms, K, min_dist, kKcFlagGroupChunks, stream, state + 1);
      if (rc != MVAL_OK) break;
      cudaError_t e = cudaMemcpyAsync(sc->host_i32 + (r & 1), state, sizeof(int32_t), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaEventRecord(ev[r & 1], stream);
      if (e == cudaSuccess && r >= 1) {
        e = cudaEventSynchronize(ev[(r - 1) & 1]);
        if (e == cudaSuccess && sc->host_i32[(r - 1) & 1] >= budget) break;
      }
      if (e != cudaSuccess) rc = cuda_fail(e, "mval_kcenter_greedy round bookkeeping");
      if (r > 4 * (budget + 8)) {
        set_error("mval_kcenter_greedy: internal error, the rounds make no progress");
        rc = MVAL_ERR_CUDA;
      }
    }
    cudaStreamSynchronize(stream);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (rc != MVAL_OK) return rc;
    return MVAL_OK;
  };
  const int rc = run();
  cudaError_t e = cudaFreeAsync(ws, stream);
  if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  return rc;
}
