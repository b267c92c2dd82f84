// Persistent "map stream" kernels (sm_100a) for the per-map heat-map scores on the reference's 64 x 64 maps:
// soft-arg-max (utils/triangulation.py:191-197), MPE / BSB (strategy.py:1149-1176, 1195-1215) and the reprojection-XE
// metric (utils/triangulation.py:236-257).
//
// These scores cost 6-20 instructions per pixel, so a warp that loads its own map from global memory alternates between
// waiting for HBM and computing and the stream stalls (measured round 1e: 0.22-0.65 of the copy bandwidth).  Here the
// two are decoupled exactly as in fused.cu: one CTA per SM, persistent over maps m = blockIdx.x + c * gridDim.x;
//
//   warp 0        TMA producer: one elected lane issues one cp.async.bulk (16 KiB, mbarrier complete_tx) per map into a
//                 ring of kStages stages, so kStages * 16 KiB = 192 KiB per SM are in flight without a single register
//                 or LSU slot spent on staging;
//   warps 1..12   consumers: warp w owns stage w - 1 (map c goes to stage c % 12 and warp c % 12), waits on full[stage],
//                 evaluates the whole map out of shared memory (conflict-free LDS.128, as many passes as it likes),
//                 writes the score and releases empty[stage].
//
// Maps of invalid joints are never read: the producer arrives on the full barrier without a copy and the consumer
// writes NaN.  Every hand-off is an mbarrier; there is no __syncthreads after set-up.  Shapes other than 64 x 64 (or
// unaligned maps) take the one-warp-per-map kernels in decode.cu / peaks.cu / xe.cu.
#include <stdlib.h>

#include "mapops.cuh"

namespace mval {

constexpr int kStreamWarps = 12;
constexpr int kStreamStages = 12;
constexpr int kStreamThreads = kWarp * (1 + kStreamWarps);  // 416
constexpr uint32_t kScratchPerWarp = 1024;  // XE: two tables of 64 doubles
// Ops that are bound by the SM (issue slots / ALU pipe: MPE, BSB) run the ring in DYNAMIC mode: 14 stages for the 12 consumer
// warps, and a warp takes the next map of the CTA (an atomic counter in shared memory) instead of owning a stage.  With one
// stage per warp a consumer idles for the whole refill of its stage after every map (producer reaction + HBM latency + 16 KiB
// at a 148th of the bandwidth, ~1.2 us of the ~6 us a BSB map takes); the two spare stages are always being refilled, so a
// warp that finishes usually finds its next map waiting.  (No aliasing of barrier phases: a warp holds one map, the producer
// fills in order and stalls at the first stage that is still held, so claims reach at most 11 + 13 < 2 x 14 maps ahead.)
template <class Op>
struct StreamCfg {
  static constexpr bool kDynamic = Op::kProducerBackoff != 0;
  static constexpr int kStages = kDynamic ? 14 : kStreamStages;
  static constexpr uint32_t kScratch = kDynamic ? 0u : kStreamWarps * kScratchPerWarp;  // only XeOp uses scratch
  static constexpr uint32_t kSmem = kStages * kMapBytes + 2u * kStages * 8u + 16u + kScratch + kMapAlign;
};

__device__ unsigned long long g_stream_abort[kWdWords];
static WatchdogHost g_stream_watchdog;

int stream_watchdog_poll() { return g_stream_watchdog.poll(g_stream_abort, "map_stream"); }
int stream_watchdog_debug(unsigned long long timeout_cycles, int stall) {
  if (int rc = g_stream_watchdog.prepare(g_stream_abort, "map_stream")) return rc;
  return g_stream_watchdog.debug_set(g_stream_abort, timeout_cycles, stall);
}

template <class Op>
__global__ void __launch_bounds__(kStreamThreads, 1)
map_stream_kernel(const float* __restrict__ hm, int64_t n_maps, const uint8_t* __restrict__ valid, int V, int J,
                  typename Op::Args args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // the staged maps sit on kMapAlign boundaries (mapops.cuh: load_row_rotated); the launch reserves the slack
  unsigned char* smem = smem_raw + ((kMapAlign - (smem_u32(smem_raw) & (kMapAlign - 1u))) & (kMapAlign - 1u));
  using Cfg = StreamCfg<Op>;
  constexpr int kStages = Cfg::kStages;
  float* ring = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kMapBytes);
  uint64_t* empty = full + kStages;
  int* next_map = reinterpret_cast<int*>(empty + kStages);  // dynamic mode: next unclaimed map of this CTA
  unsigned char* scratch = smem + kStages * kMapBytes + 2u * kStages * 8u + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    *next_map = 0;
    mbar_fence_init();
  }
  __syncthreads();
  const int64_t nm = (n_maps > (int64_t)blockIdx.x) ? (n_maps - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int64_t VJ = (int64_t)V * J;
  if (warp == 0) {
    if (lane == 0 && !watchdog_stalled(g_stream_abort)) {
      for (int64_t c = 0; c < nm; ++c) {
        const int64_t m = blockIdx.x + c * (int64_t)gridDim.x;
        const int st = (int)(c % kStages);
        const uint32_t kf = (uint32_t)(c / kStages);
        if (!mbar_wait<Op::kProducerBackoff>(&empty[st], (kf & 1u) ^ 1u, g_stream_abort, 1, c, st)) return;
        if (valid != nullptr && valid[(m / VJ) * J + m % J] == 0) {
          mbar_arrive(&full[st]);  // nothing to read for an invalid joint
        } else {
          mbar_arrive_expect_tx(&full[st], kMapBytes);
          bulk_g2s(ring + (size_t)st * kMapFloats, hm + m * kMapFloats, kMapBytes, &full[st]);
        }
      }
    }
  } else if constexpr (Cfg::kDynamic) {
    static_assert(kStreamWarps <= kStages + 1, "claims must stay less than two ring rounds ahead (barrier phase parity)");
    for (;;) {
      int claimed = 0;
      if (lane == 0) claimed = atomicAdd(next_map, 1);
      const int64_t c = __shfl_sync(kFull, claimed, 0);
      if (c >= nm) break;
      const int st = (int)(c % kStages);
      const uint32_t kf = (uint32_t)(c / kStages);
      const int64_t m = blockIdx.x + c * (int64_t)gridDim.x;
      const bool ok = valid == nullptr || valid[(m / VJ) * J + m % J] != 0;
      if (!mbar_wait(&full[st], kf & 1u, g_stream_abort, 2, c, st)) return;
      Op::run(ring + (size_t)st * kMapFloats, m, ok, lane, args, nullptr, typename Op::Pre{});
      if (Op::kWritesSmem) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
  } else {
    const int w = warp - 1;
    static_assert(kStreamStages == kStreamWarps, "stage c % S must belong to warp c % D");
    float* stage = ring + (size_t)w * kMapFloats;
    typename Op::Pre pre = Op::prefetch(w < nm ? blockIdx.x + w * (int64_t)gridDim.x : 0, args);
    for (int64_t c = w; c < nm; c += kStreamWarps) {
      const int64_t m = blockIdx.x + c * (int64_t)gridDim.x;
      const uint32_t kf = (uint32_t)(c / kStreamStages);
      const bool ok = valid == nullptr || valid[(m / VJ) * J + m % J] != 0;
      const int64_t c_next = c + kStreamWarps;
      const typename Op::Pre nxt = Op::prefetch(c_next < nm ? blockIdx.x + c_next * (int64_t)gridDim.x : m, args);
      if (!mbar_wait(&full[w], kf & 1u, g_stream_abort, 2, c, w)) return;
      Op::run(stage, m, ok, lane, args, scratch + (uint32_t)w * kScratchPerWarp, pre);
      pre = nxt;
      if (Op::kWritesSmem) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[w]);
    }
  }
}

// The per-map Op preceded by the arg-max key-point of the same staged map (the lane = row sweep of mapops.cuh, the very
// function the fused kernel's decode warps run): what a scoring pass needs from the heat maps WITHOUT the triangulation.
// mval_score_pool_scored uses it for MPE / BSB when the split path is selected (capi.cu): the float64 RANSAC then runs as
// its own launch from the key-points instead of sharing the SM with an issue-bound score.
template <class Inner>
struct ArgmaxPlusOp {
  struct Args {
    typename Inner::Args inner;
    int2* out_xy;
    int stride;
  };
  using Pre = typename Inner::Pre;
  static constexpr bool kWritesSmem = Inner::kWritesSmem;
  static constexpr int kProducerBackoff = Inner::kProducerBackoff;
  static constexpr bool kShareRowMax = Inner::kWritesSmem;  // PeaksOp<1> (BSB): the only Op with a row-softmax pass
  __device__ static __forceinline__ Pre prefetch(int64_t m, const Args& a) { return Inner::prefetch(m, a.inner); }
  __device__ static __forceinline__ void run(float* map, int64_t m, bool ok, int lane, const Args& a, unsigned char* scratch,
                                             const Pre& pre) {
    uint32_t idx = 0u;
    float2 rm = make_float2(0.f, 0.f);
    if (ok) idx = warp_argmax_map64(map, lane, kShareRowMax ? &rm : nullptr);
    if (lane == 0)  // evaluation.py:21-26: (c % H, c / H) * stride, (0, 0) for an invalid joint
      a.out_xy[m] = ok ? make_int2((int)(idx % (uint32_t)kMapDim) * a.stride, (int)(idx / (uint32_t)kMapDim) * a.stride) : make_int2(0, 0);
    __syncwarp();
    if constexpr (kShareRowMax)
      Inner::run_rows(map, m, ok, lane, a.inner, &rm);  // BSB's row softmax starts from the sweep's row maxima
    else
      Inner::run(map, m, ok, lane, a.inner, scratch, pre);
  }
};

bool map_stream_applicable(const float* hm, int H, int W) {
  const char* off = getenv("MVAL_NO_STREAM");  // A/B measurements and tests only; read on every call
  return !(off != nullptr && off[0] == '1') && H == kMapDim && W == kMapDim && (reinterpret_cast<uintptr_t>(hm) & 15) == 0;
}

template <class Op>
static int launch_map_stream(const char* name, const float* hm, int64_t n_maps, const uint8_t* valid, int V, int J,
                             const typename Op::Args& args, cudaStream_t stream) {
  if (n_maps == 0) return MVAL_OK;
  constexpr uint32_t kStreamSmem = StreamCfg<Op>::kSmem;
  MVAL_CUDA(cudaFuncSetAttribute(map_stream_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamSmem));
  const int64_t sms = num_sms();
  const int grid = (int)(n_maps < sms ? n_maps : sms);
  if (int rc = g_stream_watchdog.prepare(g_stream_abort, name)) return rc;
  map_stream_kernel<Op><<<grid, kStreamThreads, kStreamSmem, stream>>>(hm, n_maps, valid, V, J, args);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, name);
  return MVAL_OK;
}

int stream_softargmax(const float* hm, int64_t n_maps, float stride, float* out_xy, cudaStream_t stream) {
  return launch_map_stream<SoftArgmaxOp>("map_stream<softargmax>", hm, n_maps, nullptr, 1, 1, {stride, out_xy}, stream);
}
int stream_hp(const float* hm, int64_t n_maps, int V, int J, const uint8_t* valid, float* out, cudaStream_t stream) {
  return launch_map_stream<HpOp>("map_stream<HP>", hm, n_maps, valid, V, J, {out}, stream);
}
int stream_peaks(const float* hm, int64_t n_maps, int V, int J, int mode, const uint8_t* valid, float* out,
                 cudaStream_t stream) {
  if (mode == 0) return launch_map_stream<PeaksOp<0>>("map_stream<MPE>", hm, n_maps, valid, V, J, {out}, stream);
  return launch_map_stream<PeaksOp<1>>("map_stream<BSB>", hm, n_maps, valid, V, J, {out}, stream);
}
int stream_peaks_argmax(const float* hm, int64_t n_maps, int V, int J, int mode, const uint8_t* valid, int stride, float* out,
                        int32_t* out_xy, cudaStream_t stream) {
  int2* xy = reinterpret_cast<int2*>(out_xy);
  if (mode == 0)
    return launch_map_stream<ArgmaxPlusOp<PeaksOp<0>>>("map_stream<argmax+MPE>", hm, n_maps, valid, V, J, {{out}, xy, stride}, stream);
  return launch_map_stream<ArgmaxPlusOp<PeaksOp<1>>>("map_stream<argmax+BSB>", hm, n_maps, valid, V, J, {{out}, xy, stride}, stream);
}
int stream_xe(const float* hm, const double* proj, const double* xyz, int64_t n_maps, int V, int J, double inv_two_sigma2,
              double* out_map, cudaStream_t stream) {
  return launch_map_stream<XeOp>("map_stream<XE>", hm, n_maps, nullptr, V, J, {proj, xyz, V, J, inv_two_sigma2, out_map},
                                 stream);
}

}  // namespace mval
