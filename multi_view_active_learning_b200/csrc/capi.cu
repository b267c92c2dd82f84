// C-ABI plumbing: error strings, launch accounting, the fused pool-scoring entry points and the host-buffer
// streaming pipeline (include/mval_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace mval {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return e == cudaErrorMemoryAllocation ? MVAL_ERR_OUT_OF_MEMORY : MVAL_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int require_device() {
  static int cached = 1;  // 1 = unknown, 0 = ok, <0 = error
  if (cached == 1) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      (void)cudaGetLastError();
      set_error("no CUDA device available (%s); mval_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
      return MVAL_ERR_NO_DEVICE;  // not cached: a device may appear after e.g. CUDA_VISIBLE_DEVICES changes in tests
    }
    // Keep stream-ordered scratch (cudaMallocAsync) cached in the default pool: with the default release threshold
    // of 0 every synchronisation hands the memory back to the OS and the next call pays a fresh cudaMalloc
    // (measured: 5.6 ms median, up to 800 ms, per mval_topk_desc call).
    cached = 0;
  }
  // only the device the caller works on, once per device (not a process-wide side effect on every visible GPU)
  static bool tuned[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !tuned[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
    tuned[dev] = true;
  }
  return cached;
}

int num_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

int fused_watchdog_poll();
int fused_watchdog_debug(unsigned long long timeout_cycles, int stall);
int stream_watchdog_poll();
int stream_watchdog_debug(unsigned long long timeout_cycles, int stall);

int launch_decode_argmax(const float* hm, int64_t n_frames, int V, int J, int H, int W, int stride,
                         const uint8_t* valid, int32_t* out_xy, float* out_peak, cudaStream_t stream);
int triangulate_ransac(const void* xy, int xy_is_float, const double* proj, const uint8_t* valid, int64_t n_frames,
                       int V, int J, const mval_ransac_params* params, double* out_xyz, double* out_reproj,
                       int32_t* out_inliers, uint32_t* out_mask, double* out_metric, int32_t* out_inlier_count,
                       cudaStream_t stream);

int launch_score_pool_fused(const float* hm, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J,
                            int H, int W, int stride, const mval_ransac_params& prm, int map_score, int32_t* out_xy,
                            double* out_xyz, double* out_reproj, int32_t* out_inliers, double* out_metric,
                            int32_t* out_inlier_count, float* out_map_score, cudaStream_t stream);

// MVAL_FUSED=0 in the environment forces the three-launch path (A/B measurements only).
static bool fused_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("MVAL_FUSED");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

static int check_pool_args(const char* who, const void* heatmaps, const void* proj, int64_t n_frames, int V, int J,
                           int H, int W, const mval_ransac_params* params, const void* out_xyz, const void* out_metric,
                           const void* out_inlier_count) {
  MVAL_REQUIRE(params != nullptr, "%s: params is null", who);
  MVAL_REQUIRE(n_frames >= 0 && V >= 2 && J > 0 && H > 0 && W > 0, "%s: bad shape", who);
  MVAL_REQUIRE(n_frames == 0 || (heatmaps && proj && out_xyz && out_metric && out_inlier_count), "%s: null pointer", who);
  if (V > MVAL_MAX_VIEWS) {
    set_error("%s: V=%d exceeds MVAL_MAX_VIEWS=%d", who, V, MVAL_MAX_VIEWS);
    return MVAL_ERR_UNSUPPORTED;
  }
  return MVAL_OK;
}

int score_pool(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J, int H,
               int W, int stride, const mval_ransac_params* params, int map_score, int32_t* out_xy, double* out_xyz,
               double* out_reproj, int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count,
               float* out_map_score, cudaStream_t stream) {
  if (n_frames == 0) return MVAL_OK;
  if (fused_enabled()) {
    const int rc = launch_score_pool_fused(heatmaps, proj, valid, n_frames, V, J, H, W, stride, *params, map_score, out_xy,
                                           out_xyz, out_reproj, out_inliers, out_metric, out_inlier_count, out_map_score, stream);
    if (rc != MVAL_ERR_UNSUPPORTED) return rc;  // unsupported shape: fall through to the separate launches
  }
  if (map_score != MVAL_MAP_SCORE_NONE) {  // second pass over the heat maps: the stand-alone score kernel
    const int rc = map_score == MVAL_MAP_SCORE_HP
                       ? mval_score_hp(heatmaps, n_frames, V, J, H, W, valid, out_map_score, stream)
                       : mval_score_peaks(heatmaps, n_frames, V, J, H, W, map_score == MVAL_MAP_SCORE_MPE ? 0 : 1, valid,
                                          out_map_score, stream);
    if (rc != MVAL_OK) return rc;
  }
  int32_t* xy = out_xy;
  void* scratch = nullptr;
  if (xy == nullptr) {
    MVAL_CUDA(cudaMallocAsync(&scratch, sizeof(int32_t) * 2 * n_frames * V * J, stream));
    xy = static_cast<int32_t*>(scratch);
  }
  int rc = launch_decode_argmax(heatmaps, n_frames, V, J, H, W, stride, valid, xy, nullptr, stream);
  if (rc == MVAL_OK)
    rc = triangulate_ransac(xy, 0, proj, valid, n_frames, V, J, params, out_xyz, out_reproj, out_inliers, nullptr,
                            out_metric, out_inlier_count, stream);
  if (scratch) {
    cudaError_t e = cudaFreeAsync(scratch, stream);
    if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  }
  return rc;
}

// Host-buffer pipeline: two slots, each with its own stream and device staging buffers.  Slot s processes chunks
// s, s+2, ...: H2D of chunk k+1 (other slot's stream) overlaps the kernels of chunk k; results go back with
// async D2H on the same stream.  Pageable host memory still works (the copies then serialise with the host).
struct Slot {
  cudaStream_t stream = nullptr;
  float* hm = nullptr;
  double* proj = nullptr;
  uint8_t* valid = nullptr;
  int32_t* xy = nullptr;
  double* xyz = nullptr;
  double* reproj = nullptr;
  int32_t* inliers = nullptr;
  double* metric = nullptr;
  int32_t* inlier_count = nullptr;
};

static void free_slot(Slot& s) {
  cudaFree(s.hm); cudaFree(s.proj); cudaFree(s.valid); cudaFree(s.xy); cudaFree(s.xyz); cudaFree(s.reproj);
  cudaFree(s.inliers); cudaFree(s.metric); cudaFree(s.inlier_count);
  if (s.stream) cudaStreamDestroy(s.stream);
  s = Slot();
}

int score_pool_host(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J,
                    int H, int W, int stride, const mval_ransac_params* params, int64_t chunk_frames, int32_t* out_xy,
                    double* out_xyz, double* out_reproj, int32_t* out_inliers, double* out_metric,
                    int32_t* out_inlier_count) {
  if (n_frames == 0) return MVAL_OK;
  const size_t frame_hm = sizeof(float) * (size_t)V * J * H * W;
  if (chunk_frames <= 0) {
    // ~256 MiB of heat maps per chunk: large enough to amortise launches, small enough to start overlapping early
    chunk_frames = (int64_t)((256ull << 20) / frame_hm);
    if (chunk_frames < 1) chunk_frames = 1;
  }
  if (chunk_frames > n_frames) chunk_frames = n_frames;
  const int n_slots = (n_frames > chunk_frames) ? 2 : 1;
  Slot slots[2];
  int rc = MVAL_OK;
  auto fail = [&](int code) {
    cudaDeviceSynchronize();
    for (auto& s : slots) free_slot(s);
    return code;
  };
#define SLOT_CUDA(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return fail(cuda_fail(e__, #call)); \
  } while (0)
  const size_t c = (size_t)chunk_frames;
  for (int i = 0; i < n_slots; ++i) {
    Slot& s = slots[i];
    SLOT_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    SLOT_CUDA(cudaMalloc(&s.hm, frame_hm * c));
    SLOT_CUDA(cudaMalloc(&s.proj, sizeof(double) * 12 * V * c));
    if (valid) SLOT_CUDA(cudaMalloc(&s.valid, (size_t)J * c));
    SLOT_CUDA(cudaMalloc(&s.xy, sizeof(int32_t) * 2 * V * J * c));
    SLOT_CUDA(cudaMalloc(&s.xyz, sizeof(double) * 3 * J * c));
    SLOT_CUDA(cudaMalloc(&s.reproj, sizeof(double) * J * c));
    SLOT_CUDA(cudaMalloc(&s.inliers, sizeof(int32_t) * J * c));
    SLOT_CUDA(cudaMalloc(&s.metric, sizeof(double) * c));
    SLOT_CUDA(cudaMalloc(&s.inlier_count, sizeof(int32_t) * c));
  }
  int k = 0;
  for (int64_t f0 = 0; f0 < n_frames; f0 += chunk_frames, ++k) {
    Slot& s = slots[k % n_slots];
    const int64_t n = (n_frames - f0 < chunk_frames) ? (n_frames - f0) : chunk_frames;
    SLOT_CUDA(cudaMemcpyAsync(s.hm, reinterpret_cast<const char*>(heatmaps) + frame_hm * f0, frame_hm * n,
                              cudaMemcpyHostToDevice, s.stream));
    SLOT_CUDA(cudaMemcpyAsync(s.proj, proj + (size_t)12 * V * f0, sizeof(double) * 12 * V * n, cudaMemcpyHostToDevice,
                              s.stream));
    if (valid) SLOT_CUDA(cudaMemcpyAsync(s.valid, valid + (size_t)J * f0, (size_t)J * n, cudaMemcpyHostToDevice, s.stream));
    mval_ransac_params p = *params;
    p.frame_offset = params->frame_offset + f0;
    if (p.pairs) p.pairs = nullptr;  // explicit pair tables are a device-pointer feature; validated by the caller below
    if (p.frame_keys) p.frame_keys = params->frame_keys + f0;  // a DEVICE array over the whole pool
    rc = score_pool(s.hm, s.proj, valid ? s.valid : nullptr, n, V, J, H, W, stride, &p, MVAL_MAP_SCORE_NONE, s.xy, s.xyz,
                    s.reproj, s.inliers, s.metric, s.inlier_count, nullptr, s.stream);
    if (rc != MVAL_OK) return fail(rc);
    if (out_xy) SLOT_CUDA(cudaMemcpyAsync(out_xy + (size_t)2 * V * J * f0, s.xy, sizeof(int32_t) * 2 * V * J * n, cudaMemcpyDeviceToHost, s.stream));
    SLOT_CUDA(cudaMemcpyAsync(out_xyz + (size_t)3 * J * f0, s.xyz, sizeof(double) * 3 * J * n, cudaMemcpyDeviceToHost, s.stream));
    if (out_reproj) SLOT_CUDA(cudaMemcpyAsync(out_reproj + (size_t)J * f0, s.reproj, sizeof(double) * J * n, cudaMemcpyDeviceToHost, s.stream));
    if (out_inliers) SLOT_CUDA(cudaMemcpyAsync(out_inliers + (size_t)J * f0, s.inliers, sizeof(int32_t) * J * n, cudaMemcpyDeviceToHost, s.stream));
    SLOT_CUDA(cudaMemcpyAsync(out_metric + f0, s.metric, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
    SLOT_CUDA(cudaMemcpyAsync(out_inlier_count + f0, s.inlier_count, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s.stream));
  }
  for (int i = 0; i < n_slots; ++i) SLOT_CUDA(cudaStreamSynchronize(slots[i].stream));
  for (auto& s : slots) free_slot(s);
#undef SLOT_CUDA
  return MVAL_OK;
}

}  // namespace mval

extern "C" {

int mval_version(void) { return MVAL_ABI_VERSION; }
const char* mval_last_error(void) { return mval::g_error; }
uint64_t mval_launch_count(void) { return mval::g_launches.load(std::memory_order_relaxed); }

int mval_check_async(void* stream) {
  if (int rc = mval::require_device()) return rc;
  MVAL_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  if (int rc = mval::fused_watchdog_poll()) return rc;
  return mval::stream_watchdog_poll();
}

int mval_debug_watchdog(uint64_t timeout_cycles, int stall) {
  if (int rc = mval::require_device()) return rc;
  if (int rc = mval::fused_watchdog_debug(timeout_cycles, stall)) return rc;
  return mval::stream_watchdog_debug(timeout_cycles, stall);
}

int mval_score_pool(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J,
                    int H, int W, int stride, const mval_ransac_params* params, int32_t* out_xy, double* out_xyz,
                    double* out_reproj, int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count,
                    void* stream) {
  if (int rc = mval::require_device()) return rc;
  if (int rc = mval::check_pool_args("mval_score_pool", heatmaps, proj, n_frames, V, J, H, W, params, out_xyz,
                                     out_metric, out_inlier_count))
    return rc;
  return mval::score_pool(heatmaps, proj, valid, n_frames, V, J, H, W, stride, params, MVAL_MAP_SCORE_NONE, out_xy, out_xyz,
                          out_reproj, out_inliers, out_metric, out_inlier_count, nullptr, static_cast<cudaStream_t>(stream));
}

int mval_score_pool_scored(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V,
                           int J, int H, int W, int stride, const mval_ransac_params* params, int map_score,
                           int32_t* out_xy, double* out_xyz, double* out_reproj, int32_t* out_inliers,
                           double* out_metric, int32_t* out_inlier_count, float* out_map_score, void* stream) {
  if (int rc = mval::require_device()) return rc;
  if (int rc = mval::check_pool_args("mval_score_pool_scored", heatmaps, proj, n_frames, V, J, H, W, params, out_xyz,
                                     out_metric, out_inlier_count))
    return rc;
  MVAL_REQUIRE(map_score >= MVAL_MAP_SCORE_NONE && map_score <= MVAL_MAP_SCORE_BSB,
               "mval_score_pool_scored: map_score must be one of MVAL_MAP_SCORE_*");
  MVAL_REQUIRE(map_score == MVAL_MAP_SCORE_NONE || n_frames == 0 || out_map_score != nullptr,
               "mval_score_pool_scored: out_map_score is null");
  return mval::score_pool(heatmaps, proj, valid, n_frames, V, J, H, W, stride, params, map_score, out_xy, out_xyz,
                          out_reproj, out_inliers, out_metric, out_inlier_count, out_map_score,
                          static_cast<cudaStream_t>(stream));
}

int mval_score_pool_host(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V,
                         int J, int H, int W, int stride, const mval_ransac_params* params, int64_t chunk_frames,
                         int32_t* out_xy, double* out_xyz, double* out_reproj, int32_t* out_inliers, double* out_metric,
                         int32_t* out_inlier_count) {
  if (int rc = mval::require_device()) return rc;
  if (int rc = mval::check_pool_args("mval_score_pool_host", heatmaps, proj, n_frames, V, J, H, W, params, out_xyz,
                                     out_metric, out_inlier_count))
    return rc;
  MVAL_REQUIRE(params->pairs == nullptr, "mval_score_pool_host: explicit pair tables are not supported on the host path");
  return mval::score_pool_host(heatmaps, proj, valid, n_frames, V, J, H, W, stride, params, chunk_frames, out_xy, out_xyz,
                               out_reproj, out_inliers, out_metric, out_inlier_count);
}

}  // extern "C"
