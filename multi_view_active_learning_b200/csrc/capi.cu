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

int launch_score_pool_fused_segments(const float* const* seg_ptr, const int64_t* seg_frames, int n_segments,
                                     const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J, int H, int W,
                                     int stride, const mval_ransac_params& prm, int map_score, int32_t* out_xy, double* out_xyz,
                                     double* out_reproj, int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count,
                                     float* out_map_score, cudaStream_t stream);

bool map_stream_applicable(const float* hm, int H, int W);
int stream_peaks_argmax(const float* hm, int64_t n_maps, int V, int J, int mode, const uint8_t* valid, int stride, float* out,
                        int32_t* out_xy, cudaStream_t stream);

// MPE / BSB with the triangulation: one fused launch, or the stream kernel (score + arg-max key-point, one read of the heat
// maps) followed by the RANSAC launches from the key-points.  Measured per 16 384 frames (profiles/r2_summary.md section 5,
// final state): MPE 9.8 ms fused / 8.4 ms split, BSB 11.5 ms fused / 10.9 ms split -- both take the split path by default.
// MVAL_SCORED_SPLIT=0 / 1 forces one of them for both (A/B measurements and tests; read on every call).
static bool scored_split_enabled(int map_score) {
  const char* e = getenv("MVAL_SCORED_SPLIT");
  if (e != nullptr && (e[0] == '0' || e[0] == '1')) return e[0] == '1';
  return map_score == MVAL_MAP_SCORE_MPE || map_score == MVAL_MAP_SCORE_BSB;
}

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
  if ((map_score == MVAL_MAP_SCORE_MPE || map_score == MVAL_MAP_SCORE_BSB) && scored_split_enabled(map_score) && fused_enabled() &&
      params->pairs == nullptr && map_stream_applicable(heatmaps, H, W)) {
    // The heat maps are still read ONCE (score and arg-max key-point of a map come from the same staged copy); the
    // triangulation follows from the 8-byte key-points as its own launches instead of sharing the SM with an issue-bound
    // score.  Same device functions as the fused kernel, hence bit-identical results.
    int32_t* xy = out_xy;
    void* scratch = nullptr;
    if (xy == nullptr) {
      MVAL_CUDA(cudaMallocAsync(&scratch, sizeof(int32_t) * 2 * n_frames * V * J, stream));
      xy = static_cast<int32_t*>(scratch);
    }
    int rc = stream_peaks_argmax(heatmaps, n_frames * V * J, V, J, map_score == MVAL_MAP_SCORE_MPE ? 0 : 1, valid, stride,
                                 out_map_score, xy, stream);
    if (rc == MVAL_OK)
      rc = triangulate_ransac(xy, 0, proj, valid, n_frames, V, J, params, out_xyz, out_reproj, out_inliers, nullptr, out_metric,
                              out_inlier_count, stream);
    if (scratch) {
      cudaError_t e = cudaFreeAsync(scratch, stream);
      if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
    }
    return rc;
  }
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

int mval_score_pool_segments(const float* const* seg_heatmaps, const int64_t* seg_frames, int n_segments, const double* proj,
                             const uint8_t* valid, int V, int J, int H, int W, int stride,
                             const mval_ransac_params* params, int map_score, int32_t* out_xy, double* out_xyz,
                             double* out_reproj, int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count,
                             float* out_map_score, void* stream) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(seg_heatmaps && seg_frames && n_segments >= 1 && n_segments <= MVAL_MAX_SEGMENTS,
               "mval_score_pool_segments: 1..%d segments", MVAL_MAX_SEGMENTS);
  int64_t n_frames = 0;
  for (int s = 0; s < n_segments; ++s) {
    MVAL_REQUIRE(seg_frames[s] >= 0 && (seg_frames[s] == 0 || seg_heatmaps[s] != nullptr), "mval_score_pool_segments: bad segment %d", s);
    n_frames += seg_frames[s];
  }
  if (int rc = mval::check_pool_args("mval_score_pool_segments", seg_heatmaps[0], proj, n_frames, V, J, H, W, params, out_xyz,
                                     out_metric, out_inlier_count))
    return rc;
  MVAL_REQUIRE(map_score >= MVAL_MAP_SCORE_NONE && map_score <= MVAL_MAP_SCORE_BSB && (map_score == MVAL_MAP_SCORE_NONE || out_map_score),
               "mval_score_pool_segments: bad map_score / out_map_score");
  if (n_frames == 0) return MVAL_OK;
  const int rc = mval::launch_score_pool_fused_segments(seg_heatmaps, seg_frames, n_segments, proj, valid, n_frames, V, J, H, W,
                                                        stride, *params, map_score, out_xy, out_xyz, out_reproj, out_inliers,
                                                        out_metric, out_inlier_count, out_map_score,
                                                        static_cast<cudaStream_t>(stream));
  if (rc == MVAL_ERR_UNSUPPORTED) mval::set_error("mval_score_pool_segments: shape / alignment / pair table not covered by the fused kernel");
  return rc;
}

}  // extern "C"
