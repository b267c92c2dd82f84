// Host-buffer streaming pipeline (include/mval_b200.h: mval_pipeline_*, mval_score_pool_host): a pool whose heat maps
// live in HOST memory is pushed through the device in chunks.  A pipeline handle owns `n_slots` slots, each with its own
// stream and device staging buffers sized for `chunk_frames` frames; slot s processes chunks s, s + n_slots, ...: the
// host->device copy of chunk k + 1 (another slot's stream, a copy engine) overlaps the kernels of chunk k, and the
// results go back with async device->host copies on the same stream.  Nothing is allocated or freed per call: the
// handle is created once (mval_score_pool_host keeps one per device and shape).  Pageable host memory still works (the
// copies then serialise with the host); pinned memory is what makes the overlap real.
#include <mutex>

#include "common.cuh"

namespace mval {

int score_pool(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J, int H,
               int W, int stride, const mval_ransac_params* params, int map_score, int32_t* out_xy, double* out_xyz,
               double* out_reproj, int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count,
               float* out_map_score, cudaStream_t stream);

struct Slot {
  cudaStream_t stream = nullptr;
  float* hm = nullptr;
  double* proj = nullptr;
  uint8_t* valid = nullptr;
  void* xy = nullptr;  // int32 (arg-max) or float32 (soft-arg-max) [c][V][J][2]
  double* xyz = nullptr;
  double* reproj = nullptr;
  int32_t* inliers = nullptr;
  uint32_t* mask = nullptr;
  double* metric = nullptr;
  int32_t* inlier_count = nullptr;
  float* map_score = nullptr;
};

}  // namespace mval

struct mval_pipeline {
  int device = 0, V = 0, J = 0, H = 0, W = 0, n_slots = 0;
  int64_t chunk_frames = 0;
  mval::Slot slots[4];
};

namespace mval {

static void free_pipeline(mval_pipeline* p) {
  if (p == nullptr) return;
  for (int i = 0; i < p->n_slots; ++i) {
    Slot& s = p->slots[i];
    cudaFree(s.hm); cudaFree(s.proj); cudaFree(s.valid); cudaFree(s.xy); cudaFree(s.xyz); cudaFree(s.reproj);
    cudaFree(s.inliers); cudaFree(s.mask); cudaFree(s.metric); cudaFree(s.inlier_count); cudaFree(s.map_score);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  (void)cudaGetLastError();
  delete p;
}

static int create_pipeline(int V, int J, int H, int W, int64_t chunk_frames, int n_slots, mval_pipeline** out) {
  MVAL_REQUIRE(out != nullptr, "mval_pipeline_create: null pointer");
  *out = nullptr;
  MVAL_REQUIRE(V >= 2 && J > 0 && H > 0 && W > 0, "mval_pipeline_create: bad shape");
  if (V > MVAL_MAX_VIEWS) {
    set_error("mval_pipeline_create: V=%d exceeds MVAL_MAX_VIEWS=%d", V, MVAL_MAX_VIEWS);
    return MVAL_ERR_UNSUPPORTED;
  }
  const size_t frame_hm = sizeof(float) * (size_t)V * J * H * W;
  if (chunk_frames <= 0) {
    // ~256 MiB of heat maps per chunk: large enough to amortise launches, small enough to start overlapping early
    chunk_frames = (int64_t)((256ull << 20) / frame_hm);
    if (chunk_frames < 1) chunk_frames = 1;
  }
  if (n_slots <= 0) n_slots = 3;
  if (n_slots > 4) n_slots = 4;
  mval_pipeline* p = new mval_pipeline();
  MVAL_CUDA(cudaGetDevice(&p->device));
  p->V = V; p->J = J; p->H = H; p->W = W;
  p->chunk_frames = chunk_frames;
  p->n_slots = n_slots;
  const size_t c = (size_t)chunk_frames;
  cudaError_t e = cudaSuccess;
  auto take = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  for (int i = 0; i < n_slots && e == cudaSuccess; ++i) {
    Slot& s = p->slots[i];
    take(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    take(cudaMalloc(&s.hm, frame_hm * c));
    take(cudaMalloc(&s.proj, sizeof(double) * 12 * V * c));
    take(cudaMalloc(&s.valid, (size_t)J * c));
    take(cudaMalloc(&s.xy, sizeof(int32_t) * 2 * V * J * c));
    take(cudaMalloc(&s.xyz, sizeof(double) * 3 * J * c));
    take(cudaMalloc(&s.reproj, sizeof(double) * J * c));
    take(cudaMalloc(&s.inliers, sizeof(int32_t) * J * c));
    take(cudaMalloc(&s.mask, sizeof(uint32_t) * J * c));
    take(cudaMalloc(&s.metric, sizeof(double) * c));
    take(cudaMalloc(&s.inlier_count, sizeof(int32_t) * c));
    take(cudaMalloc(&s.map_score, sizeof(float) * V * J * c));
  }
  if (e != cudaSuccess) {
    free_pipeline(p);
    return cuda_fail(e, "mval_pipeline_create");
  }
  *out = p;
  return MVAL_OK;
}

static int run_pipeline(mval_pipeline* p, const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames,
                        int stride, const mval_ransac_params* params, const mval_pipeline_options* opt, void* out_xy,
                        double* out_xyz, double* out_reproj, int32_t* out_inliers, double* out_metric,
                        int32_t* out_inlier_count, float* out_map_score) {
  const int V = p->V, J = p->J, H = p->H, W = p->W;
  const size_t frame_hm = sizeof(float) * (size_t)V * J * H * W;
  const int map_score = opt ? opt->map_score : MVAL_MAP_SCORE_NONE;
  const bool soft = opt && opt->use_soft_argmax;
  const bool xe = opt && opt->use_reprojection_xe;
  const bool refine = opt && opt->direct_optimization;
  int rc = MVAL_OK;
  cudaError_t e = cudaSuccess;
  int dev = 0;
  MVAL_CUDA(cudaGetDevice(&dev));
  MVAL_REQUIRE(dev == p->device, "mval_pipeline: the handle belongs to device %d, the current device is %d", p->device, dev);
  int k = 0;
  for (int64_t f0 = 0; f0 < n_frames && rc == MVAL_OK && e == cudaSuccess; f0 += p->chunk_frames, ++k) {
    Slot& s = p->slots[k % p->n_slots];
    const int64_t n = (n_frames - f0 < p->chunk_frames) ? (n_frames - f0) : p->chunk_frames;
    auto take = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    take(cudaMemcpyAsync(s.hm, reinterpret_cast<const char*>(heatmaps) + frame_hm * f0, frame_hm * n, cudaMemcpyHostToDevice, s.stream));
    take(cudaMemcpyAsync(s.proj, proj + (size_t)12 * V * f0, sizeof(double) * 12 * V * n, cudaMemcpyHostToDevice, s.stream));
    if (valid) take(cudaMemcpyAsync(s.valid, valid + (size_t)J * f0, (size_t)J * n, cudaMemcpyHostToDevice, s.stream));
    if (e != cudaSuccess) break;
    const uint8_t* dvalid = valid ? s.valid : nullptr;
    mval_ransac_params q = *params;
    q.frame_offset = params->frame_offset + f0;
    q.pairs = nullptr;
    if (q.frame_keys) q.frame_keys = params->frame_keys + f0;  // a DEVICE array over the whole pool
    if (!soft && !refine) {
      // arg-max decode + RANSAC (+ HP / MPE / BSB of the same maps) in the fused persistent kernel
      rc = score_pool(s.hm, s.proj, dvalid, n, V, J, H, W, stride, &q, map_score, static_cast<int32_t*>(s.xy), s.xyz, s.reproj,
                      s.inliers, s.metric, s.inlier_count, s.map_score, s.stream);
    } else {
      // utils/triangulation.py:191-200 (soft-arg-max key-points) and / or :319-336 (Huber refinement): unfused kernels
      if (soft)
        rc = mval_decode_softargmax(s.hm, n, V, J, H, W, (float)stride, static_cast<float*>(s.xy), s.stream);
      else
        rc = mval_decode_argmax(s.hm, n, V, J, H, W, stride, dvalid, static_cast<int32_t*>(s.xy), nullptr, s.stream);
      if (rc == MVAL_OK)
        rc = mval_triangulate_ransac(s.xy, soft ? 1 : 0, s.proj, dvalid, n, V, J, &q, s.xyz, s.reproj, s.inliers, s.mask, s.metric,
                                     s.inlier_count, s.stream);
      if (rc == MVAL_OK && refine)
        rc = mval_refine_huber(s.xy, soft ? 1 : 0, s.proj, dvalid, s.mask, s.inliers, n, V, J, s.xyz, s.reproj, s.metric,
                               s.inlier_count, nullptr, s.stream);
      if (rc == MVAL_OK && map_score == MVAL_MAP_SCORE_HP) rc = mval_score_hp(s.hm, n, V, J, H, W, dvalid, s.map_score, s.stream);
      if (rc == MVAL_OK && (map_score == MVAL_MAP_SCORE_MPE || map_score == MVAL_MAP_SCORE_BSB))
        rc = mval_score_peaks(s.hm, n, V, J, H, W, map_score == MVAL_MAP_SCORE_MPE ? 0 : 1, dvalid, s.map_score, s.stream);
    }
    if (rc == MVAL_OK && xe)  // utils/triangulation.py:223-224: the metric becomes the reprojection XE of the same maps
      rc = mval_score_xe(s.hm, s.proj, s.xyz, n, V, J, H, W, opt->sigma, nullptr, s.metric, s.stream);
    if (rc != MVAL_OK) break;
    if (out_xy) take(cudaMemcpyAsync(static_cast<char*>(out_xy) + sizeof(int32_t) * 2 * V * J * f0, s.xy, sizeof(int32_t) * 2 * V * J * n, cudaMemcpyDeviceToHost, s.stream));
    take(cudaMemcpyAsync(out_xyz + (size_t)3 * J * f0, s.xyz, sizeof(double) * 3 * J * n, cudaMemcpyDeviceToHost, s.stream));
    if (out_reproj) take(cudaMemcpyAsync(out_reproj + (size_t)J * f0, s.reproj, sizeof(double) * J * n, cudaMemcpyDeviceToHost, s.stream));
    if (out_inliers) take(cudaMemcpyAsync(out_inliers + (size_t)J * f0, s.inliers, sizeof(int32_t) * J * n, cudaMemcpyDeviceToHost, s.stream));
    take(cudaMemcpyAsync(out_metric + f0, s.metric, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
    take(cudaMemcpyAsync(out_inlier_count + f0, s.inlier_count, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s.stream));
    if (out_map_score && map_score != MVAL_MAP_SCORE_NONE)
      take(cudaMemcpyAsync(out_map_score + (size_t)V * J * f0, s.map_score, sizeof(float) * V * J * n, cudaMemcpyDeviceToHost, s.stream));
  }
  // drain every slot even after a failure: the caller's host buffers must not be written after return
  for (int i = 0; i < p->n_slots; ++i) {
    cudaError_t r = cudaStreamSynchronize(p->slots[i].stream);
    if (e == cudaSuccess) e = r;
  }
  if (rc != MVAL_OK) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "mval_pipeline_score_pool");
  return mval_check_async(p->slots[0].stream);  // a watchdog trip of any chunk is reported by THIS call
}

static std::mutex g_cache_mutex;
static mval_pipeline* g_cached = nullptr;  // the implicit handle of mval_score_pool_host

}  // namespace mval

extern "C" {

int mval_pipeline_create(int V, int J, int H, int W, int64_t chunk_frames, int n_slots, mval_pipeline** out) {
  if (int rc = mval::require_device()) return rc;
  return mval::create_pipeline(V, J, H, W, chunk_frames, n_slots, out);
}

int mval_pipeline_destroy(mval_pipeline* p) {
  if (p == nullptr) return MVAL_OK;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaSetDevice(p->device);
  mval::free_pipeline(p);
  cudaSetDevice(dev);
  return MVAL_OK;
}

int mval_pipeline_score_pool(mval_pipeline* p, const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames,
                             int stride, const mval_ransac_params* params, const mval_pipeline_options* options, void* out_xy,
                             double* out_xyz, double* out_reproj, int32_t* out_inliers, double* out_metric,
                             int32_t* out_inlier_count, float* out_map_score) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(p != nullptr && params != nullptr && n_frames >= 0, "mval_pipeline_score_pool: null handle / params or bad size");
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(heatmaps && proj && out_xyz && out_metric && out_inlier_count, "mval_pipeline_score_pool: null pointer");
  MVAL_REQUIRE(params->pairs == nullptr, "mval_pipeline_score_pool: explicit pair tables are not supported on the host path");
  if (options) {
    MVAL_REQUIRE(options->map_score >= MVAL_MAP_SCORE_NONE && options->map_score <= MVAL_MAP_SCORE_BSB,
                 "mval_pipeline_score_pool: map_score must be one of MVAL_MAP_SCORE_*");
    MVAL_REQUIRE(options->map_score == MVAL_MAP_SCORE_NONE || out_map_score != nullptr, "mval_pipeline_score_pool: out_map_score is null");
    MVAL_REQUIRE(!options->use_reprojection_xe || options->sigma > 0.0, "mval_pipeline_score_pool: sigma must be positive");
  }
  return mval::run_pipeline(p, heatmaps, proj, valid, n_frames, stride, params, options, out_xy, out_xyz, out_reproj, out_inliers,
                            out_metric, out_inlier_count, out_map_score);
}

int mval_score_pool_host(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J, int H,
                         int W, int stride, const mval_ransac_params* params, int64_t chunk_frames, int32_t* out_xy,
                         double* out_xyz, double* out_reproj, int32_t* out_inliers, double* out_metric,
                         int32_t* out_inlier_count) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(params != nullptr && n_frames >= 0 && V >= 2 && J > 0 && H > 0 && W > 0, "mval_score_pool_host: bad arguments");
  if (n_frames == 0) return MVAL_OK;
  std::lock_guard<std::mutex> lock(mval::g_cache_mutex);
  int dev = 0;
  MVAL_CUDA(cudaGetDevice(&dev));
  mval_pipeline*& c = mval::g_cached;
  const int64_t want = chunk_frames > 0 ? chunk_frames : 0;
  if (c != nullptr && (c->device != dev || c->V != V || c->J != J || c->H != H || c->W != W || (want > 0 && c->chunk_frames != want))) {
    mval_pipeline_destroy(c);
    c = nullptr;
  }
  if (c == nullptr)
    if (int rc = mval::create_pipeline(V, J, H, W, want, 0, &c)) return rc;
  return mval_pipeline_score_pool(c, heatmaps, proj, valid, n_frames, stride, params, nullptr, out_xy, out_xyz, out_reproj,
                                  out_inliers, out_metric, out_inlier_count, nullptr);
}

}  // extern "C"
