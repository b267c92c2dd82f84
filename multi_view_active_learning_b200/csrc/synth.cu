// Synthetic heat-map renderer (benchmark / test support): a Gaussian bump per map, as the
// reference's dataset renders its ground truth (dataset/dataset.py:198-207), plus counter-based uniform noise.
#include "common.cuh"

namespace mval {

__device__ __forceinline__ uint32_t hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}

__global__ void __launch_bounds__(256)
synth_heatmaps_kernel(const float* __restrict__ centres, int64_t n_maps, int H, int W, float inv2s2, float noise,
                      uint64_t seed, float* __restrict__ out) {
  const int64_t hw = (int64_t)H * W;
  const int64_t total = n_maps * hw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t map = e / hw;
    const int r = (int)((e - map * hw) / W), c = (int)((e - map * hw) % W);
    const float cx = centres[2 * map], cy = centres[2 * map + 1];
    const float d2 = (c - cx) * (c - cx) + (r - cy) * (r - cy);
    const float u = (float)hash32(seed ^ (uint64_t)e * 0x9E3779B97F4A7C15ull) * (1.0f / 4294967296.0f) - 0.5f;
    out[e] = __expf(-d2 * inv2s2) + noise * u;
  }
}

// dataset/dataset.py:198-207: the ground-truth heat map of a joint whose projection, divided by the heat-map stride, is
// (px, py): exp(-((x - px)^2 + (y - py)^2) / (2 sigma^2)) on the pixel grid x = 0..W-1, y = 0..H-1.  The reference subtracts
// a float64 label from a float32 grid, so everything from there on is float64 (torch type promotion); so is this kernel
// (x term first, then y, like torch.sum over the last axis).
__global__ void __launch_bounds__(256)
render_gt_heatmaps_kernel(const double* __restrict__ pts, int64_t n_maps, int H, int W, double two_sigma2, double* __restrict__ out64,
                          float* __restrict__ out32) {
  const int64_t hw = (int64_t)H * W;
  const int64_t total = n_maps * hw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t map = e / hw;
    const int r = (int)((e - map * hw) / W), c = (int)((e - map * hw) % W);
    const double dx = (double)c - pts[2 * map], dy = (double)r - pts[2 * map + 1];
    const double v = exp(-__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) / two_sigma2);
    if (out64) out64[e] = v;
    if (out32) out32[e] = (float)v;
  }
}

}  // namespace mval

extern "C" int mval_render_gt_heatmaps(const double* points, int64_t n_maps, int H, int W, double sigma, double* out_f64,
                                       float* out_f32, void* stream) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(n_maps >= 0 && H > 0 && W > 0 && sigma > 0, "mval_render_gt_heatmaps: bad arguments");
  if (n_maps == 0) return MVAL_OK;
  MVAL_REQUIRE(points && (out_f64 || out_f32), "mval_render_gt_heatmaps: null pointer");
  mval::render_gt_heatmaps_kernel<<<mval::num_sms() * 16, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      points, n_maps, H, W, 2.0 * (sigma * sigma), out_f64, out_f32);
  MVAL_LAUNCH_CHECK("render_gt_heatmaps");
  return MVAL_OK;
}

extern "C" int mval_synth_heatmaps(const float* centres, int64_t n_maps, int H, int W, float sigma, float noise,
                                   uint64_t seed, float* out_heatmaps, void* stream) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(n_maps >= 0 && H > 0 && W > 0 && sigma > 0, "mval_synth_heatmaps: bad arguments");
  if (n_maps == 0) return MVAL_OK;
  MVAL_REQUIRE(centres && out_heatmaps, "mval_synth_heatmaps: null pointer");
  const int blocks = mval::num_sms() * 16;
  mval::synth_heatmaps_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      centres, n_maps, H, W, 1.0f / (2.0f * sigma * sigma), noise, seed, out_heatmaps);
  MVAL_LAUNCH_CHECK("synth_heatmaps");
  return MVAL_OK;
}
