// Reprojection "cross-entropy" metric (reference utils/triangulation.py:236-257, _compute_xe; enabled by
// AL.USE_REPROJECTION_XE): for every (view, joint) re-render a Gaussian of width sigma at the reprojection of the
// triangulated 3-D joint and add the mean squared difference to the predicted heat map.
//
// One warp per map, one pass over its 16 KiB (128-bit streaming loads).  The Gaussian is separable, so the warp first
// forms ex[x] = exp(-(x - u)^2 / 2 sigma^2) and ey[y] in float64 (W + H exponentials per map instead of W * H) in shared
// memory and a pixel costs one DMUL, one DADD and one DFMA.  The reference renders in float64 (float32 grid minus float64
// point promotes) and MSELoss promotes the float32 prediction, hence the float64 arithmetic here.  As in the reference
// the heat-map PIXEL grid is compared with a point in IMAGE pixels (no division by the stride) and joints left at
// (0, 0, 0) by triangulation() are rendered like any other.  frame metric = sum over views, then joints, in that order.
#include "common.cuh"

namespace mval {

constexpr int kXeWarps = 8;
constexpr int kXeMaxDim = 128;

__global__ void __launch_bounds__(kXeWarps * 32)
score_xe_kernel(const float* __restrict__ hm, const double* __restrict__ proj, const double* __restrict__ xyz, int64_t n_maps,
                int V, int J, int H, int W, double inv_two_sigma2, double* __restrict__ out_map) {
  __shared__ double s_ex[kXeWarps][kXeMaxDim];
  __shared__ double s_ey[kXeWarps][kXeMaxDim];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ex = s_ex[warp];
  double* ey = s_ey[warp];
  const int hw4 = (H * W) >> 2;
  const double inv_hw = 1.0 / (double)(H * W);
  for (int64_t m = (int64_t)blockIdx.x * kXeWarps + warp; m < n_maps; m += (int64_t)gridDim.x * kXeWarps) {
    const int j = (int)(m % J);
    const int64_t fv = m / J;
    const int64_t f = fv / V;
    const double* P = proj + fv * 12;
    const double* X = xyz + (f * J + j) * 3;
    const double x = X[0], y = X[1], z = X[2];
    // [X, 1] @ P^T (utils/triangulation.py:476), then dehomogenise with w == 0 -> 1 (:397-399)
    const double pu = ((x * P[0] + y * P[1]) + z * P[2]) + P[3];
    const double pv = ((x * P[4] + y * P[5]) + z * P[6]) + P[7];
    double pw = ((x * P[8] + y * P[9]) + z * P[10]) + P[11];
    if (pw == 0.0) pw = 1.0;
    const double u = pu / pw, v = pv / pw;
    __syncwarp();
    for (int i = lane; i < W; i += 32) {
      const double dx = (double)i - u;
      ex[i] = exp(-(dx * dx) * inv_two_sigma2);
    }
    for (int i = lane; i < H; i += 32) {
      const double dy = (double)i - v;
      ey[i] = exp(-(dy * dy) * inv_two_sigma2);
    }
    __syncwarp();
    const float4* map4 = reinterpret_cast<const float4*>(hm + m * (int64_t)(H * W));
    double acc0 = 0.0, acc1 = 0.0;
    for (int i = lane; i < hw4; i += 32) {
      const float4 p = ld_stream_f4(map4 + i);
      const int e = i << 2;
      const int row = e / W, col = e - row * W;  // W % 4 == 0: the four pixels share a row
      const double gy = ey[row];
      const double d0 = (double)p.x - ex[col] * gy;
      const double d1 = (double)p.y - ex[col + 1] * gy;
      const double d2 = (double)p.z - ex[col + 2] * gy;
      const double d3 = (double)p.w - ex[col + 3] * gy;
      acc0 = fma(d0, d0, acc0);
      acc1 = fma(d1, d1, acc1);
      acc0 = fma(d2, d2, acc0);
      acc1 = fma(d3, d3, acc1);
    }
    double acc = acc0 + acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if (lane == 0) out_map[m] = acc * inv_hw;
  }
}

// metric[f] = sum over views then joints, sequentially (the order of the reference's  mse_error += ...)
__global__ void xe_frame_reduce_kernel(const double* __restrict__ per_map, int64_t n_frames, int VJ, double* __restrict__ out) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const double* p = per_map + f * VJ;
  double acc = 0.0;
  for (int i = 0; i < VJ; ++i) acc += p[i];
  out[f] = acc;
}

bool map_stream_applicable(const float* hm, int H, int W);  // mapstream.cu
int stream_xe(const float* hm, const double* proj, const double* xyz, int64_t n_maps, int V, int J, double inv_two_sigma2,
              double* out_map, cudaStream_t stream);

}  // namespace mval

extern "C" int mval_score_xe(const float* heatmaps, const double* proj, const double* xyz, int64_t n_frames, int V, int J, int H,
                             int W, double sigma, double* out_map, double* out_metric, void* stream_) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && V > 0 && J > 0 && H > 0 && W > 0, "mval_score_xe: bad shape");
  MVAL_REQUIRE(W % 4 == 0 && W <= kXeMaxDim && H <= kXeMaxDim, "mval_score_xe: W must be a multiple of 4 and H, W <= %d", kXeMaxDim);
  MVAL_REQUIRE(sigma > 0.0, "mval_score_xe: sigma must be positive");
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(heatmaps && proj && xyz && out_metric, "mval_score_xe: null pointer");
  MVAL_REQUIRE((reinterpret_cast<uintptr_t>(heatmaps) & 15) == 0, "mval_score_xe: heatmaps must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t n_maps = n_frames * V * J;
  double* per_map = out_map;
  if (per_map == nullptr) MVAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&per_map), sizeof(double) * n_maps, stream));
  const int64_t want = (n_maps + kXeWarps - 1) / kXeWarps;
  const int64_t cap = (int64_t)num_sms() * 8;
  const int grid = (int)(want < cap ? want : cap);
  cudaError_t e = cudaSuccess;
  if (map_stream_applicable(heatmaps, H, W)) {
    if (stream_xe(heatmaps, proj, xyz, n_maps, V, J, 1.0 / (2.0 * sigma * sigma), per_map, stream) != MVAL_OK) e = cudaErrorUnknown;
  } else {
    score_xe_kernel<<<grid, kXeWarps * 32, 0, stream>>>(heatmaps, proj, xyz, n_maps, V, J, H, W, 1.0 / (2.0 * sigma * sigma), per_map);
    count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    xe_frame_reduce_kernel<<<(unsigned)((n_frames + 127) / 128), 128, 0, stream>>>(per_map, n_frames, V * J, out_metric);
    count_launch();
    e = cudaGetLastError();
  }
  if (out_map == nullptr) {
    cudaError_t e2 = cudaFreeAsync(per_map, stream);
    if (e == cudaSuccess) e = e2;
  }
  if (e != cudaSuccess) return cuda_fail(e, "mval_score_xe");
  return MVAL_OK;
}
