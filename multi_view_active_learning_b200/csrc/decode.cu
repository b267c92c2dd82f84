// Heat-map decode kernels (sm_100a): arg-max, soft-arg-max and the HP confidence score.
//
// One warp owns one H x W float32 map (16 KiB for the reference's 64 x 64 maps).  The map is streamed with
// 128-bit read-only loads that bypass L1 (every byte is used exactly once), kUnroll independent loads per lane
// in flight, and reduced with warp-wide REDUX / shuffle reductions -- the kernels are pure HBM streams:
// algorithmic bytes per map = H*W*4, nothing comparable is written.
//
// Semantics follow the reference literally:
//   arg-max   utils/evaluation.py:24-27   first flat index of the maximum, NaN is the maximum,
//                                         x = (c % H) * stride, y = (c / H) * stride  (H = shape[2] for both)
//   soft      utils/triangulation.py:191-197 + kornia.spatial_soft_argmax2d(normalized_coordinates=False)
//   HP        strategy.py:1185-1186       1 - max(softmax(map, dim=1))  -- the softmax is per ROW
#include "common.cuh"

namespace mval {

constexpr int kDecodeThreads = 256;
constexpr int kDecodeWarps = kDecodeThreads / kWarp;

template <int kUnroll>
__global__ void __launch_bounds__(kDecodeThreads)
decode_argmax_kernel(const float* __restrict__ hm, int64_t n_maps, int V, int J, int H, int hw4, int stride,
                     const uint8_t* __restrict__ valid, int32_t* __restrict__ out_xy, float* __restrict__ out_peak) {
  const int lane = threadIdx.x & 31;
  const int64_t map = (int64_t)blockIdx.x * kDecodeWarps + (threadIdx.x >> 5);
  if (map >= n_maps) return;
  const int j = (int)(map % J);
  const int64_t frame = map / ((int64_t)V * J);
  if (valid != nullptr && valid[frame * J + j] == 0) {  // reference :21-23 -> [0, 0]
    if (lane == 0) {
      reinterpret_cast<int2*>(out_xy)[map] = make_int2(0, 0);
      if (out_peak) out_peak[map] = __int_as_float(0x7fc00000);
    }
    return;
  }
  const float4* __restrict__ p = reinterpret_cast<const float4*>(hm) + map * hw4;
  uint32_t top;
  const uint32_t idx = warp_argmax_map<kUnroll, false>([&](int i) { return ld_stream_f4(p + i); }, hw4, lane, &top);
  if (lane == 0) {
    reinterpret_cast<int2*>(out_xy)[map] = make_int2((int)(idx % (uint32_t)H) * stride, (int)(idx / (uint32_t)H) * stride);
    if (out_peak) out_peak[map] = argmax_key_to_float(top);
  }
}

// Scalar variant for maps whose element count is not a multiple of 4 (never the case for the reference's
// 64 x 64 maps; kept so that the entry point honours every shape the Python function accepts).
__global__ void __launch_bounds__(kDecodeThreads)
decode_argmax_scalar_kernel(const float* __restrict__ hm, int64_t n_maps, int V, int J, int H, int hw, int stride,
                            const uint8_t* __restrict__ valid, int32_t* __restrict__ out_xy,
                            float* __restrict__ out_peak) {
  const int lane = threadIdx.x & 31;
  const int64_t map = (int64_t)blockIdx.x * kDecodeWarps + (threadIdx.x >> 5);
  if (map >= n_maps) return;
  const int j = (int)(map % J);
  const int64_t frame = map / ((int64_t)V * J);
  if (valid != nullptr && valid[frame * J + j] == 0) {
    if (lane == 0) {
      out_xy[2 * map] = 0;
      out_xy[2 * map + 1] = 0;
      if (out_peak) out_peak[map] = __int_as_float(0x7fc00000);
    }
    return;
  }
  const float* __restrict__ p = hm + map * hw;
  uint32_t best_key = 0u, best_idx = 0u;
  for (int i = lane; i < hw; i += kWarp) {
    const uint32_t k = argmax_key(__ldg(p + i));
    if (k > best_key) { best_key = k; best_idx = (uint32_t)i; }
  }
  const uint32_t top = __reduce_max_sync(kFull, best_key);
  const uint32_t idx = __reduce_min_sync(kFull, best_key == top ? best_idx : 0xffffffffu);
  if (lane == 0) {
    out_xy[2 * map] = (int)(idx % (uint32_t)H) * stride;
    out_xy[2 * map + 1] = (int)(idx / (uint32_t)H) * stride;
    if (out_peak) out_peak[map] = argmax_key_to_float(top);
  }
}

int launch_decode_argmax(const float* hm, int64_t n_frames, int V, int J, int H, int W, int stride,
                         const uint8_t* valid, int32_t* out_xy, float* out_peak, cudaStream_t stream) {
  const int64_t n_maps = n_frames * V * J;
  if (n_maps == 0) return MVAL_OK;
  const int64_t blocks = (n_maps + kDecodeWarps - 1) / kDecodeWarps;
  if (blocks > 0x7fffffffLL) {
    set_error("mval_decode_argmax: %lld maps exceed one launch; chunk the pool", (long long)n_maps);
    return MVAL_ERR_UNSUPPORTED;
  }
  const int hw = H * W;
  if (hw % 4 == 0 && (reinterpret_cast<uintptr_t>(hm) & 15) == 0) {
    decode_argmax_kernel<8><<<(unsigned)blocks, kDecodeThreads, 0, stream>>>(hm, n_maps, V, J, H, hw / 4, stride, valid,
                                                                            out_xy, out_peak);
  } else {
    decode_argmax_scalar_kernel<<<(unsigned)blocks, kDecodeThreads, 0, stream>>>(hm, n_maps, V, J, H, hw, stride, valid,
                                                                               out_xy, out_peak);
  }
  MVAL_LAUNCH_CHECK("decode_argmax");
  return MVAL_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// soft-arg-max: softmax over the whole map, expectation of the pixel grid (x = column, y = row), times stride.
// Online (running-max) softmax so the map is read once; weights are float32 expf like kornia's, the three sums are
// accumulated in float64 so that the result does not depend on the summation order at the 1e-6 px level.
// ---------------------------------------------------------------------------------------------------------------
template <int kUnroll>
__global__ void __launch_bounds__(kDecodeThreads)
decode_softargmax_kernel(const float* __restrict__ hm, int64_t n_maps, int W, int hw4, float stride,
                         float* __restrict__ out_xy) {
  const int lane = threadIdx.x & 31;
  const int64_t map = (int64_t)blockIdx.x * kDecodeWarps + (threadIdx.x >> 5);
  if (map >= n_maps) return;
  const float4* __restrict__ p = reinterpret_cast<const float4*>(hm) + map * hw4;
  float m = -INFINITY;
  double s = 0.0, sx = 0.0, sy = 0.0;
  for (int base = lane; base < hw4; base += kWarp * kUnroll) {
    float4 v[kUnroll];
    float bm = -INFINITY;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kWarp;
      if (i < hw4) {
        v[u] = ld_stream_f4(p + i);
        bm = fmaxf(bm, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
      }
    }
    const float m_new = fmaxf(m, bm);
    if (m_new > m) {  // rescale what has been accumulated so far
      const double sc = (m == -INFINITY) ? 0.0 : (double)expf(m - m_new);
      s *= sc; sx *= sc; sy *= sc;
      m = m_new;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kWarp;
      if (i < hw4) {
        const int e = i * 4, row = e / W, col = e - row * W;
        const float w0 = expf(v[u].x - m), w1 = expf(v[u].y - m), w2 = expf(v[u].z - m), w3 = expf(v[u].w - m);
        const float ws = (w0 + w1) + (w2 + w3);
        const float wx = fmaf(3.0f, w3, fmaf(2.0f, w2, w1));
        s += (double)ws;
        sx += fma((double)col, (double)ws, (double)wx);
        sy += (double)row * (double)ws;
      }
    }
  }
  // combine lanes: common maximum, rescale, butterfly sums
  float M = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(kFull, M, o));
  const double sc = (m == -INFINITY) ? 0.0 : (double)expf(m - M);
  s *= sc; sx *= sc; sy *= sc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(kFull, s, o);
    sx += __shfl_xor_sync(kFull, sx, o);
    sy += __shfl_xor_sync(kFull, sy, o);
  }
  if (lane == 0) {
    // the reference multiplies the float32 expectation by stride in float32 (utils/triangulation.py:193-197)
    reinterpret_cast<float2*>(out_xy)[map] = make_float2((float)(sx / s) * stride, (float)(sy / s) * stride);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// HP score (strategy.py:1185-1186): 1 - max over the map of softmax(row).  The row maximum of softmax(row) is
// exp(0) / sum_c exp(x_rc - max_r) = 1 / S_r, so the score is 1 - 1 / min_r S_r.
// W == 64 fast path: one 128-bit load covers 2 rows (16 lanes each); row max / row sum are 16-lane butterflies.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecodeThreads)
score_hp_w64_kernel(const float* __restrict__ hm, int64_t n_maps, int V, int J, int H,
                    const uint8_t* __restrict__ valid, float* __restrict__ out_hp) {
  const int lane = threadIdx.x & 31;
  const int64_t map = (int64_t)blockIdx.x * kDecodeWarps + (threadIdx.x >> 5);
  if (map >= n_maps) return;
  const int j = (int)(map % J);
  const int64_t frame = map / ((int64_t)V * J);
  if (valid != nullptr && valid[frame * J + j] == 0) {
    if (lane == 0) out_hp[map] = __int_as_float(0x7fc00000);
    return;
  }
  const float4* __restrict__ p = reinterpret_cast<const float4*>(hm) + map * (int64_t)H * 16;
  const int n4 = H * 16;
  float min_s = INFINITY;
  bool bad = false;
  constexpr int kUnroll = 8;
  for (int base0 = 0; base0 < n4; base0 += kWarp * kUnroll) {  // warp-uniform trip count: shuffles inside
    const int base = base0 + lane;
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kWarp;
      v[u] = (i < n4) ? ld_stream_f4(p + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float rm = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(kFull, rm, o));
      float rs = (expf(v[u].x - rm) + expf(v[u].y - rm)) + (expf(v[u].z - rm) + expf(v[u].w - rm));
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(kFull, rs, o);
      if (base + u * kWarp < n4) {
        bad |= !(rs >= 1.0f);  // NaN rows (a NaN or an all -inf row) poison the map like torch's softmax does
        min_s = fminf(min_s, rs);
      }
    }
  }
  min_s = fminf(min_s, __shfl_xor_sync(kFull, min_s, 16));
  bad = __any_sync(kFull, bad);
  if (lane == 0) out_hp[map] = bad ? __int_as_float(0x7fc00000) : 1.0f - 1.0f / min_s;
}

__global__ void __launch_bounds__(kDecodeThreads)
score_hp_generic_kernel(const float* __restrict__ hm, int64_t n_maps, int V, int J, int H, int W,
                        const uint8_t* __restrict__ valid, float* __restrict__ out_hp) {
  const int lane = threadIdx.x & 31;
  const int64_t map = (int64_t)blockIdx.x * kDecodeWarps + (threadIdx.x >> 5);
  if (map >= n_maps) return;
  const int j = (int)(map % J);
  const int64_t frame = map / ((int64_t)V * J);
  if (valid != nullptr && valid[frame * J + j] == 0) {
    if (lane == 0) out_hp[map] = __int_as_float(0x7fc00000);
    return;
  }
  const float* __restrict__ p = hm + map * (int64_t)H * W;
  float min_s = INFINITY;
  bool bad = false;
  for (int r = 0; r < H; ++r) {
    float rm = -INFINITY;
    for (int c = lane; c < W; c += kWarp) rm = fmaxf(rm, __ldg(p + r * W + c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(kFull, rm, o));
    float rs = 0.f;
    for (int c = lane; c < W; c += kWarp) rs += expf(__ldg(p + r * W + c) - rm);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(kFull, rs, o);
    bad |= !(rs >= 1.0f);
    min_s = fminf(min_s, rs);
  }
  if (lane == 0) out_hp[map] = bad ? __int_as_float(0x7fc00000) : 1.0f - 1.0f / min_s;
}

// mapstream.cu: persistent TMA-ring kernels for 64 x 64 maps
bool map_stream_applicable(const float* hm, int H, int W);
int stream_softargmax(const float* hm, int64_t n_maps, float stride, float* out_xy, cudaStream_t stream);
int stream_hp(const float* hm, int64_t n_maps, int V, int J, const uint8_t* valid, float* out, cudaStream_t stream);

}  // namespace mval

extern "C" int mval_decode_argmax(const float* heatmaps, int64_t n_frames, int V, int J, int H, int W, int stride,
                                  const uint8_t* valid, int32_t* out_xy, float* out_peak, void* stream) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && V > 0 && J > 0 && H > 0 && W > 0, "mval_decode_argmax: bad shape");
  MVAL_REQUIRE(n_frames == 0 || (heatmaps && out_xy), "mval_decode_argmax: null pointer");
  MVAL_REQUIRE((int64_t)H * W <= (1 << 28), "mval_decode_argmax: map too large");
  return mval::launch_decode_argmax(heatmaps, n_frames, V, J, H, W, stride, valid, out_xy, out_peak,
                                    static_cast<cudaStream_t>(stream));
}

extern "C" int mval_decode_softargmax(const float* heatmaps, int64_t n_frames, int V, int J, int H, int W, float stride,
                                      float* out_xy, void* stream) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && V > 0 && J > 0 && H > 0 && W > 0, "mval_decode_softargmax: bad shape");
  const int64_t n_maps = n_frames * V * J;
  if (n_maps == 0) return MVAL_OK;
  MVAL_REQUIRE(heatmaps && out_xy, "mval_decode_softargmax: null pointer");
  if (mval::map_stream_applicable(heatmaps, H, W))
    return mval::stream_softargmax(heatmaps, n_maps, stride, out_xy, static_cast<cudaStream_t>(stream));
  if (W % 4 != 0 || (reinterpret_cast<uintptr_t>(heatmaps) & 15) != 0) {
    mval::set_error("mval_decode_softargmax: W must be a multiple of 4 and the maps 16-byte aligned");
    return MVAL_ERR_UNSUPPORTED;
  }
  const int64_t blocks = (n_maps + mval::kDecodeWarps - 1) / mval::kDecodeWarps;
  MVAL_REQUIRE(blocks <= 0x7fffffffLL, "mval_decode_softargmax: too many maps for one launch; chunk the pool");
  mval::decode_softargmax_kernel<8><<<(unsigned)blocks, mval::kDecodeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      heatmaps, n_maps, W, H * W / 4, stride, out_xy);
  MVAL_LAUNCH_CHECK("decode_softargmax");
  return MVAL_OK;
}

extern "C" int mval_score_hp(const float* heatmaps, int64_t n_frames, int V, int J, int H, int W, const uint8_t* valid,
                             float* out_hp, void* stream) {
  if (int rc = mval::require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && V > 0 && J > 0 && H > 0 && W > 0, "mval_score_hp: bad shape");
  const int64_t n_maps = n_frames * V * J;
  if (n_maps == 0) return MVAL_OK;
  MVAL_REQUIRE(heatmaps && out_hp, "mval_score_hp: null pointer");
  const int64_t blocks = (n_maps + mval::kDecodeWarps - 1) / mval::kDecodeWarps;
  MVAL_REQUIRE(blocks <= 0x7fffffffLL, "mval_score_hp: too many maps for one launch; chunk the pool");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mval::map_stream_applicable(heatmaps, H, W)) return mval::stream_hp(heatmaps, n_maps, V, J, valid, out_hp, st);
  if (W == 64 && (reinterpret_cast<uintptr_t>(heatmaps) & 15) == 0)
    mval::score_hp_w64_kernel<<<(unsigned)blocks, mval::kDecodeThreads, 0, st>>>(heatmaps, n_maps, V, J, H, valid, out_hp);
  else
    mval::score_hp_generic_kernel<<<(unsigned)blocks, mval::kDecodeThreads, 0, st>>>(heatmaps, n_maps, V, J, H, W, valid,
                                                                                   out_hp);
  MVAL_LAUNCH_CHECK("score_hp");
  return MVAL_OK;
}
