// Peak-based heat-map uncertainty scores (sm_100a): MPE and BSB (reference strategy.py:1149-1176, 1195-1215).
//
// Both scores are functions of the local peaks skimage.feature.peak_local_max(map, min_distance=2) finds: pixels
// equal to the maximum of their 5x5 window, strictly above the map minimum, at least 2 pixels away from the border.
//   MPE  entropy of softmax over the values of ALL peaks of the raw map                      (:1168-1175)
//   BSB  |p0 - p1| for the two highest peaks of the ROW-softmaxed map (F.softmax without dim) (:1202-1208)
//
// One warp per map, one pass over HBM, no shared memory: lane l owns columns 2l and 2l+1 and walks down the rows
// (one coalesced 256-byte load per row).  The horizontal 5-max comes from four shuffles with the neighbouring
// lanes, the vertical 5-max from a 5-row sliding window held in registers, so every pixel is compared with its
// 5x5 window max two rows after it was loaded.  Peaks feed per-lane streaming accumulators (MPE: running max m,
// S = sum e^(v-m), T = sum (v-m) e^(v-m), H = log S - T/S;  BSB: the two largest values) merged across lanes at the
// end.  The "strictly above the map minimum" rule needs the global minimum, known only at the end: peaks AT the
// minimum all share one value (each lane's lowest peak class), which is therefore kept out of the aggregates until
// the minimum is known.
// Plateau pruning (skimage's ensure_spacing drops later peaks within distance < 2 of an accepted one, which only
// ever happens for exactly equal neighbouring maxima) is not replicated -- stated in DESIGN.md; parity of these two
// scores is unpinned anyway (skimage is not installed anywhere we can run the reference).
#include <math.h>

#include "common.cuh"

namespace mval {

constexpr int kPeakThreads = 128;
constexpr int kPeakWarps = kPeakThreads / kWarp;

struct MpeAcc {  // streaming softmax-entropy accumulator
  float m, S, T;
  int cnt;
};
__device__ __forceinline__ void mpe_add(MpeAcc& a, float v, int n) {  // n peaks of value v
  const float fn = (float)n;
  if (v > a.m) {
    const float sc = (a.m == -INFINITY) ? 0.f : expf(a.m - v);
    a.T = sc * (a.T + (a.m == -INFINITY ? 0.f : (a.m - v)) * a.S);
    a.S = sc * a.S + fn;
    a.m = v;
  } else {
    const float w = fn * expf(v - a.m);
    a.S += w;
    a.T += (v - a.m) * w;
  }
  a.cnt += n;
}
__device__ __forceinline__ void mpe_merge(MpeAcc& a, float m2, float S2, float T2, int c2) {
  if (c2 == 0) return;
  if (a.cnt == 0) { a.m = m2; a.S = S2; a.T = T2; a.cnt = c2; return; }
  const float M = fmaxf(a.m, m2);
  const float s1 = expf(a.m - M), s2 = expf(m2 - M);
  const float T = s1 * (a.T + (a.m - M) * a.S) + s2 * (T2 + (m2 - M) * S2);
  a.S = s1 * a.S + s2 * S2;
  a.T = T;
  a.m = M;
  a.cnt += c2;
}

// kMode 0 = MPE (raw map), 1 = BSB (row-softmaxed map)
template <int kMode>
__global__ void __launch_bounds__(kPeakThreads)
score_peaks_kernel(const float* __restrict__ hm, int64_t n_maps, int V, int J, int H, int W,
                   const uint8_t* __restrict__ valid, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t map = (int64_t)blockIdx.x * kPeakWarps + (threadIdx.x >> 5);
  if (map >= n_maps) return;
  const int j = (int)(map % J);
  const int64_t frame = map / ((int64_t)V * J);
  if (valid != nullptr && valid[frame * J + j] == 0) {
    if (lane == 0) out[map] = __int_as_float(0x7fc00000);
    return;
  }
  const float* __restrict__ p = hm + map * (int64_t)H * W;
  const int c0 = 2 * lane, c1 = c0 + 1;
  const bool in0 = c0 < W, in1 = c1 < W;
  const bool int0 = c0 >= 2 && c0 <= W - 3, int1 = c1 >= 2 && c1 <= W - 3;  // column outside the 2-pixel border
  const bool vec = (W % 2 == 0) && ((reinterpret_cast<uintptr_t>(p) & 7) == 0);

  float h0[5], h1[5], ra[5], rb[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) h0[k] = h1[k] = ra[k] = rb[k] = -INFINITY;
  float gmin = INFINITY;
  MpeAcc acc{-INFINITY, 0.f, 0.f, 0};
  float t1 = -INFINITY, t2 = -INFINITY;  // BSB: two largest peak values
  float cmin = INFINITY;                  // smallest peak value seen by this lane and how many peaks have it
  int n_cmin = 0, n_peaks = 0;

  const int rows_total = H;  // the last interior row H-3 gets its window when row H-1 arrives
  for (int r0 = 0; r0 < rows_total; r0 += 5) {
    float la[5], lb[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {  // issue the group's loads together
      const int r = r0 + k;
      la[k] = lb[k] = -INFINITY;
      if (r < H) {
        if (vec) {
          if (in0) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(p + (int64_t)r * W) + lane);
            la[k] = t.x;
            lb[k] = t.y;
          }
        } else {
          if (in0) la[k] = __ldg(p + (int64_t)r * W + c0);
          if (in1) lb[k] = __ldg(p + (int64_t)r * W + c1);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int r = r0 + k;
      float a = la[k], b = lb[k];
      if (kMode == 1 && r < H) {  // row softmax (float32, like torch): exp(x - rowmax) / rowsum
        float m = fmaxf(a, b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
        const float ea = in0 ? expf(a - m) : 0.f, eb = in1 ? expf(b - m) : 0.f;
        float s = ea + eb;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
        a = in0 ? ea / s : -INFINITY;
        b = in1 ? eb / s : -INFINITY;
      }
      if (r < H) {
        if (in0) gmin = fminf(gmin, a);
        if (in1) gmin = fminf(gmin, b);
      }
      // horizontal 5-max from the neighbouring lanes
      float al = __shfl_up_sync(kFull, a, 1), bl = __shfl_up_sync(kFull, b, 1);
      float ar = __shfl_down_sync(kFull, a, 1), br = __shfl_down_sync(kFull, b, 1);
      if (lane == 0) al = bl = -INFINITY;
      if (lane == 31) ar = br = -INFINITY;
      const float mid = fmaxf(a, b);
      h0[k] = fmaxf(fmaxf(al, bl), fmaxf(mid, ar));
      h1[k] = fmaxf(fmaxf(bl, mid), fmaxf(ar, br));
      ra[k] = a;
      rb[k] = b;
      // the row two above now has its full 5-row window (slots are a ring of 5, so all five slots are its window)
      const int rc = r - 2;
      if (rc >= 2 && rc <= H - 3) {
        const int kc = (k + 3) % 5;
        const float w0 = fmaxf(fmaxf(fmaxf(h0[0], h0[1]), fmaxf(h0[2], h0[3])), h0[4]);
        const float w1 = fmaxf(fmaxf(fmaxf(h1[0], h1[1]), fmaxf(h1[2], h1[3])), h1[4]);
        const float va = ra[kc], vb = rb[kc];
        const bool pa = int0 && va == w0, pb = int1 && vb == w1;
        if (pa || pb) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const bool pk = q == 0 ? pa : pb;
            const float v = q == 0 ? va : vb;
            if (pk) {
              ++n_peaks;
              // the lane's lowest-valued peaks wait outside the aggregate: if they turn out to sit at the map
              // minimum they are not peaks at all, and subtracting them afterwards would cancel catastrophically
              if (v < cmin) {
                if (kMode == 0 && n_cmin > 0) mpe_add(acc, cmin, n_cmin);
                cmin = v;
                n_cmin = 1;
              } else if (v == cmin) {
                ++n_cmin;
              } else if (kMode == 0) {
                mpe_add(acc, v, 1);
              }
              if (kMode == 1) {
                if (v > t1) { t2 = t1; t1 = v; } else if (v > t2) { t2 = v; }
              }
            }
          }
        }
      }
    }
  }
  // ---- merge lanes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmin = fminf(gmin, __shfl_xor_sync(kFull, gmin, o));
  // peaks sitting exactly at the map minimum are not peaks (image > image.min()): they all share the value gmin
  const bool class_is_min = (cmin == gmin);
  if (kMode == 0 && !class_is_min && n_cmin > 0) mpe_add(acc, cmin, n_cmin);  // a genuine peak class after all
  int n_bad = class_is_min ? n_cmin : 0;
  int total = n_peaks;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_bad += __shfl_xor_sync(kFull, n_bad, o);
    total += __shfl_xor_sync(kFull, total, o);
  }
  const int n_good = total - n_bad;
  if (kMode == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(kFull, acc.m, o), S2 = __shfl_xor_sync(kFull, acc.S, o),
                  T2 = __shfl_xor_sync(kFull, acc.T, o);
      const int c2 = __shfl_xor_sync(kFull, acc.cnt, o);
      mpe_merge(acc, m2, S2, T2, c2);
    }
    if (lane == 0) {
      float res = 0.f;  // no peak: the reference sums an empty list
      if (n_good > 0) res = logf(acc.S) - acc.T / acc.S;
      out[map] = res;
    }
  } else {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float u1 = __shfl_xor_sync(kFull, t1, o), u2 = __shfl_xor_sync(kFull, t2, o);
      // merge the two sorted pairs (t1 >= t2), (u1 >= u2)
      const float n1 = fmaxf(t1, u1);
      const float n2 = fmaxf(fminf(t1, u1), fmaxf(t2, u2));
      t1 = n1;
      t2 = n2;
    }
    if (lane == 0) out[map] = (n_good >= 2) ? fabsf(t1 - t2) : __int_as_float(0x7fc00000);
  }
}

bool map_stream_applicable(const float* hm, int H, int W);  // mapstream.cu
int stream_peaks(const float* hm, int64_t n_maps, int V, int J, int mode, const uint8_t* valid, float* out,
                 cudaStream_t stream);

}  // namespace mval

extern "C" int mval_score_peaks(const float* heatmaps, int64_t n_frames, int V, int J, int H, int W, int mode,
                                const uint8_t* valid, float* out_score, void* stream) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && V > 0 && J > 0 && H > 0 && W > 0, "mval_score_peaks: bad shape");
  MVAL_REQUIRE(mode == 0 || mode == 1, "mval_score_peaks: mode must be 0 (MPE) or 1 (BSB)");
  const int64_t n_maps = n_frames * V * J;
  if (n_maps == 0) return MVAL_OK;
  MVAL_REQUIRE(heatmaps && out_score, "mval_score_peaks: null pointer");
  if (map_stream_applicable(heatmaps, H, W))
    return stream_peaks(heatmaps, n_maps, V, J, mode, valid, out_score, static_cast<cudaStream_t>(stream));
  if (W > 64) {
    set_error("mval_score_peaks: maps wider than 64 pixels are not supported (W=%d)", W);
    return MVAL_ERR_UNSUPPORTED;
  }
  const int64_t blocks = (n_maps + kPeakWarps - 1) / kPeakWarps;
  MVAL_REQUIRE(blocks <= 0x7fffffffLL, "mval_score_peaks: too many maps for one launch; chunk the pool");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == 0)
    score_peaks_kernel<0><<<(unsigned)blocks, kPeakThreads, 0, st>>>(heatmaps, n_maps, V, J, H, W, valid, out_score);
  else
    score_peaks_kernel<1><<<(unsigned)blocks, kPeakThreads, 0, st>>>(heatmaps, n_maps, V, J, H, W, valid, out_score);
  MVAL_LAUNCH_CHECK("score_peaks");
  return MVAL_OK;
}
