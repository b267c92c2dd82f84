// Per-map heat-map score evaluators ("Ops") shared by the persistent kernels: map_stream_kernel (mapstream.cu) runs
// one Op per map on its own; score_pool_fused_kernel (fused.cu) runs HP / MPE / BSB in its decode warps right after the
// arg-max, so that a strategy that needs both the triangulation and a per-map score reads every heat map ONCE.
//
// An Op evaluates one 64 x 64 float32 map that already sits in shared memory, with one warp:
//   Op::run(map, m, ok, lane, args, scratch, pre)   m = global map index of the output, ok = joint is valid
//   Op::kWritesSmem                                 the Op rewrites the stage (the caller must fence before the refill)
//   Op::kProducerBackoff                            0: the stream is HBM-bound; 2: the Op is issue-bound and the producer sleeps
//                                                   between polls for a free stage (tma.cuh: mbar_wait)
//   Op::prefetch(m, args)                           per-map inputs fetched one map ahead (XE only)
#pragma once
#include "tma.cuh"

namespace mval {

constexpr int kMapDim = 64;
constexpr int kMapFloats = kMapDim * kMapDim;
constexpr uint32_t kMapBytes = kMapFloats * 4u;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {  // one FMNMX3 (sm_100+)
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float y;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
// bits |= bit where x == w, as FSETP + a predicated IMAD (bits + bit * 1: the bit is not set yet).  The C form came out
// as FSETP + SEL + LOP3 for three of the four columns of a scan step; and the scan is bound by the ALU pipe (FMNMX, LOP3,
// IADD3 issue every second cycle per scheduler, B300_MICROARCH "pipe rates"), where 26 of its 31 instructions per step ran,
// while the FMA pipe (FFMA, IMAD) idles -- so the OR is spelled as a multiply-add.
__device__ __forceinline__ void or_if_equal(uint32_t& bits, float x, float w, uint32_t bit) {
  asm("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, 1, %0;\n\t}" : "+r"(bits) : "f"(x), "f"(w), "r"(bit));
}
// Highest set bit of a peak mask, branch-free: returns its index (-1 for an empty mask: bfind) and clears it (PTX shl clamps
// a shift of 0xffffffff to 0, which C++ does not promise).
__device__ __forceinline__ int pop_highest_bit(uint32_t& b) {
  int i;
  uint32_t t;
  asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(b));
  asm("shl.b32 %0, %1, %2;" : "=r"(t) : "r"(1u), "r"(i));
  b ^= t;
  return i;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ uint32_t order_key(float x) {  // monotone float -> uint (no NaN handling needed here)
  const uint32_t u = __float_as_uint(x);
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float order_key_inv(uint32_t k) {
  return __uint_as_float(k ^ ((k & 0x80000000u) ? 0x80000000u : 0xffffffffu));
}
__device__ __forceinline__ float warp_min_f(float v) { return order_key_inv(__reduce_min_sync(kFull, order_key(v))); }
__device__ __forceinline__ float warp_max_f(float v) { return order_key_inv(__reduce_max_sync(kFull, order_key(v))); }

// ---------------------------------------------------------------------------------------------------------------
// Row-wise sweeps of a 64 x 64 map in shared memory with lane = row (rows `lane` and `lane + 32`): a lane pulls its
// whole row into 64 registers, float4 column blocks in the rotated order (k + lane) % 16, which keeps every quarter-warp
// on eight distinct 16-byte bank groups (conflict-free LDS.128).  Row statistics then need no shuffles at all.
//
// Arg-max on top of the row maxima (torch.argmax semantics: first index of the maximum, NaN is the maximum, -0.0 ==
// +0.0): every lane keeps (largest row maximum, first row attaining it); afterwards one REDUX pair picks the first such
// row of the map and half a warp re-reads that single row to find the first column.  0.75 instructions per element
// (FMNMX3 pairs + a packed-add NaN / infinity sentinel) against 2.75 for the per-vector compare-and-select scan of
// warp_argmax_map; a map whose sentinel fires (NaN, +-inf, or a sum that overflows) is re-scanned by warp_argmax_map
// with its exact monotone-key compare, so the result is the same function of the map in every case.
// ---------------------------------------------------------------------------------------------------------------
// Rows of a staged map are visited with lane = row; a lane takes the sixteen 16-byte blocks of its row in the order
// k ^ (lane & 15), which keeps every quarter-warp on eight distinct bank groups (conflict-free LDS.128 / STS.128).  The staged
// maps are 256-byte aligned (kMapAlign: both kernels align their ring), so block k sits at (row address ^ lane bits) ^ (k << 4):
// ONE LOP3 per access.  (Round 2: the (k + lane) & 15 rotation kept sixteen index registers alive and cost two IMADs per
// access, 96 instructions per map in the arg-max sweep and ~250 in BSB's softmax pass -- profiles/r2z_lines_*.)
constexpr uint32_t kMapAlign = 256;
__device__ __forceinline__ uint32_t row_block_base(const void* map, int row, int lane) {
  return ((uint32_t)__cvta_generic_to_shared(map) + (uint32_t)row * (kMapDim * 4u)) ^ ((uint32_t)(lane & 15) << 4);
}
__device__ __forceinline__ float4 lds_block(uint32_t rx, int k) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(rx ^ ((uint32_t)k << 4)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_block(uint32_t rx, int k, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rx ^ ((uint32_t)k << 4)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void load_row_rotated(const float4* p4, int row, int lane, float4 (&x)[16]) {
  const uint32_t rx = row_block_base(p4, row, lane);
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = lds_block(rx, k);
}
__device__ __forceinline__ float row_max(const float4 (&x)[16]) {
  float rm = -INFINITY;
#pragma unroll
  for (int k = 0; k < 16; k += 2)
    rm = max3(rm, max3(x[k].x, x[k].y, x[k].z), max3(x[k].w, x[k + 1].x, max3(x[k + 1].y, x[k + 1].z, x[k + 1].w)));
  return rm;
}
// (best, best_row) of every lane -> flat index of the first maximum of the map.  Only valid when the map holds no NaN.
__device__ __forceinline__ uint32_t argmax_from_row_maxima(const float4* p4, int lane, float best, int best_row) {
  const float top = warp_max_f(best + 0.0f);  // + 0.0f: -0.0 and +0.0 are the same maximum
  const uint32_t row = __reduce_min_sync(kFull, best == top ? (uint32_t)best_row : 0xffffffffu);
  const float4 w = p4[row * 16 + (lane & 15)];
  const uint32_t c = w.x == top ? 0u : (w.y == top ? 1u : (w.z == top ? 2u : (w.w == top ? 3u : 0xffffu)));
  const uint32_t col = __reduce_min_sync(kFull, c == 0xffffu ? 0xffffffffu : (uint32_t)(lane & 15) * 4u + c);
  return row * (uint32_t)kMapDim + col;
}
// row_max_out (may be nullptr): the maxima of rows lane and lane + 32 as row_max() computes them -- BSB's softmax pass over
// the same staged map starts from them instead of recomputing 32 FMNMX3 per row.
__device__ __forceinline__ uint32_t warp_argmax_map64(const float* map, int lane, float2* row_max_out = nullptr) {
  const float4* p4 = reinterpret_cast<const float4*>(map);
  float best = -INFINITY;
  int best_row = lane;
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int rr = 0; rr < 2; ++rr) {
    const int row = lane + 32 * rr;
    float4 x[16];
    load_row_rotated(p4, row, lane, x);
    const float rm = row_max(x);
    if (row_max_out != nullptr) {  // (a select, not an index: the pair stays in registers)
      if (rr == 0) row_max_out->x = rm; else row_max_out->y = rm;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = __fadd2_rn(acc, __fadd2_rn(make_float2(x[k].x, x[k].y), make_float2(x[k].z, x[k].w)));
    const bool gt = rm > best;  // strict: the lower row is kept on ties
    best = gt ? rm : best;
    best_row = gt ? row : best_row;
  }
  const float sentinel = (acc.x + acc.y) * 0.0f;  // NaN as soon as the lane saw a NaN or an infinity
  if (__any_sync(kFull, sentinel != sentinel))
    return warp_argmax_map<8, true>([&](int q) { return p4[q]; }, kMapFloats / 4, lane, nullptr);
  return argmax_from_row_maxima(p4, lane, best, best_row);
}

// ---------------------------------------------------------------------------------------------------------------
// soft-arg-max: softmax over the whole map, expectation of the pixel grid, times stride (kornia
// spatial_soft_argmax2d(normalized_coordinates=False)).  Two passes over shared memory: the map maximum, then
// w = 2^(x log2e - M log2e) (one FFMA + one MUFU per pixel; the rounding of M log2e scales every weight alike and
// cancels in the ratio).  With W = 64 a lane's float4 column is fixed (4 * (lane % 16)) and its row is 2u + lane / 16,
// so per vector only sum(w), sum(k w_k) and row * sum(w) are accumulated -- in float32 over 8 vectors, in float64
// across them and across lanes.
// ---------------------------------------------------------------------------------------------------------------
struct NoPrefetch {};

struct SoftArgmaxOp {
  struct Args {
    float stride;
    float* out_xy;
  };
  using Pre = NoPrefetch;
  static constexpr bool kWritesSmem = false;
  static constexpr int kProducerBackoff = 0;
  __device__ static __forceinline__ Pre prefetch(int64_t, const Args&) { return {}; }
  __device__ static __forceinline__ void run(float* map, int64_t m, bool ok, int lane, const Args& a, unsigned char*, const Pre&) {
    (void)ok;
    const float4* __restrict__ p = reinterpret_cast<const float4*>(map) + lane;
    float mx = -INFINITY;
#pragma unroll 8
    for (int u = 0; u < 32; ++u) {
      const float4 t = p[u * 32];
      mx = fmaxf(mx, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    constexpr float kLog2e = 1.4426950408889634f;
    const float nml = -mx * kLog2e;
    const float half = (float)(lane >> 4);
    double S = 0.0, X = 0.0, Y = 0.0;
#pragma unroll 1
    for (int blk = 0; blk < 4; ++blk) {
      float s32 = 0.f, x32 = 0.f, y32 = 0.f;
#pragma unroll
      for (int uu = 0; uu < 8; ++uu) {
        const int u = blk * 8 + uu;
        const float4 t = p[u * 32];
        const float w0 = ex2_approx(fmaf(t.x, kLog2e, nml)), w1 = ex2_approx(fmaf(t.y, kLog2e, nml));
        const float w2 = ex2_approx(fmaf(t.z, kLog2e, nml)), w3 = ex2_approx(fmaf(t.w, kLog2e, nml));
        const float ws = (w0 + w1) + (w2 + w3);
        s32 += ws;
        x32 += fmaf(3.0f, w3, fmaf(2.0f, w2, w1));
        y32 = fmaf((float)(2 * u) + half, ws, y32);
      }
      S += (double)s32;
      X += (double)x32;
      Y += (double)y32;
    }
    X = fma((double)((lane & 15) * 4), S, X);
    S = warp_sum(S);
    X = warp_sum(X);
    Y = warp_sum(Y);
    if (lane == 0)  // the reference multiplies the float32 expectation by stride in float32 (:193-197)
      reinterpret_cast<float2*>(a.out_xy)[m] = make_float2((float)(X / S) * a.stride, (float)(Y / S) * a.stride);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// HP (strategy.py:1185-1186): 1 - max over the map of softmax(row); the row maximum of softmax(row) is 1 / S_r with
// S_r = sum_c exp(x_rc - max_r), so the score is 1 - 1 / min_r S_r.  Lane = row (see above): a row's maximum and its
// S_r are formed in registers without a shuffle, in packed float32x2 arithmetic; exp = ex2.approx of the exactly formed
// difference times log2e (the maximum itself contributes exactly 1); 3.5 instructions per pixel against 10 for the
// 16-lanes-per-row butterflies of round 1e.  A NaN row (a NaN, an infinity or an all -inf row) poisons the map like
// torch's softmax.  eval<true> also returns the arg-max of the map (the row maxima are already there), which is what
// the decode warps of score_pool_fused_kernel<HP> use instead of a separate arg-max sweep.
// ---------------------------------------------------------------------------------------------------------------
struct HpOp {
  struct Args {
    float* out;
  };
  using Pre = NoPrefetch;
  static constexpr bool kWritesSmem = false;
  static constexpr int kProducerBackoff = 0;
  __device__ static __forceinline__ Pre prefetch(int64_t, const Args&) { return {}; }
  template <bool kArgmax>
  __device__ static __forceinline__ uint32_t eval(const float* map, int64_t m, int lane, const Args& a) {
    constexpr float kLog2e = 1.4426950408889634f;
    const float4* p4 = reinterpret_cast<const float4*>(map);
    float min_s = INFINITY, best = -INFINITY;
    int best_row = lane;
    bool bad = false;
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
      const int row = lane + 32 * rr;
      float4 x[16];
      load_row_rotated(p4, row, lane, x);
      const float rm = row_max(x);
      const float2 nrm = make_float2(-rm, -rm), l2e = make_float2(kLog2e, kLog2e);
      float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float2 ta = __fmul2_rn(__fadd2_rn(make_float2(x[k].x, x[k].y), nrm), l2e);
        const float2 tb = __fmul2_rn(__fadd2_rn(make_float2(x[k].z, x[k].w), nrm), l2e);
        sa = __fadd2_rn(sa, make_float2(ex2_approx(ta.x), ex2_approx(ta.y)));
        sb = __fadd2_rn(sb, make_float2(ex2_approx(tb.x), ex2_approx(tb.y)));
      }
      const float s = (sa.x + sa.y) + (sb.x + sb.y);
      bad |= (s != s);
      min_s = fminf(min_s, s);
      if (kArgmax) {
        const bool gt = rm > best;
        best = gt ? rm : best;
        best_row = gt ? row : best_row;
      }
    }
    min_s = warp_min_f(min_s);
    bad = __any_sync(kFull, bad);
    if (lane == 0) a.out[m] = bad ? __int_as_float(0x7fc00000) : 1.0f - 1.0f / min_s;
    if (!kArgmax) return 0u;
    if (bad)  // warp-uniform; the map holds a NaN, an infinity or an all -inf row: exact monotone-key scan
      return warp_argmax_map<8, true>([&](int q) { return p4[q]; }, kMapFloats / 4, lane, nullptr);
    return argmax_from_row_maxima(p4, lane, best, best_row);
  }
  __device__ static __forceinline__ void run(float* map, int64_t m, bool ok, int lane, const Args& a, unsigned char*, const Pre&) {
    if (!ok) {
      if (lane == 0) a.out[m] = __int_as_float(0x7fc00000);
      return;
    }
    eval<false>(map, m, lane, a);
  }
};

// skimage's peak_local_max ends with ensure_spacing(coords sorted by descending intensity, spacing = min_distance,
// p_norm = inf): a peak is dropped when an ACCEPTED peak lies at Chebyshev distance < 2 from it, i.e. in its 8-neighbourhood
// (skimage/_shared/coord.py).  Two 5 x 5 local maxima that touch hold the same value -- a plateau -- so the rule only acts
// among equal neighbours, in the stable order of the sort: ascending flat index.  In raster order that is the recurrence
//     accepted(r, c) = peak(r, c) and not (accepted(r-1, c-1) or accepted(r-1, c) or accepted(r-1, c+1) or accepted(r, c-1)),
// evaluated here on 64-bit row masks, warp-uniformly (every lane computes every row and keeps the bits of its own four
// columns).  Only taken when some peak bits touch (never on real-valued backbone outputs; heat maps with flat regions).
// Layout of the peak bits as in PeaksOp::run: lane = (half, l16); bit i of b[k] <-> row half * 30 + 2 + i, column 4 l16 + k.
__device__ __forceinline__ uint64_t spread_nibbles(uint32_t x16) {  // bit l of x16 -> bit 4 l
  uint64_t x = x16;
  x = (x | (x << 24)) & 0x000000ff000000ffull;
  x = (x | (x << 12)) & 0x000f000f000f000full;
  x = (x | (x << 6)) & 0x0303030303030303ull;
  x = (x | (x << 3)) & 0x1111111111111111ull;
  return x;
}

static __device__ __noinline__ uint4 prune_touching_peaks(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, const float* col, float gmin,
                                                   int lane) {
  const int half = lane >> 4, l16 = lane & 15;
  // image > image.min() is applied before the spacing rule (skimage: the mask is thresholded first)
  uint32_t f[4] = {b0, b1, b2, b3};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t b = f[k], keep = 0u;
    while (b) {
      const int i = __ffs(b) - 1;
      b &= b - 1;
      if (col[i * kMapDim + k] > gmin) keep |= 1u << i;
    }
    f[k] = keep;
  }
  uint32_t out[4] = {0u, 0u, 0u, 0u};
  uint64_t prev = 0ull;
#pragma unroll 1
  for (int r = 0; r < 60; ++r) {
    const int h = r / 30, i = r - h * 30;
    uint64_t row = 0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t bal = __ballot_sync(kFull, (f[k] >> i) & 1u);
      row |= spread_nibbles(h ? (bal >> 16) : (bal & 0xffffu)) << k;
    }
    const uint64_t cand = row & ~(prev | (prev << 1) | (prev >> 1));
    uint64_t acc = cand & ~(cand << 1);  // the first pixel of every horizontal run, then every second one
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
      const uint64_t nxt = acc | ((acc << 2) & cand & (cand << 1));
      if (nxt == acc) break;
      acc = nxt;
    }
    prev = acc;
    if (h == half) {
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k] |= (uint32_t)((acc >> (4 * l16 + k)) & 1ull) << i;
    }
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

// ---------------------------------------------------------------------------------------------------------------
// MPE (kMode 0) / BSB (kMode 1): local peaks as skimage.feature.peak_local_max(map, min_distance=2) defines them
// (equal to the maximum of their 5 x 5 window, strictly above the map minimum, 2 pixels off the border).
//
// Scan layout: lane = (half, l16): float4 column block 4 * l16 .. 4 * l16 + 3 of row stream `half` -- half 0 walks rows
// 0..33 and tests rows 2..31, half 1 walks rows 30..63 and tests rows 32..61 -- so one LDS.128 per lane feeds two rows
// per warp instruction.  Horizontal 5-max: four 16-lane shuffles bring the two neighbouring values on each side (or
// their pair maxima), three-input FMNMX does the rest; vertical 5-max: a register window of the last five horizontal
// maxima, two FMNMX3 per pixel.  A pixel equal to its window maximum sets one bit in a per-column mask (30 tested rows
// per stream), nothing else happens inside the scan: peaks are frequent on noisy maps (one per ~25 pixels), so any
// per-peak work in the scan would run, diverged, on almost every row.  Afterwards each lane walks its own set bits,
// one bit of each of its four column masks per trip (about 2-3 trips), with the map still in shared memory:
//   MPE = entropy of softmax over the peak values = log S - T / S with S = sum e^(v - c), T = sum (v - c) e^(v - c);
//         the shift c is the maximum of the whole map (known from the scan, >= every peak), so one pass suffices; should
//         every term underflow (a border pixel ~100 above every peak) a second pass shifts by the largest peak instead;
//   BSB = |p0 - p1| of the two highest peaks.
// Warp-wide minima / maxima / counts go through REDUX on monotone integer keys (one instruction instead of a five-step
// shuffle butterfly); only the two float sums of MPE use a butterfly.
// Equal neighbouring maxima (a plateau above the map minimum) all count as peaks; skimage's ensure_spacing would keep
// a subset of them that depends on an unstable argsort -- see DESIGN.md section 2.
// BSB works on the ROW-softmaxed map (F.softmax without dim on a 2-D tensor), so a first pass rewrites the stage in
// place, p = exp(x - rowmax) / rowsum.  That pass runs with lane = row (rows l and l + 32): a lane holds its whole row in
// 64 registers, so the row statistics need no shuffles at all; the float4 column blocks are visited in the order
// k ^ (lane & 15) (load_row_rotated above: conflict-free LDS.128 / STS.128).  exp = ex2.approx of the exactly formed difference times log2e (the row maximum contributes exactly 1), the
// quotient is e * (1/s) corrected by one residual step.
// ---------------------------------------------------------------------------------------------------------------
template <int kMode>
struct PeaksOp {
  struct Args {
    float* out;
  };
  using Pre = NoPrefetch;
  static constexpr bool kWritesSmem = (kMode == 1);
  static constexpr int kProducerBackoff = 2;
  __device__ static __forceinline__ Pre prefetch(int64_t, const Args&) { return {}; }
  __device__ static __forceinline__ void run(float* map, int64_t m, bool ok, int lane, const Args& a, unsigned char*, const Pre&) {
    run_rows(map, m, ok, lane, a, nullptr);
  }
  // known_row_max (BSB only, may be nullptr): maxima of rows lane and lane + 32 from an arg-max sweep over the same map
  __device__ static __forceinline__ void run_rows(float* map, int64_t m, bool ok, int lane, const Args& a, const float2* known_row_max) {
    if (!ok) {
      if (lane == 0) a.out[m] = __int_as_float(0x7fc00000);
      return;
    }
    const int half = lane >> 4, l16 = lane & 15;
    float4* p4 = reinterpret_cast<float4*>(map);
    float gmin = INFINITY, gmax = -INFINITY;
    if (kMode == 1) {
      constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {
        const uint32_t rx = row_block_base(p4, lane + 32 * rr, lane);
        float4 x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = lds_block(rx, k);
        const float rm = known_row_max != nullptr ? (rr == 0 ? known_row_max->x : known_row_max->y) : row_max(x);
        // packed float32x2 arithmetic (FADD2 / FMUL2 / FFMA2, sm_100): half the issue slots of the scalar forms
        const float2 nrm = make_float2(-rm, -rm), l2e = make_float2(kLog2e, kLog2e);
        float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 ta = __fmul2_rn(__fadd2_rn(make_float2(x[k].x, x[k].y), nrm), l2e);
          const float2 tb = __fmul2_rn(__fadd2_rn(make_float2(x[k].z, x[k].w), nrm), l2e);
          x[k].x = ex2_approx(ta.x);
          x[k].y = ex2_approx(ta.y);
          x[k].z = ex2_approx(tb.x);
          x[k].w = ex2_approx(tb.y);
          sa = __fadd2_rn(sa, make_float2(x[k].x, x[k].y));
          sb = __fadd2_rn(sb, make_float2(x[k].z, x[k].w));
        }
        const float s = (sa.x + sa.y) + (sb.x + sb.y);
        const float r = __frcp_rn(s);
        const float2 r2 = make_float2(r, r);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 ea = make_float2(x[k].x, x[k].y), eb = make_float2(x[k].z, x[k].w);
          // e * (1 / s): within 1.5 ulp of the correctly rounded quotient (the residual-corrected form of round 1 cost 4 of the
          // 22 packed instructions per float4 and changed nothing at the 2e-6 contract of this score)
          const float2 qa = __fmul2_rn(ea, r2), qb = __fmul2_rn(eb, r2);
          gmin = min3(min3(qa.x, qa.y, qb.x), qb.y, gmin);
          sts_block(rx, k, make_float4(qa.x, qa.y, qb.x, qb.y));
        }
      }
      __syncwarp();
    }
    const int rbase = half * 30;
    float h[4][5], xs[4][3];
    uint32_t bits[4] = {0u, 0u, 0u, 0u};
    // The scan: row rbase + t enters the rings (slot t % 5 of the horizontal 5-maxima, slot t % 3 of the raw values); from
    // t = 4 on, row rbase + t - 2 has its five window rows in the ring and its four columns are tested.  The ring slots
    // must be compile-time (registers): steps 0..3 only fill, steps 4..33 run as 2 x 15 rolled.  The per-line instruction
    // counts of round 2 (profiles/r2u_*) showed 42 instructions per step for 30 useful ones: a run-time "t >= 4" branch
    // with its reconvergence pair, 1 << (t - 4) rebuilt every step, and SEL + LOP3 instead of a predicated LOP3 -- hence
    // the peeled head, the carried `bit` and or_if_equal.
    auto fill = [&](const int t, const int s5, const int s3) {
      const float4 x = p4[(rbase + t) * 16 + l16];
      const float m01 = fmaxf(x.x, x.y), m23 = fmaxf(x.z, x.w);
      if (kMode == 0) {
        gmin = min3(min3(x.x, x.y, x.z), x.w, gmin);
        gmax = max3(m01, m23, gmax);
      }
      // neighbours inside the 16-lane row segment; the edge lanes get their own values back, which only ever reach the
      // windows of border columns (masked below) or are members of the window anyway
      const float Lm = __shfl_up_sync(kFull, m23, 1, 16), Lc3 = __shfl_up_sync(kFull, x.w, 1, 16);
      const float Rc0 = __shfl_down_sync(kFull, x.x, 1, 16), Rm = __shfl_down_sync(kFull, m01, 1, 16);
      h[0][s5] = max3(Lm, m01, x.z);
      h[1][s5] = max3(Lc3, m01, m23);
      h[2][s5] = max3(m01, m23, Rc0);
      h[3][s5] = max3(x.y, m23, Rm);
      xs[0][s3] = x.x; xs[1][s3] = x.y; xs[2][s3] = x.z; xs[3][s3] = x.w;
    };
    auto test = [&](const int c3, const uint32_t bit) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = max3(max3(h[k][0], h[k][1], h[k][2]), h[k][3], h[k][4]);
        or_if_equal(bits[k], xs[k][c3], w, bit);
      }
    };
#pragma unroll
    for (int t = 0; t < 4; ++t) fill(t, t % 5, t % 3);
    {
      uint32_t bit = 1u;  // bit t - 4 <-> row rbase + t - 2
#pragma unroll 1
      for (int tb = 4; tb < 34; tb += 15) {
#pragma unroll
        for (int u = 0; u < 15; ++u) {  // t = tb + u with 15 | tb - 4
          fill(tb + u, (4 + u) % 5, (4 + u) % 3);
          test((2 + u) % 3, bit);
          bit += bit;
        }
      }
    }
    if (l16 == 0) bits[0] = bits[1] = 0u;   // columns 0, 1
    if (l16 == 15) bits[2] = bits[3] = 0u;  // columns 62, 63
    gmin = warp_min_f(gmin);
    const float* col = map + (rbase + 2) * kMapDim + l16 * 4;  // bit i of bits[k] <-> col[i * 64 + k]
    {
      // do any two peak bits touch (8-neighbourhood)?  Inside a lane, across the neighbouring lane, across the two halves
      // (rows 31 | 32).  Conservative: bits at the map minimum still count here, the slow path filters them first.
      uint32_t touch = 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) touch |= bits[k] & (bits[k] << 1);
#pragma unroll
      for (int k = 0; k < 3; ++k) touch |= bits[k] & (bits[k + 1] | (bits[k + 1] << 1) | (bits[k + 1] >> 1));
      uint32_t nb = __shfl_down_sync(kFull, bits[0], 1, 16);
      if (l16 == 15) nb = 0u;
      touch |= bits[3] & (nb | (nb << 1) | (nb >> 1));
      uint32_t edge[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) edge[k] = __ballot_sync(kFull, half ? (bits[k] & 1u) : ((bits[k] >> 29) & 1u));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t lo = edge[k] & 0xffffu, hi = edge[k] >> 16;  // row 31 / row 32, column 4 l + k
        const uint32_t lo_r = (k < 3) ? (edge[k + 1] & 0xffffu) : ((edge[0] & 0xffffu) >> 1);  // column + 1
        const uint32_t hi_r = (k < 3) ? (edge[k + 1] >> 16) : ((edge[0] >> 16) >> 1);
        touch |= (lo & hi) | (lo & hi_r) | (hi & lo_r);
      }
      if (__any_sync(kFull, touch != 0u)) {
        const uint4 q = prune_touching_peaks(bits[0], bits[1], bits[2], bits[3], col, gmin, lane);
        bits[0] = q.x; bits[1] = q.y; bits[2] = q.z; bits[3] = q.w;
      }
    }
    if (kMode == 0) {
      constexpr float kLog2e = 1.4426950408889634f;
      gmax = warp_max_f(gmax);
      // Walk of the set bits, one bit of each of the four column masks per trip, branch-free inside the trip (round 2: the
      // guarded form spent 25 of its ~90 instructions per trip on branch / reconvergence pairs and register moves): an
      // exhausted mask yields row -1 of `col` (inside the map: rbase + 1 >= 1), which the predicate discards.
      const uint32_t col_s = (uint32_t)__cvta_generic_to_shared(col);
      float S = 0.f, T = 0.f;
      int n = 0;
      {
        uint32_t b[4] = {bits[0], bits[1], bits[2], bits[3]};
        while (b[0] | b[1] | b[2] | b[3]) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool any = b[k] != 0u;
            const int i = pop_highest_bit(b[k]);
            const float v = lds_f32(col_s + 4u * k + (uint32_t)(i * (kMapDim * 4)));
            if (any && v > gmin) {  // image > image.min(): peaks sitting at the map minimum are not peaks
              const float d = v - gmax;
              const float w = ex2_approx(d * kLog2e);
              S += w;
              T = fmaf(d, w, T);
              ++n;
            }
          }
        }
      }
      n = __reduce_add_sync(kFull, n);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(kFull, S, o);
        T += __shfl_xor_sync(kFull, T, o);
      }
      if (n > 0 && !(S >= 1e-30f)) {  // every term underflowed: shift by the largest peak instead (never on real maps)
        float M = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t b = bits[k];
          while (b) {
            const int i = __ffs(b) - 1;
            b &= b - 1;
            const float v = col[i * kMapDim + k];
            if (v > gmin) M = fmaxf(M, v);
          }
        }
        M = warp_max_f(M);
        S = 0.f;
        T = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t b = bits[k];
          while (b) {
            const int i = __ffs(b) - 1;
            b &= b - 1;
            const float v = col[i * kMapDim + k];
            if (v > gmin) {
              const float d = v - M;
              const float w = expf(d);
              S += w;
              T = fmaf(d, w, T);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          S += __shfl_xor_sync(kFull, S, o);
          T += __shfl_xor_sync(kFull, T, o);
        }
      }
      // H = -sum p log p with p = e^(v - c) / S  =  log S - T / S ; no peak: the reference sums an empty list
      if (lane == 0) a.out[m] = (n > 0) ? logf(S) - T / S : 0.f;
    } else {
      float t1 = -INFINITY, t2 = -INFINITY;
      int n = 0;
      {
        const uint32_t col_s = (uint32_t)__cvta_generic_to_shared(col);  // as in the MPE walk above
        uint32_t b[4] = {bits[0], bits[1], bits[2], bits[3]};
        while (b[0] | b[1] | b[2] | b[3]) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool any = b[k] != 0u;
            const int i = pop_highest_bit(b[k]);
            const float v = lds_f32(col_s + 4u * k + (uint32_t)(i * (kMapDim * 4)));
            if (any && v > gmin) {
              ++n;
              t2 = fmaxf(t2, fminf(t1, v));  // (t1, t2) = the two largest of {t1, t2, v}
              t1 = fmaxf(t1, v);
            }
          }
        }
      }
      n = __reduce_add_sync(kFull, n);
      // two largest values over all lanes: the maximum, then either the maximum again (held by two lanes, or twice by
      // one lane: then that lane's t2 equals it) or the largest remaining value
      const float top = warp_max_f(t1);
      const bool mine = (t1 == top);
      const int owners = __popc(__ballot_sync(kFull, mine));
      float second = warp_max_f(mine ? t2 : t1);
      if (owners >= 2) second = top;
      // fewer than two peaks: the reference raises IndexError (strategy.py:1208); NaN here, raised by the host wrapper
      if (lane == 0) a.out[m] = (n >= 2) ? fabsf(top - second) : __int_as_float(0x7fc00000);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Reprojection-XE term of one map (utils/triangulation.py:236-257): sum((heatmap - render)^2) / (H W) with
// render[y][x] = exp(-((x - u)^2 + (y - v)^2) / (2 sigma^2)) in float64, (u, v) the reprojection of the triangulated
// joint in IMAGE pixels (the reference compares it with the heat-map pixel grid as is).  The Gaussian is separable:
// the warp forms gx[0..63], gy[0..63] (4 double exponentials per lane) in its scratch, a lane keeps the gx of its four
// columns in registers and a pixel costs one conversion and two DFMA (r = h - gx gy; acc += r r).
// ---------------------------------------------------------------------------------------------------------------
struct XeOp {
  struct Args {
    const double* proj;
    const double* xyz;
    int V, J;
    double inv_two_sigma2;
    double* out_map;
  };
  // The projection matrix row block and the 3-D joint of a map come from global memory; they are fetched one map ahead
  // (while the warp still waits for / works on the current stage) so that their latency never sits between the
  // arrival of a map and its evaluation.
  struct Pre {
    double P[12], X[3];
  };
  static constexpr bool kWritesSmem = false;
  static constexpr int kProducerBackoff = 0;
  __device__ static __forceinline__ Pre prefetch(int64_t m, const Args& a) {
    Pre q;
    const int j = (int)(m % a.J);
    const int64_t fv = m / a.J;
    const double* __restrict__ P = a.proj + fv * 12;
    const double* __restrict__ X = a.xyz + ((fv / a.V) * a.J + j) * 3;
#pragma unroll
    for (int i = 0; i < 12; ++i) q.P[i] = __ldg(P + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) q.X[i] = __ldg(X + i);
    return q;
  }
  __device__ static __forceinline__ void run(float* map, int64_t m, bool ok, int lane, const Args& a, unsigned char* scratch,
                                             const Pre& q) {
    (void)ok;
    double* gx = reinterpret_cast<double*>(scratch);
    double* gy = gx + kMapDim;
    const double* P = q.P;
    const double x = q.X[0], y = q.X[1], z = q.X[2];
    // [X, 1] @ P^T (:476), then dehomogenise with w == 0 -> 1 (:397-399)
    const double pu = ((x * P[0] + y * P[1]) + z * P[2]) + P[3];
    const double pv = ((x * P[4] + y * P[5]) + z * P[6]) + P[7];
    double pw = ((x * P[8] + y * P[9]) + z * P[10]) + P[11];
    if (pw == 0.0) pw = 1.0;
    const double u = pu / pw, v = pv / pw;
    __syncwarp();
#pragma unroll
    for (int i = lane; i < kMapDim; i += 32) {
      const double dx = (double)i - u, dy = (double)i - v;
      gx[i] = exp(-(dx * dx) * a.inv_two_sigma2);
      gy[i] = exp(-(dy * dy) * a.inv_two_sigma2);
    }
    __syncwarp();
    const int half = lane >> 4, l16 = lane & 15;
    const double2 ga = reinterpret_cast<const double2*>(gx)[l16 * 2], gb = reinterpret_cast<const double2*>(gx)[l16 * 2 + 1];
    const float4* __restrict__ p = reinterpret_cast<const float4*>(map) + lane;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll 8
    for (int t = 0; t < 32; ++t) {
      const float4 h = p[t * 32];
      const double g = gy[2 * t + half];
      const double r0 = fma(-ga.x, g, (double)h.x), r1 = fma(-ga.y, g, (double)h.y);
      const double r2 = fma(-gb.x, g, (double)h.z), r3 = fma(-gb.y, g, (double)h.w);
      acc0 = fma(r0, r0, acc0);
      acc1 = fma(r1, r1, acc1);
      acc2 = fma(r2, r2, acc2);
      acc3 = fma(r3, r3, acc3);
    }
    const double acc = warp_sum((acc0 + acc1) + (acc2 + acc3));
    if (lane == 0) a.out_map[m] = acc * (1.0 / (double)kMapFloats);
  }
};

}  // namespace mval
