// Fused persistent pool-scoring kernel (sm_100a): heat-map decode + RANSAC/DLT triangulation + per-frame
// uncertainty in ONE launch, so that the float64 Jacobi work (FP64-pipe bound) runs underneath the heat-map
// stream (HBM bound) instead of after it, and the key-points never leave the SM.
//
// One CTA per SM, persistent over frames  f = blockIdx.x, blockIdx.x + gridDim.x, ...  Warp roles:
//
//   warp 0           TMA producer: one elected lane issues cp.async.bulk (global -> shared, mbarrier complete_tx)
//                    for every 16 KiB map of the frame into a ring of S stages (S*16 KiB in flight per SM, no
//                    registers spent on staging), plus the frame's V*96 B of projection matrices.
//   warps 1..6       decode: wait full[stage], arg-max the map out of shared memory (conflict-free LDS.128, same
//                    monotone-key / REDUX reduction as decode_argmax_kernel), release empty[stage], publish the
//                    key-point into the frame slot and arrive on kp_ready[slot].
//   warps 7..14      RANSAC: the (frame, joint) vote tasks form one stream t = i*J + j dealt round-robin to the 8
//                    warps (balanced for any J); lane = view pair as in ransac_vote_kernel.  The warp i % 8 then
//                    runs the final solves of frame i with lane = joint, reduces metric / inlier_count and frees
//                    the slot.
//
// Up to four frame slots (key-points, P, masks) decouple the roles: decode may run up to slots - 1 frames ahead of the
// finals.  A slot costs V * (96 + 8 J) bytes of shared memory, so for large rigs (20 views x 42 joints) the launcher
// trades slots for ring stages -- bytes in flight are what keeps the HBM stream saturated.
// Every hand-off is an mbarrier in shared memory; there is no __syncthreads after set-up.
//
// Arithmetic is shared with the stand-alone kernels (ransac.cuh), so the results are bit-identical to
// mval_decode_argmax + mval_triangulate_ransac (tests/test_gpu_parity.py::test_fused_equals_unfused).
//
// Scored variants (kScore = 1 HP, 2 MPE, 3 BSB; strategy.py:1149-1215): _compute_sal_dict needs BOTH the triangulation
// of every frame (sal_metric, pred_3d_keypoints; strategy.py:1037-1063) and, for those strategies, a per-map score of
// the same heat maps (strategy.py:1072-1094).  As two launches every heat-map byte crosses HBM twice.  Here the decode
// warps evaluate the map_stream Op (mapops.cuh, the very same device code, hence bit-identical scores) on the staged
// map right after its arg-max and only then hand the stage back, so the pool is read once.  The per-map work grows
// from ~400 to 1 800-2 500 warp instructions, so these variants run 12 decode warps (one per ring stage, as
// map_stream_kernel does) over 6 RANSAC warps.
#include <stdlib.h>

#include "mapops.cuh"
#include "ransac.cuh"

namespace mval {

// Warp budget and per-map evaluator of each variant.  The register file is split per scheduler (16 384 registers
// each), so 17-20 warps per CTA cap a thread at 96 registers and 13-16 warps at 128.  Round 1f, per 16 384 frames
// (profiles/r1f_summary.md): 12 decode + 6 RANSAC warps at 96 registers (ptxas spills) 10.1 / 11.0 / 18.5 ms for HP /
// MPE / BSB, 10 + 5 warps at 128 registers 9.4 / 10.6 / 13.9 ms, 12 + 3 warps at 128 registers 8.9 / 10.2 / 13.1 ms.
// Round 2, after the decode warps' instruction cuts (profiles/r2_summary.md section 5): HP 12 + 3: 7.23 ms, 11 + 4:
// 5.76 ms, 10 + 5: 6.66 ms -- HP's row sweep is short enough now that three RANSAC warps no longer keep up, so HP runs
// 11 + 4 (kShape 1); MPE / BSB show no such preference (9.3-10.1 ms for all three splits) and keep 12 + 3.
// kShape s = (12 - s) decode + (3 + s) RANSAC warps; MVAL_FUSED_SHAPE = 0 / 1 forces one split (A/B measurements).
template <int kScore, int kShape = 0> struct FusedCfg;
template <> struct FusedCfg<MVAL_MAP_SCORE_NONE, 0> { static constexpr int kD = 6, kR = 8; using Op = void; };       // 480 threads
template <int kShape> struct FusedCfg<MVAL_MAP_SCORE_HP, kShape> { static constexpr int kD = 12 - kShape, kR = 3 + kShape; using Op = HpOp; };  // 512 threads
template <int kShape> struct FusedCfg<MVAL_MAP_SCORE_MPE, kShape> { static constexpr int kD = 12 - kShape, kR = 3 + kShape; using Op = PeaksOp<0>; };
template <int kShape> struct FusedCfg<MVAL_MAP_SCORE_BSB, kShape> { static constexpr int kD = 12 - kShape, kR = 3 + kShape; using Op = PeaksOp<1>; };
// Heat maps of one launch: up to MVAL_MAX_SEGMENTS device buffers of whole frames, frame f of the launch lives in segment
// s with start[s] <= f < start[s + 1] at ptr[s] + (f - start[s]) * V * J * H * W (mval_score_pool_segments: a pool is
// usually produced batch by batch by the pose estimator and need not be contiguous).  Passed by value as a
// __grid_constant__ parameter: only the producer lane reads it, from the constant bank.
struct SegTable {
  const float* ptr[MVAL_MAX_SEGMENTS];
  int64_t start[MVAL_MAX_SEGMENTS + 1];
  int n;
};

constexpr int kMaxFrameSlots = 4;  // frame slots per CTA: 4 when they are small, fewer when V * J is large (see launcher)
constexpr int kMaxStages = 64;

// Abort record of the mbarrier watchdog (tma.cuh): [0] flag, [1] code, [2] block, [3] warp, [4] frame iter, [5] index
__device__ unsigned long long g_fused_abort[kWdWords];
static WatchdogHost g_fused_watchdog;

int fused_watchdog_poll() { return g_fused_watchdog.poll(g_fused_abort, "score_pool_fused"); }
int fused_watchdog_debug(unsigned long long timeout_cycles, int stall) {
  if (int rc = g_fused_watchdog.prepare(g_fused_abort, "score_pool_fused")) return rc;
  return g_fused_watchdog.debug_set(g_fused_abort, timeout_cycles, stall);
}

struct FusedSmem {  // byte offsets into dynamic shared memory
  uint32_t ring, proj, kp, mask, red_reproj, red_inl, pair, perm, pxy, bars, total;
  int stages, slots;
  uint32_t stage_bytes;
};

__host__ __device__ inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline FusedSmem fused_layout(int V, int J, int HW, int stages, int slots, int ransac_warps) {
  FusedSmem L;
  const int n_all = V * (V - 1) / 2;
  L.stage_bytes = (uint32_t)HW * 4u;
  L.stages = stages;
  L.slots = slots;
  uint32_t o = 0;
  L.ring = o;        o = align_up(o + (uint32_t)stages * L.stage_bytes, 128);
  L.proj = o;        o = align_up(o + (uint32_t)slots * V * 96u, 16);
  L.kp = o;          o = align_up(o + (uint32_t)slots * V * J * 8u, 16);
  L.mask = o;        o = align_up(o + (uint32_t)slots * J * 4u, 16);
  L.red_reproj = o;  o = align_up(o + (uint32_t)ransac_warps * J * 8u, 16);
  L.red_inl = o;     o = align_up(o + (uint32_t)ransac_warps * J * 4u, 16);
  L.pair = o;        o = align_up(o + 2u * n_all, 16);
  L.perm = o;        o = align_up(o + (uint32_t)ransac_warps * n_all * 2u, 16);
  L.pxy = o;         o = align_up(o + (uint32_t)ransac_warps * 2u * V * 8u, 16);
  L.bars = o;        o = o + (2u * stages + 3u * (uint32_t)slots) * 8u;
  L.total = o + kMapAlign;  // slack for aligning the ring (offset 0) to kMapAlign inside the kernel
  return L;
}

// kRowArgmax: 64 x 64 maps are arg-maxed by the lane = row sweep of mapops.cuh (warp_argmax_map64) instead of the
// generic per-vector scan; both are the same function of the map.
template <int kScore, bool kRowArgmax, int kShape = 0>
__global__ void __launch_bounds__(kWarp * (1 + FusedCfg<kScore, kShape>::kD + FusedCfg<kScore, kShape>::kR), 1)
score_pool_fused_kernel(const __grid_constant__ SegTable segs, const double* __restrict__ proj, const uint8_t* __restrict__ valid,
                        int64_t n_frames, int V, int J, int H, int HW, int stride, int stages, int slots, int n_iters, double eps,
                        uint64_t seed, int64_t frame_offset, const int64_t* __restrict__ frame_keys, int32_t* __restrict__ out_xy,
                        double* __restrict__ out_xyz,
                        double* __restrict__ out_reproj, int32_t* __restrict__ out_inliers, double* __restrict__ out_metric,
                        int32_t* __restrict__ out_inlier_count, float* __restrict__ out_map_score) {
  constexpr int kFusedDecodeWarps = FusedCfg<kScore, kShape>::kD;
  constexpr int kFusedRansacWarps = FusedCfg<kScore, kShape>::kR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // the staged maps sit on kMapAlign boundaries (mapops.cuh: load_row_rotated); L.total holds the slack
  unsigned char* smem = smem_raw + ((kMapAlign - (smem_u32(smem_raw) & (kMapAlign - 1u))) & (kMapAlign - 1u));
  const FusedSmem L = fused_layout(V, J, HW, stages, slots, kFusedRansacWarps);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + stages;
  uint64_t* kp_ready = empty + stages;
  uint64_t* kp_free = kp_ready + slots;
  uint64_t* masks_ready = kp_free + slots;
  const int VJ = V * J;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_all = V * (V - 1) / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < slots; ++s) {
      mbar_init(&kp_ready[s], (uint32_t)VJ + 1u);
      mbar_init(&kp_free[s], (uint32_t)J + 1u);
      mbar_init(&masks_ready[s], (uint32_t)J);
    }
    mbar_fence_init();
  }
  build_pair_table(smem + L.pair, V, threadIdx.x, blockDim.x);
  __syncthreads();

  // frames of this CTA: blockIdx.x + i * gridDim.x
  const int64_t nf = (n_frames > (int64_t)blockIdx.x) ? (n_frames - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------- producer
    if (lane == 0 && !watchdog_stalled(g_fused_abort)) {
      int64_t c = 0;
      int seg = 0;
      for (int64_t i = 0; i < nf; ++i) {
        const int64_t frame = blockIdx.x + i * (int64_t)gridDim.x;
        while (seg + 1 < segs.n && frame >= segs.start[seg + 1]) ++seg;  // frames only grow
        const int sl = (int)(i % slots);
        const uint32_t ku = (uint32_t)(i / slots);
        if (!mbar_wait(&kp_free[sl], (ku & 1u) ^ 1u, g_fused_abort, 1, i, sl)) return;
        mbar_arrive_expect_tx(&kp_ready[sl], (uint32_t)V * 96u);
        bulk_g2s(smem + L.proj + sl * V * 96, proj + frame * V * 12, (uint32_t)V * 96u, &kp_ready[sl]);
        const float* src = segs.ptr[seg] + (frame - segs.start[seg]) * (int64_t)VJ * HW;
        for (int m = 0; m < VJ; ++m, ++c) {
          const int st = (int)(c % stages);
          const uint32_t kf = (uint32_t)(c / stages);
          // (the MPE / BSB variants are issue-bound: the producer waits here nearly always and must not spin)
          constexpr int kProducerBackoff = (kScore == MVAL_MAP_SCORE_MPE || kScore == MVAL_MAP_SCORE_BSB) ? 2 : 0;
          if (!mbar_wait<kProducerBackoff>(&empty[st], (kf & 1u) ^ 1u, g_fused_abort, 2, i, st)) return;
          mbar_arrive_expect_tx(&full[st], L.stage_bytes);
          bulk_g2s(smem + L.ring + (uint32_t)st * L.stage_bytes, src + (int64_t)m * HW, L.stage_bytes, &full[st]);
        }
      }
    }
  } else if (warp <= kFusedDecodeWarps) {
    // ------------------------------------------------------------------------------------------- decode
    const int d = warp - 1;
    const int hw4 = HW >> 2;
    int2* kp_all = reinterpret_cast<int2*>(smem + L.kp);
    for (int64_t i = 0; i < nf; ++i) {
      const int64_t frame = blockIdx.x + i * (int64_t)gridDim.x;
      const int sl = (int)(i % slots);
      const uint32_t ku = (uint32_t)(i / slots);
      // every decode warp passes every frame's slot gate (even with no map in it) so that no warp can run a full
      // barrier phase ahead of the others
      if (!mbar_wait(&kp_free[sl], (ku & 1u) ^ 1u, g_fused_abort, 3, i, sl)) return;
      const int64_t c0 = i * VJ;
      // first map of this frame owned by this warp: smallest m with (c0 + m) % D == d
      int m = (int)(((int64_t)d - c0 % kFusedDecodeWarps + kFusedDecodeWarps) % kFusedDecodeWarps);
      for (; m < VJ; m += kFusedDecodeWarps) {
        const int64_t c = c0 + m;
        const int st = (int)(c % stages);
        const uint32_t kf = (uint32_t)(c / stages);
        if (!mbar_wait(&full[st], kf & 1u, g_fused_abort, 4, i, st)) return;
        float* stage = reinterpret_cast<float*>(smem + L.ring + (uint32_t)st * L.stage_bytes);
        const float4* p = reinterpret_cast<const float4*>(stage);
        const bool ok = valid == nullptr || valid[frame * J + m % J] != 0;
        uint32_t idx = 0u;
        if constexpr (kScore == MVAL_MAP_SCORE_HP) {
          // HP's row sweep knows every row maximum: score and arg-max come out of one pass over the stage
          if (ok)
            idx = HpOp::eval<true>(stage, frame * VJ + m, lane, HpOp::Args{out_map_score});
          else if (lane == 0)
            out_map_score[frame * VJ + m] = __int_as_float(0x7fc00000);
        } else if constexpr (kRowArgmax) {
          idx = warp_argmax_map64(stage, lane);
        } else {
          idx = warp_argmax_map<8, true>([&](int q) { return p[q]; }, hw4, lane, nullptr);
        }
        constexpr bool kScoreAfter = kScore == MVAL_MAP_SCORE_MPE || kScore == MVAL_MAP_SCORE_BSB;
        if (lane == 0) {
          // every lane's loads were consumed by the reductions above
          if (!kScoreAfter) mbar_arrive(&empty[st]);
          int2 xy = make_int2((int)(idx % (uint32_t)H) * stride, (int)(idx / (uint32_t)H) * stride);
          if (!ok) xy = make_int2(0, 0);  // evaluation.py:21-23
          kp_all[sl * VJ + m] = xy;
          if (out_xy) reinterpret_cast<int2*>(out_xy)[frame * VJ + m] = xy;
          mbar_arrive(&kp_ready[sl]);  // the RANSAC warps may start on this key-point while the score is evaluated
        }
        if constexpr (kScoreAfter) {
          using Op = typename FusedCfg<kScore, kShape>::Op;
          __syncwarp();
          Op::run(stage, frame * VJ + m, ok, lane, typename Op::Args{out_map_score}, nullptr, typename Op::Pre{});
          if (Op::kWritesSmem) fence_proxy_async_smem();  // BSB rewrote the stage; order that before the TMA refill
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[st]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------- RANSAC
    const int w = warp - 1 - kFusedDecodeWarps;
    const bool subset = n_all > n_iters;
    const int n_pairs = subset ? n_iters : n_all;
    const uint8_t* sPair = smem + L.pair;
    uint16_t* perm = reinterpret_cast<uint16_t*>(smem + L.perm) + w * n_all;
    double* px = reinterpret_cast<double*>(smem + L.pxy) + w * 2 * V;
    double* py = px + V;
    double* red_reproj = reinterpret_cast<double*>(smem + L.red_reproj) + w * J;
    int32_t* red_inl = reinterpret_cast<int32_t*>(smem + L.red_inl) + w * J;
    const int2* kp_all = reinterpret_cast<const int2*>(smem + L.kp);
    uint32_t* mask_all = reinterpret_cast<uint32_t*>(smem + L.mask);
    for (int64_t i = 0; i < nf; ++i) {
      const int64_t frame = blockIdx.x + i * (int64_t)gridDim.x;
      const int sl = (int)(i % slots);
      const uint32_t ku = (uint32_t)(i / slots);
      const double* P = reinterpret_cast<const double*>(smem + L.proj + sl * V * 96);
      const int2* kp = kp_all + sl * VJ;
      uint32_t* masks = mask_all + sl * J;
      if (!mbar_wait<true>(&kp_ready[sl], ku & 1u, g_fused_abort, 5, i, sl)) return;  // every RANSAC warp, every frame (same reason)
      const int64_t t0 = i * J;
      int j = (int)(((int64_t)w - t0 % kFusedRansacWarps + kFusedRansacWarps) % kFusedRansacWarps);
      for (; j < J; j += kFusedRansacWarps) {
        uint32_t mask = 0u;
        if (valid == nullptr || valid[frame * J + j] != 0) {
          __syncwarp();
          for (int v = lane; v < V; v += kWarp) {
            const int2 q = kp[v * J + j];
            px[v] = (double)q.x;
            py[v] = (double)q.y;
          }
          if (subset) draw_pair_subset(perm, n_all, n_iters, seed, frame_keys ? frame_keys[frame] : frame_offset + frame, j, lane);
          __syncwarp();
          mask = ransac_vote_warp(
              P, px, py, V, n_pairs, eps,
              [&](int pi, int& a, int& b) {
                const int li = subset ? (int)perm[pi] : pi;
                a = sPair[2 * li];
                b = sPair[2 * li + 1];
              },
              lane);
        }
        if (lane == 0) {
          masks[j] = mask;
          mbar_arrive(&masks_ready[sl]);
          mbar_arrive(&kp_free[sl]);
        }
      }
      if ((int)(i % kFusedRansacWarps) == w) {
        // final solves of this frame: lane = joint
        if (!mbar_wait<true>(&masks_ready[sl], ku & 1u, g_fused_abort, 7, i, sl)) return;
        for (int jb = 0; jb < J; jb += kWarp) {
          const int jj = jb + lane;
          if (jj < J) {
            const int64_t task = frame * J + jj;
            const uint32_t mask = masks[jj];
            if (valid != nullptr && valid[task] == 0) {
              out_xyz[3 * task] = 0.0;
              out_xyz[3 * task + 1] = 0.0;
              out_xyz[3 * task + 2] = 0.0;
              if (out_reproj) out_reproj[task] = __longlong_as_double(0x7ff8000000000000ll);
              if (out_inliers) out_inliers[task] = 0;
              red_inl[jj] = -1;
            } else {
              double X, Y, Z, rm;
              ransac_final_thread(
                  P,
                  [&](int v, double& x, double& y) {
                    const int2 q = kp[v * J + jj];
                    x = (double)q.x;
                    y = (double)q.y;
                  },
                  mask, V, X, Y, Z, rm);
              out_xyz[3 * task] = X;
              out_xyz[3 * task + 1] = Y;
              out_xyz[3 * task + 2] = Z;
              if (out_reproj) out_reproj[task] = rm;
              if (out_inliers) out_inliers[task] = __popc(mask);
              red_reproj[jj] = rm;
              red_inl[jj] = __popc(mask);
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&kp_free[sl]);  // slot inputs are no longer needed
          double sum = 0.0;           // same sequential order as frame_reduce_kernel
          int cnt = 0, mn = 0x7fffffff;
          for (int q = 0; q < J; ++q) {
            if (red_inl[q] >= 0) {
              sum += red_reproj[q];
              mn = min(mn, red_inl[q]);
              ++cnt;
            }
          }
          out_metric[frame] = cnt ? sum / (double)cnt : __longlong_as_double(0x7ff8000000000000ll);
          out_inlier_count[frame] = cnt ? mn : 0;
        }
        __syncwarp();
      }
    }
  }
}

// Returns MVAL_ERR_UNSUPPORTED (without setting an error) when the shape does not fit the fused kernel; the caller
// then takes the multi-launch path.
template <int kScore, bool kRowArgmax, int kShape = 0>
static int launch_fused_variant(const SegTable& segs, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J,
                                int H, int W, int stride, const mval_ransac_params& prm, int32_t* out_xy, double* out_xyz,
                                double* out_reproj, int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count,
                                float* out_map_score, cudaStream_t stream) {
  constexpr int kD = FusedCfg<kScore, kShape>::kD, kR = FusedCfg<kScore, kShape>::kR;
  const int HW = H * W;
  if (HW % 4 != 0 || (reinterpret_cast<uintptr_t>(proj) & 15) != 0 || prm.pairs != nullptr) return MVAL_ERR_UNSUPPORTED;
  for (int s = 0; s < segs.n; ++s)
    if ((reinterpret_cast<uintptr_t>(segs.ptr[s]) & 15) != 0) return MVAL_ERR_UNSUPPORTED;
  if (kScore != MVAL_MAP_SCORE_NONE && (H != kMapDim || W != kMapDim)) return MVAL_ERR_UNSUPPORTED;  // Ops are 64 x 64 only
  int dev = 0, max_smem = 0;
  MVAL_CUDA(cudaGetDevice(&dev));
  MVAL_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // The ring depth must be a multiple of the number of decode warps: map c lives in stage c % S and is decoded by warp
  // c % D, so D | S pins every stage to one warp, which visits it in order.  (Otherwise a warp can wait on a phase two
  // ahead of the barrier and the parity test aliases with the phase before -- seen as stale reads / deadlock.)
  // Deepest ring first (bytes in flight), then as many frame slots as still fit: 4 slots / 12 stages at 8 views x 19
  // joints, 2 slots / 12 stages at 20 x 42 (4 slots would leave 6 stages and the stream latency-bound).
  int stages = 0, slots = 0;
  for (int sl = kMaxFrameSlots; sl >= 2; --sl) {
    int st = kMaxStages / kD * kD;
    while (st >= kD && fused_layout(V, J, HW, st, sl, kR).total > (uint32_t)max_smem) st -= kD;
    if (st > stages) { stages = st; slots = sl; }
  }
  if (stages < kD) return MVAL_ERR_UNSUPPORTED;
  const FusedSmem L = fused_layout(V, J, HW, stages, slots, kR);
  MVAL_CUDA(cudaFuncSetAttribute(score_pool_fused_kernel<kScore, kRowArgmax, kShape>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  const int64_t sms = num_sms();
  const int grid = (int)(n_frames < sms ? n_frames : sms);
  // a watchdog trip of an EARLIER launch surfaces here (or in mval_check_async), never silently
  if (int rc = g_fused_watchdog.prepare(g_fused_abort, "score_pool_fused")) return rc;
  score_pool_fused_kernel<kScore, kRowArgmax, kShape><<<grid, kWarp * (1 + kD + kR), L.total, stream>>>(
      segs, proj, valid, n_frames, V, J, H, HW, stride, stages, slots, prm.n_iters, prm.epsilon, prm.pair_seed, prm.frame_offset,
      prm.frame_keys, out_xy, out_xyz, out_reproj, out_inliers, out_metric, out_inlier_count, out_map_score);
  MVAL_LAUNCH_CHECK("score_pool_fused");
  return MVAL_OK;
}

int launch_score_pool_fused_segments(const float* const* seg_ptr, const int64_t* seg_frames, int n_segments,
                                     const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J, int H, int W,
                                     int stride, const mval_ransac_params& prm, int map_score, int32_t* out_xy, double* out_xyz,
                                     double* out_reproj, int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count,
                                     float* out_map_score, cudaStream_t stream) {
  if (n_segments < 1 || n_segments > MVAL_MAX_SEGMENTS) return MVAL_ERR_UNSUPPORTED;
  SegTable segs;
  segs.n = n_segments;
  segs.start[0] = 0;
  for (int s = 0; s < n_segments; ++s) {
    segs.ptr[s] = seg_ptr[s];
    segs.start[s + 1] = segs.start[s] + seg_frames[s];
  }
  for (int s = n_segments; s < MVAL_MAX_SEGMENTS; ++s) {
    segs.ptr[s] = nullptr;
    segs.start[s + 1] = segs.start[n_segments];
  }
  if (segs.start[n_segments] != n_frames) return MVAL_ERR_UNSUPPORTED;
#define MVAL_FUSED_ARGS                                                                                               \
  segs, proj, valid, n_frames, V, J, H, W, stride, prm, out_xy, out_xyz, out_reproj, out_inliers, out_metric,           \
      out_inlier_count, out_map_score, stream
  // Arg-max flavour on 64 x 64 maps.  The scored variants take the lane = row sweep (170 instead of 400 instructions per map:
  // they are bound by the SM).  The unscored pass is HBM-bound either way and the choice is about POWER: on short launches the
  // generic per-vector scan is faster (API batches of 8 192 frames, A/B/A/B on one box, profiles/r2q2_*: 52.3 / 52.9 ms per
  // 125 000-frame call against 54.8 / 54.7 ms; the 20- and 31-view rigs 1.04 / 1.08 against 0.98 / 1.00 of the copy peak), but
  // in a long launch the power cap sets the clocks and the sweep's smaller instruction count buys 6 % (100 000-frame step,
  // A/B/A/B, profiles/r2p2_*: 36.2-36.5 ms against 34.0-34.2 ms).  So: the sweep from kRowArgmaxMinMaps maps per launch on
  // (~12 ms at 8 x 19 maps per frame), the scan below.  MVAL_ROW_ARGMAX = 0 / 1 forces one flavour everywhere (A/B
  // measurements and tests; read on every call).
  constexpr int64_t kRowArgmaxMinMaps = 32768ll * 152ll;
  const char* env = getenv("MVAL_ROW_ARGMAX");
  const int force = (env != nullptr && (env[0] == '0' || env[0] == '1')) ? env[0] - '0' : -1;
  const bool is64 = H == kMapDim && W == kMapDim;
  // warp split of the scored variants (FusedCfg)
  const char* shape_env = getenv("MVAL_FUSED_SHAPE");
  const int shape = (shape_env != nullptr && (shape_env[0] == '0' || shape_env[0] == '1')) ? shape_env[0] - '0'
                                                                                           : (map_score == MVAL_MAP_SCORE_HP ? 1 : 0);
  switch (map_score) {
    case MVAL_MAP_SCORE_NONE:
      out_map_score = nullptr;
      if (is64 && (force == 1 || (force < 0 && n_frames * (int64_t)V * J >= kRowArgmaxMinMaps)))
        return launch_fused_variant<MVAL_MAP_SCORE_NONE, true>(MVAL_FUSED_ARGS);
      return launch_fused_variant<MVAL_MAP_SCORE_NONE, false>(MVAL_FUSED_ARGS);
    case MVAL_MAP_SCORE_HP:  // the arg-max is a by-product of HP's own row sweep
      if (shape == 0) return launch_fused_variant<MVAL_MAP_SCORE_HP, true, 0>(MVAL_FUSED_ARGS);
      return launch_fused_variant<MVAL_MAP_SCORE_HP, true, 1>(MVAL_FUSED_ARGS);
    case MVAL_MAP_SCORE_MPE:
      if (force == 0) return launch_fused_variant<MVAL_MAP_SCORE_MPE, false>(MVAL_FUSED_ARGS);
      if (shape == 1) return launch_fused_variant<MVAL_MAP_SCORE_MPE, true, 1>(MVAL_FUSED_ARGS);
      return launch_fused_variant<MVAL_MAP_SCORE_MPE, true>(MVAL_FUSED_ARGS);
    case MVAL_MAP_SCORE_BSB:
      if (force == 0) return launch_fused_variant<MVAL_MAP_SCORE_BSB, false>(MVAL_FUSED_ARGS);
      if (shape == 1) return launch_fused_variant<MVAL_MAP_SCORE_BSB, true, 1>(MVAL_FUSED_ARGS);
      return launch_fused_variant<MVAL_MAP_SCORE_BSB, true>(MVAL_FUSED_ARGS);
    default:
      return MVAL_ERR_UNSUPPORTED;
  }
#undef MVAL_FUSED_ARGS
}

int launch_score_pool_fused(const float* hm, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J,
                            int H, int W, int stride, const mval_ransac_params& prm, int map_score, int32_t* out_xy,
                            double* out_xyz, double* out_reproj, int32_t* out_inliers, double* out_metric,
                            int32_t* out_inlier_count, float* out_map_score, cudaStream_t stream) {
  return launch_score_pool_fused_segments(&hm, &n_frames, 1, proj, valid, n_frames, V, J, H, W, stride, prm, map_score, out_xy,
                                          out_xyz, out_reproj, out_inliers, out_metric, out_inlier_count, out_map_score,
                                          stream);
}

}  // namespace mval
