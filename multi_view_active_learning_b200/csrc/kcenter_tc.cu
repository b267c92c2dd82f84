// Coreset batched update on the 5th-generation tensor cores (sm_100a): a tcgen05 TF32 GEMM as an error-bounded SCREEN,
// followed by an exact float32 re-evaluation of the few (row, centre) pairs that survive it.
//
// Goal of a batched update (kcenter.cu header):  m_i <- min(m_i, min_t dist(x_i, c_t))  for T new centres, with dist in
// the canonical float32 order.  tcgen05 has no true-float32 MMA, so the tensor cores cannot produce dist itself; they
// can, however, PROVE for almost every pair that it cannot change m_i:
//
//   dot_tc(i,t)   = TF32 tensor-core dot product (operands truncated to 10 mantissa bits, float32 accumulation in TMEM)
//   |dot_tc - dot_canonical| <= E,  2E <= (2^-9 + d 2^-21) (|x|^2 + |c|^2)            (derivation in DESIGN.md section 4)
//   => canonical d2(i,t) lies within s_i = alpha (|x_i|^2 + max_t |c_t|^2) of  D~(i,t) = |x_i|^2 - 2 val,
//      val = dot_tc - |c_t|^2 / 2,  alpha = 2^-8 + d 2^-20 (twice the bound)
//
// A pair can lower m_i only if (A) D~ - s_i < m_i^2 (it may beat the current minimum) and (B) D~ - s_i <= min_t' D~ + s_i
// (it may be the best of this batch).  The epilogue evaluates A and B straight out of TMEM -- one sweep for the row
// maximum of val, one sweep that appends the surviving (i, t) to a global list -- and kc_recheck_kernel evaluates the
// canonical distance of the survivors (typically 1-3 per row while minima are still falling, ~0 later) and applies
// atomicMin.  The result is bit-identical to the exact FFMA pass; if the list overflows, a device-side flag makes the
// FFMA pass (always launched, normally a no-op) do the work instead.  No host synchronisation.
//
// Kernel anatomy (one persistent CTA per SM, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor 2-D tiles, 128-byte swizzle, 4-stage ring
//              A = 128 feature rows x 32 floats (16 KiB), B = 256 centres x 32 floats (32 KiB; rows >= T zero-filled)
//   warp 1     allocates 512 TMEM columns, issues tcgen05.mma.cta_group::1.kind::tf32 M=128 N=256 K=8 (4 per stage),
//              tcgen05.commit -> stage-empty / accumulator-full mbarriers
//   warps 2-5  epilogue: tcgen05.ld 32x32b of the 128 x 256 float32 accumulator (two TMEM buffers, so the MMAs of the next
//              row tile overlap the epilogue of this one)
#include <cuda.h>
#include <stdlib.h>

#include "kcenter.cuh"

namespace mval {

namespace {

constexpr int kTcBlockM = 128, kTcBlockN = 256, kTcBlockK = 32, kTcStages = 4, kTcUmmaK = 8;
constexpr int kTcThreads = 192;
constexpr uint32_t kTcABytes = kTcBlockM * kTcBlockK * 4, kTcBBytes = kTcBlockN * kTcBlockK * 4;
constexpr uint32_t kTcStageBytes = kTcABytes + kTcBBytes;
constexpr uint32_t kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /*hcc*/ + 256 /*barriers etc.*/ + 1024 /*alignment*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// A wait that lasts ~4 s of SM clocks is a protocol bug: trap (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  long long t0 = 0;
  for (uint32_t polls = 0;; ++polls) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((polls & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 8000000000ll) asm volatile("trap;");
    }
  }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, float32 accumulate, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major operand tile, rows of 128 bytes, 128-byte swizzle, 8-row groups 1024 bytes apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBlockN >> 3) << 17) | ((uint32_t)(kTcBlockM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

struct TcPair {
  uint32_t row, t;
};

// Appends the surviving (row, centre) pairs of one 32-centre chunk (lane = row, v[j] = accumulator of centre t0 + j) to the
// global list.  The survivors of the whole warp take ONE atomicAdd (round 1: one per pair, which serialised the early
// passes where millions of pairs survive); a chunk without survivors costs the 32 compares and one vote.
__device__ __forceinline__ void append_survivors(const float (&v)[32], const float* __restrict__ hcc, float thr, bool rok, uint32_t row,
                                                 uint32_t t0, int lane, TcPair* __restrict__ pairs, unsigned int* __restrict__ pair_count,
                                                 unsigned int pair_capacity) {
  uint32_t m = 0u;
#pragma unroll
  for (int j = 0; j < 32; ++j) m |= ((v[j] - hcc[j]) >= thr ? 1u : 0u) << j;
  if (!rok) m = 0u;
  const int cnt = __popc(m);
  if (!__any_sync(0xffffffffu, cnt != 0)) return;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  unsigned int base = 0u;
  if (lane == 31) base = atomicAdd(pair_count, (unsigned int)incl);
  base = __shfl_sync(0xffffffffu, base, 31) + (unsigned int)(incl - cnt);
  while (m) {
    const int j = __ffs(m) - 1;
    m &= m - 1u;
    if (base < pair_capacity) pairs[base] = TcPair{row, t0 + (uint32_t)j};
    ++base;
  }
}

__global__ void __launch_bounds__(kTcThreads, 1)
kc_screen_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_c,
                    const float* __restrict__ xx, const float* __restrict__ cc, const float* __restrict__ min_dist, int64_t n,
                    int d, int T, float alpha, int batch_min, TcPair* __restrict__ pairs, unsigned int* __restrict__ pair_count,
                    unsigned int pair_capacity, uint32_t t_base, KcCount cnt) {
  T = kc_effective_T(T, cnt);
  if (T <= 0) return;  // (device-side batch size: this chunk holds no centre; every thread of every CTA leaves)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* hcc = reinterpret_cast<float*>(smem + kTcStages * kTcStageBytes);  // |c_t|^2 / 2, +inf for t >= T
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes + 1024);
  uint64_t* full = bars;                       // [kTcStages]
  uint64_t* empty = bars + kTcStages;          // [kTcStages]
  uint64_t* tmem_full = bars + 2 * kTcStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* cc_max_slot = reinterpret_cast<float*>(tmem_base_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (n + kTcBlockM - 1) / kTcBlockM;
  const int nkb = (d + kTcBlockK - 1) / kTcBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = threadIdx.x; t < kTcBlockN; t += kTcThreads) hcc[t] = (t < T) ? 0.5f * __ldg(cc + t) : INFINITY;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 2) {
    float m = 0.0f;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, __ldg(cc + t));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if (lane == 0) *cc_max_slot = m;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const float cc_max = *cc_max_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], kTcStageBytes);
          unsigned char* a_dst = smem + stage * kTcStageBytes;
          tma_load_2d(a_dst, &map_x, kb * kTcBlockK, (int)(tile * kTcBlockM), &full[stage]);
          tma_load_2d(a_dst + kTcABytes, &map_c, kb * kTcBlockK, 0, &full[stage]);
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kTcBlockN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kTcStageBytes);
          const uint32_t b_addr = a_addr + kTcABytes;
#pragma unroll
          for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k) {
            umma_tf32(tmem_d, umma_smem_desc(a_addr + k * kTcUmmaK * 4), umma_smem_desc(b_addr + k * kTcUmmaK * 4), kTcIdesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int n_chunks = (T + 31) / 32;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      const int64_t row = tile * kTcBlockM + quarter * 32 + lane;
      const bool rok = row < n;
      const float xr = rok ? __ldg(xx + row) : 0.0f;
      const float mi = rok ? __ldg(min_dist + row) : 0.0f;
      const float m2 = __fmul_ru(mi, mi);
      const float s_i = alpha * (xr + cc_max);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kTcBlockN + ((uint32_t)(quarter * 32) << 16);
      float v[32];
      float vmax = -INFINITY;
      if (batch_min) {
        for (int c = 0; c < n_chunks; ++c) {
          tmem_ld32(taddr + c * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, v[j] - hcc[c * 32 + j]);
        }
      }
      // pair (i, t) survives iff val >= max(0.5 (|x|^2 - s - m^2), max_t val - s); without batch_min (every pair
      // that may beat m_i is wanted, not only the best of the batch) the second term is dropped
      const float thr = fmaxf(0.5f * ((xr - s_i) - m2), vmax - s_i);
      for (int c = 0; c < n_chunks; ++c) {
        tmem_ld32(taddr + c * 32, v);
        append_survivors(v, hcc + c * 32, thr, rok, (uint32_t)row, (uint32_t)(c * 32) + t_base, lane, pairs, pair_count, pair_capacity);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2-CTA variant (tcgen05 cta_group::2, one CTA pair = one TPC): the 1-CTA kernel above is bound by shared-memory
// bandwidth -- per 32-float k-block a CTA writes 48 KiB (TMA) and the tensor core reads 48 KiB back, 96 KiB per 524 MMA
// cycles against 128 B/clk -- and two thirds of that is the centre tile, identical for every CTA.  Here a pair of CTAs
// screens 256 feature rows against the 256 centres with ONE M = 256, N = 256 MMA per K = 8: each CTA stages its own 128
// feature rows (16 KiB) and only HALF of the centre tile (128 centres, 16 KiB); the tensor cores of both SMs read both
// halves.  Per SM and k-block: 32 KiB written + 32 KiB read, and half the L2 -> SM traffic for the centres.
//   rank 0 (leader)  warp 0 TMA producer (own A rows + centres 0..127), warp 1 lane 0 issues every MMA and commits to the
//                    stage-empty / accumulator-full barriers of BOTH CTAs (multicast), warps 2-5 epilogue of rows 0..127
//   rank 1           warp 0 TMA producer (own A rows + centres 128..255; its transaction bytes complete on the leader's
//                    full barrier), warp 1 only allocates / frees TMEM, warps 2-5 epilogue of rows 128..255 (arriving on
//                    the leader's accumulator-empty barrier through the cluster window)
// The accumulator of a CTA is its 128 rows x 256 centres, exactly as in the 1-CTA kernel, so the epilogue is unchanged.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTc2Stages = 6;
constexpr uint32_t kTc2HalfN = kTcBlockN / 2;
constexpr uint32_t kTc2BBytes = kTc2HalfN * kTcBlockK * 4;
constexpr uint32_t kTc2StageBytes = kTcABytes + kTc2BBytes;  // per CTA
constexpr uint32_t kTc2SmemBytes = kTc2Stages * kTc2StageBytes + 1024 /*hcc*/ + 256 /*barriers etc.*/ + 1024 /*alignment*/;
constexpr uint32_t kTc2Idesc =
    (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBlockN >> 3) << 17) | ((uint32_t)((2 * kTcBlockM) >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire (arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  long long t0 = 0;
  for (uint32_t polls = 0;; ++polls) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((polls & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 8000000000ll) asm volatile("trap;");
    }
  }
}
// TMA tile load of a CTA pair: the bytes land in THIS CTA's shared memory, the transaction completes on the barrier at
// `bar_cluster_addr` (the leader's full barrier, addressed through the cluster window).
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at the same shared-memory offset in both CTAs
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
kc_screen_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_c_half,
                     const float* __restrict__ xx, const float* __restrict__ cc, const float* __restrict__ min_dist, int64_t n,
                     int d, int T, float alpha, int batch_min, TcPair* __restrict__ pairs, unsigned int* __restrict__ pair_count,
                     unsigned int pair_capacity, uint32_t t_base, KcCount cnt) {
  T = kc_effective_T(T, cnt);
  if (T <= 0) return;  // (device-side batch size: this chunk holds no centre; every thread of every CTA leaves)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* hcc = reinterpret_cast<float*>(smem + kTc2Stages * kTc2StageBytes);  // |c_t|^2 / 2, +inf for t >= T
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTc2Stages * kTc2StageBytes + 1024);
  uint64_t* full = bars;                        // [kTc2Stages]  used in the leader only: bytes of both CTAs
  uint64_t* empty = bars + kTc2Stages;          // [kTc2Stages]  per CTA, arrived by the leader's multicast commit
  uint64_t* tmem_full = bars + 2 * kTc2Stages;  // [2]           per CTA, arrived by the leader's multicast commit
  uint64_t* tmem_empty = tmem_full + 2;         // [2]           used in the leader only: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* cc_max_slot = reinterpret_cast<float*>(tmem_base_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t n_pairs = (n + 2 * kTcBlockM - 1) / (2 * kTcBlockM);
  const int64_t first = blockIdx.x >> 1, step = gridDim.x >> 1;
  const int nkb = (d + kTcBlockK - 1) / kTcBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTc2Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = threadIdx.x; t < kTcBlockN; t += kTcThreads) hcc[t] = (t < T) ? 0.5f * __ldg(cc + t) : INFINITY;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp == 2) {
    float m = 0.0f;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, __ldg(cc + t));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if (lane == 0) *cc_max_slot = m;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const float cc_max = *cc_max_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tp = first; tp < n_pairs; tp += step) {
        const int row0 = (int)(tp * 2 * kTcBlockM + rank * kTcBlockM);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2u * kTc2StageBytes);
          const uint32_t bar = mapa_rank(smem_u32(&full[stage]), 0u);
          unsigned char* a_dst = smem + stage * kTc2StageBytes;
          tma_load_2d_2sm(a_dst, &map_x, kb * kTcBlockK, row0, bar);
          tma_load_2d_2sm(a_dst + kTcABytes, &map_c_half, kb * kTcBlockK, (int)(rank * kTc2HalfN), bar);
          if (++stage == kTc2Stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int64_t tp = first; tp < n_pairs; tp += step, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kTcBlockN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_cluster(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kTc2StageBytes);
          const uint32_t b_addr = a_addr + kTcABytes;
#pragma unroll
          for (int k = 0; k < kTcBlockK / kTcUmmaK; ++k) {
            umma_tf32_2sm(tmem_d, umma_smem_desc(a_addr + k * kTcUmmaK * 4), umma_smem_desc(b_addr + k * kTcUmmaK * 4), kTc2Idesc,
                          (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty[stage]);
          if (++stage == kTc2Stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tmem_full[acc]);
      }
    }
  } else {
    // ===== epilogue (warps 2..5 of both CTAs): TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int n_chunks = (T + 31) / 32;
    uint32_t it = 0;
    for (int64_t tp = first; tp < n_pairs; tp += step, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      const int64_t row = tp * 2 * kTcBlockM + rank * kTcBlockM + quarter * 32 + lane;
      const bool rok = row < n;
      const float xr = rok ? __ldg(xx + row) : 0.0f;
      const float mi = rok ? __ldg(min_dist + row) : 0.0f;
      const float m2 = __fmul_ru(mi, mi);
      const float s_i = alpha * (xr + cc_max);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kTcBlockN + ((uint32_t)(quarter * 32) << 16);
      float v[32];
      float vmax = -INFINITY;
      if (batch_min) {
        for (int c = 0; c < n_chunks; ++c) {
          tmem_ld32(taddr + c * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, v[j] - hcc[c * 32 + j]);
        }
      }
      const float thr = fmaxf(0.5f * ((xr - s_i) - m2), vmax - s_i);  // see kc_screen_tc_kernel
      for (int c = 0; c < n_chunks; ++c) {
        tmem_ld32(taddr + c * 32, v);
        append_survivors(v, hcc + c * 32, thr, rok, (uint32_t)row, (uint32_t)(c * 32) + t_base, lane, pairs, pair_count, pair_capacity);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty[acc]), 0u));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair still works on either CTA's memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Exact canonical distance of the surviving pairs; 32 pairs per warp (lane = pair for the fma chain, which the canonical
// order forbids splitting), rows moved through a shared-memory transpose so that global loads stay coalesced.
// A steady-state launch holds only ~50 k pairs = ~1 600 warp tasks on 148 SMs, so its duration IS the latency of one
// warp's instruction stream.  Round 1 moved 32 columns per step with scalar accesses (128 pointer shuffles + 64 LDG + 64
// STS + 64 LDS + 32 FMA): ~2.2 us per 32 columns, 0.14 ms per launch, 13 of the 30 ms of a 125k x 2048 shard's selection.
// (Letting every lane stream its own rows directly is no better: 32 distinct lines per load instruction, L1-throughput
// bound, also 0.15 ms -- measured in round 2.)  Here a step moves 64 columns with 128-bit accesses (a half-warp per row) and
// the row pointers sit in shared memory: 32 LDS.64 + 32 LDG.128 + 32 STS.128 + 32 LDS.128 + 64 FMA per 64 columns, ~4x fewer
// instructions per column; 17 KiB of tiles per warp keeps 12 warps resident per SM, enough for one task per warp.
// d % 4 == 0 and 16-byte aligned rows (guaranteed by kc_tc_applicable).
//   kStore = false: min_dist[row] = min(min_dist[row], dist)
//   kStore = true : out[(t0 + t) * ld_out + row] = dist        (candidate pairwise matrix of the replay)
constexpr int kRcWarps = 4;
constexpr int kRcTileK = 64, kRcStride = kRcTileK + 4;  // +4 floats: LDS.128 of 8 consecutive lanes hit 8 distinct bank groups
constexpr uint32_t kRcSmemBytes = kRcWarps * 2 * 32 * kRcStride * 4;
template <bool kStore>
__global__ void __launch_bounds__(kRcWarps * 32)
kc_recheck_kernel(const float* __restrict__ X, const float* __restrict__ xx, int d, const float* __restrict__ C,
                  const float* __restrict__ cc, const TcPair* __restrict__ pairs, const unsigned int* __restrict__ pair_count,
                  unsigned int pair_capacity, float* __restrict__ min_dist, int t0, int64_t ld_out) {
  extern __shared__ __align__(16) float rc_smem[];
  __shared__ const float* row_ptr[kRcWarps][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tx = rc_smem + (size_t)warp * (2 * 32 * kRcStride);
  float* tc = tx + 32 * kRcStride;
  unsigned int total = *pair_count;
  if (total > pair_capacity) total = pair_capacity;  // overflow: the FFMA fallback pass redoes everything
  const unsigned int n_groups = (total + 31) / 32;
  for (unsigned int g = blockIdx.x * kRcWarps + warp; g < n_groups; g += gridDim.x * kRcWarps) {
    const unsigned int p = g * 32 + lane;
    const bool ok = p < total;
    const TcPair pr = ok ? pairs[p] : TcPair{0u, 0u};
    __syncwarp();
    row_ptr[warp][lane] = X + (int64_t)pr.row * d;
    row_ptr[warp][32 + lane] = C + (int64_t)pr.t * d;
    __syncwarp();
    float acc = 0.0f;
    const int sub = lane >> 4, c4 = (lane & 15) * 4;  // a half-warp covers the 64 columns of one row
    for (int k0 = 0; k0 < d; k0 += kRcTileK) {
      const int k = k0 + c4;
      const bool kin = k < d;
      __syncwarp();  // the previous step's chain has finished reading the tiles
#pragma unroll
      for (int r0 = 0; r0 < 32; r0 += 16) {
        float4 vx[8], vc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + 2 * u + sub;
          vx[u] = kin ? __ldg(reinterpret_cast<const float4*>(row_ptr[warp][r] + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
          vc[u] = kin ? __ldg(reinterpret_cast<const float4*>(row_ptr[warp][32 + r] + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + 2 * u + sub;
          *reinterpret_cast<float4*>(tx + r * kRcStride + c4) = vx[u];
          *reinterpret_cast<float4*>(tc + r * kRcStride + c4) = vc[u];
        }
      }
      __syncwarp();
      const int n4 = ((d - k0) < kRcTileK ? (d - k0) : kRcTileK) >> 2;
      const float4* a4 = reinterpret_cast<const float4*>(tx + lane * kRcStride);
      const float4* b4 = reinterpret_cast<const float4*>(tc + lane * kRcStride);
      if (n4 == kRcTileK / 4) {
#pragma unroll 8
        for (int q = 0; q < kRcTileK / 4; ++q) {
          const float4 a = a4[q], b = b4[q];
          acc = __fmaf_rn(a.x, b.x, acc);
          acc = __fmaf_rn(a.y, b.y, acc);
          acc = __fmaf_rn(a.z, b.z, acc);
          acc = __fmaf_rn(a.w, b.w, acc);
        }
      } else {
        for (int q = 0; q < n4; ++q) {
          const float4 a = a4[q], b = b4[q];
          acc = __fmaf_rn(a.x, b.x, acc);
          acc = __fmaf_rn(a.y, b.y, acc);
          acc = __fmaf_rn(a.z, b.z, acc);
          acc = __fmaf_rn(a.w, b.w, acc);
        }
      }
    }
    if (ok) {
      const float dist = kc_dist(acc, __ldg(xx + pr.row), __ldg(cc + pr.t));
      if (kStore) {
        min_dist[(int64_t)(t0 + (int)pr.t) * ld_out + pr.row] = dist;
      } else if (dist < min_dist[pr.row]) {
        atomicMin(reinterpret_cast<unsigned int*>(min_dist + pr.row), __float_as_uint(dist));
      }
    }
  }
}

__global__ void kc_tc_reset_kernel(unsigned int* pair_count) { *pair_count = 0u; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* map, const float* base, int64_t rows, int d, int box_rows) {
  static PFN_encodeTiled encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MVAL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return MVAL_ERR_CUDA;
    }
    encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)d * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kTcBlockK, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld, d %d)", (int)r, (long long)rows, d);
    return MVAL_ERR_CUDA;
  }
  return MVAL_OK;
}

}  // namespace

// ffma fallback that only runs when the pair list overflowed (kcenter.cu)
int kc_update_batch_exact_if(const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T,
                             float* min_dist, const unsigned int* count, unsigned int capacity, cudaStream_t stream,
                             KcCount cnt = KcCount{nullptr, 0});

bool kc_tc_applicable(const float* X, int64_t n, int d, const float* C, int T) {
  return d % 4 == 0 && d >= 64 && n >= 128 && T >= 2 && T <= kTcBlockN && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(C) & 15) == 0;
}

static int tc_prepare(KcDeviceScratch** out, size_t want_pairs) {
  KcDeviceScratch* s = nullptr;
  if (int rc = kc_scratch(&s)) return rc;
  if (want_pairs < (1u << 20)) want_pairs = 1u << 20;
  if (want_pairs > (16u << 20)) want_pairs = 16u << 20;
  if (s->tc_pairs_capacity < want_pairs) {
    if (s->tc_pairs) cudaFree(s->tc_pairs);
    s->tc_pairs = nullptr;
    s->tc_pairs_capacity = 0;
    MVAL_CUDA(cudaMalloc(&s->tc_pairs, want_pairs * sizeof(TcPair)));
    s->tc_pairs_capacity = want_pairs;
  }
  if (s->tc_count == nullptr) MVAL_CUDA(cudaMalloc(&s->tc_count, sizeof(unsigned int)));
  MVAL_CUDA(cudaFuncSetAttribute(kc_screen_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  MVAL_CUDA(cudaFuncSetAttribute(kc_recheck_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRcSmemBytes));
  MVAL_CUDA(cudaFuncSetAttribute(kc_recheck_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRcSmemBytes));
  *out = s;
  return MVAL_OK;
}

// MVAL_TC_2CTA=0 keeps the 1-CTA kernel (A/B measurements, and the tests compare the two); read on every call.
static bool tc_use_pairs(int64_t n) {
  const char* e = getenv("MVAL_TC_2CTA");
  return !(e != nullptr && e[0] == '0') && n >= 4 * kTcBlockM;
}

// t_base is added to the centre index of every appended pair; reset = false appends to the list of the previous screen (several
// 256-centre chunks then share ONE recheck launch, whose duration is a latency, not a throughput)
static int tc_screen(KcDeviceScratch* s, const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T,
                     const float* min_dist, int batch_min, cudaStream_t stream, uint32_t t_base = 0, bool reset = true,
                     KcCount cnt = KcCount{nullptr, 0}) {
  CUtensorMap map_x, map_c;
  if (int rc = make_map(&map_x, X, n, d, kTcBlockM)) return rc;
  if (reset) {
    kc_tc_reset_kernel<<<1, 1, 0, stream>>>(s->tc_count);
    MVAL_LAUNCH_CHECK("kc_tc_reset");
  }
  const float alpha2 = ldexpf(1.0f, -8) + (float)d * ldexpf(1.0f, -20);
  if (tc_use_pairs(n)) {
    if (int rc = make_map(&map_c, C, T, d, (int)kTc2HalfN)) return rc;
    MVAL_CUDA(cudaFuncSetAttribute(kc_screen_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc2SmemBytes));
    const int64_t n_pairs = (n + 2 * kTcBlockM - 1) / (2 * kTcBlockM);
    // co-resident CTA pairs (a pair needs both SMs of a TPC): asked of the driver once; every cluster walks its own tile
    // pairs, so a smaller grid is merely slower, a larger one would leave a second wave of stragglers
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(2 * (num_sms() / 2)));
      cfg.blockDim = dim3(kTcThreads);
      cfg.dynamicSmemBytes = kTc2SmemBytes;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      int q = 0;
      if (cudaOccupancyMaxActiveClusters(&q, kc_screen_tc2_kernel, &cfg) != cudaSuccess || q <= 0) {
        (void)cudaGetLastError();
        q = num_sms() / 2;
      }
      max_clusters = q;
    }
    const int grid = 2 * (int)(n_pairs < max_clusters ? n_pairs : max_clusters);
    kc_screen_tc2_kernel<<<grid, kTcThreads, kTc2SmemBytes, stream>>>(map_x, map_c, xx, cc, min_dist, n, d, T, alpha2, batch_min,
                                                                     static_cast<TcPair*>(s->tc_pairs), s->tc_count,
                                                                     (unsigned int)s->tc_pairs_capacity, t_base, cnt);
    MVAL_LAUNCH_CHECK("kc_screen_tc2");
    return MVAL_OK;
  }
  if (int rc = make_map(&map_c, C, T, d, kTcBlockN)) return rc;
  const int64_t n_tiles = (n + kTcBlockM - 1) / kTcBlockM;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  const float alpha = ldexpf(1.0f, -8) + (float)d * ldexpf(1.0f, -20);
  kc_screen_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, stream>>>(map_x, map_c, xx, cc, min_dist, n, d, T, alpha, batch_min,
                                                                  static_cast<TcPair*>(s->tc_pairs), s->tc_count,
                                                                  (unsigned int)s->tc_pairs_capacity, t_base, cnt);
  MVAL_LAUNCH_CHECK("kc_screen_tc");
  return MVAL_OK;
}

__global__ void kc_fill_inf_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = INFINITY;
}

// Candidate pairwise matrix of the replay through the tensor-core screen: out_t[s * n + j] = dist(candidate j, centre s)
// wherever that distance may be below val[j] (the only entries the replay can ever act on), +inf elsewhere.
bool kc_pairwise_tc_applicable(const float* X, int n, int d) { return kc_tc_applicable(X, n, d, X, 2) && d >= 512; }

int kc_pairwise_tc(const float* X, const float* xx, const float* val, int n, int d, float* out_t, cudaStream_t stream) {
  KcDeviceScratch* s = nullptr;
  if (int rc = tc_prepare(&s, (size_t)n * kTcBlockN)) return rc;
  kc_fill_inf_kernel<<<64, 256, 0, stream>>>(out_t, (int64_t)n * n);
  MVAL_LAUNCH_CHECK("kc_fill_inf");
  for (int t0 = 0; t0 < n; t0 += kTcBlockN) {
    const int tn = (n - t0) < kTcBlockN ? (n - t0) : kTcBlockN;
    if (int rc = tc_screen(s, X, xx, n, d, X + (int64_t)t0 * d, xx + t0, tn, val, 0, stream, (uint32_t)t0, t0 == 0)) return rc;
  }
  // every chunk's survivors in one list (centre index absolute), one exact pass
  kc_recheck_kernel<true><<<num_sms() * 3, kRcWarps * 32, kRcSmemBytes, stream>>>(X, xx, d, X, xx, static_cast<const TcPair*>(s->tc_pairs),
                                                                                   s->tc_count, (unsigned int)s->tc_pairs_capacity,
                                                                                   out_t, 0, n);
  MVAL_LAUNCH_CHECK("kc_recheck_store");
  return MVAL_OK;
}

// T centres (any number) through the tensor-core screen in chunks of 256: every chunk appends its survivors (absolute centre
// index) to one list, then ONE recheck launch and one gated FFMA fallback per chunk (it only runs if the list overflowed).
int kc_update_batch_tc(const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T, float* min_dist,
                       cudaStream_t stream, const int32_t* t_dev = nullptr, int t_dev_off = 0) {
  KcDeviceScratch* s = nullptr;
  const int n_chunks = (T + kTcBlockN - 1) / kTcBlockN;
  if (int rc = tc_prepare(&s, (size_t)n * 4 * (size_t)n_chunks)) return rc;
  for (int t0 = 0; t0 < T; t0 += kTcBlockN) {
    const int tn = (T - t0) < kTcBlockN ? (T - t0) : kTcBlockN;
    if (int rc = tc_screen(s, X, xx, n, d, C + (int64_t)t0 * d, cc + t0, tn, min_dist, 1, stream, (uint32_t)t0, t0 == 0,
                           KcCount{t_dev, t_dev_off + t0}))
      return rc;
  }
  const unsigned int cap = (unsigned int)s->tc_pairs_capacity;
  kc_recheck_kernel<false><<<num_sms() * 3, kRcWarps * 32, kRcSmemBytes, stream>>>(X, xx, d, C, cc, static_cast<const TcPair*>(s->tc_pairs),
                                                                        s->tc_count, cap, min_dist, 0, 0);
  MVAL_LAUNCH_CHECK("kc_recheck");
  for (int t0 = 0; t0 < T; t0 += kTcBlockN) {
    const int tn = (T - t0) < kTcBlockN ? (T - t0) : kTcBlockN;
    if (int rc = kc_update_batch_exact_if(X, xx, n, d, C + (int64_t)t0 * d, cc + t0, tn, min_dist, s->tc_count, cap, stream,
                                          KcCount{t_dev, t_dev_off + t0}))
      return rc;
  }
  return MVAL_OK;
}

int kc_tc_last_stats(uint64_t* survivors, uint64_t* capacity, cudaStream_t stream) {
  KcDeviceScratch* s = nullptr;
  if (int rc = kc_scratch(&s)) return rc;
  *survivors = 0;
  *capacity = s->tc_pairs_capacity;
  if (s->tc_count == nullptr) return MVAL_OK;
  unsigned int c = 0;
  MVAL_CUDA(cudaMemcpyAsync(&c, s->tc_count, sizeof(c), cudaMemcpyDeviceToHost, stream));
  MVAL_CUDA(cudaStreamSynchronize(stream));
  *survivors = c;
  return MVAL_OK;
}

int kc_update_batch(const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T, float* min_dist,
                    int flags, cudaStream_t stream, const int32_t* t_dev) {
  if (n == 0 || T == 0) return MVAL_OK;
  const bool force_exact = (flags & kKcFlagForceExact) != 0;
  const bool force_tc = (flags & kKcFlagForceTc) != 0;
  // full 256-centre chunks share one tensor-core call (one recheck launch for all of them); a short tail chunk is judged on
  // its own: the tensor-core pass costs one sweep of the features whatever tn is, the FFMA pass ~ tn / 256 of 16 sweeps
  auto eligible = [&](const float* Cb, int tn) {
    return !force_exact && kc_tc_applicable(X, n, d, Cb, tn) && (force_tc || ((int64_t)tn * d >= 16 * 1024 && n >= 16384));
  };
  // (only for the picks of a greedy round, kKcFlagGroupChunks: there the running minima are settled and few pairs survive;
  // while minima are still falling -- the labeled fold -- a chunk's screen profits from the previous chunk's update)
  int t0 = 0;
  const int full = (flags & kKcFlagGroupChunks) ? (T / kTcBlockN) * kTcBlockN : 0;
  if (full > 0 && eligible(C, kTcBlockN)) {
    if (int rc = kc_update_batch_tc(X, xx, n, d, C, cc, full, min_dist, stream, t_dev, 0)) return rc;
    t0 = full;
  }
  for (; t0 < T; t0 += kTcBlockN) {
    const int tn = (T - t0) < kTcBlockN ? (T - t0) : kTcBlockN;
    const float* Cb = C + (int64_t)t0 * d;
    int rc;
    if (eligible(Cb, tn)) rc = kc_update_batch_tc(X, xx, n, d, Cb, cc + t0, tn, min_dist, stream, t_dev, t0);
    else rc = kc_update_batch_exact(X, xx, n, d, Cb, cc + t0, tn, min_dist, stream, KcCount{t_dev, t0});
    if (rc) return rc;
  }
  return MVAL_OK;
}

}  // namespace mval
