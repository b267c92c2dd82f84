// Shared declarations of the coreset k-center kernels (kcenter.cu: exact FFMA path, select / replay;
// kcenter_tc.cu: tcgen05 TF32 screening GEMM).
#pragma once
#include "common.cuh"

namespace mval {

constexpr int kKcMaxSlots = 1024;    // candidates the replay CTA can hold (one thread each)
constexpr int kKcGreedySlots = 256;  // candidate slots per round of the single-device loop
constexpr int kKcInitChunk = 256;    // labeled centres folded in per pass over the features

// canonical distance from the canonical dot product (oracle/coreset_oracle.c: dist_f32)
__device__ __forceinline__ float kc_dist(float dot, float xx, float cc) {
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), xx), cc);
  return __fadd_rn(__fsqrt_rn(fmaxf(d2, 0.0f)), 0.0f);
}

// Number of centres of a batched update when only the DEVICE knows it (the picks of a greedy round): the kernels are launched
// for the largest possible batch and clamp it themselves, T_effective = clamp(*ptr - off, 0, T).  ptr == nullptr: T as given.
struct KcCount {
  const int32_t* ptr;
  int off;
};
__device__ __forceinline__ int kc_effective_T(int T, KcCount c) {
  if (c.ptr == nullptr) return T;
  const int r = *c.ptr - c.off;
  return r < T ? (r < 0 ? 0 : r) : T;
}

struct KcPartial {
  float val;
  int64_t idx;
};

struct KcSelectState {
  uint32_t hist1[4096];
  uint32_t hist2[1024];
  uint32_t hist3[1024];
  int32_t b1, b2;
  uint32_t k1, k2;
  uint32_t kappa, all, n_cand, first_eq;
};

struct KcDeviceScratch {
  KcSelectState* sel = nullptr;
  uint32_t* cand_idx = nullptr;
  KcPartial* partials = nullptr;
  unsigned int* counter = nullptr;
  int32_t* host_i32 = nullptr;  // pinned
  // tensor-core screening path (kcenter_tc.cu)
  void* tc_pairs = nullptr;
  size_t tc_pairs_capacity = 0;
  unsigned int* tc_count = nullptr;
};

int kc_scratch(KcDeviceScratch** out);

__host__ __device__ inline size_t kc_align256(size_t x) { return (x + 255) & ~size_t(255); }

// Candidate record block of one shard, K slots, feature dimension d (all-gathered between ranks once per round):
//   [0, 16)                 header {int32 count; float tau; 8 bytes padding}
//   [16, 16 + 4K)           float   val[K]    running minimum of the candidate (-1 = empty slot)
//   [.., + 4K)              float   xx[K]     squared norm of the candidate row
//   [.., + 8K)              int64   gidx[K]   global row index (INT64_MAX = empty slot)
//   [16 + 16K, + 4 K d)     float   rows[K][d]
struct KcRecordHead {
  int32_t count;
  float tau;
  int32_t pad[2];
};
struct KcRecordView {
  KcRecordHead* head;
  float* val;
  float* xx;
  int64_t* gidx;
  float* rows;
};
__host__ __device__ inline size_t kc_records_bytes(int K, int d) { return 16 + 16 * (size_t)K + 4 * (size_t)K * d; }
__host__ __device__ inline KcRecordView kc_record_view(char* base, int K, int d) {
  KcRecordView v;
  v.head = reinterpret_cast<KcRecordHead*>(base);
  v.val = reinterpret_cast<float*>(base + 16);
  v.xx = reinterpret_cast<float*>(base + 16 + 4 * (size_t)K);
  v.gidx = reinterpret_cast<int64_t*>(base + 16 + 8 * (size_t)K);
  v.rows = reinterpret_cast<float*>(base + 16 + 16 * (size_t)K);
  (void)d;
  return v;
}

// kcenter.cu
int kc_norms(const float* X, int64_t n, int d, float* out, cudaStream_t stream);
int kc_update_batch_exact(const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T,
                          float* min_dist, cudaStream_t stream, KcCount cnt = KcCount{nullptr, 0});
int kc_select(const float* X, const float* xx, const float* m, int64_t n, int d, int64_t index_offset, int K, void* records,
              cudaStream_t stream);
size_t kc_resolve_workspace_bytes(int n_blocks, int K, int d);
// state (device int32 [4], optional): [0] picks made so far, [1] picks of this round (out), [2] budget.  With a state the
// round's limit is budget - done, the picks go to selected_out[done ..], the counters are advanced on the device and the
// call does NOT synchronise (n_picks_host is not written).
int kc_resolve(const void* records, int n_blocks, int K, int d, int max_picks, void* workspace, float* centres,
               float* centre_norms, int64_t* selected_out, int32_t* n_picks_host, cudaStream_t stream, int32_t* state = nullptr);

// Dispatcher: exact FFMA pass, or (large aligned d, enough rows and centres) tensor-core screening + exact recheck.
// flags bit 0: force the exact FFMA pass; bit 1: force the tensor-core path when it is applicable at all.
constexpr int kKcFlagForceExact = 1;
constexpr int kKcFlagForceTc = 2;
constexpr int kKcFlagGroupChunks = 4;  // the centres are the picks of one greedy round: screen all chunks, recheck once
int kc_update_batch(const float* X, const float* xx, int64_t n, int d, const float* C, const float* cc, int T, float* min_dist,
                    int flags, cudaStream_t stream, const int32_t* t_dev = nullptr);

// candidate pairwise matrix through the tensor-core screen (kcenter_tc.cu)
bool kc_pairwise_tc_applicable(const float* X, int n, int d);
int kc_pairwise_tc(const float* X, const float* xx, const float* val, int n, int d, float* out_t, cudaStream_t stream);

// survivors of the last tensor-core screen on this device (synchronises `stream`) and the capacity of the pair list
int kc_tc_last_stats(uint64_t* survivors, uint64_t* capacity, cudaStream_t stream);

}  // namespace mval
