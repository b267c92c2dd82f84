// Per-frame bookkeeping kernels around the scoring pass (reference strategy.py:952-975, 1134-1145;
// utils/evaluation.py:198-208; utils/coreset.py:35-47).  All are O(n J) element-wise work on results that already live
// on the device; they exist so that a pool's scores never have to visit the host before the selection is made.
#include "common.cuh"

namespace mval {

// utils/coreset.py:41-46: feature row = (pose^T[0:3, :] - pose^T[0:3, root]).flatten() = x_0..x_{J-1}, y.., z.. relative to
// the root joint.  The reference builds it from sal_dict["pred_3d_keypoints"], i.e. from the float32 roundings of the
// triangulated joints (strategy.py:1046), subtracts in float64 and CoreSet's device copy rounds to float32 once more.
template <typename T>
__global__ void __launch_bounds__(256)
pose_features_kernel(const T* __restrict__ xyz, int64_t n, int J, int root, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t d = 3ll * J;
  if (i >= n * d) return;
  const int64_t f = i / d;
  const int r = (int)(i - f * d);
  const int c = r / J, j = r - c * J;
  const float a = (float)xyz[(f * J + j) * 3 + c], b = (float)xyz[(f * J + root) * 3 + c];
  out[i] = (float)((double)a - (double)b);
}

// utils/evaluation.py:198-208 compute_mkpe([pred], [gt], [valid]) for every frame (strategy.py:1134-1145), float32 like
// the reference's tensors: per joint sqrt(sum_c where(valid, (pred - gt)^2, 0)) / valid (0 / 0 = NaN for an invalid
// joint, exactly as the reference), then the mean over joints (accumulated in float64, rounded once).
// pred float32 [n][J][3]; gt float32 [n][gt_rows][J] (rows 0..2 = x, y, z; the dataset stores 4 rows); valid float32 [n][J].
__global__ void __launch_bounds__(128)
mkpe_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ valid, int64_t n, int J,
            int gt_rows, float* __restrict__ out) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const float* p = pred + f * J * 3;
  const float* g = gt + f * gt_rows * J;
  const float* v = valid + f * J;
  double acc = 0.0;
  for (int j = 0; j < J; ++j) {
    float s = 0.f;
    if (v[j] != 0.f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = __fsub_rn(p[j * 3 + c], g[c * J + j]);
        s = __fadd_rn(s, __fmul_rn(d, d));
      }
    }
    acc += (double)__fdiv_rn(sqrtf(s), v[j]);
  }
  out[f] = (float)(acc / (double)J);
}

// strategy.py:957-975: candidates for pseudo-labelling are the frames that were not picked by the AL step, are not
// already pseudo-labelled (both folded into `excluded` by the caller), have a non-NaN sal_metric and
// inlier_count > threshold; they are visited in ascending sal_metric order, ties in pool order (Python's sorted is
// stable).  Written as a descending ranking key for mval_topk_desc: NaN drops a frame, -metric + 0 orders the rest.
__global__ void __launch_bounds__(256)
sal_key_kernel(const float* __restrict__ sal_metric, const float* __restrict__ inlier_count, const uint8_t* __restrict__ excluded,
               int64_t n, float inlier_threshold, double* __restrict__ key) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float m = sal_metric[i];
  const bool keep = !(m != m) && !(excluded != nullptr && excluded[i] != 0) && inlier_count[i] > inlier_threshold;
  key[i] = keep ? (-(double)m + 0.0) : __longlong_as_double(0x7ff8000000000000ll);
}

// strategy.py:981-989: cluster_id = self.kmeans.predict([kp])[0] with kp the root-relative pose of the frame (float64
// differences of the float32 predictions, x_0..x_{J-1}, y.., z..), one sklearn call per candidate in the reference.
// sklearn's predict is argmin_c (|c|^2 - 2 x.c) in float64, first centre on ties (lloyd_iter_chunked_dense with
// update_centers=False).  One thread per frame; the dot product is a sequential float64 fma chain, so it can differ
// from sklearn's GEMM by rounding only: the margin to the runner-up is returned as well and the host re-asks sklearn
// for the frames whose margin is within rounding distance.
__global__ void __launch_bounds__(128)
kmeans_assign_kernel(const float* __restrict__ pred, int64_t n, int J, int root, const double* __restrict__ centres,
                     const double* __restrict__ centre_sq, int k, int32_t* __restrict__ label, double* __restrict__ margin) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const float* p = pred + f * J * 3;
  const double rx = (double)p[root * 3], ry = (double)p[root * 3 + 1], rz = (double)p[root * 3 + 2];
  double best = INFINITY, second = INFINITY;
  int arg = 0;
  for (int c = 0; c < k; ++c) {
    const double* cc = centres + (int64_t)c * 3 * J;
    double dot = 0.0;
    for (int j = 0; j < J; ++j) dot = fma((double)p[j * 3] - rx, __ldg(cc + j), dot);
    for (int j = 0; j < J; ++j) dot = fma((double)p[j * 3 + 1] - ry, __ldg(cc + J + j), dot);
    for (int j = 0; j < J; ++j) dot = fma((double)p[j * 3 + 2] - rz, __ldg(cc + 2 * J + j), dot);
    const double s = fma(-2.0, dot, __ldg(centre_sq + c));
    if (s < best) {
      second = best;
      best = s;
      arg = c;
    } else if (s < second) {
      second = s;
    }
  }
  label[f] = arg;
  if (margin) margin[f] = second - best;  // +inf with a single centre, NaN when the scores are not comparable
}

__global__ void __launch_bounds__(128) centre_sq_kernel(const double* __restrict__ centres, int k, int d, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  double s = 0.0;
  for (int i = 0; i < d; ++i) s = fma(centres[(int64_t)c * d + i], centres[(int64_t)c * d + i], s);
  out[c] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// Frame aggregation of the per-map scores with the REFERENCE'S OWN arithmetic (strategy.py:1151-1158, 1188-1193, 1210-1215),
// one thread per frame, so that the HP / MPE / BSB frame scores -- and with them the ranking -- never visit the host:
//   values of a frame = its per-map scores in view-major order over the VALID joints (m = V * #valid);
//   HP   (Python floats, .item()):  AVG = builtin sum() / len in double -- since Python 3.12 sum() is Neumaier-compensated
//        (Python/bltinmodule.c), before that a plain left-to-right sum: `compensated` says which;  STD = np.std of a float64
//        array;
//   MPE / BSB (np.float32 scalars): AVG = builtin sum() = one float32 add after the other, / len in float32;  STD = np.std of
//        a float32 array (NumPy >= 2 promotion).
// np.std = sqrt(pairwise_sum((x - pairwise_sum(x) / m)^2) / m) in the array's precision, with NumPy's pairwise summation:
// fewer than 8 values: left to right from 0; up to 128: eight interleaved accumulators r[k] += a[i + k], combined as
// ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), the tail added left to right; above 128: split at n / 2 rounded down to
// a multiple of 8 (numpy/core/src/umath/loops_utils.h).  Every operation is an explicitly rounded add / mul / div / sqrt:
// no fused multiply-add may sneak into (x - mean)^2 + acc.  Checked bit for bit against the host restatement
// (strategy._aggregate_map_scores), which is itself pinned to the reference's golden values.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAggMaxJoints = 128;

template <typename T> struct Rn;
template <> struct Rn<double> {
  __device__ static __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  __device__ static __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  __device__ static __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  __device__ static __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  __device__ static __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};
template <> struct Rn<float> {
  __device__ static __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  __device__ static __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  __device__ static __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  __device__ static __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  __device__ static __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};

// the i-th value of a frame (view-major over the valid joints), optionally as its squared deviation from `mean`
template <typename T>
struct FrameValues {
  const float* maps;        // the frame's [V][J] scores
  const int16_t* jidx;      // indices of its valid joints
  int J, c;                 // joints per view, valid joints
  bool squared;
  T mean;
  __device__ __forceinline__ T at(int i) const {
    const int v = i / c;
    const T x = (T)maps[v * J + jidx[i - v * c]];
    if (!squared) return x;
    const T d = Rn<T>::sub(x, mean);
    return Rn<T>::mul(d, d);
  }
};

template <typename T>
__device__ T np_pairwise_sum(const FrameValues<T>& a, int lo, int n) {
  if (n < 8) {
    T res = (T)0;
    for (int i = 0; i < n; ++i) res = Rn<T>::add(res, a.at(lo + i));
    return res;
  }
  if (n <= 128) {
    T r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = a.at(lo + k);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = Rn<T>::add(r[k], a.at(lo + i + k));
    }
    T res = Rn<T>::add(Rn<T>::add(Rn<T>::add(r[0], r[1]), Rn<T>::add(r[2], r[3])),
                       Rn<T>::add(Rn<T>::add(r[4], r[5]), Rn<T>::add(r[6], r[7])));
    for (; i < n; ++i) res = Rn<T>::add(res, a.at(lo + i));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return Rn<T>::add(np_pairwise_sum(a, lo, n2), np_pairwise_sum(a, lo + n2, n - n2));
}

template <typename T>
__device__ T np_std(FrameValues<T> a, int m) {
  a.squared = false;
  const T mean = Rn<T>::div(np_pairwise_sum(a, 0, m), (T)m);
  a.squared = true;
  a.mean = mean;
  return Rn<T>::sqrt(Rn<T>::div(np_pairwise_sum(a, 0, m), (T)m));
}

__global__ void __launch_bounds__(128)
frame_aggregate_kernel(const float* __restrict__ per_map, const uint8_t* __restrict__ valid, int64_t n_frames, int V, int J,
                       int kind_hp, int config_std, int compensated, double* __restrict__ out) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  int16_t jidx[kAggMaxJoints];
  int c = 0;
  for (int j = 0; j < J; ++j)
    if (valid == nullptr || valid[f * J + j] != 0) jidx[c++] = (int16_t)j;
  const int m = V * c;
  const float* maps = per_map + f * (int64_t)V * J;
  double res;
  if (m == 0) {
    res = __longlong_as_double(0x7ff8000000000000ll);  // the reference divides by len([]) / takes np.std([]) here
  } else if (config_std) {
    if (kind_hp) {
      FrameValues<double> a{maps, jidx, J, c, false, 0.0};
      res = np_std<double>(a, m);
    } else {
      FrameValues<float> a{maps, jidx, J, c, false, 0.0f};
      res = (double)np_std<float>(a, m);
    }
  } else if (kind_hp) {
    // builtin sum() over Python floats, then / len
    FrameValues<double> a{maps, jidx, J, c, false, 0.0};
    double s = 0.0, comp = 0.0;
    for (int i = 0; i < m; ++i) {
      const double x = a.at(i);
      const double nxt = __dadd_rn(s, x);
      if (compensated) comp = __dadd_rn(comp, fabs(s) >= fabs(x) ? __dadd_rn(__dsub_rn(s, nxt), x) : __dadd_rn(__dsub_rn(x, nxt), s));
      s = nxt;
    }
    if (compensated && comp != 0.0 && isfinite(comp)) s = __dadd_rn(s, comp);
    res = __ddiv_rn(s, (double)m);
  } else {
    // builtin sum() over np.float32 scalars: one float32 add after the other (0 + x0 first), then / len in float32
    FrameValues<float> a{maps, jidx, J, c, false, 0.0f};
    float s = 0.0f;
    for (int i = 0; i < m; ++i) s = __fadd_rn(s, a.at(i));
    res = (double)__fdiv_rn(s, (float)m);
  }
  out[f] = res;
}

}  // namespace mval

extern "C" int mval_aggregate_map_scores(const float* per_map, const uint8_t* valid, int64_t n_frames, int V, int J, int kind,
                                         int config_std, int compensated_sum, double* out, void* stream) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && V > 0 && J > 0, "mval_aggregate_map_scores: bad shape");
  MVAL_REQUIRE(kind >= MVAL_MAP_SCORE_HP && kind <= MVAL_MAP_SCORE_BSB, "mval_aggregate_map_scores: kind must be MVAL_MAP_SCORE_HP / MPE / BSB");
  if (J > kAggMaxJoints) {
    set_error("mval_aggregate_map_scores: J=%d exceeds %d", J, kAggMaxJoints);
    return MVAL_ERR_UNSUPPORTED;
  }
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(per_map && out, "mval_aggregate_map_scores: null pointer");
  frame_aggregate_kernel<<<(unsigned)((n_frames + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      per_map, valid, n_frames, V, J, kind == MVAL_MAP_SCORE_HP ? 1 : 0, config_std ? 1 : 0, compensated_sum ? 1 : 0, out);
  MVAL_LAUNCH_CHECK("frame_aggregate");
  return MVAL_OK;
}

extern "C" int mval_kmeans_assign(const float* pred, int64_t n_frames, int J, int root, const double* centres, int k,
                                  int32_t* out_label, double* out_margin, void* stream_) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && J > 0 && root >= 0 && root < J && k > 0, "mval_kmeans_assign: bad shape, root joint or k");
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(pred && centres && out_label, "mval_kmeans_assign: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  double* sq = nullptr;
  MVAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&sq), sizeof(double) * k, stream));
  centre_sq_kernel<<<(unsigned)((k + 127) / 128), 128, 0, stream>>>(centres, k, 3 * J, sq);
  kmeans_assign_kernel<<<(unsigned)((n_frames + 127) / 128), 128, 0, stream>>>(pred, n_frames, J, root, centres, sq, k, out_label,
                                                                              out_margin);
  count_launch(2);
  int rc = MVAL_OK;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) rc = cuda_fail(e, "launch kmeans_assign");
  e = cudaFreeAsync(sq, stream);
  if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  return rc;
}

extern "C" int mval_pose_features(const void* xyz, int xyz_is_double, int64_t n_frames, int J, int root, float* out_features,
                                  void* stream) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && J > 0 && root >= 0 && root < J, "mval_pose_features: bad shape or root joint");
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(xyz && out_features, "mval_pose_features: null pointer");
  const int64_t total = n_frames * 3 * J;
  const int64_t blocks = (total + 255) / 256;
  MVAL_REQUIRE(blocks <= 0x7fffffffLL, "mval_pose_features: too many frames for one launch; chunk the pool");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (xyz_is_double)
    pose_features_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const double*>(xyz), n_frames, J, root, out_features);
  else
    pose_features_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(xyz), n_frames, J, root, out_features);
  MVAL_LAUNCH_CHECK("pose_features");
  return MVAL_OK;
}

extern "C" int mval_mkpe(const float* pred, const float* gt, const float* valid, int64_t n_frames, int J, int gt_rows,
                         float* out_mkpe, void* stream) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && J > 0 && gt_rows >= 3, "mval_mkpe: bad shape (gt needs at least the x, y, z rows)");
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(pred && gt && valid && out_mkpe, "mval_mkpe: null pointer");
  mkpe_kernel<<<(unsigned)((n_frames + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(pred, gt, valid, n_frames, J,
                                                                                                gt_rows, out_mkpe);
  MVAL_LAUNCH_CHECK("mkpe");
  return MVAL_OK;
}

extern "C" int mval_sal_rank(const float* sal_metric, const float* inlier_count, const uint8_t* excluded, int64_t n,
                             float inlier_threshold, int32_t k, int64_t* out_idx, int32_t* out_count, void* stream_) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n >= 0 && k >= 0, "mval_sal_rank: bad sizes");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0 || k == 0) return mval_topk_desc(nullptr, 0, 0, k, out_idx, nullptr, out_count, stream_);
  MVAL_REQUIRE(sal_metric && inlier_count && out_idx, "mval_sal_rank: null pointer");
  double* key = nullptr;
  MVAL_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&key), sizeof(double) * n, stream));
  sal_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sal_metric, inlier_count, excluded, n, inlier_threshold, key);
  count_launch();
  int rc = MVAL_OK;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) rc = cuda_fail(e, "launch sal_key");
  if (rc == MVAL_OK) rc = mval_topk_desc(key, n, 0, k, out_idx, nullptr, out_count, stream_);
  e = cudaFreeAsync(key, stream);
  if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  return rc;
}
