// Multi-view RANSAC + DLT triangulation and reprojection uncertainty (sm_100a, float64 throughout).
//
// What the reference does per (frame, valid joint) in Python/NumPy (utils/triangulation.py:260-316), here for a
// whole batch of frames in three launches:
//
//   ransac_vote_kernel   one WARP per (frame, joint); lane = view pair (<= 32 pairs per round, 64 pairs = 2
//                        rounds).  Each lane builds the 4x4 DLT system of its pair in registers (:356-361), solves
//                        it (smallest eigenvector of A^T A by cyclic Jacobi, see below), projects the candidate
//                        into all V views (:371-384, 459-484) and votes; the first pair with the strictly largest
//                        inlier set wins (:293-300) via REDUX max / min.  Output: the inlier bit mask.
//   ransac_final_kernel  one THREAD per (frame, joint): accumulates A^T A over the inlier views in ascending
//                        view order (:306-311), solves, and averages the reprojection error over exactly those
//                        views (:312-316).
//   frame_reduce_kernel  one thread per frame: metric = mean over valid joints (:226), inlier_count = min (:231).
//
// Solver.  The reference takes vh[3] of LAPACK's SVD of A (:363-364).  We take the eigenvector of the smallest
// eigenvalue of the 4x4 matrix A^T A with cyclic Jacobi rotations in float64: unconditionally convergent, no
// data-dependent control flow beyond a warp-uniform early exit, 26 doubles of state per lane.  The rotation
// angle only steers convergence, so tan(theta) comes from the approximate rsqrt/rcp units; cos(theta) fixes the
// orthogonality of the transform and is refined to full precision by two Newton steps.  Measured against
// LAPACK on the synthetic rigs: <= 2e-10 relative on X, <= 2e-10 px on reprojection errors.
#include <math.h>

#include "common.cuh"

namespace mval {

constexpr int kVoteThreads = 256;
constexpr int kVoteWarps = kVoteThreads / kWarp;
constexpr int kMaxSweeps = 10;

__device__ __forceinline__ double rsqrt_approx(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp_approx(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
// 1/sqrt(x) for x in [1, 2.x]: hardware seed (~2^-22) + two Newton steps -> ~1 ulp.
__device__ __forceinline__ double rsqrt_refined(double x) {
  double y = rsqrt_approx(x);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double r = fma(-(x * y), y, 1.0);
    y = fma(0.5 * y, r, y);
  }
  return y;
}

// Symmetric 4x4 kept as its upper triangle m[i][j], i <= j (lower entries are never touched); all indices are
// compile-time constants after unrolling, so m and v live in registers.
#define MS(i, j) m[((i) < (j)) ? (i) : (j)][((i) < (j)) ? (j) : (i)]

template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(double (&m)[4][4], double (&v)[4][4]) {
  const double apq = m[P][Q], app = m[P][P], aqq = m[Q][Q];
  const double d = aqq - app;
  const double rad = fma(d, d, 4.0 * apq * apq);
  const double den = fabs(d) + rad * rsqrt_approx(rad);
  double t = (apq + apq) * rcp_approx(den);
  t = (d < 0.0) ? -t : t;
  t = (fabs(t) <= 2.0) ? t : 0.0;  // NaN/inf from a zero or overflowing radicand -> skip this rotation
  const double c = rsqrt_refined(fma(t, t, 1.0));
  const double s = t * c;
  const double cc = c * c, ss = s * s, cs2 = 2.0 * c * s;
  m[P][P] = fma(cc, app, fma(-cs2, apq, ss * aqq));
  m[Q][Q] = fma(ss, app, fma(cs2, apq, cc * aqq));
  m[P][Q] = fma(cc - ss, apq, 0.5 * cs2 * (app - aqq));
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (r != P && r != Q) {
      const double arp = MS(r, P), arq = MS(r, Q);
      MS(r, P) = fma(c, arp, -s * arq);
      MS(r, Q) = fma(s, arp, c * arq);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double vp = v[r][P], vq = v[r][Q];
    v[r][P] = fma(c, vp, -s * vq);
    v[r][Q] = fma(s, vp, c * vq);
  }
}

// Eigenvector of the smallest eigenvalue of the symmetric PSD matrix m (upper triangle), de-homogenised like
// utils/triangulation.py:387-399 (a 4th component of exactly 0 is replaced by 1).  kWarpUniform: all 32 lanes
// call this together and leave the sweep loop together.
template <bool kWarpUniform>
__device__ __forceinline__ void smallest_eigvec_dehom(double (&m)[4][4], double& X, double& Y, double& Z) {
  double v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
    const double off = m[0][1] * m[0][1] + m[0][2] * m[0][2] + m[0][3] * m[0][3] + m[1][2] * m[1][2] +
                       m[1][3] * m[1][3] + m[2][3] * m[2][3];
    const double dg = m[0][0] * m[0][0] + m[1][1] * m[1][1] + m[2][2] * m[2][2] + m[3][3] * m[3][3];
    const bool done = !(off > 1e-34 * dg);
    if (kWarpUniform ? __all_sync(kFull, done) : done) break;
    jacobi_rotate<0, 1>(m, v);
    jacobi_rotate<0, 2>(m, v);
    jacobi_rotate<0, 3>(m, v);
    jacobi_rotate<1, 2>(m, v);
    jacobi_rotate<1, 3>(m, v);
    jacobi_rotate<2, 3>(m, v);
  }
  double best = m[0][0];
  double e0 = v[0][0], e1 = v[1][0], e2 = v[2][0], e3 = v[3][0];
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    if (m[k][k] < best) {
      best = m[k][k];
      e0 = v[0][k]; e1 = v[1][k]; e2 = v[2][k]; e3 = v[3][k];
    }
  }
  const double w = (e3 == 0.0) ? 1.0 : e3;
  X = e0 / w;
  Y = e1 / w;
  Z = e2 / w;
}

// Adds the two DLT rows of one view to the upper triangle of A^T A.  The rows are formed with separately
// rounded multiply and subtract, exactly as NumPy evaluates  u * P[2, :] - P[0, :]  (:358-361).
__device__ __forceinline__ void accumulate_view(double (&m)[4][4], const double* __restrict__ P, double x, double y) {
  double ru[4], rv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ru[c] = __dsub_rn(__dmul_rn(x, P[8 + c]), P[c]);
    rv[c] = __dsub_rn(__dmul_rn(y, P[8 + c]), P[4 + c]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) m[i][j] = fma(ru[i], ru[j], fma(rv[i], rv[j], m[i][j]));
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t& state) {
  state += 0x9E3779B97F4A7C15ull;
  uint64_t z = state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename PT>
__global__ void __launch_bounds__(kVoteThreads)
ransac_vote_kernel(const PT* __restrict__ xy, const double* __restrict__ proj, const uint8_t* __restrict__ valid,
                   int64_t n_tasks, int V, int J, int n_iters, double eps, uint64_t seed, int64_t frame_offset,
                   const uint8_t* __restrict__ pairs_explicit, uint32_t* __restrict__ out_mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_all = V * (V - 1) / 2;
  const bool subset = n_all > n_iters;
  const int n_pairs = subset ? n_iters : n_all;
  // carve: P per warp | x,y per warp | lexicographic pair table | permutation scratch per warp
  double* sP = reinterpret_cast<double*>(smem_raw);                      // [kVoteWarps][V*12]
  double* sXY = sP + kVoteWarps * V * 12;                                // [kVoteWarps][2*V]
  uint8_t* sPair = reinterpret_cast<uint8_t*>(sXY + kVoteWarps * 2 * V);  // [n_all][2]
  uint16_t* sPerm = reinterpret_cast<uint16_t*>(sPair + ((2 * n_all + 15) & ~15));  // [kVoteWarps][n_all]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int a = threadIdx.x; a < V - 1; a += blockDim.x) {
    int idx = a * (2 * V - a - 1) / 2;
    for (int b = a + 1; b < V; ++b, ++idx) {
      sPair[2 * idx] = (uint8_t)a;
      sPair[2 * idx + 1] = (uint8_t)b;
    }
  }
  __syncthreads();

  const int64_t task = (int64_t)blockIdx.x * kVoteWarps + warp;
  if (task >= n_tasks) return;
  const int64_t frame = task / J;
  const int joint = (int)(task % J);
  if (valid != nullptr && valid[task] == 0) {
    if (lane == 0) out_mask[task] = 0u;
    return;
  }
  double* P = sP + warp * V * 12;
  double* px = sXY + warp * 2 * V;
  double* py = px + V;
  const double* gP = proj + frame * V * 12;
  for (int i = lane; i < V * 12; i += kWarp) P[i] = gP[i];
  for (int v = lane; v < V; v += kWarp) {
    const PT* q = xy + ((frame * V + v) * J + joint) * 2;
    px[v] = (double)q[0];
    py[v] = (double)q[1];
  }
  uint16_t* perm = sPerm + warp * n_all;
  if (subset && pairs_explicit == nullptr) {
    // counter-based partial Fisher-Yates, same arithmetic as oracle/triangulation_oracle.py:pair_subset_indices
    for (int i = lane; i < n_all; i += kWarp) perm[i] = (uint16_t)i;
    __syncwarp();
    if (lane == 0) {
      uint64_t state = seed + 0x9E3779B97F4A7C15ull * (uint64_t)((frame_offset + frame) * 64 + joint + 1);
      for (int i = 0; i < n_iters; ++i) {
        const uint64_t z = splitmix64(state);
        const int r = i + (int)(((z >> 32) * (uint64_t)(n_all - i)) >> 32);
        const uint16_t tmp = perm[i];
        perm[i] = perm[r];
        perm[r] = tmp;
      }
    }
  }
  __syncwarp();

  const double thr = 2.0 * eps;
  int best_cnt = -1, best_pi = 0x7fffffff;
  uint32_t best_mask = 0u;
  for (int base = 0; base < n_pairs; base += kWarp) {
    const int pi = base + lane;
    const bool active = pi < n_pairs;
    int a = 0, b = 1;
    if (active) {
      if (!subset) {
        a = sPair[2 * pi];
        b = sPair[2 * pi + 1];
      } else if (pairs_explicit != nullptr) {
        const uint8_t* e = pairs_explicit + ((int64_t)task * n_iters + pi) * 2;
        a = e[0];
        b = e[1];
      } else {
        const int li = perm[pi];
        a = sPair[2 * li];
        b = sPair[2 * li + 1];
      }
    }
    double m[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) m[i][j] = 0.0;
    if (active) {
      accumulate_view(m, P + a * 12, px[a], py[a]);
      accumulate_view(m, P + b * 12, px[b], py[b]);
    }
    double X, Y, Z;
    smallest_eigvec_dehom<true>(m, X, Y, Z);
    uint32_t mask = (1u << a) | (1u << b);
    for (int v = 0; v < V; ++v) {
      const double* Pv = P + v * 12;
      const double h0 = fma(Pv[0], X, fma(Pv[1], Y, fma(Pv[2], Z, Pv[3])));
      const double h1 = fma(Pv[4], X, fma(Pv[5], Y, fma(Pv[6], Z, Pv[7])));
      double h2 = fma(Pv[8], X, fma(Pv[9], Y, fma(Pv[10], Z, Pv[11])));
      h2 = (h2 == 0.0) ? 1.0 : h2;
      // 0.5*sqrt((x-h0/h2)^2 + (y-h1/h2)^2) < eps  <=>  (x*h2-h0)^2 + (y*h2-h1)^2 < (2*eps*h2)^2
      const double dx = fma(px[v], h2, -h0), dy = fma(py[v], h2, -h1), r = thr * h2;
      if (fma(dx, dx, dy * dy) < r * r) mask |= 1u << v;
    }
    const int cnt = __popc(mask);
    if (active && cnt > best_cnt) {  // strict: the earlier pair of this lane is kept on ties
      best_cnt = cnt;
      best_pi = pi;
      best_mask = mask;
    }
  }
  const int top = __reduce_max_sync(kFull, best_cnt);
  const int win = __reduce_min_sync(kFull, best_cnt == top ? best_pi : 0x7fffffff);
  const uint32_t mask = __shfl_sync(kFull, best_mask, win & 31);
  if (lane == 0) out_mask[task] = mask;
}

template <typename PT>
__global__ void __launch_bounds__(128)
ransac_final_kernel(const PT* __restrict__ xy, const double* __restrict__ proj, const uint8_t* __restrict__ valid,
                    const uint32_t* __restrict__ masks, int64_t n_tasks, int V, int J, double* __restrict__ out_xyz,
                    double* __restrict__ out_reproj, int32_t* __restrict__ out_inliers) {
  const int64_t task = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= n_tasks) return;
  const int64_t frame = task / J;
  const int joint = (int)(task % J);
  if (valid != nullptr && valid[task] == 0) {
    out_xyz[3 * task] = 0.0;
    out_xyz[3 * task + 1] = 0.0;
    out_xyz[3 * task + 2] = 0.0;
    if (out_reproj) out_reproj[task] = __longlong_as_double(0x7ff8000000000000ll);
    if (out_inliers) out_inliers[task] = 0;
    return;
  }
  const uint32_t mask = masks[task];
  const double* __restrict__ gP = proj + frame * V * 12;
  const PT* __restrict__ q = xy + (frame * V * J + joint) * 2;
  double m[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) m[i][j] = 0.0;
  for (int v = 0; v < V; ++v) {
    if (mask >> v & 1u) {
      double Pv[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) Pv[i] = __ldg(gP + v * 12 + i);
      accumulate_view(m, Pv, (double)q[(int64_t)v * J * 2], (double)q[(int64_t)v * J * 2 + 1]);
    }
  }
  double X, Y, Z;
  smallest_eigvec_dehom<false>(m, X, Y, Z);
  double sum = 0.0;
  for (int v = 0; v < V; ++v) {
    if (mask >> v & 1u) {
      const double* Pv = gP + v * 12;
      const double h0 = fma(__ldg(Pv + 0), X, fma(__ldg(Pv + 1), Y, fma(__ldg(Pv + 2), Z, __ldg(Pv + 3))));
      const double h1 = fma(__ldg(Pv + 4), X, fma(__ldg(Pv + 5), Y, fma(__ldg(Pv + 6), Z, __ldg(Pv + 7))));
      double h2 = fma(__ldg(Pv + 8), X, fma(__ldg(Pv + 9), Y, fma(__ldg(Pv + 10), Z, __ldg(Pv + 11))));
      h2 = (h2 == 0.0) ? 1.0 : h2;
      const double dx = (double)q[(int64_t)v * J * 2] - h0 / h2;
      const double dy = (double)q[(int64_t)v * J * 2 + 1] - h1 / h2;
      sum += 0.5 * sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    }
  }
  const int n = __popc(mask);
  out_xyz[3 * task] = X;
  out_xyz[3 * task + 1] = Y;
  out_xyz[3 * task + 2] = Z;
  if (out_reproj) out_reproj[task] = sum / (double)n;
  if (out_inliers) out_inliers[task] = n;
}

// metric = mean over valid joints of the per-joint score; inlier_count = min over valid joints.
// Recomputes nothing: reads the per-joint outputs (or, when the caller did not ask for them, the scratch copies).
__global__ void __launch_bounds__(128)
frame_reduce_kernel(const double* __restrict__ reproj, const int32_t* __restrict__ inliers,
                    const uint8_t* __restrict__ valid, int64_t n_frames, int J, double* __restrict__ out_metric,
                    int32_t* __restrict__ out_inlier_count) {
  const int64_t frame = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (frame >= n_frames) return;
  double sum = 0.0;
  int cnt = 0, mn = 0x7fffffff;
  for (int j = 0; j < J; ++j) {
    if (valid == nullptr || valid[frame * J + j]) {
      sum += reproj[frame * J + j];
      mn = min(mn, inliers[frame * J + j]);
      ++cnt;
    }
  }
  out_metric[frame] = cnt ? sum / (double)cnt : __longlong_as_double(0x7ff8000000000000ll);
  out_inlier_count[frame] = cnt ? mn : 0;
}

static size_t vote_smem_bytes(int V) {
  const int n_all = V * (V - 1) / 2;
  return sizeof(double) * kVoteWarps * V * 14 + ((2 * n_all + 15) & ~15) + sizeof(uint16_t) * kVoteWarps * n_all;
}

template <typename PT>
static int launch_ransac(const PT* xy, const double* proj, const uint8_t* valid, int64_t n_frames, int V, int J,
                         const mval_ransac_params& prm, double* out_xyz, double* out_reproj, int32_t* out_inliers,
                         uint32_t* mask, double* out_metric, int32_t* out_inlier_count, cudaStream_t stream) {
  const int64_t n_tasks = n_frames * J;
  const int64_t vote_blocks = (n_tasks + kVoteWarps - 1) / kVoteWarps;
  if (vote_blocks > 0x7fffffffLL) {
    set_error("mval_triangulate_ransac: too many (frame, joint) tasks for one launch; chunk the pool");
    return MVAL_ERR_UNSUPPORTED;
  }
  const size_t smem = vote_smem_bytes(V);
  ransac_vote_kernel<PT><<<(unsigned)vote_blocks, kVoteThreads, smem, stream>>>(
      xy, proj, valid, n_tasks, V, J, prm.n_iters, prm.epsilon, prm.pair_seed, prm.frame_offset, prm.pairs, mask);
  MVAL_LAUNCH_CHECK("ransac_vote");
  ransac_final_kernel<PT><<<(unsigned)((n_tasks + 127) / 128), 128, 0, stream>>>(xy, proj, valid, mask, n_tasks, V, J,
                                                                                out_xyz, out_reproj, out_inliers);
  MVAL_LAUNCH_CHECK("ransac_final");
  if (out_metric != nullptr) {
    frame_reduce_kernel<<<(unsigned)((n_frames + 127) / 128), 128, 0, stream>>>(out_reproj, out_inliers, valid, n_frames,
                                                                              J, out_metric, out_inlier_count);
    MVAL_LAUNCH_CHECK("frame_reduce");
  }
  return MVAL_OK;
}

int triangulate_ransac(const void* xy, int xy_is_float, const double* proj, const uint8_t* valid, int64_t n_frames,
                       int V, int J, const mval_ransac_params* params, double* out_xyz, double* out_reproj,
                       int32_t* out_inliers, uint32_t* out_mask, double* out_metric, int32_t* out_inlier_count,
                       cudaStream_t stream) {
  MVAL_REQUIRE(params != nullptr, "mval_triangulate_ransac: params is null");
  MVAL_REQUIRE(n_frames >= 0 && J > 0, "mval_triangulate_ransac: bad shape");
  // utils/triangulation.py:268  assert len(points) >= 2
  MVAL_REQUIRE(V >= 2, "mval_triangulate_ransac: need at least 2 views (reference asserts len(points) >= 2)");
  if (V > MVAL_MAX_VIEWS) {
    set_error("mval_triangulate_ransac: V=%d exceeds MVAL_MAX_VIEWS=%d", V, MVAL_MAX_VIEWS);
    return MVAL_ERR_UNSUPPORTED;
  }
  MVAL_REQUIRE(params->n_iters >= 1, "mval_triangulate_ransac: n_iters must be >= 1");
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(xy && proj && out_xyz, "mval_triangulate_ransac: null pointer");
  MVAL_REQUIRE((out_metric == nullptr) == (out_inlier_count == nullptr),
               "mval_triangulate_ransac: out_metric and out_inlier_count go together");
  const int64_t n_tasks = n_frames * J;
  // scratch for outputs the caller did not ask for but later stages need
  void* scratch = nullptr;
  size_t need = 0;
  const size_t off_mask = need;   if (!out_mask) need += (sizeof(uint32_t) * n_tasks + 15) & ~size_t(15);
  const size_t off_reproj = need; if (!out_reproj && out_metric) need += (sizeof(double) * n_tasks + 15) & ~size_t(15);
  const size_t off_inl = need;    if (!out_inliers && out_metric) need += (sizeof(int32_t) * n_tasks + 15) & ~size_t(15);
  if (need) {
    MVAL_CUDA(cudaMallocAsync(&scratch, need, stream));
    char* base = static_cast<char*>(scratch);
    if (!out_mask) out_mask = reinterpret_cast<uint32_t*>(base + off_mask);
    if (!out_reproj && out_metric) out_reproj = reinterpret_cast<double*>(base + off_reproj);
    if (!out_inliers && out_metric) out_inliers = reinterpret_cast<int32_t*>(base + off_inl);
  }
  int rc;
  if (xy_is_float)
    rc = launch_ransac<float>(static_cast<const float*>(xy), proj, valid, n_frames, V, J, *params, out_xyz, out_reproj,
                              out_inliers, out_mask, out_metric, out_inlier_count, stream);
  else
    rc = launch_ransac<int32_t>(static_cast<const int32_t*>(xy), proj, valid, n_frames, V, J, *params, out_xyz,
                                out_reproj, out_inliers, out_mask, out_metric, out_inlier_count, stream);
  if (scratch) {
    cudaError_t e = cudaFreeAsync(scratch, stream);
    if (rc == MVAL_OK && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  }
  return rc;
}

}  // namespace mval

extern "C" int mval_triangulate_ransac(const void* xy, int xy_is_float, const double* proj, const uint8_t* valid,
                                       int64_t n_frames, int V, int J, const mval_ransac_params* params,
                                       double* out_xyz, double* out_reproj, int32_t* out_inliers, uint32_t* out_mask,
                                       double* out_metric, int32_t* out_inlier_count, void* stream) {
  if (int rc = mval::require_device()) return rc;
  return mval::triangulate_ransac(xy, xy_is_float, proj, valid, n_frames, V, J, params, out_xyz, out_reproj,
                                  out_inliers, out_mask, out_metric, out_inlier_count,
                                  static_cast<cudaStream_t>(stream));
}
