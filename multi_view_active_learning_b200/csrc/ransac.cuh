// Device-side building blocks of the RANSAC / DLT triangulation shared by the stand-alone kernels
// (triangulate.cu) and the fused persistent pool kernel (fused.cu).  See triangulate.cu for the algorithm notes.
#pragma once
#include <math.h>

#include "common.cuh"

namespace mval {

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


// RANSAC vote of one (frame, joint) by one warp (utils/triangulation.py:284-300): lane = view pair, `pair_at(pi, a, b)`
// yields the pi-th pair to visit.  P: V*12 doubles, px/py: V doubles (any address space, read by every lane).
// Returns (to all lanes) the inlier bit mask of the first pair with the strictly largest inlier set.
template <typename PairFn>
__device__ __forceinline__ uint32_t ransac_vote_warp(const double* __restrict__ P, const double* __restrict__ px,
                                                     const double* __restrict__ py, int V, int n_pairs, double eps,
                                                     PairFn pair_at, int lane) {
  const double thr = 2.0 * eps;
  int best_cnt = -1, best_pi = 0x7fffffff;
  uint32_t best_mask = 0u;
  for (int base = 0; base < n_pairs; base += kWarp) {
    const int pi = base + lane;
    const bool active = pi < n_pairs;
    int a = 0, b = 1;
    if (active) pair_at(pi, a, b);
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
  return __shfl_sync(kFull, best_mask, win & 31);
}

// Final solve of one (frame, joint) by one thread (utils/triangulation.py:306-316): DLT over the inlier views in
// ascending order and the mean of 0.5*||kp - proj|| over exactly those views.  `xy_at(v, x, y)` yields the key-point
// of view v as doubles.
template <typename XyFn>
__device__ __forceinline__ void ransac_final_thread(const double* __restrict__ P, XyFn xy_at, uint32_t mask, int V,
                                                    double& X, double& Y, double& Z, double& reproj_mean) {
  double m[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) m[i][j] = 0.0;
  for (int v = 0; v < V; ++v) {
    if (mask >> v & 1u) {
      double Pv[12], x, y;
#pragma unroll
      for (int i = 0; i < 12; ++i) Pv[i] = P[v * 12 + i];
      xy_at(v, x, y);
      accumulate_view(m, Pv, x, y);
    }
  }
  smallest_eigvec_dehom<false>(m, X, Y, Z);
  double sum = 0.0;
  for (int v = 0; v < V; ++v) {
    if (mask >> v & 1u) {
      const double* Pv = P + v * 12;
      double x, y;
      xy_at(v, x, y);
      const double h0 = fma(Pv[0], X, fma(Pv[1], Y, fma(Pv[2], Z, Pv[3])));
      const double h1 = fma(Pv[4], X, fma(Pv[5], Y, fma(Pv[6], Z, Pv[7])));
      double h2 = fma(Pv[8], X, fma(Pv[9], Y, fma(Pv[10], Z, Pv[11])));
      h2 = (h2 == 0.0) ? 1.0 : h2;
      const double dx = x - h0 / h2;
      const double dy = y - h1 / h2;
      sum = fma(0.5, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), sum);
    }
  }
  reproj_mean = sum / (double)__popc(mask);
}

// Lexicographic pair table (itertools.combinations order): entry idx = a*(2V-a-1)/2 + (b-a-1).
__device__ __forceinline__ void build_pair_table(uint8_t* sPair, int V, int tid, int nthreads) {
  for (int a = tid; a < V - 1; a += nthreads) {
    int idx = a * (2 * V - a - 1) / 2;
    for (int b = a + 1; b < V; ++b, ++idx) {
      sPair[2 * idx] = (uint8_t)a;
      sPair[2 * idx + 1] = (uint8_t)b;
    }
  }
}

// Counter-based partial Fisher-Yates (same arithmetic as oracle/triangulation_oracle.py:pair_subset_indices):
// after the call perm[0..n_iters) are the indices of the pairs to visit.  Called by a whole warp.
__device__ __forceinline__ void draw_pair_subset(uint16_t* perm, int n_all, int n_iters, uint64_t seed,
                                                 int64_t global_frame, int joint, int lane) {
  for (int i = lane; i < n_all; i += kWarp) perm[i] = (uint16_t)i;
  __syncwarp();
  if (lane == 0) {
    uint64_t state = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)global_frame * 64ull + (uint64_t)(joint + 1));
    for (int i = 0; i < n_iters; ++i) {
      const uint64_t z = splitmix64(state);
      const int r = i + (int)(((z >> 32) * (uint64_t)(n_all - i)) >> 32);
      const uint16_t tmp = perm[i];
      perm[i] = perm[r];
      perm[r] = tmp;
    }
  }
  __syncwarp();
}

}  // namespace mval
