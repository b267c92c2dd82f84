// Direct reprojection-error minimisation (reference utils/triangulation.py:319-336, direct_optimization=True):
//
//   res = scipy.optimize.least_squares(residuals, x0, loss="huber", method="trf")
//
// with residuals r_v(x) = 0.5 * ||kp_v - proj_v(x)|| over the RANSAC inlier views and x0 the DLT solution on those
// views.  scipy minimises F(x) = 0.5 * sum_v rho(r_v^2), rho(z) = z for z <= 1 and 2 sqrt(z) - 1 above (f_scale = 1),
// i.e. F = sum_v phi(r_v) with phi(r) = r^2 / 2 inside |r| <= 1 and r - 1/2 outside; it stops at ftol = xtol = gtol =
// 1e-8, a few 1e-3 mm short of the minimiser on the synthetic rigs.  Here every (frame, valid joint) is one thread that
// runs a damped Newton iteration on the same F to convergence of the step (|dx| <= 1e-12 (1 + |x|)), so the result is
// the minimiser itself; tests bound the distance to the reference's output by the 1e-2 mm / 1e-3 relative contract.
//
// With e_v = proj_v(x) - kp_v (2-vector), a_v = J_v^T e_v, J_v the 2 x 3 Jacobian of the projection:
//   |e_v| <= 2 :  F += |e_v|^2 / 8,      g += a_v / 4,          H += J_v^T J_v / 4
//   |e_v| >  2 :  F += |e_v| / 2 - 1/2,  g += a_v / (2 |e_v|),  H += (J_v^T J_v - a_v a_v^T / |e_v|^2) / (2 |e_v|)
// (Gauss-Newton in the projection, exact in the Huber kink; H is positive semi-definite.)  A step solves
// (H + lambda diag H) dx = -g; it is accepted when F does not increase (lambda /= 10), otherwise lambda *= 10.
#include "common.cuh"

namespace mval {

constexpr int kRefineMaxIters = 100;

struct HuberModel {
  double F, g[3], H[6];  // H: xx xy xz yy yz zz
};

template <typename XyFn>
__device__ __forceinline__ void huber_model(const double* __restrict__ P, XyFn xy_at, uint32_t mask, int V, double X, double Y,
                                            double Z, HuberModel& m) {
  m.F = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) m.g[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) m.H[i] = 0.0;
  for (int v = 0; v < V; ++v) {
    if (!(mask >> v & 1u)) continue;
    const double* Pv = P + v * 12;
    double kx, ky;
    xy_at(v, kx, ky);
    const double h0 = fma(Pv[0], X, fma(Pv[1], Y, fma(Pv[2], Z, Pv[3])));
    const double h1 = fma(Pv[4], X, fma(Pv[5], Y, fma(Pv[6], Z, Pv[7])));
    double h2 = fma(Pv[8], X, fma(Pv[9], Y, fma(Pv[10], Z, Pv[11])));
    h2 = (h2 == 0.0) ? 1.0 : h2;
    const double iw = 1.0 / h2;
    const double pu = h0 * iw, pv = h1 * iw;
    const double ex = pu - kx, ey = pv - ky;
    double ju[3], jv[3], a[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ju[c] = (Pv[c] - pu * Pv[8 + c]) * iw;
      jv[c] = (Pv[4 + c] - pv * Pv[8 + c]) * iw;
      a[c] = fma(ju[c], ex, jv[c] * ey);
    }
    const double e2 = fma(ex, ex, ey * ey);
    if (e2 <= 4.0) {  // r = |e| / 2 <= 1: quadratic zone
      m.F = fma(0.125, e2, m.F);
#pragma unroll
      for (int c = 0; c < 3; ++c) m.g[c] = fma(0.25, a[c], m.g[c]);
      m.H[0] += 0.25 * fma(ju[0], ju[0], jv[0] * jv[0]);
      m.H[1] += 0.25 * fma(ju[0], ju[1], jv[0] * jv[1]);
      m.H[2] += 0.25 * fma(ju[0], ju[2], jv[0] * jv[2]);
      m.H[3] += 0.25 * fma(ju[1], ju[1], jv[1] * jv[1]);
      m.H[4] += 0.25 * fma(ju[1], ju[2], jv[1] * jv[2]);
      m.H[5] += 0.25 * fma(ju[2], ju[2], jv[2] * jv[2]);
    } else {  // linear zone
      const double e = sqrt(e2);
      const double s = 0.5 / e, t = 1.0 / e2;
      m.F += 0.5 * e - 0.5;
#pragma unroll
      for (int c = 0; c < 3; ++c) m.g[c] = fma(s, a[c], m.g[c]);
      m.H[0] += s * (fma(ju[0], ju[0], jv[0] * jv[0]) - t * a[0] * a[0]);
      m.H[1] += s * (fma(ju[0], ju[1], jv[0] * jv[1]) - t * a[0] * a[1]);
      m.H[2] += s * (fma(ju[0], ju[2], jv[0] * jv[2]) - t * a[0] * a[2]);
      m.H[3] += s * (fma(ju[1], ju[1], jv[1] * jv[1]) - t * a[1] * a[1]);
      m.H[4] += s * (fma(ju[1], ju[2], jv[1] * jv[2]) - t * a[1] * a[2]);
      m.H[5] += s * (fma(ju[2], ju[2], jv[2] * jv[2]) - t * a[2] * a[2]);
    }
  }
}

// Solves the symmetric 3 x 3 system A d = -g by Cramer's rule; false when A is (numerically) singular.
__device__ __forceinline__ bool solve_sym3(const double (&A)[6], const double (&g)[3], double (&d)[3]) {
  const double c00 = A[3] * A[5] - A[4] * A[4];
  const double c01 = A[2] * A[4] - A[1] * A[5];
  const double c02 = A[1] * A[4] - A[2] * A[3];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  if (!(fabs(det) > 0.0) || !isfinite(det)) return false;
  const double c11 = A[0] * A[5] - A[2] * A[2];
  const double c12 = A[1] * A[2] - A[0] * A[4];
  const double c22 = A[0] * A[3] - A[1] * A[1];
  const double id = -1.0 / det;
  d[0] = id * (c00 * g[0] + c01 * g[1] + c02 * g[2]);
  d[1] = id * (c01 * g[0] + c11 * g[1] + c12 * g[2]);
  d[2] = id * (c02 * g[0] + c12 * g[1] + c22 * g[2]);
  return true;
}

template <typename PT>
__global__ void __launch_bounds__(128)
refine_huber_kernel(const PT* __restrict__ xy, const double* __restrict__ proj, const uint8_t* __restrict__ valid,
                    const uint32_t* __restrict__ masks, int64_t n_tasks, int V, int J, double* __restrict__ xyz,
                    double* __restrict__ reproj, int32_t* __restrict__ iters_out) {
  const int64_t task = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= n_tasks) return;
  if (valid != nullptr && valid[task] == 0) return;  // invalid joints stay (0, 0, 0) / NaN as triangulation left them
  const int64_t frame = task / J;
  const int joint = (int)(task % J);
  const uint32_t mask = masks[task];
  const double* __restrict__ P = proj + frame * V * 12;
  const PT* __restrict__ q = xy + (frame * V * J + joint) * 2;
  auto xy_at = [&](int v, double& x, double& y) {
    x = (double)q[(int64_t)v * J * 2];
    y = (double)q[(int64_t)v * J * 2 + 1];
  };
  double x[3] = {xyz[3 * task], xyz[3 * task + 1], xyz[3 * task + 2]};
  HuberModel cur;
  huber_model(P, xy_at, mask, V, x[0], x[1], x[2], cur);
  double lam = 1e-4;
  int it = 0;
  for (; it < kRefineMaxIters; ++it) {
    double A[6] = {cur.H[0] * (1.0 + lam), cur.H[1], cur.H[2], cur.H[3] * (1.0 + lam), cur.H[4], cur.H[5] * (1.0 + lam)};
    double d[3];
    if (!solve_sym3(A, cur.g, d)) {
      lam *= 10.0;
      if (lam > 1e12) break;
      continue;
    }
    const double xt[3] = {x[0] + d[0], x[1] + d[1], x[2] + d[2]};
    HuberModel trial;
    huber_model(P, xy_at, mask, V, xt[0], xt[1], xt[2], trial);
    if (trial.F <= cur.F) {
      const double step2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      const double scale = 1.0 + sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      x[0] = xt[0]; x[1] = xt[1]; x[2] = xt[2];
      cur = trial;
      lam = fmax(lam * 0.1, 1e-15);
      if (step2 <= 1e-24 * scale * scale) { ++it; break; }
    } else {
      lam *= 10.0;
      if (lam > 1e12) break;
    }
  }
  xyz[3 * task] = x[0];
  xyz[3 * task + 1] = x[1];
  xyz[3 * task + 2] = x[2];
  if (iters_out) iters_out[task] = it;
  if (reproj) {  // :332-336: mean of 0.5 * ||kp - proj|| over the inlier views at the refined point
    double sum = 0.0;
    for (int v = 0; v < V; ++v) {
      if (mask >> v & 1u) {
        const double* Pv = P + v * 12;
        double kx, ky;
        xy_at(v, kx, ky);
        const double h0 = fma(Pv[0], x[0], fma(Pv[1], x[1], fma(Pv[2], x[2], Pv[3])));
        const double h1 = fma(Pv[4], x[0], fma(Pv[5], x[1], fma(Pv[6], x[2], Pv[7])));
        double h2 = fma(Pv[8], x[0], fma(Pv[9], x[1], fma(Pv[10], x[2], Pv[11])));
        h2 = (h2 == 0.0) ? 1.0 : h2;
        const double dx = kx - h0 / h2, dy = ky - h1 / h2;
        sum = fma(0.5, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), sum);
      }
    }
    reproj[task] = sum / (double)__popc(mask);
  }
}

// triangulate.cu
int launch_frame_reduce(const double* reproj, const int32_t* inliers, const uint8_t* valid, int64_t n_frames, int J,
                        double* out_metric, int32_t* out_inlier_count, cudaStream_t stream);

}  // namespace mval

extern "C" int mval_refine_huber(const void* xy, int xy_is_float, const double* proj, const uint8_t* valid,
                                 const uint32_t* inlier_mask, const int32_t* inliers, int64_t n_frames, int V, int J,
                                 double* xyz, double* reproj, double* out_metric, int32_t* out_inlier_count,
                                 int32_t* out_iters, void* stream_) {
  using namespace mval;
  if (int rc = require_device()) return rc;
  MVAL_REQUIRE(n_frames >= 0 && J > 0, "mval_refine_huber: bad shape");
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(V >= 2, "mval_refine_huber: need at least 2 views");
  if (V > MVAL_MAX_VIEWS) {
    set_error("mval_refine_huber: V=%d exceeds MVAL_MAX_VIEWS=%d", V, MVAL_MAX_VIEWS);
    return MVAL_ERR_UNSUPPORTED;
  }
  if (n_frames == 0) return MVAL_OK;
  MVAL_REQUIRE(xy && proj && inlier_mask && xyz && reproj, "mval_refine_huber: null pointer");
  MVAL_REQUIRE((out_metric == nullptr) == (out_inlier_count == nullptr) && (out_metric == nullptr || inliers != nullptr),
               "mval_refine_huber: out_metric, out_inlier_count and inliers go together");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t n_tasks = n_frames * J;
  const int64_t blocks = (n_tasks + 127) / 128;
  MVAL_REQUIRE(blocks <= 0x7fffffffLL, "mval_refine_huber: too many (frame, joint) tasks for one launch; chunk the pool");
  if (xy_is_float)
    refine_huber_kernel<float><<<(unsigned)blocks, 128, 0, stream>>>(static_cast<const float*>(xy), proj, valid, inlier_mask,
                                                                     n_tasks, V, J, xyz, reproj, out_iters);
  else
    refine_huber_kernel<int32_t><<<(unsigned)blocks, 128, 0, stream>>>(static_cast<const int32_t*>(xy), proj, valid,
                                                                       inlier_mask, n_tasks, V, J, xyz, reproj, out_iters);
  MVAL_LAUNCH_CHECK("refine_huber");
  if (out_metric != nullptr) return launch_frame_reduce(reproj, inliers, valid, n_frames, J, out_metric, out_inlier_count, stream);
  return MVAL_OK;
}
