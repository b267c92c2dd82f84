/*
 * CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT) for the float32 coreset k-center greedy selection.
 *
 * Restates reference utils/coreset.py:49-95 (update_distances / select_batch) with the distance of
 * sklearn.metrics.pairwise_distances(..., "euclidean") -- sqrt(max(|x|^2 - 2 x.c + |c|^2, 0)) -- evaluated in
 * float32 in the CANONICAL ORDER the CUDA kernels (csrc/kcenter.cu) share:
 *
 *     dot(x, c)  = fmaf(x[d-1], c[d-1], ... fmaf(x[1], c[1], fmaf(x[0], c[0], +0.0f)) ...)   one accumulator, k ascending
 *     |x|^2      = dot(x, x)
 *     d2(x, c)   = ((-2 * dot(x, c)) + |x|^2) + |c|^2        x = pool row, c = centre (not symmetric in rounding)
 *     dist(x, c) = sqrtf(fmaxf(d2, 0)) + 0.0f
 *
 * Tie-break of np.argmax (coreset.py:90): lowest index among equal maxima.  The straightforward loop below is the
 * definition; the GPU's batched / tensor-core-screened rounds must reproduce its output bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/__init__.py: build_c_oracle()).  fmaf() is the
 * correctly rounded C99 fused multiply-add (hardware FMA or glibc's exact software path -- same bits).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline float dot_f32(const float* x, const float* c, int d) {
  float acc = 0.0f;
  for (int k = 0; k < d; ++k) acc = fmaf(x[k], c[k], acc);
  return acc;
}

static inline float dist_f32(float dot, float xx, float cc) {
  const float d2 = ((-2.0f * dot) + xx) + cc;
  return sqrtf(fmaxf(d2, 0.0f)) + 0.0f;
}

void kc_norms(const float* X, int64_t n, int d, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) out[i] = dot_f32(X + i * d, X + i * d, d);
}

/* min_d[i] = min(min_d[i], dist(X[i], c)) -- coreset.py:64-69 for one centre */
void kc_update(const float* X, const float* xx, int64_t n, int d, const float* c, float cc, float* min_d) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float dist = dist_f32(dot_f32(X + i * d, c, d), xx[i], cc);
    if (dist < min_d[i]) min_d[i] = dist;
  }
}

/* dist_out[i] = dist(X[i], c) */
void kc_dist(const float* X, const float* xx, int64_t n, int d, const float* c, float cc, float* dist_out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) dist_out[i] = dist_f32(dot_f32(X + i * d, c, d), xx[i], cc);
}

static int64_t argmax_first(const float* v, int64_t n) {
  int64_t best = 0;
  for (int64_t i = 1; i < n; ++i)
    if (v[i] > v[best]) best = i;
  return best;
}

/* coreset.py:71-95.  Rows [n_unlabeled, n) are the labeled centres.  min_d [n] out, selected [budget] out.
 * Returns 0, or -1 when there is no labeled centre (np.argmax(None) in the reference). */
int kc_greedy(const float* X, int64_t n, int64_t n_unlabeled, int d, int budget, float* min_d, int64_t* selected) {
  if (n_unlabeled >= n) return -1;
  float* xx = (float*)malloc(sizeof(float) * (size_t)n);
  kc_norms(X, n, d, xx);
  for (int64_t i = 0; i < n; ++i) min_d[i] = INFINITY;
  for (int64_t c = n_unlabeled; c < n; ++c) kc_update(X, xx, n, d, X + c * d, xx[c], min_d);
  for (int t = 0; t < budget; ++t) {
    const int64_t ind = argmax_first(min_d, n);
    selected[t] = ind;
    kc_update(X, xx, n, d, X + ind * d, xx[ind], min_d);
  }
  free(xx);
  return 0;
}
