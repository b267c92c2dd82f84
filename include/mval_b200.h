/*
 * mval_b200.h -- C ABI of the B200-native (sm_100a) active-learning scoring-and-selection hot path.
 *
 * The upstream reference (facebookresearch/multi_view_active_learning) is pure Python and has no FFI
 * layer; the functions below are what a ctypes binding of its hot path binds (INTEGRATION.md shows the
 * stub).  Every entry point cites the reference interface it replaces as  file:line  relative to the
 * reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - "device" pointers are caller-owned CUDA device memory on the current device, "host" pointers are
 *     caller-owned host memory (pinned if the copy is to overlap); nothing is retained after return
 *     except by the explicit *_create / *_destroy handle pairs;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all device entry points
 *     are asynchronous with respect to the host and ordered on `stream`;
 *   - return value: 0 = MVAL_OK, negative = error (see enum); mval_last_error() returns a thread-local
 *     human readable message for the last failing call.  No exceptions cross the ABI;
 *   - there is NO CPU fallback: every compute entry point requires a CUDA device.
 *
 * Layouts (row-major, innermost last)
 *   heatmaps  float32 [n_frames][V][J][H][W]          (reference strategy.py:1035  heatmaps.view(B,-1,kp,w,h))
 *   proj      float64 [n_frames][V][3][4]             (reference dataset/dataset.py:195, strategy.py:1030)
 *   valid     uint8   [n_frames][J]  (0/1; NULL = all valid)   (strategy.py:1031 joint_valid)
 *   xy        int32 / float32 [n_frames][V][J][2]     (x, y) in image pixels = heat-map pixel * stride
 */
#ifndef MVAL_B200_H_
#define MVAL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVAL_ABI_VERSION 5
#define MVAL_MAX_VIEWS 32
#define MVAL_MAX_SEGMENTS 64

enum mval_status {
  MVAL_OK = 0,
  MVAL_ERR_INVALID_ARGUMENT = -1, /* reference raises AssertionError / TypeError (utils/triangulation.py:267-268) */
  MVAL_ERR_UNSUPPORTED = -2,      /* shape outside what the kernels cover (e.g. V > 32)                         */
  MVAL_ERR_CUDA = -3,             /* a CUDA runtime call or kernel launch failed                                */
  MVAL_ERR_NO_DEVICE = -4,        /* no CUDA device: there is deliberately no CPU fallback                      */
  MVAL_ERR_OUT_OF_MEMORY = -5
};

/* ABI version of the loaded library (== MVAL_ABI_VERSION it was built with). */
int mval_version(void);
/* Message of the last error on this thread ("" if none). */
const char* mval_last_error(void);
/* Number of kernel launches issued through this library by the calling process so far (bench bookkeeping). */
uint64_t mval_launch_count(void);

/* The persistent kernels (mval_score_pool*, the 64 x 64 forms of mval_decode_softargmax / mval_score_hp /
 * mval_score_peaks / mval_score_xe) guard every mbarrier wait with a watchdog: a wait that exceeds ~5 s of SM clocks
 * makes all roles drain and leaves a record in pinned host memory.  Because the entry points are asynchronous the
 * failure is reported by the NEXT call into the library that launches such a kernel, or by mval_check_async, as
 * MVAL_ERR_CUDA (mval_last_error() names the wait); the record is cleared by that report and later launches run
 * normally.  mval_check_async synchronises `stream` and returns the status of everything launched on it so far --
 * call it before consuming results on the host (the Python layer does so at its own synchronisation points).
 * The reference has no counterpart: its per-frame calls are synchronous and raise in place. */
int mval_check_async(void* stream);
/* Test hook: time-out of the watchdog in SM clocks (0 = default) and, with stall != 0, producers that issue no copies
 * (every consumer wait then times out).  Affects the current device until called again with (0, 0). */
int mval_debug_watchdog(uint64_t timeout_cycles, int stall);

/* ------------------------------------------------------------------------------------------------------
 * (1) heat-map decode
 * ---------------------------------------------------------------------------------------------------- */

/* Replaces utils/evaluation.py:13-30 get_scaled_pred_corrdinates (one call per frame there, V*J argmax
 * launches + .item() syncs).  Per map: c = first flat index of the maximum (NaN counts as maximum, as in
 * torch.argmax); x = (c % H) * stride, y = (c / H) * stride -- the reference divides by shape[2] = H for
 * both (evaluation.py:25-26) and so do we.  Invalid joints give (0, 0) and their maps are not read.
 * out_xy   int32   device [n_frames][V][J][2]
 * out_peak float32 device [n_frames][V][J] value at the arg-max (may be NULL). */
int mval_decode_argmax(const float* heatmaps, int64_t n_frames, int V, int J, int H, int W, int stride,
                       const uint8_t* valid, int32_t* out_xy, float* out_peak, void* stream);

/* Replaces the soft-argmax branch utils/triangulation.py:191-200 (kornia.spatial_soft_argmax2d(hm,
 * normalized_coordinates=False) * stride): softmax over the whole H*W map, expectation of the pixel grid.
 * All joints are decoded (the reference ignores valid_joints on this branch).
 * out_xy float32 device [n_frames][V][J][2]. */
int mval_decode_softargmax(const float* heatmaps, int64_t n_frames, int V, int J, int H, int W, float stride,
                           float* out_xy, void* stream);

/* Replaces strategy.py:1178-1187 (_compute_hp inner loop): per map 1 - max(softmax over each ROW of the map)
 * (F.softmax without dim on a 2-D tensor = dim 1).  Maps of invalid joints are skipped and get NaN.
 * out_hp float32 device [n_frames][V][J]. */
int mval_score_hp(const float* heatmaps, int64_t n_frames, int V, int J, int H, int W, const uint8_t* valid,
                  float* out_hp, void* stream);

/* Replaces the per-map loops of strategy.py:1160-1175 (_compute_mpes, mode 0) and :1195-1209 (_compute_bsb, mode 1):
 * local peaks as skimage.feature.peak_local_max(map, min_distance=2) defines them (equal to the 5x5 window maximum,
 * strictly above the map minimum, 2 pixels off the border), then
 *   mode 0 (MPE): entropy of softmax over all peak values of the raw map (0 when there is no peak);
 *   mode 1 (BSB): |p0 - p1| of the two highest peaks of the ROW-softmaxed map (NaN with fewer than two peaks; the
 *                 reference raises IndexError there).
 * Maps of invalid joints get NaN.  W <= 64.  out_score float32 device [n_frames][V][J]. */
int mval_score_peaks(const float* heatmaps, int64_t n_frames, int V, int J, int H, int W, int mode,
                     const uint8_t* valid, float* out_score, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * (2) multi-view RANSAC + DLT triangulation and reprojection uncertainty
 * ---------------------------------------------------------------------------------------------------- */

/* Pair-subset policy for C(V,2) > n_iters (utils/triangulation.py:279-282 uses random.shuffle there).
 *   pairs != NULL : uint8 device [n_frames][J][n_iters][2], the pairs to visit, in order, for every
 *                   (frame, joint)  (the per-frame drop-in wrapper draws them with Python's random.shuffle
 *                   exactly like the reference and passes them here);
 *   pairs == NULL : counter-based subset keyed by (pair_seed, frame key, joint) with frame key = frame_keys[frame] when
 *                   given and frame_offset + frame otherwise, see
 *                   oracle/triangulation_oracle.py:pair_subset_indices for the exact arithmetic.
 * When C(V,2) <= n_iters both are ignored and all pairs are visited in lexicographic order. */
typedef struct mval_ransac_params {
  int32_t n_iters;       /* utils/triangulation.py:176 n_iters=64 */
  double epsilon;        /* :177 reprojection_error_epsilon=5 (compared with HALF the pixel distance, :377-381) */
  uint64_t pair_seed;
  int64_t frame_offset;  /* global index of frame 0 of this call (pool sharding / chunking) */
  const uint8_t* pairs;  /* optional explicit pair table, see above */
  const int64_t* frame_keys; /* optional, int64 device [n_frames]: the key of every frame in the counter-based subset
                                (instead of frame_offset + frame), e.g. pose * 2^32 + frame_id of its guid, so that the
                                subsets do not depend on how the pool is sharded or ordered */
} mval_ransac_params;

/* Replaces utils/triangulation.py:205-232 for a whole batch of frames: per valid (frame, joint)
 * _triangulate_ransac (:260-316, direct_optimization=False) = DLT on every view pair (:341-368), inlier vote
 * with err = 0.5*||kp - proj|| < epsilon over all views (:371-384), first strictly-largest inlier set wins,
 * final DLT on the sorted inlier views, score = mean error over those views.  Then per frame
 * metric = mean over valid joints (:226) and inlier_count = min over valid joints (:231).
 *
 * xy            device [n_frames][V][J][2], int32 (xy_is_float = 0, from mval_decode_argmax) or float32 (= 1)
 * out_xyz       float64 device [n_frames][J][3]   (zeros for invalid joints, :206)
 * out_reproj    float64 device [n_frames][J]      (NaN for invalid joints)                      may be NULL
 * out_inliers   int32   device [n_frames][J]      (0 for invalid joints)                        may be NULL
 * out_mask      uint32  device [n_frames][J]      bit v set = view v in the final inlier set    may be NULL
 * out_metric    float64 device [n_frames]         (NaN when a frame has no valid joint; the reference raises)
 * out_inlier_count int32 device [n_frames]        (0 when a frame has no valid joint)
 * workspace: none (out_mask doubles as scratch when given; otherwise an internal per-call buffer is
 * allocated with cudaMallocAsync on `stream`). */
int mval_triangulate_ransac(const void* xy, int xy_is_float, const double* proj, const uint8_t* valid,
                            int64_t n_frames, int V, int J, const mval_ransac_params* params, double* out_xyz,
                            double* out_reproj, int32_t* out_inliers, uint32_t* out_mask, double* out_metric,
                            int32_t* out_inlier_count, void* stream);

/* Replaces the direct_optimization=True branch of _triangulate_ransac (utils/triangulation.py:319-336:
 * scipy.optimize.least_squares(residuals, x0, loss="huber", method="trf") on the inlier views, then the mean error at
 * the refined point).  Call after mval_triangulate_ransac with its outputs: every valid (frame, joint) minimises
 * sum_v phi(0.5 ||kp_v - proj_v(x)||), phi = Huber with f_scale 1, over the views in inlier_mask, starting from xyz,
 * by a damped Newton iteration run to convergence (csrc/refine.cu).  xyz and reproj are updated in place; when
 * out_metric / out_inlier_count are given they are recomputed (inliers = out_inliers of the triangulation call).
 * out_iters int32 device [n_frames][J] (may be NULL) receives the iteration counts. */
int mval_refine_huber(const void* xy, int xy_is_float, const double* proj, const uint8_t* valid,
                      const uint32_t* inlier_mask, const int32_t* inliers, int64_t n_frames, int V, int J, double* xyz,
                      double* reproj, double* out_metric, int32_t* out_inlier_count, int32_t* out_iters, void* stream);

/* Fused pool scoring = mval_decode_argmax + mval_triangulate_ransac on device-resident heat maps
 * (what strategy.py:1036-1045 does per frame, for a batch of frames).  out_xy may be NULL. */
int mval_score_pool(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V,
                    int J, int H, int W, int stride, const mval_ransac_params* params, int32_t* out_xy,
                    double* out_xyz, double* out_reproj, int32_t* out_inliers, double* out_metric,
                    int32_t* out_inlier_count, void* stream);

/* Per-map heat-map score evaluated inside the fused pool pass (mval_score_pool_scored). */
#define MVAL_MAP_SCORE_NONE 0
#define MVAL_MAP_SCORE_HP 1  /* mval_score_hp                 (strategy.py:1178-1193) */
#define MVAL_MAP_SCORE_MPE 2 /* mval_score_peaks, mode 0      (strategy.py:1149-1176) */
#define MVAL_MAP_SCORE_BSB 3 /* mval_score_peaks, mode 1      (strategy.py:1195-1215) */

/* mval_score_pool + one of mval_score_hp / mval_score_peaks over the SAME heat maps in a single pass: what
 * strategy.py:1036-1045 (triangulation of every frame) and :1072-1094 (the HP / MPE / BSB metric of the same frame) do
 * together inside _compute_sal_dict.  On 64 x 64 maps every heat-map byte is read once instead of twice: for HP the
 * decode warps of the fused kernel evaluate the score on the staged map together with its arg-max; for MPE / BSB one
 * streaming kernel emits the score and the arg-max key-point of every staged map and the RANSAC kernels follow from the
 * 8-byte key-points (their fused launch stays selectable, MVAL_SCORED_SPLIT=0).  Other shapes run the two calls back to
 * back.  Results are bit-identical to the separate calls.
 * map_score      one of MVAL_MAP_SCORE_* (NONE: out_map_score is ignored, same as mval_score_pool)
 * out_map_score  float32 device [n_frames][V][J], NaN for invalid joints. */
int mval_score_pool_scored(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames, int V,
                           int J, int H, int W, int stride, const mval_ransac_params* params, int map_score,
                           int32_t* out_xy, double* out_xyz, double* out_reproj, int32_t* out_inliers,
                           double* out_metric, int32_t* out_inlier_count, float* out_map_score, void* stream);

/* mval_score_pool_scored over a pool whose heat maps are NOT one contiguous buffer: `n_segments` device buffers
 * (1..MVAL_MAX_SEGMENTS), segment s holding seg_frames[s] whole frames, concatenated in order -- e.g. the outputs of
 * successive pose-estimator batches (strategy.py:1024-1035), or chunk passes over a resident buffer.  seg_heatmaps /
 * seg_frames are HOST arrays (of device pointers / frame counts) read during the call.  proj, valid and every output are
 * contiguous over all n_frames = sum(seg_frames) frames.  ONE persistent launch covers the whole pool (no per-chunk
 * launch tails).  64 x 64-style shapes that fit the fused kernel only; otherwise MVAL_ERR_UNSUPPORTED (call
 * mval_score_pool_scored per segment). */
int mval_score_pool_segments(const float* const* seg_heatmaps, const int64_t* seg_frames, int n_segments, const double* proj,
                             const uint8_t* valid, int V, int J, int H, int W, int stride,
                             const mval_ransac_params* params, int map_score, int32_t* out_xy, double* out_xyz,
                             double* out_reproj, int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count,
                             float* out_map_score, void* stream);

/* Host-buffer pipeline: every data pointer is HOST memory (pinned recommended).  A handle owns n_slots (0 = 3, at most 4)
 * slots, each with its own stream and device staging buffers for chunk_frames frames (0 = about 256 MiB of heat maps);
 * mval_pipeline_score_pool streams the pool through them -- the host->device copy of one chunk overlaps the kernels of
 * the previous one -- and returns after the last result has landed in host memory (it also reports a watchdog trip of
 * any of its launches).  Nothing is allocated per call.  This is the end-to-end entry of a caller whose heat maps are on
 * the host; it covers what triangulation() / _compute_sal_dict take as flags (utils/triangulation.py:168-179,
 * strategy.py:1036-1045, 1072-1094):
 *   map_score            MVAL_MAP_SCORE_*: HP / MPE / BSB of the same maps -> out_map_score float32 host [n][V][J]
 *   use_soft_argmax      key-points by soft-arg-max (out_xy is then float32 [n][V][J][2], otherwise int32)
 *   use_reprojection_xe  out_metric = the reprojection XE of utils/triangulation.py:236-257 with `sigma`
 *   direct_optimization  Huber refinement of :319-336
 * options == NULL: arg-max decode, no extras.  One handle serves one device and one shape; it is not re-entrant. */
typedef struct mval_pipeline mval_pipeline;
typedef struct mval_pipeline_options {
  int32_t map_score;
  int32_t use_soft_argmax;
  int32_t use_reprojection_xe;
  int32_t direct_optimization;
  double sigma;
} mval_pipeline_options;
int mval_pipeline_create(int V, int J, int H, int W, int64_t chunk_frames, int n_slots, mval_pipeline** out);
int mval_pipeline_destroy(mval_pipeline* pipeline);
int mval_pipeline_score_pool(mval_pipeline* pipeline, const float* heatmaps, const double* proj, const uint8_t* valid,
                             int64_t n_frames, int stride, const mval_ransac_params* params,
                             const mval_pipeline_options* options, void* out_xy, double* out_xyz, double* out_reproj,
                             int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count, float* out_map_score);

/* mval_pipeline_score_pool(options = NULL) on an implicit handle that the library keeps per process (re-created only when
 * the device, the shape or an explicit chunk_frames changes).  chunk_frames 0 = choose. */
int mval_score_pool_host(const float* heatmaps, const double* proj, const uint8_t* valid, int64_t n_frames,
                         int V, int J, int H, int W, int stride, const mval_ransac_params* params,
                         int64_t chunk_frames, int32_t* out_xy, double* out_xyz, double* out_reproj,
                         int32_t* out_inliers, double* out_metric, int32_t* out_inlier_count);

/* Replaces utils/triangulation.py:236-257 (_compute_xe, the metric of triangulation() when use_reprojection_xe=True /
 * AL.USE_REPROJECTION_XE): per (view, joint) the triangulated 3-D joint is reprojected (P [X,1], w == 0 -> 1), a Gaussian
 * exp(-|pixel - reprojection|^2 / (2 sigma^2)) is rendered on the H x W heat-map pixel grid in float64, and
 * sum((heatmap - render)^2) / (H W) is added up over views, then joints (all joints: the reference also renders the
 * joints triangulation() left at the origin).  xyz float64 device [n_frames][J][3] (out_xyz of the calls above).
 * out_map    float64 device [n_frames][V][J] per-map terms (may be NULL)
 * out_metric float64 device [n_frames]. */
int mval_score_xe(const float* heatmaps, const double* proj, const double* xyz, int64_t n_frames, int V, int J, int H, int W,
                  double sigma, double* out_map, double* out_metric, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * (3) ranking and coreset k-center greedy selection
 * ---------------------------------------------------------------------------------------------------- */

/* Replaces strategy.py:932-949: drop NaN scores, then heapq.nlargest(k, ...) = descending by score, ties by
 * ascending pool index.  Writes min(k, #non-NaN) entries; *out_count (device int32) receives that number; the
 * remaining slots of out_idx / out_val are set to -1 / NaN.
 * scores float64 device [n]; out_idx int64 device [k] (global index = index_offset + local);
 * out_val float64 device [k]. */
int mval_topk_desc(const double* scores, int64_t n, int64_t index_offset, int32_t k, int64_t* out_idx,
                   double* out_val, int32_t* out_count, void* stream);

/* The cross-rank half of the same ranking: every rank's mval_topk_desc output (scores descending, GLOBAL indices, unused
 * slots NaN / -1) is all-gathered rank after rank; this call merges the n = world * k candidates on the device into the
 * global top-k.  With contiguous frame sharding "rank-major, then local order" is ascending pool order among equal scores,
 * so the stable sort reproduces nlargest's tie-break.  out_idx[i] = indices[position of the i-th best candidate]. */
int mval_topk_merge(const double* scores, const int64_t* indices, int64_t n, int32_t k, int64_t* out_idx, double* out_val,
                    int32_t* out_count, void* stream);

/* Dict-insertion semantics of the guid-keyed tables strategy.py:1115-1133 builds (sal_dict[...][guid] = value, guid =
 * "%s-%s" % (pose, frame_id)) for rows that stay on the device: a guid that occurs in several rows (DistributedSampler
 * pads the last batch by repeating frames, strategy.py:753) keeps the position of its FIRST row and the value of its LAST.
 * pose / frame int64 device [n], both in [0, 2^32).
 * out_keep   uint8 device [n]  1 = first row of its guid
 * out_src    int32 device [n]  for kept rows: the last row with the same guid (take the values from there); -1 otherwise
 * out_unique int32 device [1]  number of distinct guids, or -1 when a pose / frame id does not fit 32 bits (the caller
 *                              then falls back to building the dict on the host). */
int mval_first_occurrence(const int64_t* pose, const int64_t* frame, int64_t n, uint8_t* out_keep, int32_t* out_src,
                          int32_t* out_unique, void* stream);

/* Replaces the frame aggregation of strategy.py:1151-1158 (_compute_mpe), :1188-1193 (_compute_hp) and :1210-1215
 * (_compute_bsb) for a whole pool: the frame score from the per-map scores over (view, valid joint), view-major, with the
 * reference's own arithmetic -- HP: Python floats (AVG = builtin sum() / len in double, Neumaier-compensated iff
 * compensated_sum != 0, i.e. iff the interpreter is Python >= 3.12; STD = np.std in float64); MPE / BSB: np.float32 scalars
 * (AVG = sequential float32 adds / len; STD = np.std in float32); np.std with NumPy's pairwise summation.  Bit-identical to
 * the reference's expressions (the host restatement strategy._aggregate_map_scores is pinned to its golden values).
 * per_map float32 device [n_frames][V][J] (mval_score_hp / mval_score_peaks / mval_score_pool_scored); valid uint8 device
 * [n_frames][J] or NULL; kind MVAL_MAP_SCORE_HP / MPE / BSB; config_std 0 = AVG, 1 = STD.
 * out float64 device [n_frames] (float32-typed results widened; NaN for a frame without a valid joint).  J <= 128. */
int mval_aggregate_map_scores(const float* per_map, const uint8_t* valid, int64_t n_frames, int V, int J, int kind,
                              int config_std, int compensated_sum, double* out, void* stream);

/* Replaces strategy.py:957-975 (the pseudo-label candidate filter and its sort): frames with a non-NaN sal_metric,
 * inlier_count > inlier_threshold (SAL.INLIER_THRESHOLD) and excluded[i] == 0 (the caller marks frames already picked by
 * the AL step or already pseudo-labelled), in ascending sal_metric order, ties in pool order (sorted() is stable).
 * sal_metric / inlier_count float32 device [n] (strategy.py:1061-1063 stores both as float32); excluded uint8 device [n]
 * or NULL.  Writes the first min(k, #candidates) pool indices to out_idx (int64 device [k]) and their number to
 * *out_count (device int32). */
int mval_sal_rank(const float* sal_metric, const float* inlier_count, const uint8_t* excluded, int64_t n,
                  float inlier_threshold, int32_t k, int64_t* out_idx, int32_t* out_count, void* stream);

/* Replaces the per-candidate self.kmeans.predict([kp])[0] of the cluster-balanced pseudo-label walk (strategy.py:981-989)
 * for a whole pool: kp = root-relative pose (pose^T[0:3, :] - pose^T[0:3, root]).flatten() in float64 from the float32
 * predictions of sal_dict["pred_3d_keypoints"], label = argmin_c (|c|^2 - 2 kp.c), first centre on ties -- sklearn
 * KMeans.predict.  pred float32 device [n_frames][J][3]; centres float64 device [k][3 J] (kmeans.cluster_centers_).
 * out_label  int32   device [n_frames]
 * out_margin float64 device [n_frames] runner-up score minus best score (may be NULL): the caller re-checks frames whose
 *            margin is within rounding distance of 0 with sklearn itself, so the labels are exactly the reference's. */
int mval_kmeans_assign(const float* pred, int64_t n_frames, int J, int root, const double* centres, int k,
                       int32_t* out_label, double* out_margin, void* stream);

/* Replaces utils/evaluation.py:198-208 compute_mkpe([pred], [gt], [valid]) per frame (strategy.py:1134-1145), float32:
 * mean over joints of sqrt(sum_c where(valid, (pred - gt)^2, 0)) / valid  (NaN as soon as one joint is invalid, like the
 * reference).  pred float32 device [n][J][3]; gt float32 device [n][gt_rows][J], rows 0..2 = x, y, z
 * (dataset/dataset.py:148 stores 4 rows); valid float32 device [n][J]; out_mkpe float32 device [n]. */
int mval_mkpe(const float* pred, const float* gt, const float* valid, int64_t n_frames, int J, int gt_rows, float* out_mkpe,
              void* stream);

/* Replaces utils/coreset.py:35-47 (_compute_stacked_features) for poses that are already on the device: row f =
 * (x_0..x_{J-1}, y.., z..) of frame f relative to joint `root`, from the float32 roundings of the joints
 * (strategy.py:1046), subtracted in float64 and rounded to float32.  xyz device [n][J][3], float64 (xyz_is_double = 1) or
 * float32; out_features float32 device [n][3 J]. */
int mval_pose_features(const void* xyz, int xyz_is_double, int64_t n_frames, int J, int root, float* out_features,
                       void* stream);

/* Coreset k-center greedy (utils/coreset.py:49-95) on float32 features.  features float32 device [n][d] row-major.
 * The distance is sklearn's expansion evaluated in float32 in the canonical order that oracle/coreset_oracle.c
 * defines (one fma chain over k ascending per (row, centre); d2 = ((-2 dot) + |x|^2) + |c|^2; sqrt(max(d2, 0))),
 * so selected indices and running minima are bit-reproducible on the CPU.  Ties: lowest index (np.argmax, :90).
 *
 *   mval_kcenter_norms        : row_norms[i] = dot(x_i, x_i)
 *   mval_kcenter_update       : one centre (coreset.py:64-69 with a single cluster centre):
 *                               min_dist[i] = min(min_dist[i], dist(x_i, centre)), then the arg-max of the updated
 *                               min_dist -> out_best_val float32 device[1], out_best_idx int64 device[1] (index_offset
 *                               added; -1 for n = 0) = the next `ind` of coreset.py:90.  `centre` float32 device [d].
 *   mval_kcenter_update_batch : the same fold for n_centres centres in ONE pass over the features (what coreset.py:83-84
 *                               does for the labeled set, and what a round of greedy picks needs).  centres float32
 *                               device [n_centres][d], centre_norms their canonical squared norms.  flags: 0 = choose
 *                               (tcgen05 TF32 screening GEMM + exact recheck when d is large and 16-byte aligned,
 *                               register-tiled FFMA pass otherwise), 1 = force the FFMA pass, 2 = force the tensor-core
 *                               path whenever it is applicable; + 4 = the centres are the picks of one greedy round (all
 *                               256-centre chunks are screened first and share one exact recheck).  The result does not
 *                               depend on the path. */
int mval_kcenter_norms(const float* features, int64_t n, int d, float* row_norms, void* stream);
int mval_kcenter_update(const float* features, const float* row_norms, int64_t n, int d, const float* centre,
                        float* min_dist, int64_t index_offset, float* out_best_val, int64_t* out_best_idx,
                        void* stream);
int mval_kcenter_update_batch(const float* features, const float* row_norms, int64_t n, int d, const float* centres,
                              const float* centre_norms, int n_centres, float* min_dist, int flags, void* stream);

/* Diagnostics of the tensor-core path: number of (row, centre) pairs that survived the LAST screening GEMM on the
 * current device and had to be re-evaluated exactly, and the capacity of the survivor list (beyond it the FFMA pass
 * takes over).  Host pointers; synchronises `stream`. */
int mval_kcenter_tc_stats(uint64_t* survivors, uint64_t* capacity, void* stream);

/* The greedy loop of coreset.py:86-93 in ROUNDS (csrc/kcenter.cu header): each round yields the next T >= 1 picks of
 * the sequential algorithm, exactly, for one pass over the features.  Multi-GPU (rows sharded contiguously): every
 * rank calls select on its shard, the record blocks are all-gathered (the only exchange, once per round), every rank
 * calls resolve on the gathered blocks -- all ranks compute the same picks -- and update_batch on its own shard.
 *
 *   mval_kcenter_records_bytes : size of one shard's record block for k_slots candidates of dimension d
 *   mval_kcenter_select        : writes this shard's block: up to k_slots - 1 rows whose running minimum is above the
 *                                shard's k_slots-th largest value kappa (all rows if n < k_slots), each with value,
 *                                global index (index_offset + local), squared norm and feature row, plus kappa
 *   mval_kcenter_resolve       : records = n_blocks consecutive blocks.  Replays the greedy loop on the union of the
 *                                candidates while the best candidate is provably the arg-max of the whole pool (value
 *                                above every shard's kappa; the first pick always is).  Writes the global indices of
 *                                the picks to selected_out (int64 device [max_picks]), their rows and squared norms
 *                                to centres_out (float32 device [max_picks][d]) / centre_norms_out, and the number of
 *                                picks (1..max_picks) to *n_picks_host (HOST int32) -- this call synchronises `stream`.
 *                                workspace: device, mval_kcenter_resolve_workspace_bytes(n_blocks, k_slots, d) bytes.
 *                                n_blocks * k_slots <= 1024. */
size_t mval_kcenter_records_bytes(int k_slots, int d);
int mval_kcenter_select(const float* features, const float* row_norms, const float* min_dist, int64_t n, int d,
                        int64_t index_offset, int k_slots, void* records_out, void* stream);
size_t mval_kcenter_resolve_workspace_bytes(int n_blocks, int k_slots, int d);
int mval_kcenter_resolve(const void* records, int n_blocks, int k_slots, int d, int max_picks, void* workspace,
                         float* centres_out, float* centre_norms_out, int64_t* selected_out, int32_t* n_picks_host,
                         void* stream);

/* The same round WITHOUT the host in the loop: `state` is device int32 [4] = {picks made so far, picks of this round (out),
 * budget, reserved}, initialised by the caller to {0, 0, budget, 0}.  resolve_async limits the round to budget - done picks,
 * writes them to selected_out[done ..] (int64 device [budget]), advances state[0] / sets state[1] and does not synchronise;
 * update_batch_dev folds the first *n_centres (device int32, e.g. state + 1) of `max_centres` centre rows into min_dist.  A
 * round launched after the budget is reached does nothing, so the host may read state[0] late (e.g. one round behind).
 * The result is identical to the synchronous calls. */
int mval_kcenter_resolve_async(const void* records, int n_blocks, int k_slots, int d, void* workspace, float* centres_out,
                               float* centre_norms_out, int64_t* selected_out, int32_t* state, void* stream);
int mval_kcenter_update_batch_dev(const float* features, const float* row_norms, int64_t n, int d, const float* centres,
                                  const float* centre_norms, int max_centres, const int32_t* n_centres, float* min_dist,
                                  int flags, void* stream);

/* Single-device selection (coreset.py:71-95): norms, the labeled rows [n_unlabeled, n) folded in, then `budget`
 * picks in rounds driven by device-side counters (the host reads the running total one round late).  min_dist float32
 * device [n] (out), out_selected int64 device [budget].  Synchronises `stream` before returning. */
int mval_kcenter_greedy(const float* features, int64_t n, int64_t n_unlabeled, int d, int32_t budget,
                        float* min_dist, int64_t* out_selected, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * synthetic pools (benchmark / test support)
 * ---------------------------------------------------------------------------------------------------- */

/* Renders float32 [n_maps][H][W] heat maps: Gaussian bump (sigma) at centres[m] = (x, y) heat-map pixels plus
 * uniform noise of the given amplitude from a counter-based hash of (seed, element index). */
int mval_synth_heatmaps(const float* centres, int64_t n_maps, int H, int W, float sigma, float noise,
                        uint64_t seed, float* out_heatmaps, void* stream);

/* Replaces dataset/dataset.py:198-207 (the ground-truth heat maps a frame is trained against, the format on the other side
 * of the pose estimator): points float64 device [n_maps][2] = the joint's projection divided by the heat-map stride
 * (x, y); out = exp(-((x - px)^2 + (y - py)^2) / (2 sigma^2)) on the H x W pixel grid, evaluated in float64 like the
 * reference (a float32 grid minus a float64 label promotes to float64).  out_f64 float64 device [n_maps][H][W] and / or
 * out_f32 (the same values rounded to float32, what the loss consumes); either may be NULL. */
int mval_render_gt_heatmaps(const double* points, int64_t n_maps, int H, int W, double sigma, double* out_f64, float* out_f32,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVAL_B200_H_ */
