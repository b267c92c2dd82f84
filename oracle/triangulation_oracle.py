"""CPU oracle (numpy, float64) for heatmap decode + RANSAC/DLT triangulation + reprojection score.

TEST INFRASTRUCTURE, NOT PRODUCT (see ``oracle/__init__.py``).

Restates, vectorised over (frame, joint, view pair), what the upstream reference computes one
(frame, joint, pair) at a time:

  decode_argmax        <- utils/evaluation.py:13-30   (get_scaled_pred_corrdinates)
  decode_softargmax    <- utils/triangulation.py:191-200 + kornia.spatial_soft_argmax2d  [parity unpinned:
                          kornia is not installed anywhere we can run; restated from kornia's dsnt docs]
  dlt_solve            <- utils/triangulation.py:341-368 (_triangulate_dlt) + :387-399
  reprojection_error   <- utils/triangulation.py:371-384, :459-484, :408-430
  ransac_pool          <- utils/triangulation.py:260-316 (_triangulate_ransac)
  refine_huber         <- utils/triangulation.py:319-336 (direct_optimization=True: scipy least_squares, Huber, trf)
  triangulate_pool     <- utils/triangulation.py:168-233 (triangulation) applied to every frame of a pool

The SVD is numpy's (LAPACK gesdd), exactly the routine the reference calls (:363), on matrices of
exactly the shapes the reference builds (4x4 per pair, 2n x 4 for the n sorted inlier views), so the
oracle reproduces the reference to rounding; ``tests/test_oracle_golden.py`` pins it against the
outputs of the real reference stored in ``tests/golden/``.

View-pair subsets (reference :279-282).  For C(V,2) <= n_iters the pairs are all lexicographic
pairs, in order.  Above that the reference draws ``random.shuffle`` from Python's process-global
Mersenne Twister, once per (frame, valid joint): not reproducible across ranks or runs even in the
reference itself.  The pool path defines a counter-based replacement (``pair_subset``) keyed by
(seed, global frame index, joint) that the CUDA kernel implements identically;
``oracle/make_golden.py`` runs the *reference* with its ``random`` module attribute swapped for a
shim that serves the same subsets, so the golden fixtures for V >= 12 are still produced by the
reference's own code.
"""
from itertools import combinations

import numpy as np

_MASK64 = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15


def lexicographic_pairs(n_views):
    """All (a, b), a < b, in itertools.combinations order (reference :279)."""
    return np.array(list(combinations(range(n_views), 2)), dtype=np.int64).reshape(-1, 2)


def _splitmix64(state):
    state = (state + _GOLDEN) & _MASK64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    z ^= z >> 31
    return state, z


def pair_subset_indices(n_total, n_take, seed, frame_index, joint):
    """Counter-based partial Fisher-Yates: the first ``n_take`` entries of a pseudo-random permutation of
    range(n_total).  Identical arithmetic to ``mval::pair_subset`` in csrc/triangulate.cu."""
    state = (int(seed) + _GOLDEN * (int(frame_index) * 64 + int(joint) + 1)) & _MASK64
    idx = list(range(n_total))
    for i in range(n_take):
        state, z = _splitmix64(state)
        r = i + (((z >> 32) * (n_total - i)) >> 32)
        idx[i], idx[r] = idx[r], idx[i]
    return idx[:n_take]


def pair_table(n_views, n_iters, seed=0, frame_index=0, joint=0):
    """Pairs RANSAC visits for one (frame, joint), in visiting order. [n_pairs, 2] int64."""
    allp = lexicographic_pairs(n_views)
    if len(allp) <= n_iters:
        return allp
    return allp[pair_subset_indices(len(allp), n_iters, seed, frame_index, joint)]


class DeterministicShuffle:
    """Stand-in for the ``random`` module attribute of the reference's utils/triangulation.py (:8, :281).

    ``shuffle(lst)`` reorders ``lst`` so that its first ``n_iters`` entries are the pairs
    ``pair_table`` returns for the next (frame, joint) in ``schedule`` (a list consumed in call order,
    i.e. frame-major, valid joints only -- the order the reference calls _triangulate_ransac in).
    """

    def __init__(self, seed, schedule, n_iters=64):
        self.seed, self.schedule, self.n_iters, self.calls = seed, list(schedule), n_iters, 0

    def shuffle(self, lst):
        frame_index, joint = self.schedule[self.calls]
        self.calls += 1
        take = pair_subset_indices(len(lst), self.n_iters, self.seed, frame_index, joint)
        chosen = [lst[i] for i in take]
        rest = [p for i, p in enumerate(lst) if i not in set(take)]
        lst[:] = chosen + rest


# ----------------------------------------------------------------------------------------------
# decode
# ----------------------------------------------------------------------------------------------
def decode_argmax(heatmaps, stride, valid=None):
    """heatmaps [..., J, H, W] float32 -> int64 [..., J, 2] = (x, y) * stride.

    Reference utils/evaluation.py:24-27: flat argmax (first maximum; NaN counts as maximum, as in
    torch.argmax), x = corr % shape[2], y = corr // shape[2] -- shape[2] is H, replicated literally.
    Invalid joints give [0, 0] (:21-23).  ``valid`` broadcasts against heatmaps.shape[:-2]'s joint axis.
    """
    hm = np.asarray(heatmaps)
    h, w = hm.shape[-2:]
    flat = np.argmax(hm.reshape(hm.shape[:-2] + (h * w,)), axis=-1).astype(np.int64)
    out = np.stack([(flat % h) * stride, (flat // h) * stride], axis=-1)
    if valid is not None:
        v = np.asarray(valid).astype(bool)
        # valid is [J] or [N, J]; heatmaps are [V, J, H, W] or [N, V, J, H, W]
        if v.ndim == 2:
            v = v[:, None, :]
        out = np.where(np.broadcast_to(v, flat.shape)[..., None], out, 0)
    return out


def decode_softargmax(heatmaps, stride):
    """[..., H, W] float32 -> float32 [..., 2] (x, y) * stride.  kornia itself is not installable here; pinned against
    the reference run with kornia's spatial_expectation2d / create_meshgrid as vendored verbatim by `transformers`
    (tests/golden/softargmax_v5_j6.npz, oracle/make_golden.py:case_softargmax) to float32 rounding (4.6e-5 px).

    kornia.spatial_soft_argmax2d(hm, temperature=1, normalized_coordinates=False): softmax over the
    flattened H*W map, then the expectation of the pixel grid (x = column 0..W-1, y = row 0..H-1);
    the reference multiplies by stride in float32 (utils/triangulation.py:193-197).  Computed here in
    float64 and rounded once; the CUDA kernel is compared with a tolerance stated in the tests.
    """
    hm = np.asarray(heatmaps, dtype=np.float64)
    h, w = hm.shape[-2:]
    flat = hm.reshape(hm.shape[:-2] + (h * w,))
    e = np.exp(flat - flat.max(axis=-1, keepdims=True))
    p = e / e.sum(axis=-1, keepdims=True)
    p = p.reshape(hm.shape)
    ex = (p.sum(axis=-2) * np.arange(w)).sum(axis=-1)
    ey = (p.sum(axis=-1) * np.arange(h)).sum(axis=-1)
    return (np.stack([ex, ey], axis=-1).astype(np.float32) * np.float32(stride)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# DLT + reprojection
# ----------------------------------------------------------------------------------------------
def _dehomogenise(vec):
    """Reference :387-399: divide by the last component, a last component of exactly 0 is replaced by 1."""
    w = vec[..., -1:]
    w = np.where(w == 0, np.ones_like(w), w)
    return vec[..., :-1] / w


def dlt_solve(P, pts):
    """P [..., n, 3, 4] f64, pts [..., n, 2] -> X [..., 3].  Reference :341-368.

    Rows 2j / 2j+1 of A are  u_j * P_j[2] - P_j[0]  and  v_j * P_j[2] - P_j[1]; the solution is the last
    row of vh from numpy's reduced SVD, de-homogenised.
    """
    P = np.asarray(P, dtype=np.float64)
    pts = np.asarray(pts, dtype=np.float64)
    rows_u = pts[..., 0:1] * P[..., 2, :] - P[..., 0, :]
    rows_v = pts[..., 1:2] * P[..., 2, :] - P[..., 1, :]
    A = np.stack([rows_u, rows_v], axis=-2)  # [..., n, 2, 4]
    A = A.reshape(A.shape[:-3] + (A.shape[-3] * 2, 4))
    _, _, vh = np.linalg.svd(A, full_matrices=False)
    return _dehomogenise(vh[..., 3, :])


def reprojection_error(X, pts, P):
    """X [..., 3], pts [..., n, 2], P [..., n, 3, 4] -> err [..., n] = 0.5 * ||pts - proj(X)||.

    Reference :371-384 with :459-484 ([X,1] @ P^T, divide by w with the w == 0 -> 1 rule).
    """
    Xh = np.concatenate([X, np.ones(X.shape[:-1] + (1,))], axis=-1)  # :408-418
    proj_h = np.matmul(Xh[..., None, None, :], np.swapaxes(P, -1, -2))[..., 0, :]  # [..., n, 3]
    proj = _dehomogenise(proj_h)
    return 0.5 * np.sqrt(np.sum((np.asarray(pts, dtype=np.float64) - proj) ** 2, axis=-1))


def huber_cost(X, pts, P):
    """0.5 * sum rho(r^2) with rho = Huber, f_scale = 1: the objective scipy minimises for reference :319-330."""
    r = reprojection_error(X, pts, P)
    z = r * r
    return 0.5 * np.sum(np.where(z <= 1.0, z, 2.0 * np.sqrt(z) - 1.0), axis=-1)


def refine_huber(P_inl, pts_inl, x0):
    """Reference :319-336 for one joint: the very call the reference makes (same residual function, same scipy
    routine and defaults), then the mean error at the refined point.  Returns (x [3], reproj_mean)."""
    from scipy.optimize import least_squares

    P_inl = np.asarray(P_inl, dtype=np.float64)
    pts_inl = np.asarray(pts_inl, dtype=np.float64)

    def residual_function(x):
        return reprojection_error(np.array([x]), pts_inl[None], P_inl[None])[0]

    res = least_squares(residual_function, np.array(x0, dtype=np.float64), loss="huber", method="trf")
    return res.x, float(np.mean(residual_function(res.x)))


def ransac_pool(P, pts, valid, pairs, eps=5.0, direct_optimization=False):
    """RANSAC over view pairs + final DLT on the inlier set, for every valid (frame, joint).

    P [N, V, 3, 4] f64; pts [N, V, J, 2]; valid [N, J] bool; pairs [n_pairs, 2] (shared) or
    [N, J, n_pairs, 2].  Returns keypoints_3d [N, J, 3] (zeros for invalid joints, reference :206),
    reproj_mean [N, J] (nan for invalid), inliers [N, J] int64 (0 for invalid), inlier_mask [N, J] uint32.
    Reference :284-316: a pair's inlier set is the pair itself plus every view with error < eps; the
    first pair with the strictly largest set wins; the final solve uses the sorted inlier views and the
    score is the mean error over exactly those views.
    """
    P = np.asarray(P, dtype=np.float64)
    pts = np.asarray(pts)
    valid = np.asarray(valid).astype(bool)
    N, V, J, _ = pts.shape
    assert V >= 2 and V <= 32
    pairs = np.asarray(pairs, dtype=np.int64)
    if pairs.ndim == 2:
        pairs = np.broadcast_to(pairs, (N, J) + pairs.shape)
    n_pairs = pairs.shape[2]

    kp3d = np.zeros((N, J, 3))
    reproj_mean = np.full((N, J), np.nan)
    inliers = np.zeros((N, J), dtype=np.int64)
    inlier_mask = np.zeros((N, J), dtype=np.uint32)
    nn, jj = np.nonzero(valid)
    if len(nn) == 0:
        return kp3d, reproj_mean, inliers, inlier_mask
    M = len(nn)
    Pm = P[nn]  # [M, V, 3, 4]
    pm = pts[nn, :, jj, :].astype(np.float64)  # [M, V, 2]
    pr = pairs[nn, jj]  # [M, n_pairs, 2]

    # candidate from every pair
    mi = np.arange(M)[:, None, None]
    P_pair = Pm[mi, pr]  # [M, n_pairs, 2, 3, 4]
    p_pair = pm[mi, pr]  # [M, n_pairs, 2, 2]
    X = dlt_solve(P_pair, p_pair)  # [M, n_pairs, 3]
    err = reprojection_error(X, pm[:, None], Pm[:, None])  # [M, n_pairs, V]
    member = err < eps
    mj = np.arange(M)[:, None]
    pj = np.arange(n_pairs)[None, :]
    member[mj, pj, pr[..., 0]] = True
    member[mj, pj, pr[..., 1]] = True
    counts = member.sum(axis=-1)
    best = np.argmax(counts, axis=-1)  # first maximum == "strictly greater replaces"
    best_member = member[np.arange(M), best]  # [M, V]
    n_in = best_member.sum(axis=-1)

    Xf = np.zeros((M, 3))
    rm = np.zeros(M)
    for c in np.unique(n_in):
        sel = np.nonzero(n_in == c)[0]
        views = np.nonzero(best_member[sel])[1].reshape(len(sel), c)  # sorted ascending
        Pi = Pm[sel[:, None], views]
        pi = pm[sel[:, None], views]
        Xc = dlt_solve(Pi, pi)
        ec = reprojection_error(Xc, pi, Pi)
        if direct_optimization:
            for i in range(len(sel)):
                Xc[i], m_i = refine_huber(Pi[i], pi[i], Xc[i])
                ec[i] = m_i  # broadcast: the mean below returns m_i
        Xf[sel] = Xc
        rm[sel] = np.mean(ec, axis=-1)
    kp3d[nn, jj] = Xf
    reproj_mean[nn, jj] = rm
    inliers[nn, jj] = n_in
    inlier_mask[nn, jj] = (best_member.astype(np.uint64) << np.arange(V, dtype=np.uint64)).sum(axis=-1).astype(np.uint32)
    return kp3d, reproj_mean, inliers, inlier_mask


def triangulate_pool(heatmaps, P, stride, valid, n_iters=64, eps=5.0, pair_seed=0, frame_offset=0,
                     use_soft_argmax=False, keypoints_2d=None, chunk=256, direct_optimization=False, frame_keys=None):
    """Pool-level restatement of reference utils/triangulation.py:168-233 (use_reprojection_xe=False;
    direct_optimization=True adds the Huber refinement of :319-336 through scipy, exactly as the reference does).

    heatmaps [N, V, J, H, W] f32 (or None when ``keypoints_2d`` [N, V, J, 2] is given); P [N, V, 3, 4];
    valid [N, J].  Returns dict with keypoints_3d [N,J,3] f64, keypoints_2d, metric [N] f64 (mean over valid
    joints of the per-joint mean inlier reprojection error, :226), inlier_count [N] int64 (min over valid
    joints, :231), plus per-joint reproj_mean / inliers / inlier_mask.  Frames without any valid joint
    get metric = nan and inlier_count = 0 (the reference raises there: np.min of an empty list).
    """
    P = np.asarray(P, dtype=np.float64)
    valid = np.asarray(valid).astype(bool)
    if keypoints_2d is None:
        hm = np.asarray(heatmaps)
        if use_soft_argmax:
            keypoints_2d = decode_softargmax(hm, stride)
        else:
            keypoints_2d = decode_argmax(hm, stride, valid)
    keypoints_2d = np.asarray(keypoints_2d)
    N, V, J, _ = keypoints_2d.shape
    n_all = V * (V - 1) // 2
    outs = [[], [], [], []]
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        if n_all <= n_iters:
            pairs = lexicographic_pairs(V)
        else:
            allp = lexicographic_pairs(V)
            pairs = np.zeros((e - s, J, n_iters, 2), dtype=np.int64)
            for n in range(s, e):
                for j in range(J):
                    if valid[n, j]:
                        pairs[n - s, j] = allp[pair_subset_indices(
                            n_all, n_iters, pair_seed, frame_offset + n if frame_keys is None else int(frame_keys[n]), j)]
        r = ransac_pool(P[s:e], keypoints_2d[s:e], valid[s:e], pairs, eps, direct_optimization)
        for o, x in zip(outs, r):
            o.append(x)
    kp3d, reproj_mean, inliers, inlier_mask = [np.concatenate(o, axis=0) for o in outs]
    metric = np.full(N, np.nan)
    inlier_count = np.zeros(N, dtype=np.int64)
    for n in range(N):
        v = valid[n]
        if v.any():
            metric[n] = np.mean(reproj_mean[n, v])
            inlier_count[n] = np.min(inliers[n, v])
    return {
        "keypoints_3d": kp3d,
        "keypoints_2d": keypoints_2d,
        "metric": metric,
        "inlier_count": inlier_count,
        "reproj_mean": reproj_mean,
        "inliers": inliers,
        "inlier_mask": inlier_mask,
    }


def compute_xe(keypoints_3d, P, heatmaps, sigma):
    """Restatement of reference utils/triangulation.py:236-257 (_compute_xe) for a pool.

    keypoints_3d [N, J, 3] f64 (zeros for invalid joints -- the reference renders those too), P [N, V, 3, 4] f64,
    heatmaps [N, V, J, H, W] f32.  Per (view, joint): project the 3-D point (:240-242, :459-484; w == 0 -> 1), render
    exp(-((x - u)^2 + (y - v)^2) / (2 sigma^2)) on the H x W pixel grid IN FLOAT64 (float32 grid minus float64 point
    promotes), and add sum((pred - render)^2) / (1 * H * W) (pose_estimators/loss.py:14-20; float32 pred up-cast).
    The reference compares the HEATMAP pixel grid with a point in IMAGE pixels (no division by the stride); kept.
    Returns (metric [N] f64 = sum over views then joints in that order, per_map [N, V, J] f64)."""
    kp = np.asarray(keypoints_3d, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    hm = np.asarray(heatmaps)
    N, V, J, H, W = hm.shape
    per_map = np.zeros((N, V, J))
    xs = np.arange(W, dtype=np.float32).astype(np.float64)
    ys = np.arange(H, dtype=np.float32).astype(np.float64)
    for n in range(N):
        Xh = np.hstack([kp[n], np.ones((J, 1))])
        for v in range(V):
            ph = Xh @ P[n, v].T
            z = np.where(ph[:, 2] == 0, 1.0, ph[:, 2])
            uv = ph[:, :2] / z[:, None]
            for j in range(J):
                expo = (xs[None, :] - uv[j, 0]) ** 2 + (ys[:, None] - uv[j, 1]) ** 2
                render = np.exp(-expo / (2.0 * sigma ** 2))
                per_map[n, v, j] = np.sum((hm[n, v, j].astype(np.float64) - render) ** 2) / (H * W)
    metric = np.zeros(N)
    for n in range(N):
        acc = 0.0
        for v in range(V):
            for j in range(J):
                acc += per_map[n, v, j]
        metric[n] = acc
    return metric, per_map
