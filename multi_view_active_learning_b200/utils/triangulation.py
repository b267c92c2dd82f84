"""Multi-view triangulation with the reference's entry point (utils/triangulation.py:168-233) plus the batched
pool-level entry the per-frame one is a view of."""
import itertools
import random

import numpy as np
import torch

from .. import ops


def _draw_pairs_like_reference(n_views, n_joints, valid, n_iters):
    """utils/triangulation.py:279-282, once per valid joint in joint order: consumes Python's global ``random``
    state exactly as the reference does, so a seeded caller sees the same view-pair subsets."""
    table = np.zeros((1, n_joints, n_iters, 2), dtype=np.uint8)
    for j in range(n_joints):
        if not valid[j]:
            continue
        view_pairs = list(itertools.combinations(set(range(n_views)), 2))
        random.shuffle(view_pairs)
        table[0, j] = np.asarray(view_pairs[:n_iters], dtype=np.uint8)
    return torch.from_numpy(table)


def triangulation(
    heatmaps,
    proj_matricies,
    stride,
    valid_joints,
    use_soft_argmax=False,
    use_reprojection_xe=False,
    sigma=None,
    n_iters=64,
    reprojection_error_epsilon=5,
    direct_optimization=False,
):
    """Same contract as the reference: heatmaps [V, J, H, W], proj_matricies [V, 3, 4], valid_joints [J] ->
    {"keypoints_3d": np.float64 [J, 3], "keypoints_2d": np [V, J, 2], "metric": np.float64, "inlier_count": np.int64}.
    """
    if not torch.is_tensor(heatmaps):
        heatmaps = torch.as_tensor(np.asarray(heatmaps))
    if len(proj_matricies) != heatmaps.shape[0]:
        raise AssertionError("len(proj_matricies) != number of views")  # reference :267
    if heatmaps.shape[0] < 2:
        raise AssertionError("need at least 2 views")  # reference :268
    hm = heatmaps if heatmaps.is_cuda else heatmaps.cuda()
    V, J = hm.shape[0], hm.shape[1]
    valid = np.asarray([bool(valid_joints[j]) for j in range(J)])
    P = torch.as_tensor(np.asarray(proj_matricies.detach().cpu() if torch.is_tensor(proj_matricies) else proj_matricies))
    P = P.double().unsqueeze(0)
    hm5 = hm.float().unsqueeze(0)
    if use_soft_argmax:
        kp = ops.decode_softargmax(hm5, stride)
    else:
        kp = ops.decode_argmax(hm5, stride, torch.from_numpy(valid))
    pairs = None
    if V * (V - 1) // 2 > n_iters:
        pairs = _draw_pairs_like_reference(V, J, valid, n_iters).to(hm.device)
    if not valid.any():
        # np.min([]) in the reference (:231)
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    out = ops.triangulate_ransac(kp, P, torch.from_numpy(valid), n_iters, float(reprojection_error_epsilon), pairs=pairs,
                                 direct_optimization=bool(direct_optimization))
    kp_np = kp[0].cpu().numpy()
    if use_reprojection_xe:
        # reference :223-224: the metric becomes the 0-d CUDA tensor _compute_xe returns (float64 by promotion)
        metric = ops.score_xe(hm5, P, out["keypoints_3d"], sigma)[0]
    else:
        metric = np.float64(out["metric"][0].item())
    return {
        "keypoints_3d": out["keypoints_3d"][0].cpu().numpy(),
        "keypoints_2d": kp_np if use_soft_argmax else kp_np.astype(np.int64),
        "metric": metric,
        "inlier_count": np.int64(out["inlier_count"][0].item()),
    }


def triangulation_batch(heatmaps, proj_matricies, stride, valid_joints, use_soft_argmax=False, n_iters=64,
                        reprojection_error_epsilon=5, pair_seed=0, frame_offset=0, use_reprojection_xe=False, sigma=None,
                        direct_optimization=False, map_score=None, frame_keys=None):
    """Pool-level entry: heatmaps [N, V, J, H, W] (CUDA), proj_matricies [N, V, 3, 4], valid_joints [N, J] or [J].
    Returns a dict of CUDA tensors (keypoints_3d [N,J,3] f64, keypoints_2d, metric [N] f64, inlier_count [N] i32,
    reproj_mean [N,J], inliers [N,J]).  For C(V,2) > n_iters the view-pair subsets are the counter-based ones keyed
    by (pair_seed, frame key, joint), frame key = frame_keys[frame] (int64 [N], e.g. derived from the guid, so that the
    subsets do not depend on sharding) or frame_offset + frame -- see include/mval_b200.h.
    map_score "HP" / "MPE" / "BSB": also return "map_score" float32 [N, V, J], the per-map AL score of strategy.py:1149-1215;
    on the arg-max path it comes out of the same pass over the heat maps as the triangulation."""
    if use_soft_argmax or direct_optimization:
        # the refinement (utils/triangulation.py:319-336) needs the inlier masks, which only the unfused path keeps
        kp = ops.decode_softargmax(heatmaps, stride) if use_soft_argmax else ops.decode_argmax(heatmaps, stride, valid_joints)
        out = ops.triangulate_ransac(kp, proj_matricies, valid_joints, n_iters, float(reprojection_error_epsilon),
                                     pair_seed, frame_offset, direct_optimization=bool(direct_optimization),
                                     frame_keys=frame_keys)
        if map_score is not None:
            out["map_score"] = (ops.score_hp(heatmaps, valid_joints) if map_score == "HP"
                                else ops.score_peaks(heatmaps, map_score, valid_joints))
    else:
        out = ops.score_pool(heatmaps, proj_matricies, stride, valid_joints, n_iters, float(reprojection_error_epsilon),
                             pair_seed, frame_offset, map_score=map_score, frame_keys=frame_keys)
    if use_reprojection_xe:
        # utils/triangulation.py:223-224: metric = _compute_xe(...) replaces the mean reprojection error
        out["reproj_metric"] = out["metric"]
        out["metric"] = ops.score_xe(heatmaps, proj_matricies, out["keypoints_3d"], sigma)
    return out
