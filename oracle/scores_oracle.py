"""CPU oracle for the heatmap uncertainty scores and the ranking.  TEST INFRASTRUCTURE, NOT PRODUCT.

  hp_scores / hp_metric  <- strategy.py:1178-1193 (_compute_hp): 1 - max(softmax(hm)) where the softmax has no
                            ``dim`` and therefore runs over dim=1 of the 2-D map, i.e. PER ROW
  rank_nlargest          <- strategy.py:932-949: drop NaN, heapq.nlargest(n, dict, key=dict.get)
  mkpe                   <- utils/evaluation.py:198-208 (compute_mkpe) for one frame
"""
import math
from heapq import nlargest

import numpy as np


def hp_scores(heatmaps):
    """[..., H, W] float32 -> float32 [...]: 1 - max over the map of the row-wise softmax (float32 arithmetic)."""
    import torch

    # evaluated with torch itself, as the reference does (F.softmax of the 2-D map = softmax over dim 1, then 1 - max):
    # numpy's exp and torch's vectorised exp differ in the last bit, and HP's STD is a float64 of these values
    hm = torch.from_numpy(np.ascontiguousarray(np.asarray(heatmaps, dtype=np.float32)))
    sm = torch.softmax(hm, dim=-1)
    return (1 - sm.amax(dim=(-2, -1))).numpy().astype(np.float32)


def hp_metric(heatmaps, joint_valid, config="AVG"):
    """One frame: heatmaps [V, J, H, W], joint_valid [J].  AVG = Python sum/len of the float32 items (float64
    accumulate), STD = np.std (population) -- reference :1188-1193.  Returns a Python float / np.float64."""
    s = hp_scores(heatmaps)
    vals = [float(s[v, k]) for v in range(s.shape[0]) for k in range(s.shape[1]) if joint_valid[k]]
    if config == "AVG":
        return sum(vals) / len(vals)
    if config == "STD":
        return np.std(np.array(vals))
    raise NotImplementedError


def rank_nlargest(metric_by_guid, n):
    """dict guid -> float.  NaN entries dropped first (:932-936); then the reference's exact call (:945-949)."""
    kept = {g: m for g, m in metric_by_guid.items() if not math.isnan(m)}
    return nlargest(n, kept, key=kept.get)


def mkpe(pred_3d, gt_3d, valid):
    """pred [J,3], gt [>=3, J], valid [J] (0/1) -> float32 scalar like compute_mkpe([pred],[gt],[valid]):
    per joint sqrt(sum_c (pred-gt)^2) where valid else 0, divided by valid (0/0 -> nan), then mean over J."""
    pred = np.asarray(pred_3d, dtype=np.float32)
    gt = np.asarray(gt_3d, dtype=np.float32)[:3]
    v = np.asarray(valid, dtype=np.float32)
    d = np.square(pred.T - gt)
    d = np.where(v.astype(bool), d, np.float32(0))
    d = np.sqrt(d.sum(axis=0, dtype=np.float32))
    with np.errstate(invalid="ignore", divide="ignore"):
        per_joint = d / v
    return np.float32(per_joint.mean(dtype=np.float32))


# ----------------------------------------------------------------------------------------------------------------
# MPE / BSB (strategy.py:1149-1176, 1195-1215).  parity unpinned: the reference calls
# skimage.feature.peak_local_max(..., indices=True) (skimage <= 0.19, not installed anywhere we can run); restated
# here from its published algorithm with scipy.ndimage, which IS what skimage uses underneath.
# ----------------------------------------------------------------------------------------------------------------
def peak_local_max(image, min_distance=2, num_peaks=np.inf):
    """skimage 0.18/0.19 peak_local_max(image, min_distance, indices=True, num_peaks) with the defaults the
    reference relies on (threshold_abs=None -> image.min(), threshold_rel=None, exclude_border=True -> min_distance,
    footprint = ones(2*min_distance+1), p_norm=inf): candidates = pixels equal to the maximum of their footprint
    (scipy maximum_filter, mode='constant'), strictly above the threshold, outside the border; sorted by descending
    intensity; greedy spacing (a peak removes later peaks at Chebyshev distance < min_distance); first num_peaks."""
    from scipy import ndimage as ndi

    image = np.asarray(image)
    size = 2 * min_distance + 1
    image_max = ndi.maximum_filter(image, footprint=np.ones((size, size), dtype=bool), mode="constant")
    mask = image == image_max
    if np.all(mask):
        mask[:] = False
    mask &= image > image.min()
    bw = min_distance
    mask[:bw, :] = mask[-bw:, :] = False
    mask[:, :bw] = mask[:, -bw:] = False
    coord = np.nonzero(mask)
    intensities = image[coord]
    order = np.argsort(-intensities, kind="stable")
    coord = np.transpose(coord)[order]
    keep, rejected = [], set()
    max_out = int(num_peaks) if np.isfinite(num_peaks) else None
    for i in range(len(coord)):
        if i in rejected:
            continue
        keep.append(i)
        if max_out is not None and len(keep) >= max_out:
            break
        d = np.abs(coord[i + 1:] - coord[i]).max(axis=1) if i + 1 < len(coord) else np.zeros(0)
        rejected.update((i + 1 + np.nonzero(d < min_distance)[0]).tolist())
    return coord[keep]


def mpe_scores(heatmaps):
    """[..., H, W] float32 -> float32 [...]: entropy of softmax over the local-peak values, with the reference's own
    arithmetic (strategy.py:1168-1175): peaks in peak_local_max order (descending intensity), np.exp of the float32
    values without a shift, Python sum() over float32 elements (sequential float32 adds), math.log in double, the
    products and their sum in float32 again (NumPy >= 2 promotion: np.float32 * Python float stays float32).  A map
    without peaks scores 0 (sum over an empty list).  Pinned against the reference in tests/golden/peak_scores.npz."""
    hm = np.asarray(heatmaps, dtype=np.float32)
    out = np.zeros(hm.shape[:-2], dtype=np.float32)
    for idx in np.ndindex(*hm.shape[:-2]):
        pk = peak_local_max(hm[idx], min_distance=2)
        peaks = [hm[idx][c[0]][c[1]] for c in pk]
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            e = np.exp(np.asarray(peaks, dtype=np.float32))
            probs = e / sum(e)
        out[idx] = sum(-p * math.log(p) for p in probs)
    return out


def bsb_scores(heatmaps):
    """[..., H, W] float32 -> float32 [...]: |p0 - p1| of the two highest local peaks of the ROW-softmaxed map
    (strategy.py:1202-1208; F.softmax without dim on a 2-D tensor = dim 1, evaluated with torch like the reference);
    NaN where the map has fewer than two peaks (the reference raises IndexError there)."""
    import torch

    hm = np.ascontiguousarray(np.asarray(heatmaps, dtype=np.float32))
    sm = torch.softmax(torch.from_numpy(hm), dim=-1).numpy()
    out = np.full(hm.shape[:-2], np.nan, dtype=np.float32)
    for idx in np.ndindex(*hm.shape[:-2]):
        pk = peak_local_max(sm[idx], min_distance=2, num_peaks=2)
        if len(pk) >= 2:
            out[idx] = abs(sm[idx][pk[0, 0], pk[0, 1]] - sm[idx][pk[1, 0], pk[1, 1]])
    return out


def reduce_frame_score(per_map, joint_valid, config="AVG", kind="HP"):
    """per_map [V, J] -> the frame score over (view, valid joint), view-major, as the reference forms it:
      HP       the per-map values are Python floats (.item()): AVG = sum(x) / len(x) in double, STD = np.std of a float64
               array (:1188-1193);
      MPE/BSB  the per-map values are np.float32 scalars: AVG = Python sum() = sequential float32 adds, / len in float32;
               STD = np.std of a float32 array, a float32 (:1151-1158, :1210-1215; NumPy >= 2 promotion rules)."""
    if kind == "HP":
        vals = [float(per_map[v, k]) for v in range(per_map.shape[0]) for k in range(per_map.shape[1]) if joint_valid[k]]
    else:
        vals = [np.float32(per_map[v, k]) for v in range(per_map.shape[0]) for k in range(per_map.shape[1]) if joint_valid[k]]
    if config == "AVG":
        return sum(vals) / len(vals)
    if config == "STD":
        return np.std(np.array(vals))
    raise NotImplementedError
