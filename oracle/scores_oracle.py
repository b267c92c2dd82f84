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
    hm = np.asarray(heatmaps, dtype=np.float32)
    e = np.exp(hm - hm.max(axis=-1, keepdims=True))
    p = e / e.sum(axis=-1, keepdims=True, dtype=np.float32)
    return (np.float32(1.0) - p.max(axis=(-2, -1))).astype(np.float32)


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
