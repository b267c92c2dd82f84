#!/usr/bin/env python
"""Throughput of ScoringSelectionMixin._compute_sal_dict (reference strategy.py:1004-1147) on an in-memory loader:
heat maps stand in for the images and the pose estimator is the identity, so this times the scoring pipeline around the
kernels (host->device staging, per-batch launches, the dict build), not the backbone.
    python tools/bench_sal_dict.py [frames] [batch] [strategy]"""
import os
import sys
import time
from types import SimpleNamespace as NS

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_view_active_learning_b200 import synthetic as S  # noqa: E402
from multi_view_active_learning_b200.strategy import ActiveLearningStrategy  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    strategy = sys.argv[3] if len(sys.argv) > 3 else "TRIANGULATION"
    V, J = 8, 19
    cfg = NS(EXPR_TYPE="AL", RANDOM_SEED=1307, DATA=NS(NUM_JOINTS=J, TYPE="panoptic"), POSE_ESTIMATOR=NS(STRIDE=4),
             SAL=NS(INLIER_THRESHOLD=4, CLUSTER_FILE_PATH="", NUM_CLUSTERS=10),
             AL=NS(STRATEGY=strategy, USE_SOFTARGMAX=False, USE_REPROJECTION_XE=False, REPROJECTION_SIGMA=1.0, HP_CONFIG="AVG",
                   MPE_CONFIG="AVG", BSB_CONFIG="AVG", INFERENCE=NS(BATCH_SIZE=bs, NUM_WORKERS=0)))
    pool = S.make_pool(n, V, J, seed=1, p_outlier=0.1)
    hm = S.render_heatmaps(pool["centres"][:256], noise=0.05, seed=2)
    hm = np.tile(hm, (n // 256 + 1, 1, 1, 1, 1))[:n]
    gt = np.concatenate([pool["X"].transpose(0, 2, 1), np.ones((n, 1, J))], axis=1).astype(np.float32)
    frames = [{"images": torch.from_numpy(hm[i]), "proj_matrices": torch.from_numpy(pool["P"][i]),
               "joint_valid": torch.ones(J), "3d_keypoints": torch.from_numpy(gt[i]), "pose": 160422, "frame_id": i}
              for i in range(n)]
    st = ActiveLearningStrategy(cfg)
    make = getattr(st, "_get_dataloader")
    try:
        loader = lambda: torch.utils.data.DataLoader(frames, batch_size=bs, num_workers=0, pin_memory=True)  # noqa: E731
        loader()
    except Exception:
        loader = lambda: make(frames, bs, 0)  # noqa: E731
    if len(sys.argv) > 4 and sys.argv[4] == "resident":
        # batches already collated and on the device: what is left is the per-batch cost of the scoring pipeline itself
        batches = [{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()} for b in loader()]
        loader = lambda: batches  # noqa: E731
    if len(sys.argv) > 4 and sys.argv[4] == "pinned":
        # batches already collated in pinned host memory: what is left is the upload (copy stream, one batch ahead) under the
        # scoring launches -- the host side of a loader with enough workers
        batches = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()} for b in loader()]
        loader = lambda: batches  # noqa: E731
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sal = st._compute_sal_dict(loader(), torch.nn.Identity())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("%s: %d frames, batch %d: %.3f s = %.0f frames/s (metric of frame 0: %.6f)" % (
            strategy, n, bs, dt, n / dt, list(sal["al_metric"].values())[0]))


if __name__ == "__main__":
    main()
