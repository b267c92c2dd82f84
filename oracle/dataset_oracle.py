"""CPU restatement of the data format on the far side of the pose estimator.  TEST INFRASTRUCTURE, NOT PRODUCT.

  gt_heatmaps  <- dataset/dataset.py:198-207: ground-truth heat maps of the joints of one view, evaluated with torch exactly as
                  the reference does (float32 pixel grid minus float64 labels -> float64 from there on).
"""
import numpy as np
import torch


def gt_heatmaps(points_2d, stride, width, height, sigma):
    """points_2d float64 [J, 2] image pixels (x, y) -> float64 [J, height // stride, width // stride]."""
    pt = torch.from_numpy(np.asarray(points_2d, dtype=np.float64)) / stride
    w, h = width // stride, height // stride
    grid = torch.zeros(size=(h, w, 2))
    grid[..., 0] = torch.from_numpy(np.arange(w)).unsqueeze(0)
    grid[..., 1] = torch.from_numpy(np.arange(h)).unsqueeze(1)
    grid = grid.unsqueeze(0)
    labels = pt.unsqueeze(-2).unsqueeze(-2)
    exponent = torch.sum((grid - labels) ** 2, dim=-1)
    return torch.exp(-exponent / (2.0 * (sigma ** 2))).numpy()
