"""Heat-map decode helpers with the reference's signatures (utils/evaluation.py:13-58)."""
import numpy as np
import torch

from .. import ops


def _device_heatmaps(pred_map):
    if not torch.is_tensor(pred_map):
        pred_map = torch.as_tensor(np.asarray(pred_map))
    if not pred_map.is_cuda:
        pred_map = pred_map.cuda()  # a copy, not a fallback: the arithmetic always runs on the GPU
    return pred_map.float()


def get_scaled_pred_corrdinates(pred_map, stride, num_keypoints, valid_joints):
    """Reference utils/evaluation.py:13-30.  pred_map [B, K, H, W]; returns np.int64 [B, num_keypoints, 2] with
    (x, y) = (argmax % H, argmax // H) * stride and [0, 0] for invalid joints."""
    hm = _device_heatmaps(pred_map)[:, :num_keypoints]
    valid = torch.as_tensor(np.asarray([bool(valid_joints[k]) for k in range(num_keypoints)]))
    xy = ops.decode_argmax(hm.unsqueeze(0), stride, valid)
    return xy[0].cpu().numpy().astype(np.int64)


def get_pred_coordinates(pred_map, bbox, num_keypoints, use_softargmax=False):
    """Reference utils/evaluation.py:33-58.  Arg-max branch: list [B][K][2] of 0-d tensors scaled by the bbox
    extent over the map size; soft-arg-max branch: tensor [B, K, 2] scaled by (bbox[3]-bbox[1]) / W (square boxes,
    as in the reference)."""
    hm = _device_heatmaps(pred_map)
    B, K, H, W = hm.shape
    bbox = torch.as_tensor(bbox).float().cpu()
    if use_softargmax:
        coords = ops.decode_softargmax(hm.unsqueeze(0), 1.0)[0]
        scale = ((bbox[:, 3] - bbox[:, 1]) / (1.0 * W)).to(coords.device)
        return coords * scale[:, None, None]
    flat = ops.decode_argmax(hm[:, :num_keypoints].unsqueeze(0), 1)[0].cpu()  # (x, y) = (c % H, c // H)
    out = []
    for b in range(B):
        sy = (bbox[b][2] - bbox[b][0]) / (1.0 * H)
        sx = (bbox[b][3] - bbox[b][1]) / (1.0 * W)
        out.append([[flat[b, k, 0] * sx, flat[b, k, 1] * sy] for k in range(num_keypoints)])
    return out
