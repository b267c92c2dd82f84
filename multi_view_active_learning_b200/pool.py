"""Pool-level host logic: frame sharding across ranks and the cross-rank ranking merge.

Decode, scoring and triangulation are embarrassingly parallel: rank r owns the contiguous frame range
``shard_range(n, world, r)`` and no collective runs on the data path (the reference interleaves frames with a
DistributedSampler and issues 8 all_gathers per frame, strategy.py:753,1106-1114).  Ranking needs one exchange:
every rank contributes its local top-k ``(score, global index)`` and all ranks compute the same merge.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_frames, world_size, rank):
    """Contiguous shard [start, stop) of rank ``rank``; the first n % world ranks get one extra frame."""
    base, extra = divmod(int(n_frames), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def merge_topk(values, indices, k):
    """values / indices: 1-D arrays of candidates from all ranks (any order).  Returns the indices of the k best in
    the reference's order (strategy.py:945-949): score descending, ties by ascending pool index, NaN never selected."""
    values = np.asarray(values, dtype=np.float64)
    indices = np.asarray(indices, dtype=np.int64)
    keep = ~np.isnan(values) & (indices >= 0)
    values, indices = values[keep], indices[keep]
    order = np.lexsort((indices, -values))
    return indices[order[:k]], values[order[:k]]


def distributed_topk(local_topk, k, group=None):
    """``local_topk``: (idx int64 [m], val float64 [m]) tensors of this rank's best m <= k candidates with GLOBAL
    indices (e.g. from ops.topk_desc(scores, k, index_offset=shard_start)).  One all_gather of k*(8+8) bytes per
    rank; every rank returns the same (indices, values) numpy arrays."""
    idx, val = local_topk
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_topk(val.cpu().numpy(), idx.cpu().numpy(), k)
    world = dist.get_world_size(group)
    dev = idx.device
    pad_idx = torch.full((k,), -1, dtype=torch.int64, device=dev)
    pad_val = torch.full((k,), float("nan"), dtype=torch.float64, device=dev)
    pad_idx[: idx.numel()] = idx
    pad_val[: val.numel()] = val
    all_idx = [torch.empty_like(pad_idx) for _ in range(world)]
    all_val = [torch.empty_like(pad_val) for _ in range(world)]
    dist.all_gather(all_idx, pad_idx, group=group)
    dist.all_gather(all_val, pad_val, group=group)
    return merge_topk(torch.cat(all_val).cpu().numpy(), torch.cat(all_idx).cpu().numpy(), k)
