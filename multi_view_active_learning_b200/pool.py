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


class RankingExchange:
    """Cross-rank top-k without the host: every rank's fixed-size local top-k (ops.topk_desc(..., fixed=True): global
    indices, unused slots -1 / NaN) is packed into ONE int64 buffer, exchanged with ONE all_gather_into_tensor and merged
    on the device by mval_topk_merge; the selection stays on the device until the caller reads it.  Buffers are allocated
    once per (k, device)."""

    def __init__(self, k, device, group=None):
        from . import ops  # noqa: F401

        self.k, self.group = int(k), group
        self.multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.multi else 1
        self.local = (torch.empty(self.k, dtype=torch.int64, device=device), torch.empty(self.k, dtype=torch.float64, device=device),
                      torch.empty(1, dtype=torch.int32, device=device))
        self.send = torch.empty(2 * self.k, dtype=torch.int64, device=device)
        self.recv = torch.empty(self.world * 2 * self.k, dtype=torch.int64, device=device)
        self.out = (torch.empty(self.k, dtype=torch.int64, device=device), torch.empty(self.k, dtype=torch.float64, device=device),
                    torch.empty(1, dtype=torch.int32, device=device))

    def __call__(self, scores, index_offset):
        """scores float64 CUDA [n_local] of this rank's contiguous shard starting at global index ``index_offset`` ->
        (idx int64 [k], val float64 [k], count int32 [1]) CUDA, identical on every rank."""
        from . import ops

        idx, val, cnt = ops.topk_desc(scores, self.k, index_offset=index_offset, fixed=True, out=self.local)
        if not self.multi:
            return idx, val, cnt
        self.send[: self.k].copy_(idx)
        self.send[self.k:].copy_(val.view(torch.int64))
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        r = self.recv.view(self.world, 2, self.k)
        return ops.topk_merge(r[:, 1].contiguous().view(torch.float64).reshape(-1), r[:, 0].contiguous().reshape(-1), self.k,
                              out=self.out)


# ----------------------------------------------------------------------------------------------------------------
# coreset k-center greedy over row-sharded features (utils/coreset.py:71-95 across ranks)
# ----------------------------------------------------------------------------------------------------------------
INIT_CHUNK = 256  # labeled centres folded in per pass over the features


def kcenter_fold_centres(state, centres, centre_norms=None, flags=0):
    """min_dist of every shard in ``state`` <- min over the given centre rows too (coreset.py:83-84)."""
    from . import ops

    if centres.shape[0] == 0:
        return
    centres = centres.float().contiguous()
    if centre_norms is None:
        centre_norms = ops.kcenter_norms(centres)
    for c0 in range(0, centres.shape[0], INIT_CHUNK):
        for s in state:
            ops.kcenter_update_batch(s["feat"], s["norms"], centres[c0:c0 + INIT_CHUNK], centre_norms[c0:c0 + INIT_CHUNK],
                                     s["min"], flags)


def kcenter_rounds(state, budget, group=None, k_slots=None, flags=0, stats=None):
    """``budget`` greedy picks over the shards in ``state`` (dicts with feat / norms / min / off), in rounds: every
    shard contributes its candidate record block, the blocks are exchanged with ONE all_gather per round (not per
    pick), every rank replays the greedy loop on the union -- identical picks everywhere -- and folds the new centres
    into its own shards.  Returns the selected global indices (int64 CUDA [budget])."""
    from . import ops

    state_list = state
    dev = state_list[0]["feat"].device
    d = state_list[0]["feat"].shape[1]
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    n_local = len(state)
    n_blocks = world * n_local
    if k_slots is None:
        # as many candidates per round as the replay CTA holds (1024 over all shards): fewer rounds, same picks
        k_slots = max(4, (1024 // n_blocks) // 4 * 4)
    assert k_slots % 4 == 0 and n_blocks * k_slots <= 1024, "n_blocks * k_slots must be <= 1024 and k_slots % 4 == 0"
    rb = ops.kcenter_records_bytes(k_slots, d)
    local = torch.empty(n_local * rb, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * n_local * rb, dtype=torch.uint8, device=dev) if multi else local
    resolver = ops.KcenterResolver(n_blocks, k_slots, d, dev)
    selected = torch.empty(max(int(budget), 1), dtype=torch.int64, device=dev)
    if int(budget) <= 0:
        return selected[:0]
    # The number of picks of a round is only known on the device (the replay decides it): resolve_async advances the
    # counters there and update_batch_dev reads the round's count there, so no round waits for the host.  The host learns
    # the running total one round late (pinned copy + event) and stops launching when it sees the budget reached; the round
    # launched in the meantime finds budget - done = 0 and does nothing.  Every rank sees the same totals (the replay is
    # replicated), so all ranks launch the same number of rounds -- and of all_gathers.
    state = torch.tensor([0, 0, int(budget), 0], dtype=torch.int32, device=dev)
    done_host = torch.zeros(2, dtype=torch.int32).pin_memory()
    events = [torch.cuda.Event(), torch.cuda.Event()]
    picks_log = []
    r = 0
    while True:
        for i, s in enumerate(state_list):
            ops.kcenter_select(s["feat"], s["norms"], s["min"], s["off"], k_slots, out=local[i * rb:(i + 1) * rb])
        if multi:
            dist.all_gather_into_tensor(gathered, local, group=group)
        resolver.resolve_async(gathered, selected, state)
        for s in state_list:
            ops.kcenter_update_batch_dev(s["feat"], s["norms"], resolver.centres, resolver.centre_norms, state[1:2], s["min"],
                                         flags | 4)  # 4: the picks of a round share one recheck
        if stats is not None:
            picks_log.append(state[1:2].clone())
        done_host[r & 1:(r & 1) + 1].copy_(state[0:1], non_blocking=True)
        events[r & 1].record()
        if r >= 1:
            events[(r - 1) & 1].synchronize()
            if int(done_host[(r - 1) & 1]) >= int(budget):
                break
        r += 1
        assert r <= 4 * (int(budget) + 8), "kcenter rounds make no progress"
    if stats is not None:
        stats.extend(t for t in torch.cat(picks_log).cpu().tolist() if t > 0)
    return selected[: int(budget)]


def pad_features(feat, multiple):
    """Appends zero columns up to a multiple of ``multiple``.  The canonical float32 distance (oracle/coreset_oracle.c) is
    unchanged by them: fma(0, 0, acc) = acc closes the dot-product chain and the squared norms gain exact zeros, so the
    selection and the running minima are bit-identical -- but a 16-byte aligned row of at least 64 floats is what the
    tensor-core screen (csrc/kcenter_tc.cu) needs, e.g. the reference's own d = 3 J = 57 -> 64."""
    d = feat.shape[1]
    pad = (-d) % int(multiple)
    if d + pad < 64:
        pad = 64 - d
    if pad == 0:
        return feat
    out = torch.zeros((feat.shape[0], d + pad), dtype=feat.dtype, device=feat.device)
    out[:, :d] = feat
    return out


def auto_pad(d):
    """Column multiple to zero-pad d-dimensional features to so that the batched update takes the tcgen05 screen
    (csrc/kcenter_tc.cu needs 16-byte aligned rows of at least 64 floats); 0 = leave as is.  Measured on one B200, 1M rows,
    1000 labeled, budget 10 000 (profiles/r2_summary.md): d = 57 -> 64: 61.9 -> 30.0 ms; d = 126 -> 128: 95.3 -> 22.7 ms, same
    picks -- the reference's own feature sizes (3 J = 57 Panoptic, 126 InterHand) were on the FFMA pass before."""
    d = int(d)
    if (d % 4 == 0 and d >= 64) or d <= 32:  # already eligible, or so short that the FFMA pass over d columns is cheaper
        return 0
    return 32


def kcenter_greedy_sharded(shards, labeled, budget, group=None, k_slots=None, flags=0, stats=None, pad_to="auto"):
    """Greedy k-center selection over UNLABELED feature rows that are row-sharded contiguously.

    shards : list of (features float32 CUDA [n_s, d], global_row_offset) owned by THIS process -- one entry per
             rank in a torch.distributed job, several entries to emulate ranks on one device (tests).
    labeled: float32 CUDA [L, d], the labeled centres, replicated on every rank (coreset.py:83-84).
    "First index wins" across contiguous shards is "lowest global index among equal maxima", i.e. np.argmax
    (coreset.py:90).  Returns (selected global indices int64 CUDA [budget], list of per-shard min_dist tensors).
    Labeled rows never compete: their min-distance is exactly 0 and an all-zero pool resolves to global index 0 in
    the reference too.
    """
    from . import ops

    assert len(shards) >= 1 and labeled.shape[0] >= 1, "need at least one shard and one labeled centre"
    dev = labeled.device
    if pad_to == "auto":
        pad_to = auto_pad(labeled.shape[1])
    if pad_to:
        labeled = pad_features(labeled.float(), pad_to)
    state = []
    for feat, off in shards:
        feat = feat.float().contiguous()
        if pad_to:
            feat = pad_features(feat, pad_to)
        n = feat.shape[0]
        norms = ops.kcenter_norms(feat) if n else torch.empty(0, dtype=torch.float32, device=dev)
        state.append({"feat": feat, "off": int(off), "norms": norms,
                      "min": torch.full((n,), float("inf"), dtype=torch.float32, device=dev)})
    kcenter_fold_centres(state, labeled, flags=flags)
    selected = kcenter_rounds(state, int(budget), group=group, k_slots=k_slots, flags=flags, stats=stats)
    return selected, [s["min"] for s in state]
