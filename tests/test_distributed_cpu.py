"""World-size-2 gloo tests (CPU) of the multi-rank host logic: frame sharding, the ranking merge after the
all_gather, and the per-field gather that restores the reference's dict insertion order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_view_active_learning_b200 import pool as P
from oracle import scores_oracle as SO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, scores, k, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = len(scores)
        lo, hi = P.shard_range(n, world, rank)
        local = scores[lo:hi]
        # the local top-k a rank would get from mval_topk_desc (descending, ties by index, NaN dropped)
        order = [i for i in np.lexsort((np.arange(hi - lo), -local)) if not np.isnan(local[i])][:k]
        idx = torch.tensor([lo + i for i in order], dtype=torch.int64)
        val = torch.tensor([local[i] for i in order], dtype=torch.float64)
        sel_idx, sel_val = P.distributed_topk((idx, val), k)
        # per-field gather in the reference's interleaved order
        from multi_view_active_learning_b200.strategy import ScoringSelectionMixin

        t = torch.arange(3, dtype=torch.float32) * world + rank  # local position t of rank r -> value t*world + r
        g = ScoringSelectionMixin._gather_interleaved({"x": t, "y": torch.stack([t, t], dim=1)})
        out_q.put((rank, sel_idx.tolist(), sel_val.tolist(), g["x"].tolist(), g["y"][:, 1].tolist()))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_the_pool():
    for n in (0, 1, 7, 100, 100001):
        for w in (1, 2, 3, 8):
            r = [P.shard_range(n, w, i) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_merge_topk_order():
    vals = np.array([1.0, 3.0, np.nan, 3.0, 2.0, 3.0])
    idx = np.array([10, 7, 1, 3, 4, -1])
    i, v = P.merge_topk(vals, idx, 3)
    assert i.tolist() == [3, 7, 4] and v.tolist() == [3.0, 3.0, 2.0]


def test_distributed_topk_and_gather_world2():
    rng = np.random.default_rng(1)
    n, k, world = 101, 9, 2
    scores = rng.normal(size=n).round(1)  # ties
    scores[[3, 50, 77]] = np.nan
    exp = SO.rank_nlargest({i: float(s) for i, s in enumerate(scores)}, k)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, scores, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, sel_idx, sel_val, gx, gy in results:
        assert sel_idx == exp  # every rank ends with the same selection, equal to the single-process nlargest
        assert sel_val == [float(scores[i]) for i in exp]
        assert gx == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0] and gy == gx


def _sal_worker(rank, world, port, n_frames, out_q, mode="even"):
    """One rank of _compute_sal_dict with the device calls stubbed (CPU tensors, gloo): frames are dealt round-robin like
    the reference's DistributedSampler (strategy.py:753), every rank must end with the same dicts in the order the
    reference's per-frame all_gathers insert them (frame 0, 1, 2, ...: local position t of rank r = global t * world + r)."""
    from types import SimpleNamespace as NS

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multi_view_active_learning_b200 import ops, strategy as ST

        V, J, B = 2, 4, 3
        torch.Tensor.cuda = lambda self, *a, **k: self
        ops.mkpe = lambda p, g, v: p[:, 0, 0].float()
        from conftest import first_occurrence_numpy

        ops.first_occurrence = first_occurrence_numpy

        def fake_triangulation_batch(hm, P, stride, joint_valid, **kw):
            ids = hm[:, 0, 0, 0, 0].double()  # the frame id travels in the first heat-map pixel
            scores = ids[:, None, None].float() + torch.arange(V * J, dtype=torch.float32).reshape(1, V, J) / 8
            return {"metric": ids + 0.25, "inlier_count": torch.full((len(ids),), 8, dtype=torch.int32),
                    "keypoints_3d": ids[:, None, None].expand(-1, J, 3).clone(), "map_score": scores}

        ST.triangulation.triangulation_batch = fake_triangulation_batch
        mine = list(range(rank, n_frames, world))
        if mode == "sampler_padding" and len(mine) < -(-n_frames // world):
            mine.append(rank - (n_frames % world))  # DistributedSampler repeats the head of the index list (strategy.py:753)

        def loader():
            for s in range(0, len(mine), B):
                ids = torch.tensor(mine[s:s + B])
                hm = torch.zeros(len(ids), V, J, 4, 4)
                hm[:, 0, 0, 0, 0] = ids.float()
                yield {"images": hm, "proj_matrices": torch.zeros(len(ids), V, 3, 4, dtype=torch.float64),
                       "joint_valid": torch.ones(len(ids), J), "3d_keypoints": torch.zeros(len(ids), 4, J),
                       "pose": torch.full((len(ids),), 7), "frame_id": ids}

        cfg = NS(EXPR_TYPE="AL", RANDOM_SEED=1, DATA=NS(NUM_JOINTS=J, TYPE="panoptic"), POSE_ESTIMATOR=NS(STRIDE=4),
                 SAL=NS(INLIER_THRESHOLD=4, CLUSTER_FILE_PATH="", NUM_CLUSTERS=2),
                 AL=NS(STRATEGY="HP", USE_SOFTARGMAX=False, USE_REPROJECTION_XE=False, REPROJECTION_SIGMA=1.0, HP_CONFIG="AVG",
                       MPE_CONFIG="AVG", BSB_CONFIG="AVG"))
        st = ST.ActiveLearningStrategy(cfg)
        st._compute_batch_heatmap = lambda pe, d: d["images"].reshape(-1, J, 4, 4)
        sal = st._compute_sal_dict(loader(), None)
        out_q.put((rank, list(sal["al_metric"].items()), list(sal["sal_metric"].items()), list(sal["mkpe"].values())))
    finally:
        dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("n_frames,mode", [(12, "even"), (11, "uneven"), (11, "sampler_padding")])
def test_compute_sal_dict_world2_order_and_values(n_frames, mode):
    """even: 6 frames per rank, batches of 3.  uneven: a loader that gives rank 1 one frame less (rows are padded for the
    exchange and dropped again).  sampler_padding: the DistributedSampler's repeat of frame 0 on rank 1 -- the guid keeps
    its first position and appears once, exactly like the reference's dict insertion."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sal_worker, args=(r, world, port, n_frames, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mean_offset = sum(i / 8 for i in range(8)) / 8  # HP AVG of the stub's per-map scores = frame id + mean(0 .. 7) / 8
    for rank, al, sal, mkpe in results:
        assert [g for g, _ in al] == ["7-%d" % i for i in range(n_frames)] == [g for g, _ in sal]
        assert [v for _, v in sal] == [i + 0.25 for i in range(n_frames)]
        assert [v for _, v in al] == [float(np.float32(i + mean_offset)) for i in range(n_frames)]
        assert mkpe == [float(i) for i in range(n_frames)]
