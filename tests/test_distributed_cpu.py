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
