"""Real multi-process NCCL checks (need >= 2 visible GPUs; skipped otherwise): sharded pool scoring + ranking merge,
and the sharded coreset loop, against the single-device results."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from multi_view_active_learning_b200 import ops, pool as P
        from multi_view_active_learning_b200 import synthetic as S

        N, V, J, k = 600, 8, 19, 25
        pool = S.make_pool(N, V, J, seed=4, p_outlier=0.15)
        lo, hi = P.shard_range(N, world, rank)
        centres = torch.from_numpy(pool["centres"][lo:hi]).cuda()
        hm = ops.synth_heatmaps(centres, noise=0.0, seed=0)
        out = ops.score_pool(hm, torch.from_numpy(pool["P"][lo:hi]).cuda(), 4, frame_offset=lo)
        sel, val = P.distributed_topk(ops.topk_desc(out["metric"], k, index_offset=lo), k)
        # coreset over the predicted poses, sharded the same way
        feats = out["keypoints_3d"].reshape(hi - lo, -1).float()
        labeled = torch.from_numpy(np.random.default_rng(0).normal(size=(5, J * 3)).astype(np.float32) * 300).cuda()
        csel, _ = P.kcenter_greedy_sharded([(feats, lo)], labeled, 20)
        q.put((rank, sel.tolist(), val.tolist(), csel.cpu().tolist(), out["metric"].cpu().numpy(), feats.cpu().numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_scoring_ranking_and_coreset():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from multi_view_active_learning_b200 import ops, pool as P
    from oracle import coreset_oracle as CO
    from oracle import scores_oracle as SO

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    metric = np.concatenate([r[4] for r in res])
    feats = np.concatenate([r[5] for r in res])
    exp = SO.rank_nlargest({i: float(m) for i, m in enumerate(metric)}, 25)
    labeled = np.random.default_rng(0).normal(size=(5, feats.shape[1])).astype(np.float32) * 300
    exp_c = CO.kcenter_greedy_f32(np.concatenate([feats, labeled]), len(feats), 20)[0]
    for r in res:
        assert r[1] == exp and r[2] == [float(metric[i]) for i in exp]
        assert r[3] == exp_c


def _flow_worker(rank, world, port, tmpdir, q):
    """_sal_pseudo_labeling with the mixin's own DistributedSampler loader, one NCCL rank per GPU: the packed
    all_gather_into_tensor exchange, the dict-insertion dedupe and (CORESET) the sharded greedy rounds, against the
    reference's world-size-2 fixtures; plus the device-side ranking exchange (RankingExchange) against the host merge."""
    import traceback

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from test_flow_golden import CASES, _run_ours

        from multi_view_active_learning_b200 import ops, pool as P

        os.makedirs(os.path.join(tmpdir, "r%d" % rank), exist_ok=True)
        for case in CASES:
            _run_ours(case, world, os.path.join(tmpdir, "r%d" % rank), None)
        rng = np.random.default_rng(5)
        scores = rng.normal(size=4001).round(1)
        scores[rng.integers(0, 4001, 40)] = np.nan
        lo, hi = P.shard_range(len(scores), world, rank)
        ex = P.RankingExchange(64, torch.device("cuda", rank))
        idx, val, cnt = ex(torch.from_numpy(scores[lo:hi]).cuda(), lo)
        host_idx, host_val = P.distributed_topk(ops.topk_desc(torch.from_numpy(scores[lo:hi]).cuda(), 64, index_offset=lo), 64)
        assert int(cnt.item()) == 64 and idx.cpu().tolist() == host_idx.tolist() and val.cpu().tolist() == host_val.tolist()
        q.put((rank, "ok"))
        dist.barrier()
    except Exception:
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_flow_against_reference_fixtures(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flow_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for r in range(world):
        assert res[r] == "ok", res[r]
