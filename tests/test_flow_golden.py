"""The scoring-and-selection FLOW against the unmodified reference (tests/golden/flow_*.npz, produced by
oracle/make_golden_flow.py: the reference's own _sal_pseudo_labeling / _compute_sal_dict with its DataLoader +
DistributedSampler and per-frame all_gathers, under gloo at world size 1 and 2):

  * CPU: the oracle's restatement of the flow (oracle/flow_oracle.py) reproduces the fixtures -- dict insertion order,
    every value with the reference's float32 / float64 roundings, al_guids, sal_guids;
  * GPU (-m gpu): ScoringSelectionMixin reproduces them through the C ABI, single process and as two gloo ranks sharing
    cuda:0 with the mixin's own DistributedSampler loader.
"""
import json
import math
import os
import random
import socket

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import flow_oracle as FO
from oracle.make_golden_flow import CASES, FlowDataset, cluster_file_payload, flow_cfg

WORLDS = (1, 2)


def load(name, world):
    return dict(np.load(os.path.join(GOLDEN, "flow_%s_w%d.npz" % (name, world)), allow_pickle=False))


def fit_kmeans(root=2, n_clusters=3, seed=1307):
    """strategy.py:37-52 on the generator's cluster file."""
    from sklearn.cluster import KMeans

    kp_values = []
    for kp in cluster_file_payload().values():
        kp = np.array(kp)
        kp_values.append((kp[0:3, :] - kp[0:3, root:root + 1]).flatten())
    return KMeans(n_clusters, random_state=seed).fit(kp_values)


@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_flow_reproduces_the_reference(case, world):
    name, strategy, expr, clustered, config, n_al, n_pseudo = case
    gold = load(name, world)
    pool = json.loads(str(gold["pool"]))
    ds = FlowDataset(**pool)
    n = pool["n"]
    guids = list(ds.unlabeled_data)
    gt = np.stack([f["3d_keypoints"].numpy() for f in ds.unlabeled_data.values()])
    order = FO.sampler_order(n, world)
    assert len(order) == -(-n // world) * world and sorted(set(order)) == list(range(n))
    sal = FO.compute_sal_dict(ds.hm[:n], ds.pool["P"][:n], ds.pool["valid"][:n], gt, guids, strategy, config, order)
    assert list(sal["al_metric"]) == gold["guids"].tolist()  # dict insertion order incl. the sampler's repeated frames
    for key in ("al_metric", "sal_metric", "inlier_count"):
        got = np.array(list(sal[key].values()), dtype=np.float64)
        np.testing.assert_array_equal(got, gold[key], err_msg=key)  # bit for bit, NaN == NaN
    # mkpe ends in torch.mean over J float32 values (utils/evaluation.py:208), whose summation order is torch's (and differs
    # between its CPU and CUDA reductions): defined up to float32 rounding only; it feeds a TensorBoard histogram, no selection
    mk = np.array(list(sal["mkpe"].values()), dtype=np.float64)
    assert np.array_equal(np.isnan(mk), np.isnan(gold["mkpe"]))
    np.testing.assert_allclose(mk[~np.isnan(mk)], gold["mkpe"][~np.isnan(mk)], rtol=3e-7, atol=0)
    np.testing.assert_array_equal(np.array(list(sal["pred_3d_keypoints"].values())), gold["pred_3d_keypoints"])
    random.seed(99)
    al, sg = FO.sal_pseudo_labeling(sal, strategy, expr, ds.get_al_dict_for_coreset(), 2, n_al, n_pseudo, 2,
                                    kmeans=fit_kmeans() if clustered else None, n_clusters=3)
    assert al == gold["al_guids"].tolist()
    assert sg == gold["sal_guids"].tolist()


# ------------------------------------------------------------------------------------------------------------------ GPU
def _check_against_gold(gold, al_guids, sal_guids, sal, strategy, pseudo_seen):
    """Order exact; integer-like fields exact; float fields within the documented tolerances, with the share of values that
    equal the reference bit for bit reported in the assertion message."""
    assert list(sal["al_metric"].keys()) == gold["guids"].tolist()
    inl = np.array(list(sal["inlier_count"].values()))
    np.testing.assert_array_equal(inl, gold["inlier_count"])
    salm = np.array(list(sal["sal_metric"].values()))
    np.testing.assert_allclose(salm, gold["sal_metric"], rtol=0, atol=1e-4)  # north star: reprojection errors within 1e-4 px
    assert np.mean(salm == gold["sal_metric"]) >= 0.8, "float32 roundings of the metric: %s" % (salm == gold["sal_metric"])
    pred = np.array(list(sal["pred_3d_keypoints"].values()))
    np.testing.assert_allclose(pred, gold["pred_3d_keypoints"], rtol=1e-3, atol=1e-2)
    alm = np.array(list(sal["al_metric"].values()))
    tol = {"TRIANGULATION": 1e-4, "HP": 2e-6, "CORESET": 0.0}[strategy]
    np.testing.assert_allclose(alm, gold["al_metric"], rtol=0, atol=tol)
    mk = np.array(list(sal["mkpe"].values()))
    assert np.array_equal(np.isnan(mk), np.isnan(gold["mkpe"]))
    ok = ~np.isnan(mk)
    np.testing.assert_allclose(mk[ok], gold["mkpe"][ok], rtol=1e-5, atol=1e-4)
    if strategy == "HP":  # float32 softmax of two implementations: the selection may differ inside the tie band only
        ref = dict(zip(gold["guids"].tolist(), gold["al_metric"].tolist()))
        worst = min(ref[g] for g in gold["al_guids"].tolist())
        assert len(al_guids) == len(gold["al_guids"]) and all(ref[g] >= worst - 2 * tol for g in al_guids)
    else:
        assert al_guids == gold["al_guids"].tolist()
    assert sal_guids == gold["sal_guids"].tolist()
    if len(sal_guids):
        np.testing.assert_allclose(np.array(pseudo_seen), gold["pseudo_3d_keypoints"], rtol=1e-3, atol=1e-2)


def _run_ours(case, world, tmpdir, loader_factory):
    from multi_view_active_learning_b200.strategy import ActiveLearningStrategy

    name, strategy, expr, clustered, config, n_al, n_pseudo = case
    gold = load(name, world)
    pool = json.loads(str(gold["pool"]))
    cluster_path = ""
    if clustered:
        cluster_path = os.path.join(str(tmpdir), "clusters.json")
        with open(cluster_path, "w") as f:
            json.dump(cluster_file_payload(), f)
    cfg = flow_cfg(strategy, expr, world, cluster_path, config, pool["batch"])
    st = ActiveLearningStrategy(cfg)
    ds = FlowDataset(**pool)
    if loader_factory is not None:
        st._get_dataloader = loader_factory
    random.seed(99)
    _, al_guids, sal_guids, sal = st._sal_pseudo_labeling(ds, n_al, n_pseudo, torch.nn.Identity())
    pseudo = [d["pseudo_3d_keypoints"].tolist() for d in ds.pseudo_labeled_data]
    _check_against_gold(gold, al_guids, sal_guids, sal, strategy, pseudo)
    return True


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gpu_flow_single_process(case, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = json.loads(str(load(case[0], 1)["pool"]))["n"]
    order = FO.sampler_order(n, 1)  # what DataLoader(dataset, sampler=DistributedSampler(dataset)) yields at world size 1
    _run_ours(case, 1, tmp_path, lambda ds, bs, nw: torch.utils.data.DataLoader(ds, batch_size=bs, num_workers=0, sampler=order))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, tmpdir, out_q):
    import traceback

    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for case in CASES:  # the mixin's own loader: DataLoader + DistributedSampler (strategy.py:747-760)
            _run_ours(case, world, os.path.join(tmpdir, "r%d" % rank), None)
        out_q.put((rank, "ok"))
    except Exception:
        out_q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_gpu_flow_two_ranks(tmp_path):
    """Two processes (gloo for the exchange, both on cuda:0) run _sal_pseudo_labeling with the DistributedSampler loader;
    every rank must reproduce the reference's world-size-2 run: interleaved insertion order, the repeated frame of the
    sampler's padding, identical selections."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.multiprocessing as mp

    world = 2
    for r in range(world):
        os.makedirs(os.path.join(str(tmp_path), "r%d" % r), exist_ok=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert results[r] == "ok", results[r]
