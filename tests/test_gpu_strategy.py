"""Drop-in check of the strategy hot path: ``_compute_sal_dict`` / ``_sal_pseudo_labeling`` /
``sample_next_batch`` against an emulation of the reference's per-frame loop driven by the CPU oracle."""
import math
from collections import OrderedDict
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from multi_view_active_learning_b200 import synthetic as S
from oracle import coreset_oracle as CO
from oracle import scores_oracle as SO
from oracle import triangulation_oracle as O

pytestmark = pytest.mark.gpu


def make_cfg(strategy="TRIANGULATION", expr="AL", hp="AVG"):
    return NS(EXPR_TYPE=expr, RANDOM_SEED=1307, DATA=NS(NUM_JOINTS=19, TYPE="panoptic"),
              POSE_ESTIMATOR=NS(STRIDE=4),
              SAL=NS(INLIER_THRESHOLD=4, CLUSTER_FILE_PATH="", NUM_CLUSTERS=10),
              AL=NS(STRATEGY=strategy, USE_SOFTARGMAX=False, USE_REPROJECTION_XE=False, REPROJECTION_SIGMA=1.0,
                    HP_CONFIG=hp, MPE_CONFIG=hp, BSB_CONFIG=hp, INFERENCE=NS(BATCH_SIZE=3, NUM_WORKERS=0)))


class FakeDataset(torch.utils.data.Dataset):
    """Same surface as dataset/dataset.py's ActiveLearningDataset as far as the strategy touches it."""

    def __init__(self, n, V=8, J=19, seed=0, n_labeled=6):
        pool = S.make_pool(n + n_labeled, V, J, seed=seed, valid_prob=0.95, p_outlier=0.12)
        hm = S.render_heatmaps(pool["centres"], noise=0.05, seed=seed + 1)
        gt = np.concatenate([pool["X"].transpose(0, 2, 1), np.ones((n + n_labeled, 1, J))], axis=1)  # [4, J]
        frames = [{"images": torch.from_numpy(hm[i]), "proj_matrices": torch.from_numpy(pool["P"][i]),
                   "joint_valid": torch.from_numpy(pool["valid"][i].astype(np.float32)),
                   "3d_keypoints": torch.from_numpy(gt[i].astype(np.float32)), "pose": 160422 + i % 3, "frame_id": 100 + i}
                  for i in range(n + n_labeled)]
        self.unlabeled_data = OrderedDict(("%d-%d" % (f["pose"], f["frame_id"]), f) for f in frames[:n])
        self.labeled_data = [dict(f, **{"3d_keypoints": f["3d_keypoints"].numpy()}) for f in frames[n:]]
        self.pseudo_label_guids, self.pseudo_labeled_data, self.data = [], [], []
        self.hm, self.pool = hm, pool

    def resample_unlabeled_data(self):
        self.data = list(self.unlabeled_data.values())

    def get_al_dict_for_coreset(self):
        return {i: np.array(self.labeled_data[i]["3d_keypoints"]).transpose([1, 0]) for i in range(len(self.labeled_data))}

    def label_by_frame_guids(self, guids):
        for g in guids:
            self.labeled_data.append(self.unlabeled_data[g])
            del self.unlabeled_data[g]

    def pseudo_label_by_frame_guids(self, guids, pseudo_labels):
        self.pseudo_label_guids = guids

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        return self.data[i]


def make_strategy(cfg):
    from multi_view_active_learning_b200.strategy import ActiveLearningStrategy

    st = ActiveLearningStrategy(cfg)
    # single process: plain loader in dataset order (the reference's own test swaps the sampler out the same way,
    # tests/test_strategy.py:41-43)
    st._get_dataloader = lambda ds, bs, nw: torch.utils.data.DataLoader(ds, batch_size=bs, num_workers=0)
    return st


def reference_emulation(ds, cfg):
    """What strategy.py:1004-1147 leaves in sal_dict, computed from the oracle frame by frame."""
    n = len(ds.unlabeled_data)
    ref = O.triangulate_pool(ds.hm[:n], ds.pool["P"][:n], 4, ds.pool["valid"][:n])
    sal = {k: OrderedDict() for k in ("al_metric", "sal_metric", "inlier_count", "pred_3d_keypoints", "mkpe")}
    for i, (guid, f) in enumerate(ds.unlabeled_data.items()):
        sal["sal_metric"][guid] = float(np.float32(ref["metric"][i]))
        sal["inlier_count"][guid] = float(ref["inlier_count"][i])
        pred32 = ref["keypoints_3d"][i].astype(np.float32)
        sal["pred_3d_keypoints"][guid] = pred32.tolist()
        if cfg.AL.STRATEGY == "TRIANGULATION":
            sal["al_metric"][guid] = float(ref["metric"][i])
        elif cfg.AL.STRATEGY == "HP":
            m = SO.hp_metric(ds.hm[i], ds.pool["valid"][i], cfg.AL.HP_CONFIG)
            sal["al_metric"][guid] = float(np.float32(m)) if cfg.AL.HP_CONFIG == "AVG" else float(m)
        elif cfg.AL.STRATEGY in ("MPE", "BSB"):
            per_map = (SO.mpe_scores if cfg.AL.STRATEGY == "MPE" else SO.bsb_scores)(ds.hm[i])
            config = cfg.AL.MPE_CONFIG if cfg.AL.STRATEGY == "MPE" else cfg.AL.BSB_CONFIG
            m = SO.reduce_frame_score(per_map, ds.pool["valid"][i], config, cfg.AL.STRATEGY)
            sal["al_metric"][guid] = float(np.float32(m)) if config == "AVG" else float(m)
        else:
            sal["al_metric"][guid] = 0.0
        sal["mkpe"][guid] = float(SO.mkpe(pred32, f["3d_keypoints"].numpy(), ds.pool["valid"][i]))
    return sal


@pytest.mark.parametrize("strategy,hp", [("TRIANGULATION", "AVG"), ("HP", "AVG"), ("HP", "STD"), ("CORESET", "AVG"),
                                         ("MPE", "AVG"), ("BSB", "AVG"), ("MPE", "STD")])
def test_compute_sal_dict_and_selection(strategy, hp):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = make_cfg(strategy, hp=hp)
    ds = FakeDataset(40, seed=5)
    st = make_strategy(cfg)
    ref = reference_emulation(ds, cfg)
    ds.resample_unlabeled_data()
    sal = st._compute_sal_dict(st._get_dataloader(ds, 3, 0), torch.nn.Identity())
    assert list(sal["al_metric"].keys()) == list(ref["al_metric"].keys())
    for g in ref["sal_metric"]:
        assert sal["inlier_count"][g] == ref["inlier_count"][g]
        assert abs(sal["sal_metric"][g] - ref["sal_metric"][g]) <= 1e-4
        np.testing.assert_allclose(sal["pred_3d_keypoints"][g], ref["pred_3d_keypoints"][g], rtol=1e-3, atol=1e-2)
        tol = {"TRIANGULATION": 1e-4, "MPE": 2e-5}.get(strategy, 2e-6)
        assert abs(sal["al_metric"][g] - ref["al_metric"][g]) <= tol
        a, b = sal["mkpe"][g], ref["mkpe"][g]
        assert (math.isnan(a) and math.isnan(b)) or abs(a - b) <= 1e-2 * max(1.0, abs(b))
    # selection through the reference's entry point
    n_sel = 7
    ds2 = FakeDataset(40, seed=5)
    st.sample_next_batch(ds2, n_sel, 0, torch.nn.Identity(), iteration=1, rank=0)
    if strategy == "CORESET":
        keys = list(ref["pred_3d_keypoints"])
        F = CO.stacked_features(ref["pred_3d_keypoints"].values(), FakeDataset(40, seed=5).get_al_dict_for_coreset().values(), 2)
        exp = [keys[i] for i in CO.kcenter_greedy_f32(F, len(keys), n_sel)[0]]
        assert exp == [keys[i] for i in CO.kcenter_greedy_f64(F, len(keys), n_sel)[0]]
    else:
        exp = SO.rank_nlargest(ref["al_metric"], n_sel)
    if strategy in ("MPE", "BSB", "HP"):
        # float32 scores from two float32 softmax implementations: compare the selection up to near-ties
        got = st.last_al_guids
        assert len(got) == n_sel
        worst = min(ref["al_metric"][g] for g in exp)
        assert all(ref["al_metric"][g] >= worst - 2 * tol for g in got)
    else:
        assert st.last_al_guids == exp
    assert len(ds2.unlabeled_data) == 40 - n_sel and len(ds2.labeled_data) == 6 + n_sel


def test_sal_pseudo_label_filter_and_iteration0():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import random

    cfg = make_cfg("TRIANGULATION", expr="SAL")
    st = make_strategy(cfg)
    ds = FakeDataset(30, seed=9)
    ref = reference_emulation(ds, cfg)
    random.seed(3)
    _, al_guids, sal_guids, sal_dict = st._sal_pseudo_labeling(ds, 5, 4, torch.nn.Identity())
    exp_al = SO.rank_nlargest(ref["al_metric"], 5)
    assert al_guids == exp_al
    cand = {g: m for g, m in ref["sal_metric"].items()
            if g not in exp_al and not math.isnan(m) and ref["inlier_count"][g] > cfg.SAL.INLIER_THRESHOLD}
    best = sorted(cand, key=cand.get)[:8]
    assert len(sal_guids) == 4 and set(sal_guids) <= set(best)
    assert ds.pseudo_label_guids == sal_guids
    # iteration 0 = seeded random sampling, identical to the reference's random.sample (strategy.py:868-878)
    ds0 = FakeDataset(30, seed=9)
    keys = list(ds0.unlabeled_data.keys())
    st.sample_next_batch(ds0, 6, 0, None, iteration=0)
    random.seed(cfg.RANDOM_SEED)
    assert st.last_al_guids == random.sample(keys, 6)


def test_sal_rank_mkpe_and_pose_features_kernels():
    """mval_sal_rank / mval_mkpe / mval_pose_features against literal restatements of strategy.py:957-975,
    utils/evaluation.py:198-208 and utils/coreset.py:35-47."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from multi_view_active_learning_b200 import ops

    rng = np.random.default_rng(17)
    n = 5000
    metric = rng.uniform(0, 3, size=n).astype(np.float32)
    metric[rng.integers(0, n, 200)] = np.float32(0.5)   # ties -> pool order
    metric[rng.integers(0, n, 50)] = 0.0
    metric[rng.integers(0, n, 100)] = np.nan
    inl = rng.integers(2, 9, size=n).astype(np.float32)
    excl = rng.uniform(size=n) < 0.2
    cand = {i: float(metric[i]) for i in range(n) if not excl[i] and not math.isnan(metric[i]) and inl[i] > 4}
    exp = sorted(cand, key=cand.get)
    got = ops.sal_rank(torch.from_numpy(metric).cuda(), torch.from_numpy(inl).cuda(), torch.from_numpy(excl).cuda(), 4.0, n)
    assert got.cpu().tolist() == exp
    got = ops.sal_rank(torch.from_numpy(metric).cuda(), torch.from_numpy(inl).cuda(), None, 4.0, 64)
    cand = {i: float(metric[i]) for i in range(n) if not math.isnan(metric[i]) and inl[i] > 4}
    assert got.cpu().tolist() == sorted(cand, key=cand.get)[:64]
    assert ops.sal_rank(torch.full((7,), float("nan")).cuda(), torch.ones(7).cuda(), None, 0.0, 3).numel() == 0

    N, J = 300, 19
    pred = (rng.normal(size=(N, J, 3)) * 300).astype(np.float32)
    gt = np.concatenate([(pred.transpose(0, 2, 1) + rng.normal(size=(N, 3, J)) * 20), np.ones((N, 1, J))], axis=1).astype(np.float32)
    valid = (rng.uniform(size=(N, J)) < 0.97).astype(np.float32)
    valid[:200] = 1.0
    got = ops.mkpe(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(valid).cuda()).cpu().numpy()
    exp = np.array([SO.mkpe(pred[i], gt[i], valid[i]) for i in range(N)], dtype=np.float32)
    assert np.array_equal(np.isnan(got), np.isnan(exp)) and np.isnan(exp).sum() > 0
    ok = ~np.isnan(exp)
    np.testing.assert_allclose(got[ok], exp[ok], rtol=3e-7, atol=0)  # 1-2 ulp: float32 mean in a different order

    xyz = rng.normal(size=(N, J, 3)) * 300
    feats = ops.pose_features(torch.from_numpy(xyz).cuda(), 2).cpu().numpy()
    poses32 = [xyz[i].astype(np.float32).tolist() for i in range(N)]  # what sal_dict["pred_3d_keypoints"] holds
    exp = CO.stacked_features(poses32, [], 2).astype(np.float32)
    assert feats.shape == (N, 3 * J) and np.array_equal(feats, exp)
    assert np.array_equal(ops.pose_features(torch.from_numpy(xyz.astype(np.float32)).cuda(), 2).cpu().numpy(), exp)


def test_kmeans_assign_matches_sklearn_predict():
    """mval_kmeans_assign against the reference's per-candidate call self.kmeans.predict([kp])[0] (strategy.py:981-989):
    identical labels, margins >= 0, duplicated centres resolve to the first one."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sklearn.cluster import KMeans

    from multi_view_active_learning_b200 import ops

    rng = np.random.default_rng(3)
    for J, root, k, n in ((19, 2, 7, 600), (42, 21, 4, 150)):
        train = rng.normal(size=(300, 3 * J)) * 250
        km = KMeans(k, random_state=0, n_init=2).fit(train)
        pred32 = (rng.normal(size=(n, J, 3)) * 300).astype(np.float32)
        label, margin = ops.kmeans_assign(torch.from_numpy(pred32).cuda(), km.cluster_centers_, root)
        exp = []
        for i in range(n):
            kp = np.array(pred32[i].tolist()).T  # what sal_dict["pred_3d_keypoints"][guid] holds
            kp = (kp[0:3, :] - kp[0:3, root:root + 1]).flatten()
            exp.append(int(km.predict([kp])[0]))
        assert label.cpu().tolist() == exp
        m = margin.cpu().numpy()
        assert (m >= 0).all() and np.isfinite(m).all()
        dup = np.concatenate([km.cluster_centers_[:1], km.cluster_centers_])  # centre 0 twice
        l2, m2 = ops.kmeans_assign(torch.from_numpy(pred32).cuda(), dup, root)
        l2 = l2.cpu().numpy()
        assert np.array_equal(np.where(l2 == 0, 0, l2 - 1), np.asarray(exp)) and not (l2 == 1).any()
        assert (m2.cpu().numpy()[l2 == 0] == 0).all()
    empty, _ = ops.kmeans_assign(torch.zeros(0, 19, 3).cuda(), np.zeros((3, 57)), 2)
    assert empty.numel() == 0


def test_sal_cluster_balanced_pseudo_labels(tmp_path):
    """SAL with SAL.CLUSTER_FILE_PATH (reference strategy.py:37-52, 976-992): k-means over the cluster file's
    root-relative poses, then the ascending sal_metric order (device filter + sort) is walked and every cluster takes at
    most pseudo_num_frames // NUM_CLUSTERS frames.  Checked against a literal walk over the oracle's sal_dict."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import json

    rng = np.random.default_rng(11)
    path = tmp_path / "clusters.json"
    path.write_text(json.dumps({"g%d" % i: (rng.normal(size=(4, 19)) * 300).tolist() for i in range(60)}))
    cfg = make_cfg("TRIANGULATION", expr="SAL")
    cfg.SAL.CLUSTER_FILE_PATH, cfg.SAL.NUM_CLUSTERS = str(path), 3
    st = make_strategy(cfg)
    assert st.kmeans is not None and st.kmeans.cluster_centers_.shape == (3, 57)
    ds = FakeDataset(36, seed=13)
    ref = reference_emulation(ds, cfg)
    _, al_guids, sal_guids, sal_dict = st._sal_pseudo_labeling(ds, 4, 9, torch.nn.Identity())
    exp_al = SO.rank_nlargest(ref["al_metric"], 4)
    assert al_guids == exp_al
    cand = {g: m for g, m in sal_dict["sal_metric"].items()
            if g not in exp_al and not math.isnan(m) and sal_dict["inlier_count"][g] > cfg.SAL.INLIER_THRESHOLD}
    counter, exp = [0, 0, 0], []
    for g in sorted(cand, key=cand.get):
        kp = np.array(sal_dict["pred_3d_keypoints"][g]).T
        kp = (kp[0:3, :] - kp[0:3, 2:3]).flatten()
        c = st.kmeans.predict([kp])[0]
        if counter[c] < 9 // 3:
            counter[c] += 1
            exp.append(g)
    assert sal_guids == exp and 0 < len(exp) <= 9
    assert ds.pseudo_label_guids == exp


# --------------------------------------------------------------------------------------------------------------------
# round 2: the device-resident sal_dict (table.py) against the dict path on a 10k-frame pool
# --------------------------------------------------------------------------------------------------------------------
class DevicePool:
    """Loader + dataset surface over device-resident heat maps (identity pose estimator)."""

    def __init__(self, n, V=8, J=19, seed=3, batch=2048, n_labeled=50, dup_tail=0):
        from multi_view_active_learning_b200 import ops

        pool = S.make_pool(n, V, J, seed=seed, p_outlier=0.15, valid_prob=0.97)
        self.n, self.batch = n, batch
        self.hm = ops.synth_heatmaps(torch.from_numpy(pool["centres"]).cuda(), 64, 64, 1.0, 0.05, seed)
        self.P = torch.from_numpy(pool["P"]).cuda()
        self.valid = torch.from_numpy(pool["valid"].astype(np.float32)).cuda()
        self.gt = torch.from_numpy(np.concatenate([pool["X"].transpose(0, 2, 1), np.ones((n, 1, J))], axis=1).astype(np.float32)).cuda()
        self.frame = torch.arange(n, dtype=torch.int64).cuda() * 2 + 5
        self.pose = (160000 + torch.arange(n, dtype=torch.int64) % 7).cuda()
        self.order = list(range(n)) + list(range(dup_tail))  # DistributedSampler-style repeat of the head
        rng = np.random.default_rng(seed + 1)
        self.labeled_data = [{"3d_keypoints": rng.normal(size=(4, J)) * 300} for _ in range(n_labeled)]
        self.pseudo_label_guids, self.labeled, self.pseudo_seen = [], [], None

    def resample_unlabeled_data(self):
        pass

    def get_al_dict_for_coreset(self):
        return {i: np.array(d["3d_keypoints"]).transpose([1, 0]) for i, d in enumerate(self.labeled_data)}

    def label_by_frame_guids(self, guids):
        self.labeled = list(guids)

    def pseudo_label_by_frame_guids(self, guids, pseudo_labels):
        self.pseudo_label_guids = guids
        self.pseudo_seen = [np.array(pseudo_labels[g]).transpose([1, 0]) for g in guids]

    def loader(self):
        for o in range(0, len(self.order), self.batch):
            idx = torch.as_tensor(self.order[o:o + self.batch]).cuda()
            yield {"images": self.hm[idx], "proj_matrices": self.P[idx], "joint_valid": self.valid[idx],
                   "3d_keypoints": self.gt[idx], "pose": self.pose[idx], "frame_id": self.frame[idx]}


@pytest.mark.parametrize("strategy,expr", [("TRIANGULATION", "AL"), ("TRIANGULATION", "SAL"), ("CORESET", "AL"), ("HP", "SAL")])
def test_device_table_selection_equals_dict_path(strategy, expr):
    """_sal_pseudo_labeling over the device-resident table must select exactly what the same code selects from the
    materialised reference-style dicts (the round-1 path), on a 10 000-frame pool, incl. the sampler's repeated frames."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import random

    from multi_view_active_learning_b200 import utils as U
    from multi_view_active_learning_b200.table import LazyColumn, SalDict

    cfg = make_cfg(strategy, expr=expr)
    st = make_strategy(cfg)
    ds = DevicePool(10_000, dup_tail=3)
    st._get_dataloader = lambda d, bs, nw: ds.loader()
    random.seed(11)
    _, al_guids, sal_guids, sal = st._sal_pseudo_labeling(ds, 300, 40, torch.nn.Identity())
    assert isinstance(sal, SalDict) and isinstance(sal["al_metric"], LazyColumn)
    assert len(sal["al_metric"]) == 10_000 and sal.table.n == 10_000  # the 3 repeated frames were folded into their guids
    plain = sal.to_plain()
    assert list(plain["al_metric"])[:3] == ["160000-5", "160001-7", "160002-9"]
    # the same selection from plain dicts
    if strategy == "CORESET":
        exp_al = U.coreset.CoreSet(plain["pred_3d_keypoints"], ds.get_al_dict_for_coreset(), 2).select_batch(300)
    else:
        exp_al = st._rank_nlargest(plain["al_metric"], 300)
        finite = {g: v for g, v in plain["al_metric"].items() if not math.isnan(v)}
        assert exp_al == SO.rank_nlargest(finite, 300)
    assert al_guids == exp_al and ds.labeled == exp_al and len(set(al_guids)) == 300
    if expr == "SAL":
        cand = st._sal_candidates(plain, al_guids, [], cfg.SAL.INLIER_THRESHOLD, 80)
        assert cand == st._sal_candidates(sal, al_guids, [], cfg.SAL.INLIER_THRESHOLD, 80)
        lit = {g: m for g, m in plain["sal_metric"].items()
               if g not in al_guids and not math.isnan(m) and plain["inlier_count"][g] > cfg.SAL.INLIER_THRESHOLD}
        assert cand == sorted(lit, key=lit.get)[:80]
        random.seed(11)
        assert sal_guids == random.sample(cand, 40)
        assert all(np.array_equal(a, np.array(plain["pred_3d_keypoints"][g]).T) for a, g in zip(ds.pseudo_seen, sal_guids))
        # a long exclusion list (pseudo labels of earlier iterations) goes through the vectorised guid lookup
        old = list(plain["sal_metric"])[100:700]
        assert st._sal_candidates(sal, al_guids, old, cfg.SAL.INLIER_THRESHOLD, 50) == st._sal_candidates(
            plain, al_guids, old, cfg.SAL.INLIER_THRESHOLD, 50)


def test_first_occurrence_and_topk_merge_kernels():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from multi_view_active_learning_b200 import ops

    rng = np.random.default_rng(5)
    n = 20_000
    pose = rng.integers(0, 5, size=n)
    frame = rng.integers(0, 9000, size=n)  # plenty of repeated guids
    keep, src, unique = ops.first_occurrence(torch.from_numpy(pose).cuda(), torch.from_numpy(frame).cuda())
    first, last = {}, {}
    for i, k in enumerate(zip(pose.tolist(), frame.tolist())):
        first.setdefault(k, i)
        last[k] = i
    assert int(unique.item()) == len(first)
    keep, src = keep.cpu().numpy(), src.cpu().numpy()
    assert sorted(np.nonzero(keep)[0].tolist()) == sorted(first.values())
    assert all(src[i] == last[k] for k, i in first.items()) and (src[keep == 0] == -1).all()
    _, _, u = ops.first_occurrence(torch.tensor([1 << 40, 3]).cuda(), torch.tensor([1, 2]).cuda())
    assert int(u.item()) == -1  # does not fit the packed key: the caller falls back to the host
    # cross-rank merge: 4 "ranks" with contiguous shards, ties across ranks, NaN padding
    scores = rng.normal(size=4000).round(1)
    scores[rng.integers(0, 4000, 50)] = np.nan
    k = 64
    idxs, vals = [], []
    for r in range(4):
        i, v, c = ops.topk_desc(torch.from_numpy(scores[r * 1000:(r + 1) * 1000]).cuda(), k, index_offset=r * 1000, fixed=True)
        assert int(c.item()) == k and i.shape[0] == k
        idxs.append(i)
        vals.append(v)
    mi, mv, mc = ops.topk_merge(torch.cat(vals), torch.cat(idxs), k)
    exp = SO.rank_nlargest({i: float(s) for i, s in enumerate(scores) if not math.isnan(s)}, k)
    assert int(mc.item()) == k and mi.cpu().tolist() == exp and mv.cpu().tolist() == [float(scores[i]) for i in exp]
    # fewer candidates than k: the tail is -1 / NaN
    i, v, c = ops.topk_desc(torch.tensor([1.0, float("nan"), 3.0], dtype=torch.float64).cuda(), 5, fixed=True)
    assert int(c.item()) == 2 and i.cpu().tolist() == [2, 0, -1, -1, -1] and np.isnan(v.cpu().numpy()[2:]).all()


def test_watchdog_is_reported_and_recovers():
    """A persistent kernel whose producer never issues (test hook) must not return garbage silently: the time-out is
    reported as MVAL_ERR_CUDA by mval_check_async / the next launch, the record is cleared and later launches work."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from multi_view_active_learning_b200 import _lib, ops

    pool = S.make_pool(64, 8, 19, seed=2)
    hm = ops.synth_heatmaps(torch.from_numpy(pool["centres"]).cuda(), 64, 64, 1.0, 0.05, 2)
    P = torch.from_numpy(pool["P"]).cuda()
    good = ops.score_pool(hm, P, 4)
    ops.check_async()
    lib = _lib.load()
    try:
        _lib.check(lib.mval_debug_watchdog(2_000_000, 1))  # ~1 ms time-out, producers stalled
        ops.score_pool(hm, P, 4)
        with pytest.raises(_lib.MvalError, match="watchdog"):
            ops.check_async()
        ops.score_hp(hm)
        torch.cuda.synchronize()
        with pytest.raises(_lib.MvalError, match="watchdog"):
            ops.score_hp(hm)  # reported by the NEXT launch as well
    finally:
        _lib.check(lib.mval_debug_watchdog(0, 0))
    ops.check_async()  # the record was cleared by the reports
    again = ops.score_pool(hm, P, 4)
    ops.check_async()
    assert torch.equal(again["metric"], good["metric"]) and torch.equal(again["inlier_count"], good["inlier_count"])


@pytest.mark.parametrize("V,J", [(8, 19), (5, 19), (20, 42), (2, 3), (31, 19)])
def test_frame_aggregation_kernel_equals_the_reference_arithmetic(V, J):
    """mval_aggregate_map_scores against strategy._aggregate_map_scores (the host restatement of strategy.py:1151-1158 /
    1188-1193 / 1210-1215 that is pinned to the reference's golden values): AVG and STD of HP / MPE / BSB, bit for bit, with
    ragged validity masks (1 .. J valid joints: fewer than 8, up to 128 and more than 128 values per frame), NaN scores and both
    flavours of builtin sum()."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from multi_view_active_learning_b200 import ops, strategy as ST

    rng = np.random.default_rng(V * 100 + J)
    N = 700
    per_map = (rng.random((N, V, J)) * 0.3 + 0.5).astype(np.float32)
    per_map[rng.integers(0, N, 5), rng.integers(0, V, 5), rng.integers(0, J, 5)] = np.nan
    valid = rng.random((N, J)) < 0.8
    valid[:, 0] = True
    valid[:50] = True
    valid[50:60, 1:] = False  # a single valid joint: V values
    dev_map, dev_valid = torch.from_numpy(per_map).cuda(), torch.from_numpy(valid.astype(np.float32)).cuda()
    for kind in ("HP", "MPE", "BSB"):
        for config in ("AVG", "STD"):
            exp = np.asarray(ST._aggregate_map_scores(kind, config, per_map, valid), dtype=np.float64)
            got = ops.aggregate_map_scores(dev_map, dev_valid, kind, config, ST._SUM_IS_COMPENSATED).cpu().numpy()
            assert np.array_equal(np.isnan(exp), np.isnan(got)), (kind, config)
            ok = ~np.isnan(exp)
            assert np.array_equal(exp[ok], got[ok]), (kind, config, np.abs(exp[ok] - got[ok]).max())
    # the plain left-to-right sum of Python < 3.12
    plain = ops.aggregate_map_scores(dev_map, dev_valid, "HP", "AVG", False).cpu().numpy()
    x = np.where(valid[:, None, :], per_map, 0.0).astype(np.float64)
    for f in (0, 55, 300):
        s = 0.0
        for v in x[f][np.broadcast_to(valid[f][None, :], (V, J))].tolist():
            s += v
        m = V * int(valid[f].sum())
        assert plain[f] == s / m or (np.isnan(plain[f]) and np.isnan(s))
