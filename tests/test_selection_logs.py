"""Host logic either side of the path (no GPU): the SAMPLED-GUID / SAL-GUID / SAL-DICT files sample_next_batch leaves
behind (reference strategy.py:112-134) and their replay by restore_dataset (strategy.py:315-336)."""
import json
import os
from collections import OrderedDict
from types import SimpleNamespace as NS

from multi_view_active_learning_b200.strategy import ActiveLearningStrategy


class _Dataset:
    def __init__(self, n):
        self.unlabeled_data = OrderedDict(("160422-%d" % i, {"frame_id": i}) for i in range(n))
        self.labeled_data, self.pseudo_label_guids = [], []

    def label_by_frame_guids(self, guids):
        for g in guids:
            self.labeled_data.append(self.unlabeled_data.pop(g))


def _cfg(tmp_path, expr="SAL"):
    return NS(EXPR_TYPE=expr, EXPR_NAME="run0", LOG_DIR=str(tmp_path), RANDOM_SEED=1307, DATA=NS(NUM_JOINTS=19, TYPE="panoptic"),
              SAL=NS(INLIER_THRESHOLD=7, CLUSTER_FILE_PATH="", NUM_CLUSTERS=10), AL=NS(STRATEGY="TRIANGULATION"))


def test_iteration_files_and_restore(tmp_path):
    st = ActiveLearningStrategy(_cfg(tmp_path))
    ds = _Dataset(12)
    st.sample_next_batch(ds, 3, 0, None, iteration=0, rank=0)  # iteration 0: seeded random sample (strategy.py:868-878)
    first = st.last_al_guids
    run = os.path.join(str(tmp_path), "run0")
    assert json.loads(open(os.path.join(run, "SAMPLED-GUID-ITER-0")).readline()) == first
    assert not os.path.exists(os.path.join(run, "SAL-DICT-ITER-0"))  # only written for iteration > 0 (:112-125)

    sal_dict = {"al_metric": OrderedDict(a=1.5, b=float("nan")), "sal_metric": OrderedDict(a=0.25, b=0.5),
                "inlier_count": OrderedDict(a=8.0, b=3.0), "pred_3d_keypoints": OrderedDict(a=[[0.0, 1.0, 2.0]], b=[[3.0, 4.0, 5.0]]),
                "mkpe": OrderedDict(a=1.0, b=2.0)}
    rest = [g for g in ds.unlabeled_data]
    second, pseudo = rest[:2], rest[2:3]
    st._after_sampling(1, 0, second, pseudo, sal_dict)
    assert json.loads(open(os.path.join(run, "SAMPLED-GUID-ITER-1")).read()) == second
    assert json.loads(open(os.path.join(run, "SAL-GUID-ITER-1")).read()) == pseudo
    back = json.loads(open(os.path.join(run, "SAL-DICT-ITER-1")).read())
    assert list(back) == list(sal_dict) and back["pred_3d_keypoints"]["b"] == [[3.0, 4.0, 5.0]] and back["al_metric"]["b"] != back["al_metric"]["b"]
    st._after_sampling(2, 1, ["x"], [], {})  # other ranks write nothing (:81, :126)
    assert not os.path.exists(os.path.join(run, "SAMPLED-GUID-ITER-2"))

    ds2 = _Dataset(12)
    st.restore_dataset(ds2, 2)
    assert [f["frame_id"] for f in ds2.labeled_data] == [int(g.split("-")[1]) for g in first + second]
    assert ds2.pseudo_label_guids == pseudo
    assert len(ds2.unlabeled_data) == 12 - 5


def test_al_experiment_writes_no_sal_guid_file(tmp_path):
    st = ActiveLearningStrategy(_cfg(tmp_path, expr="AL"))
    st._after_sampling(1, 0, ["a"], [], {"al_metric": {}})
    run = os.path.join(str(tmp_path), "run0")
    assert os.path.exists(os.path.join(run, "SAL-DICT-ITER-1")) and not os.path.exists(os.path.join(run, "SAL-GUID-ITER-1"))
    ds = _Dataset(3)
    ds.unlabeled_data["a"] = {"frame_id": 99}
    open(os.path.join(run, "SAMPLED-GUID-ITER-0"), "w").write(json.dumps(["160422-0"]))
    st.restore_dataset(ds, 2)
    assert ds.pseudo_label_guids == [] and len(ds.labeled_data) == 2


def test_frame_aggregation_of_precomputed_map_scores():
    """_compute_map_score_batch with the per-map scores the fused pass already produced (no kernel call): AVG is the
    reference's Python sum(x) / len(x) and STD its np.std over (view, valid joint), view-major (strategy.py:1151-1158,
    1188-1193, 1210-1215); a [J] validity vector applies to every frame."""
    import numpy as np
    import pytest
    import torch

    from multi_view_active_learning_b200.strategy import ScoringSelectionMixin as M
    from oracle import scores_oracle as SO

    rng = np.random.default_rng(5)
    B, V, J = 4, 5, 7
    scores = rng.random((B, V, J)).astype(np.float32)
    valid = rng.random((B, J)) < 0.7
    valid[:, 0] = True
    per_map = torch.from_numpy(np.where(valid[:, None, :], scores, np.nan).astype(np.float32))
    hm = torch.zeros(B, V, J, 1, 1)  # only its batch size is looked at when per_map is given
    for kind in ("HP", "MPE", "BSB"):
        for config in ("AVG", "STD"):
            got = M._compute_map_score_batch(kind, config, hm, torch.from_numpy(valid), per_map)
            exp = [SO.reduce_frame_score(scores[b], valid[b], config, kind) for b in range(B)]
            assert [float(g) for g in got] == [float(e) for e in exp]
            # HP aggregates Python floats in double; MPE / BSB aggregate np.float32 scalars in float32 (NumPy >= 2)
            assert all(isinstance(g, np.float32) for g in got) == (kind != "HP")
    one = M._compute_map_score_batch("BSB", "AVG", hm, torch.from_numpy(valid[0]), torch.from_numpy(scores))
    assert [float(g) for g in one] == [float(SO.reduce_frame_score(scores[b], valid[0], "AVG", "BSB")) for b in range(B)]
    assert M._compute_map_score_batch("HP", "MEDIAN", hm, torch.from_numpy(valid), per_map) == [None] * B
    with pytest.raises(NotImplementedError):
        M._compute_map_score_batch("MPE", "MEDIAN", hm, torch.from_numpy(valid), per_map)


def test_vectorised_frame_aggregation_is_the_reference_arithmetic():
    """_aggregate_map_scores evaluates all frames of a batch at once; it must return, bit for bit, what the reference's
    per-frame expressions return: builtin sum() of Python floats (Neumaier-compensated on Python >= 3.12) / len and
    np.std of a float64 array for HP; float32 left-to-right sums and float32 np.std for MPE / BSB -- for ragged validity
    patterns, 8 x 19 and 20 x 42 maps per frame, and values that expose the summation order."""
    import numpy as np
    import pytest

    from multi_view_active_learning_b200 import strategy as ST
    from oracle import scores_oracle as SO

    rng = np.random.default_rng(9)
    wide = (rng.random((64, 300)) * rng.choice([1e-8, 1.0, 1e8], size=(64, 300))).astype(np.float32).astype(np.float64)
    assert np.array_equal(ST._python_float_sums(np.ascontiguousarray(wide.T)), np.array([sum(r.tolist()) for r in wide]))
    assert ST._python_float_sums(np.zeros((0, 3))).tolist() == [0.0, 0.0, 0.0]
    for B, V, J in ((37, 8, 19), (9, 20, 42), (5, 2, 3)):
        scores = (rng.random((B, V, J)) * rng.choice([1e-4, 1.0, 50.0], size=(B, V, J))).astype(np.float32)
        valid = rng.random((B, J)) < 0.8
        valid[:, 0] = True
        valid[B // 2:] = valid[B // 2]  # several frames share a pattern, others are unique
        for kind in ("HP", "MPE", "BSB"):
            for config in ("AVG", "STD"):
                got = ST._aggregate_map_scores(kind, config, scores, valid)
                exp = [SO.reduce_frame_score(scores[b], valid[b], config, kind) for b in range(B)]
                assert got.dtype == (np.float64 if kind == "HP" else np.float32)
                assert [float(g) for g in got] == [float(e) for e in exp], (kind, config, B)
    none_valid = np.zeros((2, 3), dtype=bool)
    with pytest.raises(ZeroDivisionError):
        ST._aggregate_map_scores("HP", "AVG", np.ones((2, 2, 3), np.float32), none_valid)


def test_compute_sal_dict_host_flow_without_device(monkeypatch):
    """The host side of _compute_sal_dict (reference strategy.py:1004-1147) with the device calls stubbed out: per-batch
    results are only accumulated, the HP / MPE / BSB frame scores are formed once per pool from the per-map scores of
    the fused pass, dicts come out in loader order with the reference's roundings (float32 for everything except
    TRIANGULATION's metric and HP's STD)."""
    import numpy as np
    import torch

    from multi_view_active_learning_b200 import ops, strategy as ST
    from oracle import scores_oracle as SO

    V, J, B, n_batches = 2, 5, 4, 3
    rng = np.random.default_rng(21)
    maps = rng.random((n_batches, B, V, J)).astype(np.float32)
    valid = rng.random((n_batches, B, J)) < 0.7
    valid[..., 0] = True
    seen = []

    def fake_triangulation_batch(hm, P, stride, joint_valid, **kw):
        k = len(seen)
        seen.append(kw)
        out = {"metric": torch.arange(B, dtype=torch.float64) + k + 1.0 / 3.0,
               "inlier_count": torch.full((B,), 8, dtype=torch.int32),
               "keypoints_3d": torch.full((B, J, 3), 0.1, dtype=torch.float64) * (k + 1)}
        if kw.get("map_score"):
            out["map_score"] = torch.from_numpy(np.where(valid[k][:, None, :], maps[k], np.nan).astype(np.float32))
        return out

    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(ops, "mkpe", lambda p, g, v: torch.zeros(p.shape[0]))
    from conftest import first_occurrence_numpy

    monkeypatch.setattr(ops, "first_occurrence", first_occurrence_numpy)
    monkeypatch.setattr(ST.triangulation, "triangulation_batch", fake_triangulation_batch)

    def loader():
        for k in range(n_batches):
            yield {"images": torch.zeros(B, V, J, 8, 8), "proj_matrices": torch.zeros(B, V, 3, 4, dtype=torch.float64),
                   "joint_valid": torch.from_numpy(valid[k].astype(np.float32)), "3d_keypoints": torch.zeros(B, 4, J),
                   "pose": torch.full((B,), 160422), "frame_id": torch.arange(B) + 10 * k}

    for strategy, config in (("TRIANGULATION", "AVG"), ("HP", "AVG"), ("HP", "STD"), ("MPE", "AVG"), ("BSB", "STD"), ("CORESET", "AVG")):
        del seen[:]
        cfg = _cfg("/tmp")
        cfg.POSE_ESTIMATOR = NS(STRIDE=4)
        cfg.AL = NS(STRATEGY=strategy, USE_SOFTARGMAX=False, USE_REPROJECTION_XE=False, REPROJECTION_SIGMA=1.0,
                    HP_CONFIG=config, MPE_CONFIG=config, BSB_CONFIG=config)
        st = ActiveLearningStrategy(cfg)
        st._compute_batch_heatmap = lambda pe, d: d["images"].reshape(-1, J, 8, 8)
        sal = st._compute_sal_dict(loader(), None)
        guids = ["160422-%d" % (10 * k + i) for k in range(n_batches) for i in range(B)]
        assert all(list(sal[name]) == guids for name in sal)
        assert all(kw["frame_keys"] is None for kw in seen)  # 1 view pair: no subset to key (V >= 12 passes the guid keys)
        assert all(kw["map_score"] == (strategy if strategy in ("HP", "MPE", "BSB") else None) for kw in seen)
        assert sal["sal_metric"]["160422-11"] == float(np.float32(1 + 1 + 1.0 / 3.0))  # torch.Tensor([metric]) is float32
        assert sal["pred_3d_keypoints"]["160422-20"][0][0] == float(np.float32(0.1 * 3))
        for k in range(n_batches):
            for i in range(B):
                got = sal["al_metric"]["160422-%d" % (10 * k + i)]
                if strategy == "TRIANGULATION":
                    assert got == i + k + 1.0 / 3.0  # the float64 metric itself (:1074-1075)
                elif strategy == "CORESET":
                    assert got == 0.0
                else:
                    exp = SO.reduce_frame_score(maps[k, i], valid[k, i], config, strategy)
                    exact = strategy == "HP" and config == "STD"  # the only float64 tensor among them (:1081-1085)
                    assert got == (float(exp) if exact else float(np.float32(exp))), (strategy, config, k, i)
