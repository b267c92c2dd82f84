"""The drop-in recipe of INTEGRATION.md, executed: ScoringSelectionMixin / SelectionLogMixin mixed over the UNMODIFIED
reference class (imported from /root/reference -- dev container only), with both import styles of the reference satisfied
(workflow.py:20-26 imports the strategy package-relatively, strategy.py:24-25 imports ``utils`` top-level), and
``sample_next_batch`` called the way workflow.py:64-71 calls it.  The device calls are replaced by the CPU oracle, so what
runs here is exactly the host logic that ships, against the reference's own flow fixture."""
import importlib
import json
import os
import random
import re
import sys
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, first_occurrence_numpy
from oracle import flow_oracle as FO
from oracle import scores_oracle as SO
from oracle import triangulation_oracle as O
from oracle.ref_import import REFERENCE_ROOT, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs the reference tree (/root/reference)")


def _documented_recipe():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# strategy_b200\.py.*?)```", text, flags=re.S)
    assert block, "INTEGRATION.md lost its module-level recipe"
    return block.group(1)


def _oracle_device_layer(monkeypatch, ST, ops):
    """Every device entry the flow touches, served by the CPU oracle / numpy (TEST infrastructure: the product has no CPU
    path)."""
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)

    def triangulation_batch(hm, P, stride, joint_valid, **kw):
        ref = O.triangulate_pool(hm.numpy(), np.asarray(P), stride, np.asarray(joint_valid) != 0)
        return {"metric": torch.from_numpy(ref["metric"]), "inlier_count": torch.from_numpy(ref["inlier_count"].astype(np.int32)),
                "keypoints_3d": torch.from_numpy(ref["keypoints_3d"])}

    def topk_desc(scores, k, index_offset=0, **_):
        s = scores.numpy()
        order = [i for i in np.lexsort((np.arange(len(s)), -s)) if not np.isnan(s[i])][:k]
        return torch.tensor(order, dtype=torch.int64) + index_offset, torch.from_numpy(s[order])

    def sal_rank(metric, inliers, excluded, thr, k):
        m, c = metric.numpy(), inliers.numpy()
        ex = np.zeros(len(m), bool) if excluded is None else excluded.numpy().astype(bool)
        cand = [i for i in range(len(m)) if not ex[i] and not np.isnan(m[i]) and c[i] > thr]
        return torch.tensor(sorted(cand, key=lambda i: m[i])[:k], dtype=torch.int64)

    monkeypatch.setattr(ST.triangulation, "triangulation_batch", triangulation_batch)
    monkeypatch.setattr(ops, "topk_desc", topk_desc)
    monkeypatch.setattr(ops, "sal_rank", sal_rank)
    monkeypatch.setattr(ops, "first_occurrence", first_occurrence_numpy)
    monkeypatch.setattr(ops, "mkpe", lambda p, g, v: torch.tensor(
        [SO.mkpe(p[i].numpy(), g[i].numpy(), v[i].numpy()) for i in range(p.shape[0])], dtype=torch.float32))


def test_mixin_over_the_reference_class(monkeypatch, tmp_path):
    from multi_view_active_learning_b200 import ops, strategy as ST
    from oracle.make_golden import load_strategy
    from oracle.make_golden_flow import FlowDataset, flow_cfg
    from oracle.ref_import import load_reference

    load_reference()
    ref_strategy = load_strategy()  # top-level ``strategy`` of the reference, itself importing top-level ``utils``
    assert ref_strategy.__file__.startswith(REFERENCE_ROOT) and sys.modules["utils"].__file__.startswith(REFERENCE_ROOT)
    sys.modules["iopath.common.file_io"].PathManager.open = staticmethod(lambda path, mode="r": open(path, mode))

    class CfgNode(dict):  # enough of yacs for ``import config``
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        __setattr__ = dict.__setitem__

        def clone(self):
            return self
    sys.modules["yacs.config"].CfgNode = CfgNode

    # 1. the reference as the PACKAGE workflow.py expects (``from .strategy import ActiveLearningStrategy``)
    pkg = types.ModuleType("mval_reference_pkg")
    pkg.__path__ = [REFERENCE_ROOT]
    monkeypatch.setitem(sys.modules, "mval_reference_pkg", pkg)
    # 2. the documented strategy_b200.py, next to strategy.py, and the one changed import of workflow.py
    shim = types.ModuleType("mval_reference_pkg.strategy_b200")
    exec(compile(_documented_recipe(), "INTEGRATION.md", "exec"), shim.__dict__)
    monkeypatch.setitem(sys.modules, "mval_reference_pkg.strategy_b200", shim)
    workflow = importlib.import_module("mval_reference_pkg.workflow")
    assert workflow.ActiveLearningStrategy.__module__ == "mval_reference_pkg.strategy"  # the unmodified import
    monkeypatch.setattr(workflow, "ActiveLearningStrategy", shim.ActiveLearningStrategy)
    Mixed = workflow.ActiveLearningStrategy
    mro = [c.__name__ for c in Mixed.__mro__]
    assert mro[:4] == ["ActiveLearningStrategy", "SelectionLogMixin", "ScoringSelectionMixin", "ActiveLearningStrategy"]
    assert Mixed.sample_next_batch is ST.ScoringSelectionMixin.sample_next_batch
    assert Mixed._compute_sal_dict is ST.ScoringSelectionMixin._compute_sal_dict
    assert Mixed._evaluate_all is ref_strategy.ActiveLearningStrategy._evaluate_all  # training / evaluation: the reference's

    _oracle_device_layer(monkeypatch, ST, ops)
    gold = dict(np.load(os.path.join(GOLDEN, "flow_triangulation_sal_w1.npz"), allow_pickle=False))
    pool = json.loads(str(gold["pool"]))
    cfg = flow_cfg("TRIANGULATION", "SAL", 1, "", "AVG", pool["batch"])
    cfg.LOG_DIR, cfg.EXPR_NAME = str(tmp_path), "expr"
    os.makedirs(os.path.join(str(tmp_path), "expr"))
    strategy = Mixed(cfg)  # the reference's constructor (logger, PathManager, joint_root_index)
    assert strategy.joint_root_index == 2 and hasattr(strategy, "_pathmgr")
    order = FO.sampler_order(pool["n"], 1)
    strategy._get_dataloader = lambda ds, bs, nw: torch.utils.data.DataLoader(ds, batch_size=bs, num_workers=0, sampler=order)
    hist = []
    strategy.al_writer = types.SimpleNamespace(add_histogram=lambda tag, values, it: hist.append((tag, len(values))),
                                               add_scalar=lambda *a: None)
    ds = FlowDataset(**pool)
    random.seed(99)
    # workflow.py:64-71
    out = strategy.sample_next_batch(ds, 4, 3, torch.nn.Identity(), 1, rank=0)
    assert out is ds
    assert strategy.last_al_guids == gold["al_guids"].tolist()
    assert strategy.last_sal_guids == gold["sal_guids"].tolist()
    sal = strategy.last_sal_dict
    assert list(sal["al_metric"].keys()) == gold["guids"].tolist()
    np.testing.assert_array_equal(np.array(list(sal["al_metric"].values())), gold["al_metric"])
    np.testing.assert_array_equal(np.array(list(sal["sal_metric"].values())), gold["sal_metric"])
    np.testing.assert_array_equal(np.array(list(sal["pred_3d_keypoints"].values())), gold["pred_3d_keypoints"])
    assert len(ds.unlabeled_data) == pool["n"] - 4 and ds.pseudo_label_guids == gold["sal_guids"].tolist()
    np.testing.assert_array_equal(np.array([d["pseudo_3d_keypoints"] for d in ds.pseudo_labeled_data]), gold["pseudo_3d_keypoints"])
    # the rank-0 files of strategy.py:112-134, readable by the reference's restore_dataset (:315-336)
    base = os.path.join(str(tmp_path), "expr")
    assert json.load(open(os.path.join(base, "SAMPLED-GUID-ITER-1"))) == gold["al_guids"].tolist()
    assert json.load(open(os.path.join(base, "SAL-GUID-ITER-1"))) == gold["sal_guids"].tolist()
    dumped = json.load(open(os.path.join(base, "SAL-DICT-ITER-1")))
    assert list(dumped) == ["al_metric", "sal_metric", "inlier_count", "pred_3d_keypoints", "mkpe"]
    assert list(dumped["sal_metric"]) == gold["guids"].tolist()
    assert sorted(t for t, _ in hist) == ["sal/al_metric", "sal/inlier_count", "sal/mkpe", "sal/sal_metric"]
