"""CPU tests of the device-resident sal_dict views (multi_view_active_learning_b200/table.py): a PoolTable over CPU tensors
must behave exactly like the five OrderedDicts strategy.py:1115-1133 builds, entry by entry."""
import json
import math
from collections import OrderedDict

import numpy as np
import pytest
import torch

from multi_view_active_learning_b200.table import COLUMNS, LazyColumn, PoolTable, SalDict, dumps


def make_table(n=50, J=4, seed=0):
    rng = np.random.default_rng(seed)
    pose = torch.from_numpy(rng.integers(160000, 160004, size=n))
    frame = torch.arange(n, dtype=torch.int64) * 3 + 7
    al = torch.from_numpy(rng.normal(size=n))
    al[5] = float("nan")
    sal = torch.from_numpy(rng.normal(size=n).astype(np.float32))
    inl = torch.from_numpy(rng.integers(2, 9, size=n).astype(np.float32))
    pred = torch.from_numpy(rng.normal(size=(n, J, 3)).astype(np.float32) * 100)
    mkpe = torch.from_numpy(rng.uniform(size=n).astype(np.float32))
    t = PoolTable(pose, frame, al, sal, inl, pred, mkpe)
    # what the reference's loop leaves behind (values are .item() / .tolist() of float32 / float64 tensors)
    ref = {k: OrderedDict() for k in COLUMNS}
    for i in range(n):
        g = "%s-%s" % (pose[i].item(), frame[i].item())
        ref["al_metric"][g] = al[i].item()
        ref["sal_metric"][g] = sal[i].item()
        ref["inlier_count"][g] = inl[i].item()
        ref["pred_3d_keypoints"][g] = pred[i].numpy().tolist()
        ref["mkpe"][g] = mkpe[i].item()
    return t, ref


def same(a, b):
    if isinstance(a, float):
        return (math.isnan(a) and math.isnan(b)) or a == b
    return a == b


def test_columns_equal_the_reference_dicts():
    t, ref = make_table()
    sd = t.as_sal_dict()
    assert isinstance(sd, dict) and list(sd) == list(COLUMNS) and sd.table is t
    for name in COLUMNS:
        col = sd[name]
        assert isinstance(col, LazyColumn) and len(col) == len(ref[name])
        assert list(col) == list(ref[name]) == list(col.keys())
        for g, v in ref[name].items():
            assert g in col and same(col[g], v) and type(col[g]) is type(v)
        assert all(same(a, b) for a, b in zip(col.values(), ref[name].values()))
        assert [k for k, _ in col.items()] == list(ref[name])
        d = col.to_dict()
        assert isinstance(d, OrderedDict) and list(d) == list(ref[name])
    assert sd["sal_metric"] == ref["sal_metric"] and sd["pred_3d_keypoints"] == ref["pred_3d_keypoints"]
    assert "1-2" not in sd["mkpe"] and "nonsense" not in sd["mkpe"] and 17 not in sd["mkpe"]
    with pytest.raises(KeyError):
        sd["mkpe"]["160000-999999"]
    # the SAL-DICT-ITER-k payload (strategy.py:123-128)
    assert dumps(sd) == json.dumps(ref)
    assert json.loads(dumps(sd.to_plain()).replace("NaN", "null")) == json.loads(json.dumps(ref).replace("NaN", "null"))
    with pytest.raises(TypeError):
        json.dumps(sd)  # a LazyColumn is not silently written as something else


def test_row_lookup_and_prefetch():
    t, ref = make_table(n=300, J=3, seed=4)
    guids = t.guid_at(torch.tensor([17, 3, 250]))
    assert guids == [list(ref["mkpe"])[i] for i in (17, 3, 250)]
    assert t.rows_of(guids) == [17, 3, 250]
    t2, _ = make_table(n=300, J=3, seed=4)  # a fresh table: lookups go through the packed-key index
    assert t2.rows_of(list(ref["mkpe"])[100:110]) == list(range(100, 110))
    assert t2.rows_of(["9-9", list(ref["mkpe"])[5]], missing_ok=True) == [5]
    t3, _ = make_table(n=300, J=3, seed=4)  # long lists take one vectorised search
    everything = list(ref["mkpe"])
    assert t3.rows_of(everything[::-1]) == list(range(299, -1, -1))
    assert t3.rows_of(everything + ["9-9"], missing_ok=True) == list(range(300))
    with pytest.raises(KeyError):
        t3.rows_of(everything + ["9-9"])
    col = t2.column("pred_3d_keypoints").prefetch(list(ref["mkpe"])[40:44])
    assert "pred_3d_keypoints" not in t2._host  # the column itself was not copied
    for g in list(ref["mkpe"])[40:44]:
        assert col[g] == ref["pred_3d_keypoints"][g]
    assert "pred_3d_keypoints" not in t2._host
    assert col[list(ref["mkpe"])[0]] == ref["pred_3d_keypoints"][list(ref["mkpe"])[0]]  # falls back to the host copy


def test_ids_beyond_32_bits_fall_back_to_a_host_index():
    n = 6
    pose = torch.tensor([1 << 40, 5, -3, 7, 7, 2])
    frame = torch.tensor([1, 1 << 33, 4, -8, 9, 0])
    z = torch.zeros(n)
    t = PoolTable(pose, frame, z.double(), z, z, torch.zeros(n, 2, 3), z)
    guids = ["%s-%s" % (p, f) for p, f in zip(pose.tolist(), frame.tolist())]
    assert t.guids() == guids
    fresh = PoolTable(pose, frame, z.double(), z, z, torch.zeros(n, 2, 3), z)
    assert fresh.rows_of(guids) == list(range(n))
    assert "-3-4" in fresh.column("mkpe") and "7--8" in fresh.column("mkpe")


def test_empty_table():
    z = torch.zeros(0)
    t = PoolTable(z.long(), z.long(), z.double(), z, z, torch.zeros(0, 19, 3), z)
    sd = SalDict(t)
    assert len(sd["al_metric"]) == 0 and list(sd["pred_3d_keypoints"].items()) == [] and dumps(sd) == json.dumps(
        {k: {} for k in COLUMNS})
