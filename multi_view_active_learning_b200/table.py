"""Device-resident form of the reference's ``sal_dict`` (strategy.py:1009-1015, 1115-1133).

The reference fills five guid-keyed ``OrderedDict``s frame by frame; for a pool of a million frames that is millions of
Python objects (the ``pred_3d_keypoints`` lists alone are N * J * 3 floats) built on every rank before the selection
looks at a few hundred of them.  Here the pool's results stay where the kernels left them -- one row per guid, in the
reference's insertion order, as CUDA tensors -- and the dicts are *views*:

  PoolTable   the rows: pose / frame ids (the guid is "%s-%s" % (pose, frame)), al_metric, sal_metric, inlier_count,
              pred_3d_keypoints, mkpe;  row <-> guid translation on demand.
  LazyColumn  a read-only ``Mapping`` guid -> value over one column with the reference's value types (Python floats,
              [J][3] lists); materialises on access (``to_dict()`` gives the reference's OrderedDict).
  SalDict     the ``dict`` of the five columns that ``_compute_sal_dict`` returns; ``.table`` is what the selection code
              in strategy.py / utils/coreset.py uses to stay on the device.
"""
import json
from collections import OrderedDict
from collections.abc import Mapping

import numpy as np
import torch

COLUMNS = ("al_metric", "sal_metric", "inlier_count", "pred_3d_keypoints", "mkpe")


class PoolTable:
    def __init__(self, pose, frame, al, sal, inl, pred, mkpe):
        """All tensors on one device, row i = the i-th guid the reference would have inserted.
        pose / frame int64 [N]; al float64 [N] (already holding the float32 roundings where the reference's tensor is
        float32); sal / inl / mkpe float32 [N]; pred float32 [N, J, 3]."""
        self.pose, self.frame = pose, frame
        self.values = {"al_metric": al, "sal_metric": sal, "inlier_count": inl, "pred_3d_keypoints": pred, "mkpe": mkpe}
        self.n = int(pose.shape[0])
        self._ids_host = None
        self._guids = None
        self._known = {}  # guid -> row for every guid formatted so far
        self._index = None  # (sorted packed keys, rows) or dict, built on the first foreign lookup
        self._host = {}

    # ---- ids -------------------------------------------------------------------------------------------------------
    def ids_host(self):
        if self._ids_host is None:
            self._ids_host = (self.pose.cpu().numpy(), self.frame.cpu().numpy())
        return self._ids_host

    def guid_at(self, rows):
        """rows: iterable of ints / 1-D tensor -> list of guid strings (formatted like strategy.py:1121-1122)."""
        if torch.is_tensor(rows):
            rows = rows.cpu().tolist()
        pose, frame = self.ids_host()
        out = []
        for r in rows:
            g = "%s-%s" % (int(pose[r]), int(frame[r]))
            self._known.setdefault(g, int(r))
            out.append(g)
        return out

    def guids(self):
        """All guids in row order (built once; O(N) Python strings -- nothing on the selection path needs it)."""
        if self._guids is None:
            pose, frame = self.ids_host()
            self._guids = ["%s-%s" % pf for pf in zip(pose.tolist(), frame.tolist())]
        return self._guids

    @staticmethod
    def _parse(guid):
        i = guid.index("-", 1)  # a leading "-" belongs to a negative pose id
        return int(guid[:i]), int(guid[i + 1:])

    def _build_index(self):
        pose, frame = self.ids_host()
        if self.n and (pose.min() < 0 or frame.min() < 0 or pose.max() >= 1 << 31 or frame.max() >= 1 << 32):
            self._index = {g: i for i, g in enumerate(self.guids())}
            return
        keys = (pose.astype(np.int64) << 32) | frame.astype(np.int64)
        order = np.argsort(keys, kind="stable")
        self._index = (keys[order], order)

    def row_of(self, guid):
        r = self._known.get(guid)
        if r is not None:
            return r
        if self._index is None:
            self._build_index()
        if isinstance(self._index, dict):
            return self._index[guid]
        try:
            p, f = self._parse(guid)
        except (ValueError, AttributeError):
            raise KeyError(guid)
        keys, order = self._index
        if not (0 <= p < 1 << 31 and 0 <= f < 1 << 32):
            raise KeyError(guid)
        k = (p << 32) | f
        i = int(np.searchsorted(keys, k))
        if i >= len(keys) or int(keys[i]) != k:
            raise KeyError(guid)
        r = int(order[i])
        self._known[guid] = r
        return r

    def rows_of(self, guids, missing_ok=False):
        """Rows of the given guids, in their order.  Guids this table formatted itself (guid_at: the selections) come out of
        a dict; only foreign ones (e.g. pseudo labels of earlier iterations) need the sorted index -- one vectorised search
        for all of them."""
        guids = list(guids)
        rows = [self._known.get(g) for g in guids]
        unknown = [i for i, r in enumerate(rows) if r is None]
        if unknown:
            if len(unknown) > 64 and self._index is None:
                self._build_index()
            if len(unknown) > 64 and not isinstance(self._index, dict):
                try:
                    parsed = np.array([self._parse(guids[i]) for i in unknown], dtype=np.int64).reshape(-1, 2)
                except (ValueError, AttributeError):
                    parsed = None
                if parsed is not None and (parsed >= 0).all() and (parsed[:, 0] < 1 << 31).all() and (parsed[:, 1] < 1 << 32).all():
                    keys, order = self._index
                    k = (parsed[:, 0] << 32) | parsed[:, 1]
                    pos = np.minimum(np.searchsorted(keys, k), max(len(keys) - 1, 0))
                    hit = (keys[pos] == k) if len(keys) else np.zeros(len(k), dtype=bool)
                    found = order[pos]
                    for n_, i in enumerate(unknown):
                        if hit[n_]:
                            rows[i] = int(found[n_])
                    unknown = [i for n_, i in enumerate(unknown) if not hit[n_]]
                    if unknown and not missing_ok:
                        raise KeyError(guids[unknown[0]])
                    unknown = []
            for i in unknown:
                try:
                    rows[i] = self.row_of(guids[i])
                except KeyError:
                    if not missing_ok:
                        raise
        return [r for r in rows if r is not None]

    # ---- values ----------------------------------------------------------------------------------------------------
    def host(self, name):
        if name not in self._host:
            self._host[name] = self.values[name].cpu().numpy()
        return self._host[name]

    def column(self, name):
        return LazyColumn(self, name)

    def as_sal_dict(self):
        return SalDict(self)


class LazyColumn(Mapping):
    """Read-only guid -> value view of one PoolTable column; equal to (and convertible into) the OrderedDict the
    reference builds.  ``device_values`` is the CUDA tensor behind it, in key order."""

    def __init__(self, table, name):
        self.table, self.name = table, name
        self._rows = {}  # row -> value for rows fetched with prefetch()

    @property
    def device_values(self):
        return self.table.values[self.name]

    def __len__(self):
        return self.table.n

    def __iter__(self):
        return iter(self.table.guids())

    def __contains__(self, guid):
        try:
            self.table.row_of(guid)
            return True
        except (KeyError, TypeError):
            return False

    def _value(self, row):
        v = self._rows.get(row)
        if v is not None:
            return v
        return self.table.host(self.name)[row].tolist()

    def __getitem__(self, guid):
        return self._value(self.table.row_of(guid))

    def prefetch(self, guids):
        """Fetches the values of these guids with ONE device gather (instead of the whole column), e.g. the pseudo labels
        the dataset is about to read with ``pseudo_labels[guid]`` (strategy.py:998-1000)."""
        if self.name in self.table._host:
            return self
        rows = self.table.rows_of(guids, missing_ok=True)
        if rows:
            dev = self.device_values
            picked = dev[torch.as_tensor(rows, dtype=torch.int64, device=dev.device)].cpu().numpy()
            for r, v in zip(rows, picked):
                self._rows[r] = v.tolist()
        return self

    def keys(self):
        return self.table.guids()

    def values(self):
        return self.table.host(self.name).tolist()

    def items(self):
        return zip(self.table.guids(), self.values())

    def to_dict(self):
        return OrderedDict(self.items())

    def __eq__(self, other):
        if isinstance(other, Mapping):
            return dict(self.items()) == dict(other.items())
        return NotImplemented

    __hash__ = None

    def __repr__(self):
        return "LazyColumn(%s, %d guids)" % (self.name, len(self))


class SalDict(dict):
    """{"al_metric": ..., "sal_metric": ..., "inlier_count": ..., "pred_3d_keypoints": ..., "mkpe": ...} with LazyColumn
    values over one PoolTable (``.table``)."""

    def __init__(self, table):
        super().__init__((name, table.column(name)) for name in COLUMNS)
        self.table = table

    def to_plain(self):
        """The reference's structure: a dict of five OrderedDicts of Python values."""
        return {k: (v.to_dict() if isinstance(v, LazyColumn) else v) for k, v in self.items()}


def dumps(obj):
    """json.dumps that writes LazyColumns as the dicts they stand for: the SAL-DICT-ITER-k payload (strategy.py:123-128)."""
    return json.dumps(obj, default=lambda o: o.to_dict() if isinstance(o, LazyColumn) else json.JSONEncoder().default(o))
