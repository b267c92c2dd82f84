"""Coreset k-center greedy selection with the reference's class interface (utils/coreset.py:13-95)."""
from collections import OrderedDict

import numpy as np
import torch

from .. import ops
from ..table import LazyColumn


class CoreSet:
    """Same constructor, attributes and methods as the reference.  The distances are evaluated on the GPU in float32 in
    the canonical summation order (csrc/kcenter.cu), first-index tie-break like np.argmax.

    ``sal_dict`` may be the reference's guid -> [J][3] dict or the ``pred_3d_keypoints`` column of the device-resident
    table ``_compute_sal_dict`` returns (table.py).  In the second case the unlabeled feature rows are formed on the
    device from the CUDA column (mval_pose_features) -- no per-pose Python, no host copy -- and ``features`` /
    ``sal_keys`` (the reference's host-side attributes) are only materialised if somebody reads them.  Inside a
    torch.distributed job every rank holds the same table; the rows are then split contiguously over the ranks and the
    greedy rounds run sharded with one all_gather per round (pool.kcenter_greedy_sharded), every rank ending with the
    same picks -- the reference runs the identical selection redundantly on every rank."""

    def __init__(self, sal_dict, al_dict, joint_root_index, metric="euclidean", device=None, sharded=None):
        if metric != "euclidean":
            raise NotImplementedError("only the euclidean metric of the reference's call sites is built")
        self._column = sal_dict if isinstance(sal_dict, LazyColumn) else None
        self.sal_dict = sal_dict if self._column is not None else OrderedDict(sal_dict)
        self.al_dict = OrderedDict(al_dict)
        self._root = joint_root_index
        self.name = "kcenter"
        self.metric = metric
        self.max_distances = None
        self._n_unl, self._n_lab = len(sal_dict), len(al_dict)
        self.n_obs = self._n_unl + self._n_lab
        self.al_indices = list(range(self._n_unl, self.n_obs))
        self.already_selected = []
        self._features = None
        self._sal_keys = None
        lab = self._labeled_features()
        if self._column is not None:
            dev_col = self._column.device_values
            self._device = dev_col.device
            unl = ops.pose_features(dev_col, joint_root_index) if self._n_unl else torch.zeros((0, lab.shape[1]), device=self._device)
            self._feat = torch.cat([unl, torch.from_numpy(lab.astype(np.float32)).to(self._device)]).contiguous()
        else:
            self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            self._features = np.concatenate([self._host_features(list(self.sal_dict.values())), lab]) if self._n_unl else lab
            self._feat = torch.from_numpy(np.ascontiguousarray(self._features, dtype=np.float32)).to(self._device)
        from .. import pool

        pad = pool.auto_pad(self._feat.shape[1]) if self._feat.shape[0] else 0
        if pad:  # zero columns: bit-identical distances, and the tensor-core screen applies (pool.auto_pad)
            self._feat = pool.pad_features(self._feat, pad)
        self._sharded = sharded
        self._norms = None
        self._min_dist = None

    # reference :35-47: rows = unlabeled poses (dict order) then labeled; root-relative x.., y.., z.. per row
    def _host_features(self, poses):
        if not poses:
            return np.zeros((0, 0))
        try:
            p = np.asarray(poses, dtype=np.float64)  # [n, J, >= 3] when every pose has the same shape
            if p.ndim != 3:
                raise ValueError
        except ValueError:
            rows = []
            for pose in poses:
                q = np.array(pose).transpose([1, 0])[0:3, :]
                rows.append((q - q[:, self._root:self._root + 1]).flatten())
            return np.stack(rows)
        p = p.transpose(0, 2, 1)[:, 0:3, :]
        return (p - p[:, :, self._root:self._root + 1]).reshape(p.shape[0], -1)

    def _labeled_features(self):
        lab = self._host_features(list(self.al_dict.values()))
        if lab.shape[0] == 0 and self._n_unl:
            J = self._column.device_values.shape[1] if self._column is not None else len(next(iter(self.sal_dict.values())))
            lab = np.zeros((0, 3 * J))
        return lab

    def _compute_stacked_features(self, root_idx):
        """Reference :35-47 (host, float64)."""
        poses = list(self.sal_dict.values()) + list(self.al_dict.values())
        return self._host_features(poses)

    @property
    def features(self):
        if self._features is None:
            self._features = self._compute_stacked_features(self._root)
        return self._features

    @property
    def sal_keys(self):
        if self._sal_keys is None:
            self._sal_keys = list(self.sal_dict.keys())
        return self._sal_keys

    def _keys_at(self, rows):
        if self._column is not None:
            return self._column.table.guid_at(rows)
        keys = self.sal_keys
        return [keys[i] for i in rows]

    @property
    def min_distances(self):
        if self._min_dist is None:
            return None
        return self._min_dist.cpu().numpy().astype(np.float64).reshape(-1, 1)

    def update_distances(self, cluster_centers, only_new=True, reset_dist=False):
        """Reference :49-69.  (With several centres and an existing min_distances the reference broadcasts to an
        [n, c] matrix; here the minimum over all given centres is folded into the [n, 1] vector.)"""
        from .. import pool

        if reset_dist:
            self._min_dist = None
        if only_new:
            cluster_centers = [d for d in cluster_centers if d not in self.already_selected]
        if cluster_centers:
            if self._norms is None:
                self._norms = ops.kcenter_norms(self._feat)
            if self._min_dist is None:
                self._min_dist = torch.full((self.n_obs,), float("inf"), dtype=torch.float32, device=self._device)
            idx = torch.as_tensor([int(c) for c in cluster_centers], dtype=torch.int64, device=self._device)
            pool.kcenter_fold_centres([self._state()], self._feat[idx].contiguous(), self._norms[idx].contiguous())

    def _state(self):
        return {"feat": self._feat, "norms": self._norms, "min": self._min_dist, "off": 0}

    def _use_shards(self):
        import torch.distributed as dist

        if self._sharded is not None:
            return bool(self._sharded) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        return (self._column is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                and self._n_unl >= 4096 * dist.get_world_size())

    def select_batch(self, N, **kwargs):
        """Reference :71-95: fold in the labeled set, then N times {argmax, assert, update} -- executed in exact
        rounds on the device (csrc/kcenter.cu)."""
        from .. import pool

        already_selected = self.al_indices
        fresh = self._min_dist is None and len(already_selected) > 0 and not self.already_selected
        if fresh and self._use_shards():
            import torch.distributed as dist

            lo, hi = pool.shard_range(self._n_unl, dist.get_world_size(), dist.get_rank())
            sel, _ = pool.kcenter_greedy_sharded([(self._feat[lo:hi], lo)], self._feat[self._n_unl:], int(N), pad_to=None)
            new_batch = [int(i) for i in sel.cpu().tolist()]
        elif fresh:
            # whole selection in one C call
            sel, self._min_dist = ops.kcenter_greedy(self._feat, self._n_unl, N)
            new_batch = [int(i) for i in sel.cpu().tolist()]
        else:
            self.update_distances(already_selected, only_new=True, reset_dist=False)
            if self._min_dist is None:  # no labeled centre at all: np.argmax(None) in the reference, undefined
                raise ValueError("CoreSet.select_batch needs at least one labeled pose")
            sel = pool.kcenter_rounds([self._state()], int(N))
            new_batch = [int(i) for i in sel.cpu().tolist()]
        lo_lab = self._n_unl
        for ind in new_batch:
            assert ind < lo_lab  # reference :91 ``assert ind not in already_selected`` (the labeled rows)
        self.already_selected = already_selected
        return self._keys_at(new_batch)
