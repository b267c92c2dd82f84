"""Coreset k-center greedy selection with the reference's class interface (utils/coreset.py:13-95)."""
from collections import OrderedDict

import numpy as np
import torch

from .. import ops


class CoreSet:
    """Same constructor, attributes and methods as the reference.  ``features`` is the reference's float64 host
    array; the distances are evaluated on the GPU in float32 in the canonical summation order (see
    csrc/kcenter.cu), first-index tie-break like np.argmax."""

    def __init__(self, sal_dict, al_dict, joint_root_index, metric="euclidean", device=None):
        if metric != "euclidean":
            raise NotImplementedError("only the euclidean metric of the reference's call sites is built")
        self.sal_dict = OrderedDict(sal_dict)
        self.al_dict = OrderedDict(al_dict)
        self.features = self._compute_stacked_features(joint_root_index)
        self.sal_keys = list(self.sal_dict.keys())
        self.name = "kcenter"
        self.metric = metric
        self.max_distances = None
        self.n_obs = len(sal_dict) + len(al_dict)
        self.al_indices = list(range(len(sal_dict), len(sal_dict) + len(al_dict)))
        self.already_selected = []
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._feat = torch.from_numpy(np.ascontiguousarray(self.features, dtype=np.float32)).to(self._device)
        self._norms = None
        self._min_dist = None

    def _compute_stacked_features(self, root_idx):
        # reference :35-47: rows = unlabeled poses (dict order) then labeled; root-relative x.., y.., z.. per row
        poses = list(self.sal_dict.values()) + list(self.al_dict.values())
        rows = []
        for pose in poses:
            p = np.array(pose).transpose([1, 0])[0:3, :]
            rows.append((p - p[:, root_idx:root_idx + 1]).flatten())
        return np.stack(rows)

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

    def select_batch(self, N, **kwargs):
        """Reference :71-95: fold in the labeled set, then N times {argmax, assert, update} -- executed in exact
        rounds on the device (csrc/kcenter.cu)."""
        from .. import pool

        already_selected = self.al_indices
        if self._min_dist is None and len(already_selected) > 0 and not self.already_selected:
            # whole selection in one C call
            sel, self._min_dist = ops.kcenter_greedy(self._feat, len(self.sal_dict), N)
            new_batch = [int(i) for i in sel.cpu().tolist()]
        else:
            self.update_distances(already_selected, only_new=True, reset_dist=False)
            if self._min_dist is None:  # no labeled centre at all: np.argmax(None) in the reference, undefined
                raise ValueError("CoreSet.select_batch needs at least one labeled pose")
            sel = pool.kcenter_rounds([self._state()], int(N))
            new_batch = [int(i) for i in sel.cpu().tolist()]
        for ind in new_batch:
            assert ind not in already_selected
        self.already_selected = already_selected
        return [self.sal_keys[i] for i in new_batch]
