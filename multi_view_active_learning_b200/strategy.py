"""Scoring-and-selection half of the reference's ``ActiveLearningStrategy`` (strategy.py:54-135, 915-1215),
re-built around batched sm_100a kernels.

``ScoringSelectionMixin`` carries the hot-path methods with the reference's names and signatures so that it can be
mixed over the reference class and ``workflow.py`` runs unchanged; ``ActiveLearningStrategy`` is the
stand-alone form used by the tests.  Training, evaluation, checkpoints and TensorBoard (strategy.py:137-914) stay
with the reference -- they are outside SURVEY.md section 8.

What changes relative to the reference's ``_compute_sal_dict`` (strategy.py:1004-1147):
  * one batched kernel pass per data-loader batch instead of a Python loop over frames, joints and view pairs -- for the
    HP / MPE / BSB strategies the same pass also yields the per-map scores (mval_score_pool_scored), and nothing in the
    loop waits for the device;
  * no per-frame collectives: every rank accumulates its frames on the device and the 8 per-frame all_gathers
    (:1106-1114) become one all_gather per field at the end;
  * the five guid-keyed dicts are views (table.py) of one device-resident table whose rows are in the order the
    reference would have inserted them (for each local position, rank 0..G-1; a repeated guid keeps its first position
    and its last value) and whose values carry the reference's float32 / float64 roundings; Python
    objects are only made for what is looked at -- the selected guids, or everything when SAL-DICT-ITER-k is written.
"""
import json
import math
import os
import random
from collections import OrderedDict

import numpy as np
import torch

from . import ops
from .table import LazyColumn, PoolTable, SalDict, dumps
from .utils import coreset, triangulation


def _item32(x):
    return float(np.float32(x))


# builtin sum() over Python floats is Neumaier-compensated since Python 3.12 and a plain left-to-right sum before; the
# reference's ``sum(hps) / len(hps)`` (strategy.py:1188-1190) inherits whichever the interpreter does, and so do we
_SUM_IS_COMPENSATED = sum([1e16, 1.0, -1e16]) == 1.0


def _python_float_sums(xt):
    """xt float64 [m, n] (one COLUMN per sequence) -> float64 [n]: what ``sum(column.tolist())`` returns for every
    column, evaluated for all of them at once (m vector steps instead of n Python-level sums).  Mirrors builtin sum():
    start 0, left to right; on Python >= 3.12 with the Neumaier compensation term of Python/bltinmodule.c."""
    s = np.zeros(xt.shape[1], dtype=np.float64)
    c = np.zeros(xt.shape[1], dtype=np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        for row in xt:
            nxt = s + row
            if _SUM_IS_COMPENSATED:
                c += np.where(np.abs(s) >= np.abs(row), (s - nxt) + row, (row - nxt) + s)
            s = nxt
        if _SUM_IS_COMPENSATED:
            s = np.where((c != 0) & np.isfinite(c), s + c, s)
    return s


def _aggregate_map_scores(kind, config, per_map, valid):
    """per_map float32 [B, V, J], valid bool [B, J] -> the frame scores of strategy.py:1151-1158 / 1188-1193 / 1210-1215
    for all B frames, with the reference's own arithmetic (values taken view-major over the valid joints):
      HP       per-map values are Python floats (.item()): AVG = builtin sum() / len in double, STD = np.std of a float64
               array  -> float64 [B];
      MPE/BSB  per-map values are np.float32 scalars: AVG = builtin sum() = one float32 add after the other, / len in
               float32; STD = np.std of a float32 array (NumPy >= 2 promotion)  -> float32 [B].
    Frames are grouped by their NUMBER of valid joints (at most J + 1 groups): both the left-to-right sums and NumPy's
    pairwise reduction depend on the values and their count only, so a group is one dense [n, m] array, summed value by
    value (AVG) or reduced with np.std(axis=1), which runs the pairwise summation per row exactly like the 1-D call."""
    B, V, J = per_map.shape
    dtype = np.float64 if kind == "HP" else np.float32
    out = np.empty(B, dtype=dtype)
    counts = valid.sum(axis=1)
    for c in np.unique(counts).tolist():
        rows = slice(None) if bool((counts == c).all()) else np.nonzero(counts == c)[0]
        x, v = per_map[rows], valid[rows]
        n = x.shape[0]
        if c != J:  # compact every frame's valid entries (row-major = view-major, joints ascending)
            x = x[np.broadcast_to(v[:, None, :], x.shape)]
        x = x.reshape(n, V * c)
        m = x.shape[1]
        if config == "AVG":
            if m == 0:
                raise ZeroDivisionError("division by zero")  # sum([]) / len([]) in the reference
            xt = np.ascontiguousarray(x.T, dtype=dtype)
            if kind == "HP":
                out[rows] = _python_float_sums(xt) / m
            else:
                acc = np.zeros(n, dtype=np.float32)
                with np.errstate(invalid="ignore", over="ignore"):
                    for row in xt:
                        acc = acc + row
                out[rows] = acc / np.float32(m)
        else:
            out[rows] = np.std(x.astype(dtype, copy=False), axis=1)
    return out


class ScoringSelectionMixin:
    # ------------------------------------------------------------------------------------------ entry point
    def sample_next_batch(self, train_dataset, al_num_frames, sal_num_frames, pose_estimator, iteration, rank=-1):
        """Reference strategy.py:54-135 without the rank-0 JSON / TensorBoard side effects (those stay in the
        reference class; mixing this class over it keeps them via ``_after_sampling``)."""
        sal_guids, sal_dict = [], {}
        if iteration == 0:
            train_dataset, al_guids = self._random_sample_frames(train_dataset, al_num_frames)
        else:
            train_dataset, al_guids, sal_guids, sal_dict = self._sal_pseudo_labeling(
                train_dataset, al_num_frames, sal_num_frames, pose_estimator)
        hook = getattr(self, "_after_sampling", None)
        if hook is not None:
            hook(iteration, rank, al_guids, sal_guids, sal_dict)
        self.last_al_guids, self.last_sal_guids, self.last_sal_dict = al_guids, sal_guids, sal_dict
        return train_dataset

    def _random_sample_frames(self, train_dataset, num_frames, seed=None):
        """Reference strategy.py:868-878."""
        if seed is None:
            seed = self.al_cfg.RANDOM_SEED
        random.seed(seed)
        guids = random.sample(list(train_dataset.unlabeled_data.keys()), num_frames)
        train_dataset.label_by_frame_guids(guids)
        return train_dataset, guids

    # ------------------------------------------------------------------------------------------ selection
    def _sal_pseudo_labeling(self, train_dataset, al_num_frames, pseudo_num_frames, pose_estimator):
        """Reference strategy.py:915-1002.  ``sal_dict`` is the device-resident table of ``_compute_sal_dict`` (table.py):
        ranking, coreset features, the pseudo-label filter and the cluster ids are all computed from its CUDA columns and
        only the selected rows are turned into guid strings -- nothing here is O(pool) in Python."""
        cfg = self.al_cfg
        if cfg.AL.STRATEGY == "RANDOM" and cfg.EXPR_TYPE == "AL":
            train_dataset, al_guids = self._random_sample_frames(train_dataset, al_num_frames, seed=cfg.RANDOM_SEED)
            return train_dataset, al_guids, [], {}
        train_dataset.resample_unlabeled_data()
        data_loader = self._get_dataloader(train_dataset, cfg.AL.INFERENCE.BATCH_SIZE, cfg.AL.INFERENCE.NUM_WORKERS)
        sal_dict = self._compute_sal_dict(data_loader, pose_estimator)
        if cfg.AL.STRATEGY == "CORESET":
            cs = coreset.CoreSet(sal_dict["pred_3d_keypoints"], train_dataset.get_al_dict_for_coreset(),
                                 self.joint_root_index)
            al_guids = cs.select_batch(al_num_frames)
        else:
            al_guids = self._rank_nlargest(sal_dict["al_metric"], al_num_frames)
        train_dataset.label_by_frame_guids(al_guids)
        sal_sampled_guids = []
        if cfg.EXPR_TYPE == "SAL":
            clustered = cfg.SAL.CLUSTER_FILE_PATH != ""
            # reference :957-975: filter + ascending sort, on the device (mval_sal_rank).  Without clusters only the best
            # 2n candidates are ever looked at (:993-995); the cluster-balanced walk may need the whole order.
            sal_guids = self._sal_candidates(sal_dict, al_guids, train_dataset.pseudo_label_guids, cfg.SAL.INLIER_THRESHOLD,
                                             None if clustered else 2 * pseudo_num_frames)
            if clustered:
                # reference :976-992: walk the candidates in order, every cluster takes its first per_cluster_count
                cluster_ids = self._cluster_ids(sal_dict["pred_3d_keypoints"], sal_guids)
                counter = [0 for _ in range(cfg.SAL.NUM_CLUSTERS)]
                per_cluster_count = pseudo_num_frames // cfg.SAL.NUM_CLUSTERS
                for guid, cluster_id in zip(sal_guids, cluster_ids):
                    if counter[cluster_id] < per_cluster_count:
                        counter[cluster_id] += 1
                        sal_sampled_guids.append(guid)
            else:
                sal_sampled_guids = random.sample(sal_guids[:2 * pseudo_num_frames], pseudo_num_frames)
            pseudo_labels = sal_dict["pred_3d_keypoints"]
            if isinstance(pseudo_labels, LazyColumn):
                pseudo_labels.prefetch(sal_sampled_guids)  # one device gather for the rows the dataset reads (:998-1000)
            train_dataset.pseudo_label_by_frame_guids(sal_sampled_guids, pseudo_labels)
        return train_dataset, al_guids, sal_sampled_guids, sal_dict

    def _cluster_ids(self, pred_3d_keypoints, guids, margin_rtol=1e-9):
        """strategy.py:981-989 ``self.kmeans.predict([kp])[0]`` for every candidate at once (mval_kmeans_assign).  The
        device evaluates sklearn's predict rule in float64; a frame whose two best centres are closer than rounding
        could separate (relative margin <= margin_rtol; also NaN) is handed to sklearn itself, literally as the
        reference does, so the labels are the reference's."""
        if not guids:
            return []
        if isinstance(pred_3d_keypoints, LazyColumn):
            dev = pred_3d_keypoints.device_values
            rows = torch.as_tensor(pred_3d_keypoints.table.rows_of(guids), dtype=torch.int64, device=dev.device)
            pred = dev[rows]
        else:
            pred = torch.tensor([pred_3d_keypoints[g] for g in guids], dtype=torch.float32).cuda()
        centres = np.asarray(self.kmeans.cluster_centers_, dtype=np.float64)
        label, margin = ops.kmeans_assign(pred, centres, self.joint_root_index)
        label, margin = label.cpu().numpy(), margin.cpu().numpy()
        cmax = float(np.abs(centres).max())
        scale = cmax * max(cmax, 2.0 * float(pred.abs().max())) * centres.shape[1] + 1.0  # bound on |x||c| d
        close = np.nonzero(~(margin > margin_rtol * scale))[0].tolist()
        if close and isinstance(pred_3d_keypoints, LazyColumn):
            pred_3d_keypoints.prefetch([guids[i] for i in close])
        for i in close:
            kp = np.array(pred_3d_keypoints[guids[i]]).T
            kp = (kp[0:3, :] - kp[0:3, self.joint_root_index:self.joint_root_index + 1]).flatten()
            label[i] = self.kmeans.predict([kp])[0]
        return label.tolist()

    @staticmethod
    def _sal_candidates(sal_dict, al_guids, pseudo_label_guids, inlier_threshold, limit=None):
        """strategy.py:957-975: guids with a non-NaN sal_metric and inlier_count > threshold that were neither picked by
        the AL step nor pseudo-labelled before, by ascending sal_metric (ties in dict order); the first ``limit``."""
        table = getattr(sal_dict, "table", None)
        if table is not None:
            if table.n == 0:
                return []
            metric, inliers = table.values["sal_metric"], table.values["inlier_count"]
            excluded = torch.zeros(table.n, dtype=torch.uint8, device=metric.device)
            rows = table.rows_of(list(al_guids) + list(pseudo_label_guids), missing_ok=True)
            if rows:
                excluded[torch.as_tensor(rows, dtype=torch.int64, device=metric.device)] = 1
            idx = ops.sal_rank(metric, inliers, excluded, float(inlier_threshold), table.n if limit is None else int(limit))
            return table.guid_at(idx)
        keys = list(sal_dict["sal_metric"].keys())
        if not keys:
            return []
        taken = set(al_guids) | set(pseudo_label_guids)
        excluded = torch.tensor([k in taken for k in keys], dtype=torch.uint8).cuda()
        metric = torch.tensor(list(sal_dict["sal_metric"].values()), dtype=torch.float32).cuda()
        inliers = torch.tensor([sal_dict["inlier_count"][k] for k in keys], dtype=torch.float32).cuda()
        idx = ops.sal_rank(metric, inliers, excluded, float(inlier_threshold), len(keys) if limit is None else int(limit))
        return [keys[i] for i in idx.cpu().tolist()]

    @staticmethod
    def _rank_nlargest(al_metric, n):
        """strategy.py:932-949 (NaN filter + heapq.nlargest) as a device top-k: descending score, ties by dict
        insertion order.  A LazyColumn is ranked straight from its CUDA column."""
        if n <= 0 or len(al_metric) == 0:
            return []
        if isinstance(al_metric, LazyColumn):
            idx, _ = ops.topk_desc(al_metric.device_values, n)
            return al_metric.table.guid_at(idx)
        keys = list(al_metric.keys())
        scores = torch.tensor(list(al_metric.values()), dtype=torch.float64).cuda()
        idx, _ = ops.topk_desc(scores, n)
        return [keys[i] for i in idx.cpu().tolist()]

    # ------------------------------------------------------------------------------------------ scoring
    def _get_dataloader(self, dataset, batch_size, num_workers):
        """Reference strategy.py:747-760."""
        from torch.utils.data import DataLoader, DistributedSampler

        # pinned batches: the uploads of _device_batches are then truly asynchronous (the reference's loader is the same
        # DataLoader + DistributedSampler without pinning, strategy.py:753-759)
        return DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, sampler=DistributedSampler(dataset),
                          pin_memory=torch.cuda.is_available())

    @staticmethod
    def _device_batches(data_loader):
        """the reference uploads a batch with blocking .cuda() calls, runs the forward, then scores
        (strategy.py:1024-1035).  Here the tensors of batch k + 1 are uploaded on a COPY stream while the forward and the
        scoring launch of batch k run on the compute stream; the compute stream waits on the upload's event only.  Batches
        whose tensors already live on the device pass through untouched."""
        if not torch.cuda.is_available():
            yield from data_loader
            return
        copy_stream = None

        def upload(dp):
            nonlocal copy_stream
            if not isinstance(dp, dict) or not any(torch.is_tensor(v) and not v.is_cuda for v in dp.values()):
                return dp, None
            if copy_stream is None:
                copy_stream = torch.cuda.Stream()
            with torch.cuda.stream(copy_stream):
                out = {k: (v.cuda(non_blocking=True) if torch.is_tensor(v) and not v.is_cuda else v) for k, v in dp.items()}
                done = torch.cuda.Event()
                done.record(copy_stream)
            return out, done

        it = iter(data_loader)
        try:
            nxt = upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, done = nxt
            try:
                nxt = upload(next(it))  # enqueue the next upload before this batch's kernels are launched
            except StopIteration:
                nxt = None
            if done is not None:
                torch.cuda.current_stream().wait_event(done)
                for v in cur.values():
                    if torch.is_tensor(v) and v.is_cuda:
                        v.record_stream(torch.cuda.current_stream())
            yield cur

    @staticmethod
    def _compute_batch_heatmap(pose_estimator, data):
        """Reference strategy.py:771-782 (the forward stays in PyTorch/cuDNN)."""
        images = data["images"].cuda()
        return pose_estimator(images.reshape([-1, images.shape[2], images.shape[3], images.shape[4]]))

    def _batch_al_metric(self, B, tri):
        """What strategy.py:1072-1094 contributes for the B frames of one batch, WITHOUT synchronising with the device:
        ("al", float64 CUDA [B]) for the strategies whose metric is final per frame, or ("map", float32 CUDA [B, V, J])
        for HP / MPE / BSB, whose per-map scores (from the same fused pass as the triangulation) are reduced to frame
        scores once per pool in ``_pool_al_metric``."""
        strategy = self.al_cfg.AL.STRATEGY
        if strategy == "RANDOM":
            # torch.rand(1).cuda() per frame in the reference: same draws from the global generator, float32 values
            draws = [torch.rand(1) for _ in range(B)]
            return "al", (torch.cat(draws) if draws else torch.zeros(0)).double().cuda()
        if strategy == "TRIANGULATION":
            return "al", tri["metric"]
        if strategy == "CORESET":
            return "al", torch.zeros(B, dtype=torch.float64, device=tri["metric"].device)
        if strategy in ("HP", "MPE", "BSB"):
            return "map", tri["map_score"]
        raise NotImplementedError()

    def _pool_al_metric(self, per_map, valid):
        """HP / MPE / BSB: per-map scores float32 CUDA [N, V, J] + validity float CUDA [N, J] of this rank's frames -> (frame
        scores float64 CUDA [N], whether the reference's tensor is float64) -- strategy.py:1076-1090 for the whole pool."""
        cfg = self.al_cfg.AL
        config = {"HP": cfg.HP_CONFIG, "MPE": cfg.MPE_CONFIG, "BSB": cfg.BSB_CONFIG}[cfg.STRATEGY]
        # torch.tensor(...) of the reference: float64 only for HP's np.std of Python floats (:1081-1085)
        al_is_f64 = config == "STD" and cfg.STRATEGY == "HP"
        if per_map.is_cuda and config in ("AVG", "STD") and per_map.shape[2] <= 128:
            # the reference's summation rules evaluated per frame on the device (mval_aggregate_map_scores): the per-map
            # scores of a 125 k-frame shard are 76 MB that used to travel to the host for ~0.2 s of numpy
            return ops.aggregate_map_scores(per_map, valid, cfg.STRATEGY, config, _SUM_IS_COMPENSATED), al_is_f64
        vals = self._compute_map_score_batch(cfg.STRATEGY, config, per_map, valid, per_map)
        return torch.from_numpy(np.asarray(vals, dtype=np.float64)).to(per_map.device), al_is_f64

    @staticmethod
    def _compute_map_score_batch(kind, config, heatmaps, joint_valid, per_map=None):
        """strategy.py:1149-1215 for a batch: the per-map score (HP / MPE / BSB) on the device, then the reference's AVG /
        STD over (view, valid joint) for every frame at once (``_aggregate_map_scores``).
        per_map: float32 [B, V, J] scores that were already computed (ops.score_pool(..., map_score=kind))."""
        valid = (torch.as_tensor(joint_valid) != 0)
        if valid.dim() == 1:
            valid = valid.unsqueeze(0).expand(heatmaps.shape[0], -1)
        if per_map is None:
            per_map = ops.score_hp(heatmaps, valid) if kind == "HP" else ops.score_peaks(heatmaps, kind, valid)
        if config not in ("AVG", "STD"):
            if kind == "MPE":
                raise NotImplementedError("AL.MPE_CONFIG should be either AVG or STD.")  # reference :1157-1158
            return [None] * per_map.shape[0]  # the reference falls off the end of _compute_hp / _compute_bsb: None
        return _aggregate_map_scores(kind, config, per_map.cpu().numpy(), valid.cpu().numpy())

    def _one_frame(self, kind, config, heatmaps, joint_valid):
        hm = heatmaps if heatmaps.is_cuda else heatmaps.cuda()
        res = self._compute_map_score_batch(kind, config, hm.unsqueeze(0), torch.as_tensor(joint_valid).unsqueeze(0))[0]
        if kind == "BSB" and res is not None and np.isnan(res):
            raise IndexError("list index out of range")  # reference :1208 probs[1] with fewer than two peaks
        return res

    def _compute_hp(self, heatmaps, joint_valid):
        """Reference signature (strategy.py:1178): heatmaps [V, J, H, W] of one frame."""
        return self._one_frame("HP", self.al_cfg.AL.HP_CONFIG, heatmaps, joint_valid)

    def _compute_mpe(self, heatmaps, joint_valid):
        """Reference signature (strategy.py:1149)."""
        return self._one_frame("MPE", self.al_cfg.AL.MPE_CONFIG, heatmaps, joint_valid)

    def _compute_bsb(self, heatmaps, joint_valid):
        """Reference signature (strategy.py:1195)."""
        return self._one_frame("BSB", self.al_cfg.AL.BSB_CONFIG, heatmaps, joint_valid)

    def _compute_sal_dict(self, data_loader, pose_estimator):
        """Reference strategy.py:1004-1147, batched (see module docstring).  Nothing in the loop waits for the device:
        every batch enqueues its forward and ONE fused scoring launch and keeps the results as CUDA tensors, so the
        kernels of batch k run while the loader collates and uploads batch k + 1; the frame scores of HP / MPE / BSB and
        the MKPE are formed once per pool, the ranks exchange two packed tensors, and the result is the device-resident
        table of table.py viewed through the reference's five guid-keyed dicts."""
        cfg = self.al_cfg
        acc = {k: [] for k in ("sal", "inl", "al", "map", "pred", "gt", "valid", "pose", "frame")}
        for dp in self._device_batches(data_loader):
            with torch.no_grad():
                heatmaps = self._compute_batch_heatmap(pose_estimator, dp)
                _, kp, w, h = heatmaps.shape
                B = dp["proj_matrices"].shape[0]
                heatmaps = heatmaps.reshape([B, -1, kp, w, h]).float()
                joint_valid = dp["joint_valid"]
                pose = torch.as_tensor(dp["pose"]).reshape(-1).cuda(non_blocking=True).long()
                frame = torch.as_tensor(dp["frame_id"]).reshape(-1).cuda(non_blocking=True).long()
                n_views = heatmaps.shape[1]
                # more than 64 view pairs (utils/triangulation.py:279-282): the pair subsets are keyed by the frame's guid,
                # not by its position, so that the result does not depend on world size, rank or loader order
                keys = (pose << 32) + frame if n_views * (n_views - 1) // 2 > 64 else None
                tri = triangulation.triangulation_batch(
                    heatmaps, dp["proj_matrices"], cfg.POSE_ESTIMATOR.STRIDE, joint_valid,
                    use_soft_argmax=cfg.AL.USE_SOFTARGMAX, pair_seed=getattr(cfg, "RANDOM_SEED", 0),
                    frame_keys=keys, use_reprojection_xe=cfg.AL.USE_REPROJECTION_XE, sigma=cfg.AL.REPROJECTION_SIGMA,
                    map_score=cfg.AL.STRATEGY if cfg.AL.STRATEGY in ("HP", "MPE", "BSB") else None)
                field, value = self._batch_al_metric(B, tri)
                acc[field].append(value)
                acc["sal"].append(tri["metric"].float())  # torch.Tensor([metric]) -> float32 (:1061)
                acc["inl"].append(tri["inlier_count"].float())
                acc["pred"].append(tri["keypoints_3d"].float())  # torch.Tensor(keypoints_3d) -> float32 (:1046)
                acc["gt"].append(dp["3d_keypoints"].cuda(non_blocking=True).float())
                acc["valid"].append(torch.as_tensor(joint_valid).cuda(non_blocking=True).float())
                acc["pose"].append(pose)
                acc["frame"].append(frame)
        fields = {k: (torch.cat(v) if v else torch.zeros(0).cuda()) for k, v in acc.items()}
        al_is_f64 = cfg.AL.STRATEGY == "TRIANGULATION"
        per_map = fields.pop("map")
        if getattr(ops, "check_async", None) is not None and fields["sal"].is_cuda:
            ops.check_async()  # a tripped watchdog of any launch above raises here, before the scores are used
        if acc["sal"]:
            self._raise_reference_errors(fields["valid"], per_map if acc["map"] else None)
        if acc["map"]:
            fields["al"], al_is_f64 = self._pool_al_metric(per_map, fields["valid"])
        return self._build_table(fields, al_is_f64).as_sal_dict()

    def _raise_reference_errors(self, valid, per_map):
        """The batched pass cannot raise in the middle of a frame; the reference's per-frame errors are raised here, once
        per pool: a frame without a valid joint is np.min([]) in triangulation() (utils/triangulation.py:231), and a BSB map
        with fewer than two peaks is probs[1] of a one-element list (strategy.py:1208)."""
        flags = [(valid.sum(dim=1) == 0).any()]
        if per_map is not None and self.al_cfg.AL.STRATEGY == "BSB":
            flags.append((torch.isnan(per_map) & (valid != 0)[:, None, :]).any())
        flags = torch.stack(flags).cpu().tolist()
        if flags[0]:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        if len(flags) > 1 and flags[1]:
            raise IndexError("list index out of range")

    @staticmethod
    def _gather_rows(packs):
        """packs: tensors [n_local, ...] of this rank's rows (loader order).  Returns them for ALL ranks in the order the
        reference's per-frame all_gathers insert rows into its dicts: local position t of rank r lands at t * world + r
        (strategy.py:1106-1133).  One all_gather_into_tensor per pack (the reference: 8 list all_gathers per FRAME)."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return packs
        world = dist.get_world_size()
        n = int(packs[0].shape[0])
        dev = packs[0].device
        if dist.get_backend() == "gloo" and dev.type == "cuda":
            # gloo (CPU tests of the multi-rank flow on one GPU) exchanges host buffers; NCCL takes the CUDA tensors as is
            return [g.to(dev) for g in ScoringSelectionMixin._gather_rows([t.cpu() for t in packs])]
        sizes = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([n], dtype=torch.int64, device=dev))
        sizes = sizes.cpu().tolist()
        n_max = max(sizes)
        out = []
        for t in packs:
            t = t.contiguous()
            if n < n_max:  # a loader without the DistributedSampler's padding: pad, gather, drop the padding below
                pad = torch.zeros((n_max - n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                t = torch.cat([t, pad])
            g = torch.empty((world * n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(g, t)
            g = g.reshape((world, n_max) + tuple(t.shape[1:])).transpose(0, 1).reshape((n_max * world,) + tuple(t.shape[1:]))
            out.append(g)
        if min(sizes) < n_max:
            dev = out[0].device
            live = (torch.arange(n_max, device=dev)[:, None] < torch.tensor(sizes, device=dev)[None, :]).reshape(-1)
            out = [g[live] for g in out]
        return out

    @staticmethod
    def _gather_interleaved(fields):
        """Dict form of ``_gather_rows`` (kept for callers of the round-1 name)."""
        names = list(fields)
        return dict(zip(names, ScoringSelectionMixin._gather_rows([fields[k] for k in names])))

    def _build_table(self, f, al_is_f64):
        """Local rows -> MKPE (utils/evaluation.py:198-208, on the device, before the exchange: the reference gathers the
        ground truth of every frame to every rank only to compute it there) -> two packed exchanges -> dict-insertion
        semantics for repeated guids (mval_first_occurrence) -> PoolTable."""
        n, J = f["sal"].shape[0], (f["pred"].shape[1] if f["pred"].dim() == 3 else 0)
        dev = f["sal"].device
        if n:
            mkpe = ops.mkpe(f["pred"], f["gt"], f["valid"])
            al = f["al"].double()
            if not al_is_f64:
                al = al.float().double()  # torch.tensor(python float) of the reference is float32 (:1077-1090)
            pack32 = torch.cat([f["sal"].reshape(n, 1), f["inl"].reshape(n, 1), mkpe.reshape(n, 1), f["pred"].reshape(n, -1)], dim=1)
            pack64 = torch.stack([f["pose"], f["frame"], al.view(torch.int64)], dim=1)
        else:
            pack32 = torch.zeros((0, 3 + 3 * J), dtype=torch.float32, device=dev)
            pack64 = torch.zeros((0, 3), dtype=torch.int64, device=dev)
        pack32, pack64 = self._gather_rows([pack32, pack64])
        if pack64.shape[0]:
            pose, frame = pack64[:, 0].contiguous(), pack64[:, 1].contiguous()
            keep, src, unique = ops.first_occurrence(pose, frame)
            unique = int(unique.cpu().item())
            if unique < 0:
                src = self._first_occurrence_host(pose, frame)  # ids beyond 32 bits: dict semantics on the host
                rows = torch.nonzero(src >= 0).reshape(-1)
                pack32, pack64 = pack32[src[rows].long()], torch.cat([pack64[rows, :2], pack64[src[rows].long(), 2:]], dim=1)
            elif unique != pack64.shape[0]:
                rows = torch.nonzero(keep).reshape(-1)
                last = src[rows].long()
                pack32, pack64 = pack32[last], torch.cat([pack64[rows, :2], pack64[last, 2:]], dim=1)
        N = pack64.shape[0]
        return PoolTable(pose=pack64[:, 0].contiguous(), frame=pack64[:, 1].contiguous(),
                         al=pack64[:, 2].contiguous().view(torch.float64), sal=pack32[:, 0].contiguous(),
                         inl=pack32[:, 1].contiguous(), mkpe=pack32[:, 2].contiguous(),
                         pred=pack32[:, 3:].contiguous().reshape(N, J, 3))

    @staticmethod
    def _first_occurrence_host(pose, frame):
        first, last = {}, {}
        for i, key in enumerate(zip(pose.cpu().tolist(), frame.cpu().tolist())):
            first.setdefault(key, i)
            last[key] = i
        src = torch.full((pose.shape[0],), -1, dtype=torch.int32)
        for key, i in first.items():
            src[i] = last[key]
        return src.to(pose.device)


class SelectionLogMixin:
    """The files either side of the path: what ``sample_next_batch`` leaves on disk for rank 0
    (strategy.py:112-134) and how ``restore_dataset`` replays them (strategy.py:315-336).  Same names, same JSON
    payloads -- a run started with the reference can be resumed here and vice versa.  TensorBoard histograms
    (:82-108) stay with the reference class."""

    def _log_path(self, name):
        return os.path.join(self.al_cfg.LOG_DIR, self.al_cfg.EXPR_NAME, name)

    def _open(self, path, mode):
        mgr = getattr(self, "_pathmgr", None)
        if mgr is not None:
            return mgr.open(path, mode)
        if "w" in mode:
            os.makedirs(os.path.dirname(path), exist_ok=True)
        return open(path, mode)

    def _after_sampling(self, iteration, rank, al_guids, sal_guids, sal_dict):
        if rank != 0 or not hasattr(self.al_cfg, "LOG_DIR"):
            return
        if iteration != 0:
            if len(sal_guids) != 0:
                with self._open(self._log_path("SAL-GUID-ITER-%d" % iteration), "w") as f:
                    f.write(json.dumps(sal_guids))
            with self._open(self._log_path("SAL-DICT-ITER-%d" % iteration), "w") as f:
                f.write(dumps(sal_dict))  # json.dumps of the five dicts the LazyColumns stand for
        with self._open(self._log_path("SAMPLED-GUID-ITER-%d" % iteration), "w") as f:
            f.write(json.dumps(al_guids))

    def restore_dataset(self, train_dataset, iteration):
        """Reference strategy.py:315-336."""
        for i in range(0, iteration):
            with self._open(self._log_path("SAMPLED-GUID-ITER-%d" % i), "r") as f:
                guids = json.loads(f.readline())
            train_dataset.label_by_frame_guids(guids)
        if self.al_cfg.EXPR_TYPE == "SAL" and iteration > 1:
            with self._open(self._log_path("SAL-GUID-ITER-%d" % (iteration - 1)), "r") as f:
                train_dataset.pseudo_label_guids = json.loads(f.readline())
        return train_dataset


class ActiveLearningStrategy(SelectionLogMixin, ScoringSelectionMixin):
    """Stand-alone strategy object exposing only the scoring-and-selection path (reference strategy.py:28-52 for
    the constructor fields that path reads)."""

    def __init__(self, al_cfg):
        self.al_cfg = al_cfg
        self.num_joints = al_cfg.DATA.NUM_JOINTS
        self.joint_root_index = 2 if al_cfg.DATA.TYPE == "panoptic" else 21
        self.kmeans = None
        if al_cfg.EXPR_TYPE == "SAL" and getattr(al_cfg.SAL, "CLUSTER_FILE_PATH", "") != "":
            # reference :37-52: k-means over the root-relative poses of the cluster file
            from sklearn.cluster import KMeans

            with self._open(al_cfg.SAL.CLUSTER_FILE_PATH, "r") as f:
                clusters = json.load(f)
            kp_values = []
            for guid in clusters:
                kp = np.array(clusters[guid])
                kp_values.append((kp[0:3, :] - kp[0:3, self.joint_root_index:self.joint_root_index + 1]).flatten())
            self.kmeans = KMeans(al_cfg.SAL.NUM_CLUSTERS, random_state=al_cfg.RANDOM_SEED).fit(kp_values)
