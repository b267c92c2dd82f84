"""Generates tests/golden/flow_*.npz by running the UNMODIFIED reference strategy.py (dev container only):
ActiveLearningStrategy._sal_pseudo_labeling (strategy.py:915-1002) -> _compute_sal_dict (:1004-1147) with its own
DataLoader + DistributedSampler (:747-760), its per-frame triangulation() calls and its 8 all_gathers per frame, on a
small synthetic pool, under torch.distributed with the gloo backend at world size 1 and 2.

TEST INFRASTRUCTURE.  Usage:  python oracle/make_golden_flow.py   (outputs are committed; re-run only when cases change).

What makes the reference runnable here without touching it:
  * stand-ins for the modules that are not installed and not on the arithmetic path (colorlog, kornia, matplotlib, iopath,
    yacs, tensorboard; skimage's peak finder is not called by these strategies) -- oracle/ref_import.py;
  * ``torch.Tensor.cuda`` patched to the identity (no GPU in the dev container; the arithmetic is device independent);
  * the pose estimator is ``torch.nn.Identity()`` and the dataset's "images" ARE the heat maps -- the forward is outside
    the path under test;
  * a dataset object with the methods of dataset/dataset.py's ActiveLearningDataset that the path calls (:47-74, 98-110),
    restated literally.

Every fixture stores the pool's generator arguments (the heat maps are re-rendered from them by the tests) and what the
reference left behind on rank 0: the five dicts of sal_dict as (ordered guid list, value arrays), al_guids, sal_guids.
"""
import json
import os
import random
import socket
import sys
from collections import OrderedDict
from types import SimpleNamespace as NS

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multi_view_active_learning_b200 import synthetic as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
POOL = dict(n=23, n_labeled=6, V=5, J=19, seed=5, valid_prob=0.95, p_outlier=0.04, noise=0.05, batch=3)


class FlowDataset(torch.utils.data.Dataset):
    """dataset/dataset.py:47-74, 98-110 over a synthetic pool (shared by the generator and the tests)."""

    def __init__(self, n, n_labeled, V, J, seed, valid_prob, p_outlier, noise, **_):
        pool = S.make_pool(n + n_labeled, V, J, seed=seed, valid_prob=valid_prob, p_outlier=p_outlier)
        hm = S.render_heatmaps(pool["centres"], noise=noise, seed=seed + 1)
        gt = np.concatenate([pool["X"].transpose(0, 2, 1), np.ones((n + n_labeled, 1, J))], axis=1)  # [4, J]
        frames = [{"images": torch.from_numpy(hm[i]), "proj_matrices": torch.from_numpy(pool["P"][i]),
                   "joint_valid": torch.from_numpy(pool["valid"][i].astype(np.float32)),
                   "3d_keypoints": torch.from_numpy(gt[i].astype(np.float32)), "pose": 160422 + i % 3, "frame_id": 100 + i}
                  for i in range(n + n_labeled)]
        self.unlabeled_data = OrderedDict(("%d-%d" % (f["pose"], f["frame_id"]), f) for f in frames[:n])
        self.labeled_data = [dict(f, **{"3d_keypoints": f["3d_keypoints"].numpy()}) for f in frames[n:]]
        self.pseudo_label_guids, self.pseudo_labeled_data, self.data = [], [], []
        self.hm, self.pool = hm, pool

    def get_al_dict_for_coreset(self):  # :47-51
        return {idx: np.array(self.labeled_data[idx]["3d_keypoints"]).transpose([1, 0]) for idx in range(len(self.labeled_data))}

    def label_by_frame_guids(self, guids):  # :61-64
        for guid in guids:
            self.labeled_data.append(self.unlabeled_data[guid])
            del self.unlabeled_data[guid]

    def pseudo_label_by_frame_guids(self, guids, pseudo_labels):  # :66-74
        self.pseudo_label_guids = guids
        self.pseudo_labeled_data = list()
        for guid in guids:
            frame = self.unlabeled_data[guid].copy()
            frame["pseudo_3d_keypoints"] = np.array(pseudo_labels[guid]).transpose([1, 0])
            self.pseudo_labeled_data.append(frame)

    def resample_unlabeled_data(self):  # :98-102
        self.data = [self.unlabeled_data[guid] for guid in self.unlabeled_data]

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return self.data[idx].copy()


def flow_cfg(strategy, expr, world, cluster_file="", hp="AVG", batch=3):
    return NS(EXPR_TYPE=expr, RANDOM_SEED=1307, NUM_GPUS=world, DATA=NS(NUM_JOINTS=19, TYPE="panoptic"),
              POSE_ESTIMATOR=NS(STRIDE=4), SAL=NS(INLIER_THRESHOLD=2, CLUSTER_FILE_PATH=cluster_file, NUM_CLUSTERS=3),
              AL=NS(STRATEGY=strategy, USE_SOFTARGMAX=False, USE_REPROJECTION_XE=False, REPROJECTION_SIGMA=1.0, HP_CONFIG=hp,
                    MPE_CONFIG=hp, BSB_CONFIG=hp, INFERENCE=NS(BATCH_SIZE=batch, NUM_WORKERS=0)))


CASES = [  # name, strategy, EXPR_TYPE, clustered, HP config, al_num_frames, pseudo_num_frames
    ("triangulation_al", "TRIANGULATION", "AL", False, "AVG", 5, 0),
    ("hp_avg_al", "HP", "AL", False, "AVG", 5, 0),
    ("hp_std_al", "HP", "AL", False, "STD", 5, 0),
    ("coreset_al", "CORESET", "AL", False, "AVG", 5, 0),
    ("triangulation_sal", "TRIANGULATION", "SAL", False, "AVG", 4, 3),
    ("triangulation_sal_clustered", "TRIANGULATION", "SAL", True, "AVG", 4, 6),
]


def cluster_file_payload(J=19):
    rng = np.random.default_rng(11)
    return {"g%d" % i: (rng.normal(size=(4, J)) * 300).tolist() for i in range(40)}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmpdir, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import traceback

        from oracle.make_golden import load_strategy
        from oracle.ref_import import load_reference

        load_reference()
        st_mod = load_strategy()
        sys.modules["iopath.common.file_io"].PathManager.open = staticmethod(lambda path, mode="r": open(path, mode))
        st_mod.PathManager.open = staticmethod(lambda path, mode="r": open(path, mode))
        torch.Tensor.cuda = lambda self, *a, **k: self
        results = {}
        cluster_path = os.path.join(tmpdir, "clusters.json")
        for name, strategy, expr, clustered, hp, n_al, n_pseudo in CASES:
            cfg = flow_cfg(strategy, expr, world, cluster_path if clustered else "", hp, POOL["batch"])
            st = st_mod.ActiveLearningStrategy(cfg)
            ds = FlowDataset(**POOL)
            random.seed(99)  # the non-clustered SAL branch draws random.sample (strategy.py:993-995)
            _, al_guids, sal_guids, sal_dict = st._sal_pseudo_labeling(ds, n_al, n_pseudo, torch.nn.Identity())
            results[name] = (al_guids, sal_guids, sal_dict, [d["pseudo_3d_keypoints"].tolist() for d in ds.pseudo_labeled_data])
        out_q.put((rank, results))
    except Exception:
        out_q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def run_world(world, tmpdir):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, tmpdir, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r, v in got.items():
        if isinstance(v, str):
            raise RuntimeError("rank %d failed:\n%s" % (r, v))
    return got


def main():
    import tempfile

    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "clusters.json"), "w") as f:
            json.dump(cluster_file_payload(), f)
        for world in (1, 2):
            got = run_world(world, tmp)
            for name, *_ in CASES:
                per_rank = [got[r][name] for r in range(world)]
                for other in per_rank[1:]:  # every rank ends with the same selection and the same dicts
                    assert other[0] == per_rank[0][0] and other[1] == per_rank[0][1]
                    assert json.dumps(other[2]) == json.dumps(per_rank[0][2])
                al_guids, sal_guids, sal_dict, pseudo = per_rank[0]
                guids = list(sal_dict["al_metric"])
                assert all(list(sal_dict[k]) == guids for k in sal_dict)
                np.savez_compressed(
                    os.path.join(OUT, "flow_%s_w%d.npz" % (name, world)), pool=json.dumps(POOL), world=world,
                    guids=np.array(guids), al_metric=np.array(list(sal_dict["al_metric"].values()), dtype=np.float64),
                    sal_metric=np.array(list(sal_dict["sal_metric"].values()), dtype=np.float64),
                    inlier_count=np.array(list(sal_dict["inlier_count"].values()), dtype=np.float64),
                    mkpe=np.array(list(sal_dict["mkpe"].values()), dtype=np.float64),
                    pred_3d_keypoints=np.array(list(sal_dict["pred_3d_keypoints"].values()), dtype=np.float64),
                    al_guids=np.array(al_guids), sal_guids=np.array(sal_guids, dtype="<U32"),
                    pseudo_3d_keypoints=np.array(pseudo, dtype=np.float64).reshape(len(pseudo), 3, -1) if pseudo else np.zeros((0, 3, 19)))
                print("flow_%s_w%d" % (name, world), len(guids), "guids; al", al_guids[:3], "sal", sal_guids[:3])


if __name__ == "__main__":
    main()
