"""Generates tests/golden/*.npz by running the UNMODIFIED reference from /root/reference (dev container only).

TEST INFRASTRUCTURE.  Usage:  python oracle/make_golden.py   (re-run only when cases change; outputs are
committed).  Each fixture stores the inputs (small: key-points, projection matrices, sparse heatmaps) and the
outputs the reference produced for them, so that tests on the GPU box -- which has no /root/reference -- can
pin both the oracle and the CUDA path against the real thing.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multi_view_active_learning_b200 import synthetic as S  # noqa: E402
from oracle import triangulation_oracle as O  # noqa: E402
from oracle.ref_import import _stub, load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_strategy():
    for name in ["iopath", "iopath.common", "iopath.common.file_io", "skimage", "skimage.feature", "yacs",
                 "yacs.config", "torch.utils.tensorboard"]:
        _stub(name)
    sys.modules["torch.utils.tensorboard"].summary_writer = None
    sys.modules["iopath.common.file_io"].PathManager = type("PathManager", (), {})
    sys.modules["skimage.feature"].peak_local_max = lambda *a, **k: None
    import importlib

    return importlib.import_module("strategy")


def run_reference_triangulation(tri, heatmaps, P, stride, valid, pair_seed, **kw):
    """Calls the reference's triangulation() frame by frame, exactly like strategy.py:1036-1045."""
    N, V, J = heatmaps.shape[:3]
    sched = [(n, j) for n in range(N) for j in range(J) if valid[n, j]]
    tri.random = O.DeterministicShuffle(pair_seed, sched)
    res = [tri.triangulation(torch.from_numpy(heatmaps[n]), torch.from_numpy(P[n]), stride,
                             torch.from_numpy(valid[n]), **kw) for n in range(N)]
    return {
        "keypoints_3d": np.stack([r["keypoints_3d"] for r in res]),
        "keypoints_2d": np.stack([r["keypoints_2d"] for r in res]),
        "metric": np.array([r["metric"] for r in res]),
        "inlier_count": np.array([r["inlier_count"] for r in res]),
    }


def case_reference_unit_test(tri):
    """Inputs of the reference's tests/test_triangulation.py:15-71 (the test itself asserts shapes only)."""
    sys.path.insert(0, os.path.join("/root/reference", "tests"))
    import importlib
    import unittest

    mod = importlib.import_module("test_triangulation")
    captured = {}
    real = tri.triangulation

    def spy(heatmaps, proj, stride, valid, *a, **k):
        captured.update(heatmaps=heatmaps.numpy().copy(), P=proj.numpy().copy(), stride=stride,
                        valid=valid.numpy().copy())
        out = real(heatmaps, proj, stride, valid, *a, **k)
        captured.update(out=out)
        return out

    mod.triangulation = spy
    with contextlib.redirect_stdout(io.StringIO()):
        r = unittest.TextTestRunner(stream=io.StringIO()).run(
            unittest.defaultTestLoader.loadTestsFromTestCase(mod.TestTriangulation))
    assert r.wasSuccessful()
    o = captured["out"]
    np.savez_compressed(os.path.join(OUT, "ref_unit_triangulation.npz"),
                        P=captured["P"], stride=captured["stride"], valid=captured["valid"],
                        bump_rc=np.array([[11, 11, 1.0], [10, 11, 0.5], [11, 10, 0.5], [11, 12, 0.5], [12, 11, 0.5],
                                          [12, 12, 0.3], [10, 10, 0.3], [10, 12, 0.3], [12, 10, 0.3]]),
                        keypoints_3d=o["keypoints_3d"], keypoints_2d=o["keypoints_2d"], metric=o["metric"],
                        inlier_count=o["inlier_count"])
    print("ref_unit_triangulation", o["keypoints_3d"][0], float(o["metric"]), int(o["inlier_count"]))


def case_pool(tri, name, N, V, J, seed, valid_prob, pair_seed, p_outlier=0.12):
    pool = S.make_pool(N, V, J, seed=seed, valid_prob=valid_prob, p_outlier=p_outlier)
    # integer bump centres -> one-hot heatmaps the test can rebuild from keypoints_2d alone
    kp = (np.round(pool["centres"]).astype(np.int64)) * S.STRIDE
    hm = S.onehot_heatmaps(kp)
    out = run_reference_triangulation(tri, hm, pool["P"], S.STRIDE, pool["valid"], pair_seed)
    kp_masked = np.where(pool["valid"][:, None, :, None], kp, 0)
    assert np.array_equal(out["keypoints_2d"], kp_masked)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), P=pool["P"], valid=pool["valid"], stride=S.STRIDE,
                        pair_seed=pair_seed, keypoints_2d_unmasked=kp, **out)
    print(name, "metric[:3]", out["metric"][:3], "inlier_count", out["inlier_count"][:8])


def case_huber(tri):
    """triangulation(..., direct_optimization=True): the Huber refinement of utils/triangulation.py:319-336 run by
    the reference itself (scipy least_squares) on a pool with outlier views, so that many residuals sit in the
    linear zone of the loss."""
    N, V, J = 6, 8, 19
    pool = S.make_pool(N, V, J, seed=301, valid_prob=0.9, p_outlier=0.15)
    rng = np.random.default_rng(302)
    kp = (np.round(pool["centres"] + rng.normal(scale=0.6, size=pool["centres"].shape)).astype(np.int64)).clip(0, 63) * S.STRIDE
    hm = S.onehot_heatmaps(kp)
    out = run_reference_triangulation(tri, hm, pool["P"], S.STRIDE, pool["valid"], 0, direct_optimization=True)
    base = run_reference_triangulation(tri, hm, pool["P"], S.STRIDE, pool["valid"], 0)
    np.savez_compressed(os.path.join(OUT, "huber_v8_j19.npz"), P=pool["P"], valid=pool["valid"], stride=S.STRIDE,
                        keypoints_2d_unmasked=kp, keypoints_3d_dlt=base["keypoints_3d"], metric_dlt=base["metric"], **out)
    print("huber: max shift from DLT (mm)", np.abs(out["keypoints_3d"] - base["keypoints_3d"]).max(), "metric", out["metric"][:3])


def case_decode(ev):
    rng = np.random.default_rng(11)
    V, J = 3, 5
    hm = rng.normal(size=(V, J, 64, 64)).astype(np.float32)
    hm[0, 0] = 0.0  # all equal -> index 0
    hm[0, 1, 5, 7] = hm[0, 1, 40, 2] = 9.0  # exact tie -> first
    hm[0, 2, 63, 63] = 50.0  # last element
    hm[1, 0, 20, 30] = np.nan  # NaN is the maximum
    hm[1, 1, 3, 3] = np.nan
    hm[1, 1, 2, 60] = np.nan  # first NaN wins
    hm[1, 2, 9, 9] = np.inf
    hm[1, 3] = -np.inf
    hm[1, 4, 0, 0] = -0.0
    hm[2, 0] = np.abs(hm[2, 0]) * -1.0
    hm[2, 0, 17, 5] = 0.0
    hm[2, 0, 33, 8] = -0.0  # -0.0 == 0.0 -> earlier index
    valid = np.array([1, 1, 1, 0, 1], dtype=bool)
    out = ev.get_scaled_pred_corrdinates(torch.from_numpy(hm), 4, J, torch.from_numpy(valid))
    out_all = ev.get_scaled_pred_corrdinates(torch.from_numpy(hm), 4, J, torch.ones(J).bool())
    boxes = torch.tensor([[0.0, 0.0, 128.0, 128.0], [10.0, 20.0, 210.0, 220.0], [0.0, 0.0, 64.0, 64.0]])
    pc = ev.get_pred_coordinates(torch.from_numpy(hm), boxes, J)
    pc = np.array([[[float(c) for c in k] for k in b] for b in pc], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "decode_edge_cases.npz"), heatmaps=hm, valid=valid, stride=4,
                        scaled=out, scaled_all_valid=out_all, boxes=boxes.numpy(), pred_coordinates=pc)
    print("decode_edge_cases", out.dtype, out[0, :3].tolist(), out[1, :3].tolist())


def case_hp(st):
    rng = np.random.default_rng(5)
    V, J = 4, 6
    centres = rng.uniform(4, 60, size=(V, J, 2)).astype(np.float32)
    hm = S.render_heatmaps(centres, noise=0.05, seed=6) * np.float32(4.0)
    valid = np.array([1, 1, 0, 1, 1, 1], dtype=np.float32)

    class NS:
        pass

    self_ = NS()
    self_.al_cfg = NS()
    self_.al_cfg.AL = NS()
    import warnings

    res = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for cfg in ("AVG", "STD"):
            self_.al_cfg.AL.HP_CONFIG = cfg
            res[cfg] = st.ActiveLearningStrategy._compute_hp(self_, torch.from_numpy(hm), torch.from_numpy(valid))
        per_map = np.array([[float(1 - torch.max(torch.nn.functional.softmax(torch.from_numpy(hm[v, k]), dim=1)))
                             for k in range(J)] for v in range(V)])
    np.savez_compressed(os.path.join(OUT, "hp_scores.npz"), heatmaps=hm, valid=valid, hp_avg=res["AVG"],
                        hp_std=res["STD"], hp_per_map=per_map)
    print("hp_scores", res)


def case_peaks(st):
    """_compute_mpe / _compute_mpes / _compute_bsb of the unmodified reference (strategy.py:1149-1176, 1195-1215) with the
    one call it makes into scikit-image -- peak_local_max(map, min_distance=2, indices=True[, num_peaks=2]), skimage <= 0.19,
    not installable here -- served by the restatement in oracle/scores_oracle.py.  This pins everything the reference
    itself does around the peaks (which values are read, the float32 softmax over them, math.log, the order of the two
    best peaks, AVG / STD, the joint_valid skip); the peak finder alone stays "parity unpinned"."""
    from oracle import scores_oracle as SO

    rng = np.random.default_rng(15)
    V, J = 2, 4
    centres = rng.uniform(6, 58, size=(V, J, 2)).astype(np.float32)
    hm = S.render_heatmaps(centres, noise=0.05, seed=16) * np.float32(3.0)
    second = S.render_heatmaps(rng.uniform(6, 58, size=(V, J, 2)).astype(np.float32), noise=0.0, seed=0) * np.float32(1.5)
    hm = (hm + second).astype(np.float32)  # two bumps per map + noise: several local peaks everywhere
    valid = np.array([1, 0, 1, 1], dtype=np.float32)

    def served_peaks(image, min_distance=1, indices=True, num_peaks=np.inf, **kw):
        assert indices is True and not kw
        return SO.peak_local_max(image, min_distance=min_distance, num_peaks=num_peaks)

    st.peak_local_max = served_peaks

    class NS:
        pass

    self_ = NS()
    self_.al_cfg = NS()
    self_.al_cfg.AL = NS()
    self_._compute_mpes = lambda h, v: st.ActiveLearningStrategy._compute_mpes(self_, h, v)
    import warnings

    res = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ents = st.ActiveLearningStrategy._compute_mpes(self_, torch.from_numpy(hm), torch.from_numpy(valid))
        for cfg in ("AVG", "STD"):
            self_.al_cfg.AL.MPE_CONFIG = self_.al_cfg.AL.BSB_CONFIG = cfg
            res["mpe_" + cfg] = st.ActiveLearningStrategy._compute_mpe(self_, torch.from_numpy(hm), torch.from_numpy(valid))
            res["bsb_" + cfg] = st.ActiveLearningStrategy._compute_bsb(self_, torch.from_numpy(hm), torch.from_numpy(valid))
    np.savez_compressed(os.path.join(OUT, "peak_scores.npz"), heatmaps=hm, valid=valid,
                        mpe_per_map=np.array(ents, dtype=np.float64), numpy_version=np.__version__,
                        **{k: np.float64(v) for k, v in res.items()})
    print("peak_scores", res, "n maps", len(ents))


def case_softargmax(tri):
    """triangulation(..., use_soft_argmax=True) of the unmodified reference (utils/triangulation.py:191-200).  kornia is
    not installable here, but `transformers` ships a verbatim copy of kornia's `spatial_expectation2d` and
    `create_meshgrid` ("Copied from kornia library: kornia/geometry/subpix/dsnt.py:76", transformers/models/efficientloftr/
    modeling_efficientloftr.py); kornia.spatial_soft_argmax2d(x, temperature=1, normalized_coordinates) is
    spatial_expectation2d(spatial_softmax2d(x, temperature), normalized_coordinates) with spatial_softmax2d =
    F.softmax(x.view(B, C, -1) * temperature, dim=-1).view_as(x) (dsnt.py), served to the reference's `kornia` name that way.
    Pins the key-points (float32), the 3-D joints and the metric of the soft-arg-max path."""
    import torch.nn.functional as F
    from transformers.models.efficientloftr.modeling_efficientloftr import spatial_expectation2d

    def spatial_soft_argmax2d(input, temperature=torch.tensor(1.0), normalized_coordinates=True):
        b, c, h, w = input.shape
        soft = F.softmax(input.view(b, c, -1) * temperature.to(input.dtype), dim=-1).view(b, c, h, w)
        return spatial_expectation2d(soft, normalized_coordinates)

    tri.kornia.spatial_soft_argmax2d = spatial_soft_argmax2d
    N, V, J = 4, 5, 6
    pool = S.make_pool(N, V, J, seed=401, valid_prob=0.85, p_outlier=0.1)
    hm = S.render_heatmaps(pool["centres"], noise=0.05, seed=402) * np.float32(8.0)
    out = run_reference_triangulation(tri, hm, pool["P"], S.STRIDE, pool["valid"], 0, use_soft_argmax=True)
    assert out["keypoints_2d"].dtype == np.float32
    np.savez_compressed(os.path.join(OUT, "softargmax_v5_j6.npz"), P=pool["P"], valid=pool["valid"], stride=S.STRIDE,
                        centres=pool["centres"], heatmap_seed=402, noise=0.05, gain=8.0, **out)
    print("softargmax kp[0,0,:2]", out["keypoints_2d"][0, 0, :2], "metric", out["metric"])


def case_pred_coordinates_soft(ev):
    """get_pred_coordinates(..., use_softargmax=True) of the unmodified reference (utils/evaluation.py:37-43), kornia served
    by the same vendored expectation code as case_softargmax."""
    import torch.nn.functional as F
    from transformers.models.efficientloftr.modeling_efficientloftr import spatial_expectation2d

    def spatial_soft_argmax2d(input, temperature=torch.tensor(1.0), normalized_coordinates=True):
        b, c, h, w = input.shape
        soft = F.softmax(input.view(b, c, -1) * temperature.to(input.dtype), dim=-1).view(b, c, h, w)
        return spatial_expectation2d(soft, normalized_coordinates)

    ev.kornia.spatial_soft_argmax2d = spatial_soft_argmax2d
    B, K = 3, 5
    pool = S.make_pool(1, B, K, seed=411)
    hm = S.render_heatmaps(pool["centres"][0], noise=0.05, seed=412) * np.float32(6.0)  # [B, K, 64, 64]
    boxes = np.array([[10, 20, 266, 276], [0, 0, 128, 128], [5, 7, 325, 327]], dtype=np.float32)  # square boxes (:40)
    out = ev.get_pred_coordinates(torch.from_numpy(hm), torch.from_numpy(boxes), K, use_softargmax=True)
    np.savez_compressed(os.path.join(OUT, "pred_coordinates_soft.npz"), centres=pool["centres"][0], heatmap_seed=412, noise=0.05,
                        gain=6.0, boxes=boxes, coords=out.numpy())
    print("pred_coordinates_soft", out.numpy()[0, 0], out.dtype)


def case_coreset(cs):
    import uuid

    rng = np.random.default_rng(3)
    N, L, J, budget = 400, 25, 19, 40
    sal = {"%d-%d" % (i % 7, i): (rng.normal(size=(J, 3)) * 100).astype(np.float32).tolist() for i in range(N)}
    al = {i: rng.normal(size=(J, 4)) * 100 for i in range(L)}
    with contextlib.redirect_stdout(io.StringIO()):
        c = cs.CoreSet(sal, al, 2)
        picked = c.select_batch(budget)
    keys = list(sal)
    np.savez_compressed(os.path.join(OUT, "coreset_random.npz"), sal_poses=np.array(list(sal.values())),
                        al_poses=np.array(list(al.values())), sal_keys=np.array(keys), root=2, budget=budget,
                        features=c.features, picked=np.array([keys.index(k) for k in picked]),
                        min_distances=c.min_distances)
    # the reference's own tests/test_coreset.py:15-17 input: identical poses everywhere -> index 0 five times
    sal2 = {str(uuid.UUID(int=i)): [[0, 1, 2] for _ in range(19)] for i in range(20)}
    al2 = {i: [[0, 1, 2] for _ in range(19)] for i in range(5)}
    with contextlib.redirect_stdout(io.StringIO()):
        c2 = cs.CoreSet(sal2, al2, 2)
        picked2 = c2.select_batch(5)
    keys2 = list(sal2)
    np.savez_compressed(os.path.join(OUT, "ref_unit_coreset.npz"), picked=np.array([keys2.index(k) for k in picked2]),
                        n_sal=20, n_al=5, n_joints=19, root=2)
    print("coreset", [keys.index(k) for k in picked][:8], [keys2.index(k) for k in picked2])


def case_xe(tri):
    """_compute_xe (utils/triangulation.py:236-257).  The function moves its rendered maps with .cuda(); there is no
    GPU in the dev container, so Tensor.cuda is patched to the identity for this call (the arithmetic -- float64
    render, float32 prediction promoted by MSELoss -- is device independent)."""
    N, V, J, sigma = 6, 4, 5, 2.5
    pool = S.make_pool(N, V, J, seed=77, valid_prob=1.0)
    hm = S.render_heatmaps(pool["centres"], noise=0.05, seed=78)
    P = pool["P"].copy()
    P[:3, :, :2, :] /= S.STRIDE  # first three frames: projections land inside the 64 x 64 grid (non-trivial renders)
    kp3d = pool["X"].copy()
    kp3d[:, 1] = 0.0  # an "invalid joint": keypoints_3d row left at zero by triangulation()
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        xe = np.array([float(tri._compute_xe(kp3d[n], P[n], torch.from_numpy(hm[n]), sigma)) for n in range(N)])
    finally:
        torch.Tensor.cuda = real_cuda
    m, _ = O.compute_xe(kp3d, P, hm, sigma)
    np.testing.assert_allclose(m, xe, rtol=1e-12)
    np.savez_compressed(os.path.join(OUT, "xe_metric.npz"), P=P, keypoints_3d=kp3d, centres=pool["centres"], sigma=sigma,
                        noise=0.05, heatmap_seed=78, xe=xe)
    print("xe", xe)


def main():
    os.makedirs(OUT, exist_ok=True)
    tri, ev, cs = load_reference()
    st = load_strategy()
    case_reference_unit_test(tri)
    case_pool(tri, "pool_v5_j19", N=24, V=5, J=19, seed=101, valid_prob=1.0, pair_seed=0)
    case_pool(tri, "pool_v8_j19", N=24, V=8, J=19, seed=102, valid_prob=1.0, pair_seed=0)
    case_pool(tri, "pool_v20_j42", N=4, V=20, J=42, seed=103, valid_prob=0.8, pair_seed=1234)
    case_pool(tri, "pool_v31_j19", N=3, V=31, J=19, seed=104, valid_prob=1.0, pair_seed=99)
    case_pool(tri, "pool_v2_j3", N=8, V=2, J=3, seed=105, valid_prob=0.7, pair_seed=0, p_outlier=0.3)
    case_huber(tri)
    case_decode(ev)
    case_hp(st)
    case_peaks(st)
    case_softargmax(tri)
    case_pred_coordinates_soft(ev)
    case_coreset(cs)
    case_xe(tri)


if __name__ == "__main__":
    main()
