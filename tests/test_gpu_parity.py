"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed golden
fixtures produced by the unmodified reference.  Run on the B200 box with  pytest -m gpu.

Tolerances (BASELINE.json north_star): arg-max key-points, inlier sets/counts, ranking order and coreset index
sets bit-exact; 3-D joints within 1e-3 relative / 1e-2 mm; reprojection errors within 1e-4 px.
"""
import itertools
import random

import numpy as np
import pytest
import torch

from multi_view_active_learning_b200 import synthetic as S
from oracle import coreset_oracle as CO
from oracle import scores_oracle as SO
from oracle import triangulation_oracle as O

pytestmark = pytest.mark.gpu

POOLS = ["pool_v5_j19", "pool_v8_j19", "pool_v20_j42", "pool_v31_j19", "pool_v2_j3"]
XYZ_ATOL_MM, XYZ_RTOL, REPROJ_ATOL_PX = 1e-2, 1e-3, 1e-4


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from multi_view_active_learning_b200 import ops as _ops

    return _ops


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _check_tri(out, ref, valid):
    """out: dict of numpy from the CUDA path; ref: oracle / golden dict."""
    assert np.array_equal(out["inlier_count"], ref["inlier_count"])
    if "inliers" in ref:
        assert np.array_equal(out["inliers"], ref["inliers"])
    if "inlier_mask" in ref and "inlier_mask" in out:
        assert np.array_equal(out["inlier_mask"].astype(np.uint32), ref["inlier_mask"])
    np.testing.assert_allclose(out["keypoints_3d"], ref["keypoints_3d"], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    # in practice the float64 Jacobi solve agrees far better than the contract; keep an eye on it
    assert np.abs(out["keypoints_3d"] - ref["keypoints_3d"]).max() < 1e-5
    np.testing.assert_allclose(out["metric"], ref["metric"], rtol=0, atol=REPROJ_ATOL_PX)
    assert (out["keypoints_3d"][~valid] == 0).all()


# ------------------------------------------------------------------------------------------------ decode
def test_decode_edge_cases_golden(ops, golden):
    g = golden("decode_edge_cases")
    hm = _cuda(g["heatmaps"])[None]
    xy = ops.decode_argmax(hm, int(g["stride"]), torch.from_numpy(g["valid"]))
    assert np.array_equal(xy[0].cpu().numpy(), g["scaled"])
    xy, peak = ops.decode_argmax(hm, int(g["stride"]), None, return_peak=True)
    assert np.array_equal(xy[0].cpu().numpy(), g["scaled_all_valid"])
    flat = g["heatmaps"].reshape(3, 5, -1)
    exp_peak = np.take_along_axis(flat, np.argmax(flat, -1)[..., None], -1)[..., 0]
    np.testing.assert_array_equal(peak[0].cpu().numpy(), exp_peak)


def test_decode_reference_api(ops, golden):
    from multi_view_active_learning_b200.utils import evaluation

    g = golden("decode_edge_cases")
    hm = torch.from_numpy(g["heatmaps"])
    out = evaluation.get_scaled_pred_corrdinates(hm, int(g["stride"]), 5, torch.from_numpy(g["valid"]))
    assert out.dtype == np.int64 and np.array_equal(out, g["scaled"])
    pc = evaluation.get_pred_coordinates(hm.cuda(), torch.from_numpy(g["boxes"]), 5)
    pc = np.array([[[float(c) for c in k] for k in b] for b in pc])
    np.testing.assert_allclose(pc, g["pred_coordinates"], rtol=1e-6)


def test_pred_coordinates_softargmax_branch_vs_reference_golden(ops, golden):
    """get_pred_coordinates(..., use_softargmax=True) (utils/evaluation.py:37-43) against the unmodified reference run with
    kornia's vendored expectation code: bbox-scaled float32 coordinates, tensor [B, K, 2]."""
    from multi_view_active_learning_b200.utils import evaluation

    g = golden("pred_coordinates_soft")
    hm = S.render_heatmaps(g["centres"], noise=float(g["noise"]), seed=int(g["heatmap_seed"])) * np.float32(g["gain"])
    out = evaluation.get_pred_coordinates(torch.from_numpy(hm).cuda(), torch.from_numpy(g["boxes"]), hm.shape[1], use_softargmax=True)
    assert torch.is_tensor(out) and tuple(out.shape) == g["coords"].shape and out.dtype == torch.float32
    scale = (g["boxes"][:, 3] - g["boxes"][:, 1]) / 64.0
    np.testing.assert_allclose(out.cpu().numpy(), g["coords"], rtol=0, atol=float(2e-4 * scale.max()))


@pytest.mark.parametrize("shape", [(3, 2, 5, 64, 64), (2, 3, 4, 48, 48), (1, 2, 3, 7, 9), (2, 2, 2, 32, 64)])
def test_decode_argmax_random_vs_oracle(ops, shape):
    rng = np.random.default_rng(sum(shape))
    hm = rng.normal(size=shape).astype(np.float32)
    hm[0, 0, 0].flat[[3, 77 % hm[0, 0, 0].size]] = 11.0  # a tie
    valid = rng.uniform(size=(shape[0], shape[2])) < 0.7
    got = ops.decode_argmax(_cuda(hm), 4, torch.from_numpy(valid)).cpu().numpy()
    assert np.array_equal(got, O.decode_argmax(hm, 4, valid))


def test_softargmax_vs_oracle(ops):
    pool = S.make_pool(3, 4, 6, seed=9)
    hm = S.render_heatmaps(pool["centres"], noise=0.05, seed=1) * np.float32(8.0)
    got = ops.decode_softargmax(_cuda(hm), 4).cpu().numpy()
    exp = O.decode_softargmax(hm, 4)
    # float32 expectation of a 0..252 px coordinate: a few 1e-5 px of rounding is inherent (kornia's own float32
    # output has the same spread); 2e-4 px bounds it
    np.testing.assert_allclose(got, exp, rtol=0, atol=2e-4)


def test_softargmax_path_vs_reference_golden(ops, golden):
    """Soft-arg-max decode + triangulation against the unmodified reference run with kornia's vendored expectation code
    (tests/golden/softargmax_v5_j6.npz): key-points within 2e-4 px, inlier counts equal, 3-D joints within 1e-3 rel /
    1e-2 mm, metric within 1e-4 px -- through the batched entry and through the reference's per-frame signature."""
    from multi_view_active_learning_b200.utils import triangulation as T

    g = golden("softargmax_v5_j6")
    hm = S.render_heatmaps(g["centres"], noise=float(g["noise"]), seed=int(g["heatmap_seed"])) * np.float32(g["gain"])
    stride = int(g["stride"])
    kp = ops.decode_softargmax(_cuda(hm), stride).cpu().numpy()
    np.testing.assert_allclose(kp, g["keypoints_2d"], rtol=0, atol=2e-4)
    out = T.triangulation_batch(_cuda(hm), _cuda(g["P"]), stride, torch.from_numpy(g["valid"]), use_soft_argmax=True)
    assert np.array_equal(out["inlier_count"].cpu().numpy(), g["inlier_count"])
    np.testing.assert_allclose(out["keypoints_3d"].cpu().numpy(), g["keypoints_3d"], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    np.testing.assert_allclose(out["metric"].cpu().numpy(), g["metric"], rtol=0, atol=REPROJ_ATOL_PX)
    one = T.triangulation(torch.from_numpy(hm[1]), torch.from_numpy(g["P"][1]), stride, torch.from_numpy(g["valid"][1]),
                          use_soft_argmax=True)
    assert one["keypoints_2d"].dtype == np.float32 and int(one["inlier_count"]) == int(g["inlier_count"][1])
    np.testing.assert_allclose(one["keypoints_3d"], g["keypoints_3d"][1], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    assert abs(float(one["metric"]) - float(g["metric"][1])) <= REPROJ_ATOL_PX


def test_hp_scores(ops, golden):
    g = golden("hp_scores")
    valid = g["valid"] != 0
    got = ops.score_hp(_cuda(g["heatmaps"])[None], torch.from_numpy(valid)).cpu().numpy()[0]
    assert np.isnan(got[:, ~valid]).all()
    np.testing.assert_allclose(got[:, valid], g["hp_per_map"][:, valid], rtol=0, atol=2e-6)
    np.testing.assert_allclose(got[:, valid], SO.hp_scores(g["heatmaps"])[:, valid], rtol=0, atol=2e-6)
    # non-64-wide maps take the generic kernel
    hm = np.random.default_rng(0).normal(size=(1, 2, 3, 20, 24)).astype(np.float32) * 3
    np.testing.assert_allclose(ops.score_hp(_cuda(hm)).cpu().numpy(), SO.hp_scores(hm), rtol=0, atol=2e-6)


def test_mpe_bsb_scores_vs_reference_golden(ops, golden):
    """Kernels against what the unmodified reference's _compute_mpes / _compute_bsb returned for tests/golden/
    peak_scores.npz (skimage's peak finder served by the restatement, everything else the reference's own code)."""
    g = golden("peak_scores")
    hm, valid = g["heatmaps"], g["valid"].astype(bool)
    V, J = hm.shape[:2]
    mpe = ops.score_peaks(_cuda(hm[None]), "MPE", torch.from_numpy(valid)).cpu().numpy()[0]
    got = np.array([mpe[v, k] for v in range(V) for k in range(J) if valid[k]], dtype=np.float64)
    np.testing.assert_allclose(got, g["mpe_per_map"], rtol=0, atol=2e-5)
    assert np.isnan(mpe[:, ~valid]).all()
    bsb = ops.score_peaks(_cuda(hm[None]), "BSB", torch.from_numpy(valid)).cpu().numpy()[0]
    for cfg, tol in (("AVG", 2e-6), ("STD", 2e-6)):
        assert abs(float(SO.reduce_frame_score(bsb, valid, cfg, "BSB")) - float(g["bsb_" + cfg])) <= tol
        assert abs(float(SO.reduce_frame_score(mpe, valid, cfg, "MPE")) - float(g["mpe_" + cfg])) <= 2e-5


def test_mpe_bsb_scores_vs_restatement(ops):
    """MPE / BSB against the scipy restatement of skimage.peak_local_max (parity unpinned: skimage is not installed).
    Tolerance: float32 softmax / entropy of ~130 peak values -> 2e-5 absolute."""
    pool = S.make_pool(3, 4, 5, seed=12)
    hm = S.render_heatmaps(pool["centres"], noise=0.05, seed=13)
    hm[0, 0, 0] = 0.0                      # flat map: no peak at all
    hm[0, 0, 1] = 0.0
    hm[0, 0, 1, 20, 30] = 1.0              # one-hot: a single peak
    hm[0, 0, 2] = 0.0
    hm[0, 0, 2, 10, 10], hm[0, 0, 2, 40, 50], hm[0, 0, 2, 1, 1] = 2.0, 1.5, 9.0  # the 9.0 sits on the excluded border
    valid = np.ones((3, 5), dtype=bool)
    valid[1, 3] = False
    mpe = ops.score_peaks(_cuda(hm), "MPE", torch.from_numpy(valid)).cpu().numpy()
    bsb = ops.score_peaks(_cuda(hm), "BSB", torch.from_numpy(valid)).cpu().numpy()
    exp_m, exp_b = SO.mpe_scores(hm), SO.bsb_scores(hm)
    v = np.broadcast_to(valid[:, None, :], mpe.shape)
    assert np.isnan(mpe[~v]).all() and np.isnan(bsb[~v]).all()
    np.testing.assert_allclose(mpe[v], exp_m[v], rtol=0, atol=2e-5)
    assert mpe[0, 0, 0] == 0.0 and mpe[0, 0, 1] == 0.0 and abs(mpe[0, 0, 2] - exp_m[0, 0, 2]) < 1e-6
    both = v & ~np.isnan(exp_b)
    np.testing.assert_allclose(bsb[both], exp_b[both], rtol=0, atol=2e-6)
    assert np.array_equal(np.isnan(bsb[v]), np.isnan(exp_b[v]))  # fewer than two peaks -> NaN on both sides
    # narrower maps go through the same kernel
    hm2 = np.random.default_rng(3).normal(size=(1, 2, 3, 24, 20)).astype(np.float32)
    np.testing.assert_allclose(ops.score_peaks(_cuda(hm2), "MPE").cpu().numpy(), SO.mpe_scores(hm2), rtol=0, atol=2e-5)
    np.testing.assert_allclose(ops.score_peaks(_cuda(hm2), "BSB").cpu().numpy(), SO.bsb_scores(hm2), rtol=0, atol=2e-6)


# ------------------------------------------------------------------------------------------------ triangulation
def test_reference_unit_test_known_answer(ops, golden):
    from multi_view_active_learning_b200.utils.triangulation import triangulation

    g = golden("ref_unit_triangulation")
    hm = torch.zeros(8, 19, 64, 64)
    for r, c, v in g["bump_rc"]:
        hm[:, :, int(r), int(c)] = float(v)
    res = triangulation(hm, torch.from_numpy(g["P"]), int(g["stride"]), torch.from_numpy(g["valid"]))
    assert list(res["keypoints_3d"].shape) == [19, 3] and list(res["keypoints_2d"].shape) == [8, 19, 2]
    assert res["keypoints_2d"].dtype == np.int64 and (res["keypoints_2d"] == 88).all()
    assert res["inlier_count"] == 3 and isinstance(res["inlier_count"], np.int64)
    np.testing.assert_allclose(res["keypoints_3d"], g["keypoints_3d"], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    np.testing.assert_allclose(res["keypoints_3d"][0], [-37.2230772535, -125.6428366493, -23.8942096679], atol=1e-6)
    np.testing.assert_allclose(res["metric"], 3.1619149383944163, atol=REPROJ_ATOL_PX)


@pytest.mark.parametrize("name", POOLS)
def test_pool_matches_reference_golden(ops, golden, name):
    g = golden(name)
    stride, valid = int(g["stride"]), g["valid"]
    hm = S.onehot_heatmaps(g["keypoints_2d_unmasked"], stride)
    out = ops.to_numpy(ops.score_pool(_cuda(hm), _cuda(g["P"]), stride, torch.from_numpy(valid),
                                      pair_seed=int(g["pair_seed"])))
    assert np.array_equal(out["keypoints_2d"], g["keypoints_2d"])
    _check_tri(out, g, valid)


@pytest.mark.parametrize("N,V,J,seed,vp", [(48, 8, 19, 1, 1.0), (32, 5, 19, 2, 0.8), (6, 20, 42, 3, 0.8),
                                           (6, 31, 19, 4, 1.0), (16, 12, 5, 5, 0.9), (16, 3, 4, 6, 1.0)])
def test_noisy_pool_vs_oracle(ops, N, V, J, seed, vp):
    pool = S.make_pool(N, V, J, seed=seed, valid_prob=vp, p_outlier=0.15)
    hm = S.render_heatmaps(pool["centres"], noise=0.05, seed=seed + 100)
    ref = O.triangulate_pool(hm, pool["P"], 4, pool["valid"], pair_seed=77, frame_offset=1000)
    out = ops.to_numpy(ops.score_pool(_cuda(hm), _cuda(pool["P"]), 4, torch.from_numpy(pool["valid"]), pair_seed=77,
                                      frame_offset=1000))
    assert np.array_equal(out["keypoints_2d"], ref["keypoints_2d"])
    _check_tri(out, ref, pool["valid"])
    np.testing.assert_allclose(out["reproj_mean"][pool["valid"]], ref["reproj_mean"][pool["valid"]], atol=REPROJ_ATOL_PX)
    assert np.isnan(out["reproj_mean"][~pool["valid"]]).all()
    # ranking order of the uncertainty metric is identical (strategy.py:945-949)
    assert np.array_equal(np.argsort(-out["metric"], kind="stable"), np.argsort(-ref["metric"], kind="stable"))


def test_triangulate_from_float_keypoints_and_explicit_pairs(ops):
    N, V, J = 5, 14, 6
    pool = S.make_pool(N, V, J, seed=21, p_outlier=0.1)
    kp = (pool["centres"] * 4).astype(np.float32)  # sub-pixel float key-points (soft-arg-max path)
    rng = random.Random(5)
    allp = list(itertools.combinations(range(V), 2))
    pairs = np.zeros((N, J, 64, 2), dtype=np.uint8)
    for n in range(N):
        for j in range(J):
            p = list(allp)
            rng.shuffle(p)
            pairs[n, j] = p[:64]
    kp3, rm, inl, mask = O.ransac_pool(pool["P"], kp, np.ones((N, J), bool), pairs.astype(np.int64))
    out = ops.to_numpy(ops.triangulate_ransac(_cuda(kp), _cuda(pool["P"]), None, pairs=_cuda(pairs)))
    assert np.array_equal(out["inliers"], inl) and np.array_equal(out["inlier_mask"].astype(np.uint32), mask)
    np.testing.assert_allclose(out["keypoints_3d"], kp3, rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    np.testing.assert_allclose(out["reproj_mean"], rm, atol=REPROJ_ATOL_PX)


def test_per_frame_api_consumes_python_random_like_reference(ops):
    """V >= 12: the drop-in triangulation() draws its pair subsets with random.shuffle exactly like the reference
    (utils/triangulation.py:279-282), so seeding ``random`` reproduces the reference's choice."""
    from multi_view_active_learning_b200.utils.triangulation import triangulation

    V, J = 13, 4
    pool = S.make_pool(1, V, J, seed=33, p_outlier=0.2)
    kp = np.round(pool["centres"]).astype(np.int64) * 4
    hm = S.onehot_heatmaps(kp)
    valid = np.array([1, 0, 1, 1], dtype=bool)
    random.seed(1234)
    res = triangulation(torch.from_numpy(hm[0]), torch.from_numpy(pool["P"][0]), 4, torch.from_numpy(valid))
    state_after = random.getstate()
    random.seed(1234)
    pairs = np.zeros((1, J, 64, 2), dtype=np.int64)
    for j in range(J):
        if valid[j]:
            p = list(itertools.combinations(set(range(V)), 2))
            random.shuffle(p)
            pairs[0, j] = p[:64]
    assert random.getstate() == state_after
    kp_masked = np.where(valid[None, :, None], kp[0], 0)
    kp3, rm, inl, _ = O.ransac_pool(pool["P"], kp_masked[None], valid[None], pairs)
    assert np.array_equal(res["keypoints_2d"], kp_masked)
    np.testing.assert_allclose(res["keypoints_3d"], kp3[0], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    assert res["inlier_count"] == inl[0][valid].min()
    np.testing.assert_allclose(res["metric"], rm[0][valid].mean(), atol=REPROJ_ATOL_PX)


def test_error_behaviour(ops):
    from multi_view_active_learning_b200._lib import MvalError
    from multi_view_active_learning_b200.utils.triangulation import triangulation

    hm = torch.zeros(1, 3, 64, 64)
    with pytest.raises(AssertionError):  # reference :268 assert len(points) >= 2
        triangulation(hm, torch.zeros(1, 3, 4), 4, torch.ones(3).bool())
    with pytest.raises(ValueError):  # reference :231 np.min([])
        triangulation(torch.zeros(2, 3, 64, 64), torch.zeros(2, 3, 4), 4, torch.zeros(3).bool())
    with pytest.raises(MvalError):
        ops.triangulate_ransac(torch.zeros(1, 40, 2, 2, dtype=torch.int32).cuda(), torch.zeros(1, 40, 3, 4).cuda())
    # empty pool is fine
    out = ops.score_pool(torch.zeros(0, 4, 3, 64, 64).cuda(), torch.zeros(0, 4, 3, 4).double().cuda(), 4)
    assert out["metric"].numel() == 0


def test_host_pipeline_matches_device_path(ops):
    N, V, J = 37, 8, 19
    pool = S.make_pool(N, V, J, seed=8, valid_prob=0.9)
    hm = torch.from_numpy(S.render_heatmaps(pool["centres"], noise=0.05, seed=3)).pin_memory()
    P, valid = torch.from_numpy(pool["P"]), torch.from_numpy(pool["valid"])
    dev = ops.to_numpy(ops.score_pool(hm.cuda(), P.cuda(), 4, valid, pair_seed=5, frame_offset=10))
    host = ops.to_numpy(ops.score_pool_host(hm, P, 4, valid, pair_seed=5, frame_offset=10, chunk_frames=8))
    for k in ("keypoints_2d", "keypoints_3d", "inliers", "inlier_count"):
        assert np.array_equal(dev[k], host[k]), k
    np.testing.assert_array_equal(dev["metric"], host["metric"])
    # the same handle serves the next call (no staging allocation per call) and an unpinned, ragged last chunk
    host2 = ops.to_numpy(ops.score_pool_host(hm[:13].clone(), P[:13], 4, valid[:13], pair_seed=5, frame_offset=10, chunk_frames=8))
    np.testing.assert_array_equal(dev["metric"][:13], host2["metric"])
    assert len(ops._host_pipelines) == 1


def test_host_pipeline_flags_match_triangulation_batch(ops):
    """The host-buffer entry with the flags of triangulation() / _compute_sal_dict (utils/triangulation.py:168-179,
    strategy.py:1072-1094): every variant equals the device-resident triangulation_batch bit for bit."""
    from multi_view_active_learning_b200.utils.triangulation import triangulation_batch

    N, V, J = 29, 5, 19
    pool = S.make_pool(N, V, J, seed=18, valid_prob=0.9)
    hm = torch.from_numpy(S.render_heatmaps(pool["centres"], noise=0.05, seed=5)).pin_memory()
    P, valid = torch.from_numpy(pool["P"]), torch.from_numpy(pool["valid"])
    pipe = ops.HostPipeline(V, J, 64, 64, chunk_frames=8, n_slots=2)
    for flags in ({"map_score": "HP"}, {"map_score": "MPE"}, {"map_score": "BSB"}, {"use_soft_argmax": True},
                  {"use_reprojection_xe": True, "sigma": 1.5}, {"direct_optimization": True},
                  {"use_soft_argmax": True, "map_score": "HP", "use_reprojection_xe": True, "sigma": 2.0}):
        host = ops.to_numpy(pipe.score_pool(hm, P, 4, valid, **flags))
        dev = ops.to_numpy(triangulation_batch(hm.cuda(), P.cuda(), 4, valid, **flags))
        for k in ("keypoints_2d", "keypoints_3d", "inlier_count", "metric") + (("map_score",) if "map_score" in flags else ()):
            np.testing.assert_array_equal(dev[k], host[k], err_msg="%s %s" % (flags, k))
    pipe.close()
    with pytest.raises(ValueError):
        ops.HostPipeline(8, 19).score_pool(hm, P, 4, valid)  # shape of the handle


def test_score_pool_segments_equals_one_buffer(ops):
    """One persistent launch over a pool given as several device buffers (mval_score_pool_segments) == the single-buffer
    call over their concatenation, for every scored variant, with reused outputs."""
    N, V, J = 301, 8, 19
    pool = S.make_pool(N, V, J, seed=31, valid_prob=0.9)
    hm = ops.synth_heatmaps(_cuda(pool["centres"]), noise=0.05, seed=6)
    P, valid = _cuda(pool["P"]), torch.from_numpy(pool["valid"])
    cuts = [0, 1, 150, 150, 299, 301]  # incl. an empty segment
    segs = [hm[a:b].clone() for a, b in zip(cuts[:-1], cuts[1:])]
    out = None
    for ms in (None, "HP", "MPE", "BSB"):
        one = ops.to_numpy(ops.score_pool(hm, P, 4, valid, frame_offset=7, map_score=ms))
        out = ops.score_pool_segments(segs, P, 4, valid, frame_offset=7, map_score=ms, return_keypoints_2d=True,
                                      out=out if ms is None else None)
        got = ops.to_numpy(out)
        for k in got:
            np.testing.assert_array_equal(one[k], got[k], err_msg="%s %s" % (ms, k))
    # the same resident buffer passed several times = chunk passes of the benchmark
    rep = ops.to_numpy(ops.score_pool_segments([hm, hm, hm[:11]], torch.cat([P, P, P[:11]]), 4, torch.cat([valid, valid, valid[:11]])))
    one = ops.to_numpy(ops.score_pool(hm, P, 4, valid))
    np.testing.assert_array_equal(rep["metric"], np.concatenate([one["metric"], one["metric"], one["metric"][:11]]))


def test_full_size_properties(ops):
    """C2-sized chunk (8 views x 19 joints): size-independent properties on 2048 frames -- invariance to the
    chunking / frame offset, duplicated frames score identically, scores are finite and inlier counts in [2, V]."""
    N, V, J = 2048, 8, 19
    pool = S.make_pool(N, V, J, seed=77)
    hm = ops.synth_heatmaps(_cuda(pool["centres"]), noise=0.05, seed=4)
    P = _cuda(pool["P"])
    full = ops.score_pool(hm, P, 4)
    a = ops.score_pool(hm[:1000], P[:1000], 4)
    b = ops.score_pool(hm[1000:], P[1000:], 4, frame_offset=1000)
    assert torch.equal(torch.cat([a["metric"], b["metric"]]), full["metric"])
    assert torch.equal(torch.cat([a["keypoints_3d"], b["keypoints_3d"]]), full["keypoints_3d"])
    m = full["metric"].cpu().numpy()
    ic = full["inlier_count"].cpu().numpy()
    assert np.isfinite(m).all() and (m >= 0).all() and (ic >= 2).all() and (ic <= V).all()
    dup = ops.score_pool(torch.cat([hm[:8], hm[:8]]), torch.cat([P[:8], P[:8]]), 4)
    assert torch.equal(dup["metric"][:8], dup["metric"][8:])
    # and a sample of it against the oracle
    sample = slice(500, 532)
    ref = O.triangulate_pool(hm[sample].cpu().numpy(), pool["P"][sample], 4, pool["valid"][sample])
    out = ops.to_numpy({k: v[sample] for k, v in full.items()})
    assert np.array_equal(out["keypoints_2d"], ref["keypoints_2d"])
    _check_tri(out, ref, pool["valid"][sample])


@pytest.mark.parametrize("N,V,J,vp", [(1500, 8, 19, 1.0), (700, 5, 19, 0.8), (300, 20, 42, 0.8), (200, 31, 19, 1.0),
                                      (333, 2, 3, 0.7), (5, 8, 19, 1.0)])
def test_fused_equals_unfused(ops, N, V, J, vp):
    """mval_score_pool (one persistent fused kernel: TMA ring + decode warps + RANSAC warps) must be bit-identical
    to mval_decode_argmax followed by mval_triangulate_ransac (three launches), for every output."""
    pool = S.make_pool(N, V, J, seed=N + V, valid_prob=vp, p_outlier=0.15)
    hm = ops.synth_heatmaps(_cuda(pool["centres"]), noise=0.05, seed=11)
    P, valid = _cuda(pool["P"]), torch.from_numpy(pool["valid"])
    fused = ops.score_pool(hm, P, 4, valid, pair_seed=3, frame_offset=12345)
    xy = ops.decode_argmax(hm, 4, valid)
    ref = ops.triangulate_ransac(xy, P, valid, pair_seed=3, frame_offset=12345)
    assert torch.equal(fused["keypoints_2d"], xy)
    for k in ("keypoints_3d", "inliers", "inlier_count"):
        assert torch.equal(fused[k], ref[k]), k
    for k in ("metric", "reproj_mean"):
        assert torch.equal(torch.nan_to_num(fused[k], nan=-1.0), torch.nan_to_num(ref[k], nan=-1.0)), k
    # optional outputs may be omitted
    lean = ops.score_pool(hm, P, 4, valid, pair_seed=3, frame_offset=12345, return_keypoints_2d=False)
    assert torch.equal(torch.nan_to_num(lean["metric"], nan=-1.0), torch.nan_to_num(ref["metric"], nan=-1.0))


# ------------------------------------------------------------------------------------------------ ranking
def test_topk_matches_nlargest(ops):
    rng = np.random.default_rng(0)
    n = 5000
    s = rng.normal(size=n).round(2)  # many exact ties
    s[rng.integers(0, n, 50)] = np.nan
    s[10], s[20], s[30] = np.inf, -np.inf, -0.0
    d = {"g%d" % i: float(s[i]) for i in range(n)}
    for k in (1, 7, 100, 4000, n + 5):
        exp = SO.rank_nlargest(d, k)
        idx, val = ops.topk_desc(_cuda(s), k, index_offset=0)
        assert ["g%d" % i for i in idx.cpu().tolist()] == exp
        assert np.array_equal(val.cpu().numpy(), np.array([d[g] for g in exp]))
    idx, _ = ops.topk_desc(_cuda(s), 5, index_offset=1000)
    assert idx.cpu().tolist() == [1000 + int(g[1:]) for g in SO.rank_nlargest(d, 5)]


# ------------------------------------------------------------------------------------------------ reprojection XE
def test_reprojection_xe_golden_and_oracle(ops, golden):
    """utils/triangulation.py:236-257 (_compute_xe): golden values of the reference, then a larger random pool against
    the oracle, then the per-frame drop-in with use_reprojection_xe=True.  float64 sums of 4096 terms in a different
    order: rtol 1e-11."""
    from multi_view_active_learning_b200.utils import triangulation as T

    g = golden("xe_metric")
    hm = S.render_heatmaps(g["centres"], noise=float(g["noise"]), seed=int(g["heatmap_seed"]))
    xe, per_map = ops.score_xe(_cuda(hm), _cuda(g["P"]), _cuda(g["keypoints_3d"]), float(g["sigma"]), return_per_map=True)
    np.testing.assert_allclose(xe.cpu().numpy(), g["xe"], rtol=1e-11, atol=0)
    exp, exp_map = O.compute_xe(g["keypoints_3d"], g["P"], hm, float(g["sigma"]))
    np.testing.assert_allclose(per_map.cpu().numpy(), exp_map, rtol=1e-11, atol=1e-300)
    # random pool, projections inside and outside the grid, a joint at the origin, w == 0
    N, V, J = 9, 5, 7
    pool = S.make_pool(N, V, J, seed=5)
    P = pool["P"].copy()
    P[:5, :, :2, :] /= S.STRIDE
    X = pool["X"].copy()
    X[:, 2] = 0.0
    P[8, 0, 2, :] = 0.0  # w == 0 for every point of that view -> w := 1
    hm = S.render_heatmaps(pool["centres"], noise=0.1, seed=6)
    for sigma in (1.0, 3.0):
        got = ops.score_xe(_cuda(hm), _cuda(P), _cuda(X), sigma).cpu().numpy()
        np.testing.assert_allclose(got, O.compute_xe(X, P, hm, sigma)[0], rtol=1e-11, atol=0)
    # drop-in: metric is the XE of the triangulated joints (a 0-d float64 CUDA tensor, like the reference's)
    n = 2
    r = T.triangulation(_cuda(hm[n]), torch.from_numpy(P[n]), 4, torch.from_numpy(pool["valid"][n]), False, True, 2.0)
    ref = O.triangulate_pool(hm[n:n + 1], P[n:n + 1], 4, pool["valid"][n:n + 1])
    exp = O.compute_xe(r["keypoints_3d"][None], P[n:n + 1], hm[n:n + 1], 2.0)[0][0]
    assert torch.is_tensor(r["metric"]) and r["metric"].is_cuda and r["metric"].dtype == torch.float64
    np.testing.assert_allclose(float(r["metric"].item()), exp, rtol=1e-11)
    np.testing.assert_allclose(r["keypoints_3d"], ref["keypoints_3d"][0], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    # batch entry
    b = T.triangulation_batch(_cuda(hm), _cuda(P), 4, _cuda(pool["valid"]), use_reprojection_xe=True, sigma=2.0)
    expb = O.compute_xe(b["keypoints_3d"].cpu().numpy(), P, hm, 2.0)[0]
    np.testing.assert_allclose(b["metric"].cpu().numpy(), expb, rtol=1e-11)


# ------------------------------------------------------------------------------------------------ coreset
@pytest.mark.parametrize("n,L,d,budget", [(700, 30, 57, 40), (300, 10, 126, 25), (513, 7, 2048, 30), (200, 5, 130, 20),
                                          (100, 3, 3, 10), (5000, 300, 57, 400), (3000, 20, 128, 300)])
def test_kcenter_bit_exact_vs_f32_oracle(ops, n, L, d, budget):
    rng = np.random.default_rng(n + d)
    F = (rng.normal(size=(n + L, d)) * 50).astype(np.float32)
    F[5] = F[3]  # duplicate rows: exact ties in min-distance
    exp_sel, exp_min = CO.kcenter_greedy_f32(F, n, budget)
    sel, min_d = ops.kcenter_greedy(_cuda(F), n, budget)
    assert sel.cpu().tolist() == exp_sel
    assert np.array_equal(min_d.cpu().numpy(), exp_min)  # bit-exact float32 distances, not only indices
    assert np.array_equal(ops.kcenter_norms(_cuda(F)).cpu().numpy(), CO.canonical_dot_f32(F))


@pytest.mark.parametrize("kind", ["clustered", "lattice", "all_equal", "few_distinct"])
def test_kcenter_rounds_on_hard_pools(ops, kind):
    """Pools where a greedy round must stop early (clusters: one pick collapses its neighbours' minima) or where
    almost everything ties (integer lattice, identical rows): the rounds still equal the sequential loop."""
    rng = np.random.default_rng(11)
    n, L, d, budget = 2500, 6, 57, 120
    if kind == "clustered":
        centres = rng.normal(size=(12, d)) * 100
        F = centres[rng.integers(0, 12, n + L)] + rng.normal(size=(n + L, d)) * 0.5
    elif kind == "lattice":
        F = rng.integers(0, 3, size=(n + L, d)).astype(np.float64)
    elif kind == "all_equal":
        F = np.ones((n + L, d))
        budget = 7
    else:
        F = rng.normal(size=(5, d))[rng.integers(0, 5, n + L)] * 10
        budget = 12
    F = F.astype(np.float32)
    exp_sel, exp_min = CO.kcenter_greedy_f32(F, n, budget)
    sel, min_d = ops.kcenter_greedy(_cuda(F), n, budget)
    assert sel.cpu().tolist() == exp_sel
    assert np.array_equal(min_d.cpu().numpy(), exp_min)


def test_kcenter_update_paths_agree(ops):
    """Single-centre update, batched FFMA update (all tile widths, aligned and unaligned d) and the oracle give the same
    running minima bit for bit."""
    rng = np.random.default_rng(3)
    for d in (57, 64, 130, 256):
        n = 1000
        F = (rng.normal(size=(n, d)) * 20).astype(np.float32)
        X = _cuda(F)
        norms = ops.kcenter_norms(X)
        xx = CO.canonical_dot_f32(F)
        assert np.array_equal(norms.cpu().numpy(), xx)
        for T in (1, 2, 16, 17, 64, 65, 200):
            cidx = rng.integers(0, n, T)
            exp = np.full(n, np.inf, dtype=np.float32)
            for c in cidx:
                exp = np.minimum(exp, CO.canonical_dist_f32(F, xx, F[c], xx[c]))
            m = torch.full((n,), float("inf"), dtype=torch.float32, device="cuda")
            ci = torch.as_tensor(cidx, device="cuda")
            ops.kcenter_update_batch(X, norms, X[ci].contiguous(), norms[ci].contiguous(), m, flags=1)
            assert np.array_equal(m.cpu().numpy(), exp), (d, T)
        m = torch.full((n,), float("inf"), dtype=torch.float32, device="cuda")
        best = ops.kcenter_update(X, norms, X[17], m)
        exp = CO.canonical_dist_f32(F, xx, F[17], xx[17])
        assert np.array_equal(m.cpu().numpy(), exp)
        assert int(best[1].item()) == int(np.argmax(exp)) and float(best[0].item()) == float(exp.max())


@pytest.mark.parametrize("n,d", [(5000, 64), (20000, 256), (3000, 2048), (4097, 100)])
def test_kcenter_tensor_core_screen_equals_exact_pass(ops, n, d):
    """tcgen05 TF32 screening GEMM + exact recheck must give the same running minima, bit for bit, as the FFMA pass:
    from +inf (every row has survivors), after some centres (few survivors), with duplicate rows (distance exactly 0)."""
    rng = np.random.default_rng(d)
    F = (rng.normal(size=(n, d)) * 3).astype(np.float32)
    F[n // 2] = F[11]
    X = _cuda(F)
    norms = ops.kcenter_norms(X)
    m_tc = torch.full((n,), float("inf"), dtype=torch.float32, device="cuda")
    m_ex = m_tc.clone()
    for T in (256, 37, 3, 200):
        ci = torch.as_tensor(rng.integers(0, n, T), device="cuda")
        ci[0] = 11
        C, cn = X[ci].contiguous(), norms[ci].contiguous()
        ops.kcenter_update_batch(X, norms, C, cn, m_tc, flags=2)
        survivors, capacity = ops.kcenter_tc_stats()
        ops.kcenter_update_batch(X, norms, C, cn, m_ex, flags=1)
        assert torch.equal(m_tc, m_ex), (n, d, T)
        # the screen really screens: far fewer survivors than pairs, and never the overflow fallback
        assert 0 < survivors <= capacity and survivors <= max(4 * n, n * T // 8), (n, d, T, survivors)
    assert float(m_tc[n // 2].item()) == 0.0
    # against the oracle for one batch
    xx = CO.canonical_dot_f32(F)
    exp = np.full(n, np.inf, dtype=np.float32)
    cidx = rng.integers(0, n, 40)
    for c in cidx:
        exp = np.minimum(exp, CO.canonical_dist_f32(F, xx, F[c], xx[c]))
    m = torch.full((n,), float("inf"), dtype=torch.float32, device="cuda")
    ci = torch.as_tensor(cidx, device="cuda")
    ops.kcenter_update_batch(X, norms, X[ci].contiguous(), norms[ci].contiguous(), m, flags=2)
    assert np.array_equal(m.cpu().numpy(), exp)


def test_kcenter_greedy_with_tensor_core_updates(ops):
    """Whole selection with the tensor-core update path (auto-selected at this size) against the C oracle."""
    from multi_view_active_learning_b200 import pool as P

    rng = np.random.default_rng(21)
    n, L, d, budget = 20000, 40, 128, 300
    F = (rng.normal(size=(n + L, d)) * 2).astype(np.float32)
    exp_sel, exp_min = CO.kcenter_greedy_f32(F, n, budget)
    sel, min_d = ops.kcenter_greedy(_cuda(F), n, budget)
    assert sel.cpu().tolist() == exp_sel
    assert np.array_equal(min_d.cpu().numpy(), exp_min)
    sel2, mins = P.kcenter_greedy_sharded([(_cuda(F[:9000]), 0), (_cuda(F[9000:n]), 9000)], _cuda(F[n:]), budget, flags=2)
    assert sel2.cpu().tolist() == exp_sel
    assert np.array_equal(torch.cat(mins).cpu().numpy(), exp_min[:n])


def test_coreset_class_matches_reference_golden(ops, golden):
    from multi_view_active_learning_b200.utils.coreset import CoreSet

    g = golden("coreset_random")
    keys = [str(k) for k in g["sal_keys"]]
    sal = {k: p.tolist() for k, p in zip(keys, g["sal_poses"])}
    al = {i: p for i, p in enumerate(g["al_poses"])}
    cs = CoreSet(sal, al, int(g["root"]))
    assert np.array_equal(cs.features, g["features"])
    picked = cs.select_batch(int(g["budget"]))
    assert picked == [keys[i] for i in g["picked"]]
    np.testing.assert_allclose(cs.min_distances, g["min_distances"], rtol=1e-4, atol=1e-2)
    assert cs.n_obs == 425 and cs.al_indices == list(range(400, 425))
    # incremental API (update_distances + repeated select_batch) continues the same greedy sequence
    cs2 = CoreSet(sal, al, int(g["root"]))
    first = cs2.select_batch(10)
    cs2.update_distances([keys.index(k) for k in first])
    assert first == picked[:10]

    g2 = golden("ref_unit_coreset")
    pose = [[0, 1, 2] for _ in range(19)]
    sal2 = {"k%d" % i: pose for i in range(20)}
    cs3 = CoreSet(sal2, {i: pose for i in range(5)}, 2)
    assert cs3.select_batch(5) == ["k%d" % i for i in g2["picked"]] == ["k0"] * 5


def test_kcenter_sharded_loop_bit_exact(ops):
    """The multi-GPU greedy rounds (per-shard candidate records -> replay on the union -> batched update) with the ranks
    emulated as shards on one device: uneven shards, an empty shard, duplicate rows across shards."""
    from multi_view_active_learning_b200 import pool as P

    rng = np.random.default_rng(7)
    n, L, d, budget = 900, 12, 57, 60
    F = (rng.normal(size=(n + L, d)) * 30).astype(np.float32)
    F[700] = F[10]  # the same row in two different shards: the lower global index must win the tie
    exp_sel, exp_min = CO.kcenter_greedy_f32(F, n, budget)
    cuts = [0, 250, 250, 610, 900]  # shard 1 is empty
    shards = [(_cuda(F[a:b]), a) for a, b in zip(cuts[:-1], cuts[1:])]
    sel, mins = P.kcenter_greedy_sharded(shards, _cuda(F[n:]), budget)
    assert sel.cpu().tolist() == exp_sel
    assert np.array_equal(torch.cat(mins).cpu().numpy(), exp_min[:n])
    # one shard == the single-device loop
    stats = []
    sel1, _ = P.kcenter_greedy_sharded([(_cuda(F[:n]), 0)], _cuda(F[n:]), budget, stats=stats)
    assert sel1.cpu().tolist() == exp_sel
    assert sum(stats) == budget and len(stats) < budget  # rounds really batch several picks
    # tiny candidate sets (k_slots = 4): many short rounds, same answer
    sel2, _ = P.kcenter_greedy_sharded(shards, _cuda(F[n:]), budget, k_slots=4)
    assert sel2.cpu().tolist() == exp_sel


# ------------------------------------------------------------------------------------------------ fused pass with AL scores
def _with_env(name, value, fn):
    import os

    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.parametrize("N,V,J,vp", [(1500, 8, 19, 1.0), (700, 5, 19, 0.8), (300, 20, 42, 0.8), (200, 31, 19, 1.0),
                                      (333, 2, 3, 0.7), (5, 8, 19, 0.9)])
def test_fused_scored_pass_equals_separate_passes(ops, N, V, J, vp):
    """mval_score_pool_scored: the decode warps of the fused kernel evaluate HP / MPE / BSB on the staged map right after
    its arg-max (one pass over the pool instead of two).  Every triangulation output must be bit-identical to the
    unscored fused pass and the per-map scores must equal those of mval_score_hp / mval_score_peaks (same device code:
    bit-identical), over several ring rounds per SM, invalid joints, 2 and 4 frame slots, and C(V,2) > 64.  The
    passes are run with both arg-max flavours (MVAL_ROW_ARGMAX: generic per-vector scan / lane = row sweep)."""
    pool = S.make_pool(N, V, J, seed=N + V, valid_prob=vp, p_outlier=0.15)
    hm = ops.synth_heatmaps(_cuda(pool["centres"]), noise=0.05, seed=11)
    P, valid = _cuda(pool["P"]), torch.from_numpy(pool["valid"])
    v = np.broadcast_to(pool["valid"][:, None, :], (N, V, J))
    plain = ops.score_pool(hm, P, 4, valid, pair_seed=3, frame_offset=12345)
    for flavour in ("0", "1"):  # generic per-vector scan / lane = row sweep (the default on 64 x 64 maps)
        rows = _with_env("MVAL_ROW_ARGMAX", flavour, lambda: ops.score_pool(hm, P, 4, valid, pair_seed=3, frame_offset=12345))
        assert torch.equal(plain["keypoints_2d"], rows["keypoints_2d"]) and torch.equal(plain["keypoints_3d"], rows["keypoints_3d"])
    assert torch.equal(plain["keypoints_2d"], ops.decode_argmax(hm, 4, valid))
    # MPE / BSB: the fused kernel with either arg-max flavour (MVAL_SCORED_SPLIT=0) and the default split path (stream kernel
    # with score + arg-max key-point, then RANSAC from the key-points) must all give the same bits
    for kind, flavour, split in (("HP", "1", "1"), ("MPE", "0", "0"), ("MPE", "1", "0"), ("BSB", "0", "0"), ("BSB", "1", "0"),
                                 ("MPE", "1", "1"), ("BSB", "1", "1")):
        both = _with_env("MVAL_SCORED_SPLIT", split, lambda: _with_env(
            "MVAL_ROW_ARGMAX", flavour, lambda: ops.score_pool(hm, P, 4, valid, pair_seed=3, frame_offset=12345, map_score=kind)))
        for k in ("keypoints_2d", "keypoints_3d", "inliers", "inlier_count"):
            assert torch.equal(both[k], plain[k]), (kind, k)
        for k in ("metric", "reproj_mean"):
            assert torch.equal(torch.nan_to_num(both[k], nan=-1.0), torch.nan_to_num(plain[k], nan=-1.0)), (kind, k)
        sep = (ops.score_hp(hm, valid) if kind == "HP" else ops.score_peaks(hm, kind, valid)).cpu().numpy()
        got = both["map_score"].cpu().numpy()
        assert got.shape == (N, V, J)
        assert np.isnan(got[~v]).all() and np.array_equal(np.isnan(got), np.isnan(sep)), kind
        assert np.array_equal(got[v], sep[v]), (kind, np.nanmax(np.abs(got[v] - sep[v])))


def test_fused_argmax_on_reference_decode_edge_cases(ops, golden):
    """The arg-max of the fused pass (lane = row sweep; HP variant: row maxima of the softmax sweep) on the maps of
    tests/golden/decode_edge_cases.npz -- ties (first index wins), NaN (counts as the maximum), +-inf, -0.0 vs +0.0,
    invalid joints -- against what the unmodified reference's get_scaled_pred_corrdinates returned."""
    g = golden("decode_edge_cases")
    hm = _cuda(np.stack([g["heatmaps"], g["heatmaps"][::-1]]))  # [2 frames, 3 views, 5 joints, 64, 64]
    P = _cuda(S.make_pool(2, 3, 5, seed=9)["P"])
    exp = np.stack([g["scaled"], g["scaled"][::-1]])
    exp_all = np.stack([g["scaled_all_valid"], g["scaled_all_valid"][::-1]])
    valid = torch.from_numpy(g["valid"])
    for kind, flavour in ((None, "0"), (None, "1"), ("HP", "1"), ("MPE", "0"), ("MPE", "1"), ("BSB", "0"), ("BSB", "1")):
        run = lambda v: _with_env("MVAL_ROW_ARGMAX", flavour, lambda: ops.score_pool(hm, P, int(g["stride"]), v, map_score=kind))
        assert np.array_equal(run(valid)["keypoints_2d"].cpu().numpy(), exp), (kind, flavour)
        assert np.array_equal(run(None)["keypoints_2d"].cpu().numpy(), exp_all), (kind, flavour)
    # ties across the two rows of one lane (rows r and r + 32), across lanes, and inside one row; -0.0 before +0.0
    t = np.full((1, 2, 4, 64, 64), -1.0, dtype=np.float32)
    t[0, :, 0, 40, 7] = t[0, :, 0, 8, 9] = t[0, :, 0, 8, 50] = 2.0    # rows 8 and 40 belong to lane 8: first is (8, 9)
    t[0, :, 1, 63, 63] = t[0, :, 1, 31, 0] = 5.0                       # lanes 31 (row 63) and 31 (row 31): row 31 first
    t[0, :, 2] = -0.0
    t[0, :, 2, 3, 3] = 0.0                                             # all zeros of either sign: index 0 wins
    t[0, :, 3, 0, 1] = t[0, :, 3, 0, 0] = 7.0
    Pt = _cuda(S.make_pool(1, 2, 4, seed=10)["P"])
    want = np.array([[9, 8], [0, 31], [0, 0], [0, 0]]) * 4
    for kind, flavour in ((None, "0"), (None, "1"), ("HP", "1"), ("MPE", "1"), ("BSB", "1")):
        got = _with_env("MVAL_ROW_ARGMAX", flavour, lambda: ops.score_pool(_cuda(t), Pt, 4, None, map_score=kind))
        got = got["keypoints_2d"].cpu().numpy()
        assert np.array_equal(got[0, 0], want) and np.array_equal(got[0, 1], want), (kind, flavour)
    assert np.array_equal(ops.decode_argmax(_cuda(t), 4).cpu().numpy()[0, 0], want)


def test_fused_scored_pass_edge_cases(ops):
    """NaN / infinity / constant maps, maps with fewer than two peaks, zero frames and a shape the fused kernel does not
    take (48 x 48: the entry point then runs the two passes back to back) through mval_score_pool_scored."""
    rng = np.random.default_rng(2)
    hm = rng.normal(size=(3, 2, 6, 64, 64)).astype(np.float32)
    hm[0, :, 1, 10, 10] = np.nan
    hm[0, :, 2, 5, 5] = np.inf
    hm[1, :, 3, 7, :] = -np.inf
    hm[1, :, 4] = 3.0
    hm[2, 0, 5] = -1.0
    hm[2, 0, 5, 20, 20] = 0.5  # one peak only: BSB is NaN there
    P = _cuda(S.make_pool(3, 2, 6, seed=4)["P"])
    valid = torch.ones(3, 6, dtype=torch.bool)
    valid[2, 0] = False
    g = _cuda(hm)
    plain = ops.score_pool(g, P, 4, valid)
    # the lane = row arg-max sweep falls back to the exact scan on maps with NaN / infinities: same key-points as the
    # stand-alone decode kernel and as the generic scan
    assert torch.equal(plain["keypoints_2d"], ops.decode_argmax(g, 4, valid))
    assert torch.equal(plain["keypoints_2d"], _with_env("MVAL_ROW_ARGMAX", "1", lambda: ops.score_pool(g, P, 4, valid))["keypoints_2d"])
    for kind in ("HP", "MPE", "BSB"):
        both = ops.score_pool(g, P, 4, valid, map_score=kind)
        assert torch.equal(both["keypoints_2d"], plain["keypoints_2d"]), kind
        assert torch.equal(both["inlier_count"], plain["inlier_count"]), kind
        assert torch.equal(torch.nan_to_num(both["keypoints_3d"], nan=-1.0), torch.nan_to_num(plain["keypoints_3d"], nan=-1.0)), kind
        sep = (ops.score_hp(g, valid) if kind == "HP" else ops.score_peaks(g, kind, valid)).cpu().numpy()
        got = both["map_score"].cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(sep)), kind
        assert np.array_equal(got[~np.isnan(got)], sep[~np.isnan(sep)]), kind
        assert np.isnan(got[2, :, 0]).all()  # invalid joint
        if kind == "HP":  # a NaN / an infinity poisons every row softmax it takes part in
            assert np.isnan(got[0, :, 1]).all() and np.isnan(got[0, :, 2]).all()
        empty = ops.score_pool(g[:0], P[:0], 4, map_score=kind)
        assert empty["map_score"].shape == (0, 2, 6) and empty["metric"].numel() == 0
    small = _cuda(rng.normal(size=(7, 3, 4, 48, 48)).astype(np.float32))
    Ps = _cuda(S.make_pool(7, 3, 4, seed=5)["P"])
    for kind in ("HP", "MPE", "BSB"):
        both = ops.score_pool(small, Ps, 4, map_score=kind)
        sep = ops.score_hp(small) if kind == "HP" else ops.score_peaks(small, kind)
        assert torch.equal(torch.nan_to_num(both["map_score"], nan=-9.0), torch.nan_to_num(sep, nan=-9.0)), kind
        assert torch.equal(both["keypoints_2d"], ops.decode_argmax(small, 4))
    with pytest.raises(ValueError):
        ops.score_pool(g, P, 4, map_score="XE")


# ------------------------------------------------------------------------------------------------ map-stream kernels
def _legacy(fn):
    """Runs fn with the persistent TMA-ring kernels (csrc/mapstream.cu) switched off -> one-warp-per-map kernels."""
    import os

    os.environ["MVAL_NO_STREAM"] = "1"
    try:
        return fn()
    finally:
        os.environ["MVAL_NO_STREAM"] = "0"


def test_map_stream_kernels_match_per_map_kernels(ops):
    """The persistent kernels (64 x 64 maps) against the one-warp-per-map kernels on a pool large enough for several
    rounds of the 12-stage ring on every SM (148 x 12 x 3 maps), with invalid joints (never read) and a ragged tail."""
    N, V, J = 41, 8, 19  # 6232 maps
    pool = S.make_pool(N, V, J, seed=21, valid_prob=0.85)
    hm_np = S.render_heatmaps(pool["centres"], noise=0.05, seed=22)
    hm, valid = _cuda(hm_np), torch.from_numpy(pool["valid"])
    v = np.broadcast_to(pool["valid"][:, None, :], (N, V, J))

    soft_s = ops.decode_softargmax(hm, 4).cpu().numpy()
    soft_l = _legacy(lambda: ops.decode_softargmax(hm, 4)).cpu().numpy()
    np.testing.assert_allclose(soft_s, soft_l, rtol=0, atol=1e-4)
    np.testing.assert_allclose(soft_s[:4], O.decode_softargmax(hm_np[:4], 4), rtol=0, atol=2e-4)

    hp_s = ops.score_hp(hm, valid).cpu().numpy()
    hp_l = _legacy(lambda: ops.score_hp(hm, valid)).cpu().numpy()
    assert np.isnan(hp_s[~v]).all() and not np.isnan(hp_s[v]).any()
    np.testing.assert_allclose(hp_s[v], hp_l[v], rtol=0, atol=1e-6)

    for mode, atol in (("MPE", 2e-5), ("BSB", 2e-6)):
        a = ops.score_peaks(hm, mode, valid).cpu().numpy()
        b = _legacy(lambda: ops.score_peaks(hm, mode, valid)).cpu().numpy()
        assert np.isnan(a[~v]).all() and not np.isnan(a[v]).any()
        np.testing.assert_allclose(a[v], b[v], rtol=0, atol=atol)
    sub = slice(0, 2)
    np.testing.assert_allclose(ops.score_peaks(hm[sub], "MPE").cpu().numpy(), SO.mpe_scores(hm_np[sub]), rtol=0, atol=2e-5)
    np.testing.assert_allclose(ops.score_peaks(hm[sub], "BSB").cpu().numpy(), SO.bsb_scores(hm_np[sub]), rtol=0, atol=2e-6)

    P, X = _cuda(pool["P"]), _cuda(pool["X"])
    P4 = P.clone()
    P4[: N // 2, :, :2, :] /= 4  # half of the frames project onto the 64 x 64 grid (non-trivial renders)
    xe_s, map_s = ops.score_xe(hm, P4, X, 2.0, return_per_map=True)
    xe_l, map_l = _legacy(lambda: ops.score_xe(hm, P4, X, 2.0, return_per_map=True))
    np.testing.assert_allclose(map_s.cpu().numpy(), map_l.cpu().numpy(), rtol=1e-12, atol=0)
    np.testing.assert_allclose(xe_s.cpu().numpy(), xe_l.cpu().numpy(), rtol=1e-12, atol=0)
    np.testing.assert_allclose(xe_s[:3].cpu().numpy(), O.compute_xe(pool["X"][:3], P4[:3].cpu().numpy(), hm_np[:3], 2.0)[0],
                               rtol=1e-11, atol=0)


def test_map_stream_peak_edge_cases(ops):
    """Peaks next to the excluded 2-pixel border, on the seam between the two row streams (rows 30..33), plateaus at the
    map minimum, and a map with exactly two peaks -- persistent kernel vs the scipy restatement."""
    hm = np.zeros((1, 1, 8, 64, 64), dtype=np.float32)
    hm[0, 0, 0, 2, 2], hm[0, 0, 0, 61, 61], hm[0, 0, 0, 1, 30], hm[0, 0, 0, 30, 62] = 3.0, 2.0, 9.0, 9.0  # corners in, border out
    for r in (29, 30, 31, 32, 33, 34):
        hm[0, 0, 1, r, 3 * (r - 28) + 5] = 1.0 + 0.1 * r  # one peak per seam row, far apart
    hm[0, 0, 2, 31, 31], hm[0, 0, 2, 32, 33] = 2.0, 2.5  # neighbours across the seam: only the larger survives
    hm[0, 0, 3] = 1.0
    hm[0, 0, 3, 10:20, 10:20] = 0.5  # a plateau of 1.0 above the minimum: touching plateau pixels are pruned (ensure_spacing)
    hm[0, 0, 4] = -1.0
    hm[0, 0, 4, 20, 20], hm[0, 0, 4, 40, 40] = 0.0, 0.25
    rng = np.random.default_rng(8)
    hm[0, 0, 5] = rng.normal(size=(64, 64)).astype(np.float32)
    hm[0, 0, 6] = rng.uniform(size=(64, 64)).astype(np.float32) * 10
    hm[0, 0, 7, 15, 15], hm[0, 0, 7, 15, 19], hm[0, 0, 7, 15, 17] = 1.0, 1.0, 0.5  # equal peaks, 4 apart
    mpe = ops.score_peaks(_cuda(hm), "MPE").cpu().numpy()
    bsb = ops.score_peaks(_cuda(hm), "BSB").cpu().numpy()
    exp_m, exp_b = SO.mpe_scores(hm), SO.bsb_scores(hm)
    keep = np.ones(8, dtype=bool)  # map 3 (the plateau) is part of the comparison since round 2
    np.testing.assert_allclose(mpe[0, 0, keep], exp_m[0, 0, keep], rtol=0, atol=2e-5)
    ok = keep & ~np.isnan(exp_b[0, 0])
    np.testing.assert_allclose(bsb[0, 0, ok], exp_b[0, 0, ok], rtol=0, atol=2e-6)
    assert np.array_equal(np.isnan(bsb[0, 0, keep]), np.isnan(exp_b[0, 0, keep]))
    legacy_m = _legacy(lambda: ops.score_peaks(_cuda(hm), "MPE")).cpu().numpy()
    keep[3] = False  # the generic-shape kernel (one warp streaming a map of any size) does not prune plateaus
    np.testing.assert_allclose(mpe[0, 0, keep], legacy_m[0, 0, keep], rtol=0, atol=2e-5)


def test_peak_plateaus_follow_ensure_spacing(ops):
    """skimage's peak_local_max drops a peak that touches an already accepted one (ensure_spacing, Chebyshev distance < 2);
    peaks are visited by descending value, equal values in flat-index order (strategy.py:1168-1170, 1204-1206).  Touching
    5 x 5 maxima are always equal, so this is about plateaus: pairs, runs, blocks, across the kernel's lane boundary (columns
    4l+3 | 4l+4) and across its two row streams (rows 31 | 32).  MPE sees the raw maps, BSB their row softmax."""
    rng = np.random.default_rng(3)
    M = 10
    hm = (rng.uniform(size=(1, 1, M, 64, 64)) * 0.1).astype(np.float32)
    hm[0, 0, 0, 20, 20] = hm[0, 0, 0, 20, 21] = 2.0          # horizontal pair + a lower peak
    hm[0, 0, 0, 40, 40] = 1.5
    hm[0, 0, 1, 31, 25] = hm[0, 0, 1, 32, 25] = 2.0          # vertical pair across the two row streams
    hm[0, 0, 1, 31, 40] = hm[0, 0, 1, 32, 41] = 1.75         # diagonal pair across them
    hm[0, 0, 1, 50, 10] = 1.25
    hm[0, 0, 2, 12, 7] = hm[0, 0, 2, 12, 8] = 2.0            # pair across a lane boundary (columns 7 | 8)
    hm[0, 0, 2, 13, 11] = hm[0, 0, 2, 12, 12] = 1.5          # anti-diagonal pair across one (row 12 col 12 comes first)
    hm[0, 0, 3, 30, 30:33] = 2.0                             # run of three: first and third survive
    hm[0, 0, 3, 45, 10:17] = 1.5                             # run of seven: every second one
    hm[0, 0, 4, 10:12, 10:12] = 2.0                          # 2 x 2 block: only its first pixel
    hm[0, 0, 4, 30:33, 40:43] = 1.5                          # 3 x 3 block: its four corners
    hm[0, 0, 5, 20:26, 20] = 2.0                             # vertical run of six
    hm[0, 0, 6, 29:35, 29:35] = 3.0                          # block over the seam rows
    hm[0, 0, 7] = 1.0                                        # everything a plateau except a dent ...
    hm[0, 0, 7, 5:9, 50:60] = 0.25
    hm[0, 0, 8, 2, 2] = hm[0, 0, 8, 2, 3] = 2.0              # pair at the edge of the admissible region
    hm[0, 0, 8, 61, 60] = hm[0, 0, 8, 61, 61] = 2.0
    hm[0, 0, 9, 20, 20] = hm[0, 0, 9, 20, 22] = 2.0          # NOT touching (distance 2): both stay
    mpe = ops.score_peaks(_cuda(hm), "MPE").cpu().numpy()
    np.testing.assert_allclose(mpe[0, 0], SO.mpe_scores(hm)[0, 0], rtol=0, atol=2e-5)
    # the same maps through the fused pool pass (decode warps evaluate the score on the staged map)
    pool = S.make_pool(1, 2, M, seed=1)
    hm2 = np.concatenate([hm, hm], axis=1)
    fused = ops.score_pool(_cuda(hm2), _cuda(pool["P"]), 4, map_score="MPE")["map_score"].cpu().numpy()
    np.testing.assert_array_equal(fused[0, 0], mpe[0, 0])
    # BSB: equal values in one row stay equal after the row softmax; the two best SURVIVING peaks are compared
    b = np.zeros((1, 1, 4, 64, 64), dtype=np.float32)
    b[0, 0, 0, 20, 20] = b[0, 0, 0, 20, 21] = 6.0            # touching pair: one survives, second best is the 4.0 below
    b[0, 0, 0, 40, 40] = 4.0
    b[0, 0, 1, 20, 7] = b[0, 0, 1, 20, 8] = 6.0              # the same across a lane boundary
    b[0, 0, 1, 40, 40] = 5.0
    b[0, 0, 2, 30, 30:33] = 6.0                              # run of three: first and third survive -> difference 0
    b[0, 0, 3, 31, 25] = b[0, 0, 3, 32, 25] = 6.0            # identical rows: vertical pair across the seam
    b[0, 0, 3, 50, 50] = 3.0
    bsb = ops.score_peaks(_cuda(b), "BSB").cpu().numpy()
    exp_b = SO.bsb_scores(b)
    np.testing.assert_allclose(bsb[0, 0], exp_b[0, 0], rtol=0, atol=2e-6)
    assert bsb[0, 0, 2] == 0.0 and bsb[0, 0, 0] > 1e-5 and bsb[0, 0, 1] > 1e-3  # not the 0 of two touching plateau pixels


# ------------------------------------------------------------------------------------------------ Huber refinement
def test_huber_refinement_golden_and_oracle(ops, golden):
    """direct_optimization=True (utils/triangulation.py:319-336).  scipy stops at ftol = xtol = 1e-8, the kernel runs
    its damped Newton iteration to convergence, so the two differ by scipy's termination slack (observed < 5e-3 mm):
    inside the 1e-2 mm / 1e-3 relative contract, and the kernel's Huber cost is never above scipy's."""
    g = golden("huber_v8_j19")
    valid, P = g["valid"], g["P"]
    kp = np.where(valid[:, None, :, None], g["keypoints_2d_unmasked"], 0)
    out = ops.triangulate_ransac(_cuda(kp.astype(np.int32)), _cuda(P), torch.from_numpy(valid), direct_optimization=True)
    got = ops.to_numpy(out)
    np.testing.assert_allclose(got["keypoints_3d"], g["keypoints_3d"], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    assert np.abs(got["keypoints_3d"] - g["keypoints_3d"]).max() < 1e-2
    np.testing.assert_allclose(got["metric"], g["metric"], rtol=0, atol=REPROJ_ATOL_PX)
    assert np.array_equal(got["inlier_count"], g["inlier_count"])
    assert (got["keypoints_3d"][~valid] == 0).all()
    assert got["refine_iters"][valid].max() < 100 and got["refine_iters"][valid].min() >= 1
    # cost at the kernel's point <= cost at the reference's point, joint by joint
    base = O.triangulate_pool(None, P, 4, valid, keypoints_2d=kp)
    N, V, J = kp.shape[:3]
    for n in range(N):
        for j in range(J):
            if not valid[n, j]:
                continue
            views = [v for v in range(V) if base["inlier_mask"][n, j] >> v & 1]
            pts, Pi = kp[n, views, j].astype(np.float64), P[n, views]
            assert O.huber_cost(got["keypoints_3d"][n, j], pts, Pi) <= O.huber_cost(g["keypoints_3d"][n, j], pts, Pi) + 1e-12
    # drop-in entry point and the batched entry
    from multi_view_active_learning_b200.utils import triangulation as T

    hm = S.onehot_heatmaps(g["keypoints_2d_unmasked"][:2], 4)
    r = T.triangulation(torch.from_numpy(hm[1]), torch.from_numpy(P[1]), 4, torch.from_numpy(valid[1]), direct_optimization=True)
    np.testing.assert_allclose(r["keypoints_3d"], g["keypoints_3d"][1], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    np.testing.assert_allclose(r["metric"], g["metric"][1], rtol=0, atol=REPROJ_ATOL_PX)
    b = T.triangulation_batch(_cuda(hm), _cuda(P[:2]), 4, _cuda(valid[:2]), direct_optimization=True)
    np.testing.assert_allclose(b["keypoints_3d"].cpu().numpy(), got["keypoints_3d"][:2], rtol=0, atol=1e-9)
    # float key-points (soft-arg-max path) against the scipy oracle
    pool = S.make_pool(5, 6, 7, seed=41, p_outlier=0.2)
    kpf = (pool["centres"] * 4 + np.random.default_rng(1).normal(scale=1.5, size=pool["centres"].shape)).astype(np.float32)
    o2 = ops.to_numpy(ops.triangulate_ransac(_cuda(kpf), _cuda(pool["P"]), torch.from_numpy(pool["valid"]), direct_optimization=True))
    e2 = O.triangulate_pool(None, pool["P"], 4, pool["valid"], keypoints_2d=kpf, direct_optimization=True)
    np.testing.assert_allclose(o2["keypoints_3d"], e2["keypoints_3d"], rtol=XYZ_RTOL, atol=XYZ_ATOL_MM)
    np.testing.assert_allclose(o2["metric"], e2["metric"], rtol=0, atol=REPROJ_ATOL_PX)


def test_empty_inputs_and_nan_maps(ops):
    """Zero-frame pools through every entry point (the data loader's last batch can be empty on some ranks) and the
    NaN / infinity conventions of the per-map scores on the persistent kernels."""
    z = torch.zeros(0, 4, 3, 64, 64).cuda()
    P0 = torch.zeros(0, 4, 3, 4).double().cuda()
    assert ops.decode_argmax(z, 4).shape == (0, 4, 3, 2) and ops.decode_softargmax(z, 4).shape == (0, 4, 3, 2)
    assert ops.score_hp(z).numel() == 0 and ops.score_peaks(z, "MPE").numel() == 0 and ops.score_peaks(z, "BSB").numel() == 0
    assert ops.score_xe(z, P0, torch.zeros(0, 3, 3).double().cuda(), 2.0).numel() == 0
    out = ops.triangulate_ransac(torch.zeros(0, 4, 3, 2, dtype=torch.int32).cuda(), P0, direct_optimization=True)
    assert out["keypoints_3d"].shape == (0, 3, 3) and out["metric"].numel() == 0
    assert ops.sal_rank(torch.zeros(0).cuda(), torch.zeros(0).cuda(), None, 7.0, 5).numel() == 0
    assert ops.mkpe(torch.zeros(0, 3, 3).cuda(), torch.zeros(0, 4, 3).cuda(), torch.zeros(0, 3).cuda()).numel() == 0
    assert ops.pose_features(torch.zeros(0, 19, 3).double().cuda(), 2).shape == (0, 57)

    rng = np.random.default_rng(2)
    hm = rng.normal(size=(1, 1, 6, 64, 64)).astype(np.float32)
    hm[0, 0, 1, 10, 10] = np.nan       # NaN poisons softmax-based scores of that map only
    hm[0, 0, 2, 5, 5] = np.inf         # inf - inf = NaN inside the softmax
    hm[0, 0, 3, 7, :] = -np.inf        # an all -inf row: NaN row sum
    hm[0, 0, 4] = 3.0                  # constant map
    hp = ops.score_hp(_cuda(hm)).cpu().numpy()[0, 0]
    exp = SO.hp_scores(hm)[0, 0]
    assert np.array_equal(np.isnan(hp), np.isnan(exp)) and np.isnan(hp[[1, 2, 3]]).all()
    np.testing.assert_allclose(hp[[0, 4, 5]], exp[[0, 4, 5]], rtol=0, atol=2e-6)
    assert abs(hp[4] - (1 - 1 / 64)) < 1e-6
    soft = ops.decode_softargmax(_cuda(hm), 4).cpu().numpy()[0, 0]
    assert np.isnan(soft[1]).all() and np.isnan(soft[2]).all() and np.isfinite(soft[[0, 3, 4, 5]]).all()
    np.testing.assert_allclose(soft[4], [31.5 * 4, 31.5 * 4], atol=1e-3)  # uniform weights -> the grid centre
    legacy = _legacy(lambda: ops.decode_softargmax(_cuda(hm), 4)).cpu().numpy()[0, 0]
    assert np.array_equal(np.isnan(soft), np.isnan(legacy))
    mpe = ops.score_peaks(_cuda(hm), "MPE").cpu().numpy()[0, 0]
    assert mpe[4] == 0.0  # constant map: no peak (every pixel sits at the map minimum)
    np.testing.assert_allclose(mpe[[0, 5]], SO.mpe_scores(hm[:, :, [0, 5]])[0, 0], rtol=0, atol=2e-5)


# ------------------------------------------------------------------------------------------------ inlier threshold
def _planted_threshold_pool(n_cases, V, delta, rng):
    """Frames with one joint whose views are EXACT (integer key-points, principal points shifted so that the joint
    projects onto them to ~1e-13 px) except one planted view whose key-point sits 2 * (5 + delta) px from the projection:
    its half-distance reprojection error is 5 + delta for every view pair that does not contain it."""
    base = S.ring_cameras(V, rng=rng)
    P = np.broadcast_to(base, (n_cases,) + base.shape).copy()
    X = rng.uniform(-300, 300, size=(n_cases, 1, 3))
    uv = S.project(P, X)  # [n, V, 1, 2]
    kp = np.round(uv)
    planted = rng.integers(0, V, size=n_cases)
    ang = rng.uniform(0, 2 * np.pi, size=n_cases)
    want = kp.copy()  # where the joint must project to
    for i in range(n_cases):
        want[i, planted[i], 0] -= 2.0 * (5.0 + delta) * np.array([np.cos(ang[i]), np.sin(ang[i])])
    shift = want - uv  # move every projection onto its target through the principal point (exact: P[0] += s * P[2])
    P[:, :, 0, :] += shift[:, :, 0, 0:1] * P[:, :, 2, :]
    P[:, :, 1, :] += shift[:, :, 0, 1:2] * P[:, :, 2, :]
    return P, kp.astype(np.int32), planted


def _single_vote(ops, P, kp, planted):
    """One RANSAC vote per joint from an explicit view pair that does NOT contain the planted view (n_iters = 1), so the
    planted view's membership of the final inlier set is decided by exactly that one comparison with the threshold."""
    n, V = kp.shape[:2]
    pairs = np.zeros((n, 1, 1, 2), dtype=np.uint8)
    for i in range(n):
        others = [v for v in range(V) if v != planted[i]]
        pairs[i, 0, 0] = others[:2]
    valid = np.ones((n, 1), dtype=bool)
    kp3d, reproj, inliers, mask = O.ransac_pool(P, kp, valid, pairs.astype(np.int64), 5.0)
    out = ops.to_numpy(ops.triangulate_ransac(_cuda(kp), _cuda(P), torch.from_numpy(valid), n_iters=1, pairs=_cuda(pairs)))
    return out, {"inlier_mask": mask, "inliers": inliers, "reproj_mean": reproj}


@pytest.mark.parametrize("delta", [1e-3, -1e-3, 1e-6, -1e-6, 1e-8, -1e-8])
def test_inlier_votes_next_to_the_threshold(ops, delta):
    """utils/triangulation.py:293-300 votes with  0.5 * sqrt(dx^2 + dy^2) < 5 ; the kernel votes division-free with
    dx'^2 + dy'^2 < (10 h)^2 on the homogeneous residuals.  Planted errors of 5 + delta px: down to |delta| = 1e-8 px
    (1e4 times the ~1e-12 px that separates LAPACK's SVD from the Jacobi solve on these rigs) every vote equals the
    oracle's, on both sides of the threshold."""
    rng = np.random.default_rng(int(abs(np.log10(abs(delta)))) * 7 + (delta > 0))
    n, V = 1500, 6
    P, kp, planted = _planted_threshold_pool(n, V, delta, rng)
    out, ref = _single_vote(ops, P, kp, planted)
    assert np.array_equal(out["inlier_mask"].astype(np.uint32), ref["inlier_mask"])
    assert np.array_equal(out["inliers"], ref["inliers"])
    inside = (ref["inlier_mask"][:, 0] >> planted) & 1
    assert inside.all() if delta < 0 else not inside.any()  # the planted view is in (delta < 0) or out (delta > 0)
    np.testing.assert_allclose(out["reproj_mean"], ref["reproj_mean"], rtol=0, atol=REPROJ_ATOL_PX)


def test_inlier_votes_at_the_rounding_limit_are_counted(ops):
    """At |delta| = 1e-12 px the two eigen-solvers legitimately disagree now and then; count it instead of
    pretending: the disagreements must stay confined to the planted view."""
    rng = np.random.default_rng(12)
    n, V = 3000, 6
    P, kp, planted = _planted_threshold_pool(n, V, 1e-12, rng)
    out, ref = _single_vote(ops, P, kp, planted)
    diff = out["inlier_mask"][:, 0].astype(np.uint32) ^ ref["inlier_mask"][:, 0]
    assert ((diff & ~(1 << planted).astype(np.uint32)) == 0).all()
    print("votes that differ from LAPACK at |delta| = 1e-12 px: %d of %d" % (int((diff != 0).sum()), n))


def test_gt_heatmap_renderer(ops):
    """mval_render_gt_heatmaps against dataset/dataset.py:198-207 restated with the reference's own torch expressions: float64
    maps, equal to rounding of exp (<= 2 ulp), incl. joints that project outside the map."""
    from oracle import dataset_oracle as DO

    rng = np.random.default_rng(6)
    pts = rng.uniform(-40, 300, size=(5, 19, 2))  # image pixels, 256 x 256 crop, some outside
    for sigma in (1.0, 2.5):
        exp = np.stack([DO.gt_heatmaps(pts[v], 4, 256, 256, sigma) for v in range(5)])
        got = ops.render_gt_heatmaps(_cuda(pts / 4), 64, 64, sigma)
        assert got.dtype == torch.float64 and tuple(got.shape) == (5, 19, 64, 64) and exp.dtype == np.float64
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=5e-16, atol=1e-300)
        got32 = ops.render_gt_heatmaps(_cuda(pts / 4), 64, 64, sigma, dtype=torch.float32).cpu().numpy()
        np.testing.assert_allclose(got32, exp.astype(np.float32), rtol=2e-7, atol=1e-45)
