"""Pins the CPU oracle against outputs of the unmodified reference (tests/golden, made by oracle/make_golden.py)."""
import numpy as np
import pytest

from multi_view_active_learning_b200 import synthetic as S
from oracle import coreset_oracle as C
from oracle import scores_oracle as SC
from oracle import triangulation_oracle as O

POOLS = ["pool_v5_j19", "pool_v8_j19", "pool_v20_j42", "pool_v31_j19", "pool_v2_j3"]


def test_reference_unit_test_known_answer(golden):
    g = golden("ref_unit_triangulation")
    hm = np.zeros((8, 19, 64, 64), dtype=np.float32)
    for r, c, v in g["bump_rc"]:
        hm[:, :, int(r), int(c)] = v
    out = O.triangulate_pool(hm[None], g["P"][None].astype(np.float64), int(g["stride"]), g["valid"][None])
    assert np.array_equal(out["keypoints_2d"][0], g["keypoints_2d"])
    assert (out["keypoints_2d"] == 88).all()
    assert int(out["inlier_count"][0]) == int(g["inlier_count"]) == 3
    # values the survey recorded from the reference
    np.testing.assert_allclose(out["keypoints_3d"][0, 0], [-37.2230772535, -125.6428366493, -23.8942096679], rtol=1e-10)
    np.testing.assert_allclose(out["metric"][0], 3.1619149383944163, rtol=1e-13)
    np.testing.assert_allclose(out["keypoints_3d"][0], g["keypoints_3d"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(out["metric"][0], g["metric"], rtol=1e-13)


@pytest.mark.parametrize("name", POOLS)
def test_ransac_pool_matches_reference(golden, name):
    g = golden(name)
    kp = g["keypoints_2d"]
    V = kp.shape[1]
    out = O.triangulate_pool(None, g["P"], int(g["stride"]), g["valid"], pair_seed=int(g["pair_seed"]), keypoints_2d=kp)
    assert np.array_equal(out["inlier_count"], g["inlier_count"])
    np.testing.assert_allclose(out["keypoints_3d"], g["keypoints_3d"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(out["metric"], g["metric"], rtol=1e-12)
    # invalid joints stay exactly zero (reference :206-211)
    assert (out["keypoints_3d"][~g["valid"]] == 0).all()
    # decode of the rebuilt one-hot heatmaps reproduces the reference's keypoints_2d (incl. invalid -> [0,0])
    hm = S.onehot_heatmaps(g["keypoints_2d_unmasked"][:2], int(g["stride"]))
    assert np.array_equal(O.decode_argmax(hm, int(g["stride"]), g["valid"][:2]), kp[:2])
    assert V == g["P"].shape[1]


def test_pair_tables():
    assert O.pair_table(8, 64).tolist() == [list(p) for p in __import__("itertools").combinations(range(8), 2)]
    t = O.pair_table(20, 64, seed=5, frame_index=3, joint=2)
    assert t.shape == (64, 2) and len({tuple(p) for p in t.tolist()}) == 64 and (t[:, 0] < t[:, 1]).all()
    assert not np.array_equal(t, O.pair_table(20, 64, seed=5, frame_index=3, joint=3))
    # the shim hands the reference the same subset, in the same order
    lst = list(__import__("itertools").combinations(range(20), 2))
    O.DeterministicShuffle(5, [(3, 2)]).shuffle(lst)
    assert [list(p) for p in lst[:64]] == t.tolist() and len(set(lst)) == 190


def test_decode_edge_cases(golden):
    g = golden("decode_edge_cases")
    hm, stride = g["heatmaps"], int(g["stride"])
    assert np.array_equal(O.decode_argmax(hm, stride, g["valid"]), g["scaled"])
    assert np.array_equal(O.decode_argmax(hm, stride), g["scaled_all_valid"])
    # get_pred_coordinates (utils/evaluation.py:46-57): same argmax scaled by bbox extent / map size
    flat = O.decode_argmax(hm, 1)
    boxes = g["boxes"].astype(np.float64)
    sx = (boxes[:, 3] - boxes[:, 1]) / 64.0
    sy = (boxes[:, 2] - boxes[:, 0]) / 64.0
    exp = np.stack([flat[..., 0] * sx[:, None], flat[..., 1] * sy[:, None]], axis=-1)
    np.testing.assert_allclose(exp, g["pred_coordinates"], rtol=1e-6)


def test_hp_scores(golden):
    g = golden("hp_scores")
    s = SC.hp_scores(g["heatmaps"])
    np.testing.assert_allclose(s, g["hp_per_map"], atol=2e-6)
    np.testing.assert_allclose(SC.hp_metric(g["heatmaps"], g["valid"], "AVG"), float(g["hp_avg"]), atol=2e-6)
    np.testing.assert_allclose(SC.hp_metric(g["heatmaps"], g["valid"], "STD"), float(g["hp_std"]), atol=2e-6)


def test_softargmax_matches_reference_with_vendored_kornia(golden):
    """Soft-arg-max path: the unmodified reference's triangulation(use_soft_argmax=True) with kornia's
    spatial_expectation2d / create_meshgrid as vendored verbatim by `transformers` standing in for the missing package
    (oracle/make_golden.py:case_softargmax).  The float64 restatement agrees with the reference's float32 key-points to
    float32 rounding; from the reference's own key-points the RANSAC restatement reproduces 3-D joints, metric and
    inlier counts."""
    g = golden("softargmax_v5_j6")
    hm = S.render_heatmaps(g["centres"], noise=float(g["noise"]), seed=int(g["heatmap_seed"])) * np.float32(g["gain"])
    kp = O.decode_softargmax(hm, int(g["stride"]))
    assert kp.dtype == np.float32 and g["keypoints_2d"].dtype == np.float32
    np.testing.assert_allclose(kp, g["keypoints_2d"], rtol=0, atol=1e-4)  # px; observed 4.6e-5
    out = O.triangulate_pool(None, g["P"], int(g["stride"]), g["valid"], keypoints_2d=g["keypoints_2d"])
    assert np.array_equal(out["inlier_count"], g["inlier_count"])
    np.testing.assert_allclose(out["keypoints_3d"], g["keypoints_3d"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(out["metric"], g["metric"], rtol=1e-12)
    own = O.triangulate_pool(hm, g["P"], int(g["stride"]), g["valid"], use_soft_argmax=True)
    assert np.array_equal(own["inlier_count"], g["inlier_count"])
    np.testing.assert_allclose(own["keypoints_3d"], g["keypoints_3d"], rtol=1e-3, atol=1e-2)


def test_peak_scores_match_reference(golden):
    """MPE / BSB: the oracle against the unmodified reference's _compute_mpes / _compute_mpe / _compute_bsb
    (strategy.py:1149-1176, 1195-1215) run with the restated peak finder standing in for skimage (oracle/make_golden.py:
    case_peaks) -- bit for bit, per map and per frame, AVG and STD."""
    g = golden("peak_scores")
    hm, valid = g["heatmaps"], g["valid"].astype(bool)
    mpe, bsb = SC.mpe_scores(hm), SC.bsb_scores(hm)
    ents = [mpe[v, k] for v in range(hm.shape[0]) for k in range(hm.shape[1]) if valid[k]]
    assert np.array_equal(np.asarray(ents, dtype=np.float64), g["mpe_per_map"])
    assert len(ents) == 6 and min(ents) > 1.0  # several peaks per map: the entropy is far from 0
    for cfg in ("AVG", "STD"):
        assert float(SC.reduce_frame_score(mpe, valid, cfg, "MPE")) == float(g["mpe_" + cfg])
        assert float(SC.reduce_frame_score(bsb, valid, cfg, "BSB")) == float(g["bsb_" + cfg])
    assert isinstance(SC.reduce_frame_score(mpe, valid, "AVG", "MPE"), np.float32)


def test_coreset_matches_reference(golden):
    g = golden("coreset_random")
    F = C.stacked_features(g["sal_poses"], g["al_poses"], int(g["root"]))
    assert np.array_equal(F, g["features"])
    n = len(g["sal_poses"])
    picked, min_d = C.kcenter_greedy_f64(F, n, int(g["budget"]))
    assert picked == g["picked"].tolist()
    assert np.array_equal(min_d, g["min_distances"])
    # the float32 canonical-order variant selects the same rows on this well separated pool
    picked32, min32 = C.kcenter_greedy_f32(F, n, int(g["budget"]))
    assert picked32 == picked
    np.testing.assert_allclose(min32, min_d[:, 0], rtol=1e-4, atol=1e-2)


def test_coreset_reference_unit_test(golden):
    g = golden("ref_unit_coreset")
    pose = [[0, 1, 2] for _ in range(int(g["n_joints"]))]
    F = C.stacked_features([pose] * int(g["n_sal"]), [pose] * int(g["n_al"]), int(g["root"]))
    assert F.shape == (25, 57)
    picked, _ = C.kcenter_greedy_f64(F, 20, 5)
    assert picked == g["picked"].tolist() == [0] * 5
    assert C.kcenter_greedy_f32(F, 20, 5)[0] == [0] * 5


def test_canonical_dot_is_order_defined():
    rng = np.random.default_rng(0)
    for d in (3, 57, 126, 128, 300, 2048):
        X = rng.normal(size=(5, d)).astype(np.float32)
        c = rng.normal(size=d).astype(np.float32)
        got = C.canonical_dot_f32(X, c)
        np.testing.assert_allclose(got, X.astype(np.float64) @ c.astype(np.float64), rtol=2e-4, atol=1e-4)
        assert got.dtype == np.float32
    x = rng.normal(size=(1, 2048)).astype(np.float32)
    xx = C.canonical_dot_f32(x)
    assert C.canonical_dist_f32(x, xx, x[0])[0] == 0.0  # a row is at distance exactly 0 from itself


def test_c_oracle_equals_numpy_fma_emulation():
    """Two independent statements of the canonical float32 arithmetic (C99 fmaf vs float64 + TwoSum + round-to-odd)
    give the same bits: norms, distances, selected rows, final minima."""
    rng = np.random.default_rng(1)
    for n, L, d, b in [(300, 5, 57, 20), (120, 3, 126, 12), (64, 2, 2048, 6), (50, 1, 7, 8)]:
        F = (rng.standard_normal((n + L, d)) * rng.choice([1e-2, 1.0, 300.0])).astype(np.float32)
        F[7] = F[2]
        assert np.array_equal(C.canonical_dot_f32(F), C.fma_dot_f32_numpy(F, F))
        s1, m1 = C.kcenter_greedy_f32(F, n, b)
        s2, m2 = C.kcenter_greedy_f32_numpy(F, n, b)
        assert s1 == s2 and np.array_equal(m1, m2)
    # the emulated fma itself on operands of very different magnitude (where double rounding would bite)
    x = rng.standard_normal(200000).astype(np.float32)
    c = rng.standard_normal(200000).astype(np.float32)
    a = (rng.standard_normal(200000) * rng.choice([1e-6, 1e-3, 1.0, 1e3, 1e6], 200000)).astype(np.float32)
    got = C._fma_f32(x, c, a)
    import fractions

    for i in range(0, 200000, 997):
        exact = fractions.Fraction(float(x[i])) * fractions.Fraction(float(c[i])) + fractions.Fraction(float(a[i]))
        lo, hi = np.nextafter(got[i], np.float32(-np.inf)), np.nextafter(got[i], np.float32(np.inf))
        err = abs(fractions.Fraction(float(got[i])) - exact)
        assert err <= abs(fractions.Fraction(float(lo)) - exact) and err <= abs(fractions.Fraction(float(hi)) - exact)


def test_xe_oracle_matches_reference(golden):
    """oracle compute_xe against the values the reference's _compute_xe produced (oracle/make_golden.py:case_xe)."""
    from multi_view_active_learning_b200 import synthetic as S
    from oracle import triangulation_oracle as O

    g = golden("xe_metric")
    hm = S.render_heatmaps(g["centres"], noise=float(g["noise"]), seed=int(g["heatmap_seed"]))
    m, per_map = O.compute_xe(g["keypoints_3d"], g["P"], hm, float(g["sigma"]))
    np.testing.assert_allclose(m, g["xe"], rtol=1e-12, atol=0)
    assert per_map.shape == (6, 4, 5) and (per_map > 0).all()


def test_ranking():
    d = {"a": 1.0, "b": float("nan"), "c": 3.0, "d": 3.0, "e": 2.0}
    assert SC.rank_nlargest(d, 3) == ["c", "d", "e"]
    assert SC.rank_nlargest(d, 10) == ["c", "d", "e", "a"]


def test_huber_refinement_matches_reference(golden):
    """direct_optimization=True (utils/triangulation.py:319-336): the oracle makes the reference's own scipy call, so it
    reproduces the reference's refined joints to rounding; the refinement really moves the points (several mm)."""
    g = golden("huber_v8_j19")
    kp = np.where(g["valid"][:, None, :, None], g["keypoints_2d_unmasked"], 0)
    out = O.triangulate_pool(None, g["P"], int(g["stride"]), g["valid"], keypoints_2d=kp, direct_optimization=True)
    np.testing.assert_allclose(out["keypoints_3d"], g["keypoints_3d"], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(out["metric"], g["metric"], rtol=1e-9)
    assert np.array_equal(out["inlier_count"], g["inlier_count"])
    assert np.abs(g["keypoints_3d"] - g["keypoints_3d_dlt"]).max() > 1.0


def test_zero_padding_leaves_the_canonical_coreset_arithmetic_unchanged():
    """pool.pad_features: zero columns close the fma chain with fma(0, 0, acc) = acc, so the float32 selection and the
    running minima of the C oracle are bit-identical with and without them (what makes 57 -> 64 padding legal)."""
    import torch

    from multi_view_active_learning_b200 import pool as P

    rng = np.random.default_rng(4)
    F = (rng.normal(size=(600, 57)) * 100).astype(np.float32)
    Fp = P.pad_features(torch.from_numpy(F), 64).numpy()
    assert Fp.shape == (600, 64) and np.array_equal(Fp[:, :57], F) and not Fp[:, 57:].any()
    assert P.pad_features(torch.from_numpy(Fp), 32) is not None and P.pad_features(torch.from_numpy(Fp), 32).shape == (600, 64)
    sel, mins = C.kcenter_greedy_f32(F, 560, 30)
    sel_p, mins_p = C.kcenter_greedy_f32(Fp, 560, 30)
    assert sel == sel_p and np.array_equal(mins, mins_p)
    assert np.array_equal(C.canonical_dot_f32(F), C.canonical_dot_f32(Fp))


def test_coreset_float32_contract_against_float64_on_near_ties():
    """The contract of the coreset arithmetic (north star: "bit-exact when distances are computed in fp32 with the reference's
    tie-break order"): on pools without near-ties the float32 canonical selection equals the
    reference's float64 selection; when two candidates' distances differ by less than float32 can resolve, float32 sees a
    tie and takes the LOWER index (np.argmax's rule on equal values), float64 the farther one."""
    from oracle import coreset_oracle as CO

    rng = np.random.default_rng(23)
    # (1) generic pools: identical selections
    for n, L, d in ((300, 10, 57), (200, 5, 126)):
        F = rng.normal(size=(n + L, d)) * 300
        a, _ = CO.kcenter_greedy_f32(F, n, 25)
        b, _ = CO.kcenter_greedy_f64(F, n, 25)
        assert a == b
    # (2) an adversarial near-tie: rows 0 and 1 at distances r and r * (1 + 1e-9) from the single labeled centre
    d = 57
    u = rng.normal(size=d)
    u /= np.linalg.norm(u)
    w = rng.normal(size=d)
    w -= w.dot(u) * u
    w /= np.linalg.norm(w)
    centre = np.zeros(d)
    F = np.stack([1000.0 * u, 1000.0 * (1 + 1e-9) * w, 10.0 * u, centre])
    pick64, _ = CO.kcenter_greedy_f64(F, 3, 1)
    pick32, _ = CO.kcenter_greedy_f32(F, 3, 1)
    assert pick64 == [1]  # the farther row, by 1e-6 units
    assert pick32 == [0]  # float32 cannot tell them apart: first index wins
