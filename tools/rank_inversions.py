#!/usr/bin/env python
"""Ranking agreement of the HP / MPE / BSB strategies with the reference arithmetic on a large pool (GPU box).

north_star: "uncertainty ranking order must be bit-exact".  TRIANGULATION's metric is float64 and its order is compared
exactly in the tests.  HP / MPE / BSB are float32 softmax / exp sums: the reference itself evaluates them with torch on
whatever device it runs on (its CPU and CUDA softmax already differ in the last bit), so there is no bit pattern to hit;
this tool measures what that means for the ORDER: frame scores of N frames from the fused kernel (float32, the reference's
AVG aggregation) against the CPU oracle (torch softmax / the reference's own numpy expressions around the restated peak
finder), then
  * the largest score difference and the smallest gap between neighbouring oracle scores,
  * the number of pair inversions between the two rankings (merge count) and how many of them involve a pair whose oracle
    scores are further apart than the tie band (2 x the largest score difference) -- those would be real errors,
  * how the top-k selections differ.
    python tools/rank_inversions.py [frames] > profiles/r2_rank_inversions.json
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_view_active_learning_b200 import ops  # noqa: E402
from multi_view_active_learning_b200 import strategy as ST  # noqa: E402
from multi_view_active_learning_b200 import synthetic as S  # noqa: E402
from oracle import scores_oracle as SO  # noqa: E402

V, J = 2, 3


_HM = None  # host heat maps, inherited by the forked workers (no pickling of 10 GB)


def _oracle_chunk(args):
    kind, lo, hi = args
    hm = _HM[lo:hi]
    torch.set_num_threads(1)
    if kind == "HP":
        return SO.hp_scores(hm)
    return (SO.mpe_scores if kind == "MPE" else SO.bsb_scores)(hm)


def count_inversions(a):
    """Number of pairs i < j with a[i] > a[j] (merge sort)."""
    a = list(a)
    n = len(a)
    inv = 0
    width = 1
    buf = [0] * n
    while width < n:
        for lo in range(0, n, 2 * width):
            mid, hi = min(lo + width, n), min(lo + 2 * width, n)
            i, j, k = lo, mid, lo
            while i < mid and j < hi:
                if a[i] <= a[j]:
                    buf[k] = a[i]
                    i += 1
                else:
                    buf[k] = a[j]
                    j += 1
                    inv += mid - i
                k += 1
            while i < mid:
                buf[k] = a[i]
                i += 1
                k += 1
            while j < hi:
                buf[k] = a[j]
                j += 1
                k += 1
        a, buf = buf, a
        width *= 2
    return inv


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    topk = 1000
    torch.cuda.set_device(0)
    pool = S.make_pool(n, V, J, seed=2024, p_outlier=0.1)
    centres = torch.from_numpy(pool["centres"]).cuda()
    hm = ops.synth_heatmaps(centres, 64, 64, 1.0, 0.05, 77)
    P = torch.from_numpy(pool["P"]).cuda()
    valid = np.ones((n, J), dtype=bool)
    global _HM
    hm_host = _HM = hm.cpu().numpy()
    out = {"frames": n, "views": V, "joints": J, "topk": topk, "strategies": {}}
    with mp.get_context("fork").Pool(os.cpu_count()) as workers:
        for kind in ("HP", "MPE", "BSB"):
            per_map = ops.score_pool(hm, P, 4, map_score=kind)["map_score"]
            got = ST._aggregate_map_scores(kind, "AVG", per_map.cpu().numpy(), valid).astype(np.float32).astype(np.float64)
            cuts = np.linspace(0, n, os.cpu_count() * 8 + 1).astype(int)
            ref_maps = np.concatenate(workers.map(_oracle_chunk, [(kind, a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]))
            ref = ST._aggregate_map_scores(kind, "AVG", ref_maps.astype(np.float32), valid).astype(np.float32).astype(np.float64)
            ok = ~(np.isnan(got) | np.isnan(ref))
            err = float(np.abs(got[ok] - ref[ok]).max())
            exact = float(np.mean(got[ok] == ref[ok]))
            order_ref = np.lexsort((np.arange(n), -ref))
            order_got = np.lexsort((np.arange(n), -got))
            pos_in_ref = np.empty(n, dtype=np.int64)
            pos_in_ref[order_ref] = np.arange(n)
            inv = count_inversions(pos_in_ref[order_got].tolist())
            # inversions between frames whose oracle scores are further apart than the tie band
            band = 2.0 * err
            seq = ref[order_got]  # oracle scores in the kernel's order: should be non-increasing
            real = 0
            running_min = np.inf
            for s in seq:  # an earlier frame with an oracle score more than `band` BELOW a later one is a real inversion
                if s - running_min > band:
                    real += 1
                running_min = min(running_min, s)
            sorted_ref = np.sort(ref[ok])
            gaps = np.diff(sorted_ref)
            sel_ref, sel_got = set(order_ref[:topk].tolist()), set(order_got[:topk].tolist())
            cut = ref[order_ref[topk - 1]]
            out["strategies"][kind] = {
                "max_abs_score_difference": err, "share_of_scores_equal_bit_for_bit": exact,
                "pair_inversions": inv, "pairs_total": n * (n - 1) // 2,
                "frames_ranked_below_a_frame_they_beat_by_more_than_the_tie_band": real, "tie_band": band,
                "median_gap_between_neighbouring_scores": float(np.median(gaps)), "share_of_neighbour_gaps_inside_tie_band":
                float(np.mean(gaps <= band)),
                "topk_selection_differs_in": len(sel_ref - sel_got),
                "topk_differences_all_within_tie_band_of_the_cut": bool(all(abs(ref[i] - cut) <= band for i in sel_ref ^ sel_got)),
                "nan_scores": int((~ok).sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
