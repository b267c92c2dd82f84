"""CPU oracle for coreset k-center greedy selection.  TEST INFRASTRUCTURE, NOT PRODUCT.

  stacked_features     <- utils/coreset.py:35-47  (root-relative 3J pose features, unlabeled rows first)
  kcenter_greedy_f64   <- utils/coreset.py:49-95  (update_distances / select_batch) with the very call the
                          reference makes: sklearn.metrics.pairwise_distances(features, centres) in float64
                          (sklearn 1.9.0 in this image; the reference pins no version)
  kcenter_greedy_f32   <- the same greedy loop with the distance evaluated in float32 in a *fixed summation
                          order* (``canonical_dot_f32``).  BASELINE.json asks for bit-exact selected indices
                          "when distances are computed in fp32 with the reference's tie-break order"; float32
                          addition is not associative, so the order is part of the contract and is shared with
                          csrc/kcenter.cu.  Tie-break = np.argmax = lowest index among equal maxima (:90).

Canonical float32 squared distance between rows x and c of dimension D (zero-padded to Dp = 128*K):
    element e lives in (chunk k, lane l, slot s) = (e // 128, (e % 128) // 4, e % 4)
    acc[l][s] = (((0 + x*c at k=0) + x*c at k=1) + ...)           # separate IEEE multiply and add, no FMA
    lane[l]   = (acc[l][0] + acc[l][1]) + (acc[l][2] + acc[l][3])
    butterfly: for m in 16, 8, 4, 2, 1: lane[l] = lane[l] + lane[l ^ m]   -> dot = lane[0]
    d2 = ((-2 * dot(x, c)) + dot(x, x)) + dot(c, c);  d = sqrt(max(d2, 0))   # sklearn's expansion, in float32
"""
import numpy as np

LANES, SLOTS = 32, 4
CHUNK = LANES * SLOTS


def stacked_features(sal_poses, al_poses, root_idx):
    """sal_poses: iterable of [J][>=3] predicted poses (dict order), al_poses: iterable of [J][>=3] labeled
    poses.  Row = (pose^T[0:3] - pose^T[0:3, root]).flatten(), i.e. x_0..x_J-1, y_0.., z_0.. (reference :40-47)."""
    rows = []
    for pose in list(sal_poses) + list(al_poses):
        p = np.array(pose).transpose([1, 0])[0:3, :]
        rows.append((p - p[:, root_idx:root_idx + 1]).flatten())
    return np.stack(rows)


def kcenter_greedy_f64(features, n_unlabeled, budget):
    """Reference utils/coreset.py:71-95 on ``features`` whose rows >= n_unlabeled are the labeled centres.
    Returns (selected row indices in order, final min_distances [n,1])."""
    from sklearn.metrics import pairwise_distances

    feats = np.asarray(features)
    labeled = list(range(n_unlabeled, feats.shape[0]))
    min_d = None
    if labeled:
        min_d = np.min(pairwise_distances(feats, feats[labeled], metric="euclidean"), axis=1).reshape(-1, 1)
    picked = []
    for _ in range(budget):
        ind = int(np.argmax(min_d))
        assert ind not in labeled
        d = pairwise_distances(feats, feats[[ind]], metric="euclidean")
        min_d = d if min_d is None else np.minimum(min_d, d)
        picked.append(ind)
    return picked, min_d


def _pad(X):
    X = np.ascontiguousarray(X, dtype=np.float32)
    d = X.shape[-1]
    dp = (d + CHUNK - 1) // CHUNK * CHUNK
    if dp != d:
        X = np.concatenate([X, np.zeros(X.shape[:-1] + (dp - d,), dtype=np.float32)], axis=-1)
    return X


def canonical_dot_f32(X, c):
    """X [n, D] float32, c [D] or [n, D] float32 -> [n] float32 in the canonical summation order."""
    X = _pad(X)
    c = _pad(np.broadcast_to(np.asarray(c, dtype=np.float32), X.shape[:-1] + (np.shape(c)[-1],)))
    n, dp = X.shape
    k = dp // CHUNK
    prod = (X * c).reshape(n, k, LANES, SLOTS)
    acc = np.zeros((n, LANES, SLOTS), dtype=np.float32)
    for i in range(k):
        acc = acc + prod[:, i]
    lane = (acc[..., 0] + acc[..., 1]) + (acc[..., 2] + acc[..., 3])
    idx = np.arange(LANES)
    for m in (16, 8, 4, 2, 1):
        lane = lane + lane[:, idx ^ m]
    return lane[:, 0]


def canonical_dist_f32(X, xx, c):
    """Distances of every row of X (row norms ``xx`` precomputed canonically) to the single row c."""
    c = np.asarray(c, dtype=np.float32)
    cc = canonical_dot_f32(c[None], c)[0]
    dot = canonical_dot_f32(X, c)
    d2 = ((np.float32(-2.0) * dot) + xx) + cc
    return np.sqrt(np.maximum(d2, np.float32(0.0)))


def kcenter_greedy_f32(features, n_unlabeled, budget):
    """float32 canonical-order variant of the reference greedy loop.  Returns (indices, min_d [n] float32)."""
    X = np.ascontiguousarray(features, dtype=np.float32)
    n = X.shape[0]
    xx = canonical_dot_f32(X, X)
    min_d = np.full(n, np.inf, dtype=np.float32)
    for ci in range(n_unlabeled, n):
        min_d = np.minimum(min_d, canonical_dist_f32(X, xx, X[ci]))
    picked = []
    for _ in range(budget):
        ind = int(np.argmax(min_d))
        picked.append(ind)
        min_d = np.minimum(min_d, canonical_dist_f32(X, xx, X[ind]))
    return picked, min_d
