"""CPU oracle for coreset k-center greedy selection.  TEST INFRASTRUCTURE, NOT PRODUCT.

  stacked_features     <- utils/coreset.py:35-47  (root-relative 3J pose features, unlabeled rows first)
  kcenter_greedy_f64   <- utils/coreset.py:49-95  (update_distances / select_batch) with the very call the
                          reference makes: sklearn.metrics.pairwise_distances(features, centres) in float64
                          (sklearn 1.9.0 in this image; the reference pins no version)
  kcenter_greedy_f32   <- the same greedy loop with the distance evaluated in float32 in a *fixed order*.
                          BASELINE.json asks for bit-exact selected indices "when distances are computed in fp32
                          with the reference's tie-break order"; float32 arithmetic is not associative, so the order
                          is part of the contract and is shared with csrc/kcenter.cu.  Tie-break = np.argmax =
                          lowest index among equal maxima (:90).

Canonical float32 distance between pool row x and centre c (dimension D):
    dot   = fma(x[D-1], c[D-1], ... fma(x[1], c[1], fma(x[0], c[0], +0)) ...)     one accumulator, k ascending, fused
    d2    = ((-2 * dot(x, c)) + dot(x, x)) + dot(c, c)                            sklearn's expansion, in float32
    dist  = sqrt(max(d2, 0)) + 0

Two independent implementations are kept and tested against each other (tests/test_oracle_golden.py):
  * oracle/coreset_oracle.c (C99 fmaf, built by build_c_oracle() into oracle/_coreset_oracle.so) -- the fast one,
    also the CPU arm of bench.py's coreset workload;
  * fma_dot_f32_numpy below: float64 products (exact) + TwoSum + round-to-odd, i.e. an exact emulation of the
    float32 fused multiply-add without any C.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_C_SRC = os.path.join(_HERE, "coreset_oracle.c")
_C_LIB = os.path.join(_HERE, "_coreset_oracle.so")
_lib = None


def build_c_oracle(force=False):
    """gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle/coreset_oracle.c -> oracle/_coreset_oracle.so"""
    if not force and os.path.isfile(_C_LIB) and os.path.getmtime(_C_LIB) >= os.path.getmtime(_C_SRC):
        return _C_LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _C_LIB + ".tmp", _C_SRC, "-lm"]
    subprocess.run(cmd, check=True)
    os.replace(_C_LIB + ".tmp", _C_LIB)
    return _C_LIB


def _c():
    global _lib
    if _lib is None:
        if not os.path.isfile(_C_LIB):
            build_c_oracle()
        lib = ctypes.CDLL(_C_LIB)
        fp, i64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        lib.kc_norms.argtypes = [fp, ctypes.c_int64, ctypes.c_int, fp]
        lib.kc_norms.restype = None
        lib.kc_update.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_int, fp, ctypes.c_float, fp]
        lib.kc_update.restype = None
        lib.kc_dist.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_int, fp, ctypes.c_float, fp]
        lib.kc_dist.restype = None
        lib.kc_greedy.argtypes = [fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, fp, i64p]
        lib.kc_greedy.restype = ctypes.c_int
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


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


# ---------------------------------------------------------------------------------------------- C implementation
def canonical_dot_f32(X, c=None):
    """Row norms |x_i|^2 (c is None or c is X) or dots <x_i, c> for a single row c, canonical order, via C."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, d = X.shape
    out = np.empty(n, dtype=np.float32)
    if c is None or c is X:
        _c().kc_norms(_fp(X), n, d, _fp(out))
        return out
    c = np.ascontiguousarray(c, dtype=np.float32)
    if c.shape == X.shape:
        return fma_dot_f32_numpy(X, c)
    # dot only: dist with xx = cc = 0 is sqrt(max(-2 dot, 0)); go through the numpy emulation instead
    return fma_dot_f32_numpy(X, np.broadcast_to(c, X.shape))


def canonical_dist_f32(X, xx, c, cc=None):
    """Distances of every row of X (row norms ``xx`` precomputed canonically) to the single centre row c."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    xx = np.ascontiguousarray(xx, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    n, d = X.shape
    if cc is None:
        cc = canonical_dot_f32(c[None])[0]
    out = np.empty(n, dtype=np.float32)
    _c().kc_dist(_fp(X), _fp(xx), n, d, _fp(c), ctypes.c_float(float(cc)), _fp(out))
    return out


def kcenter_greedy_f32(features, n_unlabeled, budget):
    """float32 canonical-order variant of the reference greedy loop (C).  Returns (indices, min_d [n] float32)."""
    X = np.ascontiguousarray(features, dtype=np.float32)
    n, d = X.shape
    min_d = np.empty(n, dtype=np.float32)
    sel = np.empty(max(int(budget), 1), dtype=np.int64)
    rc = _c().kc_greedy(_fp(X), n, int(n_unlabeled), d, int(budget), _fp(min_d),
                        sel.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if rc != 0:
        raise ValueError("kcenter_greedy_f32 needs at least one labeled row")
    return [int(i) for i in sel[: int(budget)]], min_d


# ------------------------------------------------------------------------------------ numpy emulation (cross-check)
def _fma_f32(x, c, acc):
    """Exact float32 fma(x, c, acc) for float32 arrays: the product is exact in float64, the sum is formed with
    TwoSum and rounded to odd before the final rounding to float32 (no double-rounding error)."""
    p = x.astype(np.float64) * c.astype(np.float64)
    a = acc.astype(np.float64)
    s = p + a
    bb = s - p
    err = (p - (s - bb)) + (a - bb)
    bits = s.view(np.int64)
    inexact_even = (err != 0) & ((bits & 1) == 0) & np.isfinite(s)
    toward = np.where((err > 0) == (s > 0), 1, -1)  # move away from / toward zero in the integer representation
    toward = np.where(s == 0, 0, toward)
    s = np.where(inexact_even, (bits + toward).view(np.float64), s)
    return s.astype(np.float32)


def fma_dot_f32_numpy(X, C):
    """Row-wise canonical dot of two float32 [n, D] arrays without the C library."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    C = np.ascontiguousarray(np.broadcast_to(np.asarray(C, dtype=np.float32), X.shape))
    acc = np.zeros(X.shape[0], dtype=np.float32)
    for k in range(X.shape[1]):
        acc = _fma_f32(X[:, k], C[:, k], acc)
    return acc


def kcenter_greedy_f32_numpy(features, n_unlabeled, budget):
    """Same greedy loop as kcenter_greedy_f32 on the numpy emulation (small inputs only)."""
    X = np.ascontiguousarray(features, dtype=np.float32)
    n = X.shape[0]
    xx = fma_dot_f32_numpy(X, X)

    def dist(ci):
        dot = fma_dot_f32_numpy(X, X[ci][None])
        d2 = ((np.float32(-2.0) * dot) + xx) + xx[ci]
        return np.sqrt(np.maximum(d2, np.float32(0.0))) + np.float32(0.0)

    min_d = np.full(n, np.inf, dtype=np.float32)
    for ci in range(n_unlabeled, n):
        min_d = np.minimum(min_d, dist(ci))
    picked = []
    for _ in range(budget):
        ind = int(np.argmax(min_d))
        picked.append(ind)
        min_d = np.minimum(min_d, dist(ind))
    return picked, min_d
