"""Tensor-level wrappers over the C ABI (include/mval_b200.h).

torch is used for device memory and streams only: every function takes CUDA tensors, passes raw
pointers + the current stream to libmval_b200.so and returns CUDA tensors.  There is no CPU
implementation here on purpose -- a CPU tensor raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import RansacParams, check

DEFAULT_N_ITERS = 64  # utils/triangulation.py:176
DEFAULT_EPSILON = 5.0  # utils/triangulation.py:177


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _cuda(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: mval_b200 has no CPU fallback, move it to the GPU" % (name, t.device))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _valid_u8(valid, n_frames, J, device):
    """valid: None | [J] | [N, J] of anything truthy -> uint8 CUDA [N, J] (or None)."""
    if valid is None:
        return None
    v = torch.as_tensor(valid)
    v = (v != 0).to(torch.uint8)
    if v.dim() == 1:
        v = v.unsqueeze(0).expand(n_frames, J)
    if tuple(v.shape) != (n_frames, J):
        raise ValueError("valid has shape %s, expected [%d, %d]" % (tuple(v.shape), n_frames, J))
    return v.to(device).contiguous()


def ransac_params(n_iters=DEFAULT_N_ITERS, epsilon=DEFAULT_EPSILON, pair_seed=0, frame_offset=0, pairs=None, frame_keys=None):
    p = RansacParams()
    p.n_iters = int(n_iters)
    p.epsilon = float(epsilon)
    p.pair_seed = int(pair_seed) & ((1 << 64) - 1)
    p.frame_offset = int(frame_offset)
    p.pairs = None if pairs is None else pairs.data_ptr()
    p.frame_keys = None if frame_keys is None else frame_keys.data_ptr()
    return p


def _frame_keys(frame_keys, n_frames, device):
    if frame_keys is None:
        return None
    k = torch.as_tensor(frame_keys).to(device=device, dtype=torch.int64).contiguous().reshape(-1)
    if k.numel() != n_frames:
        raise ValueError("frame_keys must hold one int64 per frame")
    return k


def decode_argmax(heatmaps, stride, valid=None, return_peak=False):
    """heatmaps [N, V, J, H, W] float32 CUDA -> int32 [N, V, J, 2] (x, y)*stride  (utils/evaluation.py:13-30)."""
    hm = _cuda(heatmaps, torch.float32, "heatmaps")
    N, V, J, H, W = hm.shape
    v = _valid_u8(valid, N, J, hm.device)
    out = torch.empty((N, V, J, 2), dtype=torch.int32, device=hm.device)
    peak = torch.empty((N, V, J), dtype=torch.float32, device=hm.device) if return_peak else None
    with torch.cuda.device(hm.device):
        check(_lib.load().mval_decode_argmax(_ptr(hm), N, V, J, H, W, int(stride), _ptr(v), _ptr(out), _ptr(peak), _stream()))
    return (out, peak) if return_peak else out


def decode_softargmax(heatmaps, stride):
    """heatmaps [N, V, J, H, W] -> float32 [N, V, J, 2]  (utils/triangulation.py:191-197)."""
    hm = _cuda(heatmaps, torch.float32, "heatmaps")
    N, V, J, H, W = hm.shape
    out = torch.empty((N, V, J, 2), dtype=torch.float32, device=hm.device)
    with torch.cuda.device(hm.device):
        check(_lib.load().mval_decode_softargmax(_ptr(hm), N, V, J, H, W, float(stride), _ptr(out), _stream()))
    return out


def score_hp(heatmaps, valid=None):
    """heatmaps [N, V, J, H, W] -> float32 [N, V, J]: 1 - max(row-wise softmax)  (strategy.py:1185-1186)."""
    hm = _cuda(heatmaps, torch.float32, "heatmaps")
    N, V, J, H, W = hm.shape
    v = _valid_u8(valid, N, J, hm.device)
    out = torch.empty((N, V, J), dtype=torch.float32, device=hm.device)
    with torch.cuda.device(hm.device):
        check(_lib.load().mval_score_hp(_ptr(hm), N, V, J, H, W, _ptr(v), _ptr(out), _stream()))
    return out


def score_peaks(heatmaps, mode, valid=None):
    """heatmaps [N, V, J, H, W] -> float32 [N, V, J].  mode "MPE": entropy of softmax over the local-peak values
    (strategy.py:1160-1175); mode "BSB": |p0 - p1| of the two best peaks of the row-softmaxed map (:1195-1209)."""
    hm = _cuda(heatmaps, torch.float32, "heatmaps")
    N, V, J, H, W = hm.shape
    v = _valid_u8(valid, N, J, hm.device)
    out = torch.empty((N, V, J), dtype=torch.float32, device=hm.device)
    code = {"MPE": 0, "BSB": 1}[mode]
    with torch.cuda.device(hm.device):
        check(_lib.load().mval_score_peaks(_ptr(hm), N, V, J, H, W, code, _ptr(v), _ptr(out), _stream()))
    return out


def _alloc_tri_outputs(N, J, device):
    return {
        "keypoints_3d": torch.empty((N, J, 3), dtype=torch.float64, device=device),
        "reproj_mean": torch.empty((N, J), dtype=torch.float64, device=device),
        "inliers": torch.empty((N, J), dtype=torch.int32, device=device),
        "metric": torch.empty((N,), dtype=torch.float64, device=device),
        "inlier_count": torch.empty((N,), dtype=torch.int32, device=device),
    }


def triangulate_ransac(keypoints_2d, proj, valid=None, n_iters=DEFAULT_N_ITERS, epsilon=DEFAULT_EPSILON, pair_seed=0,
                       frame_offset=0, pairs=None, direct_optimization=False, frame_keys=None):
    """keypoints_2d [N, V, J, 2] int32/float32 CUDA, proj [N, V, 3, 4] -> dict of CUDA tensors
    (utils/triangulation.py:205-232 for N frames at once).  direct_optimization=True adds the Huber refinement of
    :319-336 on the inlier views (keypoints_3d, reproj_mean and metric are then those of the refined points;
    "refine_iters" [N, J] holds the iteration counts)."""
    kp = keypoints_2d
    if not kp.is_cuda:
        raise RuntimeError("keypoints_2d is on %s: mval_b200 has no CPU fallback" % kp.device)
    is_float = kp.dtype.is_floating_point
    kp = kp.to(torch.float32 if is_float else torch.int32).contiguous()
    N, V, J, _ = kp.shape
    P = _cuda(proj.to(kp.device) if not proj.is_cuda else proj, torch.float64, "proj")
    if tuple(P.shape) != (N, V, 3, 4):
        raise ValueError("proj has shape %s, expected [%d, %d, 3, 4]" % (tuple(P.shape), N, V))
    v = _valid_u8(valid, N, J, kp.device)
    if pairs is not None:
        pairs = _cuda(pairs, torch.uint8, "pairs")
        if tuple(pairs.shape) != (N, J, int(n_iters), 2):
            raise ValueError("pairs must be uint8 [N, J, n_iters, 2]")
    out = _alloc_tri_outputs(N, J, kp.device)
    out["inlier_mask"] = torch.empty((N, J), dtype=torch.int32, device=kp.device)
    fk = _frame_keys(frame_keys, N, kp.device)
    prm = ransac_params(n_iters, epsilon, pair_seed, frame_offset, pairs, fk)
    with torch.cuda.device(kp.device):
        check(_lib.load().mval_triangulate_ransac(_ptr(kp), int(is_float), _ptr(P), _ptr(v), N, V, J, C.byref(prm),
                                                  _ptr(out["keypoints_3d"]), _ptr(out["reproj_mean"]),
                                                  _ptr(out["inliers"]), _ptr(out["inlier_mask"]), _ptr(out["metric"]),
                                                  _ptr(out["inlier_count"]), _stream()))
        if direct_optimization:
            out["refine_iters"] = torch.zeros((N, J), dtype=torch.int32, device=kp.device)
            check(_lib.load().mval_refine_huber(_ptr(kp), int(is_float), _ptr(P), _ptr(v), _ptr(out["inlier_mask"]),
                                                _ptr(out["inliers"]), N, V, J, _ptr(out["keypoints_3d"]),
                                                _ptr(out["reproj_mean"]), _ptr(out["metric"]), _ptr(out["inlier_count"]),
                                                _ptr(out["refine_iters"]), _stream()))
    out["keypoints_2d"] = kp
    return out


def score_pool(heatmaps, proj, stride, valid=None, n_iters=DEFAULT_N_ITERS, epsilon=DEFAULT_EPSILON, pair_seed=0,
               frame_offset=0, return_keypoints_2d=True, map_score=None, frame_keys=None):
    """Device-resident pool scoring: decode (arg-max) + RANSAC triangulation + per-frame uncertainty.
    map_score "HP" / "MPE" / "BSB" additionally returns out["map_score"] float32 [N, V, J] -- the per-map score of
    score_hp / score_peaks, evaluated in the same pass over the heat maps (mval_score_pool_scored)."""
    if map_score not in _lib.MAP_SCORE:
        raise ValueError("map_score must be None, 'HP', 'MPE' or 'BSB'")
    hm = _cuda(heatmaps, torch.float32, "heatmaps")
    N, V, J, H, W = hm.shape
    P = _cuda(proj.to(hm.device) if not proj.is_cuda else proj, torch.float64, "proj")
    if tuple(P.shape) != (N, V, 3, 4):
        raise ValueError("proj has shape %s, expected [%d, %d, 3, 4]" % (tuple(P.shape), N, V))
    v = _valid_u8(valid, N, J, hm.device)
    out = _alloc_tri_outputs(N, J, hm.device)
    xy = torch.empty((N, V, J, 2), dtype=torch.int32, device=hm.device) if return_keypoints_2d else None
    fk = _frame_keys(frame_keys, N, hm.device)
    prm = ransac_params(n_iters, epsilon, pair_seed, frame_offset, frame_keys=fk)
    per_map = torch.empty((N, V, J), dtype=torch.float32, device=hm.device) if map_score is not None else None
    with torch.cuda.device(hm.device):
        check(_lib.load().mval_score_pool_scored(_ptr(hm), _ptr(P), _ptr(v), N, V, J, H, W, int(stride), C.byref(prm),
                                                 _lib.MAP_SCORE[map_score], _ptr(xy), _ptr(out["keypoints_3d"]),
                                                 _ptr(out["reproj_mean"]), _ptr(out["inliers"]), _ptr(out["metric"]),
                                                 _ptr(out["inlier_count"]), _ptr(per_map), _stream()))
    if xy is not None:
        out["keypoints_2d"] = xy
    if per_map is not None:
        out["map_score"] = per_map
    return out


def score_pool_segments(segments, proj, stride, valid=None, n_iters=DEFAULT_N_ITERS, epsilon=DEFAULT_EPSILON, pair_seed=0,
                        frame_offset=0, return_keypoints_2d=False, map_score=None, frame_keys=None, out=None):
    """score_pool over a pool given as a LIST of CUDA heat-map buffers [n_s, V, J, H, W] (successive backbone batches, or
    chunk passes over a resident buffer) in ONE persistent launch (mval_score_pool_segments).  proj [N, V, 3, 4] / valid /
    the outputs cover all N = sum(n_s) frames.  ``out``: a dict from an earlier call whose tensors are reused (no
    allocation in steady state)."""
    if map_score not in _lib.MAP_SCORE:
        raise ValueError("map_score must be None, 'HP', 'MPE' or 'BSB'")
    segs = [_cuda(s, torch.float32, "segment") for s in segments]
    if not 1 <= len(segs) <= _lib.MAX_SEGMENTS:
        raise ValueError("1..%d segments per call" % _lib.MAX_SEGMENTS)
    _, V, J, H, W = segs[0].shape
    counts = [int(s.shape[0]) for s in segs]
    N = sum(counts)
    dev = segs[0].device
    P = _cuda(proj.to(dev) if not proj.is_cuda else proj, torch.float64, "proj")
    if tuple(P.shape) != (N, V, 3, 4):
        raise ValueError("proj has shape %s, expected [%d, %d, 3, 4]" % (tuple(P.shape), N, V))
    v = _valid_u8(valid, N, J, dev)
    if out is None or out["metric"].shape[0] != N:
        out = _alloc_tri_outputs(N, J, dev)
        if return_keypoints_2d:
            out["keypoints_2d"] = torch.empty((N, V, J, 2), dtype=torch.int32, device=dev)
        if map_score is not None:
            out["map_score"] = torch.empty((N, V, J), dtype=torch.float32, device=dev)
    fk = _frame_keys(frame_keys, N, dev)
    prm = ransac_params(n_iters, epsilon, pair_seed, frame_offset, frame_keys=fk)
    ptrs = (C.c_void_p * len(segs))(*[s.data_ptr() for s in segs])
    cnts = (C.c_int64 * len(segs))(*counts)
    with torch.cuda.device(dev):
        check(_lib.load().mval_score_pool_segments(ptrs, cnts, len(segs), _ptr(P), _ptr(v), V, J, H, W, int(stride), C.byref(prm),
                                                   _lib.MAP_SCORE[map_score], _ptr(out.get("keypoints_2d")),
                                                   _ptr(out["keypoints_3d"]), _ptr(out["reproj_mean"]), _ptr(out["inliers"]),
                                                   _ptr(out["metric"]), _ptr(out["inlier_count"]), _ptr(out.get("map_score")),
                                                   _stream()))
    return out


def _host_outputs(N, V, J, pin, soft, map_score):
    out = {
        "keypoints_2d": torch.empty((N, V, J, 2), dtype=torch.float32 if soft else torch.int32, pin_memory=pin),
        "keypoints_3d": torch.empty((N, J, 3), dtype=torch.float64, pin_memory=pin),
        "reproj_mean": torch.empty((N, J), dtype=torch.float64, pin_memory=pin),
        "inliers": torch.empty((N, J), dtype=torch.int32, pin_memory=pin),
        "metric": torch.empty((N,), dtype=torch.float64, pin_memory=pin),
        "inlier_count": torch.empty((N,), dtype=torch.int32, pin_memory=pin),
    }
    if map_score is not None:
        out["map_score"] = torch.empty((N, V, J), dtype=torch.float32, pin_memory=pin)
    return out


class HostPipeline:
    """Handle of the host-buffer streaming pipeline (include/mval_b200.h: mval_pipeline_*): device staging slots and
    streams are allocated once here and reused by every ``score_pool`` call."""

    def __init__(self, V, J, H=64, W=64, chunk_frames=0, n_slots=0, device=None):
        self.shape = (int(V), int(J), int(H), int(W))
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = C.c_void_p(None)
        with torch.cuda.device(self.device):
            check(_lib.load().mval_pipeline_create(*self.shape, int(chunk_frames), int(n_slots), C.byref(self._h)))

    def close(self):
        if self._h is not None and self._h.value:
            _lib.load().mval_pipeline_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def score_pool(self, heatmaps, proj, stride, valid=None, n_iters=DEFAULT_N_ITERS, epsilon=DEFAULT_EPSILON, pair_seed=0,
                   frame_offset=0, map_score=None, use_soft_argmax=False, use_reprojection_xe=False, sigma=None,
                   direct_optimization=False, out=None):
        """Host tensors in (heatmaps float32 [N, V, J, H, W], proj float64 [N, V, 3, 4], valid [N, J] or [J]), host tensors
        out: H2D, kernels and D2H all inside the call (chunked, overlapped).  The keyword flags are those of
        triangulation() / _compute_sal_dict (utils/triangulation.py:168-179, strategy.py:1072-1094)."""
        if heatmaps.is_cuda:
            raise RuntimeError("HostPipeline.score_pool takes host tensors; use score_pool for device-resident heat maps")
        if map_score not in _lib.MAP_SCORE:
            raise ValueError("map_score must be None, 'HP', 'MPE' or 'BSB'")
        hm = heatmaps.to(torch.float32).contiguous()
        N, V, J, H, W = hm.shape
        if (V, J, H, W) != self.shape:
            raise ValueError("this pipeline was created for %s, got %s" % (self.shape, (V, J, H, W)))
        P = proj.to(torch.float64).contiguous()
        if tuple(P.shape) != (N, V, 3, 4):
            raise ValueError("proj has shape %s, expected [%d, %d, 3, 4]" % (tuple(P.shape), N, V))
        v = None
        if valid is not None:
            v = (torch.as_tensor(valid) != 0).to(torch.uint8)
            if v.dim() == 1:
                v = v.unsqueeze(0).expand(N, J)
            v = v.contiguous()
        if out is None:
            out = _host_outputs(N, V, J, hm.is_pinned(), use_soft_argmax, map_score)
        opt = _lib.PipelineOptions(_lib.MAP_SCORE[map_score], int(bool(use_soft_argmax)), int(bool(use_reprojection_xe)),
                                   int(bool(direct_optimization)), float(sigma) if sigma is not None else 0.0)
        prm = ransac_params(n_iters, epsilon, pair_seed, frame_offset)
        with torch.cuda.device(self.device):
            check(_lib.load().mval_pipeline_score_pool(self._h, _ptr(hm), _ptr(P), _ptr(v), N, int(stride), C.byref(prm),
                                                       C.byref(opt), _ptr(out.get("keypoints_2d")), _ptr(out["keypoints_3d"]),
                                                       _ptr(out.get("reproj_mean")), _ptr(out.get("inliers")), _ptr(out["metric"]),
                                                       _ptr(out["inlier_count"]), _ptr(out.get("map_score"))))
        return out


_host_pipelines = {}


def score_pool_host(heatmaps, proj, stride, valid=None, n_iters=DEFAULT_N_ITERS, epsilon=DEFAULT_EPSILON, pair_seed=0,
                    frame_offset=0, chunk_frames=0, out=None, device=None, **flags):
    """End-to-end entry for HOST heat maps (CPU tensors, pinned for overlap): streams the pool through the GPU in
    chunks (H2D, kernels, D2H all inside the call) and returns CPU tensors.  ``flags``: map_score / use_soft_argmax /
    use_reprojection_xe + sigma / direct_optimization as in HostPipeline.score_pool.  The staging buffers live in a
    handle that is kept per (device, shape, chunk) -- nothing is allocated per call."""
    if heatmaps.is_cuda:
        raise RuntimeError("score_pool_host takes host tensors; use score_pool for device-resident heat maps")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _, V, J, H, W = heatmaps.shape
    key = (dev.index, V, J, H, W, int(chunk_frames))
    pipe = _host_pipelines.get(key)
    if pipe is None:
        for k in [k for k in _host_pipelines if k[0] == dev.index]:  # one shape at a time per device: free the old slots
            _host_pipelines.pop(k).close()
        pipe = _host_pipelines[key] = HostPipeline(V, J, H, W, chunk_frames, 0, dev)
    return pipe.score_pool(heatmaps, proj, stride, valid, n_iters, epsilon, pair_seed, frame_offset, out=out, **flags)


def score_xe(heatmaps, proj, keypoints_3d, sigma, return_per_map=False):
    """Reprojection-XE metric (utils/triangulation.py:236-257) for a batch: heatmaps [N, V, J, H, W] float32 CUDA,
    proj [N, V, 3, 4], keypoints_3d [N, J, 3] float64 -> float64 [N] (and the per-map terms [N, V, J])."""
    hm = _cuda(heatmaps, torch.float32, "heatmaps")
    N, V, J, H, W = hm.shape
    P = _cuda(proj.to(hm.device) if not proj.is_cuda else proj, torch.float64, "proj")
    X = _cuda(keypoints_3d, torch.float64, "keypoints_3d")
    if tuple(P.shape) != (N, V, 3, 4) or tuple(X.shape) != (N, J, 3):
        raise ValueError("proj / keypoints_3d shapes do not match the heat maps")
    out = torch.empty((N,), dtype=torch.float64, device=hm.device)
    per_map = torch.empty((N, V, J), dtype=torch.float64, device=hm.device) if return_per_map else None
    with torch.cuda.device(hm.device):
        check(_lib.load().mval_score_xe(_ptr(hm), _ptr(P), _ptr(X), N, V, J, H, W, float(sigma), _ptr(per_map), _ptr(out), _stream()))
    return (out, per_map) if return_per_map else out


def check_async():
    """Synchronises the current stream and raises if a persistent kernel launched on it tripped its mbarrier watchdog
    (include/mval_b200.h: mval_check_async).  Call before results are consumed on the host."""
    check(_lib.load().mval_check_async(_stream()))


def topk_desc(scores, k, index_offset=0, fixed=False, out=None):
    """scores float64 CUDA [n] -> (idx int64 [m], val float64 [m]), m = min(k, #non-NaN): descending score, ties by
    ascending index, NaN dropped (strategy.py:932-949).
    fixed=True: no host synchronisation -- returns (idx [k], val [k], count int32 [1]) with the slots beyond the count set
    to -1 / NaN, the form the cross-rank merge (topk_merge) consumes.  ``out`` = (idx, val, count) buffers to reuse."""
    s = _cuda(scores, torch.float64, "scores").reshape(-1)
    n = s.numel()
    k = int(k) if fixed else int(min(k, n))
    if out is None:
        out = (torch.empty((max(k, 1),), dtype=torch.int64, device=s.device),
               torch.empty((max(k, 1),), dtype=torch.float64, device=s.device),
               torch.empty((1,), dtype=torch.int32, device=s.device))
    idx, val, cnt = out
    with torch.cuda.device(s.device):
        check(_lib.load().mval_topk_desc(_ptr(s), n, int(index_offset), k, _ptr(idx), _ptr(val), _ptr(cnt), _stream()))
    if fixed:
        return idx, val, cnt
    m = int(cnt.item())
    return idx[:m], val[:m]


def topk_merge(scores, indices, k, out=None):
    """Candidates of all ranks (scores float64 [n] with NaN in unused slots, indices int64 [n] global pool indices, gathered
    rank after rank) -> (idx int64 [k], val float64 [k], count int32 [1]) on the device: the global top-k in the
    reference's order (strategy.py:945-949); no host synchronisation."""
    s = _cuda(scores, torch.float64, "scores").reshape(-1)
    g = _cuda(indices, torch.int64, "indices").reshape(-1)
    assert s.numel() == g.numel()
    k = int(k)
    if out is None:
        out = (torch.empty((max(k, 1),), dtype=torch.int64, device=s.device),
               torch.empty((max(k, 1),), dtype=torch.float64, device=s.device),
               torch.empty((1,), dtype=torch.int32, device=s.device))
    with torch.cuda.device(s.device):
        check(_lib.load().mval_topk_merge(_ptr(s), _ptr(g), s.numel(), k, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream()))
    return out


def first_occurrence(pose, frame):
    """Dict-insertion semantics for rows keyed by guid = (pose, frame) (include/mval_b200.h: mval_first_occurrence):
    -> (keep uint8 [n], src int32 [n], unique int32 [1]) CUDA tensors."""
    p = _cuda(pose, torch.int64, "pose").reshape(-1)
    f = _cuda(frame, torch.int64, "frame").reshape(-1)
    n = p.numel()
    keep = torch.empty((n,), dtype=torch.uint8, device=p.device)
    src = torch.empty((n,), dtype=torch.int32, device=p.device)
    unique = torch.empty((1,), dtype=torch.int32, device=p.device)
    with torch.cuda.device(p.device):
        check(_lib.load().mval_first_occurrence(_ptr(p), _ptr(f), n, _ptr(keep), _ptr(src), _ptr(unique), _stream()))
    return keep, src, unique


def aggregate_map_scores(per_map, valid, kind, config, compensated_sum):
    """strategy.py:1151-1158 / 1188-1193 / 1210-1215 for a pool: per_map float32 CUDA [N, V, J], valid [N, J] (anything truthy)
    or None -> float64 CUDA [N] frame scores with the reference's arithmetic (include/mval_b200.h:
    mval_aggregate_map_scores).  kind "HP" / "MPE" / "BSB", config "AVG" / "STD"."""
    pm = _cuda(per_map, torch.float32, "per_map")
    N, V, J = pm.shape
    v = _valid_u8(valid, N, J, pm.device)
    out = torch.empty((N,), dtype=torch.float64, device=pm.device)
    with torch.cuda.device(pm.device):
        check(_lib.load().mval_aggregate_map_scores(_ptr(pm), _ptr(v), N, V, J, _lib.MAP_SCORE[kind], int(config == "STD"),
                                                    int(bool(compensated_sum)), _ptr(out), _stream()))
    return out


def sal_rank(sal_metric, inlier_count, excluded, inlier_threshold, k):
    """strategy.py:957-975 on the device: pool indices (int64 CUDA [m]) of the pseudo-label candidates -- non-NaN
    sal_metric, inlier_count > threshold, not excluded -- in ascending sal_metric order (ties in pool order), first k."""
    m = _cuda(sal_metric, torch.float32, "sal_metric").reshape(-1)
    c = _cuda(inlier_count, torch.float32, "inlier_count").reshape(-1)
    n = m.numel()
    ex = None if excluded is None else _cuda(excluded, torch.uint8, "excluded").reshape(-1)
    k = int(min(k, n))
    idx = torch.empty((max(k, 1),), dtype=torch.int64, device=m.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=m.device)
    with torch.cuda.device(m.device):
        check(_lib.load().mval_sal_rank(_ptr(m), _ptr(c), _ptr(ex), n, float(inlier_threshold), k, _ptr(idx), _ptr(cnt), _stream()))
    return idx[: int(cnt.item())]


def mkpe(pred, gt, valid):
    """utils/evaluation.py:198-208 per frame: pred [N, J, 3], gt [N, R >= 3, J], valid [N, J] (CUDA) -> float32 [N]."""
    p = _cuda(pred, torch.float32, "pred")
    g = _cuda(gt, torch.float32, "gt")
    v = _cuda(valid, torch.float32, "valid")
    N, J, _ = p.shape
    if g.shape[0] != N or g.shape[2] != J or tuple(v.shape) != (N, J):
        raise ValueError("gt / valid shapes do not match pred")
    out = torch.empty((N,), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        check(_lib.load().mval_mkpe(_ptr(p), _ptr(g), _ptr(v), N, J, int(g.shape[1]), _ptr(out), _stream()))
    return out


def kmeans_assign(pred, centres, root):
    """strategy.py:981-989 for a whole pool: pred float32 CUDA [N, J, 3] (sal_dict["pred_3d_keypoints"]), centres float64
    [k, 3 J] (kmeans.cluster_centers_) -> (label int32 [N], margin float64 [N]); see include/mval_b200.h."""
    p = _cuda(pred, torch.float32, "pred")
    N, J, _ = p.shape
    c = _cuda(centres if torch.is_tensor(centres) and centres.is_cuda else torch.as_tensor(np.asarray(centres)).to(p.device),
              torch.float64, "centres")
    if c.dim() != 2 or c.shape[1] != 3 * J:
        raise ValueError("centres has shape %s, expected [k, %d]" % (tuple(c.shape), 3 * J))
    label = torch.empty((N,), dtype=torch.int32, device=p.device)
    margin = torch.empty((N,), dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        check(_lib.load().mval_kmeans_assign(_ptr(p), N, J, int(root), _ptr(c), int(c.shape[0]), _ptr(label), _ptr(margin),
                                             _stream()))
    return label, margin


def pose_features(keypoints_3d, root):
    """utils/coreset.py:35-47 for device-resident poses: [N, J, 3] float64/float32 CUDA -> float32 [N, 3 J]."""
    x = keypoints_3d
    if not x.is_cuda:
        raise RuntimeError("keypoints_3d is on %s: mval_b200 has no CPU fallback" % x.device)
    x = x.contiguous() if x.dtype in (torch.float32, torch.float64) else x.double().contiguous()
    N, J, _ = x.shape
    out = torch.empty((N, 3 * J), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().mval_pose_features(_ptr(x), int(x.dtype == torch.float64), N, J, int(root), _ptr(out), _stream()))
    return out


def kcenter_norms(features):
    f = _cuda(features, torch.float32, "features")
    n, d = f.shape
    out = torch.empty((n,), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        check(_lib.load().mval_kcenter_norms(_ptr(f), n, d, _ptr(out), _stream()))
    return out


def kcenter_update(features, norms, centre, min_dist, index_offset=0, out_best=None):
    """In-place min_dist = min(min_dist, dist(features, centre)); returns (best_val [1] f32, best_idx [1] i64) CUDA."""
    f = _cuda(features, torch.float32, "features")
    n, d = f.shape
    c = _cuda(centre, torch.float32, "centre").reshape(-1)
    assert c.numel() == d and min_dist.dtype == torch.float32 and min_dist.is_contiguous() and min_dist.numel() == n
    if out_best is None:
        out_best = (torch.empty((1,), dtype=torch.float32, device=f.device),
                    torch.empty((1,), dtype=torch.int64, device=f.device))
    with torch.cuda.device(f.device):
        check(_lib.load().mval_kcenter_update(_ptr(f), _ptr(norms), n, d, _ptr(c), _ptr(min_dist), int(index_offset),
                                              _ptr(out_best[0]), _ptr(out_best[1]), _stream()))
    return out_best


def kcenter_greedy(features, n_unlabeled, budget):
    """features float32 CUDA [n, d] (rows >= n_unlabeled are the labeled centres) -> (selected int64 [budget],
    min_dist float32 [n]); the whole greedy loop of utils/coreset.py:83-93 on the device."""
    f = _cuda(features, torch.float32, "features")
    n, d = f.shape
    min_dist = torch.empty((n,), dtype=torch.float32, device=f.device)
    sel = torch.empty((max(int(budget), 1),), dtype=torch.int64, device=f.device)
    with torch.cuda.device(f.device):
        check(_lib.load().mval_kcenter_greedy(_ptr(f), n, int(n_unlabeled), d, int(budget), _ptr(min_dist), _ptr(sel), _stream()))
    return sel[: int(budget)], min_dist


def synth_heatmaps(centres, H=64, W=64, sigma=1.0, noise=0.05, seed=0, out=None):
    """centres float32 CUDA [..., 2] heat-map pixel (x, y) -> float32 [..., H, W]."""
    c = _cuda(centres, torch.float32, "centres")
    n_maps = c.numel() // 2
    if out is None:
        out = torch.empty(tuple(c.shape[:-1]) + (H, W), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        check(_lib.load().mval_synth_heatmaps(_ptr(c), n_maps, H, W, float(sigma), float(noise), int(seed), _ptr(out), _stream()))
    return out


def render_gt_heatmaps(points, H=64, W=64, sigma=1.0, dtype=torch.float64):
    """dataset/dataset.py:198-207: points float64 CUDA [..., 2] (projection / stride, (x, y)) -> [..., H, W] ground-truth heat
    maps, float64 like the reference's (dtype=torch.float32: the same values rounded once)."""
    p = _cuda(points, torch.float64, "points")
    n_maps = p.numel() // 2
    out = torch.empty(tuple(p.shape[:-1]) + (H, W), dtype=dtype, device=p.device)
    with torch.cuda.device(p.device):
        check(_lib.load().mval_render_gt_heatmaps(_ptr(p), n_maps, H, W, float(sigma), _ptr(out if dtype == torch.float64 else None),
                                                  _ptr(out if dtype == torch.float32 else None), _stream()))
    return out


def to_numpy(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def kcenter_update_batch(features, norms, centres, centre_norms, min_dist, flags=0):
    """In place: min_dist[i] = min(min_dist[i], dist(features[i], centres[t]) for every t) in one pass over the features
    (utils/coreset.py:64-69 for a list of cluster centres).  flags: 0 choose, 1 force the FFMA pass, 2 force tcgen05."""
    f = _cuda(features, torch.float32, "features")
    n, d = f.shape
    c = _cuda(centres, torch.float32, "centres").reshape(-1, d)
    cn = _cuda(centre_norms, torch.float32, "centre_norms").reshape(-1)
    assert cn.numel() == c.shape[0] and min_dist.dtype == torch.float32 and min_dist.is_contiguous() and min_dist.numel() == n
    with torch.cuda.device(f.device):
        check(_lib.load().mval_kcenter_update_batch(_ptr(f), _ptr(norms), n, d, _ptr(c), _ptr(cn), c.shape[0], _ptr(min_dist),
                                                    int(flags), _stream()))


def kcenter_tc_stats():
    """(survivors of the last tensor-core screen on the current device, capacity of the survivor list)."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    check(_lib.load().mval_kcenter_tc_stats(C.byref(a), C.byref(b), _stream()))
    return int(a.value), int(b.value)


def kcenter_records_bytes(k_slots, d):
    return int(_lib.load().mval_kcenter_records_bytes(int(k_slots), int(d)))


def kcenter_select(features, norms, min_dist, index_offset, k_slots, out=None):
    """Candidate record block of this shard for one greedy round (include/mval_b200.h:mval_kcenter_select)."""
    n, d = features.shape
    if out is None:
        out = torch.empty(kcenter_records_bytes(k_slots, d), dtype=torch.uint8, device=features.device)
    with torch.cuda.device(features.device):
        check(_lib.load().mval_kcenter_select(_ptr(features), _ptr(norms), _ptr(min_dist), n, d, int(index_offset),
                                              int(k_slots), _ptr(out), _stream()))
    return out


class KcenterResolver:
    """Workspace + outputs of mval_kcenter_resolve for a fixed (n_blocks, k_slots, d)."""

    def __init__(self, n_blocks, k_slots, d, device):
        self.n_blocks, self.k_slots, self.d = int(n_blocks), int(k_slots), int(d)
        kc = self.n_blocks * self.k_slots
        lib = _lib.load()
        self.workspace = torch.empty(int(lib.mval_kcenter_resolve_workspace_bytes(self.n_blocks, self.k_slots, self.d)),
                                     dtype=torch.uint8, device=device)
        self.centres = torch.empty((kc, self.d), dtype=torch.float32, device=device)
        self.centre_norms = torch.empty((kc,), dtype=torch.float32, device=device)
        self._n = C.c_int32(0)

    def resolve(self, records, max_picks, selected_out):
        """records: the n_blocks gathered record blocks (uint8 CUDA).  Writes the picks' global indices to
        selected_out[:T] and returns (T, centres[:T], centre_norms[:T]).  Synchronises the current stream."""
        with torch.cuda.device(records.device):
            check(_lib.load().mval_kcenter_resolve(_ptr(records), self.n_blocks, self.k_slots, self.d, int(max_picks),
                                                   _ptr(self.workspace), _ptr(self.centres), _ptr(self.centre_norms),
                                                   _ptr(selected_out), C.byref(self._n), _stream()))
        t = int(self._n.value)
        return t, self.centres[:t], self.centre_norms[:t]

    def resolve_async(self, records, selected_out, state):
        """The same round without the host in the loop (mval_kcenter_resolve_async): ``state`` int32 CUDA [4] = {picks so
        far, picks of this round, budget, -}; the picks go to selected_out[state[0] ..], nothing is synchronised.  Fold them
        in with kcenter_update_batch_dev(..., self.centres, self.centre_norms, state[1:2], ...)."""
        with torch.cuda.device(records.device):
            check(_lib.load().mval_kcenter_resolve_async(_ptr(records), self.n_blocks, self.k_slots, self.d, _ptr(self.workspace),
                                                         _ptr(self.centres), _ptr(self.centre_norms), _ptr(selected_out),
                                                         _ptr(state), _stream()))


def kcenter_update_batch_dev(features, norms, centres, centre_norms, n_centres, min_dist, flags=0):
    """kcenter_update_batch for a batch whose size only the device knows: the first ``n_centres[0]`` (int32 CUDA [1]) of the
    rows of ``centres`` are folded into min_dist (include/mval_b200.h: mval_kcenter_update_batch_dev)."""
    f = _cuda(features, torch.float32, "features")
    n, d = f.shape
    c = _cuda(centres, torch.float32, "centres").reshape(-1, d)
    cn = _cuda(centre_norms, torch.float32, "centre_norms").reshape(-1)
    assert cn.numel() == c.shape[0] and min_dist.dtype == torch.float32 and min_dist.is_contiguous() and min_dist.numel() == n
    assert n_centres.dtype == torch.int32 and n_centres.is_cuda
    with torch.cuda.device(f.device):
        check(_lib.load().mval_kcenter_update_batch_dev(_ptr(f), _ptr(norms), n, d, _ptr(c), _ptr(cn), c.shape[0], _ptr(n_centres),
                                                        _ptr(min_dist), int(flags), _stream()))
