"""ctypes binding of libmval_b200.so (include/mval_b200.h).  No CPU fallback: a missing library or device raises."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libmval_b200.so")

MVAL_OK = 0
MVAL_ERR_INVALID_ARGUMENT = -1
MVAL_ERR_UNSUPPORTED = -2
MVAL_ERR_CUDA = -3
MVAL_ERR_NO_DEVICE = -4
MVAL_ERR_OUT_OF_MEMORY = -5
MAX_VIEWS = 32
MAX_SEGMENTS = 64
ABI_VERSION = 5
MAP_SCORE = {None: 0, "HP": 1, "MPE": 2, "BSB": 3}  # MVAL_MAP_SCORE_*


class MvalError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("mval_b200 error %d: %s" % (status, message))
        self.status = status


class RansacParams(C.Structure):
    _fields_ = [("n_iters", C.c_int32), ("epsilon", C.c_double), ("pair_seed", C.c_uint64),
                ("frame_offset", C.c_int64), ("pairs", C.c_void_p), ("frame_keys", C.c_void_p)]


class PipelineOptions(C.Structure):
    _fields_ = [("map_score", C.c_int32), ("use_soft_argmax", C.c_int32), ("use_reprojection_xe", C.c_int32),
                ("direct_optimization", C.c_int32), ("sigma", C.c_double)]


_p, _i, _i64, _f, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64

PROTOTYPES = {
    "mval_version": (C.c_int, []),
    "mval_last_error": (C.c_char_p, []),
    "mval_launch_count": (C.c_uint64, []),
    "mval_check_async": (C.c_int, [_p]),
    "mval_debug_watchdog": (C.c_int, [_u64, _i]),
    "mval_decode_argmax": (C.c_int, [_p, _i64, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "mval_decode_softargmax": (C.c_int, [_p, _i64, _i, _i, _i, _i, _f, _p, _p]),
    "mval_score_hp": (C.c_int, [_p, _i64, _i, _i, _i, _i, _p, _p, _p]),
    "mval_score_peaks": (C.c_int, [_p, _i64, _i, _i, _i, _i, _i, _p, _p, _p]),
    "mval_triangulate_ransac": (C.c_int, [_p, _i, _p, _p, _i64, _i, _i, C.POINTER(RansacParams), _p, _p, _p, _p, _p, _p, _p]),
    "mval_refine_huber": (C.c_int, [_p, _i, _p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p, _p, _p]),
    "mval_score_pool": (C.c_int, [_p, _p, _p, _i64, _i, _i, _i, _i, _i, C.POINTER(RansacParams), _p, _p, _p, _p, _p, _p, _p]),
    "mval_score_pool_scored": (C.c_int, [_p, _p, _p, _i64, _i, _i, _i, _i, _i, C.POINTER(RansacParams), _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mval_score_pool_segments": (C.c_int, [_p, _p, _i, _p, _p, _i, _i, _i, _i, _i, C.POINTER(RansacParams), _i, _p, _p, _p, _p, _p, _p,
                                           _p, _p]),
    "mval_score_pool_host": (C.c_int, [_p, _p, _p, _i64, _i, _i, _i, _i, _i, C.POINTER(RansacParams), _i64, _p, _p, _p, _p, _p, _p]),
    "mval_pipeline_create": (C.c_int, [_i, _i, _i, _i, _i64, _i, C.POINTER(C.c_void_p)]),
    "mval_pipeline_destroy": (C.c_int, [_p]),
    "mval_pipeline_score_pool": (C.c_int, [_p, _p, _p, _p, _i64, _i, C.POINTER(RansacParams), C.POINTER(PipelineOptions), _p, _p, _p,
                                           _p, _p, _p, _p]),
    "mval_score_xe": (C.c_int, [_p, _p, _p, _i64, _i, _i, _i, _i, _d, _p, _p, _p]),
    "mval_topk_desc": (C.c_int, [_p, _i64, _i64, C.c_int32, _p, _p, _p, _p]),
    "mval_topk_merge": (C.c_int, [_p, _p, _i64, C.c_int32, _p, _p, _p, _p]),
    "mval_first_occurrence": (C.c_int, [_p, _p, _i64, _p, _p, _p, _p]),
    "mval_aggregate_map_scores": (C.c_int, [_p, _p, _i64, _i, _i, _i, _i, _i, _p, _p]),
    "mval_sal_rank": (C.c_int, [_p, _p, _p, _i64, _f, C.c_int32, _p, _p, _p]),
    "mval_mkpe": (C.c_int, [_p, _p, _p, _i64, _i, _i, _p, _p]),
    "mval_kmeans_assign": (C.c_int, [_p, _i64, _i, _i, _p, _i, _p, _p, _p]),
    "mval_pose_features": (C.c_int, [_p, _i, _i64, _i, _i, _p, _p]),
    "mval_kcenter_norms": (C.c_int, [_p, _i64, _i, _p, _p]),
    "mval_kcenter_update": (C.c_int, [_p, _p, _i64, _i, _p, _p, _i64, _p, _p, _p]),
    "mval_kcenter_update_batch": (C.c_int, [_p, _p, _i64, _i, _p, _p, _i, _p, _i, _p]),
    "mval_kcenter_tc_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _p]),
    "mval_kcenter_records_bytes": (C.c_size_t, [_i, _i]),
    "mval_kcenter_select": (C.c_int, [_p, _p, _p, _i64, _i, _i64, _i, _p, _p]),
    "mval_kcenter_resolve_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "mval_kcenter_resolve": (C.c_int, [_p, _i, _i, _i, _i, _p, _p, _p, _p, C.POINTER(C.c_int32), _p]),
    "mval_kcenter_resolve_async": (C.c_int, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "mval_kcenter_update_batch_dev": (C.c_int, [_p, _p, _i64, _i, _p, _p, _i, _p, _p, _i, _p]),
    "mval_kcenter_greedy": (C.c_int, [_p, _i64, _i64, _i, C.c_int32, _p, _p, _p]),
    "mval_render_gt_heatmaps": (C.c_int, [_p, _i64, _i, _i, _d, _p, _p, _p]),
    "mval_synth_heatmaps": (C.c_int, [_p, _i64, _i, _i, _f, _f, _u64, _p, _p]),
}

_lib = None


def load():
    """Loads (building first if it is missing and nvcc exists) and returns the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    stale = False
    try:
        stale = _build.needs_build()  # missing, or a source / header is newer than the library (never load a stale ABI)
    except OSError:
        pass
    if stale or os.environ.get("MVAL_REBUILD") == "1":
        try:
            _build.build(force=os.environ.get("MVAL_REBUILD") == "1")
        except Exception:
            if not os.path.isfile(LIB_PATH):
                raise
    if not os.path.isfile(LIB_PATH):
        raise MvalError(MVAL_ERR_NO_DEVICE, "libmval_b200.so is missing and could not be built; there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mval_version() != ABI_VERSION:
        raise MvalError(MVAL_ERR_UNSUPPORTED, "ABI version mismatch")
    _lib = lib
    return lib


def check(status):
    if status != MVAL_OK:
        raise MvalError(status, load().mval_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(load().mval_launch_count())
