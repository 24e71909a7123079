"""ctypes binding of libmonovo_b200.so (the C ABI in include/monovo_b200.h).

Fails loudly when the CUDA library has not been built: there is no CPU fallback in this package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MVO_B200_LIB: an alternative build of the same library (kernel variants for A/B timing, scripts/lk_variants.py)
LIB_PATH = os.environ.get("MVO_B200_LIB") or os.path.join(HERE, "libmonovo_b200.so")

MVO_OK, MVO_ERR_INVALID, MVO_ERR_CUDA, MVO_ERR_CAPACITY, MVO_ERR_UNSUPPORTED, MVO_ERR_DEGENERATE = 0, -1, -2, -3, -4, -5


class MvoKeypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


class MvoDMatch(C.Structure):
    _fields_ = [("query_idx", C.c_int32), ("train_idx", C.c_int32), ("img_idx", C.c_int32), ("distance", C.c_float)]


class MvoConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_width", C.c_int32), ("max_height", C.c_int32),
                ("nfeatures", C.c_int32), ("batch", C.c_int32), ("max_points", C.c_int32),
                ("ransac_seed", C.c_uint64), ("cuda_stream", C.c_void_p)]


class MvoFrameResult(C.Structure):
    _fields_ = [("n_keypoints", C.c_int32), ("n_matches", C.c_int32), ("n_tracked", C.c_int32),
                ("score_h", C.c_int32), ("score_f", C.c_int32), ("n_inliers_e", C.c_int32),
                ("n_pose_good", C.c_int32), ("n_triangulated", C.c_int32),
                ("R", C.c_double * 9), ("t", C.c_double * 3)]


class MvoGroupConfig(C.Structure):
    _fields_ = [("channels", C.c_int32), ("outputs", C.c_uint32)]


class MvoStreamOutputs(C.Structure):
    _fields_ = [("n_keypoints", C.c_int32), ("n_matches", C.c_int32), ("n_prev", C.c_int32), ("n_tracked", C.c_int32),
                ("keypoints", C.c_void_p), ("descriptors", C.c_void_p), ("matches", C.c_void_p),
                ("track_xy", C.c_void_p), ("track_status", C.c_void_p), ("track_err", C.c_void_p),
                ("mask_h", C.c_void_p), ("mask_f", C.c_void_p), ("mask_e", C.c_void_p), ("mask_pose", C.c_void_p),
                ("X4", C.c_void_p), ("x4_stride", C.c_int64),
                ("H", C.c_double * 9), ("F", C.c_double * 9), ("E", C.c_double * 9),
                ("flags", C.c_int32), ("n_cloud", C.c_int32), ("cloud_xyz", C.c_void_p),
                ("occupied_cells", C.c_int32), ("total_cells", C.c_int32)]


class MvoTrackResult(C.Structure):
    _fields_ = [("n_prev", C.c_int32), ("n_tracked", C.c_int32), ("n_pnp_inliers", C.c_int32), ("pnp_ok", C.c_int32),
                ("rvec", C.c_double * 3), ("tvec", C.c_double * 3)]


MVO_OUT_KEYPOINTS, MVO_OUT_MATCHES, MVO_OUT_TRACKS, MVO_OUT_MODELS, MVO_OUT_POINTS3D, MVO_OUT_CLOUD, MVO_OUT_ALL = 1, 2, 4, 8, 16, 32, 63

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/monovo_b200.h
SIGNATURES = {
    "mvo_create": (C.c_int, [C.POINTER(_vp), C.POINTER(MvoConfig)]),
    "mvo_destroy": (None, [_vp]),
    "mvo_last_error": (C.c_char_p, [_vp]),
    "mvo_version": (C.c_char_p, []),
    "mvo_cuda_stream": (_vp, [_vp]),
    "mvo_batch": (C.c_int, [_vp]),
    "mvo_launch_count": (C.c_uint64, [_vp]),
    "mvo_orb_detect_and_compute": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _i32p]),
    "mvo_orb_compute": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp]),
    "mvo_orb_num_levels": (C.c_int, []),
    "mvo_orb_level_size": (C.c_int, [_vp, C.c_int, _i32p, _i32p]),
    "mvo_orb_get_level": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int]),
    "mvo_lk_get_level": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mvo_orb_get_fast": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _i32p]),
    "mvo_knn_ratio": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_double, _vp, _i32p]),
    "mvo_measure_popc_peak": (C.c_int, [_vp, _f64p]),
    "mvo_knn2": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "mvo_lk_track": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp]),
    "mvo_find_homography": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_double, _vp, _vp, _i32p]),
    "mvo_find_fundamental": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_double, C.c_double, _vp, _vp, _i32p]),
    "mvo_find_essential": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, C.c_double, C.c_double, _vp, _vp, _i32p]),
    "mvo_recover_pose": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _i32p]),
    "mvo_triangulate": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "mvo_solve_pnp_ransac": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp,
                                       _vp, _i32p]),
    "mvo_rodrigues": (C.c_int, [_vp, _vp]),
    "mvo_pnp_get_hypotheses": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "mvo_score_hypotheses": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_double, C.c_int, _vp, _vp, _vp]),
    "mvo_group_step": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "mvo_group_submit": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "mvo_group_collect": (C.c_int, [_vp, _vp]),
    "mvo_group_reset": (C.c_int, [_vp]),
    "mvo_group_configure": (C.c_int, [_vp, C.POINTER(MvoGroupConfig)]),
    "mvo_group_output_bytes": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "mvo_group_outputs": (C.c_int, [_vp, C.c_int, C.POINTER(MvoStreamOutputs)]),
    "mvo_group_set_tracks": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int]),
    "mvo_group_track": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "mvo_group_get_tracks": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _i32p, _i32p]),
    "mvo_cache_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "mvo_graph_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "mvo_set_occupancy_grid": (C.c_int, [_vp, C.c_int]),
    "mvo_orb_occupancy": (C.c_int, [_vp, C.c_int, _i32p, _i32p]),
    "mvo_pack_pointcloud": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int]),
    "mvo_stage_ms": (C.c_int, [_vp, C.c_char_p, _f32p]),
    "mvo_stage_span_ms": (C.c_int, [_vp, C.c_char_p, _f32p, _f32p]),
    "mvo_debug_set": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "mvo_debug_time": (C.c_int, [_vp, C.c_char_p, C.c_int, _f32p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ros2_mono_vo_b200.build` "
            "(ros2_mono_vo_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
