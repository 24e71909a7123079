"""Python mirror of the reference's hot-path interface on top of the C ABI (tests + bench harness).

FeatureProcessor keeps the reference's method names and argument meaning
(/root/reference/include/mono_vo/feature_processor.hpp:14-31):
    detect(image) -> keypoints
    detect_and_compute(image) -> (keypoints, descriptors)
    find_matches(descriptors1, descriptors2, lowes_distance_ratio) -> matches
Everything runs through libmonovo_b200.so; a missing library or GPU raises (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

KP_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("size", np.float32), ("angle", np.float32),
                     ("response", np.float32), ("octave", np.int32), ("class_id", np.int32)])
DMATCH_DTYPE = np.dtype([("query_idx", np.int32), ("train_idx", np.int32), ("img_idx", np.int32),
                         ("distance", np.float32)])
assert KP_DTYPE.itemsize == C.sizeof(_lib.MvoKeypoint) == 28
assert DMATCH_DTYPE.itemsize == C.sizeof(_lib.MvoDMatch) == 16


class MvoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"monovo_b200 error {code}: {msg}")
        self.code = code


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One stream group (mvo_ctx): `batch` camera streams in lock step on one GPU."""

    def __init__(self, max_width: int, max_height: int, nfeatures: int = 1000, batch: int = 1, device: int = 0,
                 max_points: int = 0, ransac_seed: int = 0, cuda_stream: int | None = None):
        self.lib = _lib.load()
        cfg = _lib.MvoConfig(device, max_width, max_height, nfeatures, batch, max_points, ransac_seed,
                             C.c_void_p(cuda_stream) if cuda_stream else None)
        h = C.c_void_p()
        rc = self.lib.mvo_create(C.byref(h), C.byref(cfg))
        if rc != 0:
            raise MvoError(rc, (self.lib.mvo_last_error(None) or b"").decode())
        self.h = h
        self.nfeatures = nfeatures
        self.batch = batch
        self.kp_cap = nfeatures + nfeatures // 4 + 64

    def close(self):
        if getattr(self, "h", None):
            self.lib.mvo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise MvoError(rc, (self.lib.mvo_last_error(self.h) or b"").decode())

    @property
    def launch_count(self) -> int:
        return int(self.lib.mvo_launch_count(self.h))

    @property
    def cuda_stream(self) -> int:
        return int(self.lib.mvo_cuda_stream(self.h) or 0)

    # ---- ORB ---------------------------------------------------------------------------------
    def orb_detect_and_compute(self, img: np.ndarray, want_desc: bool = True):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        kps = np.zeros(self.kp_cap, KP_DTYPE)
        desc = np.zeros((self.kp_cap, 32), np.uint8) if want_desc else None
        n = C.c_int32(0)
        self._check(self.lib.mvo_orb_detect_and_compute(self.h, _ptr(img), w, h, img.strides[0], ch, _ptr(kps),
                                                        _ptr(desc) if want_desc else None, self.kp_cap, C.byref(n)))
        return kps[:n.value].copy(), (desc[:n.value].copy() if want_desc else None)

    def orb_compute(self, img: np.ndarray, kps: np.ndarray):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        desc = np.zeros((len(kps), 32), np.uint8)
        valid = np.zeros(len(kps), np.uint8)
        self._check(self.lib.mvo_orb_compute(self.h, _ptr(img), w, h, img.strides[0], ch, _ptr(kps), len(kps),
                                             _ptr(desc), _ptr(valid)))
        return desc, valid.astype(bool)

    def orb_level(self, level: int, blurred: bool = False) -> np.ndarray:
        w, h = C.c_int32(), C.c_int32()
        self._check(self.lib.mvo_orb_level_size(self.h, level, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        self._check(self.lib.mvo_orb_get_level(self.h, level, int(blurred), _ptr(out), out.strides[0]))
        return out

    def orb_fast(self, level: int):
        cap = 1 << 20
        xy = np.zeros(cap, np.uint32)
        sc = np.zeros(cap, np.int32)
        n = C.c_int32()
        self._check(self.lib.mvo_orb_get_fast(self.h, level, _ptr(xy), _ptr(sc), cap, C.byref(n)))
        xy, sc = xy[:n.value], sc[:n.value]
        return (xy & 0xffff).astype(np.int32), (xy >> 16).astype(np.int32), sc.copy()

    # ---- LK ----------------------------------------------------------------------------------
    def lk_track(self, prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray):
        prev = np.ascontiguousarray(prev, np.uint8)
        nxt = np.ascontiguousarray(nxt, np.uint8)
        assert prev.shape == nxt.shape
        h, w = prev.shape[:2]
        ch = 1 if prev.ndim == 2 else prev.shape[2]
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = len(pts)
        out = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32)
        self._check(self.lib.mvo_lk_track(self.h, _ptr(prev), _ptr(nxt), w, h, prev.strides[0], ch, _ptr(pts), n,
                                          _ptr(out), _ptr(status), _ptr(err)))
        return out, status, err

    # ---- kNN ---------------------------------------------------------------------------------
    def knn_ratio(self, q: np.ndarray, t: np.ndarray, ratio: float) -> np.ndarray:
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        out = np.zeros(max(len(q), 1), DMATCH_DTYPE)
        n = C.c_int32()
        self._check(self.lib.mvo_knn_ratio(self.h, _ptr(q), len(q), _ptr(t), len(t), float(ratio), _ptr(out),
                                           C.byref(n)))
        return out[:n.value].copy()

    def knn2(self, q: np.ndarray, t: np.ndarray):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        idx = np.full((len(q), 2), -1, np.int32)
        dist = np.full((len(q), 2), -1, np.int32)
        self._check(self.lib.mvo_knn2(self.h, _ptr(q), len(q), _ptr(t), len(t), _ptr(idx), _ptr(dist)))
        return idx, dist


def calc_optical_flow_pyr_lk(ctx: "Context", prev_img, next_img, prev_pts):
    """Mirror of cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err) with default arguments
    (reference call site src/tracker.cpp:68-69)."""
    return ctx.lk_track(prev_img, next_img, prev_pts)


class FeatureProcessor:
    """Mirror of mono_vo::FeatureProcessor (reference include/mono_vo/feature_processor.hpp:14-31)."""

    def __init__(self, num_features: int = 1000, max_width: int = 1920, max_height: int = 1080, device: int = 0):
        self.ctx = Context(max_width, max_height, nfeatures=num_features, batch=1, device=device)

    def detect(self, image: np.ndarray) -> np.ndarray:
        return self.ctx.orb_detect_and_compute(image, want_desc=False)[0]

    def detect_and_compute(self, image: np.ndarray):
        return self.ctx.orb_detect_and_compute(image, want_desc=True)

    def find_matches(self, descriptors1: np.ndarray, descriptors2: np.ndarray, lowes_distance_ratio: float):
        return self.ctx.knn_ratio(descriptors1, descriptors2, lowes_distance_ratio)
