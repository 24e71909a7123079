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
RESULT_DTYPE = np.dtype([("n_keypoints", np.int32), ("n_matches", np.int32), ("n_tracked", np.int32),
                         ("score_h", np.int32), ("score_f", np.int32), ("n_inliers_e", np.int32),
                         ("n_pose_good", np.int32), ("n_triangulated", np.int32),
                         ("R", np.float64, (9,)), ("t", np.float64, (3,))])
TRACK_RESULT_DTYPE = np.dtype([("n_prev", np.int32), ("n_tracked", np.int32), ("n_pnp_inliers", np.int32),
                               ("pnp_ok", np.int32), ("rvec", np.float64, (3,)), ("tvec", np.float64, (3,))])
assert TRACK_RESULT_DTYPE.itemsize == C.sizeof(_lib.MvoTrackResult)
assert RESULT_DTYPE.itemsize == C.sizeof(_lib.MvoFrameResult)
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

    def measure_popc_peak(self) -> float:
        """Measured 32-bit popc per second of this GPU (roofline denominator of the matching kernel)."""
        v = C.c_double()
        self._check(self.lib.mvo_measure_popc_peak(self.h, C.byref(v)))
        return float(v.value)

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

    # ---- stream-group front-end step ---------------------------------------------------------
    STAGES = ("orb", "knn", "lk", "ransac_h", "ransac_f", "ransac_e", "pose", "triangulate", "total", "orb_dense")

    channels = 1

    def group_configure(self, channels: int = 1, outputs: int = 0):
        """Frame format of the group entry points (1 = gray8, 3 = BGR8) and the MVO_OUT_* outputs copied to pinned host
        memory inside every step (mvo_group_configure)."""
        cfg = _lib.MvoGroupConfig(channels, outputs)
        self._check(self.lib.mvo_group_configure(self.h, C.byref(cfg)))
        self.channels = channels

    def group_output_bytes(self) -> int:
        v = C.c_size_t()
        self._check(self.lib.mvo_group_output_bytes(self.h, C.byref(v)))
        return int(v.value)

    def _frames_arg(self, images: np.ndarray):
        """(pointer, w, h, row stride in bytes) of a batch x h x w [x 3] u8 host array"""
        assert images.dtype == np.uint8 and images.shape[0] == self.batch and images.strides[-1] == 1
        assert images.ndim == (3 if self.channels == 1 else 4)
        if self.channels == 3:
            assert images.shape[3] == 3 and images.strides[2] == 3
        h, w = images.shape[1:3]
        assert self.batch == 1 or images.strides[0] == h * images.strides[1]   # numpy leaves the stride of a length-1 axis arbitrary
        return _ptr(images), w, h, images.strides[1]

    def group_step(self, images: np.ndarray, K, device_ptr: int | None = None, shape=None):
        """One front-end frame for every stream of the group.  images: batch x h x w u8 (host; batch x h x w x 3 after
        group_configure(channels=3)), or pass device_ptr (+ shape=(h, w)) for frames already resident in HBM.  Returns a
        structured array."""
        Kp = np.ascontiguousarray(K, np.float64).reshape(9)
        res = np.zeros(self.batch, RESULT_DTYPE)
        if device_ptr is None:
            images = np.ascontiguousarray(images, np.uint8)
            p, w, h, stride = self._frames_arg(images)
            self._check(self.lib.mvo_group_step(self.h, p, w, h, stride, 0, _ptr(Kp), _ptr(res)))
        else:
            h, w = shape
            self._check(self.lib.mvo_group_step(self.h, C.c_void_p(device_ptr), w, h, w * self.channels, 1, _ptr(Kp), _ptr(res)))
        return res

    def group_submit(self, images: np.ndarray, K, device_ptr: int | None = None, shape=None):
        """Pipelined group_step: enqueue one step on (pinned) host frames -- or frames resident in HBM (device_ptr +
        shape=(h, w)) -- and return at once (<= 2 in flight).  Host frames must stay valid and unchanged until the step
        has been collected."""
        Kp = np.ascontiguousarray(K, np.float64).reshape(9)
        if device_ptr is not None:
            h, w = shape
            self._check(self.lib.mvo_group_submit(self.h, C.c_void_p(device_ptr), w, h, w * self.channels, 1, _ptr(Kp)))
            return
        p, w, h, stride = self._frames_arg(images)
        self._check(self.lib.mvo_group_submit(self.h, p, w, h, stride, 0, _ptr(Kp)))

    def group_outputs(self, stream: int, copy: bool = True) -> dict:
        """Full outputs of one stream for the step most recently returned by group_step / group_collect (the sections
        chosen with group_configure).  copy=False returns views into the context's pinned memory (valid until the next
        submit / step)."""
        o = _lib.MvoStreamOutputs()
        self._check(self.lib.mvo_group_outputs(self.h, int(stream), C.byref(o)))

        def view(ptr, dtype, shape):
            if not ptr:
                return None
            n = int(np.prod(shape)) * np.dtype(dtype).itemsize
            a = np.frombuffer((C.c_uint8 * n).from_address(ptr), dtype=dtype).reshape(shape)
            return a.copy() if copy else a

        nk, nm, npv, nt = o.n_keypoints, o.n_matches, o.n_prev, o.n_tracked
        out = {"n_keypoints": nk, "n_matches": nm, "n_prev": npv, "n_tracked": nt, "flags": o.flags,
               "keypoints": view(o.keypoints, KP_DTYPE, (nk,)), "descriptors": view(o.descriptors, np.uint8, (nk, 32)),
               "matches": view(o.matches, DMATCH_DTYPE, (nm,)),
               "track_xy": view(o.track_xy, np.float32, (npv, 2)), "track_status": view(o.track_status, np.uint8, (npv,)),
               "track_err": view(o.track_err, np.float32, (npv,)),
               "mask_h": view(o.mask_h, np.uint8, (nt,)), "mask_f": view(o.mask_f, np.uint8, (nt,)),
               "mask_e": view(o.mask_e, np.uint8, (nt,)), "mask_pose": view(o.mask_pose, np.uint8, (nt,)),
               "H": np.array(o.H).reshape(3, 3), "F": np.array(o.F).reshape(3, 3), "E": np.array(o.E).reshape(3, 3)}
        out["occupied_cells"], out["total_cells"] = int(o.occupied_cells), int(o.total_cells)
        out["cloud_xyz"] = view(o.cloud_xyz, np.float32, (int(o.n_cloud), 3))
        if o.X4:
            full = view(o.X4, np.float32, (4, int(o.x4_stride)))
            out["X4"] = full[:, :nt].copy() if copy else full[:, :nt]
        else:
            out["X4"] = None
        return out

    # ---- stream-group tracking frame (Tracker::update) ------------------------------------------
    def group_set_tracks(self, stream: int, xy, xyz):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        assert len(xy) == len(xyz)
        self._check(self.lib.mvo_group_set_tracks(self.h, int(stream), _ptr(xy), _ptr(xyz), len(xy)))

    def group_track(self, images: np.ndarray, K):
        """LK from every stream's latest frame to the new one on its tracked observations + solvePnPRansac against
        their landmarks (mvo_group_track).  Returns a structured array of mvo_track_result."""
        Kp = np.ascontiguousarray(K, np.float64).reshape(9)
        res = np.zeros(self.batch, TRACK_RESULT_DTYPE)
        images = np.ascontiguousarray(images, np.uint8)
        p, w, h, stride = self._frames_arg(images)
        self._check(self.lib.mvo_group_track(self.h, p, w, h, stride, 0, _ptr(Kp), _ptr(res)))
        return res

    def group_get_tracks(self, stream: int):
        """(xy, index into the previous observation list, PnP inlier indices) of one stream after group_track."""
        nt, ni = C.c_int32(), C.c_int32()
        self._check(self.lib.mvo_group_get_tracks(self.h, int(stream), None, None, None, 0, C.byref(nt), C.byref(ni)))
        cap = max(nt.value, ni.value, 1)
        xy = np.zeros((cap, 2), np.float32)
        src = np.zeros(cap, np.int32)
        inl = np.zeros(cap, np.int32)
        self._check(self.lib.mvo_group_get_tracks(self.h, int(stream), _ptr(xy), _ptr(src), _ptr(inl), cap, C.byref(nt),
                                                  C.byref(ni)))
        return xy[:nt.value], src[:nt.value], inl[:ni.value]

    def group_collect(self):
        """Results of the oldest submitted step."""
        res = np.zeros(self.batch, RESULT_DTYPE)
        self._check(self.lib.mvo_group_collect(self.h, _ptr(res)))
        return res

    def cache_stats(self) -> dict:
        """Hits / misses of the content-keyed device caches (descriptor blocks, LK pyramids) since creation."""
        v = (C.c_uint64 * 4)()
        self._check(self.lib.mvo_cache_stats(self.h, v))
        return {"desc_hits": int(v[0]), "desc_misses": int(v[1]), "pyr_hits": int(v[2]), "pyr_misses": int(v[3])}

    def graph_stats(self) -> dict:
        """CUDA-graph form of the synchronous group step: graphs captured, steps replayed, captures abandoned."""
        v = (C.c_uint64 * 3)()
        self._check(self.lib.mvo_graph_stats(self.h, v))
        return {"captures": int(v[0]), "replays": int(v[1]), "fallbacks": int(v[2])}

    # ---- SURVEY 8(f) #4 by-products -------------------------------------------------------------
    def set_occupancy_grid(self, grid_div: int):
        """Cell size of the keypoint-distribution grid (Initializer::good_keypoint_distribution); 0 = off."""
        self._check(self.lib.mvo_set_occupancy_grid(self.h, int(grid_div)))

    def orb_occupancy(self, stream: int = 0):
        """(occupied cells, total cells) of the last orb_detect_and_compute / collected group step."""
        a, b = C.c_int32(), C.c_int32()
        self._check(self.lib.mvo_orb_occupancy(self.h, int(stream), C.byref(a), C.byref(b)))
        return a.value, b.value

    def pack_pointcloud(self, points_xyz) -> np.ndarray:
        """points3d_to_pointcloud_msg's PointCloud2.data (n * 12 bytes, ROS axes) for n x 3 f32 points."""
        pts = np.ascontiguousarray(points_xyz, np.float32).reshape(-1, 3)
        out = np.zeros(len(pts) * 12, np.uint8)
        self._check(self.lib.mvo_pack_pointcloud(self.h, _ptr(pts), len(pts), 0, _ptr(out), 0))
        return out

    def group_reset(self):
        self._check(self.lib.mvo_group_reset(self.h))

    def debug_set(self, key: str, value: int):
        """Profiling / parity knob (mvo_debug_set): e.g. ("lk_impl", 1) selects the first-generation LK kernel."""
        self._check(self.lib.mvo_debug_set(self.h, key.encode(), int(value)))

    def debug_time(self, what: str, reps: int = 10) -> float:
        """Average device ms of one stage re-run on the state of the last group step (mvo_debug_time)."""
        v = C.c_float()
        self._check(self.lib.mvo_debug_time(self.h, what.encode(), int(reps), C.byref(v)))
        return float(v.value)

    def stage_ms(self) -> dict:
        out = {}
        v = C.c_float()
        for s in self.STAGES:
            self._check(self.lib.mvo_stage_ms(self.h, s.encode(), C.byref(v)))
            out[s] = float(v.value)
        return out

    def lk_level(self, which: int, level: int, plane: int = 0):
        """Parity tap: pyramid level of the previous (0) / next (1) image of the last lk_track call; None past the top."""
        w, h = C.c_int(), C.c_int()
        if self.lib.mvo_lk_get_level(self.h, which, level, plane, None, 0, C.byref(w), C.byref(h)) != _lib.MVO_OK:
            return None
        out = np.zeros((h.value, w.value), np.uint8)
        self._check(self.lib.mvo_lk_get_level(self.h, which, level, plane, _ptr(out), out.strides[0], C.byref(w), C.byref(h)))
        return out

    def stage_spans_ms(self) -> dict:
        """(start, end) of every stage of the last enqueued step, relative to the start of that step."""
        out = {}
        a, b = C.c_float(), C.c_float()
        for s in self.STAGES:
            self._check(self.lib.mvo_stage_span_ms(self.h, s.encode(), C.byref(a), C.byref(b)))
            out[s] = (float(a.value), float(b.value))
        return out

    # ---- two-view geometry -------------------------------------------------------------------
    @staticmethod
    def _pts(p):
        return np.ascontiguousarray(p, np.float32).reshape(-1, 2)

    def find_homography(self, p1, p2, thr: float = 3.0):
        """cv::findHomography(p1, p2, RANSAC, thr) -> (H 3x3, mask u8, n_inliers)."""
        p1, p2 = self._pts(p1), self._pts(p2)
        H = np.zeros(9)
        mask = np.zeros(len(p1), np.uint8)
        n = C.c_int32()
        self._check(self.lib.mvo_find_homography(self.h, _ptr(p1), _ptr(p2), len(p1), float(thr), _ptr(H), _ptr(mask),
                                                 C.byref(n)))
        return H.reshape(3, 3), mask, n.value

    def find_fundamental(self, p1, p2, thr: float = 3.0, conf: float = 0.99):
        """cv::findFundamentalMat(p1, p2, FM_RANSAC, thr, conf) -> (F, mask, n_inliers)."""
        p1, p2 = self._pts(p1), self._pts(p2)
        F = np.zeros(9)
        mask = np.zeros(len(p1), np.uint8)
        n = C.c_int32()
        self._check(self.lib.mvo_find_fundamental(self.h, _ptr(p1), _ptr(p2), len(p1), float(thr), float(conf),
                                                  _ptr(F), _ptr(mask), C.byref(n)))
        return F.reshape(3, 3), mask, n.value

    def find_essential(self, p1, p2, K, conf: float = 0.999, thr: float = 1.0):
        """cv::findEssentialMat(p1, p2, K, RANSAC, conf, thr) -> (E, mask, n_inliers)."""
        p1, p2 = self._pts(p1), self._pts(p2)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        E = np.zeros(9)
        mask = np.zeros(len(p1), np.uint8)
        n = C.c_int32()
        self._check(self.lib.mvo_find_essential(self.h, _ptr(p1), _ptr(p2), len(p1), _ptr(K), float(conf), float(thr),
                                                _ptr(E), _ptr(mask), C.byref(n)))
        return E.reshape(3, 3), mask, n.value

    def recover_pose(self, E, p1, p2, K, mask=None):
        """cv::recoverPose(E, p1, p2, K, R, t, mask) -> (R, t, mask, n_good)."""
        p1, p2 = self._pts(p1), self._pts(p2)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        E = np.ascontiguousarray(E, np.float64).reshape(9)
        R = np.zeros(9)
        t = np.zeros(3)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).copy()
        n = C.c_int32()
        self._check(self.lib.mvo_recover_pose(self.h, _ptr(E), _ptr(p1), _ptr(p2), len(p1), _ptr(K), _ptr(R), _ptr(t),
                                              _ptr(m) if m is not None else None, C.byref(n)))
        return R.reshape(3, 3), t, m, n.value

    def triangulate(self, P0, P1, p0, p1):
        """cv::triangulatePoints(P0, P1, p0, p1) -> 4 x N f32."""
        p0, p1 = self._pts(p0), self._pts(p1)
        P0 = np.ascontiguousarray(P0, np.float64).reshape(12)
        P1 = np.ascontiguousarray(P1, np.float64).reshape(12)
        X = np.zeros((4, len(p0)), np.float32)
        self._check(self.lib.mvo_triangulate(self.h, _ptr(P0), _ptr(P1), _ptr(p0), _ptr(p1), len(p0), _ptr(X)))
        return X

    def solve_pnp_ransac(self, obj, img, K, dist=None, iterations: int = 100, reproj_err: float = 8.0,
                         confidence: float = 0.99):
        """cv::solvePnPRansac(obj, img, K, dist, rvec, tvec, false, iterations, reproj_err, confidence, inliers)
        -> (ok, rvec, tvec, inliers).  Arguments default to the reference's (src/tracker.cpp:309)."""
        obj = np.ascontiguousarray(obj, np.float32).reshape(-1, 3)
        img = self._pts(img)
        assert len(obj) == len(img)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        d = None if dist is None else np.ascontiguousarray(dist, np.float64).ravel()
        rvec, tvec = np.zeros(3), np.zeros(3)
        inl = np.zeros(max(len(obj), 1), np.int32)
        n = C.c_int32()
        rc = self.lib.mvo_solve_pnp_ransac(self.h, _ptr(obj), _ptr(img), len(obj), _ptr(K),
                                           _ptr(d) if d is not None else None, 0 if d is None else len(d), iterations,
                                           reproj_err, confidence, _ptr(rvec), _ptr(tvec), _ptr(inl), C.byref(n))
        if rc == _lib.MVO_ERR_DEGENERATE:
            return False, rvec, tvec, inl[:0]
        self._check(rc)
        return True, rvec, tvec, inl[:n.value].copy()

    def pnp_hypotheses(self, iterations: int = 100):
        """Parity hook: (subsets, models R|t, counts) of the last solve_pnp_ransac call."""
        sub = np.zeros((iterations, 5), np.int32)
        mdl = np.zeros((iterations, 12), np.float64)
        cnt = np.zeros(iterations, np.int32)
        self._check(self.lib.mvo_pnp_get_hypotheses(self.h, iterations, _ptr(sub), _ptr(mdl), _ptr(cnt)))
        return sub, mdl, cnt

    def rodrigues(self, rvec):
        """cv::Rodrigues(rvec) -> R (3 x 3)."""
        r = np.ascontiguousarray(rvec, np.float64).reshape(3)
        R = np.zeros(9)
        self._check(self.lib.mvo_rodrigues(_ptr(r), _ptr(R)))
        return R.reshape(3, 3)

    def score_hypotheses(self, model: int, p1, p2, m: int, thr: float = 1.0, K=None, want_models: bool = False):
        """C4 sweep: m minimal samples from the OpenCV RNG stream, solved and scored with no early exit."""
        p1, p2 = self._pts(p1), self._pts(p2)
        kk, mm = {0: (4, 1), 1: (7, 3), 2: (5, 10)}[model]
        idx = np.zeros((m, kk), np.int32)
        counts = np.zeros((m, mm), np.int32)
        models = np.zeros((m, mm, 9)) if want_models else None
        Kp = None if K is None else np.ascontiguousarray(K, np.float64).reshape(9)
        self._check(self.lib.mvo_score_hypotheses(self.h, model, _ptr(p1), _ptr(p2), len(p1),
                                                  _ptr(Kp) if Kp is not None else None, float(thr), m, _ptr(idx),
                                                  _ptr(counts), _ptr(models) if want_models else None))
        return idx, counts, models

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
