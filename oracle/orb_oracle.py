"""CPU oracle (numpy restatement) of cv::ORB::detectAndCompute as called by the reference.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.

Reference call site: /root/reference/src/feature_processor.cpp:5-23
  cv::ORB::create(num_features)  -> all other parameters default
  detector_->detectAndCompute(image, cv::noArray(), keypoints, descriptors)
The arithmetic lives in un-vendored OpenCV (CMakeLists.txt:17 find_package(OpenCV REQUIRED), no
pin).  Pinned here against cv2 4.13.0 (opencv-python-headless 4.13.0.92): tests/golden/orb_*.npz
were produced by tests/golden/gen_golden.py calling cv2 in the build container, and
tests/test_oracle_orb.py checks this restatement against them (keypoint set, response, angle
bit-exact; descriptors byte-exact).  Spec: SURVEY.md Appendix A.1.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

NLEVELS = 8
EDGE = 31            # edgeThreshold
PATCH = 31           # patchSize
HALF_PATCH = 15
FAST_THR = 20
HARRIS_BLOCK = 7
HARRIS_K = f32(0.04)
SCALE_FACTOR_D = float(f32(1.2))      # (double)1.2f

UMAX = np.array([15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3], dtype=np.int32)

# FAST ring (dx, dy), in OpenCV order
RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
        (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]

import os as _os
_PATTERN = np.load(_os.path.join(_os.path.dirname(__file__), "brief_pattern_31.npy")).astype(np.int32)  # 256x4


def rne(x):
    """cvRound: round half to even."""
    return np.rint(x).astype(np.int64)


# ----------------------------------------------------------------------------------------------
# A.1.2  level geometry
def level_scales(nlevels: int = NLEVELS) -> np.ndarray:
    return np.array([f32(SCALE_FACTOR_D ** l) for l in range(nlevels)], dtype=f32)


def level_sizes(w: int, h: int, nlevels: int = NLEVELS):
    out = []
    for s in level_scales(nlevels):
        inv = f32(1.0) / s
        out.append((int(rne(f32(w) * inv)), int(rne(f32(h) * inv))))
    return out


# ----------------------------------------------------------------------------------------------
# A.1.3  INTER_LINEAR_EXACT resize (8.8 fixed point coefficients)
def _exact_coeffs(src: int, dst: int):
    v = np.arange(dst, dtype=np.float64)
    scale = float(src) / float(dst)
    f = scale * (v + 0.5) - 0.5
    i = np.floor(f).astype(np.int64)
    c1 = rne((f - i) * 256.0)
    off = i.copy()
    lo = i < 0
    hi = i >= src - 1
    off[lo] = 0
    c1[lo] = 0
    off[hi] = src - 1
    c1[hi] = 0
    return off, c1.astype(np.int64)


def resize_linear_exact(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape
    ox, cx1 = _exact_coeffs(sw, dw)
    oy, cy1 = _exact_coeffs(sh, dh)
    cx0 = 256 - cx1
    cy0 = 256 - cy1
    s = src.astype(np.int64)
    ox1 = np.minimum(ox + 1, sw - 1)
    oy1 = np.minimum(oy + 1, sh - 1)
    hrow = s[:, ox] * cx0[None, :] + s[:, ox1] * cx1[None, :]          # 8.8, every source row
    out = (hrow[oy, :] * cy0[:, None] + hrow[oy1, :] * cy1[:, None] + 32768) >> 16
    return out.astype(np.uint8)


def build_pyramid(img: np.ndarray, nlevels: int = NLEVELS):
    h, w = img.shape
    sizes = level_sizes(w, h, nlevels)
    pyr = [img]
    for l in range(1, nlevels):
        pyr.append(resize_linear_exact(pyr[-1], sizes[l][0], sizes[l][1]))
    return pyr


def bgr_to_gray(bgr: np.ndarray) -> np.ndarray:
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# A.1.4  FAST-9/16 + NMS
def fast_score_map(img: np.ndarray, thr: int = FAST_THR) -> np.ndarray:
    """score = (max over 16 arcs of 9 of max(min d, min -d)) - 1 where > thr else 0; rows/cols 3..size-4."""
    h, w = img.shape
    I = img.astype(np.int16)
    c = I[3:h - 3, 3:w - 3]
    d = np.stack([c - I[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in RING])   # 16 x H x W
    d = np.concatenate([d, d[:8]], axis=0)                                            # wrap: 24
    best = np.full(c.shape, -32768, np.int16)
    for k in range(16):
        arc = d[k:k + 9]
        best = np.maximum(best, np.maximum(arc.min(axis=0), (-arc).min(axis=0)))
    score = np.zeros((h, w), np.int32)
    score[3:h - 3, 3:w - 3] = np.where(best > thr, best.astype(np.int32) - 1, 0)
    return score


def fast_nms(score: np.ndarray):
    """Keep iff score > all 8 neighbours (strict); returns (x, y, score) row-major."""
    h, w = score.shape
    s = score
    c = s[1:-1, 1:-1]
    keep = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            keep &= c > s[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
    ys, xs = np.nonzero(keep)
    ys += 1
    xs += 1
    return xs.astype(np.int32), ys.astype(np.int32), s[ys, xs].astype(np.int32)


# ----------------------------------------------------------------------------------------------
# A.1.6  quotas + retainBest
def level_quotas(nfeatures: int, nlevels: int = NLEVELS):
    factor = f32(1.0 / SCALE_FACTOR_D)
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(float(factor) ** nlevels))
    nd = f32(nd)
    out, total = [], 0
    for l in range(nlevels - 1):
        n = int(rne(nd))
        out.append(n)
        total += n
        nd = f32(nd * factor)
    out.append(max(nfeatures - total, 0))
    return out


def retain_best_mask(resp: np.ndarray, k: int) -> np.ndarray:
    """KeyPointsFilter::retainBest: keep ALL with response >= k-th largest (ties kept)."""
    n = len(resp)
    if k <= 0:
        return np.zeros(n, bool)
    if n <= k:
        return np.ones(n, bool)
    r = np.sort(resp)[::-1][k - 1]
    return resp >= r


# ----------------------------------------------------------------------------------------------
# A.1.7  Harris response (blockSize 7, k 0.04), float32 non-fused
def harris_responses(img: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    I = img.astype(np.int64)
    n = len(xs)
    a = np.zeros(n, np.int64)
    b = np.zeros(n, np.int64)
    c = np.zeros(n, np.int64)
    r = HARRIS_BLOCK // 2
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            y = ys + dy
            x = xs + dx
            ix = (I[y, x + 1] - I[y, x - 1]) * 2 + (I[y - 1, x + 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y + 1, x - 1])
            iy = (I[y + 1, x] - I[y - 1, x]) * 2 + (I[y + 1, x - 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y - 1, x + 1])
            a += ix * ix
            b += iy * iy
            c += ix * iy
    scale = f32(1.0) / (f32(4 * HARRIS_BLOCK) * f32(255.0))
    s4 = f32(f32(f32(scale * scale) * scale) * scale)
    af, bf, cf = a.astype(f32), b.astype(f32), c.astype(f32)
    with np.errstate(over="ignore"):
        t1 = f32(af * bf) - f32(cf * cf)
        apb = af + bf
        t2 = f32(HARRIS_K * apb) * apb
        return ((t1 - t2) * s4).astype(f32)


# ----------------------------------------------------------------------------------------------
# A.1.8  intensity-centroid angle
_DEG = f32(180.0 / np.pi)
_P1 = f32(f32(0.9997878412794807) * _DEG)       # float * float, as in the C source
_P3 = f32(f32(-0.3258083974640975) * _DEG)
_P5 = f32(f32(0.1555786518463281) * _DEG)
_P7 = f32(f32(-0.04432655554792128) * _DEG)
_EPS = f32(2.220446049250313e-16)


def fast_atan2(y, x) -> np.ndarray:
    y = np.asarray(y, f32)
    x = np.asarray(x, f32)
    ax, ay = np.abs(x), np.abs(y)
    mn, mx = np.minimum(ax, ay), np.maximum(ax, ay)
    c = (mn / (mx + _EPS)).astype(f32)
    c2 = (c * c).astype(f32)
    a = ((((_P7 * c2 + _P5).astype(f32) * c2 + _P3).astype(f32) * c2 + _P1).astype(f32) * c).astype(f32)
    a = np.where(ax < ay, f32(90.0) - a, a).astype(f32)
    a = np.where(x < 0, f32(180.0) - a, a).astype(f32)
    a = np.where(y < 0, f32(360.0) - a, a).astype(f32)
    return a


def ic_moments(img: np.ndarray, xs: np.ndarray, ys: np.ndarray):
    I = img.astype(np.int64)
    m10 = np.zeros(len(xs), np.int64)
    m01 = np.zeros(len(xs), np.int64)
    for v in range(-HALF_PATCH, HALF_PATCH + 1):
        um = int(UMAX[abs(v)])
        for u in range(-um, um + 1):
            p = I[ys + v, xs + u]
            m10 += u * p
            m01 += v * p
    return m10, m01


def ic_angles(img, xs, ys) -> np.ndarray:
    m10, m01 = ic_moments(img, xs, ys)
    return fast_atan2(m01.astype(f32), m10.astype(f32))


# ----------------------------------------------------------------------------------------------
# A.1.9  ORB-internal Gaussian blur (7x7, sigma 2, float sepFilter, FMA order of the AVX2 path)
def gaussian_kernel7() -> np.ndarray:
    x = np.arange(7, dtype=np.float64) - 3
    k = np.exp(-(x * x) / (2.0 * 2.0 * 2.0))
    k /= k.sum()
    return k.astype(f32)


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def blur_orb(img: np.ndarray) -> np.ndarray:
    k = gaussian_kernel7()
    h, w = img.shape
    p = np.pad(img, 3, mode="reflect").astype(f32)          # REFLECT_101
    # row pass (all padded rows), s = k0*p0; s = fma(kj, pj, s)
    s = (k[0] * p[:, 0:w]).astype(f32)
    for j in range(1, 7):
        s = _fma32(np.full_like(s, k[j]), p[:, j:j + w], s)
    r = s                                                    # (h+6) x w
    # column pass, symmetric pairing
    kk = lambda i: np.full((h, w), k[i], f32)
    o = (k[3] * r[3:3 + h]).astype(f32)
    o = _fma32(kk(2), (r[2:2 + h] + r[4:4 + h]).astype(f32), o)
    o = _fma32(kk(1), (r[1:1 + h] + r[5:5 + h]).astype(f32), o)
    o = _fma32(kk(0), (r[0:0 + h] + r[6:6 + h]).astype(f32), o)
    return np.clip(np.rint(o), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# A.1.10  rotated BRIEF
def brief_descriptors(blurred_levels, kp_x, kp_y, kp_angle, kp_octave, scales) -> np.ndarray:
    """kp_x/kp_y are full-resolution coordinates (level coords * scale), angle in degrees."""
    n = len(kp_x)
    desc = np.zeros((n, 32), np.uint8)
    pat = _PATTERN.reshape(512, 2)
    px = pat[:, 0].astype(f32)
    py = pat[:, 1].astype(f32)
    for i in range(n):
        l = int(kp_octave[i])
        img = blurred_levels[l]
        inv = f32(1.0) / scales[l]
        cx = int(rne(f32(kp_x[i]) * inv))
        cy = int(rne(f32(kp_y[i]) * inv))
        ang = f32(kp_angle[i]) * f32(np.pi / 180.0)
        a = f32(np.cos(np.float64(ang)))
        b = f32(np.sin(np.float64(ang)))
        x = (px * a).astype(f32) - (py * b).astype(f32)
        y = (px * b).astype(f32) + (py * a).astype(f32)
        ix = rne(x.astype(f32))
        iy = rne(y.astype(f32))
        v = img[cy + iy, cx + ix].astype(np.int32)
        bits = (v[0::2] < v[1::2]).astype(np.uint8)          # 256 bits
        desc[i] = np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").ravel()
    return desc


# ----------------------------------------------------------------------------------------------
KP_DTYPE = np.dtype([("x", f32), ("y", f32), ("size", f32), ("angle", f32), ("response", f32),
                     ("octave", np.int32), ("class_id", np.int32)])


def detect_level(level_img: np.ndarray, n_l: int):
    """FAST+NMS -> edge filter -> retainBest(2 n_l) by FAST score -> Harris -> retainBest(n_l)."""
    h, w = level_img.shape
    xs, ys, sc = fast_nms(fast_score_map(level_img))
    m = (xs >= EDGE) & (xs < w - EDGE) & (ys >= EDGE) & (ys < h - EDGE)
    xs, ys, sc = xs[m], ys[m], sc[m]
    m = retain_best_mask(sc.astype(f32), 2 * n_l)
    xs, ys, sc = xs[m], ys[m], sc[m]
    resp = harris_responses(level_img, xs, ys)
    m = retain_best_mask(resp, n_l)
    return xs[m], ys[m], resp[m]


def canonical_order(octave, resp, y, x):
    """The product's canonical keypoint order: level, Harris response desc, y, x."""
    return np.lexsort((x, y, -resp.astype(np.float64), octave))


def orb_detect_and_compute(img: np.ndarray, nfeatures: int, nlevels: int = NLEVELS, want_desc: bool = True):
    """Returns (keypoints[KP_DTYPE] in canonical order, descriptors N x 32 u8)."""
    if img.ndim == 3:
        img = bgr_to_gray(img)
    pyr = build_pyramid(img, nlevels)
    scales = level_scales(nlevels)
    quotas = level_quotas(nfeatures, nlevels)
    recs = []
    for l in range(nlevels):
        xs, ys, resp = detect_level(pyr[l], quotas[l])
        ang = ic_angles(pyr[l], xs, ys)
        k = np.zeros(len(xs), KP_DTYPE)
        k["x"] = xs.astype(f32) * scales[l]
        k["y"] = ys.astype(f32) * scales[l]
        k["size"] = f32(PATCH) * scales[l]
        k["angle"] = ang
        k["response"] = resp
        k["octave"] = l
        k["class_id"] = -1
        order = canonical_order(np.zeros(len(xs)), resp, ys, xs)
        recs.append(k[order])
    kps = np.concatenate(recs) if recs else np.zeros(0, KP_DTYPE)
    if not want_desc:
        return kps, None
    blurred = [blur_orb(p) for p in pyr]
    desc = brief_descriptors(blurred, kps["x"], kps["y"], kps["angle"], kps["octave"], scales)
    return kps, desc


def orb_compute(img: np.ndarray, kps: np.ndarray, nlevels: int = NLEVELS) -> np.ndarray:
    """Descriptors for GIVEN keypoints (parity hook == cv::ORB::compute with kp.angle/octave/pt honoured)."""
    if img.ndim == 3:
        img = bgr_to_gray(img)
    pyr = build_pyramid(img, nlevels)
    blurred = [blur_orb(p) for p in pyr]
    return brief_descriptors(blurred, kps["x"], kps["y"], kps["angle"], kps["octave"], level_scales(nlevels))
