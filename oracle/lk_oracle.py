"""CPU oracle (numpy restatement) of cv::calcOpticalFlowPyrLK with all-default arguments.

TEST INFRASTRUCTURE ONLY.
Reference call site: /root/reference/src/tracker.cpp:68-69
    cv::calcOpticalFlowPyrLK(prev_frame_.image, new_frame.image, prev_pts_2d, new_pts_2d, status, err)
defaults: winSize 21x21, maxLevel 3, criteria COUNT|EPS 30 / 0.01, flags 0, minEigThreshold 1e-4.
Arithmetic: pyrDown integer 5x5, Scharr int16 derivatives, Q14 bilinear weights,
patches scaled by 32, float 2x2 solve.  Pinned by tests/golden/lk.npz (cv2 4.13.0): status identical,
positions within 1e-3 px (the residual is float summation order inside OpenCV's SIMD loops).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
WIN = 21
MAX_LEVEL = 3
MAX_ITERS = 30
EPS2 = 0.01 * 0.01                     # criteria.epsilon *= criteria.epsilon (double in OpenCV)
MIN_EIG = 1e-4
W_BITS = 14
FLT_SCALE = f32(1.0 / (1 << 20))
FLT_EPSILON = f32(1.1920929e-07)


def pyr_down(img: np.ndarray) -> np.ndarray:
    """cv::pyrDown: [1 4 6 4 1]^2, REFLECT_101, (sum + 128) >> 8, size ((w+1)/2, (h+1)/2)."""
    if img.ndim == 3:       # cn > 1: every channel on its own
        return np.stack([pyr_down(img[:, :, c]) for c in range(img.shape[2])], 2)
    h, w = img.shape
    dh, dw = (h + 1) // 2, (w + 1) // 2
    p = np.pad(img.astype(np.int64), 2, mode="reflect")
    k = (1, 4, 6, 4, 1)
    # horizontal at even columns for all padded rows
    cols = 2 * np.arange(dw)
    hrow = sum(k[i] * p[:, cols + i] for i in range(5))
    rows = 2 * np.arange(dh)
    out = sum(k[j] * hrow[rows + j, :] for j in range(5))
    return ((out + 128) >> 8).astype(np.uint8)


def build_pyramid(img: np.ndarray, max_level: int = MAX_LEVEL):
    """Levels exist while the NEXT level stays larger than the window in both dimensions."""
    pyr = [img]
    for _ in range(max_level):
        h, w = pyr[-1].shape[:2]
        if (w + 1) // 2 <= WIN or (h + 1) // 2 <= WIN:
            break
        pyr.append(pyr_down(pyr[-1]))
    return pyr


def scharr(img: np.ndarray):
    """calcScharrDeriv: int16 dx, dy with REFLECT_101 borders."""
    if img.ndim == 3:
        d = [scharr(img[:, :, c]) for c in range(img.shape[2])]
        return np.stack([a for a, _ in d], 2), np.stack([b for _, b in d], 2)
    p = np.pad(img.astype(np.int32), 1, mode="reflect")
    t0 = (p[:-2, :] + p[2:, :]) * 3 + p[1:-1, :] * 10       # vertical smooth, all padded columns
    t1 = p[2:, :] - p[:-2, :]                                # vertical diff
    dx = t0[:, 2:] - t0[:, :-2]
    dy = (t1[:, 2:] + t1[:, :-2]) * 3 + t1[:, 1:-1] * 10
    return dx.astype(np.int16), dy.astype(np.int16)


def _weights(a, b):
    one = f32(1.0)
    s = f32(1 << W_BITS)
    iw00 = np.rint(((one - a) * (one - b)).astype(f32) * s).astype(np.int64)
    iw01 = np.rint((a * (one - b)).astype(f32) * s).astype(np.int64)
    iw10 = np.rint(((one - a) * b).astype(f32) * s).astype(np.int64)
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


def _interp(P, ix, iy, w, shift, pad):
    """Bilinear Q14 interpolation of the (WIN x WIN) window whose top-left integer corner is (ix, iy)."""
    iw00, iw01, iw10, iw11 = w
    yy = (iy + pad)[:, None, None] + np.arange(WIN)[None, :, None]
    xx = (ix + pad)[:, None, None] + np.arange(WIN)[None, None, :]
    e = (slice(None), None, None) + ((None,) if P.ndim == 3 else ())
    v = (P[yy, xx] * iw00[e] + P[yy, xx + 1] * iw01[e] + P[yy + 1, xx] * iw10[e] + P[yy + 1, xx + 1] * iw11[e])
    return (v + (1 << (shift - 1))) >> shift


def lk_track(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray):
    """Returns next_pts (N x 2 f32), status (N u8), err (N f32).  Images are H x W (cn = 1) or H x W x cn (the node
    feeds BGR8, /root/reference/src/mono_vo.cpp:94): the window then spans all channels; only err is normalised by cn."""
    cn = 1 if prev.ndim == 2 else prev.shape[2]
    sum_axes = (1, 2) if cn == 1 else (1, 2, 3)
    pts = np.asarray(pts, f32).reshape(-1, 2)
    n = len(pts)
    status = np.ones(n, np.uint8)
    err = np.zeros(n, f32)
    next_pts = np.zeros((n, 2), f32)
    if n == 0:
        return next_pts, status, err
    pyr_i = build_pyramid(prev)
    pyr_j = build_pyramid(nxt)
    top = len(pyr_i) - 1
    half = f32((WIN - 1) * 0.5)
    pad = WIN + 1
    for level in range(top, -1, -1):
        I, J = pyr_i[level], pyr_j[level]
        h, w = I.shape[:2]
        dx, dy = scharr(I)
        pw = ((pad, pad), (pad, pad)) + (((0, 0),) if cn > 1 else ())
        Ip = np.pad(I.astype(np.int64), pw, mode="reflect")
        Jp = np.pad(J.astype(np.int64), pw, mode="reflect")
        dxp = np.pad(dx.astype(np.int64), pw, mode="constant")
        dyp = np.pad(dy.astype(np.int64), pw, mode="constant")
        prev_pt = (pts * f32(1.0 / (1 << level))).astype(f32)
        if level == top:
            next_pt = prev_pt.copy()
        else:
            next_pt = (next_pts * f32(2.0)).astype(f32)
        next_pts = next_pt.copy()
        pp = (prev_pt - half).astype(f32)
        # cvFloor of a NaN is INT_MIN (cvtss2si): such a point is out of range on every level
        ip = np.where(np.isnan(pp), -2 ** 31, np.floor(np.nan_to_num(pp))).astype(np.int64)
        oob = (ip[:, 0] < -WIN) | (ip[:, 0] >= w) | (ip[:, 1] < -WIN) | (ip[:, 1] >= h)
        if level == 0:
            status[oob] = 0
            err[oob] = 0
        act = np.nonzero(~oob)[0]
        if len(act) == 0:
            continue
        a = (pp[act, 0] - ip[act, 0].astype(f32)).astype(f32)
        b = (pp[act, 1] - ip[act, 1].astype(f32)).astype(f32)
        wts = _weights(a, b)
        Iw = _interp(Ip, ip[act, 0], ip[act, 1], wts, W_BITS - 5, pad)
        Ix = _interp(dxp, ip[act, 0], ip[act, 1], wts, W_BITS, pad)
        Iy = _interp(dyp, ip[act, 0], ip[act, 1], wts, W_BITS, pad)
        A11 = ((Ix * Ix).sum(sum_axes).astype(f32) * FLT_SCALE).astype(f32)
        A12 = ((Ix * Iy).sum(sum_axes).astype(f32) * FLT_SCALE).astype(f32)
        A22 = ((Iy * Iy).sum(sum_axes).astype(f32) * FLT_SCALE).astype(f32)
        D = (A11 * A22 - A12 * A12).astype(f32)
        min_eig = ((A22 + A11 - np.sqrt(((A11 - A22) * (A11 - A22) + f32(4.0) * A12 * A12).astype(f32))).astype(f32)
                   / f32(2 * WIN * WIN)).astype(f32)
        bad = (min_eig < MIN_EIG) | (D < FLT_EPSILON)
        if level == 0:
            status[act[bad]] = 0
        good = ~bad
        act, Iw, Ix, Iy = act[good], Iw[good], Ix[good], Iy[good]
        A11, A12, A22, D = A11[good], A12[good], A22[good], D[good]
        if len(act) == 0:
            continue
        Dinv = (f32(1.0) / D).astype(f32)
        npt = (next_pt[act] - half).astype(f32)
        prev_delta = np.zeros((len(act), 2), f32)
        running = np.ones(len(act), bool)
        for j in range(MAX_ITERS):
            r = np.nonzero(running)[0]
            if len(r) == 0:
                break
            inx = np.floor(npt[r]).astype(np.int64)
            oob = (inx[:, 0] < -WIN) | (inx[:, 0] >= w) | (inx[:, 1] < -WIN) | (inx[:, 1] >= h)
            if oob.any():
                if level == 0:
                    status[act[r[oob]]] = 0
                running[r[oob]] = False
                r, inx = r[~oob], inx[~oob]
                if len(r) == 0:
                    break
            a = (npt[r, 0] - inx[:, 0].astype(f32)).astype(f32)
            b = (npt[r, 1] - inx[:, 1].astype(f32)).astype(f32)
            Jw = _interp(Jp, inx[:, 0], inx[:, 1], _weights(a, b), W_BITS - 5, pad)
            diff = Jw - Iw[r]
            b1 = ((diff * Ix[r]).sum(sum_axes).astype(f32) * FLT_SCALE).astype(f32)
            b2 = ((diff * Iy[r]).sum(sum_axes).astype(f32) * FLT_SCALE).astype(f32)
            delta = np.stack([((A12[r] * b2 - A22[r] * b1).astype(f32) * Dinv[r]).astype(f32),
                              ((A12[r] * b1 - A11[r] * b2).astype(f32) * Dinv[r]).astype(f32)], 1)
            npt[r] = (npt[r] + delta).astype(f32)
            next_pts[act[r]] = (npt[r] + half).astype(f32)
            dd = delta.astype(np.float64)
            conv = (dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) <= float(EPS2)
            if j > 0:
                osc = (~conv) & (np.abs(delta[:, 0] + prev_delta[r, 0]) < 0.01) & \
                      (np.abs(delta[:, 1] + prev_delta[r, 1]) < 0.01)
                next_pts[act[r[osc]]] = (next_pts[act[r[osc]]] - delta[osc] * f32(0.5)).astype(f32)
                conv = conv | osc
            running[r[conv]] = False
            prev_delta[r] = delta
        if level == 0:
            ok = np.nonzero(status[act] != 0)[0]
            if len(ok):
                q = (next_pts[act[ok]] - half).astype(f32)
                iq = np.floor(q).astype(np.int64)
                oob = (iq[:, 0] < -WIN) | (iq[:, 0] >= w) | (iq[:, 1] < -WIN) | (iq[:, 1] >= h)
                status[act[ok[oob]]] = 0
                ok, q, iq = ok[~oob], q[~oob], iq[~oob]
                if len(ok):
                    a = (q[:, 0] - iq[:, 0].astype(f32)).astype(f32)
                    b = (q[:, 1] - iq[:, 1].astype(f32)).astype(f32)
                    Jw = _interp(Jp, iq[:, 0], iq[:, 1], _weights(a, b), W_BITS - 5, pad)
                    e = np.abs(Jw - Iw[ok]).sum(sum_axes).astype(f32)
                    err[act[ok]] = ((e * f32(1.0)) / f32(32 * WIN * cn * WIN)).astype(f32)   # errval * 1.f/(32*w*cn*h)
    return next_pts, status, err
