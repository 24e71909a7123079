"""Seeded synthetic inputs for the VO front-end hot path (test + bench infrastructure).

numpy-only so the very same arrays can be regenerated on the GPU box (no cv2 dependence).
The frames are "KITTI-shaped": smooth multi-scale texture + many grey rectangles (corner
sources) + mild blur + sensor noise, so that FAST finds more corners per pyramid level than
the ORB per-level quota and every quota saturates.

THIS IS TEST INFRASTRUCTURE: only tests/, bench.py and __graft_entry__.smoke() import it.
"""
from __future__ import annotations

import numpy as np


def _upsample_bilinear(a: np.ndarray, h: int, w: int) -> np.ndarray:
    sh, sw = a.shape
    ys = (np.arange(h) + 0.5) * sh / h - 0.5
    xs = (np.arange(w) + 0.5) * sw / w - 0.5
    y0 = np.clip(np.floor(ys).astype(int), 0, sh - 2)
    x0 = np.clip(np.floor(xs).astype(int), 0, sw - 2)
    fy = np.clip(ys - y0, 0, 1)[:, None]
    fx = np.clip(xs - x0, 0, 1)[None, :]
    a00 = a[y0][:, x0]
    a01 = a[y0][:, x0 + 1]
    a10 = a[y0 + 1][:, x0]
    a11 = a[y0 + 1][:, x0 + 1]
    return (a00 * (1 - fx) + a01 * fx) * (1 - fy) + (a10 * (1 - fx) + a11 * fx) * fy


def _gauss_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    r = max(1, int(np.ceil(3 * sigma)))
    x = np.arange(-r, r + 1)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    p = np.pad(img, r, mode="reflect")
    t = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(2 * r + 1))
    return sum(k[i] * t[i:i + img.shape[0], :] for i in range(2 * r + 1))


def synth_frame(h: int, w: int, seed: int) -> np.ndarray:
    """One grayscale u8 frame, deterministic in (h, w, seed)."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128.0)
    for s in (2, 4, 8, 16, 32):
        coarse = rng.uniform(-1.0, 1.0, size=(h // s + 2, w // s + 2))
        img += 12.0 * np.sqrt(s) * _upsample_bilinear(coarse, h, w)
    img = np.clip(img, 0, 255)
    nrect = h * w // 600
    ys = rng.integers(0, h, nrect)
    xs = rng.integers(0, w, nrect)
    hs = rng.integers(4, 49, nrect)
    ws = rng.integers(4, 49, nrect)
    gs = rng.integers(0, 256, nrect)
    for y, x, rh, rw, g in zip(ys, xs, hs, ws, gs):
        img[y:y + rh, x:x + rw] = g
    img = _gauss_blur(img, 0.8)
    img += rng.normal(0.0, 2.0, size=(h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def warp_similarity(img: np.ndarray, angle_deg: float, tx: float, ty: float) -> np.ndarray:
    """Bilinear warp of a u8 image by a small rotation about the centre + shift (reflect-101)."""
    h, w = img.shape
    a = np.deg2rad(angle_deg)
    ca, sa = np.cos(a), np.sin(a)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    # dst -> src (inverse map)
    xd, yd = xx - cx - tx, yy - cy - ty
    xs = ca * xd + sa * yd + cx
    ys = -sa * xd + ca * yd + cy
    x0 = np.floor(xs).astype(int)
    y0 = np.floor(ys).astype(int)
    fx, fy = xs - x0, ys - y0

    def refl(i, n):
        i = np.abs(i)
        i = np.where(i >= n, 2 * (n - 1) - i, i)
        return np.clip(i, 0, n - 1)

    x0r, x1r = refl(x0, w), refl(x0 + 1, w)
    y0r, y1r = refl(y0, h), refl(y0 + 1, h)
    f = img.astype(np.float64)
    out = (f[y0r, x0r] * (1 - fx) + f[y0r, x1r] * fx) * (1 - fy) + \
          (f[y1r, x0r] * (1 - fx) + f[y1r, x1r] * fx) * fy
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def synth_pair(h: int, w: int, seed: int):
    """Frame t and frame t+1 = small seeded similarity warp of it (C1/C3 pairs)."""
    rng = np.random.default_rng(seed + 7919)
    f0 = synth_frame(h, w, seed)
    ang = rng.uniform(-0.4, 0.4)
    tx, ty = rng.uniform(-3, 3, 2)
    return f0, warp_similarity(f0, ang, tx, ty)


def synth_pair_bgr(h: int, w: int, seed: int):
    """BGR8 version of synth_pair (the node feeds BGR8 images, /root/reference/src/mono_vo.cpp:94): per-channel
    gain / offset plus independent sensor noise, so the three planes are correlated but not identical."""
    f0, f1 = synth_pair(h, w, seed)
    rng = np.random.default_rng(seed + 104729)
    gains = ((0.9, 10.0), (1.0, 0.0), (0.8, 25.0))
    out = []
    for f in (f0, f1):
        noise = rng.normal(0.0, 1.5, (h, w, 3))
        out.append(np.stack([np.clip(np.rint(f * g + o + noise[:, :, i]), 0, 255) for i, (g, o) in enumerate(gains)],
                            2).astype(np.uint8))
    return out[0], out[1]


def pnp_scene(n: int, seed: int, noise: float = 0.5, outlier_frac: float = 0.1, planar: bool = False):
    """3-D landmarks + their pixel observations under a seeded pose (the Tracker's solvePnPRansac input,
    /root/reference/src/tracker.cpp:298-309).  Returns obj (n x 3 f32), img (n x 2 f32), K, rvec, tvec."""
    rng = np.random.default_rng(seed + 15485863)
    K = np.array([[718.856, 0, 620.5], [0, 718.856, 188.0], [0, 0, 1.0]])
    X = np.stack([rng.uniform(-10, 10, n), rng.uniform(-3, 3, n), rng.uniform(5, 40, n)], 1)
    if planar:
        X[:, 2] = 20.0 + 0.3 * X[:, 0] - 0.2 * X[:, 1]
    rvec = np.array([0.01, 0.03, -0.005]) * rng.uniform(0.5, 3.0)
    tvec = np.array([0.1, -0.02, 0.8]) * rng.uniform(0.5, 2.0)
    th = np.linalg.norm(rvec)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)
    Xc = X @ R.T + tvec
    uv = Xc[:, :2] / Xc[:, 2:3] * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]]
    uv += rng.normal(0, noise, uv.shape)
    no = int(outlier_frac * n)
    oi = rng.choice(n, no, replace=False)
    uv[oi] = np.stack([rng.uniform(0, 1241, no), rng.uniform(0, 376, no)], 1)
    return X.astype(np.float32), uv.astype(np.float32), K, rvec, tvec


def _smooth_field(h: int, w: int, rng, scales=(16, 64)) -> np.ndarray:
    f = np.zeros((h, w))
    for s in scales:
        f += _upsample_bilinear(rng.uniform(-1.0, 1.0, size=(h // s + 2, w // s + 2)), h, w)
    f -= f.min()
    return f / max(f.max(), 1e-9)


def warp_parallax(img: np.ndarray, T, inv_depth: np.ndarray, K: np.ndarray) -> np.ndarray:
    """View of a rigid scene after the camera translated by T (no rotation): the new frame's pixel p' with
    inverse depth rho sees the old frame at x = (x' + T_xy rho) / (1 + T_z rho) (normalised coordinates)."""
    h, w = img.shape
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    xn, yn = (xx - cx) / fx, (yy - cy) / fy
    den = 1.0 + T[2] * inv_depth
    xs = (xn + T[0] * inv_depth) / den * fx + cx
    ys = (yn + T[1] * inv_depth) / den * fy + cy
    x0 = np.floor(xs).astype(int)
    y0 = np.floor(ys).astype(int)
    fxr, fyr = xs - x0, ys - y0

    def refl(i, n):
        i = np.abs(i)
        i = np.where(i >= n, 2 * (n - 1) - i, i)
        return np.clip(i, 0, n - 1)

    x0r, x1r = refl(x0, w), refl(x0 + 1, w)
    y0r, y1r = refl(y0, h), refl(y0 + 1, h)
    f = img.astype(np.float64)
    out = (f[y0r, x0r] * (1 - fxr) + f[y0r, x1r] * fxr) * (1 - fyr) + \
          (f[y1r, x0r] * (1 - fxr) + f[y1r, x1r] * fxr) * fyr
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def sequence_camera(h: int, w: int) -> np.ndarray:
    """KITTI-like intrinsics scaled to the frame size."""
    f = 718.856 * w / 1241.0
    return np.array([[f, 0.0, (w - 1) / 2.0], [0.0, f, (h - 1) / 2.0], [0.0, 0.0, 1.0]])


def synth_sequence(h: int, w: int, stream: int, nframes: int, return_poses: bool = False):
    """A seeded camera translating through a rigid non-planar scene (C2/C5 sequences): every frame is one
    resampling of the base frame with a smooth inverse-depth map (depths 6..26, ~0.5 units of forward motion
    per frame, i.e. KITTI-like depth/baseline ratios of 12..52: nearly every point lies inside recoverPose's
    50-baseline distance threshold, so n_pose_good ~ n_inliers_e).  The depth map is attached to the pixel grid of
    every rendered frame, so the scene is rigid only to first order in the per-frame motion (exact for a plane).
    Returns (frames, K), or (frames, K, T) with return_poses: T[k] is the position of camera k in the coordinates
    of camera 0 (no rotation)."""
    rng = np.random.default_rng(1000 * stream + 17)
    base = synth_frame(h, w, 1000 * stream)
    K = sequence_camera(h, w)
    depth = 6.0 + 20.0 * _smooth_field(h, w, rng)
    inv_depth = 1.0 / depth
    frames = [base]
    T = np.zeros(3)
    poses = [T.copy()]
    for _ in range(1, nframes):
        T = T + np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.02, 0.02), rng.uniform(0.4, 0.6)])
        frames.append(warp_parallax(base, T, inv_depth, K))
        poses.append(T.copy())
    if return_poses:
        return frames, K, np.stack(poses)
    return frames, K


def sequence_depth(h: int, w: int, stream: int) -> np.ndarray:
    """The depth map synth_sequence(h, w, stream, ...) renders with (same generator state)."""
    rng = np.random.default_rng(1000 * stream + 17)
    return 6.0 + 20.0 * _smooth_field(h, w, rng)


def _sample_bilinear(f: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """f (float, h x w) at fractional positions, reflect-101 outside."""
    h, w = f.shape
    x0 = np.floor(xs).astype(int)
    y0 = np.floor(ys).astype(int)
    ax, ay = xs - x0, ys - y0

    def refl(i, n):
        i = np.abs(i)
        i = np.where(i >= n, 2 * (n - 1) - i, i)
        return np.clip(i, 0, n - 1)

    x0r, x1r = refl(x0, w), refl(x0 + 1, w)
    y0r, y1r = refl(y0, h), refl(y0 + 1, h)
    return (f[y0r, x0r] * (1 - ax) + f[y0r, x1r] * ax) * (1 - ay) + (f[y1r, x0r] * (1 - ax) + f[y1r, x1r] * ax) * ay


def room_planes(stream: int) -> np.ndarray:
    """Walls of a seeded convex room around camera 0 (rows n_x, n_y, n_z, d with n.X = d, camera axes: x right, y down,
    z forward): a road 1.65 below the camera, a ceiling, two side walls and a front wall, each slightly slanted.  From
    any point inside a convex room every ray meets exactly one wall first and nothing is ever occluded."""
    rng = np.random.default_rng(1000 * stream + 31)
    j = lambda a: rng.uniform(-a, a)
    return np.array([[j(0.03), 1.0, j(0.02), 1.65],
                     [j(0.03), -1.0, j(0.03), rng.uniform(4.0, 6.0)],
                     [-1.0, j(0.05), j(0.08), rng.uniform(5.0, 8.0)],
                     [1.0, j(0.05), j(0.08), rng.uniform(5.0, 8.0)],
                     [j(0.15), j(0.05), 1.0, rng.uniform(28.0, 34.0)]])


def render_rigid(base: np.ndarray, planes: np.ndarray, K: np.ndarray, R_wc: np.ndarray, C: np.ndarray):
    """Exact view of the room `planes`, textured by camera 0's frame `base` (a wall point has the colour of the pixel of
    camera 0 it projects to), from the camera with orientation R_wc (camera -> world = camera 0) at position C: ray /
    wall intersection in closed form per pixel, one bilinear sample of `base`.  The scene is rigid for any motion,
    rotation included (synth_sequence's is rigid to first order only).  Returns the u8 frame and the depth map."""
    h, w = base.shape
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    D = np.tensordot(R_wc @ np.linalg.inv(K), np.stack([xx, yy, np.ones_like(xx)]), 1)     # ray directions (world)
    s = np.full((h, w), np.inf)
    for n0, n1, n2, d in planes:
        den = n0 * D[0] + n1 * D[1] + n2 * D[2]
        num = d - (n0 * C[0] + n1 * C[1] + n2 * C[2])
        si = np.where(den > 1e-12, num / np.where(den > 1e-12, den, 1.0), np.inf)
        s = np.minimum(s, np.where(si > 0, si, np.inf))
    X = C[:, None, None] + s * D
    q = np.tensordot(K, X, 1)
    out = _sample_bilinear(base.astype(np.float64), q[0] / q[2], q[1] / q[2])
    depth = np.tensordot(R_wc.T, X - C[:, None, None], 1)[2]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8), depth


def synth_rigid_sequence(h: int, w: int, stream: int, nframes: int):
    """A seeded camera moving (forward translation, a little sideways drift and a slow yaw / pitch / roll) through an
    exactly rigid scene: every frame is render_rigid of the same textured room (depths 6 .. 33).
    Returns frames, K, R_wc (nframes x 3 x 3), C (nframes x 3): X_world = R_wc[k] X_cam_k + C[k], world = camera 0."""
    rng = np.random.default_rng(1000 * stream + 29)
    base = synth_frame(h, w, 1000 * stream)
    K = sequence_camera(h, w)
    planes = room_planes(stream)
    frames, Rs, Cs = [base], [np.eye(3)], [np.zeros(3)]
    C = np.zeros(3)
    rv = np.zeros(3)
    for _ in range(1, nframes):
        rv = rv + np.array([rng.uniform(-0.0005, 0.0005), rng.uniform(0.001, 0.003), rng.uniform(-0.0003, 0.0003)])
        R = rodrigues(rv)
        C = C + R @ np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.02, 0.02), rng.uniform(0.4, 0.6)])
        frames.append(render_rigid(base, planes, K, R, C)[0])
        Rs.append(R)
        Cs.append(C.copy())
    return frames, K, np.stack(Rs), np.stack(Cs)


def rodrigues(rvec) -> np.ndarray:
    rvec = np.asarray(rvec, dtype=np.float64)
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


KITTI_K = np.array([[718.856, 0.0, 620.5], [0.0, 718.856, 188.0], [0.0, 0.0, 1.0]])


def scene_correspondences(n: int, seed: int, outlier_frac: float = 0.3, noise_px: float = 0.3,
                          planar: bool = False, K: np.ndarray = KITTI_K,
                          rvec=(0.01, 0.03, -0.005), tvec=(0.1, -0.02, 0.8), w: int = 1241, h: int = 376):
    """Two-view correspondences of a seeded 3-D scene (C4): returns p1, p2 (n x 2 f32), R, t, inlier flags."""
    rng = np.random.default_rng(seed)
    R = rodrigues(rvec)
    t = np.asarray(tvec, dtype=np.float64)
    # sample image points in view 1 uniformly, give them a depth, back-project
    u = rng.uniform(0, w, n)
    v = rng.uniform(0, h, n)
    if planar:
        # a slanted plane n.X = d
        nrm = np.array([0.1, -0.2, 1.0]); nrm /= np.linalg.norm(nrm)
        rays = np.linalg.inv(K) @ np.vstack([u, v, np.ones(n)])
        depth = 12.0 / (nrm @ rays)
        X = rays * depth
    else:
        depth = rng.uniform(5, 40, n)
        X = (np.linalg.inv(K) @ np.vstack([u, v, np.ones(n)])) * depth
    x2 = K @ (R @ X + t[:, None])
    p1 = np.stack([u, v], 1)
    p2 = (x2[:2] / x2[2]).T
    p1 = p1 + rng.normal(0, noise_px, p1.shape)
    p2 = p2 + rng.normal(0, noise_px, p2.shape)
    nout = int(round(outlier_frac * n))
    out_idx = rng.permutation(n)[:nout]
    p2[out_idx, 0] = rng.uniform(0, w, nout)
    p2[out_idx, 1] = rng.uniform(0, h, nout)
    inl = np.ones(n, bool)
    inl[out_idx] = False
    return p1.astype(np.float32), p2.astype(np.float32), R, t, inl
