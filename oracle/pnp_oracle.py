"""CPU oracle (numpy restatement) of cv::solvePnPRansac as the reference calls it.

TEST INFRASTRUCTURE ONLY.
Reference call site: /root/reference/src/tracker.cpp:309
    cv::solvePnPRansac(points_3d, points_2d, K, d, rvec, tvec, false, 100, 8.0, 0.99, inliers)
followed by cv::Rodrigues (src/tracker.cpp:315).  SURVEY.md section 8(f), "next #1".

The arithmetic lives in un-vendored OpenCV (calib3d: solvepnp.cpp, epnp.cpp, calibration.cpp, ptsetreg.cpp,
compat_ptsetreg.cpp); restated here from the published algorithms and pinned against cv2 4.13.0:
  * RANSAC registrator with 5-point minimal sets, RNG(2^64-1), <= 100 iterations, threshold 8 px, confidence 0.99
    (the loop of oracle/ransac_oracle.py; the callback has no subset check),
  * minimal solver = EPnP (Lepetit, Moreno-Noguer, Fua 2009) on K-normalised image points, with OpenCV's structure:
    PCA control points, 12x12 M^T M, three beta approximations + 5 Gauss-Newton steps each, absolute orientation by
    SVD, the candidate with the least reprojection error wins.  The control-point axes carry the SIGNS of cv::SVD's
    left singular vectors (one-sided Jacobi, emulated in jacobi_svd); with noisy points the EPnP result depends on them,
  * error = squared reprojection distance in float, inlier iff <= 64,
  * final pose = solvePnP(inliers, SOLVEPNP_ITERATIVE): DLT initialisation (non-planar object) and Levenberg-Marquardt
    as CvLevMarq runs it (<= 20 iterations, eps FLT_EPSILON).
Distortion coefficients must be zero (the node subscribes to a rectified image topic).
"""
from __future__ import annotations

import math

import numpy as np

from oracle.ransac_oracle import CvRNG, update_iters

f32 = np.float32
DBL_EPS = float(np.finfo(np.float64).eps)
FLT_EPS = float(np.finfo(np.float32).eps)
MODEL_POINTS = 5


# --------------------------------------------------------------------------------------------------
def jacobi_svd(A):
    """cv::SVD::compute for a double matrix with rows >= cols: one-sided (Hestenes) Jacobi on the rows of A^T,
    pairs (i, j) in lexicographic order, eps = 10 DBL_EPSILON, singular values sorted by selection.
    Returns w (descending), U (columns = left singular vectors, OpenCV's signs), Vt."""
    A = np.asarray(A, np.float64)
    m, n = A.shape
    At = A.T.copy()
    Vt = np.eye(n)
    eps = DBL_EPS * 10
    W = (At * At).sum(1)
    for _ in range(max(m, 30)):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                a, b = W[i], W[j]
                p = float(At[i] @ At[j])
                if abs(p) <= eps * math.sqrt(a * b):
                    continue
                p *= 2
                beta = a - b
                gamma = math.hypot(p, beta)
                if beta < 0:
                    delta = (gamma - beta) * 0.5
                    s = math.sqrt(delta / gamma)
                    c = p / (gamma * s * 2)
                else:
                    c = math.sqrt((gamma + beta) / (gamma * 2))
                    s = p / (gamma * c * 2)
                t0 = c * At[i] + s * At[j]
                t1 = -s * At[i] + c * At[j]
                At[i], At[j] = t0, t1
                W[i], W[j] = (t0 * t0).sum(), (t1 * t1).sum()
                v0 = c * Vt[i] + s * Vt[j]
                v1 = -s * Vt[i] + c * Vt[j]
                Vt[i], Vt[j] = v0, v1
                changed = True
        if not changed:
            break
    W = np.sqrt((At * At).sum(1))
    for i in range(n - 1):
        j = i
        for k in range(i + 1, n):
            if W[j] < W[k]:
                j = k
        if i != j:
            W[[i, j]] = W[[j, i]]
            At[[i, j]] = At[[j, i]]
            Vt[[i, j]] = Vt[[j, i]]
    U = np.zeros((m, n))
    tiny = float(np.finfo(np.float64).tiny)
    for i in range(n):
        if W[i] > tiny:
            U[:, i] = At[i] / W[i]
    return W, U, Vt


def rodrigues_to_matrix(r, want_jac=False):
    """cv::Rodrigues(rvec) -> R (and dR/dr as 3 x 9, row i = d vec(R) / d r_i)."""
    r = np.asarray(r, np.float64).ravel()
    theta = float(np.linalg.norm(r))
    if theta < DBL_EPS:
        R = np.eye(3)
        if not want_jac:
            return R
        J = np.zeros((3, 9))
        J[0, 5], J[0, 7] = -1, 1
        J[1, 2], J[1, 6] = 1, -1
        J[2, 1], J[2, 3] = -1, 1
        return R, J
    c, s = math.cos(theta), math.sin(theta)
    c1 = 1.0 - c
    it = 1.0 / theta
    k = r * it
    rrt = np.outer(k, k)
    rx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = c * np.eye(3) + c1 * rrt + s * rx
    if not want_jac:
        return R
    I = np.eye(3).ravel()
    drrt = np.array([[k[0] + k[0], k[1], k[2], k[1], 0, 0, k[2], 0, 0],
                     [0, k[0], 0, k[0], k[1] + k[1], k[2], 0, k[2], 0],
                     [0, 0, k[0], 0, 0, k[1], k[0], k[1], k[2] + k[2]]], np.float64)
    d_r_x = np.array([[0, 0, 0, 0, 0, -1, 0, 1, 0],
                      [0, 0, 1, 0, 0, 0, -1, 0, 0],
                      [0, -1, 0, 1, 0, 0, 0, 0, 0]], np.float64)
    J = np.zeros((3, 9))
    for i in range(3):
        ri = k[i]
        a0, a1, a2 = -s * ri, (s - 2 * c1 * it) * ri, c1 * it
        a3, a4 = (c - s * it) * ri, s * it
        J[i] = a0 * I + a1 * rrt.ravel() + a2 * drrt[i] + a3 * rx.ravel() + a4 * d_r_x[i]
    return R, J


def matrix_to_rodrigues(R):
    """cv::Rodrigues(R) -> rvec (R is first projected onto SO(3) by SVD, as OpenCV does)."""
    U, _, Vt = np.linalg.svd(np.asarray(R, np.float64))
    R = U @ Vt
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = math.sqrt((r * r).sum() * 0.25)
    c = min(max((R[0, 0] + R[1, 1] + R[2, 2] - 1) * 0.5, -1.0), 1.0)
    theta = math.acos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        rr = np.zeros(3)
        rr[0] = math.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        rr[1] = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        rr[2] = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(rr[0]) < abs(rr[1]) and abs(rr[0]) < abs(rr[2]) and (R[1, 2] > 0) != (rr[1] * rr[2] > 0):
            rr[2] = -rr[2]
        theta /= np.linalg.norm(rr)
        return rr * theta
    return r * (theta / (2 * s))


def project(X, rvec, tvec, K, want_jac=False):
    """cv::projectPoints without distortion.  Returns N x 2 (and d proj / d (r, t): 2N x 6)."""
    X = np.asarray(X, np.float64)
    if want_jac:
        R, dRdr = rodrigues_to_matrix(rvec, True)
    else:
        R = rodrigues_to_matrix(rvec)
    t = np.asarray(tvec, np.float64).ravel()
    Xc = X @ R.T + t
    z = np.where(Xc[:, 2] != 0, 1.0 / Xc[:, 2], 1.0)
    x, y = Xc[:, 0] * z, Xc[:, 1] * z
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    uv = np.stack([x * fx + cx, y * fy + cy], 1)
    if not want_jac:
        return uv
    n = len(X)
    J = np.zeros((2 * n, 6))
    # d(x, y) / d Xc
    dxdX = np.stack([z, np.zeros(n), -x * z], 1)
    dydX = np.stack([np.zeros(n), z, -y * z], 1)
    J[0::2, 3:] = fx * dxdX
    J[1::2, 3:] = fy * dydX
    # d Xc / d r_i = (dR/dr_i) X
    for i in range(3):
        dR = dRdr[i].reshape(3, 3)
        dXc = X @ dR.T
        J[0::2, i] = fx * (dxdX * dXc).sum(1)
        J[1::2, i] = fy * (dydX * dXc).sum(1)
    return uv, J


# --------------------------------------------------------------------------------------------------
def epnp(pws, us):
    """EPnP on normalised image coordinates (fu = fv = 1, uc = vc = 0).  Returns R, t."""
    pws = np.asarray(pws, np.float64)
    us = np.asarray(us, np.float64)
    n = len(pws)
    cws = np.zeros((4, 3))
    cws[0] = pws.sum(0) / n
    PW0 = pws - cws[0]
    dc, U, _ = jacobi_svd(PW0.T @ PW0)
    for i in range(1, 4):
        cws[i] = cws[0] + math.sqrt(dc[i - 1] / n) * U[:, i - 1]
    CC = (cws[1:] - cws[0]).T
    ci = np.linalg.inv(CC)
    a = (ci @ (pws - cws[0]).T).T
    alphas = np.concatenate([1.0 - a.sum(1, keepdims=True), a], 1)
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = alphas[:, j]
        M[0::2, 3 * j + 2] = alphas[:, j] * (0.0 - us[:, 0])
        M[1::2, 3 * j + 1] = alphas[:, j]
        M[1::2, 3 * j + 2] = alphas[:, j] * (0.0 - us[:, 1])
    _, Um, _ = jacobi_svd(M.T @ M)
    ut = Um.T                          # rows: singular vectors, descending singular value
    v = [ut[11], ut[10], ut[9], ut[8]]
    dv = np.zeros((4, 6, 3))
    for i in range(4):
        a_, b_ = 0, 1
        for j in range(6):
            dv[i, j] = v[i][3 * a_:3 * a_ + 3] - v[i][3 * b_:3 * b_ + 3]
            b_ += 1
            if b_ > 3:
                a_ += 1
                b_ = a_ + 1
    L = np.zeros((6, 10))
    for i in range(6):
        d = dv[:, i]
        L[i] = [d[0] @ d[0], 2 * d[0] @ d[1], d[1] @ d[1], 2 * d[0] @ d[2], 2 * d[1] @ d[2], d[2] @ d[2],
                2 * d[0] @ d[3], 2 * d[1] @ d[3], 2 * d[2] @ d[3], d[3] @ d[3]]

    def d2(p, q):
        return float(((p - q) ** 2).sum())

    rho = np.array([d2(cws[0], cws[1]), d2(cws[0], cws[2]), d2(cws[0], cws[3]),
                    d2(cws[1], cws[2]), d2(cws[1], cws[3]), d2(cws[2], cws[3])])

    def lstsq(A, b):
        return np.linalg.lstsq(A, b, rcond=None)[0]

    def approx1():
        b4 = lstsq(L[:, [0, 1, 3, 6]], rho)
        if b4[0] < 0:
            b0 = math.sqrt(-b4[0])
            return np.array([b0, -b4[1] / b0, -b4[2] / b0, -b4[3] / b0])
        b0 = math.sqrt(b4[0])
        return np.array([b0, b4[1] / b0, b4[2] / b0, b4[3] / b0])

    def approx2():
        b3 = lstsq(L[:, [0, 1, 2]], rho)
        if b3[0] < 0:
            b = [math.sqrt(-b3[0]), math.sqrt(-b3[2]) if b3[2] < 0 else 0.0]
        else:
            b = [math.sqrt(b3[0]), math.sqrt(b3[2]) if b3[2] > 0 else 0.0]
        if b3[1] < 0:
            b[0] = -b[0]
        return np.array([b[0], b[1], 0.0, 0.0])

    def approx3():
        b5 = lstsq(L[:, [0, 1, 2, 3, 4]], rho)
        if b5[0] < 0:
            b = [math.sqrt(-b5[0]), math.sqrt(-b5[2]) if b5[2] < 0 else 0.0]
        else:
            b = [math.sqrt(b5[0]), math.sqrt(b5[2]) if b5[2] > 0 else 0.0]
        if b5[1] < 0:
            b[0] = -b[0]
        return np.array([b[0], b[1], b5[3] / b[0], 0.0])

    def gauss_newton(b):
        b = b.copy()
        for _ in range(5):
            A = np.zeros((6, 4))
            bb = np.zeros(6)
            for i in range(6):
                l = L[i]
                A[i] = [2 * l[0] * b[0] + l[1] * b[1] + l[3] * b[2] + l[6] * b[3],
                        l[1] * b[0] + 2 * l[2] * b[1] + l[4] * b[2] + l[7] * b[3],
                        l[3] * b[0] + l[4] * b[1] + 2 * l[5] * b[2] + l[8] * b[3],
                        l[6] * b[0] + l[7] * b[1] + l[8] * b[2] + 2 * l[9] * b[3]]
                bb[i] = rho[i] - (l[0] * b[0] * b[0] + l[1] * b[0] * b[1] + l[2] * b[1] * b[1] + l[3] * b[0] * b[2] +
                                  l[4] * b[1] * b[2] + l[5] * b[2] * b[2] + l[6] * b[0] * b[3] + l[7] * b[1] * b[3] +
                                  l[8] * b[2] * b[3] + l[9] * b[3] * b[3])
            b = b + lstsq(A, bb)
        return b

    def pose(b):
        ccs = np.zeros((4, 3))
        for i in range(4):
            vv = ut[11 - i]
            for j in range(4):
                ccs[j] += b[i] * vv[3 * j:3 * j + 3]
        pcs = alphas @ ccs
        if pcs[0, 2] < 0:
            ccs, pcs = -ccs, -pcs
        pc0, pw0 = pcs.sum(0) / n, pws.sum(0) / n
        ABt = (pcs - pc0).T @ (pws - pw0)
        Ua, _, Vta = np.linalg.svd(ABt)
        R = Ua @ Vta
        if np.linalg.det(R) < 0:
            R[2] = -R[2]
        t = pc0 - R @ pw0
        Xc = pws @ R.T + t
        err = float(np.sqrt((us[:, 0] - Xc[:, 0] / Xc[:, 2]) ** 2 + (us[:, 1] - Xc[:, 1] / Xc[:, 2]) ** 2).sum() / n)
        return R, t, err

    cands = [pose(gauss_newton(f())) for f in (approx1, approx2, approx3)]
    N = 0
    if cands[1][2] < cands[0][2]:
        N = 1
    if cands[2][2] < cands[N][2]:
        N = 2
    return cands[N][0], cands[N][1]


def normalize(uv, K):
    """undistortPoints with zero distortion."""
    uv = np.asarray(uv, np.float64)
    return np.stack([(uv[:, 0] - K[0, 2]) / K[0, 0], (uv[:, 1] - K[1, 2]) / K[1, 1]], 1)


def epnp_pose(obj, img, K):
    """solvePnP(..., SOLVEPNP_EPNP): (rvec, tvec).  undistortPoints returns float for float image points: the
    normalised coordinates are rounded to f32 before EPnP sees them (verified: 1e-13 agreement with cv2 for n >= 6)."""
    R, t = epnp(np.asarray(obj, np.float64), normalize(img, K).astype(f32).astype(np.float64))
    return matrix_to_rodrigues(R), t


def reproj_errors(obj, img, rvec, tvec, K):
    """PnPRansacCallback::computeError: float squared distance between float projections and float image points."""
    proj = project(obj, rvec, tvec, K).astype(f32)
    d = (np.asarray(img, f32) - proj).astype(f32)
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(f32)


def sample_subsets(n: int, count: int, seed: int = 0xFFFFFFFFFFFFFFFF):
    """The first `count` 5-point samples of the registrator loop (duplicates inside a sample are redrawn)."""
    rng = CvRNG(seed)
    out = []
    for _ in range(count):
        idx = []
        while len(idx) < MODEL_POINTS:
            v = rng.uniform(n)
            if v not in idx:
                idx.append(v)
        out.append(idx)
    return np.array(out, np.int32)


# --------------------------------------------------------------------------------------------------
def dlt_init(obj, xn):
    """cvFindExtrinsicCameraParams2, non-planar branch: DLT on normalised points -> (rvec, tvec)."""
    n = len(obj)
    L = np.zeros((2 * n, 12))
    X = np.asarray(obj, np.float64)
    L[0::2, 0:3], L[0::2, 3] = X, 1
    L[0::2, 8:11], L[0::2, 11] = -xn[:, 0:1] * X, -xn[:, 0]
    L[1::2, 4:7], L[1::2, 7] = X, 1
    L[1::2, 8:11], L[1::2, 11] = -xn[:, 1:2] * X, -xn[:, 1]
    _, _, Vt = jacobi_svd(L.T @ L)
    RR = Vt[11].reshape(3, 4).copy()
    if np.linalg.det(RR[:, :3]) < 0:
        RR = -RR
    sc = np.linalg.norm(RR[:, :3])
    U, _, Vt3 = np.linalg.svd(RR[:, :3])
    R = U @ Vt3
    t = RR[:, 3] * (np.linalg.norm(R) / sc)
    return matrix_to_rodrigues(R), t


def is_planar(obj):
    X = np.asarray(obj, np.float64)
    Mc = X.mean(0)
    MM = (X - Mc).T @ (X - Mc)
    w = np.linalg.svd(MM, compute_uv=False)
    return w[2] / w[1] < 1e-3


def lm_refine(obj, img, K, rvec, tvec, max_iter=20, eps=FLT_EPS):
    """CvLevMarq as driven by cvFindExtrinsicCameraParams2 (6 parameters, residual = projection - measurement)."""
    obj = np.asarray(obj, np.float64)
    m = np.asarray(img, np.float64)
    param = np.concatenate([np.asarray(rvec, np.float64).ravel(), np.asarray(tvec, np.float64).ravel()])
    lambda_lg10 = -3
    iters = 0

    def step(JtJ, JtErr, prev):
        lam = math.exp(lambda_lg10 * math.log(10.0))
        A = JtJ.copy()
        A[np.diag_indices(6)] *= 1.0 + lam
        delta = np.linalg.lstsq(A, JtErr, rcond=None)[0]       # solve(..., DECOMP_SVD)
        return prev - delta

    while True:
        # CALC_J at param
        uv, J = project(obj, param[:3], param[3:], K, True)
        err = (uv - m).ravel()
        JtJ, JtErr = J.T @ J, J.T @ err
        prev = param.copy()
        prev_norm = float(np.linalg.norm(err))
        param = step(JtJ, JtErr, prev)
        while True:
            # CHECK_ERR at the new param
            e2 = float(np.linalg.norm((project(obj, param[:3], param[3:], K) - m).ravel()))
            if e2 > prev_norm:
                lambda_lg10 += 1
                if lambda_lg10 <= 16:
                    param = step(JtJ, JtErr, prev)
                    continue
            break
        lambda_lg10 = max(lambda_lg10 - 1, -16)
        iters += 1
        if iters >= max_iter or np.linalg.norm(param - prev) / np.linalg.norm(prev) < eps:
            return param[:3], param[3:]


def solve_pnp_iterative(obj, img, K):
    """solvePnP(..., SOLVEPNP_ITERATIVE) without an extrinsic guess, non-planar object (>= 6 points)."""
    xn = normalize(img, K)
    r, t = dlt_init(obj, xn)
    return lm_refine(obj, img, K, r, t)


def solve_pnp_ransac(obj, img, K, iterations=100, reproj_err=8.0, confidence=0.99, seed=0xFFFFFFFFFFFFFFFF):
    """Returns (ok, rvec, tvec, inlier_indices, info).  info: iterations run, winning iteration, best model."""
    obj32 = np.asarray(obj, f32).reshape(-1, 3)
    img32 = np.asarray(img, f32).reshape(-1, 2)
    n = len(obj32)
    assert n > MODEL_POINTS, "n == 4 / 5 take other OpenCV branches (P3P / direct solve)"
    t2 = f32(reproj_err * reproj_err)
    rng = CvRNG(seed)
    niters = iterations
    best, best_mask, best_count, win = None, np.zeros(n, bool), 0, -1
    it = 0
    while it < niters:
        idx = []
        while len(idx) < MODEL_POINTS:
            v = rng.uniform(n)
            if v not in idx:
                idx.append(v)
        r, t = epnp_pose(obj32[idx], img32[idx], K)
        if np.all(np.isfinite(r)) and np.all(np.isfinite(t)):
            mask = reproj_errors(obj32, img32, r, t, K) <= t2
            cnt = int(mask.sum())
            if cnt > max(best_count, MODEL_POINTS - 1):
                best, best_mask, best_count, win = (r, t), mask, cnt, it
                niters = update_iters(confidence, (n - cnt) / n, MODEL_POINTS, niters)
        it += 1
    info = {"iters": it, "win": win, "model": best}
    if best is None:
        return False, np.zeros(3), np.zeros(3), np.zeros(0, np.int32), info
    inl = np.nonzero(best_mask)[0].astype(np.int32)
    r, t = solve_pnp_iterative(obj32[inl].astype(np.float64), img32[inl].astype(np.float64), K)
    return True, r, t, inl, info
