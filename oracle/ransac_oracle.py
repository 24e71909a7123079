"""CPU oracle (numpy restatement) of the RANSAC two-view geometry the reference calls.

TEST INFRASTRUCTURE ONLY.
Reference call sites (all arithmetic is in un-vendored OpenCV; restated from its published algorithms and
pinned against cv2 4.13.0 golden vectors in tests/golden/ransac.npz):
  cv::findHomography(p1, p2, RANSAC, thr, mask)          src/initializer.cpp:82,  src/tracker.cpp:243
  cv::findFundamentalMat(p1, p2, FM_RANSAC, thr, .99)    src/initializer.cpp:87,  src/tracker.cpp:248
  cv::findEssentialMat(p1, p2, K, RANSAC, .99, 1.0)      src/initializer.cpp:228-229
  cv::recoverPose(E, p1, p2, K, R, t, mask)              src/initializer.cpp:236
  cv::triangulatePoints(P0, P1, p0, p1, X)               src/initializer.cpp:125, src/tracker.cpp:149
Spec: SURVEY.md Appendix A.4.  The sequential OpenCV loop (fresh RNG(2^64-1) per call, redraw rules,
strict '>' best-model update, shrinking niters) is reproduced exactly; because the RNG stream depends only
on the data-only subset checks, all subsets can also be generated up front (sample_subsets) -- that is what
the CUDA path does before it solves and scores every hypothesis in parallel.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
FLT_EPSILON = float(np.finfo(np.float32).eps)
DBL_EPSILON = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)
MWC_A = 4164903690


# --------------------------------------------------------------------------------------------------
class CvRNG:
    """cv::RNG multiply-with-carry generator."""

    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF):
        self.state = state & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * MWC_A + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, n: int) -> int:
        return self.next() % n


def rng_stream(count: int, seed: int = 0xFFFFFFFFFFFFFFFF) -> np.ndarray:
    r = CvRNG(seed)
    return np.array([r.next() for _ in range(count)], np.uint32)


# --------------------------------------------------------------------------------------------------
def _have_collinear(pts: np.ndarray) -> bool:
    """haveCollinearPoints(m, count): is the LAST point on a line through two earlier ones (float diffs)."""
    p = pts.astype(f32)
    i = len(p) - 1
    for j in range(i):
        dx1 = float(f32(p[j, 0] - p[i, 0]))
        dy1 = float(f32(p[j, 1] - p[i, 1]))
        for k in range(j):
            dx2 = float(f32(p[k, 0] - p[i, 0]))
            dy2 = float(f32(p[k, 1] - p[i, 1]))
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _det3(a):
    return (a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2]) +
            a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]))


def check_subset(model: str, s1: np.ndarray, s2: np.ndarray) -> bool:
    if model == "E":
        return True
    if _have_collinear(s1) or _have_collinear(s2):
        return False
    if model == "H":
        neg = 0
        for t in ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3)):
            A = [[float(s1[i, 0]), float(s1[i, 1]), 1.0] for i in t]
            B = [[float(s2[i, 0]), float(s2[i, 1]), 1.0] for i in t]
            neg += 1 if _det3(A) * _det3(B) < 0 else 0
        if neg not in (0, 4):
            return False
    return True


MODEL_POINTS = {"H": 4, "F": 7, "E": 5}
MAX_ITERS = {"H": 2000, "F": 1000, "E": 1000}


def get_subset(model: str, p1, p2, rng: CvRNG, max_attempts: int = 10000):
    k = MODEL_POINTS[model]
    n = len(p1)
    for _ in range(max_attempts):
        idx = []
        for i in range(k):
            while True:
                v = rng.uniform(n)
                if v not in idx:
                    break
            idx.append(v)
        if check_subset(model, p1[idx], p2[idx]):
            return idx
    return None


def sample_subsets(model: str, p1, p2, count: int, seed: int = 0xFFFFFFFFFFFFFFFF) -> np.ndarray:
    """The first `count` minimal samples the OpenCV loop would draw (no early exit)."""
    rng = CvRNG(seed)
    out = []
    for _ in range(count):
        s = get_subset(model, p1, p2, rng)
        if s is None:
            break
        out.append(s)
    return np.array(out, np.int32).reshape(-1, MODEL_POINTS[model])


# --------------------------------------------------------------------------------------------------
# minimal / least-squares solvers
def h_kernel(m1: np.ndarray, m2: np.ndarray):
    """HomographyEstimatorCallback::runKernel: per-axis L1-normalised DLT, smallest eigenvector of LtL."""
    M = m1.astype(np.float64)
    m = m2.astype(np.float64)
    n = len(M)
    cM, cm = M.mean(0), m.mean(0)
    sM = np.abs(M - cM).sum(0)
    sm = np.abs(m - cm).sum(0)
    if (np.abs(sM) < DBL_EPSILON).any() or (np.abs(sm) < DBL_EPSILON).any():
        return None
    sM, sm = n / sM, n / sm
    invHnorm = np.array([[1 / sm[0], 0, cm[0]], [0, 1 / sm[1], cm[1]], [0, 0, 1]])
    Hnorm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1]])
    x, y = (m[:, 0] - cm[0]) * sm[0], (m[:, 1] - cm[1]) * sm[1]
    X, Y = (M[:, 0] - cM[0]) * sM[0], (M[:, 1] - cM[1]) * sM[1]
    z, o = np.zeros(n), np.ones(n)
    Lx = np.stack([X, Y, o, z, z, z, -x * X, -x * Y, -x], 1)
    Ly = np.stack([z, z, z, X, Y, o, -y * X, -y * Y, -y], 1)
    LtL = Lx.T @ Lx + Ly.T @ Ly
    w, V = np.linalg.eigh(LtL)
    H0 = V[:, 0].reshape(3, 3)
    H = invHnorm @ H0 @ Hnorm2
    if H[2, 2] == 0:
        return None
    return H / H[2, 2]


def _solve_cubic_real(c):
    """Real roots of c[0] x^3 + c[1] x^2 + c[2] x + c[3] (degenerating like cv::solveCubic)."""
    c = np.asarray(c, np.float64)
    if c[0] == 0:
        if c[1] == 0:
            if c[2] == 0:
                return []
            return [-c[3] / c[2]]
        d = c[2] * c[2] - 4 * c[1] * c[3]
        if d < 0:
            return []
        d = np.sqrt(d)
        return [(-c[2] + d) / (2 * c[1]), (-c[2] - d) / (2 * c[1])]
    r = np.roots(c)
    return [float(v.real) for v in r if abs(v.imag) < 1e-9 * max(1.0, abs(v.real))]


def f7_kernel(m1: np.ndarray, m2: np.ndarray):
    """FMEstimatorCallback::run7Point: 1..3 fundamental matrices, each scaled to F[2,2] = 1."""
    a1 = m1.astype(np.float64)
    a2 = m2.astype(np.float64)
    c1, c2 = a1.mean(0), a2.mean(0)
    s1 = np.sqrt(((a1 - c1) ** 2).sum(1)).mean()
    s2 = np.sqrt(((a2 - c2) ** 2).sum(1)).mean()
    if s1 < FLT_EPSILON or s2 < FLT_EPSILON:
        return []
    s1, s2 = np.sqrt(2.0) / s1, np.sqrt(2.0) / s2
    x0, y0 = (a1[:, 0] - c1[0]) * s1, (a1[:, 1] - c1[1]) * s1
    x1, y1 = (a2[:, 0] - c2[0]) * s2, (a2[:, 1] - c2[1]) * s2
    A = np.stack([x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, np.ones(7)], 1)
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    f1, f2 = Vt[7].copy(), Vt[8].copy()
    f1 -= f2                               # F = lambda*f1' + f2 with f1' = f1 - f2

    def det_poly(fa, fb):
        # coefficients of det(l*fa + fb) in l via 4 evaluations
        ls = np.array([0.0, 1.0, -1.0, 2.0])
        v = [np.linalg.det((l * fa + fb).reshape(3, 3)) for l in ls]
        return np.linalg.solve(np.vander(ls, 4), v)   # [l^3, l^2, l, 1]

    roots = _solve_cubic_real(det_poly(f1, f2))
    T1 = np.array([[s1, 0, -s1 * c1[0]], [0, s1, -s1 * c1[1]], [0, 0, 1]])
    T2 = np.array([[s2, 0, -s2 * c2[0]], [0, s2, -s2 * c2[1]], [0, 0, 1]])
    out = []
    for lam in roots:
        mu = 1.0
        s = f1[8] * lam + f2[8]
        F = np.zeros(9)
        if abs(s) > DBL_EPSILON:
            mu = 1.0 / s
            lam_s = lam * mu
            F[8] = 1.0
        else:
            lam_s = lam
            F[8] = 0.0
        F[:8] = f1[:8] * lam_s + f2[:8] * mu
        F = T2.T @ F.reshape(3, 3) @ T1
        if abs(F[2, 2]) > FLT_EPSILON:
            F = F / F[2, 2]
        out.append(F)
    return out


# ---- 5-point (Nister) ---------------------------------------------------------------------------
def _pmul(a, b):
    """product of polynomials in (x, y, z) stored as coefficient cubes [i, j, k]."""
    out = np.zeros((a.shape[0] + b.shape[0] - 1,) * 3)
    for i, j, k in zip(*np.nonzero(a)):
        out[i:i + b.shape[0], j:j + b.shape[1], k:k + b.shape[2]] += a[i, j, k] * b
    return out


_MONO = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0), (1, 1, 1), (1, 1, 0),
         (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2), (0, 0, 1), (0, 0, 0)]


def e5_kernel(q1: np.ndarray, q2: np.ndarray):
    """EMEstimatorCallback::runKernel (Nister 5-point): list of unit-Frobenius-norm essential matrices.
    q1, q2: 5 x 2 float64 K-normalised points; the constraint is q2^T E q1 = 0."""
    x1, y1 = q1[:, 0], q1[:, 1]
    x2, y2 = q2[:, 0], q2[:, 1]
    A = np.stack([x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, np.ones(5)], 1)
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    EE = Vt[5:9]                                       # 4 x 9 basis, E = x*EE0 + y*EE1 + z*EE2 + EE3
    # entries of E as linear polynomials in x, y, z
    Ep = np.zeros((3, 3, 2, 2, 2))
    for r in range(3):
        for c in range(3):
            Ep[r, c, 1, 0, 0] = EE[0, 3 * r + c]
            Ep[r, c, 0, 1, 0] = EE[1, 3 * r + c]
            Ep[r, c, 0, 0, 1] = EE[2, 3 * r + c]
            Ep[r, c, 0, 0, 0] = EE[3, 3 * r + c]
    # EEt = E E^T (quadratic), trace
    EEt = [[sum(_pmul(Ep[r, k], Ep[c, k]) for k in range(3)) for c in range(3)] for r in range(3)]
    tr = EEt[0][0] + EEt[1][1] + EEt[2][2]
    eqs = []
    # det(E) = 0
    det = (_pmul(Ep[0, 0], _pmul(Ep[1, 1], Ep[2, 2]) - _pmul(Ep[1, 2], Ep[2, 1])) -
           _pmul(Ep[0, 1], _pmul(Ep[1, 0], Ep[2, 2]) - _pmul(Ep[1, 2], Ep[2, 0])) +
           _pmul(Ep[0, 2], _pmul(Ep[1, 0], Ep[2, 1]) - _pmul(Ep[1, 1], Ep[2, 0])))
    eqs.append(det)
    # 2 E E^T E - tr(E E^T) E = 0
    for r in range(3):
        for c in range(3):
            t = sum(_pmul(EEt[r][k], Ep[k, c]) for k in range(3))
            eqs.append(2 * t - _pmul(tr, Ep[r, c]))
    M = np.array([[e[m] for m in _MONO] for e in eqs])           # 10 x 20
    try:
        G = np.linalg.solve(M[:, :10], M[:, 10:])                 # reduced rows: mono_i + G[i] . tail = 0
    except np.linalg.LinAlgError:
        return []
    # rows 4..9: x^2 z, x^2, y^2 z, y^2, x y z, x y ; tail = [x z^2, x z, x, y z^2, y z, y, z^3, z^2, z, 1]
    def row_polys(r):
        g = G[r]
        return (np.array([g[0], g[1], g[2]]), np.array([g[3], g[4], g[5]]), np.array([g[6], g[7], g[8], g[9]]))

    B = []
    for ra, rb in ((4, 5), (6, 7), (8, 9)):
        ax, ay, a1 = row_polys(ra)
        bx, by, b1 = row_polys(rb)
        # <row a> - z * <row b>, polynomials in z, highest power first
        B.append((np.concatenate([[0.0], ax]) - np.concatenate([bx, [0.0]]),
                  np.concatenate([[0.0], ay]) - np.concatenate([by, [0.0]]),
                  np.concatenate([[0.0], a1]) - np.concatenate([b1, [0.0]])))
    pm = np.polymul
    detp = (pm(B[0][0], np.polysub(pm(B[1][1], B[2][2]), pm(B[1][2], B[2][1]))) -
            pm(B[0][1], np.polysub(pm(B[1][0], B[2][2]), pm(B[1][2], B[2][0]))) +
            pm(B[0][2], np.polysub(pm(B[1][0], B[2][1]), pm(B[1][1], B[2][0]))))
    roots = np.roots(detp)
    out = []
    for rt in roots:
        if abs(rt.imag) > 1e-10:
            continue
        z = float(rt.real)
        Bz = np.array([[np.polyval(B[i][j], z) for j in range(3)] for i in range(3)])
        _, _, vt = np.linalg.svd(Bz)
        xy1 = vt[2]
        if abs(xy1[2]) < 1e-10:
            continue
        x, y = xy1[0] / xy1[2], xy1[1] / xy1[2]
        E = x * EE[0] + y * EE[1] + z * EE[2] + EE[3]
        E = E / np.linalg.norm(E)
        out.append(E.reshape(3, 3))
    return out


# --------------------------------------------------------------------------------------------------
# errors
def h_errors(H, p1, p2) -> np.ndarray:
    """float32 arithmetic, non-fused, exactly as HomographyEstimatorCallback::computeError."""
    Hf = H.astype(f32).ravel()
    M = p1.astype(f32)
    m = p2.astype(f32)
    one = f32(1.0)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        ww = one / ((Hf[6] * M[:, 0]).astype(f32) + (Hf[7] * M[:, 1]).astype(f32) + one).astype(f32)
    dx = (((Hf[0] * M[:, 0]).astype(f32) + (Hf[1] * M[:, 1]).astype(f32)).astype(f32) + Hf[2]).astype(f32) * ww - m[:, 0]
    dy = (((Hf[3] * M[:, 0]).astype(f32) + (Hf[4] * M[:, 1]).astype(f32)).astype(f32) + Hf[5]).astype(f32) * ww - m[:, 1]
    dx, dy = dx.astype(f32), dy.astype(f32)
    return ((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32)


def f_errors(F, p1, p2) -> np.ndarray:
    F = np.asarray(F, np.float64).ravel()
    x1, y1 = p1[:, 0].astype(np.float64), p1[:, 1].astype(np.float64)
    x2, y2 = p2[:, 0].astype(np.float64), p2[:, 1].astype(np.float64)
    a = F[0] * x1 + F[1] * y1 + F[2]
    b = F[3] * x1 + F[4] * y1 + F[5]
    c = F[6] * x1 + F[7] * y1 + F[8]
    s2 = 1.0 / (a * a + b * b)
    d2 = x2 * a + y2 * b + c
    a = F[0] * x2 + F[3] * y2 + F[6]
    b = F[1] * x2 + F[4] * y2 + F[7]
    c = F[2] * x2 + F[5] * y2 + F[8]
    s1 = 1.0 / (a * a + b * b)
    d1 = x1 * a + y1 * b + c
    return np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(f32)


def e_errors(E, q1, q2) -> np.ndarray:
    """Sampson error on K-normalised float64 points."""
    E = np.asarray(E, np.float64).reshape(3, 3)
    X1 = np.column_stack([q1, np.ones(len(q1))])
    X2 = np.column_stack([q2, np.ones(len(q2))])
    Ex1 = X1 @ E.T
    Etx2 = X2 @ E
    x2tEx1 = (X2 * Ex1).sum(1)
    den = Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2
    return (x2tEx1 * x2tEx1 / den).astype(f32)


def update_iters(p: float, ep: float, k: int, max_iters: int) -> int:
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    den = 1.0 - (1.0 - ep) ** k
    if den < DBL_MIN:
        return 0
    num, den = np.log(num), np.log(den)
    if den >= 0 or -num >= max_iters * (-den):
        return max_iters
    return int(np.rint(num / den))


# --------------------------------------------------------------------------------------------------
def ransac(model: str, p1, p2, thr: float, conf: float, seed: int = 0xFFFFFFFFFFFFFFFF, pts_solver=None):
    """RANSACPointSetRegistrator::run.  p1/p2 are the arrays scored (f32 pixels for H/F, f64 normalised for E).
    Returns (best_model or None, mask u8, n_iterations_run, winning_iteration)."""
    k = MODEL_POINTS[model]
    n = len(p1)
    niters = MAX_ITERS[model]
    kernel = {"H": lambda a, b: [m for m in [h_kernel(a, b)] if m is not None], "F": f7_kernel, "E": e5_kernel}[model]
    errfn = {"H": h_errors, "F": f_errors, "E": e_errors}[model]
    t = f32(thr * thr)
    rng = CvRNG(seed)
    best, best_mask, best_count, win = None, np.zeros(n, np.uint8), 0, -1
    it = 0
    while it < niters:
        idx = get_subset(model, p1, p2, rng)
        if idx is None:
            break
        for mdl in kernel(p1[idx], p2[idx]):
            mask = errfn(mdl, p1, p2) <= t
            cnt = int(mask.sum())
            if cnt > max(best_count, k - 1):
                best, best_mask, best_count, win = mdl, mask.astype(np.uint8), cnt, it
                niters = update_iters(conf, (n - cnt) / n, k, niters)
        it += 1
    return best, best_mask, it, win


# ---- homography refinement: DLT on inliers + 10 LM iterations (cv::LMSolver) -------------------
def _h_residual_jac(h8, M, m, want_j=True):
    Mx, My = M[:, 0], M[:, 1]
    ww = h8[6] * Mx + h8[7] * My + 1.0
    ww = np.where(np.abs(ww) > DBL_EPSILON, 1.0 / ww, 0.0)
    xi = (h8[0] * Mx + h8[1] * My + h8[2]) * ww
    yi = (h8[3] * Mx + h8[4] * My + h8[5]) * ww
    r = np.empty(2 * len(M))
    r[0::2] = xi - m[:, 0]
    r[1::2] = yi - m[:, 1]
    if not want_j:
        return r, None
    J = np.zeros((2 * len(M), 8))
    J[0::2, 0], J[0::2, 1], J[0::2, 2] = Mx * ww, My * ww, ww
    J[0::2, 6], J[0::2, 7] = -Mx * ww * xi, -My * ww * xi
    J[1::2, 3], J[1::2, 4], J[1::2, 5] = Mx * ww, My * ww, ww
    J[1::2, 6], J[1::2, 7] = -Mx * ww * yi, -My * ww * yi
    return r, J


def _solve_sym(A, b):
    """cv::solve(..., DECOMP_EIG): pseudo-inverse through the symmetric eigen-decomposition."""
    w, V = np.linalg.eigh(A)
    winv = np.where(np.abs(w) > DBL_EPSILON * np.abs(w).max() * 0 + 1e-300, 1.0 / w, 0.0)
    return V @ (winv * (V.T @ b))


def lm_refine_h(H, p1, p2, max_iters: int = 10, eps: float = FLT_EPSILON):
    M = p1.astype(np.float64)
    m = p2.astype(np.float64)
    x = (H / H[2, 2]).ravel()[:8].copy()
    r, J = _h_residual_jac(x, M, m)
    S = float(r @ r)
    A = J.T @ J
    v = J.T @ r
    D = np.diag(A).copy()
    Rlo, Rhi = 0.25, 0.75
    lam, lc = 1.0, 0.75
    it = 0
    while True:
        Ap = A + np.diag(lam * D)
        d = _solve_sym(Ap, v)
        xd = x - d
        rd, _ = _h_residual_jac(xd, M, m, want_j=False)
        Sd = float(rd @ rd)
        temp_d = 2 * v - A @ d
        dS = float(d @ temp_d)
        R = (S - Sd) / (dS if abs(dS) > DBL_EPSILON else 1.0)
        if R > Rhi:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif R < Rlo:
            t = float(d @ v)
            nu = (Sd - S) / (t if abs(t) > DBL_EPSILON else 1.0) + 2.0
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0:
                Ainv = np.linalg.pinv(A)
                maxval = max(DBL_EPSILON, np.abs(np.diag(Ainv)).max())
                lam = lc = 1.0 / maxval
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S = Sd
            x = xd
            r, J = _h_residual_jac(x, M, m)
            A = J.T @ J
            v = J.T @ r
        it += 1
        if not (it < max_iters and np.abs(d).max() >= eps and np.abs(r).max() >= eps):
            break
    return np.append(x, 1.0).reshape(3, 3)


def find_homography(p1, p2, thr: float, seed: int = 0xFFFFFFFFFFFFFFFF):
    """cv::findHomography(p1, p2, RANSAC, thr): returns (H, mask, iterations)."""
    p1 = np.asarray(p1, f32)
    p2 = np.asarray(p2, f32)
    H, mask, iters, _ = ransac("H", p1, p2, thr, 0.995, seed)
    if H is None:
        return None, np.zeros(len(p1), np.uint8), iters
    inl = mask.astype(bool)
    H = h_kernel(p1[inl], p2[inl])
    H = lm_refine_h(H, p1[inl], p2[inl])
    mask = (h_errors(H, p1, p2) <= f32(thr * thr)).astype(np.uint8)     # 4.13: mask of the refined H
    return H, mask, iters


def lmeds(model: str, p1, p2, conf: float, seed: int = 0xFFFFFFFFFFFFFFFF):
    """LMeDSPointSetRegistrator::run (OpenCV modules/calib3d/src/ptsetreg.cpp): the same subsets as RANSAC, the model
    with the least median error wins (strict '<', models of one sample in solver order), then
    sigma = 2.5 * 1.4826 * (1 + 5 / (N - k)) * sqrt(median), at least 0.001; inliers = err <= sigma^2.
    Returns (model or None, mask u8, iterations run, success)."""
    k = MODEL_POINTS[model]
    n = len(p1)
    kernel = {"F": f7_kernel}[model]
    errfn = {"F": f_errors}[model]
    niters = max(update_iters(conf, 0.45, k, MAX_ITERS[model]), 3)
    rng = CvRNG(seed)
    best, min_median = None, np.inf
    it = 0
    while it < niters:
        idx = get_subset(model, p1, p2, rng)
        if idx is None:
            break
        for mdl in kernel(p1[idx], p2[idx]):
            err = np.sort(errfn(mdl, p1, p2))
            median = float(err[n // 2])
            if median < min_median:
                min_median, best = median, mdl
        it += 1
    if best is None:
        return None, np.zeros(n, np.uint8), it, False
    sigma = max(2.5 * 1.4826 * (1 + 5.0 / (n - k)) * np.sqrt(min_median), 0.001)
    mask = (errfn(best, p1, p2) <= f32(sigma * sigma)).astype(np.uint8)
    return best, mask, it, int(mask.sum()) >= k


def find_fundamental(p1, p2, thr: float, conf: float = 0.99, seed: int = 0xFFFFFFFFFFFFFFFF):
    """cv::findFundamentalMat(p1, p2, FM_RANSAC, thr, conf, mask): N >= 15 RANSAC; 8 <= N < 15 silently LMedS
    (the Tracker can get there: /root/reference/src/tracker.cpp:239-248 with min_tracked_points = 10); N == 7 the
    7-point solution itself with an all-ones mask; N < 7 nothing.  Returns (F or None, mask, iterations)."""
    p1 = np.asarray(p1, f32)
    p2 = np.asarray(p2, f32)
    n = len(p1)
    if n < 7:
        return None, np.zeros(n, np.uint8), 0
    if n == 7:
        models = f7_kernel(p1, p2)
        return (models[0] if models else None), np.ones(n, np.uint8), 1
    if n < 15:
        F, mask, iters, ok = lmeds("F", p1, p2, conf, seed)
        return (F if ok else None), mask, iters
    F, mask, iters, _ = ransac("F", p1, p2, thr, conf, seed)
    return F, mask, iters


def normalize_points(p, K):
    p = np.asarray(p, np.float64)
    return np.stack([(p[:, 0] - K[0, 2]) / K[0, 0], (p[:, 1] - K[1, 2]) / K[1, 1]], 1)


def find_essential(p1, p2, K, conf: float = 0.99, thr: float = 1.0, seed: int = 0xFFFFFFFFFFFFFFFF):
    K = np.asarray(K, np.float64)
    q1, q2 = normalize_points(p1, K), normalize_points(p2, K)
    t = thr / ((K[0, 0] + K[1, 1]) / 2.0)
    E, mask, iters, _ = ransac("E", q1, q2, t, conf, seed)
    return E, mask, iters


# --------------------------------------------------------------------------------------------------
def triangulate(P0, P1, p0, p1) -> np.ndarray:
    """cv::triangulatePoints: 4 x N homogeneous points (unit columns, sign unspecified)."""
    P0 = np.asarray(P0, np.float64)
    P1 = np.asarray(P1, np.float64)
    p0 = np.asarray(p0, np.float64)
    p1 = np.asarray(p1, np.float64)
    out = np.zeros((4, len(p0)))
    for i in range(len(p0)):
        A = np.stack([p0[i, 0] * P0[2] - P0[0], p0[i, 1] * P0[2] - P0[1],
                      p1[i, 0] * P1[2] - P1[0], p1[i, 1] * P1[2] - P1[1]])
        _, _, vt = np.linalg.svd(A)
        out[:, i] = vt[3]
    return out


def decompose_essential(E):
    U, _, Vt = np.linalg.svd(np.asarray(E, np.float64).reshape(3, 3))
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 1.0]])
    return U @ W @ Vt, U @ W.T @ Vt, U[:, 2].copy()


def recover_pose(E, p1, p2, K, mask=None, dist: float = 50.0):
    """cv::recoverPose(E, p1, p2, K, R, t, mask): returns (R, t, mask, good)."""
    K = np.asarray(K, np.float64)
    q1, q2 = normalize_points(p1, K), normalize_points(p2, K)
    R1, R2, t = decompose_essential(E)
    P0 = np.eye(3, 4)
    masks, cands = [], [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    for R, tt in cands:
        P = np.column_stack([R, tt])
        Q = triangulate(P0, P, q1, q2)
        m = Q[2] * Q[3] > 0
        Q = Q / Q[3]
        m &= Q[2] < dist
        Q2 = P @ Q
        m &= (Q2[2] > 0) & (Q2[2] < dist)
        if mask is not None:
            m &= np.asarray(mask).ravel() != 0
        masks.append(m)
    good = [int(m.sum()) for m in masks]
    for i in range(4):
        if all(good[i] >= g for g in good):
            R, tt = cands[i]
            return R, tt, masks[i].astype(np.uint8), good[i]
    raise AssertionError
