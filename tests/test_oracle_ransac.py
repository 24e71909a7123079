"""Pin the RANSAC / two-view geometry oracle against cv2 4.13.0 golden vectors."""
import numpy as np
import pytest

from conftest import load_golden, sha
from oracle import ransac_oracle as ro
from oracle import synth


def _scene(g, tag):
    n, seed, planar = g[f"{tag}_args"].tolist()
    p1, p2, R, t, inl = synth.scene_correspondences(n, seed, planar=bool(planar), outlier_frac=0.25)
    assert sha(np.concatenate([p1, p2])) == str(g[f"{tag}_sha"]), "synthetic generator drifted"
    return p1, p2


def _same_up_to_sign(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    return min(np.abs(a - b).max(), np.abs(a + b).max()) < tol


def test_rng_stream_known_answer():
    # cv::RNG(0xffffffffffffffff): first outputs of the MWC generator
    r = ro.CvRNG()
    first = [r.next() for _ in range(4)]
    s = 0xFFFFFFFFFFFFFFFF
    exp = []
    for _ in range(4):
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        exp.append(s & 0xFFFFFFFF)
    assert first == exp
    # closed form: state_n == A^n * state_0 (mod A*2^32 - 1) -- the jump-ahead the CUDA sampler relies on
    m = 4164903690 * (1 << 32) - 1
    r = ro.CvRNG()
    for _ in range(50):
        r.next()
    assert r.state % m == (pow(4164903690, 50, m) * (0xFFFFFFFFFFFFFFFF % m)) % m


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_ransac_models_and_masks(tag):
    g = load_golden("ransac.npz")
    K = g["K"]
    p1, p2 = _scene(g, tag)
    H, mh, _ = ro.find_homography(p1, p2, 1.0)
    assert np.array_equal(mh, g[f"{tag}_mask_h"])
    assert np.abs(H - g[f"{tag}_H"]).max() < 1e-6
    F, mf, _ = ro.find_fundamental(p1, p2, 1.0, 0.99)
    assert np.array_equal(mf, g[f"{tag}_mask_f"])
    assert np.abs(F - g[f"{tag}_F"]).max() < 1e-9
    E, me, _ = ro.find_essential(p1, p2, K, 0.99, 1.0)
    assert np.array_equal(me, g[f"{tag}_mask_e"])
    assert _same_up_to_sign(E, g[f"{tag}_E"], 1e-6)
    R, t, mp, good = ro.recover_pose(g[f"{tag}_E"], p1, p2, K, mask=g[f"{tag}_mask_e"])
    assert good == int(g[f"{tag}_good"])
    assert np.array_equal(mp != 0, g[f"{tag}_mask_pose"] != 0)
    assert np.abs(R - g[f"{tag}_R"]).max() < 1e-12 and np.abs(t - g[f"{tag}_t"]).max() < 1e-12
    X = ro.triangulate(K @ np.eye(3, 4), K @ np.column_stack([g[f"{tag}_R"], g[f"{tag}_t"]]), p1, p2)
    Xg = g[f"{tag}_X"].astype(np.float64)
    ok = np.abs(Xg[3]) > 1e-3
    a, b = (X[:3] / X[3])[:, ok], (Xg[:3] / Xg[3])[:, ok]
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-5


def test_minimal_solvers():
    g = load_golden("ransac.npz")
    K = g["K"]
    p1, p2 = g["hyp_p1"], g["hyp_p2"]
    for j, i in enumerate(g["hyp_i4"]):
        assert np.abs(ro.h_kernel(p1[i], p2[i]) - g["hyp_H"][j]).max() < 1e-9 * np.abs(g["hyp_H"][j]).max() + 1e-9
    for j, i in enumerate(g["hyp_i7"]):
        ref = g[f"hyp_F{j}"]
        got = ro.f7_kernel(p1[i], p2[i])
        assert len(got) == len(ref)
        for f in ref:
            assert min(np.abs(f - o).max() for o in got) < 1e-9
    q1, q2 = ro.normalize_points(p1, K), ro.normalize_points(p2, K)
    for j, i in enumerate(g["hyp_i5"]):
        ref = g[f"hyp_E{j}"]
        got = ro.e5_kernel(q1[i], q2[i])
        assert len(got) == len(ref)
        for e in ref:
            # spurious roots of the degree-10 polynomial can be ill-conditioned: 1e-4 still separates solutions
            assert min(min(np.abs(e - o).max(), np.abs(e + o).max()) for o in got) < 1e-4


def test_sample_subsets_prefix_of_sequential_loop():
    p1, p2, *_ = synth.scene_correspondences(300, 3)
    for model in ("H", "F", "E"):
        a = ro.sample_subsets(model, p1, p2, 40)
        rng = ro.CvRNG()
        b = [ro.get_subset(model, p1, p2, rng) for _ in range(40)]
        assert np.array_equal(a, np.array(b))
        assert all(len(set(r)) == len(r) for r in a.tolist())


def test_update_iters():
    assert ro.update_iters(0.99, 0.0, 7, 1000) == 0
    assert ro.update_iters(0.99, 1.0, 7, 1000) == 1000
    assert ro.update_iters(0.995, 0.3, 4, 2000) == int(np.rint(np.log(0.005) / np.log(1 - 0.7 ** 4)))
    assert ro.update_iters(0.99, 0.9, 5, 17) == 17            # never grows past the current cap


def test_find_fundamental_below_15_points_vs_cv2():
    """cv::findFundamentalMat(FM_RANSAC) silently changes algorithm below 15 correspondences (LMedS; the 7-point solution
    itself at N == 7).  N == 14: mask and F identical to cv2.  8 <= N <= 13: the median of a model that fits its 7
    sample points exactly is numerically zero for EVERY sample, so which sample wins is rounding noise inside OpenCV
    too -- but the result is always "exactly one minimal sample": 7 inliers, which is all the reference consumes
    (countNonZero, src/tracker.cpp:249)."""
    g = load_golden("f_small.npz")
    for n, seed in g["cases"].tolist():
        p1, p2, *_ = synth.scene_correspondences(n, seed, outlier_frac=0.2, noise_px=0.3)
        assert sha(np.concatenate([p1, p2])) == str(g[f"n{n}_s{seed}_sha"])
        F, mask, iters = ro.find_fundamental(p1, p2, 1.0, 0.99)
        ref_mask, ref_F = g[f"n{n}_s{seed}_mask"], g[f"n{n}_s{seed}_F"]
        assert int(mask.sum()) == int(ref_mask.sum())
        if n == 7:
            assert mask.all() and iters == 1
            assert min(np.abs(F - f).max() for f in ref_F.reshape(-1, 3, 3)) < 1e-8
        elif n == 14:
            assert iters == 300 and np.array_equal(mask, ref_mask)
            assert np.abs(F - ref_F).max() < 1e-8
        else:
            assert int(mask.sum()) == 7
    assert ro.find_fundamental(p1[:6], p2[:6], 1.0)[0] is None
