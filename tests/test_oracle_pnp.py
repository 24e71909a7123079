"""Pin the solvePnPRansac oracle (oracle/pnp_oracle.py) against cv2 4.13.0 golden vectors (tests/golden/pnp.npz).

Reference call site: /root/reference/src/tracker.cpp:309 (100 iterations, 8 px, confidence 0.99)."""
import numpy as np
import pytest

from conftest import load_golden, sha
from oracle import pnp_oracle as po
from oracle import synth

CASES = ["easy", "outl30", "outl50", "noisy", "few"]


def _scene(g, tag):
    n, seed, noise, outl, planar = g[f"{tag}_args"].tolist()
    obj, img, K, _, _ = synth.pnp_scene(int(n), int(seed), noise, outl, bool(planar))
    assert sha(np.concatenate([obj.ravel(), img.ravel()])) == str(g[f"{tag}_sha"]), "synthetic generator drifted"
    return obj, img, K


@pytest.mark.parametrize("tag", CASES)
def test_pnp_ransac_oracle_vs_cv2(tag):
    g = load_golden("pnp.npz")
    obj, img, K = _scene(g, tag)
    ok, r, t, inl, info = po.solve_pnp_ransac(obj, img, K)
    assert ok == bool(g[f"{tag}_ok"])
    assert np.array_equal(inl, g[f"{tag}_inliers"])                 # the winning hypothesis' inlier set, in order
    assert np.abs(r - g[f"{tag}_rvec"]).max() < 1e-7
    assert np.abs(t - g[f"{tag}_tvec"]).max() < 1e-7


def test_epnp_and_iterative_vs_cv2():
    g = load_golden("pnp.npz")
    obj, img, K, _, _ = synth.pnp_scene(400, 9, 0.5, 0.0)
    for i, want in zip(g["hyp_i6"], g["hyp_epnp"]):
        r, t = po.epnp_pose(obj[i], img[i], K)
        assert np.abs(np.concatenate([r, t]) - want).max() < 1e-9   # needs cv::SVD's signs for the control axes
    r, t = po.solve_pnp_iterative(obj.astype(np.float64), img.astype(np.float64), K)
    assert np.abs(np.concatenate([r, t]) - g["iter_rt"]).max() < 1e-8


def test_jacobi_svd_signs_and_sampler():
    rng = np.random.default_rng(0)
    P = rng.normal(size=(30, 3)) * [5, 1, 0.2]
    w, U, Vt = po.jacobi_svd(P.T @ P)
    assert np.allclose(U @ np.diag(w) @ Vt, P.T @ P, atol=1e-9) and np.all(np.diff(w) <= 0)
    s = po.sample_subsets(50, 20)
    assert s.shape == (20, 5) and all(len(set(r)) == 5 for r in s.tolist()) and s.max() < 50
    # Rodrigues round trip
    r = np.array([0.3, -0.2, 0.5])
    assert np.abs(po.matrix_to_rodrigues(po.rodrigues_to_matrix(r)) - r).max() < 1e-12
