"""Pin the kNN oracle against cv2 BFMatcher golden vectors (ties -> lowest train index)."""
import numpy as np

from conftest import load_golden
from oracle import knn_oracle as ko


def test_knn2_low_entropy_ties():
    g = load_golden("knn.npz")
    idx, dist = ko.knn2(g["q"], g["t"])
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(dist.astype(np.float32), g["dist"])
    assert (dist[:, 0] == dist[:, 1]).sum() > 20          # the fixture really exercises ties


def test_knn2_orb_descriptors_and_ratio():
    g = load_golden("knn.npz")
    idx, dist = ko.knn2(g["d0"], g["d1"])
    assert np.array_equal(idx, g["idx2"])
    assert np.array_equal(dist.astype(np.float32), g["dist2"])
    qi, ti, d = ko.find_matches(g["d0"], g["d1"], 0.7)
    good = g["good"]
    assert np.array_equal(qi, good[:, 0].astype(np.int32))
    assert np.array_equal(ti, good[:, 1].astype(np.int32))
    assert np.array_equal(d, good[:, 2])


def test_knn_c3_golden():
    """configs[2]: 5000 x 5000 keyframe matching against cv2's knnMatch."""
    g = load_golden("knn_c3.npz")
    idx, dist = ko.knn2(g["d0"], g["d1"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist.astype(np.float32), g["dist"])
    qi, ti, d = ko.find_matches(g["d0"], g["d1"], 0.7)
    good = g["good"]
    assert np.array_equal(qi, good[:, 0].astype(np.int32)) and np.array_equal(ti, good[:, 1].astype(np.int32))


def test_knn_edge_cases():
    rng = np.random.default_rng(1)
    q = rng.integers(0, 256, (5, 32)).astype(np.uint8)
    idx, dist = ko.knn2(q, q[:0])
    assert (idx == -1).all()
    idx, dist = ko.knn2(q, q[:1])
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all()
    assert len(ko.find_matches(q, q[:1], 0.7)[0]) == 0      # match.size() != 2 -> dropped
    assert len(ko.find_matches(q[:0], q, 0.7)[0]) == 0
    # ratio 0.7 is the integer test 10*d0 < 7*d1 for all distances 0..256
    d0, d1 = np.meshgrid(np.arange(257), np.arange(257), indexing="ij")
    lhs = d0.astype(np.float32).astype(np.float64) < 0.7 * d1.astype(np.float32).astype(np.float64)
    assert np.array_equal(lhs, 10 * d0 < 7 * d1)
