"""GPU parity: Hamming kNN(k=2) + ratio through the C ABI vs the oracle / cv2 golden.  Bit-exact indices."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import knn_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["tcgen05", "popc"])
def ctx(request):
    """Every test runs on both matching kernels: the tensor-core one (knn_mma_kernel, the default) and the xor + popc
    one (knn_top2_kernel, the in-tree cross-check)."""
    from ros2_mono_vo_b200 import Context
    c = Context(640, 480, nfeatures=1000)
    c.debug_set("knn_impl", 1 if request.param == "popc" else 0)
    yield c
    c.close()


def test_knn_golden(ctx):
    g = load_golden("knn.npz")
    idx, dist = ctx.knn2(g["q"], g["t"])
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(dist.astype(np.float32), g["dist"])
    idx, dist = ctx.knn2(g["d0"], g["d1"])
    assert np.array_equal(idx, g["idx2"]) and np.array_equal(dist.astype(np.float32), g["dist2"])
    m = ctx.knn_ratio(g["d0"], g["d1"], 0.7)
    good = g["good"]
    assert np.array_equal(m["query_idx"], good[:, 0].astype(np.int32))
    assert np.array_equal(m["train_idx"], good[:, 1].astype(np.int32))
    assert np.array_equal(m["distance"], good[:, 2])
    assert (m["img_idx"] == 0).all()


def test_knn_c3_vs_cv2_golden(ctx):
    """configs[2] as worded: 5000 x 5000 keyframe matching with Lowe ratio 0.7, against cv2's own knnMatch output."""
    g = load_golden("knn_c3.npz")
    idx, dist = ctx.knn2(g["d0"], g["d1"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist.astype(np.float32), g["dist"])
    m = ctx.knn_ratio(g["d0"], g["d1"], 0.7)
    good = g["good"]
    assert np.array_equal(m["query_idx"], good[:, 0].astype(np.int32))
    assert np.array_equal(m["train_idx"], good[:, 1].astype(np.int32))
    assert np.array_equal(m["distance"], good[:, 2])


@pytest.mark.parametrize("nq,nt,bits", [(1000, 1000, 256), (2000, 2000, 256), (5000, 5000, 256), (777, 1313, 4),
                                        (33, 257, 2), (1, 2, 256), (3000, 255, 256)])
def test_knn_vs_oracle(ctx, nq, nt, bits):
    rng = np.random.default_rng(nq * 7 + nt)
    q = rng.integers(0, bits, (nq, 32)).astype(np.uint8)
    t = rng.integers(0, bits, (nt, 32)).astype(np.uint8)
    # plant near-duplicates so that the ratio test accepts a good fraction
    k = min(nq, nt) // 2
    t[:k] = q[:k] ^ (rng.integers(0, 256, (k, 32)) < 8).astype(np.uint8)
    idx, dist = ctx.knn2(q, t)
    oidx, odist = ko.knn2(q, t)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    for ratio in (0.7, 0.85, 1.0):
        m = ctx.knn_ratio(q, t, ratio)
        qi, ti, d = ko.find_matches(q, t, ratio)
        assert np.array_equal(m["query_idx"], qi) and np.array_equal(m["train_idx"], ti)
        assert np.array_equal(m["distance"], d)


def test_knn_edge_cases(ctx):
    rng = np.random.default_rng(0)
    q = rng.integers(0, 256, (10, 32)).astype(np.uint8)
    assert len(ctx.knn_ratio(q, q[:0], 0.7)) == 0          # empty train set
    assert len(ctx.knn_ratio(q, q[:1], 0.7)) == 0          # one train row: match.size() != 2
    assert len(ctx.knn_ratio(q[:0], q, 0.7)) == 0          # empty query set
    idx, dist = ctx.knn2(q, q[:1])
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all()
    # identical train rows: ties resolve to the lowest train index
    t = np.repeat(q[:1], 5, axis=0)
    idx, dist = ctx.knn2(q[:1], t)
    assert idx.tolist() == [[0, 1]] and dist.tolist() == [[0, 0]]


def test_knn_linearity_property(ctx):
    """Size-independent property at the full C3 size: xor-ing every descriptor with one mask leaves all
    Hamming distances, hence the whole result, unchanged; permuting train rows permutes indices."""
    rng = np.random.default_rng(9)
    q = rng.integers(0, 256, (5000, 32)).astype(np.uint8)
    t = rng.integers(0, 256, (5000, 32)).astype(np.uint8)
    mask = rng.integers(0, 256, (1, 32)).astype(np.uint8)
    i0, d0 = ctx.knn2(q, t)
    i1, d1 = ctx.knn2(q ^ mask, t ^ mask)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    perm = rng.permutation(len(t))
    i2, d2 = ctx.knn2(q, t[perm])
    assert np.array_equal(d0, d2)
    untied = d0[:, 0] != d0[:, 1]
    assert np.array_equal(perm[i2[untied, 0]], i0[untied, 0])


def test_knn_extreme_distances_and_cross_tile_ties(ctx):
    """Distances 0 and 256 (the ends of the tensor kernel's 16-bit key range), and exact ties between train rows that
    sit in different 128-row tiles and in both halves of a split train range: lowest index wins."""
    rng = np.random.default_rng(21)
    q = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    t = rng.integers(0, 256, (1500, 32)).astype(np.uint8)
    t[5] = q[0]; t[700] = q[0]; t[1400] = q[0]              # three exact copies of query 0 in three tiles
    t[130] = ~q[1]                                          # distance 256 from query 1
    t[900] = q[2]; t[901] = q[2]                            # adjacent columns (the two packed 16-bit lanes)
    idx, dist = ctx.knn2(q, t)
    oidx, odist = ko.knn2(q, t)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    assert idx[0].tolist() == [5, 700] and dist[0].tolist() == [0, 0]
    assert idx[2].tolist() == [900, 901]
    # a train set of complements only: every distance is large, one is exactly 256
    tc = ~q[:200]
    idx, dist = ctx.knn2(q[:200], tc)
    oidx, odist = ko.knn2(q[:200], tc)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    # all-identical train rows across many tiles
    tt = np.repeat(q[:1], 1000, axis=0)
    idx, dist = ctx.knn2(q[:3], tt)
    assert idx[0].tolist() == [0, 1] and dist[0].tolist() == [0, 0]
    oidx, odist = ko.knn2(q[:3], tt)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
