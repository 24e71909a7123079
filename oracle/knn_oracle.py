"""CPU oracle (numpy restatement) of cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) + Lowe ratio filter.

TEST INFRASTRUCTURE ONLY.
Reference: /root/reference/src/feature_processor.cpp:25-40 (find_matches): knnMatch k=2, keep m[0]
iff m.size()==2 && m[0].distance < ratio*m[1].distance (float < double*float, evaluated in double).
OpenCV semantics (pinned by tests/golden/knn_*.npz from cv2 4.13.0): top-2 by
(distance, trainIdx) ascending == stable argsort; rows have min(2, Nt) entries.
"""
from __future__ import annotations

import numpy as np

_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.int32)


def hamming_matrix(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    out = np.zeros((len(q), len(t)), np.int32)
    for b in range(q.shape[1]):
        out += _POP8[q[:, b][:, None] ^ t[:, b][None, :]]
    return out


def knn2(q: np.ndarray, t: np.ndarray):
    """Returns idx (Nq x 2 int32, -1 where absent) and dist (Nq x 2 int32, -1 where absent)."""
    nq, nt = len(q), len(t)
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), -1, np.int32)
    if nq == 0 or nt == 0:
        return idx, dist
    d = hamming_matrix(q, t)
    order = np.argsort(d, axis=1, kind="stable")[:, :2]
    k = order.shape[1]
    idx[:, :k] = order
    dist[:, :k] = np.take_along_axis(d, order, axis=1)
    return idx, dist


def find_matches(q: np.ndarray, t: np.ndarray, ratio: float):
    """(queryIdx, trainIdx, distance f32) of accepted matches in query order."""
    idx, dist = knn2(q, t)
    ok = (idx[:, 1] >= 0) & (dist[:, 0].astype(np.float32).astype(np.float64)
                            < np.float64(ratio) * dist[:, 1].astype(np.float32).astype(np.float64))
    qi = np.nonzero(ok)[0].astype(np.int32)
    return qi, idx[qi, 0], dist[qi, 0].astype(np.float32)
