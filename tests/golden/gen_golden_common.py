"""Helpers shared by gen_golden.py (needs cv2) and the tests (numpy only)."""
import numpy as np


def distort_pixels(img, K, d):
    """Forward plumb-bob model (k1 k2 p1 p2 k3) applied to ideal pixel positions."""
    k1, k2, p1, p2, k3 = d
    x = (img[:, 0].astype(np.float64) - K[0, 2]) / K[0, 0]
    y = (img[:, 1].astype(np.float64) - K[1, 2]) / K[1, 1]
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([xd * K[0, 0] + K[0, 2], yd * K[1, 1] + K[1, 2]], 1).astype(np.float32)
