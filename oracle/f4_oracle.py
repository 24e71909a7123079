"""Oracle (test infrastructure): the two small host loops of SURVEY.md 8(f) #4, restated in numpy.

* keypoint_occupancy  -- Initializer::good_keypoint_distribution, /root/reference/src/initializer.cpp:52-75
* pack_pointcloud     -- points3d_to_pointcloud_msg,              /root/reference/src/utils.cpp:184-243
Parity: the reference has no test for either; the loops are 10 lines of integer / copy arithmetic restated literally
(parity unpinned by the reference; pinned by construction to its source lines).
"""
import numpy as np


def keypoint_occupancy(kx, ky, rows: int, cols: int, div: int):
    """(occupied cells, total cells).  The reference builds a (rows / div) x (cols / div) CV_8U grid and marks
    grid.at<uchar>(int(pt.y / div), int(pt.x / div)) (initializer.cpp:57-66); for a continuous Mat that is the byte at
    r * grid_cols + c.  Indices past the end of the grid (undefined behaviour in the reference) are ignored."""
    gr, gc = rows // div, cols // div
    r = (np.asarray(ky, np.float32) / np.float32(div)).astype(np.int64)
    c = (np.asarray(kx, np.float32) / np.float32(div)).astype(np.int64)
    cell = r * gc + c
    cell = cell[(cell >= 0) & (cell < gr * gc)]
    return int(len(np.unique(cell))), int(gr * gc)


def good_keypoint_distribution(kx, ky, rows, cols, div=50, thresh=0.5) -> bool:
    occ, tot = keypoint_occupancy(kx, ky, rows, cols, div)
    return occ / tot > thresh            # initializer.cpp:69-74


def pack_pointcloud(points_xyz) -> np.ndarray:
    """PointCloud2.data of points3d_to_pointcloud_msg: per point float32 (z, -x, -y), 12 bytes, little endian."""
    p = np.asarray(points_xyz, np.float32).reshape(-1, 3)
    out = np.stack([p[:, 2], -p[:, 0], -p[:, 1]], 1).astype("<f4")
    return out.reshape(-1).view(np.uint8)
