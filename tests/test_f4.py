"""keypoint-distribution grid as a by-product of ORB, PointCloud2 packing on the device."""
import numpy as np
import pytest

from oracle import f4_oracle as fo
from oracle import synth


def test_oracle_occupancy_literal_loop():
    """The numpy restatement equals the reference's loop written out literally (src/initializer.cpp:57-68)."""
    rng = np.random.default_rng(0)
    rows, cols, div = 376, 1241, 50
    kx = rng.uniform(31, cols - 31, 500).astype(np.float32)
    ky = rng.uniform(31, rows - 31, 500).astype(np.float32)
    gr, gc = rows // div, cols // div
    grid = np.zeros(gr * gc + 64, np.uint8)      # flat CV_8U storage (+ slack where the reference would overrun)
    occupied = 0
    for x, y in zip(kx, ky):
        r, c = int(np.float32(y) / np.float32(div)), int(np.float32(x) / np.float32(div))
        if r * gc + c < gr * gc and not grid[r * gc + c]:
            grid[r * gc + c] = 1
            occupied += 1
    assert fo.keypoint_occupancy(kx, ky, rows, cols, div) == (occupied, gr * gc)
    assert fo.good_keypoint_distribution(kx, ky, rows, cols) == (occupied / (gr * gc) > 0.5)


def test_oracle_pack_pointcloud():
    p = np.array([[1.0, 2.0, 3.0], [-4.5, 0.25, 7.0]], np.float32)
    d = fo.pack_pointcloud(p)
    assert d.dtype == np.uint8 and len(d) == 24
    assert np.array_equal(d.view("<f4"), np.array([3.0, -1.0, -2.0, 7.0, 4.5, -0.25], np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,n,div", [(376, 1241, 2000, 50), (480, 640, 1000, 50), (240, 320, 300, 16), (1080, 1920, 5000, 50)])
def test_gpu_occupancy_by_product(h, w, n, div):
    from ros2_mono_vo_b200 import Context
    img = synth.synth_frame(h, w, 5)
    ctx = Context(w, h, nfeatures=n)
    ctx.set_occupancy_grid(div)
    kps, _ = ctx.orb_detect_and_compute(img)
    assert ctx.orb_occupancy() == fo.keypoint_occupancy(kps["x"], kps["y"], h, w, div)
    # a mostly empty image: few occupied cells; and the switch
    img2 = np.full((h, w), 90, np.uint8)
    img2[40:90, 40:90] = synth.synth_frame(50, 50, 1)
    kps2, _ = ctx.orb_detect_and_compute(img2)
    occ2 = ctx.orb_occupancy()
    assert occ2 == fo.keypoint_occupancy(kps2["x"], kps2["y"], h, w, div) and occ2[0] <= (50 // div + 2) ** 2
    ctx.set_occupancy_grid(0)
    ctx.orb_detect_and_compute(img)
    from ros2_mono_vo_b200.api import MvoError
    with pytest.raises(MvoError):
        ctx.orb_occupancy()
    ctx.close()


@pytest.mark.gpu
def test_gpu_pack_pointcloud_and_group_cloud():
    from ros2_mono_vo_b200 import Context, _lib
    rng = np.random.default_rng(1)
    ctx = Context(320, 240, nfeatures=300)
    for n in (0, 1, 7, 5000):
        pts = rng.normal(0, 10, (n, 3)).astype(np.float32)
        assert np.array_equal(ctx.pack_pointcloud(pts), fo.pack_pointcloud(pts))
    ctx.close()
    # group step: MVO_OUT_CLOUD == packing of the chirality-valid triangulated points, occupancy per stream
    h, w, n, batch = 240, 320, 300, 3
    seqs = [synth.synth_sequence(h, w, s, 3) for s in range(batch)]
    K = seqs[0][1]
    grp = Context(w, h, nfeatures=n, batch=batch)
    grp.group_configure(channels=1, outputs=_lib.MVO_OUT_ALL)
    for t in range(3):
        res = grp.group_step(np.stack([seqs[s][0][t] for s in range(batch)]), K)
        for s in range(batch):
            o = grp.group_outputs(s)
            assert (o["occupied_cells"], o["total_cells"]) == fo.keypoint_occupancy(o["keypoints"]["x"], o["keypoints"]["y"], h, w, 50)
            assert grp.orb_occupancy(s) == (o["occupied_cells"], o["total_cells"])
            if t == 0:
                continue
            X = o["X4"].astype(np.float32)
            sc = np.where(X[3] != 0, np.float32(1) / X[3], np.float32(1)).astype(np.float32)
            xyz = (X[:3] * sc).T                                   # convertPointsFromHomogeneous
            R, tt = res[s]["R"].reshape(3, 3), res[s]["t"]
            z2 = R[2, 0] * xyz[:, 0].astype(np.float64) + R[2, 1] * xyz[:, 1].astype(np.float64) + R[2, 2] * xyz[:, 2].astype(np.float64) + tt[2]
            ok = (o["mask_pose"] != 0) & (xyz[:, 2] > 0) & (z2 > 0)
            assert int(ok.sum()) == int(res[s]["n_triangulated"]) == len(o["cloud_xyz"])
            assert np.array_equal(o["cloud_xyz"].reshape(-1).view(np.uint8), fo.pack_pointcloud(xyz[ok]))
    grp.close()
