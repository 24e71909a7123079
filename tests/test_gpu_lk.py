"""GPU parity: pyramidal LK through the C ABI vs the oracle and the cv2 golden.  Tolerance 0.01 px (north_star)."""
import numpy as np
import pytest

from conftest import load_golden, sha
from oracle import lk_oracle as lo
from oracle import synth

pytestmark = pytest.mark.gpu
POS_TOL = 0.01      # px, BASELINE.json north_star
ERR_TOL = 0.02      # mean abs patch difference (grey levels)


@pytest.fixture(scope="module")
def ctx():
    from ros2_mono_vo_b200 import Context
    c = Context(1920, 1080, nfeatures=5000)
    yield c
    c.close()


@pytest.mark.parametrize("tag", ["small", "c2", "c3"])
def test_lk_vs_cv2_golden(ctx, tag):
    g = load_golden("lk_c3.npz" if tag == "c3" else "lk.npz")
    h, w, seed = g[f"{tag}_hw_seed"].tolist()
    f0, f1 = synth.synth_pair(h, w, seed)
    assert sha(f0) == str(g[f"{tag}_sha0"])
    nxt, st, err = ctx.lk_track(f0, f1, g[f"{tag}_pts"])
    assert np.array_equal(st, g[f"{tag}_status"])
    m = st == 1
    assert np.abs(nxt[m] - g[f"{tag}_next"][m]).max() < POS_TOL
    assert np.abs(err[m] - g[f"{tag}_err"][m]).max() < ERR_TOL
    assert np.abs(nxt[~m] - g[f"{tag}_next"][~m]).max() < POS_TOL
    # reference filter (src/tracker.cpp:70-77): status && err < 30 keeps the same set
    assert np.array_equal(m & (err < 30.0), (g[f"{tag}_status"] == 1) & (g[f"{tag}_err"] < 30.0))


@pytest.mark.parametrize("h,w,seed,n", [(480, 640, 41, 1000), (1080, 1920, 43, 5000), (100, 47, 44, 50)])
def test_lk_vs_oracle(ctx, h, w, seed, n):
    f0, f1 = synth.synth_pair(h, w, seed)
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-25, w + 25, n), rng.uniform(-25, h + 25, n)], 1).astype(np.float32)
    nxt, st, err = ctx.lk_track(f0, f1, pts)
    onxt, ost, oerr = lo.lk_track(f0, f1, pts)
    assert np.array_equal(st, ost)                 # exact: no point of these sets sits on a threshold (scripts/diag_parity_soft.py)
    m = st == 1
    assert np.abs(nxt[m] - onxt[m]).max() < POS_TOL
    assert np.abs(err[m] - oerr[m]).max() < ERR_TOL


def test_lk_flat_and_empty(ctx):
    g = load_golden("lk.npz")
    f0 = synth.synth_frame(120, 160, 4)
    nxt, st, err = ctx.lk_track(np.full_like(f0, 100), f0, g["flat_pts"])
    assert np.array_equal(st, g["flat_status"]) and st.sum() == 0
    assert np.allclose(nxt, g["flat_next"], atol=1e-3)
    nxt, st, err = ctx.lk_track(f0, f0, np.zeros((0, 2), np.float32))
    assert nxt.shape == (0, 2)


def test_lk_identity_property(ctx):
    """Size-independent property at full C3 size: tracking an image onto itself returns the input points."""
    f0 = synth.synth_frame(1080, 1920, 7)
    rng = np.random.default_rng(7)
    pts = np.stack([rng.uniform(30, 1890, 5000), rng.uniform(30, 1050, 5000)], 1).astype(np.float32)
    nxt, st, err = ctx.lk_track(f0, f0, pts)
    ok = st == 1
    assert ok.mean() > 0.99
    assert np.abs(nxt[ok] - pts[ok]).max() < 1e-3
    assert err[ok].max() < 1e-3


@pytest.mark.parametrize("tag", ["small", "c2"])
def test_lk_bgr_vs_cv2_golden(ctx, tag):
    """BGR8 input (cn = 3), as the reference node feeds it (src/mono_vo.cpp:94 -> src/tracker.cpp:68)."""
    g = load_golden("lk_bgr.npz")
    h, w, seed = g[f"{tag}_hw_seed"].tolist()
    f0, f1 = synth.synth_pair_bgr(h, w, seed)
    assert sha(f0) == str(g[f"{tag}_sha0"])
    nxt, st, err = ctx.lk_track(f0, f1, g[f"{tag}_pts"])
    assert np.array_equal(st, g[f"{tag}_status"])
    m = st == 1
    assert np.abs(nxt[m] - g[f"{tag}_next"][m]).max() < POS_TOL
    assert np.abs(err[m] - g[f"{tag}_err"][m]).max() < ERR_TOL
    assert np.abs(nxt[~m] - g[f"{tag}_next"][~m]).max() < POS_TOL
    assert np.array_equal(m & (err < 30.0), (g[f"{tag}_status"] == 1) & (g[f"{tag}_err"] < 30.0))
    # a gray frame replicated to three channels goes through the same kernel
    g0, g1 = synth.synth_pair(h, w, seed)
    pts = g[f"{tag}_pts"]
    n3, s3, e3 = ctx.lk_track(np.repeat(g0[:, :, None], 3, 2), np.repeat(g1[:, :, None], 3, 2), pts)
    o3, os3, oe3 = lo.lk_track(np.repeat(g0[:, :, None], 3, 2), np.repeat(g1[:, :, None], 3, 2), pts)
    assert np.array_equal(s3, os3)
    mm = s3 == 1
    assert np.abs(n3[mm] - o3[mm]).max() < POS_TOL


@pytest.mark.parametrize("h,w,seed", [(240, 320, 12), (480, 640, 5), (121, 203, 9)])
def test_lk_bgr_team_kernel_matches_first_generation(ctx, h, w, seed):
    """The three-warp-team BGR8 kernel against the one-warp-per-point kernel (lk_impl = 1): identical outputs, on a point
    set that is dense along the image borders (border template / border J region on one to four levels) and includes
    points whose windows leave the image."""
    f0, f1 = synth.synth_pair_bgr(h, w, seed)
    rng = np.random.default_rng(seed)
    inner = rng.uniform([0, 0], [w - 1, h - 1], (1500, 2))
    edge = np.concatenate([np.stack([rng.uniform(-3, 45, 400), rng.uniform(0, h, 400)], 1),
                           np.stack([rng.uniform(w - 45, w + 3, 400), rng.uniform(0, h, 400)], 1),
                           np.stack([rng.uniform(0, w, 400), rng.uniform(-3, 45, 400)], 1),
                           np.stack([rng.uniform(0, w, 400), rng.uniform(h - 45, h + 3, 400)], 1)])
    grid = np.stack(np.meshgrid(np.arange(20, w - 19, 40.0), np.arange(20, h - 19, 40.0)), -1).reshape(-1, 2)
    pts = np.concatenate([inner, edge, grid]).astype(np.float32)
    for frames in ((f0, f1), (np.repeat(f0[:, :, :1], 3, 2), np.repeat(f1[:, :, :1], 3, 2))):
        ctx.debug_set("cache", 0)
        try:
            new = ctx.lk_track(*frames, pts)
            ctx.debug_set("lk_impl", 1)
            old = ctx.lk_track(*frames, pts)
        finally:
            ctx.debug_set("lk_impl", 0)
            ctx.debug_set("cache", 1)
        assert np.array_equal(new[1], old[1])
        assert np.array_equal(new[0], old[0], equal_nan=True)
        assert np.array_equal(new[2], old[2])


def test_lk_bgr_gray_streams_take_the_gray_path_with_identical_results():
    """A BGR8 group in which some streams are gray cameras behind a BGR8 conversion (identical planes, what cv_bridge
    makes of a mono8 KITTI frame: /root/reference/src/mono_vo.cpp:94) and some have real colour: the identical-plane
    streams are tracked on one plane with tripled sums, the others by the three-warp teams, and every output equals the
    all-teams run (lk_bgr_gray = 0) and the one-warp-per-point kernel (lk_impl = 1).  Frame pairs where only one of the
    two frames is gray must not take the gray path.  Persistent kernels (34 streams x 400 points)."""
    from ros2_mono_vo_b200 import Context, _lib
    h, w, B, n = 240, 320, 34, 400
    frames = []
    for t in range(3):
        fs = []
        for s in range(B):
            g0, g1 = synth.synth_pair(h, w, 100 + s)
            c0, c1 = synth.synth_pair_bgr(h, w, 100 + s)
            gray3 = [np.repeat(g[:, :, None], 3, 2) for g in (g0, g1, g0)]
            col = [c0, c1, c0]
            if s % 3 == 0:
                f = gray3[t]                      # gray camera throughout
            elif s % 3 == 1:
                f = col[t]                        # colour throughout
            else:
                f = gray3[t] if t != 1 else col[t]   # gray, colour, gray: both pairs are mixed
            fs.append(f)
        frames.append(np.stack(fs))
    K = synth.sequence_camera(h, w)
    outs = {}
    for mode in ("default", "teams", "first"):
        c = Context(w, h, nfeatures=n, batch=B)
        c.group_configure(channels=3, outputs=_lib.MVO_OUT_TRACKS)
        if mode == "teams":
            c.debug_set("lk_bgr_gray", 0)
        if mode == "first":
            c.debug_set("lk_impl", 1)
        got = []
        for t in range(3):
            res = c.group_step(frames[t], K)
            if t:
                got.append([{k: v.copy() for k, v in c.group_outputs(s).items() if k.startswith("track") or k.startswith("lk")} for s in range(B)])
                assert (res["n_tracked"] > 50).all()
        outs[mode] = got
        c.close()
    keys = outs["default"][0][0].keys()
    assert len(keys) >= 2
    for mode in ("teams", "first"):
        for t in range(2):
            for s in range(B):
                for k in keys:
                    assert np.array_equal(outs["default"][t][s][k], outs[mode][t][s][k], equal_nan=True), (mode, t, s, k)


def test_lk_outside_and_nan_points(ctx):
    """Points outside the image and a NaN coordinate (values from cv2 4.13.0 on the same pair: status 0 1 0 0 1 1; OpenCV
    floors a NaN to INT_MIN, i.e. out of range on every level, and still returns the propagated coordinates)."""
    f0, f1 = synth.synth_pair(480, 640, 2)
    pts = np.array([[np.nan, 5], [100, 100], [-50, -50], [700, 500], [639.9, 479.9], [0, 0]], np.float32)
    nxt, st, err = ctx.lk_track(f0, f1, pts)
    onxt, ost, oerr = lo.lk_track(f0, f1, pts)
    assert st.tolist() == [0, 1, 0, 0, 1, 1] and np.array_equal(st, ost)
    assert np.isnan(nxt[0, 0]) and nxt[0, 1] == 5.0 and err[0] == 0.0
    assert np.abs(nxt[1:] - onxt[1:]).max() < POS_TOL
    assert np.abs(nxt[1:] - np.array([[101.367386, 99.107544], [-48.246277, -51.015106], [700.7065, 500.23755],
                                      [640.7122, 480.0799], [1.4291534, -1.006732]])).max() < POS_TOL


@pytest.mark.parametrize("h,w,bgr", [(376, 1241, False), (480, 640, False), (1080, 1920, False), (97, 131, False), (200, 333, True)])
def test_lk_pyramid_levels_bit_exact(h, w, bgr):
    """Every pyramid level equals cv::pyrDown's chain (oracle lo.pyr_down, pinned to cv2.pyrDown in tests/test_oracle_lk.py):
    image sizes that end inside a tile, widths that are / are not multiples of the tile, odd sizes, BGR planes."""
    from ros2_mono_vo_b200 import Context
    f0, f1 = (synth.synth_pair_bgr if bgr else synth.synth_pair)(h, w, 7)
    c = Context(w, h, nfeatures=100)
    try:
        c.lk_track(f0, f1, np.array([[w / 2, h / 2]], np.float32))
        for which, img in ((0, f0), (1, f1)):
            for plane in range(3 if bgr else 1):
                ref = np.ascontiguousarray(img[:, :, plane]) if bgr else img
                level = 0
                while True:
                    got = c.lk_level(which, level, plane)
                    if got is None:
                        break
                    assert got.shape == ref.shape and np.array_equal(got, ref), (which, plane, level)
                    ref = lo.pyr_down(ref)
                    level += 1
                assert level >= 3
    finally:
        c.close()
