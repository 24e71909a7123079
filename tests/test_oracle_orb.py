"""Pin the ORB oracle (oracle/orb_oracle.py) against golden vectors produced by cv2 4.13.0."""
import numpy as np
import pytest

from conftest import golden_image, kp_dict, load_golden, sha
from oracle import orb_oracle as oo


@pytest.mark.parametrize("name", ["orb_small.npz", "orb_c1.npz", "orb_c2.npz", "orb_c3.npz"])
def test_oracle_orb_matches_cv2_golden(name):
    g = load_golden(name)
    img = golden_image(g)
    kps, desc = oo.orb_detect_and_compute(img, int(g["nfeatures"]))
    gd = kp_dict(g["octave"], g["x"], g["y"])
    od = kp_dict(kps["octave"], kps["x"], kps["y"])
    assert set(gd) == set(od)                     # keypoint SET is bit-exact
    gi = np.array([gd[k] for k in od])            # golden index for each oracle keypoint
    oi = np.array([od[k] for k in od])
    for f in ("response", "angle", "size"):
        assert np.array_equal(g[f][gi].view(np.uint32), kps[f][oi].view(np.uint32)), f
    assert np.array_equal(g["desc"][gi], desc[oi])  # descriptors byte-exact


def test_oracle_pyramid_fast_blur_small():
    g = load_golden("orb_small.npz")
    img = golden_image(g)
    pyr = oo.build_pyramid(img)
    for l, lvl in enumerate(pyr):
        assert sha(lvl) == str(g[f"level{l}_sha"]), f"pyramid level {l}"
        assert sha(oo.blur_orb(lvl)) == str(g[f"blur{l}_sha"]), f"blur level {l}"
        xs, ys, sc = oo.fast_nms(oo.fast_score_map(lvl))
        ref = g[f"fast{l}"]
        assert set(zip(xs.tolist(), ys.tolist(), sc.tolist())) == set(map(tuple, ref.tolist())), f"FAST level {l}"


def test_quotas_and_sizes():
    assert oo.level_quotas(1000) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert oo.level_quotas(2000) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oo.level_quotas(5000) == [1086, 905, 754, 628, 524, 436, 364, 303]
    assert oo.level_sizes(1241, 376) == [(1241, 376), (1034, 313), (862, 261), (718, 218), (598, 181), (499, 151),
                                         (416, 126), (346, 105)]


def test_bgr_equals_gray_conversion():
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (64, 80, 3)).astype(np.uint8)
    g = oo.bgr_to_gray(bgr)
    assert g.shape == (64, 80) and g.dtype == np.uint8
    # luma weights sum to 2^15
    assert 3735 + 19235 + 9798 == 32768
    grey3 = np.repeat(rng.integers(0, 256, (8, 8, 1)).astype(np.uint8), 3, axis=2)
    assert np.array_equal(oo.bgr_to_gray(grey3), grey3[..., 0])
