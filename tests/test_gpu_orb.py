"""GPU parity: CUDA ORB (through the C ABI) vs the oracle and the cv2 golden vectors.  Bit-exact."""
import numpy as np
import pytest

from conftest import golden_image, kp_dict, load_golden, sha
from oracle import orb_oracle as oo
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx_cache():
    from ros2_mono_vo_b200 import Context
    cache = {}

    def get(w, h, n):
        key = (w, h, n)
        if key not in cache:
            cache[key] = Context(w, h, nfeatures=n)
        return cache[key]
    yield get
    for c in cache.values():
        c.close()


def _compare(kps, desc, ref_oct, ref_x, ref_y, ref_fields, ref_desc):
    gd = kp_dict(ref_oct, ref_x, ref_y)
    od = kp_dict(kps["octave"], kps["x"], kps["y"])
    assert len(od) == len(kps), "duplicate keypoints"
    assert set(gd) == set(od), f"keypoint set differs: {len(set(gd) ^ set(od))} of {len(gd)}"
    gi = np.array([gd[k] for k in od])
    oi = np.array([od[k] for k in od])
    for f, ref in ref_fields.items():
        assert np.array_equal(np.asarray(ref)[gi].view(np.uint32), kps[f][oi].view(np.uint32)), f
    if desc is not None:
        assert np.array_equal(np.asarray(ref_desc)[gi], desc[oi]), "descriptors differ"


@pytest.mark.parametrize("name", ["orb_small.npz", "orb_c1.npz", "orb_c2.npz", "orb_c3.npz"])
def test_orb_vs_cv2_golden(ctx_cache, name):
    g = load_golden(name)
    img = golden_image(g)
    ctx = ctx_cache(int(g["w"]), int(g["h"]), int(g["nfeatures"]))
    kps, desc = ctx.orb_detect_and_compute(img)
    _compare(kps, desc, g["octave"], g["x"], g["y"],
             {"response": g["response"], "angle": g["angle"], "size": g["size"]}, g["desc"])
    assert (kps["class_id"] == -1).all()
    # canonical order: octave, response desc, y, x
    o = oo.canonical_order(kps["octave"], kps["response"], kps["y"], kps["x"])
    assert np.array_equal(o, np.arange(len(kps)))
    # deterministic run to run
    kps2, desc2 = ctx.orb_detect_and_compute(img)
    assert np.array_equal(kps, kps2) and np.array_equal(desc, desc2)


def test_pyramid_fast_blur_taps(ctx_cache):
    g = load_golden("orb_small.npz")
    img = golden_image(g)
    ctx = ctx_cache(int(g["w"]), int(g["h"]), int(g["nfeatures"]))
    ctx.orb_detect_and_compute(img)
    for l in range(8):
        lvl = ctx.orb_level(l, blurred=False)
        assert sha(lvl) == str(g[f"level{l}_sha"]), f"INTER_LINEAR_EXACT level {l}"
        assert sha(ctx.orb_level(l, blurred=True)) == str(g[f"blur{l}_sha"]), f"blur level {l}"
        xs, ys, sc = ctx.orb_fast(l)
        ref = g[f"fast{l}"]
        h, w = lvl.shape
        m = (ref[:, 0] >= 31) & (ref[:, 0] < w - 31) & (ref[:, 1] >= 31) & (ref[:, 1] < h - 31)
        assert set(zip(xs.tolist(), ys.tolist(), sc.tolist())) == set(map(tuple, ref[m].tolist())), f"FAST level {l}"


@pytest.mark.parametrize("h,w,n,seed", [(480, 640, 1000, 21), (376, 1241, 2000, 22), (1080, 1920, 5000, 23),
                                        (200, 333, 500, 24)])
def test_orb_vs_oracle_seeded(ctx_cache, h, w, n, seed):
    img = synth.synth_frame(h, w, seed)
    okp, odesc = oo.orb_detect_and_compute(img, n)
    ctx = ctx_cache(w, h, n)
    kps, desc = ctx.orb_detect_and_compute(img)
    _compare(kps, desc, okp["octave"], okp["x"], okp["y"],
             {"response": okp["response"], "angle": okp["angle"], "size": okp["size"]}, odesc)
    assert np.array_equal(kps, okp)          # same canonical order, every field bit-equal
    assert np.array_equal(desc, odesc)


def test_orb_noise_frame_grows_the_candidate_lists():
    """A frame of pure noise has ~10 % FAST corners after NMS, more than the default candidate capacity (1 / 16 of the
    pixels): the detect call doubles the lists and runs again instead of failing; the result is still cv::ORB's."""
    from ros2_mono_vo_b200 import Context
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (240, 320)).astype(np.uint8)
    ctx = Context(320, 240, nfeatures=500)
    kps, desc = ctx.orb_detect_and_compute(img)
    okp, odesc = oo.orb_detect_and_compute(img, 500)
    assert len(kps) == len(okp) > 300
    assert np.array_equal(kps, okp) and np.array_equal(desc, odesc)
    # and the context keeps working on ordinary frames afterwards
    img2 = synth.synth_frame(240, 320, 5)
    k2, d2 = ctx.orb_detect_and_compute(img2)
    o2, od2 = oo.orb_detect_and_compute(img2, 500)
    assert np.array_equal(k2, o2) and np.array_equal(d2, od2)
    ctx.close()


def test_orb_bgr_input(ctx_cache):
    img = synth.synth_frame(240, 320, 31)
    rng = np.random.default_rng(3)
    bgr = np.stack([np.clip(img.astype(int) + rng.integers(-20, 20, img.shape), 0, 255).astype(np.uint8)
                    for _ in range(3)], axis=2)
    ctx = ctx_cache(320, 240, 300)
    kps, desc = ctx.orb_detect_and_compute(bgr)
    okp, odesc = oo.orb_detect_and_compute(bgr, 300)
    assert np.array_equal(kps, okp) and np.array_equal(desc, odesc)


def test_orb_compute_given_keypoints(ctx_cache):
    """Descriptors given identical keypoints/angles are bit-exact (the north_star wording)."""
    g = load_golden("orb_c1.npz")
    img = golden_image(g)
    ctx = ctx_cache(640, 480, 1000)
    from ros2_mono_vo_b200.api import KP_DTYPE
    kin = np.zeros(len(g["x"]), KP_DTYPE)
    for f in ("x", "y", "size", "angle", "response", "octave"):
        kin[f] = g[f]
    kin["class_id"] = -1
    desc, valid = ctx.orb_compute(img, kin)
    assert valid.all()
    assert np.array_equal(desc, g["desc"])
    # keypoints near the border are flagged invalid, like cv2's compute() drops them
    kin2 = kin[:4].copy()
    kin2["x"][0] = 5
    kin2["y"][0] = 5
    _, valid2 = ctx.orb_compute(img, kin2)
    assert not valid2[0] and valid2[1:].all()


def test_orb_featureless_and_tiny(ctx_cache):
    ctx = ctx_cache(320, 240, 300)
    kps, desc = ctx.orb_detect_and_compute(np.full((240, 320), 127, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    ctx2 = ctx_cache(96, 80, 100)
    img = synth.synth_frame(80, 96, 5)
    kps, desc = ctx2.orb_detect_and_compute(img)
    okp, odesc = oo.orb_detect_and_compute(img, 100)
    assert np.array_equal(kps, okp) and np.array_equal(desc, odesc)


def test_feature_processor_mirror():
    from ros2_mono_vo_b200 import FeatureProcessor
    fp = FeatureProcessor(300, max_width=320, max_height=240)
    f0, f1 = synth.synth_pair(240, 320, 11)
    k0, d0 = fp.detect_and_compute(f0)
    k1, d1 = fp.detect_and_compute(f1)
    assert np.array_equal(fp.detect(f0), k0)
    m = fp.find_matches(d0, d1, 0.7)
    assert len(m) > 50
    from oracle import knn_oracle as ko
    qi, ti, d = ko.find_matches(d0, d1, 0.7)
    assert np.array_equal(m["query_idx"], qi) and np.array_equal(m["train_idx"], ti) and np.array_equal(m["distance"], d)
