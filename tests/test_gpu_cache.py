"""GPU: the content-keyed device caches of the synchronous single-call path: descriptor blocks and LK
pyramids are recognised by the content of the host buffer; a hit gives the same results as an upload; a modified buffer
never aliases a stale device copy."""
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


def test_descriptor_cache_hits_and_invalidation():
    from ros2_mono_vo_b200 import Context
    f0, f1 = synth.synth_pair(240, 320, 11)
    c = Context(320, 240, nfeatures=300)
    k0, d0 = c.orb_detect_and_compute(f0)
    k1, d1 = c.orb_detect_and_compute(f1)
    s0 = c.cache_stats()
    # the reference gathers the rows into a fresh Mat before matching (src/frame.cpp:50-64): new buffers, same content
    m = c.knn_ratio(d0.copy(), d1.copy(), 0.7)
    s1 = c.cache_stats()
    assert s1["desc_hits"] - s0["desc_hits"] == 2 and s1["desc_misses"] == s0["desc_misses"]
    # same results as a context that always uploads
    ref = Context(320, 240, nfeatures=300)
    ref.debug_set("cache", 0)
    mr = ref.knn_ratio(d0, d1, 0.7)
    assert np.array_equal(m, mr)
    assert ref.cache_stats()["desc_hits"] == 0
    # a mutated host buffer is a different block: miss, and the result follows the new content
    d1m = d1.copy()
    d1m[5] ^= 0xFF
    m2 = c.knn_ratio(d0, d1m, 0.7)
    s2 = c.cache_stats()
    assert s2["desc_misses"] - s1["desc_misses"] == 1 and s2["desc_hits"] - s1["desc_hits"] == 1
    assert np.array_equal(m2, ref.knn_ratio(d0, d1m, 0.7)) and not np.array_equal(m2, m)
    # the original block is still resident (4 entries), and a subset (the keyframe keeps fewer observations) is its own block
    assert np.array_equal(c.knn_ratio(d0, d1, 0.7), mr)
    sub = np.ascontiguousarray(d1[::2])
    assert np.array_equal(c.knn_ratio(d0, sub, 0.7), ref.knn_ratio(d0, sub, 0.7))
    # more blocks than entries: least recently used ones are replaced, results stay right
    rng = np.random.default_rng(0)
    for i in range(6):
        q = rng.integers(0, 256, (200 + i, 32)).astype(np.uint8)
        assert np.array_equal(c.knn_ratio(q, d1, 0.7), ref.knn_ratio(q, d1, 0.7))
    assert np.array_equal(c.knn_ratio(d0, d1, 0.7), mr)
    c.close()
    ref.close()


@pytest.mark.parametrize("bgr", [False, True])
def test_lk_pyramid_cache(bgr):
    from ros2_mono_vo_b200 import Context
    h, w = 240, 320
    frames, K = synth.synth_sequence(h, w, 1, 5)
    if bgr:
        frames = [np.ascontiguousarray(np.stack([f, np.roll(f, 1, 1) // 2 + f // 2, 255 - f], 2)) for f in frames]
    pts = np.stack(np.meshgrid(np.arange(40, 280, 24.0), np.arange(40, 200, 20.0)), -1).reshape(-1, 2).astype(np.float32)
    c = Context(w, h, nfeatures=300)
    ref = Context(w, h, nfeatures=300)
    ref.debug_set("cache", 0)
    for t in range(1, 5):
        # the tracker passes copies (Frame clones its image): same content as the previous call's "next", other buffer
        a = c.lk_track(frames[t - 1].copy(), frames[t].copy(), pts)
        b = ref.lk_track(frames[t - 1], frames[t], pts)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        # the parity tap follows the slots: level 1 of the "prev" pyramid is the pyrDown of the prev image
        assert np.array_equal(c.lk_level(0, 1, 0), ref.lk_level(0, 1, 0)) and np.array_equal(c.lk_level(1, 1, 0), ref.lk_level(1, 1, 0))
    s = c.cache_stats()
    assert s["pyr_hits"] == 3 and s["pyr_misses"] == 5          # first call builds two, every later call only the new frame
    # a modified previous image must be rebuilt
    fm = frames[4].copy()
    fm[100:120, 100:120] = 255 - fm[100:120, 100:120]
    a = c.lk_track(fm, frames[3], pts)
    b = ref.lk_track(fm, frames[3], pts)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    s2 = c.cache_stats()
    assert s2["pyr_misses"] == s["pyr_misses"] + 1 and s2["pyr_hits"] == s["pyr_hits"] + 1   # fm rebuilt, frames[3] resident
    # group steps rewrite the pyramids: the cache must not survive them
    if not bgr:
        c.group_step(frames[0][None], K)
    a = c.lk_track(fm, frames[3], pts)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    c.close()
    ref.close()
