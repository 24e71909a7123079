"""A/B of the two homography-refinement kernels (mvo_debug_set "h_refine_impl" 1 | 2): results and time per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context

ctx = Context(1241, 376, nfeatures=2000, max_points=5200)
for (n, seed, planar, outl) in [(2000, 1, True, 0.3), (2000, 2, False, 0.3), (5000, 3, True, 0.2), (300, 4, True, 0.5), (40, 5, True, 0.1)]:
    p1, p2, R, t, inl = synth.scene_correspondences(n, seed, outlier_frac=outl, planar=planar)
    out = {}
    for impl in (1, 2):
        ctx.debug_set("h_refine_impl", impl)
        H, mask, cnt = ctx.find_homography(p1, p2, 1.0)
        t0 = time.perf_counter()
        for _ in range(50):
            ctx.find_homography(p1, p2, 1.0)
        out[impl] = (H, mask, cnt, (time.perf_counter() - t0) / 50 * 1e3)
    H1, m1, c1, t1 = out[1]
    H2, m2, c2, t2 = out[2]
    print(f"n={n} planar={planar}: inliers {c1} / {c2}, mask equal {np.array_equal(m1, m2)}, H bit-identical {np.array_equal(H1, H2)}, "
          f"max |dH| {np.abs(np.asarray(H1) - np.asarray(H2)).max():.3e}, call ms {t1:.3f} -> {t2:.3f}")
ctx.close()
