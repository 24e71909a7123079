"""A/B of the 5-point root finders (mvo_debug_set "e5_roots_impl" 1 = bracketing, 2 = Ehrlich-Aberth first): models, masks, time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context
K = synth.KITTI_K
ctx = Context(1241, 376, nfeatures=2000, max_points=5200)
tot = {1: 0.0, 2: 0.0}
for (n, seed, planar, outl) in [(2000, 1, False, 0.3), (2000, 2, True, 0.3), (5000, 3, False, 0.2), (300, 4, False, 0.5), (64, 5, False, 0.1), (1000, 6, False, 0.6)]:
    p1, p2, R, t, inl = synth.scene_correspondences(n, seed, outlier_frac=outl, planar=planar)
    out = {}
    for impl in (1, 2):
        ctx.debug_set("e5_roots_impl", impl)
        E, mask, cnt = ctx.find_essential(p1, p2, K, 0.99, 1.0)
        t0 = time.perf_counter()
        for _ in range(30):
            ctx.find_essential(p1, p2, K, 0.99, 1.0)
        out[impl] = (E, mask, cnt, (time.perf_counter() - t0) / 30 * 1e3)
        # hypothesis sweep: models of 512 samples
        idx, counts, models = ctx.score_hypotheses(2, p1, p2, 512, thr=1.0, K=K, want_models=True)
        out[impl] += (counts, models)
    a, b = out[1], out[2]
    same_counts = np.array_equal(a[4], b[4])
    nm = int((a[4] >= 0).sum())
    md = np.nanmax(np.abs(np.where(a[4][..., None] >= 0, a[5] - b[5], 0.0))) if same_counts else float("nan")
    print(f"n={n} planar={planar}: inliers {a[2]} / {b[2]}, mask equal {np.array_equal(a[1], b[1])}, E max diff {np.abs(np.asarray(a[0]) - np.asarray(b[0])).max():.2e}, "
          f"sweep: model counts per hypothesis equal {same_counts} ({nm} models), max |dE| {md:.2e}, inlier counts equal {np.array_equal(a[4], b[4])}, call ms {a[3]:.3f} -> {b[3]:.3f}")
ctx.close()
