"""A/B of the two forms of pnp_epnp_kernel's eigen-decomposition (mvo_debug_set("pnp_epnp_impl", 0 | 1)): call time of
mvo_solve_pnp_ransac, the batched tracking step, and the results against form 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context
ctx = Context(1241, 376, nfeatures=1000, max_points=2000)
ref = {}
for impl in (0, 1):
    ctx.debug_set("pnp_epnp_impl", impl)
    line = []
    for n, seed, outl in ((300, 3, 0.2), (2000, 5, 0.1), (2000, 7, 0.4), (50, 9, 0.3)):
        obj, img, K, _, _ = synth.pnp_scene(n, seed, 0.5, outl)
        for _ in range(3):
            r = ctx.solve_pnp_ransac(obj, img, K)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter(); r = ctx.solve_pnp_ransac(obj, img, K); ts.append(time.perf_counter() - t0)
        key = (n, seed)
        if impl == 0:
            ref[key] = r
            same = "ref"
        else:
            a, b = ref[key], r
            same = "inliers %s  |drvec| %.2e |dtvec| %.2e" % (np.array_equal(a[3], b[3]) if len(a) > 3 else "?", np.abs(np.asarray(a[1]) - np.asarray(b[1])).max(), np.abs(np.asarray(a[2]) - np.asarray(b[2])).max())
        line.append("n=%d: %.3f ms (%s)" % (n, np.median(ts) * 1e3, same))
    print("impl", impl, " | ".join(line))
ctx.close()
