"""solvePnPRansac: the library against cv2 on many seeded scenes (landmark counts, outlier rates, noise): inlier lists
and poses.  The winning hypothesis comes out of a 5-point EPnP whose null-space basis is decided by rounding noise, so
identity with cv2 is expected for most scenes, not all; this prints how many, for both forms of the
12 x 12 eigen-solver (pnp_epnp_impl 1 = shipped, 0 = round 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import synth
from ros2_mono_vo_b200 import Context
cv2.setNumThreads(1)
ctx = Context(1241, 376, nfeatures=1000, max_points=2000)
cases = [(n, seed, noise, outl) for n in (30, 100, 300, 1000, 2000) for seed in range(20) for noise, outl in ((0.3, 0.05), (0.5, 0.2), (1.0, 0.4))]
for impl in (1, 0):
    ctx.debug_set("pnp_epnp_impl", impl)
    same = 0; tot = 0; worst_r = 0.0; worst_t = 0.0; diff_sizes = []; fail = 0
    for n, seed, noise, outl in cases:
        obj, img, K, _, _ = synth.pnp_scene(n, 1000 + seed, noise, outl)
        cv2.setRNGSeed(0)
        ok_c, r_c, t_c, inl_c = cv2.solvePnPRansac(obj.astype(np.float32), img.astype(np.float32), K, None, iterationsCount=100, reprojectionError=8.0, confidence=0.99)
        ok, r, t, inl = ctx.solve_pnp_ransac(obj, img, K)
        tot += 1
        if bool(ok_c) != bool(ok):
            fail += 1
            continue
        if not ok:
            same += 1
            continue
        a = set(inl.tolist()); b = set(inl_c.ravel().tolist())
        if a == b:
            same += 1
            worst_r = max(worst_r, float(np.abs(r - r_c.ravel()).max())); worst_t = max(worst_t, float(np.abs(t - t_c.ravel()).max()))
        else:
            diff_sizes.append((n, seed, outl, len(a ^ b), len(b)))
    print(f"impl {impl}: {same} / {tot} scenes with cv2's inlier list (ok flag differs: {fail}); on those max |drvec| {worst_r:.2e} |dtvec| {worst_t:.2e}")
    print("   different lists (n, seed, outlier rate, |symmetric difference|, |cv2 list|):", diff_sizes[:12])
ctx.close()
