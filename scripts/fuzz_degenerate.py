"""Degenerate / edge inputs through the synchronous API: nothing may hang or crash; results are printed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context
from ros2_mono_vo_b200.api import MvoError
ctx = Context(640, 480, nfeatures=500)
K = synth.KITTI_K
def run(name, f):
    t0 = time.perf_counter()
    try:
        r = f()
        msg = "ok " + str(r)[:90].replace("\n", " ")
    except MvoError as e:
        msg = "MvoError " + str(e)[:90]
    print(f"{name:34s} {1e3 * (time.perf_counter() - t0):7.2f} ms  {msg}")
rng = np.random.default_rng(0)
same = np.full((200, 2), 100.0, np.float32)
line = np.stack([np.linspace(0, 600, 200), np.linspace(0, 300, 200)], 1).astype(np.float32)
rand1, rand2 = rng.uniform(0, 600, (200, 2)).astype(np.float32), rng.uniform(0, 400, (200, 2)).astype(np.float32)
for nm, a, b in (("identical points", same, same), ("collinear points", line, line + 1), ("pure noise", rand1, rand2)):
    run(f"H  {nm}", lambda: ctx.find_homography(a, b, 1.0)[2])
    run(f"F  {nm}", lambda: ctx.find_fundamental(a, b, 1.0, 0.99)[2])
    run(f"E  {nm}", lambda: ctx.find_essential(a, b, K, 0.99, 1.0)[2])
for n in (4, 5, 7, 8, 14, 15, 16):
    p1, p2, *_ = synth.scene_correspondences(n, 3, outlier_frac=0.0)
    run(f"H  n={n}", lambda: ctx.find_homography(p1, p2, 1.0)[2])
    run(f"F  n={n}", lambda: ctx.find_fundamental(p1, p2, 1.0, 0.99)[2])
    run(f"E  n={n}", lambda: ctx.find_essential(p1, p2, K, 0.99, 1.0)[2])
nanp = rand1.copy(); nanp[3, 0] = np.nan
run("H  NaN coordinate", lambda: ctx.find_homography(nanp, rand2, 1.0)[2])
run("E  NaN coordinate", lambda: ctx.find_essential(nanp, rand2, K, 0.99, 1.0)[2])
flat = np.full((480, 640), 77, np.uint8)
run("ORB flat image", lambda: len(ctx.orb_detect_and_compute(flat)[0]))
run("ORB 64x64 image", lambda: len(ctx.orb_detect_and_compute(synth.synth_frame(64, 64, 1))[0]))
run("ORB 40x40 image", lambda: len(ctx.orb_detect_and_compute(synth.synth_frame(40, 40, 1))[0]))
d = rng.integers(0, 256, (10, 32), dtype=np.uint8)
run("kNN empty query", lambda: len(ctx.knn_ratio(d[:0], d, 0.7)))
run("kNN one train row", lambda: len(ctx.knn_ratio(d, d[:1], 0.7)))
run("kNN empty train", lambda: len(ctx.knn_ratio(d, d[:0], 0.7)))
f0, f1 = synth.synth_pair(480, 640, 2)
run("LK no points", lambda: ctx.lk_track(f0, f1, np.zeros((0, 2), np.float32))[1].sum())
run("LK points outside", lambda: ctx.lk_track(f0, f1, np.array([[-50, -50], [700, 500], [639.9, 479.9], [0, 0]], np.float32))[1])
run("LK NaN point", lambda: ctx.lk_track(f0, f1, np.array([[np.nan, 5], [100, 100]], np.float32))[1])
obj, img, Kp, _, _ = synth.pnp_scene(50, 3, 0.5, 0.2)
run("PnP n=6", lambda: ctx.solve_pnp_ransac(obj[:6], img[:6], Kp)[0])
run("PnP identical 3-D points", lambda: ctx.solve_pnp_ransac(np.tile(obj[:1], (50, 1)), img, Kp)[0])
run("PnP NaN", lambda: ctx.solve_pnp_ransac(np.where(np.arange(150).reshape(50, 3) == 4, np.nan, obj).astype(np.float32), img, Kp)[0])
ctx.close()
print("done")
