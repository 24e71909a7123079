"""Wall-clock cost of the synchronous single-call entry points (what the reference's host code pays per call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context
h, w = 376, 1241
frames, K = synth.synth_sequence(h, w, 0, 4)
ctx = Context(w, h, nfeatures=1000)
def t(f, reps=30):
    for _ in range(3): f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return 1e3 * (time.perf_counter() - t0) / reps
k0, d0 = ctx.orb_detect_and_compute(frames[0]); k1, d1 = ctx.orb_detect_and_compute(frames[1])
pts = np.stack([k0["x"], k0["y"]], 1)
bgr = [np.ascontiguousarray(np.repeat(f[:, :, None], 3, axis=2)) for f in frames]
print("orb gray  %.3f ms" % t(lambda: ctx.orb_detect_and_compute(frames[1])))
print("orb bgr   %.3f ms" % t(lambda: ctx.orb_detect_and_compute(bgr[1])))
print("knn       %.3f ms" % t(lambda: ctx.knn_ratio(d0, d1, 0.7)))
for n in (50, 1000):
    print("lk gray n=%d  %.3f ms" % (n, t(lambda: ctx.lk_track(frames[0], frames[1], pts[:n]))))
    print("lk bgr  n=%d  %.3f ms" % (n, t(lambda: ctx.lk_track(bgr[0], bgr[1], pts[:n]))))
nxt, st, err = ctx.lk_track(frames[0], frames[1], pts)
p1, p2 = pts[st == 1], nxt[st == 1]
print("H   %.3f ms" % t(lambda: ctx.find_homography(p1, p2, 1.0)))
print("F   %.3f ms" % t(lambda: ctx.find_fundamental(p1, p2, 1.0, 0.99)))
print("E   %.3f ms" % t(lambda: ctx.find_essential(p1, p2, K, 0.99, 1.0)))
obj, img, Kp, _, _ = synth.pnp_scene(300, 3, 0.5, 0.2)
print("pnp %.3f ms" % t(lambda: ctx.solve_pnp_ransac(obj, img, Kp)))
ctx.close()
