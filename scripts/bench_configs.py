"""Timings of the BASELINE.json configs that are parity cases rather than the bench line: C1 (640x480 two-view init),
C3 (1920x1080, 5000 ORB, keyframe matching) through the synchronous single-call API next to cv2 on the host, and the C4
hypothesis sweep (512 - 16384 hypotheses x 5000 correspondences)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context

def t(f, reps=10, warm=2):
    for _ in range(warm): f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return 1e3 * (time.perf_counter() - t0) / reps

out = {}
try:
    import cv2
    cv2.setNumThreads(os.cpu_count())
except ImportError:
    cv2 = None
for tag, (h, w, n) in {"C1": (480, 640, 1000), "C3": (1080, 1920, 5000)}.items():
    f0, f1 = synth.synth_pair(h, w, 1)
    ctx = Context(w, h, nfeatures=n, max_points=2 * n)
    k0, d0 = ctx.orb_detect_and_compute(f0); k1, d1 = ctx.orb_detect_and_compute(f1)
    m = ctx.knn_ratio(d0, d1, 0.7)
    p1 = np.stack([k0["x"][m["query_idx"]], k0["y"][m["query_idx"]]], 1); p2 = np.stack([k1["x"][m["train_idx"]], k1["y"][m["train_idx"]]], 1)
    r = {"keypoints": int(len(k1)), "matches": int(len(m)),
         "gpu_ms": {"orb": t(lambda: ctx.orb_detect_and_compute(f1)), "knn_ratio": t(lambda: ctx.knn_ratio(d0, d1, 0.7)),
                    "find_homography": t(lambda: ctx.find_homography(p1, p2, 1.0)), "find_fundamental": t(lambda: ctx.find_fundamental(p1, p2, 1.0, 0.99))}}
    ctx.close()
    if cv2 is not None:
        orb = cv2.ORB_create(n); bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        ck0, cd0 = orb.detectAndCompute(f0, None); ck1, cd1 = orb.detectAndCompute(f1, None)
        r["cv2_ms"] = {"orb": t(lambda: orb.detectAndCompute(f1, None), 5, 1), "knn_ratio": t(lambda: bf.knnMatch(cd0, cd1, 2), 5, 1),
                       "find_homography": t(lambda: cv2.findHomography(p1, p2, cv2.RANSAC, 1.0), 5, 1),
                       "find_fundamental": t(lambda: cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99), 5, 1)}
        r["cv2_threads"] = cv2.getNumThreads()
    out[tag] = r
# C4
p1, p2, R, tt, inl = synth.scene_correspondences(5000, 4, outlier_frac=0.3, noise_px=0.3)
ctx = Context(1241, 376, nfeatures=2000, max_points=8192)
c4 = {}
for model, name in ((0, "H"), (1, "F"), (2, "E")):
    for mhyp in (512, 4096, 16384):
        ms = t(lambda: ctx.score_hypotheses(model, p1, p2, mhyp, 1.0, K=synth.KITTI_K), 5, 1)
        c4[f"{name}_{mhyp}"] = {"ms": ms, "hypotheses_per_s": mhyp / (ms * 1e-3), "point_scores_per_s": mhyp * 5000 / (ms * 1e-3)}
ctx.close()
out["C4"] = c4
print(json.dumps(out, indent=1))
