import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context
h, w, n = 376, 1241, 2000
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
f0, f1 = synth.synth_pair(h, w, 2)
ctx = Context(w, h, nfeatures=n)
for _ in range(3):
    k0, d0 = ctx.orb_detect_and_compute(f0)
    k1, d1 = ctx.orb_detect_and_compute(f1)
    m = ctx.knn_ratio(d0, d1, 0.7)
t = time.perf_counter()
for _ in range(reps):
    k1, d1 = ctx.orb_detect_and_compute(f1)
t1 = time.perf_counter()
for _ in range(reps):
    m = ctx.knn_ratio(d0, d1, 0.7)
t2 = time.perf_counter()
print(f"ORB e2e {1e3*(t1-t)/reps:.3f} ms/frame, kNN e2e {1e3*(t2-t1)/reps:.3f} ms, kps {len(k1)} matches {len(m)}")
try:
    import cv2
    orb = cv2.ORB_create(n); bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    for _ in range(2): ck, cd = orb.detectAndCompute(f1, None)
    t = time.perf_counter()
    for _ in range(5): ck, cd = orb.detectAndCompute(f1, None)
    t1 = time.perf_counter()
    for _ in range(5): mm = bf.knnMatch(d0, d1, 2)
    t2 = time.perf_counter()
    print(f"cv2 ORB {1e3*(t1-t)/5:.2f} ms, knn {1e3*(t2-t1)/5:.2f} ms, threads {cv2.getNumThreads()} cpus {os.cpu_count()}")
except Exception as e:
    print("cv2 unavailable", e)
