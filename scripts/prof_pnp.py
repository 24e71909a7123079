"""Small driver for ncu: mvo_solve_pnp_ransac single calls (n landmarks, default 2000 as in a tracking frame)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from ros2_mono_vo_b200 import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
obj, img, K, _, _ = synth.pnp_scene(n, 3, 0.5, 0.2)
ctx = Context(1241, 376, nfeatures=1000, max_points=max(n, 2000))
import time
for _ in range(4):
    t0 = time.perf_counter(); ctx.solve_pnp_ransac(obj, img, K); t1 = time.perf_counter()
print("call ms", (t1 - t0) * 1e3)
ctx.close()
