import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from ros2_mono_vo_b200 import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
obj, img, K, _, _ = synth.pnp_scene(n, 3, 0.5, 0.2)
ctx = Context(1241, 376, nfeatures=1000)
for _ in range(4):
    ctx.solve_pnp_ransac(obj, img, K)
ctx.close()
