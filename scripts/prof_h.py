"""Small driver for ncu: mvo_find_homography single calls on a 2000-point planar scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from ros2_mono_vo_b200 import Context
ctx = Context(1241, 376, nfeatures=2000, max_points=5200)
p1, p2, R, t, inl = synth.scene_correspondences(2000, 1, outlier_frac=0.3, planar=True)
for _ in range(4):
    H, mask, cnt = ctx.find_homography(p1, p2, 1.0)
print(cnt)
ctx.close()
