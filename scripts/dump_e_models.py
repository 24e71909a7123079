"""Dump the E / F / H models and counts of the first m hypotheses (C4 sweep API) for build-to-build comparisons."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from ros2_mono_vo_b200 import Context
tag = sys.argv[1]
ctx = Context(1241, 376, nfeatures=2000, max_points=8192)
out = {}
for seed in (5, 6):
    p1, p2, R, t, inl = synth.scene_correspondences(3000, seed, outlier_frac=0.3)
    for model in (0, 1, 2):
        idx, counts, models = ctx.score_hypotheses(model, p1, p2, 512, 1.0, K=synth.KITTI_K, want_models=True)
        out[f"s{seed}_m{model}_idx"] = idx; out[f"s{seed}_m{model}_counts"] = counts; out[f"s{seed}_m{model}_models"] = models
np.savez(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"models_{tag}.npz"), **out)
ctx.close()
print("dumped", tag)
